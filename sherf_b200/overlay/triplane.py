"""Resolved as `training.triplane` by sherf_b200.overlay: the generator shell the reference constructs by dotted name
(/root/reference/sherf/train.py:310, training_loop.py:193) with the render hot path on sherf_b200's CUDA kernels.

`TriPlaneGenerator` keeps the reference's constructor signature, sub-module / parameter names, `mapping` / `synthesis` / `forward`
signatures and the returned dict (triplane.py:29-172,232-236).  What it owns itself:
  * the observation preparation of triplane.py:105-137 -> `ImportanceRenderer.prepare_observation` (csrc/observation.cu),
  * the render call of :156-157 -> `ImportanceRenderer.forward` (libsherf_b200.so),
  * the output reshape of :160-172.
The StyleGAN2 backbone, the two ResNet-18 encoders and the super-resolution module are reference code outside this repo's scope
(SURVEY.md section 2): they are imported from the reference tree when it is importable, else built by the callables registered with
`sherf_b200.overlay.set_factories`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from sherf_b200 import overlay as _overlay
from sherf_b200.triplane import NeRFDecoder                                            # noqa: F401  (re-exported name)
from training.volumetric_rendering.renderer import ImportanceRenderer, SMPL_to_tensor, read_pickle    # noqa: F401  (triplane.py:19)

try:                                                                                   # reference infrastructure, when present
    from torch_utils import persistence as _persistence
    _persistent_class = _persistence.persistent_class
except ImportError:
    def _persistent_class(cls):
        return cls
try:
    import dnnlib as _dnnlib
except ImportError:
    _dnnlib = None
try:
    from training.networks_stylegan2 import FullyConnectedLayer, Generator as StyleGAN2Backbone
except ImportError:
    StyleGAN2Backbone = None

    class FullyConnectedLayer(nn.Module):
        """Parameter container with networks_stylegan2.FullyConnectedLayer's parameter names / shapes (weight [out,in], bias [out])."""

        def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
            super().__init__()
            self.weight = nn.Parameter(torch.randn(out_features, in_features) / lr_multiplier)
            self.bias = nn.Parameter(torch.full([out_features], float(bias_init))) if bias else None
try:
    from training.volumetric_rendering.ray_sampler import RaySampler
except ImportError:
    RaySampler = None


class ResNet18Classifier(nn.Module):
    """torchvision ResNet-18 wrapper with the reference's attribute name (`backbone`, triplane.py:320-343): the whole network
    gives the 512-d latent, conv1..layer1 without the max-pool gives the half-resolution 64-channel feature map."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        from torchvision.models import resnet18
        try:
            self.backbone = resnet18(weights='IMAGENET1K_V1')
        except Exception:                      # no network / no cached weights: the checkpoint overwrites them (copy_params_and_buffers)
            self.backbone = resnet18(weights=None)

    def forward(self, x, extract_feature=False):
        b = self.backbone
        x = b.relu(b.bn1(b.conv1(x)))
        if not extract_feature:
            x = b.maxpool(x)
        x = b.layer1(x)
        if extract_feature:
            return x
        x = b.layer4(b.layer3(b.layer2(x)))
        return torch.flatten(b.avgpool(x), 1)


class OSGDecoder(nn.Module):
    """Import-surface stand-in for triplane.py:242-265 (parameter names `net.0.*`, `net.2.*`).  Every shipped SHERF script sets
    `--use_nerf_decoder True`; the CUDA path implements that configuration only, so evaluating this decoder raises."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = nn.Sequential(FullyConnectedLayer(n_features, self.hidden_dim, lr_multiplier=options['decoder_lr_mul']), nn.Softplus(),
                                 FullyConnectedLayer(self.hidden_dim, 1 + options['decoder_output_dim'], lr_multiplier=options['decoder_lr_mul']))

    def forward(self, sampled_features, ray_directions):
        raise NotImplementedError('sherf_b200 renders with the NeRF decoder only (use_NeRF_decoder=True, as every shipped SHERF script does)')


def _build(kind, reference_ctor, *args, **kwargs):
    factory = _overlay.FACTORIES.get(kind)
    if factory is not None:
        return factory(*args, **kwargs)
    if reference_ctor is None:
        raise RuntimeError(f'the reference tree is not importable and no `{kind}` factory is registered '
                           f'(sherf_b200.overlay.set_factories): the {kind} is reference code outside this repo\'s scope')
    return reference_ctor(*args, **kwargs)


@_persistent_class
class TriPlaneGenerator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, use_1d_feature, use_2d_feature, use_3d_feature, use_trans, use_NeRF_decoder, img_resolution,
                 img_channels, sr_num_fp16_res=0, mapping_kwargs={}, rendering_kwargs={}, sr_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.renderer = ImportanceRenderer(use_1d_feature=use_1d_feature, use_2d_feature=use_2d_feature, use_3d_feature=use_3d_feature,
                                           use_trans=use_trans, use_NeRF_decoder=use_NeRF_decoder)
        self.ray_sampler = RaySampler() if RaySampler is not None else None          # parameter-free; its call is commented out upstream (:91)
        self.encoder_2d = _build('encoder_2d', ResNet18Classifier)
        self.encoder_2d_feature = _build('encoder_2d', ResNet18Classifier)
        self.conv1d_projection = nn.Conv1d(96, 32, 1)
        self.backbone = _build('backbone', StyleGAN2Backbone, z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3,
                               mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        sr_args = dict(channels=32, img_resolution=img_resolution, sr_num_fp16_res=sr_num_fp16_res,
                       sr_antialias=rendering_kwargs.get('sr_antialias', True), **sr_kwargs)
        if _overlay.FACTORIES.get('superresolution') is not None:
            self.superresolution = _overlay.FACTORIES['superresolution'](**sr_args)
        elif _dnnlib is not None and rendering_kwargs.get('superresolution_module'):
            self.superresolution = _dnnlib.util.construct_class_by_name(class_name=rendering_kwargs['superresolution_module'], **sr_args)
        else:
            self.superresolution = None                                               # every shipped script runs with --use_sr_module False
        self.decoder = NeRFDecoder(32) if use_NeRF_decoder else OSGDecoder(
            32, {'decoder_lr_mul': rendering_kwargs.get('decoder_lr_mul', 1), 'decoder_output_dim': 3})
        self.neural_rendering_resolution = 64
        self.rendering_kwargs = rendering_kwargs
        self.use_1d_feature, self.use_2d_feature, self.use_3d_feature = use_1d_feature, use_2d_feature, use_3d_feature
        self._last_planes = None

    def mapping(self, z, c, input_img=None, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        z = self.encoder_2d(input_img)                                                # the latent IS the image encoding (triplane.py:75)
        if self.rendering_kwargs['c_gen_conditioning_zero']:
            c = torch.zeros_like(c)
        return self.backbone.mapping(z, c * self.rendering_kwargs.get('c_scale', 0), truncation_psi=truncation_psi,
                                     truncation_cutoff=truncation_cutoff, update_emas=update_emas)

    def synthesis(self, ws, input_data, c, neural_rendering_resolution=None, use_sr_module=True, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, test_flag=False, **synthesis_kwargs):
        if neural_rendering_resolution is None:
            neural_rendering_resolution = self.neural_rendering_resolution
        else:
            self.neural_rendering_resolution = neural_rendering_resolution
        ray_origins, ray_directions = input_data['ray_o_all'][:, 0], input_data['ray_d_all'][:, 0]        # rays come from the dataset (:92)
        near, far = input_data['near_all'][:, 0], input_data['far_all'][:, 0]
        N = ray_origins.shape[0]
        if use_cached_backbone and self._last_planes is not None:
            planes = self._last_planes
        else:
            planes = self.backbone.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)            # [1, 96, 256, 256]
        if cache_backbone:
            self._last_planes = planes
        if not self.use_3d_feature:
            raise NotImplementedError('sherf_b200 implements use_1d/2d/3d_feature all True (every shipped SHERF script)')
        obs_input_img = input_data['obs_img_all'][:, 0]
        obs_input_feature = self.encoder_2d_feature(obs_input_img, extract_feature=True)
        # triplane.py:111-137 on the device: vertex features, visibility, canonical vertices, voxel coordinates
        canonical_sp_conv_volume, obs_sp_input, obs_smpl_vertex_mask = self.renderer.prepare_observation(
            input_data, obs_input_img, obs_input_feature, self.conv1d_projection)
        planes = planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
        if test_flag:
            self.rendering_kwargs.update({'density_noise': 0})
        feature_samples, depth_samples, weights_samples = self.renderer(
            planes, obs_input_img, obs_input_feature, canonical_sp_conv_volume, obs_smpl_vertex_mask, obs_sp_input, self.decoder,
            ray_origins, ray_directions, near, far, input_data, self.rendering_kwargs)
        return self.images_from_samples(feature_samples, depth_samples, weights_samples, input_data['obs_img_all'].shape[-2:], N, ws,
                                        use_sr_module, synthesis_kwargs)

    def images_from_samples(self, feature_samples, depth_samples, weights_samples, hw, N, ws=None, use_sr_module=False, synthesis_kwargs=None):
        """triplane.py:160-172: [N, H*W, C] channels-last samples -> image tensors; super-resolution only on request."""
        H, W = hw
        feature_image = feature_samples.permute(0, 2, 1).reshape(N, feature_samples.shape[-1], H, W).contiguous()
        depth_image = depth_samples.permute(0, 2, 1).reshape(N, 1, H, W)
        weights_image = weights_samples.permute(0, 2, 1).reshape(N, 1, H, W)
        rgb_image = feature_image[:, :3]
        if use_sr_module:
            if self.superresolution is None:
                raise RuntimeError('use_sr_module=True but no super-resolution module was built (reference tree absent)')
            kw = synthesis_kwargs or {}
            sr_image = self.superresolution(rgb_image, feature_image, ws, noise_mode=self.rendering_kwargs['superresolution_noise_mode'],
                                            **{k: kw[k] for k in kw if k != 'noise_mode'})
        else:
            sr_image = rgb_image
        return {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image, 'weights_image': weights_image}

    def sample(self, coordinates, directions, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        raise NotImplementedError('TriPlaneGenerator.sample calls run_model with EG3D\'s 5-argument signature upstream (triplane.py:219-224 vs '
                                  'renderer.py:400) and cannot execute there either; shape extraction is not part of SHERF')

    sample_mixed = sample

    def forward(self, input_data, z, c, truncation_psi=1, truncation_cutoff=None, neural_rendering_resolution=None, use_sr_module=True,
                update_emas=False, cache_backbone=False, use_cached_backbone=False, test_flag=False, **synthesis_kwargs):
        input_img = input_data['obs_img_all'][:, 0]
        ws = self.mapping(z, c, input_img=input_img, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, input_data, c, update_emas=update_emas, neural_rendering_resolution=neural_rendering_resolution,
                              use_sr_module=use_sr_module, cache_backbone=cache_backbone, use_cached_backbone=use_cached_backbone,
                              test_flag=test_flag, **synthesis_kwargs)
