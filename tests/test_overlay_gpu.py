"""GPU test of the generator-level boundary: `training.triplane.TriPlaneGenerator` resolved through sherf_b200.overlay, built by dotted
name (dnnlib.util.construct_class_by_name where the reference tree exists, its restatement otherwise), against the golden produced by
the REFERENCE's own TriPlaneGenerator.synthesis (oracle/gen_golden_synthesis.py): observation preparation (triplane.py:105-137),
render call (:156-157), output reshape (:160-172); then the per-tick deepcopy / pickle of training_loop.py:196,572-579.

Tolerances: box / grid shape bit-exact; visibility mask and voxel coordinates >= 99.9 % equal (sign of a near-zero dot product / a
rounding tie on a float result); canonical vertices 5e-6 m; vertex features 1e-4; images: rgb / acc 1e-4 and depth 1e-3 x span on
>= 99.5 % of the pixels (a flipped voxel coordinate moves one 5 mm voxel of the 3-D feature volume)."""
import copy
import io
import os
import pickle
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN_DIR, scene_to
from sherf_b200 import overlay, synthetic as S

pytestmark = pytest.mark.gpu


class FixedPlanes(nn.Module):
    """Stand-in for the StyleGAN2 backbone (out of scope): returns the scene's tri-planes."""

    def __init__(self, planes):
        super().__init__()
        self.register_buffer('planes', planes.reshape(1, 96, 256, 256).clone())

    def mapping(self, z, c, **k):
        return None

    def synthesis(self, ws, update_emas=False, **k):
        return self.planes


class FixedFeature(nn.Module):
    """Stand-in for the ResNet-18 encoder (out of scope): returns the scene's half-resolution feature map."""

    def __init__(self, feat):
        super().__init__()
        self.register_buffer('feat', feat.clone())

    def forward(self, x, extract_feature=False):
        return self.feat if extract_feature else x.new_zeros(x.shape[0], 512)


@pytest.fixture()
def installed_overlay():
    overlay.install()
    yield
    overlay.set_factories()
    overlay.uninstall()


def test_synthesis_by_dotted_name_against_reference_golden(installed_overlay, smpl_model):
    from oracle.gen_golden_synthesis import encoder_state, state_checksum
    g = np.load(os.path.join(GOLDEN_DIR, 'synthesis_32x32x16.npz'))
    H, W, samples, seed, rgr, wb = [int(v) for v in g['spec']]
    spec = S.SceneSpec(H=H, W=W, samples=samples, seed=seed, random_global_R=bool(rgr), white_back=bool(wb))
    dev = torch.device('cuda:0')
    cpu_scene = S.make_scene(spec, smpl_model)
    scene = scene_to(cpu_scene, dev)
    overlay.set_factories(backbone=lambda *a, **k: FixedPlanes(cpu_scene['planes']), encoder_2d=lambda: FixedFeature(cpu_scene['obs_input_feature']))
    rendering = dict(cpu_scene['rendering_options'], c_gen_conditioning_zero=True, superresolution_noise_mode='none')
    kwargs = dict(class_name='training.triplane.TriPlaneGenerator', z_dim=512, c_dim=0, w_dim=512, use_1d_feature=True, use_2d_feature=True,
                  use_3d_feature=True, use_trans=True, use_NeRF_decoder=True, img_resolution=512, img_channels=3, rendering_kwargs=rendering)
    try:
        import dnnlib                                                     # the reference's own factory, where the tree exists
        construct = dnnlib.util.construct_class_by_name
    except ImportError:
        construct = overlay.construct_class_by_name
    cwd = os.getcwd()
    os.chdir('/tmp')                                                      # no assets/SMPL_NEUTRAL.pkl here: the synthetic body is injected
    try:
        G = construct(**kwargs)
    finally:
        os.chdir(cwd)
    assert type(G).__module__ in ('training.triplane',) or 'overlay' in sys.modules[type(G).__module__].__file__
    G.renderer.set_smpl_model(smpl_model)
    # ---- golden weights: hot path + projection conv stored, sparse encoder regenerated from its seed (checksum proves identity) ----
    w = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('w/')}
    enc_sd = encoder_state(int(g['enc_seed']))
    assert state_checksum(enc_sd) == str(g['enc_sha256'])
    sd = G.state_dict()
    for k, v in w.items():
        sd[k].copy_(v)
    for k, v in enc_sd.items():
        sd['renderer.encoder_3d.' + k].copy_(v)
    sd['conv1d_projection.weight'].copy_(torch.from_numpy(g['proj_w']))
    sd['conv1d_projection.bias'].copy_(torch.from_numpy(g['proj_b']))
    G = G.to(dev).eval().requires_grad_(False)
    # ---- observation preparation against the reference's own intermediates ----
    vol, sp_input, vmask, can = G.renderer.prepare_observation(scene['input_data'], scene['obs_input_img'], scene['obs_input_feature'],
                                                               G.conv1d_projection, return_canonical=True)
    V = can.shape[1]
    assert sp_input['out_sh'] == [int(v) for v in g['out_sh']]
    assert np.array_equal(sp_input['bounds'].cpu().numpy(), g['bounds']), 'canonical box'
    gm = np.unpackbits(g['vmask'])[:V].astype(bool)
    same_mask = float((vmask.cpu().numpy()[0] == gm).mean())
    e_can = float((can.cpu()[0] - torch.from_numpy(g['can'])).abs().max())
    same_coord = float((vol.indices.cpu().numpy() == g['coord']).all(1).mean())
    both = torch.from_numpy((vmask.cpu().numpy()[0] == gm))
    e_feat = float((vol.features.cpu() - torch.from_numpy(g['vert_feat']))[both].abs().max())
    print(f'\n[observation] mask equal {same_mask:.5f} canonical {e_can:.2e} coord equal {same_coord:.5f} vertex features {e_feat:.2e} '
          f'visible {int(vmask.sum())}/{V}')
    assert same_mask >= 0.999 and same_coord >= 0.999
    assert e_can <= 5e-6 and e_feat <= 1e-4
    # ---- the generator call (loss.py:82 / test_loop.py:189 shape) ----
    out = G.synthesis(None, scene['input_data'], None, use_sr_module=False, test_flag=True)
    assert set(out) == {'image', 'image_raw', 'image_depth', 'weights_image'}
    assert out['image'].shape == (1, 3, H, W) and out['image_depth'].shape == (1, 1, H, W) and out['weights_image'].shape == (1, 1, H, W)
    assert out['image'].data_ptr() == out['image_raw'].data_ptr()         # image IS image_raw without the SR module (triplane.py:170)
    span = float((cpu_scene['far'] - cpu_scene['near']).abs().max())
    d_rgb = (out['image'].cpu() - torch.from_numpy(g['image'])).abs().amax(1).reshape(-1)
    d_acc = (out['weights_image'].cpu() - torch.from_numpy(g['weights_image'])).abs().reshape(-1)
    d_dep = (out['image_depth'].cpu() - torch.from_numpy(g['image_depth'])).abs().reshape(-1) / span
    bad = (d_rgb > 1e-4) | (d_acc > 1e-4) | (d_dep > 1e-3)
    print(f'[synthesis] rgb={float(d_rgb.max()):.2e} acc={float(d_acc.max()):.2e} depth/span={float(d_dep.max()):.2e} bad pixels {float(bad.float().mean()):.4%} '
          f'acc.max={float(out["weights_image"].max()):.3f}')
    assert float(bad.float().mean()) <= 5e-3
    # ---- training_loop.py:196 (deepcopy) and :572-579 (pickle) after a forward: identical renders ----
    G2 = copy.deepcopy(G)
    buf = io.BytesIO()
    pickle.dump(dict(G_ema=G), buf)
    buf.seek(0)
    G3 = pickle.load(buf)['G_ema']
    for other in (G2, G3):
        o = other.synthesis(None, scene['input_data'], None, use_sr_module=False, test_flag=True)
        assert all(torch.equal(o[k], out[k]) for k in out)
