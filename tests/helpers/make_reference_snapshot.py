"""Run in a FRESH interpreter (tests/test_overlay.py): writes a `network-snapshot`-style pickle of the REFERENCE's own generator
(training_loop.py:572-579 shape: dict(G=..., G_ema=..., training_set_kwargs=None, augment_pipe=None)) plus its tensors as a plain
state dict.  The spconv stand-ins are given spconv's real module paths so that the pickle refers to `spconv.pytorch.conv.SubMConv3d` /
`spconv.pytorch.modules.SparseSequential` exactly like a snapshot written with spconv installed.  usage: <out.pkl> <out_state.pt>"""
import copy
import os
import pickle
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torchvision.models as tvm  # noqa: E402

_orig_resnet18 = tvm.resnet18
tvm.resnet18 = lambda *a, pretrained=False, **k: _orig_resnet18(weights=None)

from oracle import ref_shim  # noqa: E402
from sherf_b200 import synthetic as S  # noqa: E402

model_t = S.smpl_model_to_torch(S.make_smpl_model(0))
ref_shim.load(model_t)
conv_mod, mods_mod = types.ModuleType('spconv.pytorch.conv'), types.ModuleType('spconv.pytorch.modules')
SubMConv3d = type('SubMConv3d', (ref_shim.SpconvConvStub,), {'__module__': 'spconv.pytorch.conv'})
SparseConv3d = type('SparseConv3d', (ref_shim.SpconvConvStub,), {'__module__': 'spconv.pytorch.conv'})
SparseSequential = type('SparseSequential', (ref_shim.SpconvSequentialStub,), {'__module__': 'spconv.pytorch.modules'})
conv_mod.SubMConv3d, conv_mod.SparseConv3d, mods_mod.SparseSequential = SubMConv3d, SparseConv3d, SparseSequential
sys.modules.update({'spconv.pytorch.conv': conv_mod, 'spconv.pytorch.modules': mods_mod})
sp = sys.modules['spconv.pytorch']
sp.SubMConv3d, sp.SparseConv3d, sp.SparseSequential = SubMConv3d, SparseConv3d, SparseSequential
import dnnlib  # noqa: E402

rendering = {'image_resolution': 512, 'disparity_space_sampling': False, 'clamp_mode': 'relu',
             'superresolution_module': 'training.superresolution.SuperresolutionHybrid8XDC', 'white_back': False,
             'c_gen_conditioning_zero': True, 'gpc_reg_prob': None, 'c_scale': 1.0, 'superresolution_noise_mode': 'none', 'density_reg': 0.25,
             'density_reg_p_dist': 0.004, 'density_noise': 1.0, 'reg_type': 'l1', 'decoder_lr_mul': 1.0, 'sr_antialias': True,
             'depth_resolution': 48, 'depth_resolution_importance': 0}
G_kwargs = dnnlib.EasyDict(class_name='training.triplane.TriPlaneGenerator', z_dim=512, w_dim=512, use_1d_feature=True, use_2d_feature=True,
                           use_3d_feature=True, use_trans=True, use_NeRF_decoder=True, mapping_kwargs=dnnlib.EasyDict(num_layers=2),
                           channel_base=32768, channel_max=512, fused_modconv_default='inference_only', rendering_kwargs=rendering,
                           num_fp16_res=0, sr_num_fp16_res=4, conv_clamp=None,
                           sr_kwargs=dnnlib.EasyDict(channel_base=32768, channel_max=512, fused_modconv_default='inference_only'))
torch.manual_seed(3)
G = dnnlib.util.construct_class_by_name(**G_kwargs, c_dim=0, img_resolution=512, img_channels=3)
with torch.no_grad():                                       # spconv weights are zero-filled in the stand-ins: make them distinguishable
    for n, p in G.named_parameters():
        if 'encoder_3d' in n and p.dim() == 5:
            p.copy_(torch.randn(p.shape, generator=torch.Generator().manual_seed(len(n))) * 0.05)
G_ema = copy.deepcopy(G).eval().requires_grad_(False)
with open(sys.argv[1], 'wb') as f:
    pickle.dump(dict(G=G, G_ema=G_ema, training_set_kwargs=None, augment_pipe=None), f)
torch.save({k: v.detach().clone() for k, v in G_ema.state_dict().items()}, sys.argv[2])
print('SNAPSHOT_OK', len(G_ema.state_dict()))
