"""Run in a FRESH interpreter (tests/test_overlay.py): builds the reference's own TriPlaneGenerator (under the oracle's shims) and
this repo's overlay class through the reference's own `dnnlib.util.construct_class_by_name('training.triplane.TriPlaneGenerator')`,
and checks the resume contract of training_loop.py:193-208 between them.  Needs /root/reference.  Prints one JSON line."""
import copy
import io
import json
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torchvision.models as tvm  # noqa: E402

_orig_resnet18 = tvm.resnet18
tvm.resnet18 = lambda *a, pretrained=False, **k: _orig_resnet18(weights=None)         # no network: skip the ImageNet download (triplane.py:323)

from oracle import ref_shim  # noqa: E402
from sherf_b200 import overlay, synthetic as S  # noqa: E402

model_t = S.smpl_model_to_torch(S.make_smpl_model(0))
ref_shim.load(model_t)                                                                # imports the REFERENCE training.* under shims
import dnnlib  # noqa: E402
from torch_utils import misc  # noqa: E402

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
common = dict(c_dim=0, img_resolution=512, img_channels=3)                            # training_loop.py:192

torch.manual_seed(0)
ref_G = dnnlib.util.construct_class_by_name(**G_kwargs, **common)                     # the reference's class (reference module in sys.modules)
ref_file = sys.modules['training.triplane'].__file__

saved = {n: sys.modules.get(n) for n in overlay.SHADOWED}
overlay.install()
torch.manual_seed(1)
G = dnnlib.util.construct_class_by_name(**G_kwargs, **common)                         # training_loop.py:193, now resolving to the overlay
our_file = sys.modules['training.triplane'].__file__

ref_named = {n: tuple(t.shape) for n, t in misc.named_params_and_buffers(ref_G)}
our_named = {n: tuple(t.shape) for n, t in misc.named_params_and_buffers(G)}
out = {
    'ref_file': ref_file, 'our_file': our_file,
    'only_ref': sorted(set(ref_named) - set(our_named)), 'only_ours': sorted(set(our_named) - set(ref_named)),
    'shape_mismatch': sorted(n for n in set(ref_named) & set(our_named) if ref_named[n] != our_named[n]),
    'n_tensors': len(our_named),
    'state_dict_equal': set(ref_G.state_dict()) == set(G.state_dict()),
}
with torch.no_grad():
    misc.copy_params_and_buffers(ref_G, G, require_all=True)                          # resume direction (training_loop.py:207-208)
    out['copied_equal'] = all(torch.equal(t, dict(misc.named_params_and_buffers(ref_G))[n]) for n, t in misc.named_params_and_buffers(G))
    misc.copy_params_and_buffers(G, ref_G, require_all=True)                          # and back: nothing of ours is missing upstream
# the per-tick deepcopy / pickle of training_loop.py:196,572-579 (persistence stores this module's source in the pickle)
G2 = copy.deepcopy(G).eval().requires_grad_(False)
buf = io.BytesIO()
pickle.dump(dict(G=G, G_ema=G2), buf)
buf.seek(0)
back = pickle.load(buf)
out['pickle_roundtrip_names_equal'] = set(back['G'].state_dict()) == set(G.state_dict())
out['pickle_roundtrip_values_equal'] = all(torch.equal(v, back['G_ema'].state_dict()[k]) for k, v in G2.state_dict().items())
out['pickle_bytes'] = buf.getbuffer().nbytes
out['init_kwargs_kept'] = back['G'].init_kwargs['rendering_kwargs']['depth_resolution'] == 48
out['hot_path_params'] = sum(t.numel() for n, t in G.named_parameters()
                             if n.startswith(('renderer.conv1d', 'renderer.transformer', 'decoder.')))
print('RESULT ' + json.dumps(out))
