"""Times SparseConvNet.forward (csrc/sparse_encoder.cu) at the real size: the 6 890 canonical SMPL vertices voxelised at 5 mm like
prepare_sp_input (triplane.py:174-217), three dense output levels.  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sherf_b200 import synthetic as S                              # noqa: E402
from sherf_b200.renderer import SparseConvNet, SparseConvTensor    # noqa: E402

dev = torch.device('cuda:0')
model = S.make_smpl_model(0)
scene = S.make_scene(S.SceneSpec(H=8, W=8, samples=4, seed=0), model)
tv = scene['input_data']['t_vertices'][0]
bounds, out_sh = scene['obs_sp_input']['bounds'][0], scene['obs_sp_input']['out_sh']
coord = torch.round((tv[:, [2, 1, 0]] - bounds[0][[2, 1, 0]]) / 0.005).to(torch.int32)
idx = torch.cat([torch.zeros(coord.shape[0], 1, dtype=torch.int32), coord], 1).to(dev)
feat = torch.randn(coord.shape[0], 32, device=dev)
torch.manual_seed(0)
enc = SparseConvNet(4).to(dev).eval()
sp = SparseConvTensor(feat, idx, out_sh, 1)
torch.set_grad_enabled(False)                                      # the evaluation forward (under autograd the differentiable kernels run instead)
for _ in range(3):
    vols = enc(sp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    vols = enc(sp)
e1.record()
e1.synchronize()
print(json.dumps({'sparse_encoder_ms': e0.elapsed_time(e1) / 10, 'vertices': int(coord.shape[0]), 'out_sh': out_sh,
                  'active_sites': [int((v[0] != 0).any(0).sum()) for v in vols],
                  'dense_output_MB': sum(v.numel() * 4 for v in vols) / 1e6}))
