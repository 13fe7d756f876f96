"""Per-tensor norms of all 44 gradients of one 512x512x64 training view (render forward + loss + backward): run once per setting
(e.g. SHERF_BWD_CHUNK_CAP=131072 and the default) and compare the lines -- chunking must not change the result beyond summation order."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sherf_b200 import synthetic as S                      # noqa: E402
from sherf_b200.triplane import hot_path_modules           # noqa: E402

dev = torch.device('cuda:0')
model = S.make_smpl_model(0)
scene = S.make_scene(S.SceneSpec(H=512, W=512, samples=64, seed=0), model)


def mv(x):
    if torch.is_tensor(x):
        return x.to(dev)
    if isinstance(x, dict):
        return {k: mv(v) for k, v in x.items()}
    if isinstance(x, list):
        return [mv(v) for v in x]
    return x


scene = {k: mv(v) for k, v in scene.items()}
ren, dec = hot_path_modules(model, seed=0, dense_sigma=True)
ren, dec = ren.to(dev).requires_grad_(True), dec.to(dev).requires_grad_(True)
leaves = {'planes': scene['planes'].requires_grad_(True), 'obs_input_feature': scene['obs_input_feature'].requires_grad_(True)}
for l, v in enumerate(scene['volumes']):
    leaves[f'vol{l}'] = v.requires_grad_(True)
torch.manual_seed(1)
tgt = torch.rand(1, 512 * 512, 3, device=dev)
rgb, depth, acc = ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None, scene['obs_sp_input'], dec,
                      scene['ray_origins'], scene['ray_directions'], scene['near'], scene['far'], scene['input_data'], scene['rendering_options'])
(100 * ((rgb / 2 + 0.5 - tgt) ** 2).mean() + 10 * ((acc - 1) ** 2).mean()).backward()
out = {}
for k, p in list(ren.named_parameters()) + list(dec.named_parameters()) + list(leaves.items()):
    if p.grad is not None:
        out[k] = [float(p.grad.double().norm()), float(p.grad.double().sum())]
print(json.dumps(out))
