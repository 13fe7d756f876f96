"""Times one training-style step of the hot path on the GPU box: forward + loss + backward through sherf_b200.ImportanceRenderer at
BASELINE configs[1] (512x512 rays x 64 samples), all 39 hot-path parameters and the five feature tensors requiring grad.
Prints one JSON line (CUDA events, max of nothing: single GPU)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sherf_b200 import synthetic as S                      # noqa: E402
from sherf_b200.triplane import hot_path_modules           # noqa: E402


def main():
    H = W = int(os.environ.get('BWD_RES', '512'))
    samples = int(os.environ.get('BWD_SAMPLES', '64'))
    steps, warmup = int(os.environ.get('BWD_STEPS', '5')), int(os.environ.get('BWD_WARMUP', '2'))
    dev = torch.device('cuda:0')
    model = S.make_smpl_model(0)
    scene = S.make_scene(S.SceneSpec(H=H, W=W, samples=samples, seed=0), model)

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
    ren, dec = ren.to(dev), dec.to(dev)
    ren.requires_grad_(True); dec.requires_grad_(True)
    inputs_grad = os.environ.get('BWD_INPUT_GRADS', '1') != '0'
    if inputs_grad:
        scene['planes'].requires_grad_(True); scene['obs_input_feature'].requires_grad_(True)
        for v in scene['volumes']:
            v.requires_grad_(True)
    N = H * W
    tgt = torch.rand(1, N, 3, device=dev)

    def step():
        rgb, depth, acc = ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None, scene['obs_sp_input'],
                              dec, scene['ray_origins'], scene['ray_directions'], scene['near'], scene['far'], scene['input_data'],
                              scene['rendering_options'])
        loss = 100 * ((rgb / 2 + 0.5 - tgt) ** 2).mean() + 10 * ((acc - 1) ** 2).mean()
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({'what': 'forward + loss + backward per view', 'H': H, 'W': W, 'samples': samples, 'surviving_points': ren.last_num_points,
                      'ms_per_step': ms, 'ray_samples_per_sec_training': N * samples / ms * 1e3, 'input_grads': inputs_grad,
                      'backward_launches': getattr(ren, 'last_backward_launches', None), 'arithmetic': 'fp32 simt' if os.environ.get('SHERF_BWD_SIMT') == '1' else '3xTF32 tcgen05', 'loss': float(loss),
                      'peak_mem_GB': torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == '__main__':
    main()
