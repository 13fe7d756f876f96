"""Novel-view sampler in SHERF's calling convention: what `test_loop.py:197-251` does around the renderer --
render target views of one observed subject and write `frame{pose:04d}_view{view:04d}.png` (`to8b(image / 2 + 0.5)`, test_loop.py:197,
218-222) plus the depth / accumulated-weight maps as `.npy` -- driven entirely through the public API (device ray generation,
`ImportanceRenderer.forward`, `dist.render_sequence`).  The reference's own `gen_samples.py` / `gen_videos.py` call `G.synthesis(ws,
camera_params)` with EG3D's signature and cannot run against SHERF's generator; this replaces them for the
render half.  No dataset or checkpoint is available offline, so the CLI renders the seeded synthetic subject; `render_orbit` takes
any scene dict with the layout of `sherf_b200.synthetic.make_scene`.

    python -m sherf_b200.sample --out out_dir --views 8 --res 512 --samples 64 [--importance 64] [--weights hot_path_state_dict.pt]
    python -m sherf_b200.sample --out out_dir --network network-snapshot-000400.pkl --reference-root /path/to/SHERF/sherf

`--network` loads a reference snapshot through `sherf_b200.checkpoint.load_generator` (legacy.load_network_pkl + the overlay) and renders
with ITS renderer / decoder parameters; the observation tensors (planes, feature map, volumes) are still the seeded synthetic ones
because no dataset is reachable offline.
"""
from __future__ import annotations

import argparse
import math
import os

import numpy as np
import torch


def to8b(x):
    """test_loop.py:27: (255 * clip(x, 0, 1)).astype(uint8)."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def orbit_cameras(scene: dict, n_views: int, H: int, W: int, dist_m: float = 3.0):
    """`n_views` cameras on a circle around the target body, looking at its centroid (the synthetic scenes' camera model)."""
    from .synthetic import _look_at
    centre = scene['input_data']['vertices'][0].double().mean(0).cpu().numpy()
    cams = []
    for v in range(n_views):
        a = 2 * math.pi * v / n_views
        eye = centre + dist_m * np.array([math.sin(a), 0.15, math.cos(a)])
        R, T = _look_at(eye, centre)
        K = np.array([[1.2 * W, 0, W / 2], [0, 1.2 * W, H / 2], [0, 0, 1]], np.float64)
        cams.append({'K': K, 'R': R, 'T': T, 'bounds': scene['camera']['bounds']})
    return cams


def render_orbit(renderer, decoder, scene: dict, n_views: int, H: int, W: int, out_dir: str | None = None, pose_index: int = 0):
    """Renders `n_views` novel views of the scene's target pose (device-made rays, one frame per view through dist.render_sequence, so
    it also runs frame-parallel under torch.distributed) and, on rank 0, writes the files test_loop.py writes for its predictions.
    Returns the list of [H*W,5] tensors (rgb in (-1,1) | depth | acc)."""
    from . import dist as sd
    frames = [{'params': scene['input_data']['params'], 'vertices': scene['input_data']['vertices'], 'camera': cam}
              for cam in orbit_cameras(scene, n_views, H, W)]
    outs = sd.render_sequence(renderer, decoder, scene, frames, H, W)
    rank0 = not (torch.distributed.is_available() and torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0
    if out_dir is not None and rank0:
        from PIL import Image
        os.makedirs(out_dir, exist_ok=True)
        for v, o in enumerate(outs):
            img = (o[:, :3].reshape(H, W, 3) / 2 + 0.5).cpu().numpy()                       # test_loop.py:197
            stem = os.path.join(out_dir, 'frame{:04d}_view{:04d}'.format(pose_index, v))    # test_loop.py:218
            Image.fromarray(to8b(img)).save(stem + '.png')
            np.save(stem + '_depth.npy', o[:, 3].reshape(H, W).cpu().numpy())
            np.save(stem + '_acc.npy', o[:, 4].reshape(H, W).cpu().numpy())
    return outs


def main():
    ap = argparse.ArgumentParser(description=__doc__.split('\n\n')[0])
    ap.add_argument('--out', required=True)
    ap.add_argument('--views', type=int, default=8)
    ap.add_argument('--res', type=int, default=512)
    ap.add_argument('--samples', type=int, default=64)
    ap.add_argument('--importance', type=int, default=0)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--weights', default=None, help="torch-saved state dict with 'renderer.*' / 'decoder.*' hot-path names (default: seeded random init)")
    ap.add_argument('--network', default=None, help='reference snapshot (network-snapshot-*.pkl, training_loop.py:563-579); needs --reference-root')
    ap.add_argument('--reference-root', default=None, help="the reference's `sherf/` directory (dnnlib, torch_utils, legacy.py)")
    ap.add_argument('--which', default='G_ema', help='entry of the snapshot dict to load (G or G_ema)')
    args = ap.parse_args()
    from . import synthetic as S
    from .triplane import hot_path_modules
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    model = S.make_smpl_model(0)
    scene = S.make_scene(S.SceneSpec(H=args.res, W=args.res, samples=args.samples, seed=args.seed), model)
    scene['rendering_options']['depth_resolution_importance'] = args.importance

    def mv(x):
        if torch.is_tensor(x):
            return x.to(dev)
        if isinstance(x, dict):
            return {k: mv(v) for k, v in x.items()}
        if isinstance(x, list):
            return [mv(v) for v in x]
        return x
    scene = {k: mv(v) for k, v in scene.items()}
    if args.network:
        if not args.reference_root:
            ap.error('--network needs --reference-root (the snapshot is unpickled with the reference\'s legacy.py / torch_utils)')
        from .checkpoint import load_generator
        G = load_generator(args.network, args.reference_root, which=args.which)
        ren, dec = G.renderer, G.decoder
        if ren.SMPL_NEUTRAL is None:
            ren.set_smpl_model(model)
    else:
        ren, dec = hot_path_modules(model, seed=args.seed, dense_sigma=args.weights is None)
    if args.weights:
        sd = torch.load(args.weights, map_location='cpu')
        ren.load_state_dict({k[len('renderer.'):]: v for k, v in sd.items() if k.startswith('renderer.')}, strict=False)
        dec.load_state_dict({k[len('decoder.'):]: v for k, v in sd.items() if k.startswith('decoder.')})
    outs = render_orbit(ren.to(dev), dec.to(dev), scene, args.views, args.res, args.res, args.out)
    print(f'wrote {len(outs)} views of {args.res}x{args.res} x {args.samples}+{args.importance} samples to {args.out}')


if __name__ == '__main__':
    main()
