"""Launched by tests/test_dist_gpu.py under torchrun (one rank per GPU, NCCL): renders one view (a) sharded over the ranks in
interleaved 256-ray tiles with ONE all-gather (sherf_b200.dist.render_sharded) and (b) whole on every rank, and checks on the hardware
that the gathered image equals the single-GPU image bit for bit -- coarse only and with the importance pass.  Prints `DIST_OK` on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from conftest import scene_to  # noqa: E402
from sherf_b200 import dist as sd, synthetic as S  # noqa: E402
from sherf_b200.triplane import hot_path_modules  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    model = S.make_smpl_model(0)
    ren, dec = hot_path_modules(model, seed=0, dense_sigma=True)
    ren, dec = ren.to(dev), dec.to(dev)
    ok = True
    for spec, n_imp in ((S.SceneSpec(H=128, W=160, samples=32, seed=3), 0), (S.SceneSpec(H=96, W=96, samples=24, seed=4, random_global_R=True), 16)):
        scene = scene_to(S.make_scene(spec, model), dev)
        scene['rendering_options']['depth_resolution_importance'] = n_imp
        N = spec.H * spec.W
        u = torch.rand(N, n_imp, generator=torch.Generator().manual_seed(1)).to(dev) if n_imp else None

        whole = ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None, scene['obs_sp_input'], dec,
                    scene['ray_origins'], scene['ray_directions'], scene['near'], scene['far'], scene['input_data'], scene['rendering_options'],
                    importance_u=u)
        got = sd.render_sharded(ren, dec, scene, importance_u=u)
        same = all(torch.equal(a, b) for a, b in zip(whole, got))
        flag = torch.tensor([0.0 if same else 1.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        ok &= float(flag) == 0.0
        if rank == 0:
            print(f'[nccl x{world}] {spec.H}x{spec.W}x{spec.samples}+{n_imp}: gathered == single-GPU image on every rank: {float(flag) == 0.0}', flush=True)
    dist.destroy_process_group()
    if rank == 0:
        print('DIST_OK' if ok else 'DIST_MISMATCH', flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
