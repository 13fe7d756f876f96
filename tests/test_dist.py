"""CPU (gloo, world_size 2) tests of the N>1 host logic: tile sharding, the single all-gather, reassembly."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sherf_b200 import dist as sd


@pytest.mark.parametrize('n,world', [(1, 2), (255, 2), (256, 2), (1000, 2), (512 * 512, 8), (230400, 4), (777, 3)])
def test_shards_partition_the_rays(n, world):
    seen = torch.zeros(n, dtype=torch.int32)
    for r in range(world):
        idx = sd.shard_indices(n, r, world)
        assert idx.numel() <= sd.padded_shard_size(n, world)
        assert torch.all(idx[1:] > idx[:-1]) if idx.numel() > 1 else True
        seen[idx] += 1
    assert torch.all(seen == 1)


def test_depth_range_matches_reference_formula():
    torch.manual_seed(0)
    near, far = torch.rand(1, 500, 1) * 3, torch.rand(1, 500, 1) * 3 + 3
    S = 7
    steps = torch.arange(S, dtype=torch.float32) / (S - 1)
    depths = near[0] + steps[None] * (far - near)[0]              # math_utils.py:101-118
    lo, hi = sd.depth_range(near, far, S)
    assert lo == float(depths.min()) and hi == float(depths.max())


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        idx = sd.shard_indices(n, rank, world)
        local = torch.stack([idx.float() * (k + 1) for k in range(5)], 1)       # "rendered" value = f(ray index)
        full = sd.all_gather_tiles(local, n)
        want = torch.stack([torch.arange(n).float() * (k + 1) for k in range(5)], 1)
        q.put((rank, bool(torch.equal(full, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [1000, 64 * 64, 300])
def test_all_gather_tiles_gloo_world2(n):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _seq_worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        H, W = 6, 7
        rendered = []

        def fake_render(frame):                                  # "image" = f(frame id, ray index); records who rendered what
            rendered.append(frame['id'])
            return torch.arange(H * W).float()[:, None] * torch.arange(1, 6).float()[None] + 1000.0 * frame['id']
        frames = [{'id': f} for f in range(n_frames)]
        scene = {'planes': torch.zeros(1)}
        outs = sd.render_sequence(None, None, scene, frames, H, W, render_fn=fake_render)
        ok = len(outs) == n_frames and all(torch.equal(o, fake_render({'id': f})) for f, o in enumerate(outs))
        q.put((rank, bool(ok), sorted(set(rendered[:len(rendered) - n_frames]))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_frames', [4, 5])
def test_render_sequence_round_robin_gloo_world2(n_frames):
    """configs[3] host logic: frame f goes to rank f % world, one all-gather per round, partial last round handled."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() + n_frames) % 2000
    procs = [ctx.Process(target=_seq_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[:2] for r in res] == [(0, True), (1, True)]
    assert res[0][2] == [f for f in range(n_frames) if f % 2 == 0] and res[1][2] == [f for f in range(n_frames) if f % 2 == 1]


def _sharded_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        scene = {'ray_origins': torch.randn(1, n, 3), 'ray_directions': torch.randn(1, n, 3), 'near': torch.rand(1, n, 1),
                 'far': torch.rand(1, n, 1) + 2, 'planes': None, 'obs_input_img': None, 'obs_input_feature': None, 'volumes': None,
                 'obs_sp_input': None, 'input_data': None, 'rendering_options': {'depth_resolution': 8, 'depth_resolution_importance': 4}}
        u = torch.rand(n, 4)
        seen = {}

        def fake_renderer(planes, img, feat, vol, mask, sp, decoder, o, d, near, far, idt, opts, depth_clamp=None, importance_u=None):
            # a "render" that depends on the ray, on ITS row of the importance draws and on the global depth range
            seen['clamp'] = depth_clamp
            val = o.sum(-1, keepdim=True) + importance_u.sum(-1)[None, :, None] + depth_clamp[1]
            return val.expand(-1, -1, 3).contiguous(), near + val, far + val
        rgb, depth, acc = sd.render_sharded(fake_renderer, None, scene, importance_u=u)
        clamp = sd.depth_range(scene['near'], scene['far'], 8)
        val = scene['ray_origins'].sum(-1, keepdim=True) + u.sum(-1)[None, :, None] + clamp[1]
        ok = torch.equal(rgb, val.expand(-1, -1, 3)) and torch.equal(depth, scene['near'] + val) and torch.equal(acc, scene['far'] + val)
        q.put((rank, bool(ok), seen['clamp'] == clamp))
    finally:
        dist.destroy_process_group()


def test_render_sharded_passes_each_rank_its_rows_of_the_importance_draws():
    """dist.render_sharded on CPU tensors with a stand-in renderer: every rank gets its rays, ITS rows of the full view's draws and the
    full view's depth range; the reassembled image equals the unsharded one."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, 1000, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]
