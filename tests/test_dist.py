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
