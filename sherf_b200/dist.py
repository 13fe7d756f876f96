"""Multi-GPU ray sharding.  Rays are independent; the only cross-ray coupling in the reference is the
global depth clamp `torch.min(depths) / torch.max(depths)` (ray_marcher.py:57), which every rank computes from the
replicated near/far.  One process per GPU renders an interleaved set of ray tiles and the rendered tiles
(rgb 3 + depth 1 + acc 1 = 20 B per ray) are exchanged with exactly ONE all-gather; nothing else crosses ranks.
The reference has no such path (its evaluation runs on rank 0 only, training_loop.py:311-328).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TILE = 256          # consecutive rays per tile; tiles are dealt round-robin so that body / background pixels balance


def shard_indices(n_rays: int, rank: int, world: int, tile: int = TILE, device='cpu') -> torch.Tensor:
    """Ray indices owned by `rank`: tiles t with t % world == rank, in ascending order."""
    idx = torch.arange(n_rays, device=device)
    return idx[(idx // tile) % world == rank]


def padded_shard_size(n_rays: int, world: int, tile: int = TILE) -> int:
    """Largest shard over all ranks (ranks pad to it so one fixed-size all-gather suffices)."""
    n_tiles = (n_rays + tile - 1) // tile
    most = (n_tiles + world - 1) // world
    return most * tile


def depth_range(near: torch.Tensor, far: torch.Tensor, n_samples: int):
    """(min, max) over ALL sample depths of the full view, bit-identical to torch.min/max(depths_coarse):
    t_0 = near + 0 * (far - near) = near and t_{S-1} = near + 1 * (far - near) (math_utils.py:101-118), t monotone in i."""
    first = near.float().reshape(-1)
    last = first + (far.float().reshape(-1) - first)
    lo = torch.minimum(first, last).min()
    hi = torch.maximum(first, last).max()
    return float(lo), float(hi)


def shard_scene(scene: dict, rank: int, world: int, tile: int = TILE):
    """A shallow copy of a scene dict (synthetic.make_scene layout) that carries only `rank`'s rays."""
    n = scene['ray_origins'].shape[1]
    idx = shard_indices(n, rank, world, tile, device=scene['ray_origins'].device)
    sh = dict(scene)
    for k in ('ray_origins', 'ray_directions', 'near', 'far'):
        sh[k] = scene[k][:, idx].contiguous()
    return sh, idx


_perm_cache: dict = {}


def _gather_permutation(n_rays: int, world: int, tile: int, device) -> torch.Tensor:
    """perm[ray] = row of that ray in the [world * pad] all-gather buffer (built once per shape, cached)."""
    key = (n_rays, world, tile, str(device))
    perm = _perm_cache.get(key)
    if perm is None:
        pad = padded_shard_size(n_rays, world, tile)
        perm = torch.empty(n_rays, dtype=torch.long)
        for r in range(world):
            idx = shard_indices(n_rays, r, world, tile)
            perm[idx] = r * pad + torch.arange(idx.numel())
        perm = perm.to(device)
        _perm_cache[key] = perm
    return perm


def all_gather_tiles(local: torch.Tensor, n_rays: int, group=None, tile: int = TILE) -> torch.Tensor:
    """local: [n_local, C] rendered values of this rank's rays (shard_indices order).  Returns the full [n_rays, C]
    on every rank using a single all_gather of equally padded shards followed by one index_select."""
    world = dist.get_world_size(group)
    pad = padded_shard_size(n_rays, world, tile)
    if local.shape[0] == pad:
        buf = local.contiguous()
    else:
        buf = local.new_zeros(pad, local.shape[1])
        buf[:local.shape[0]] = local
    gathered = local.new_empty(world * pad, local.shape[1])
    dist.all_gather_into_tensor(gathered, buf, group=group)
    return gathered.index_select(0, _gather_permutation(n_rays, world, tile, local.device))


def render_sharded(renderer, decoder, scene: dict, group=None, tile: int = TILE, importance_u: torch.Tensor | None = None):
    """Every rank holds the replicated scene, renders its ray tiles through the CUDA path and receives the full image.
    Returns (rgb[1,N,3], depth[1,N,1], acc[1,N,1]) on every rank.  `importance_u` ([N,S_f], replicated) holds the
    importance-pass draws of the FULL view (renderer.py:526); each rank uses the rows of its own rays, so the sharded
    render equals the single-GPU render of the same draws."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = scene['ray_origins'].shape[1]
    S = int(scene['rendering_options']['depth_resolution'])
    clamp = depth_range(scene['near'], scene['far'], S)
    sh, idx = shard_scene(scene, rank, world, tile)
    if idx.numel() > 0:
        rgb, depth, acc = renderer(sh['planes'], sh['obs_input_img'], sh['obs_input_feature'], sh['volumes'], None,
                                   sh['obs_sp_input'], decoder, sh['ray_origins'], sh['ray_directions'], sh['near'], sh['far'],
                                   sh['input_data'], sh['rendering_options'], depth_clamp=clamp,
                                   importance_u=None if importance_u is None else importance_u[idx.to(importance_u.device)].contiguous())
        local = torch.cat([rgb[0], depth[0], acc[0]], dim=-1)
    else:
        local = scene['ray_origins'].new_zeros(0, 5)
    full = all_gather_tiles(local, n, group, tile)
    return full[None, :, :3], full[None, :, 3:4], full[None, :, 4:5]


def _render_frame(renderer, decoder, scene: dict, frame: dict, H: int, W: int):
    """One frame of a sequence through the CUDA path: rays made on the device (sherf_b200.rays), the frame's target pose
    swapped into a shallow copy of input_data.  Returns [N,5] = rgb | depth | acc."""
    from .rays import generate_rays
    dev = scene['planes'].device
    cam = frame['camera']
    rays = generate_rays(H, W, cam['K'], cam['R'], cam['T'], cam['bounds'], dev)
    idt = dict(scene['input_data'])
    idt['params'] = {k: torch.as_tensor(v).to(dev) for k, v in frame['params'].items()}
    idt['vertices'] = torch.as_tensor(frame['vertices']).to(dev)
    rgb, depth, acc = renderer(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None,
                               scene['obs_sp_input'], decoder, rays['ray_o_all'][:, 0], rays['ray_d_all'][:, 0], rays['near_all'][:, 0],
                               rays['far_all'][:, 0], idt, scene['rendering_options'])
    return torch.cat([rgb[0], depth[0], acc[0]], dim=-1)


def render_sequence(renderer, decoder, scene: dict, frames: list, H: int, W: int, group=None, render_fn=None):
    """BASELINE configs[3]: a streamed novel-pose / novel-view sequence of one observed subject over the ranks of `group`.

    `scene` holds the static observation (planes, 2-D feature map, 3-D volumes, obs_* and t_* entries of input_data, replicated on
    every rank); `frames[f]` = {'params': {poses, shapes, R, Th}, 'vertices': [1,V,3], 'camera': {K, R, T, bounds}} is all that is
    uploaded per frame -- the rays are generated on the device (RenderPeople_dataset.py:14-27,68-101 on the GPU).  Frames are dealt
    round-robin: frame f is rendered by rank f % world; after every round of `world` frames exactly ONE all-gather (20 B per ray)
    hands the round's images to every rank.  Whole frames have no cross-ray coupling, so no depth-range exchange is needed.
    Returns a list of [N,5] tensors (rgb | depth | acc), one per frame, on every rank."""
    use_dist = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if use_dist else 1
    rank = dist.get_rank(group) if use_dist else 0
    fn = render_fn or (lambda fr: _render_frame(renderer, decoder, scene, fr, H, W))
    N = H * W
    out = []
    for f0 in range(0, len(frames), world):
        mine = f0 + rank
        local = fn(frames[mine]) if mine < len(frames) else None
        if world == 1:
            out.append(local)
            continue
        if local is None:                                   # last, partial round: ranks without a frame contribute a dummy tile
            ref = out[-1] if out else None
            local = torch.zeros(N, 5, device=ref.device if ref is not None else scene['planes'].device)
        gathered = local.new_empty(world * N, 5)
        dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)
        for r in range(min(world, len(frames) - f0)):
            out.append(gathered[r * N:(r + 1) * N])
    return out
