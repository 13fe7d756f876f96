"""TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this.  (Device-agnostic: on CUDA tensors it is "the reference written as eager
PyTorch on the same GPU", the bar BASELINE.md 3.4 asks to report next to the CUDA path -- bench.py --gpu-eager-baseline.)

CPU restatement (torch, fp32) of the reference's hot path -- ImportanceRenderer.forward
(renderer.py:286-398) + NeRFDecoder.forward (triplane.py:285-316) -- for per-GPU batch 1.  It exists
because the reference is Python that cannot travel to the GPU box; it is validated against the
reference's own code (oracle/ref_shim.py) in tests/test_oracle.py and against tests/golden/*.npz.

PARITY STATUS: "parity unpinned" -- the reference ships no tests / golden vectors / checkpoints
(SURVEY.md §4), and its KNN comes from pytorch3d (un-vendored, version unpinned, README.md:48).  The
pin is: this file == reference code under shims, on identical synthetic inputs, stage by stage.

Index bookkeeping (depth indices, cull mask, nearest-vertex ids, compaction order) is defined here
with an explicit rounding order so that the CUDA path can be bit-exact against it:
  step_i = fl(i / (S-1));  t = fl(near + fl(step_i * fl(far - near)))          math_utils.py:101-118
  x = fl(o + fl(t * d))                                                        renderer.py:304
  q_j = fma(p2, R2j, fma(p1, R1j, fl(p0 * R0j))),  p = fl(x - Th)              renderer.py:309 (what torch's CPU matmul does)
  d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)); argmin, smallest index wins  renderer.py:315 (our statement of knn_points)
  mask = d2 < fl(0.05**2 as python double -> compared in fp32)                 renderer.py:316-319
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

CULL_THRESHOLD = 0.05 ** 2          # python double; `distance < threshold` compares fp32 against it


# ----------------------------------------------------------------------------- helpers

def fma32(a, b, c):
    """Single-rounded a*b+c in fp32 (product of two fp32 is exact in fp64)."""
    return (a.double() * b.double() + c.double()).float()


def matvec3_rows(p, M):
    """p[...,3] @ M[3,3] with the k-ordered FMA chain torch's CPU sgemm uses for K=3."""
    cols = []
    for j in range(3):
        acc = p[..., 0] * M[0, j]
        acc = fma32(p[..., 1], M[1, j], acc)
        acc = fma32(p[..., 2], M[2, j], acc)
        cols.append(acc)
    return torch.stack(cols, -1)


def knn1(q, v, chunk=8192):
    """K=1 nearest neighbour, squared L2 in x,y,z order, smallest index on ties.  q[P,3], v[V,3]."""
    d2_out = torch.empty(q.shape[0], dtype=torch.float32, device=q.device)
    id_out = torch.empty(q.shape[0], dtype=torch.long, device=q.device)
    vx, vy, vz = v[:, 0][None], v[:, 1][None], v[:, 2][None]
    for s in range(0, q.shape[0], chunk):
        c = q[s:s + chunk]
        dx, dy, dz = c[:, 0:1] - vx, c[:, 1:2] - vy, c[:, 2:3] - vz
        d2 = (dx * dx + dy * dy) + dz * dz
        m = d2.min(dim=1).values
        d2_out[s:s + chunk] = m
        id_out[s:s + chunk] = (d2 == m[:, None]).float().argmax(dim=1)
    return d2_out, id_out


def rodrigues(rvec):
    """renderer.py:76-94 / :159-190 (both variants are the same arithmetic). rvec[n,3] -> [n,3,3]."""
    angle = torch.norm(rvec + 1e-8, p=2, dim=1, keepdim=True)
    k = rvec / angle
    c, s = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    z = torch.zeros_like(k[:, :1])
    K = torch.cat([z, -k[:, 2:3], k[:, 1:2], k[:, 2:3], z, -k[:, 0:1], -k[:, 1:2], k[:, 0:1], z], 1).reshape(-1, 3, 3)
    return torch.eye(3, device=rvec.device)[None] + s * K + (1 - c) * torch.matmul(K, K)


def lbs_transforms(smpl, poses, shapes):
    """renderer.py:96-157: shape blend -> joints -> Rodrigues -> kinematic chain -> 24 rigid [4,4]."""
    v_shaped = smpl['v_template'] + (smpl['shapedirs'] * shapes.reshape(1, 1, 10)).sum(-1)
    joints = smpl['J_regressor'] @ v_shaped
    rot = rodrigues(poses.reshape(-1, 3))
    parents = smpl['kintree_table'][0]
    rel = joints.clone()
    rel[1:] -= joints[parents[1:]]
    local = torch.zeros(24, 4, 4, device=joints.device)
    local[:, :3, :3] = rot
    local[:, :3, 3] = rel
    local[:, 3, 3] = 1
    chain = [local[0]]
    for i in range(1, 24):
        chain.append(chain[int(parents[i])] @ local[i])
    A = torch.stack(chain, 0)
    jh = torch.cat([joints, torch.zeros(24, 1, device=joints.device)], 1)
    A[..., 3] = A[..., 3] - (A * jh[:, None, :]).sum(-1)
    return A


def pose_offsets(smpl, poses):
    """renderer.py:578-584: posedirs @ (R(theta)[1:] - I)."""
    rot = rodrigues(poses.reshape(-1, 3))
    feat = (rot[1:] - torch.eye(3, device=rot.device)).reshape(1, -1)
    return (feat @ smpl['posedirs'].reshape(6890 * 3, -1).t()).reshape(-1, 3)


def shape_offsets(smpl, shapes):
    """renderer.py:590-591."""
    return torch.matmul(smpl['shapedirs'], shapes.reshape(10, 1)).squeeze(-1)


def positional_encoding(x, num_freqs):
    """renderer.py:875-916: [x, sin(x f0), sin(x f0 + pi/2), sin(x f1), ...] with the reference's interleave:
    embed[:, 2k+b, c] = sin(phase_b + x_c * 2^k) flattened as (freq-pair major, coordinate minor)."""
    freqs = torch.repeat_interleave(2. ** torch.linspace(0., num_freqs - 1, steps=num_freqs), 2).view(1, -1, 1).to(x.device)
    phases = torch.zeros(2 * num_freqs, device=x.device)
    phases[1::2] = torch.pi * 0.5
    phases = phases.view(1, -1, 1)
    e = x.unsqueeze(1).repeat(1, num_freqs * 2, 1)
    e = torch.sin(torch.addcmul(phases, e, freqs)).view(x.shape[0], num_freqs * 2 * x.shape[1])
    return torch.cat((x, e), dim=-1)


# ----------------------------------------------------------------------------- stages

def sample_depths(near, far, S):
    """renderer.py:458-481 (tensor branch, jitter commented out) -> [N,S]."""
    steps = torch.arange(S, dtype=torch.float32, device=near.device) / (S - 1)
    return near.reshape(-1, 1) + steps[None] * (far - near).reshape(-1, 1)


def cull(scene_in, S, depths=None):
    """renderer.py:299-321: depths, SMPL-space queries, nearest posed vertex, 5 cm mask.  `depths` [N,S'] overrides the
    stratified depths (the repaired fine pass pushes the importance samples through the same lines)."""
    idt = scene_in['input_data']
    o, d = scene_in['ray_origins'][0], scene_in['ray_directions'][0]
    if depths is None:
        depths = sample_depths(scene_in['near'][0], scene_in['far'][0], S)        # [N,S]
    S = depths.shape[1]
    x = (o[:, None, :] + depths[..., None] * d[:, None, :]).reshape(-1, 3)
    dirs = d[:, None, :].expand(-1, S, -1).reshape(-1, 3)
    R, Th = idt['params']['R'][0], idt['params']['Th'][0]
    q = matvec3_rows(x - Th, R)
    vdir = matvec3_rows(dirs, R)
    verts = matvec3_rows(idt['vertices'][0] - Th, R)
    d2, vid = knn1(q, verts)
    mask = d2 < CULL_THRESHOLD
    return {'depths': depths, 'q': q, 'vdir': vdir, 'verts_smpl': verts, 'd2': d2, 'id1': vid, 'mask': mask}


def warp_to_canonical(smpl, params, t_params, q, vdir, vid):
    """renderer.py:558-621 with nearest-vertex blend weights (vid = nearest posed vertex)."""
    bw = smpl['weights'][vid]                                                     # [P,24]
    A = (bw @ lbs_transforms(smpl, params['poses'][0], params['shapes'][0]).reshape(24, 16)).reshape(-1, 4, 4)
    Rinv = torch.inverse(A[:, :3, :3])
    can = torch.matmul(Rinv, (q - A[:, :3, 3])[..., None]).squeeze(-1)
    cdir = torch.matmul(Rinv, vdir[..., None]).squeeze(-1)
    can = can - pose_offsets(smpl, params['poses'][0])[vid]
    can = can - shape_offsets(smpl, params['shapes'][0])[vid]
    can = can + pose_offsets(smpl, t_params['poses'][0])[vid]
    Ab = (bw @ lbs_transforms(smpl, t_params['poses'][0], t_params['shapes'][0]).reshape(24, 16)).reshape(-1, 4, 4)
    can = torch.matmul(Ab[:, :3, :3], can[..., None]).squeeze(-1) + Ab[:, :3, 3]
    cdir = torch.matmul(Ab[:, :3, :3], cdir[..., None]).squeeze(-1)
    return can, cdir


def warp_to_observation(smpl, obs_params, t_params, t_vertices, can):
    """renderer.py:623-684: canonical -> T -> observation pose -> world; nearest canonical vertex (knn #3)."""
    _, vid = knn1(can, t_vertices)
    bw = smpl['weights'][vid]
    bw = bw / bw.sum(-1, keepdim=True)
    Ab = (bw @ lbs_transforms(smpl, t_params['poses'][0], t_params['shapes'][0]).reshape(24, 16)).reshape(-1, 4, 4)
    p = torch.matmul(torch.inverse(Ab[:, :3, :3]), (can - Ab[:, :3, 3])[..., None]).squeeze(-1)
    p = p - pose_offsets(smpl, t_params['poses'][0])[vid]
    p = p + shape_offsets(smpl, obs_params['shapes'][0])[vid]
    p = p + pose_offsets(smpl, obs_params['poses'][0])[vid]
    Ao = (bw @ lbs_transforms(smpl, obs_params['poses'][0], obs_params['shapes'][0]).reshape(24, 16)).reshape(-1, 4, 4)
    p = torch.matmul(Ao[:, :3, :3], p[..., None]).squeeze(-1) + Ao[:, :3, 3]
    world = torch.matmul(p, torch.inverse(obs_params['R'][0])) + obs_params['Th'][0]
    return world, vid


def project(world, Rc, Tc, Kc):
    """renderer.py:686-704 (face=None)."""
    cam = torch.matmul(Rc, world[..., None]) + Tc
    pix = torch.matmul(Kc, cam)[..., 0]
    return pix[:, :2] / (pix[:, 2:] + 1e-5)


def gather_2d(obs_img, obs_feat, uv):
    """renderer.py:331-340: uv normalised by the IMAGE size, sampled (align_corners=True) from both maps."""
    g = 2.0 * uv[None, :, None, :] / torch.tensor([obs_img.shape[-1], obs_img.shape[-2]], dtype=torch.float32, device=uv.device) - 1.0
    feat = F.grid_sample(obs_feat, g, align_corners=True)[0, :, :, 0].t()
    rgb = F.grid_sample(obs_img, g, align_corners=True)[0, :, :, 0].t()
    return torch.cat([feat, positional_encoding(rgb, 5)[:, :32]], -1)


def gather_3d(volumes, sp_bounds, out_sh, can):
    """renderer.py:544-556 + :762-797: voxel coords (0.005 m), 3 trilinear gathers, concat 192."""
    dhw = (can[:, [2, 1, 0]] - sp_bounds[0][[2, 1, 0]]) / torch.tensor([0.005, 0.005, 0.005], device=can.device)
    dhw = dhw / torch.tensor(out_sh, dtype=torch.float32, device=can.device) * 2 - 1
    g = dhw[:, [2, 1, 0]][None, None, None]
    fs = [F.grid_sample(v, g, padding_mode='zeros', align_corners=True) for v in volumes]
    fs = torch.cat(fs, 1)
    return fs.view(1, -1, fs.size(4))[0].t()


def gather_triplane(planes, can, box):
    """renderer.py:192-243: normalise by t_world_bounds; planes (x,y), (x,z), (z,y); align_corners=False."""
    c = 2 * (can - box[0]) / (box[1] - box[0]) - 1
    coords = torch.stack([c[:, [0, 1]], c[:, [0, 2]], c[:, [2, 1]]], 0)          # [3,P,2]
    out = F.grid_sample(planes[0], coords[:, None], mode='bilinear', padding_mode='zeros', align_corners=False)
    return out[:, :, 0].permute(0, 2, 1)                                          # [3,P,32]


def fuse_and_decode(w, tri, f2d, f3d, can, cdir):
    """renderer.py:423-432 + Transformer (:920-993) + NeRFDecoder.forward (triplane.py:285-316)."""
    P = can.shape[0]
    comb = torch.cat([tri, f2d.reshape(P, 3, 32).permute(1, 0, 2), f3d.reshape(P, 3, 32).permute(1, 0, 2)], -1)   # [3,P,96]
    tok = F.linear(comb, w['renderer.conv1d_reprojection.weight'][:, :, 0], w['renderer.conv1d_reprojection.bias'])
    x = tok.permute(1, 0, 2)                                                      # [P,3,32]
    pre = x
    t = 'renderer.transformer.layers.0.'
    h = F.layer_norm(x, (32,), w[t + '0.fn.norm.weight'], w[t + '0.fn.norm.bias'])
    qkv = F.linear(h, w[t + '0.fn.fn.to_qkv.weight']).reshape(P, 3, 3, 3, 16)    # [P, tok, (q|k|v), head, 16]
    qh, kh, vh = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))             # [P, head, tok, 16]
    att = torch.softmax(torch.matmul(qh, kh.transpose(-1, -2)) * (16 ** -0.5), dim=-1)
    o = torch.matmul(att, vh).permute(0, 2, 1, 3).reshape(P, 3, 48)
    x = F.linear(o, w[t + '0.fn.fn.to_out.0.weight'], w[t + '0.fn.fn.to_out.0.bias']) + x
    h = F.layer_norm(x, (32,), w[t + '1.fn.norm.weight'], w[t + '1.fn.norm.bias'])
    h = F.linear(F.gelu(F.linear(h, w[t + '1.fn.fn.net.0.weight'], w[t + '1.fn.fn.net.0.bias'])),
                 w[t + '1.fn.fn.net.3.weight'], w[t + '1.fn.fn.net.3.bias'])
    x = h + x
    tok0, tok1 = x[:, 0], x[:, 1]
    xin = torch.cat([positional_encoding(can, 6), tok0], -1)                      # 39 + 32 = 71
    h = xin
    for i in range(8):
        h = F.relu(F.linear(h, w[f'decoder.pts_linears.{i}.weight'], w[f'decoder.pts_linears.{i}.bias']))
        if i == 4:
            h = torch.cat([xin, h], -1)
    sigma = F.linear(h, w['decoder.alpha_linear.weight'], w['decoder.alpha_linear.bias'])
    feat = F.linear(h, w['decoder.feature_linear.weight'], w['decoder.feature_linear.bias'])
    h = torch.cat([feat, positional_encoding(cdir, 4), tok1], -1)                 # 128 + 27 + 32 = 187
    h = F.relu(F.linear(h, w['decoder.views_linear.weight'], w['decoder.views_linear.bias']))
    rgb = torch.sigmoid(F.linear(h, w['decoder.rgb_linear.weight'], w['decoder.rgb_linear.bias'])) * (1 + 2 * 0.001) - 0.001
    return {'tok_pre': pre, 'tok_post': x, 'sigma': sigma[:, 0], 'rgb': rgb}


def composite(colors, sigma, depths, rays_d, white_back, depth_clamp=None):
    """ray_marcher.py:25-64, clamp_mode='relu'.  colors[N,S,3] sigma[N,S] depths[N,S] rays_d[N,3].
    `depth_clamp` = (min, max) of the depths of the FULL view when `depths` holds only a subset of its rays (:57 takes
    torch.min / torch.max over all rays of the view)."""
    deltas = torch.cat([depths[:, 1:] - depths[:, :-1], torch.full_like(depths[:, :1], 1e10)], 1)
    deltas = deltas * torch.norm(rays_d, dim=-1, keepdim=True)
    alpha = 1 - torch.exp(-(F.relu(sigma) * deltas))
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], 1), 1)[:, :-1]
    wts = alpha * trans
    rgb = (wts[..., None] * colors).sum(1)
    wsum = wts.sum(1, keepdim=True)
    depth = (wts * depths).sum(1, keepdim=True) / wsum
    lo, hi = (depths.min(), depths.max()) if depth_clamp is None else depth_clamp
    depth = torch.clamp(torch.nan_to_num(depth, float('inf')), lo, hi)
    if white_back:
        rgb = rgb + 1 - wsum
    return rgb * 2 - 1, depth, wts


def sample_importance(depths, wts, n_importance, u):
    """renderer.py:483-501 (sample_importance) + :503-542 (sample_pdf, det=False) with the uniform draws `u` [N,S_f]
    supplied by the caller in place of torch.rand (:526).  depths[N,S], wts[N,S] (ray-marcher weights) -> (t_fine[N,S_f],
    bin index `inds` [N,S_f] = searchsorted(cdf, u, right=True))."""
    w = F.max_pool1d(wts.unsqueeze(1).float(), 2, 1, padding=1)
    w = F.avg_pool1d(w, 2, 1).squeeze(1)
    w = w + 0.01
    bins = 0.5 * (depths[:, :-1] + depths[:, 1:])                                 # [N,S-1]
    w = w[:, 1:-1]                                                                # [N,S-2]
    n_bins = w.shape[1]
    w = w + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)                      # [N,S-1]
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_bins)
    idx = torch.stack([below, above], -1).view(u.shape[0], 2 * n_importance)
    cdf_g = torch.gather(cdf, 1, idx).view(u.shape[0], n_importance, 2)
    bins_g = torch.gather(bins, 1, idx).view(u.shape[0], n_importance, 2)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom[denom < 1e-5] = 1
    return bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0]), inds


def unify_samples(d1, c1, s1, d2, c2, s2):
    """renderer.py:446-456: concatenate coarse and fine samples and sort every ray by depth."""
    d = torch.cat([d1, d2], 1)
    c = torch.cat([c1, c2], 1)
    s_ = torch.cat([s1, s2], 1)
    d, idx = torch.sort(d, dim=1)
    return d, torch.gather(c, 1, idx[..., None].expand(-1, -1, 3)), torch.gather(s_, 1, idx)


def evaluate_samples(weights: dict, smpl: dict, scene: dict, depths=None, density_noise_points=None):
    """renderer.py:299-371: everything between the depth samples and the dense [N,S] colour / density arrays
    (cull, warps, three gathers, fusion, transformer, decoder, scatter-back with sigma = -80 where culled)."""
    idt, opts = scene['input_data'], scene['rendering_options']
    st = cull(scene, opts['depth_resolution'], depths)
    N, S = st['depths'].shape
    sel = st['mask'].nonzero()[:, 0]                                              # row-major [N,S] order
    st['sel'] = sel
    colors = torch.zeros(N * S, 3, device=sel.device)
    sigma = torch.full((N * S,), -80.0, device=sel.device)
    if sel.numel() > 0:                                                           # else every sample keeps sigma = -80
        q, vdir, vid = st['q'][sel], st['vdir'][sel], st['id1'][sel]
        can, cdir = warp_to_canonical(smpl, idt['params'], idt['t_params'], q, vdir, vid)
        world, id3 = warp_to_observation(smpl, idt['obs_params'], idt['t_params'], idt['t_vertices'][0], can)
        uv = project(world, idt['obs_R_all'][0, 0], idt['obs_T_all'][0, 0], idt['obs_K_all'][0, 0])
        f2d = gather_2d(scene['obs_input_img'], scene['obs_input_feature'], uv)
        f3d_raw = gather_3d(scene['volumes'], scene['obs_sp_input']['bounds'][0], scene['obs_sp_input']['out_sh'], can)
        f3d = F.linear(f3d_raw, weights['renderer.conv1d_projection.weight'][:, :, 0], weights['renderer.conv1d_projection.bias'])
        tri = gather_triplane(scene['planes'], can, idt['t_world_bounds'][0])
        dec = fuse_and_decode(weights, tri, f2d, f3d, can, cdir)
        colors[sel] = dec['rgb']
        # renderer.py:435-436: sigma += randn_like(sigma) * density_noise on the surviving points (the draws are an input here)
        sigma[sel] = dec['sigma'] if density_noise_points is None else dec['sigma'] + density_noise_points.reshape(-1)[:sel.numel()]
        st.update({'can': can, 'cdir': cdir, 'id3': id3, 'world_src': world, 'uv': uv, 'f2d': f2d, 'f3d_raw': f3d_raw,
                   'f3d': f3d, 'tri': tri, **dec})
    return colors.view(N, S, 3), sigma.view(N, S), st


@torch.no_grad()
def render_forward(weights: dict, smpl: dict, scene: dict, return_stages: bool = False, importance_u=None, depth_clamp=None,
                   density_noise_points=None):
    """The whole hot path.  `weights`: state-dict names prefixed 'renderer.' / 'decoder.' (SURVEY §8b).
    Returns rgb[1,N,3], depth[1,N,1], acc[1,N,1] (+ stages dict).

    rendering_options['depth_resolution_importance'] = S_f > 0 runs the REPAIRED fine pass (the reference's
    own lines renderer.py:375-393 cannot execute: :376 omits rays_d, :383 calls run_model with the 5-argument EG3D
    signature).  The repair keeps every reference function and changes only the two broken call sites: ray directions
    are passed to the marcher at :376, and the fine depths are pushed through :304-371 again (same cull, warps,
    gathers and decoder as the coarse samples).  `importance_u` [N,S_f] stands for torch.rand at :526.  Parity for
    this row is pinned only by composition of the reference's own sample_importance / unify_samples / ray marcher
    (oracle/ref_shim.render_importance), not by a run of the reference's forward."""
    opts = scene['rendering_options']
    assert opts['clamp_mode'] == 'relu'
    rays_d = scene['ray_directions'][0]
    colors, sigma, st = evaluate_samples(weights, smpl, scene, density_noise_points=density_noise_points)
    n_imp = int(opts.get('depth_resolution_importance', 0) or 0)
    if n_imp > 0:
        assert importance_u is not None, 'the fine pass needs the uniform draws (torch.rand at renderer.py:526)'
        _, _, wts_c = composite(colors, sigma, st['depths'], rays_d, opts['white_back'])
        t_fine, inds = sample_importance(st['depths'], wts_c, n_imp, importance_u)
        colors_f, sigma_f, st_f = evaluate_samples(weights, smpl, scene, t_fine)
        d_all, c_all, s_all = unify_samples(st['depths'], colors, sigma, t_fine, colors_f, sigma_f)
        rgb, depth, wts = composite(c_all, s_all, d_all, rays_d, opts['white_back'], depth_clamp)
        st.update({'coarse_weights': wts_c, 't_fine': t_fine, 'fine_bins': inds, 'fine': st_f})
    else:
        rgb, depth, wts = composite(colors, sigma, st['depths'], rays_d, opts['white_back'], depth_clamp)
    out = (rgb[None], depth[None], wts.sum(1, keepdim=True)[None])
    if not return_stages:
        return out
    st['weights'] = wts
    return out + (st,)


def hot_path_state_dict(renderer_module, decoder_module) -> dict:
    """Flatten the two modules' parameters under the names the checkpoint uses."""
    sd = {'renderer.' + k: v.detach().float() for k, v in renderer_module.state_dict().items()
          if not k.startswith('encoder_3d')}
    sd.update({'decoder.' + k: v.detach().float() for k, v in decoder_module.state_dict().items()})
    return sd


FLOP_PER_DECODED_SAMPLE = 429_248        # a12 + conv1d_projection, MAC x 2
