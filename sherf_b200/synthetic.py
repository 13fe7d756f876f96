"""Seeded synthetic SMPL-shaped body and render scenes.

The real SMPL_NEUTRAL.pkl is licensed and absent, and there are no datasets or
checkpoints, so every test / bench input comes from here.  Shapes, dtypes, key
names and conventions follow what the reference's dataset hands to the renderer:

* body dict keys    -- renderer.py:65-74 (SMPL_to_tensor): v_template[V,3], shapedirs[V,3,10],
                       J_regressor[24,V], kintree_table[2,24], f[F,3], weights[V,24], posedirs[V,3,207]
* input_data keys   -- RenderPeople_dataset.py:362-391 (after DataLoader collation, batch 1)
* rays              -- RenderPeople_dataset.py:14-27 (get_rays; directions un-normalised)
* near / far        -- RenderPeople_dataset.py:68-101,129-134 (bbox slab test; (0,1) for misses)
* canonical pose    -- RenderPeople_dataset.py:222-235 (the "big pose")
* 3-D volume shapes -- triplane.py:174-217 (prepare_sp_input: out_sh = (ceil(extent/0.005) | 31) + 1)

Everything is numpy float64 internally and cast to float32 at the end, so the
same seed yields bit-identical inputs in this container and on the GPU box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

V = 6890
J = 24
PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], dtype=np.int64)

# T-pose joints of a ~1.7 m figure, pelvis at the origin, y up, x to the figure's left.
_TPOSE = np.array([
    [0.00, 0.00, 0.00],     # 0 pelvis
    [0.07, -0.09, 0.00],    # 1 l hip
    [-0.07, -0.09, 0.00],   # 2 r hip
    [0.00, 0.11, -0.01],    # 3 spine1
    [0.10, -0.47, 0.00],    # 4 l knee
    [-0.10, -0.47, 0.00],   # 5 r knee
    [0.00, 0.25, 0.00],     # 6 spine2
    [0.09, -0.87, -0.03],   # 7 l ankle
    [-0.09, -0.87, -0.03],  # 8 r ankle
    [0.00, 0.31, 0.01],     # 9 spine3
    [0.11, -0.92, 0.09],    # 10 l foot
    [-0.11, -0.92, 0.09],   # 11 r foot
    [0.00, 0.51, -0.02],    # 12 neck
    [0.08, 0.42, -0.02],    # 13 l collar
    [-0.08, 0.42, -0.02],   # 14 r collar
    [0.00, 0.60, 0.02],     # 15 head
    [0.19, 0.45, -0.02],    # 16 l shoulder
    [-0.19, 0.45, -0.02],   # 17 r shoulder
    [0.45, 0.45, -0.03],    # 18 l elbow
    [-0.45, 0.45, -0.03],   # 19 r elbow
    [0.70, 0.45, -0.03],    # 20 l wrist
    [-0.70, 0.45, -0.03],   # 21 r wrist
    [0.79, 0.44, -0.02],    # 22 l hand
    [-0.79, 0.44, -0.02],   # 23 r hand
], dtype=np.float64)

# capsule radius per driving joint (bone = parent joint -> child joint, skinned to the parent)
_RADIUS = {0: 0.12, 3: 0.12, 6: 0.12, 9: 0.11, 1: 0.08, 2: 0.08, 4: 0.06, 5: 0.06, 7: 0.04, 8: 0.04,
           12: 0.055, 13: 0.06, 14: 0.06, 15: 0.10, 16: 0.05, 17: 0.05, 18: 0.04, 19: 0.04, 20: 0.035, 21: 0.035}


def _bones():
    """(driving joint, a, b, radius) segments: every parent->child link plus end caps."""
    segs = []
    for c in range(1, J):
        p = int(PARENTS[c])
        segs.append((p, _TPOSE[p], _TPOSE[c], _RADIUS.get(p, 0.05)))
    segs.append((15, _TPOSE[15], _TPOSE[15] + np.array([0.0, 0.16, 0.0]), 0.10))      # skull
    segs.append((22, _TPOSE[22], _TPOSE[22] + np.array([0.08, 0.0, 0.0]), 0.03))       # l fingers
    segs.append((23, _TPOSE[23], _TPOSE[23] + np.array([-0.08, 0.0, 0.0]), 0.03))      # r fingers
    segs.append((10, _TPOSE[10], _TPOSE[10] + np.array([0.0, 0.0, 0.08]), 0.035))      # l toes
    segs.append((11, _TPOSE[11], _TPOSE[11] + np.array([0.0, 0.0, 0.08]), 0.035))      # r toes
    return segs


def _seg_dist(p, a, b):
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(float(ab @ ab), 1e-12), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * ab), axis=1)


def make_smpl_model(seed: int = 0) -> dict:
    """SMPL-shaped dict of numpy arrays (float32 / int64), deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    segs = _bones()
    area = np.array([2 * math.pi * r * (np.linalg.norm(b - a) + 2 * r) for (_, a, b, r) in segs])
    counts = np.floor(area / area.sum() * V).astype(int)
    counts[0] += V - counts.sum()
    pts = []
    for (_, a, b, r), n in zip(segs, counts):
        axis = b - a
        L = np.linalg.norm(axis)
        axis = axis / L
        ref = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
        u = np.cross(axis, ref); u /= np.linalg.norm(u)
        w = np.cross(axis, u)
        s = rng.uniform(-r, L + r, n)                 # along the capsule incl. caps
        th = rng.uniform(0, 2 * math.pi, n)
        over = np.where(s < 0, -s, np.where(s > L, s - L, 0.0))
        rad = np.sqrt(np.maximum(r * r - over * over, 0.0))
        pts.append(a + np.clip(s, 0, L)[:, None] * axis + np.sign(s - np.clip(s, 0, L))[:, None] * over[:, None] * axis
                   + rad[:, None] * (np.cos(th)[:, None] * u + np.sin(th)[:, None] * w))
    v_template = np.concatenate(pts, 0)
    v_template = v_template[rng.permutation(V)]

    dist = np.stack([_seg_dist(v_template, a, b) for (_, a, b, _) in segs], 1)       # [V, nseg]
    drive = np.array([j for (j, _, _, _) in segs])
    tau = 0.04 ** 2
    logits = -dist ** 2 / tau
    keep = np.argsort(-logits, axis=1)[:, :4]
    weights = np.zeros((V, J))
    for k in range(4):
        idx = keep[:, k]
        e = np.exp(logits[np.arange(V), idx] - logits[np.arange(V), keep[:, 0]])
        np.add.at(weights, (np.arange(V), drive[idx]), e)
    weights /= weights.sum(1, keepdims=True)

    # joint regressor: normalised incidence of the 48 vertices nearest to each T-pose joint
    # (joints are *defined* as J_regressor @ v_shaped, so they sit within a few cm of _TPOSE).
    J_regressor = np.zeros((J, V))
    for j in range(J):
        d = np.linalg.norm(v_template - _TPOSE[j], axis=1)
        idx = np.argsort(d)[:48]
        J_regressor[j, idx] = 1.0 / 48
    shapedirs = rng.normal(0, 0.01, (V, 3, 10))
    posedirs = rng.normal(0, 0.001, (V, 3, 207))
    f = rng.integers(0, V, (13776, 3))                       # random index triples, made pairwise distinct
    f[:, 1] = (f[:, 0] + 1 + f[:, 1] % (V - 2)) % V
    f[:, 2] = (f[:, 1] + 1 + f[:, 2] % (V - 3)) % V
    clash = f[:, 2] == f[:, 0]
    f[clash, 2] = (f[clash, 2] + 1) % V
    kintree = np.stack([PARENTS.copy(), np.arange(J)], 0)
    kintree[0, 0] = 4294967295
    return {
        'v_template': v_template.astype(np.float32),
        'shapedirs': shapedirs.astype(np.float32),
        'posedirs': posedirs.astype(np.float32),
        'J_regressor': J_regressor.astype(np.float32),
        'weights': weights.astype(np.float32),
        'kintree_table': kintree.astype(np.int64),
        'f': f.astype(np.int64),
    }


def smpl_model_to_torch(model: dict, device='cpu') -> dict:
    """Same result as renderer.py:65-74 applied to an in-memory (dense) model."""
    out = {}
    for k, v in model.items():
        t = torch.as_tensor(np.asarray(v))
        out[k] = (t.long() if k in ('kintree_table', 'f') else t.float()).to(device)
    return out


def rodrigues_np(rvec):
    """cv2.Rodrigues semantics for [n,3] axis-angle (smpl_numpy.py:60-64)."""
    rvec = np.asarray(rvec, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((len(rvec), 3, 3))
    for i, r in enumerate(rvec):
        th = np.linalg.norm(r)
        if th < 1e-12:
            out[i] = np.eye(3)
            continue
        k = r / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        out[i] = np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)
    return out


def smpl_forward_np(model: dict, poses, shapes):
    """Posed vertices in SMPL space; the dataset-side SMPL (smpl_numpy.py:46-98), float64."""
    vt = model['v_template'].astype(np.float64)
    v_shaped = vt + model['shapedirs'].astype(np.float64).reshape(-1, 10).dot(np.asarray(shapes, np.float64).reshape(10)).reshape(V, 3)
    Jt = model['J_regressor'].astype(np.float64).dot(v_shaped)
    R = rodrigues_np(np.asarray(poses).reshape(24, 3))
    lrot = (R[1:] - np.eye(3)).reshape(-1)
    v_posed = v_shaped + model['posedirs'].astype(np.float64).reshape(-1, 207).dot(lrot).reshape(V, 3)
    G = np.zeros((J, 4, 4))
    for j in range(J):
        loc = np.eye(4)
        loc[:3, :3] = R[j]
        loc[:3, 3] = Jt[j] - (Jt[PARENTS[j]] if j > 0 else 0)
        G[j] = loc if j == 0 else G[PARENTS[j]] @ loc
    for j in range(J):
        G[j, :3, 3] -= G[j, :3, :3] @ Jt[j]
    T = model['weights'].astype(np.float64).dot(G.reshape(J, 16)).reshape(V, 4, 4)
    return np.einsum('vij,vj->vi', T[:, :3, :3], v_posed) + T[:, :3, 3]


def big_pose():
    """RenderPeople_dataset.py:222-235."""
    p = np.zeros((1, 72), np.float32)
    p[0, 5] = 45 / 180 * np.pi
    p[0, 8] = -45 / 180 * np.pi
    p[0, 23] = -30 / 180 * np.pi
    p[0, 26] = 30 / 180 * np.pi
    return {'R': np.ones((3, 3), np.float32), 'Th': np.zeros((1, 3), np.float32),
            'shapes': np.zeros((1, 10), np.float32), 'poses': p}


def _look_at(eye, target, up=np.array([0.0, 1.0, 0.0])):
    """World->camera (R, T) of an OpenCV-style camera (x right, y down, z forward)."""
    z = target - eye; z /= np.linalg.norm(z)
    x = np.cross(z, up); x /= np.linalg.norm(x)          # right-handed with y down
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)
    return R, (-R @ eye).reshape(3, 1)


def get_rays_np(H, W, K, R, T):
    rays_o = -np.dot(R.T, T).ravel()
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    cam = np.dot(np.stack([i, j, np.ones_like(i)], 2), np.linalg.inv(K).T)
    rays_d = np.dot(cam - T.ravel(), R) - rays_o[None, None]
    return np.broadcast_to(rays_o, rays_d.shape), rays_d


def near_far_np(bounds, ray_o, ray_d):
    """Slab test with the dataset's exact two-hit rule; returns near, far (0/1 on misses), hit mask."""
    bounds = bounds + np.array([-0.01, 0.01])[:, None]
    ray_d = ray_d.copy()
    ray_d[ray_d == 0.0] = 1e-8
    d_int = ((bounds[None] - ray_o[:, None]) / ray_d[:, None]).reshape(-1, 6)
    p_int = d_int[..., None] * ray_d[:, None] + ray_o[:, None]
    lo, hi = bounds[0] - 1e-6, bounds[1] + 1e-6
    inside = np.all((p_int >= lo) & (p_int <= hi), axis=-1)
    hit = inside.sum(-1) == 2
    near = np.zeros(len(ray_o), np.float32)
    far = np.ones(len(ray_o), np.float32)
    pi = p_int[hit][inside[hit]].reshape(-1, 2, 3)
    nrm = np.linalg.norm(ray_d[hit], axis=1)
    d0 = np.linalg.norm(pi[:, 0] - ray_o[hit], axis=1) / nrm
    d1 = np.linalg.norm(pi[:, 1] - ray_o[hit], axis=1) / nrm
    near[hit] = np.minimum(d0, d1).astype(np.float32)
    far[hit] = np.maximum(d0, d1).astype(np.float32)
    return near, far, hit


def volume_shapes(t_vertices: np.ndarray):
    """out_sh (z,y,x) of triplane.py:174-217 and the three pyramid levels renderer.py:762-782 densifies."""
    mn = t_vertices.min(0) - 0.05
    mx = t_vertices.max(0) + 0.05
    out = np.ceil(((mx - mn)[[2, 1, 0]]).astype(np.float32) / np.float32(0.005)).astype(np.int32)
    out = (out | 31) + 1
    lv = [tuple(int(s) // (2 ** k) for s in out) for k in (1, 2, 3)]
    return [int(s) for s in out], lv, np.stack([mn, mx], 0).astype(np.float32)


@dataclass
class SceneSpec:
    H: int = 64
    W: int = 64
    samples: int = 16
    seed: int = 0
    random_global_R: bool = False      # HuMMan / ZJU style params['R'] != I
    white_back: bool = False
    cam_dist: float = 3.0
    cam_azim_deg: float = 25.0
    obs_azim_deg: float = -40.0


def make_scene(spec: SceneSpec, model: dict | None = None, device='cpu', rays_only: bool = False) -> dict:
    """Everything ImportanceRenderer.forward consumes, as torch tensors on `device`.

    Returns dict with: input_data (the reference's dict), planes, obs_input_img, obs_input_feature,
    volumes (3 dense [1,C,D,H,W]), obs_sp_input {'bounds','out_sh'}, ray_origins, ray_directions, near, far,
    rendering_options, mask_at_box.  `rays_only`: stop after the target camera's rays (same values as the full scene's; skips the
    ~600 MB of random feature tensors) and return only ray_origins / ray_directions / near / far / mask_at_box / camera.
    """
    model = model or make_smpl_model(0)
    rng = np.random.default_rng(1000 + spec.seed)
    H, W = spec.H, spec.W

    def subject(pose_sigma):
        poses = rng.normal(0, pose_sigma, (1, 72)).astype(np.float32)
        if spec.random_global_R:
            R = rodrigues_np(poses[0, :3] * 3.0)[0].astype(np.float32)
            poses[0, :3] = 0
        else:
            R = np.eye(3, dtype=np.float32)
        Th = rng.uniform(-0.1, 0.1, (1, 3)).astype(np.float32)
        return poses, R, Th

    shapes = rng.normal(0, 0.5, (1, 10)).astype(np.float32)
    p_pose, p_R, p_Th = subject(0.2)
    o_pose, o_R, o_Th = subject(0.2)
    params = {'poses': p_pose, 'shapes': shapes, 'R': p_R, 'Th': p_Th}
    obs_params = {'poses': o_pose, 'shapes': shapes.copy(), 'R': o_R, 'Th': o_Th}
    t_params = big_pose()

    vertices = (smpl_forward_np(model, p_pose, shapes) @ p_R.astype(np.float64).T + p_Th).astype(np.float32)
    obs_vertices = (smpl_forward_np(model, o_pose, shapes) @ o_R.astype(np.float64).T + o_Th).astype(np.float32)
    t_vertices = smpl_forward_np(model, t_params['poses'], t_params['shapes']).astype(np.float32)
    mn = t_vertices.min(0) - 0.05
    mx = t_vertices.max(0) + 0.05
    mn[2] -= 0.1
    mx[2] += 0.1
    t_world_bounds = np.stack([mn, mx], 0).astype(np.float32)

    def camera(azim_deg, centre):
        a = math.radians(azim_deg)
        eye = centre + spec.cam_dist * np.array([math.sin(a), 0.15, math.cos(a)])
        R, T = _look_at(eye, centre)
        K = np.array([[1.2 * W, 0, W / 2], [0, 1.2 * W, H / 2], [0, 0, 1]], np.float64)
        return K, R, T

    K, R, T = camera(spec.cam_azim_deg, vertices.mean(0).astype(np.float64))
    oK, oR, oT = camera(spec.obs_azim_deg, obs_vertices.mean(0).astype(np.float64))
    ray_o, ray_d = get_rays_np(H, W, K, R, T)
    ray_o = ray_o.reshape(-1, 3).astype(np.float32)
    ray_d = ray_d.reshape(-1, 3).astype(np.float32)
    wb = np.stack([vertices.min(0) - 0.05, vertices.max(0) + 0.05], 0)
    near, far, hit = near_far_np(wb, ray_o, ray_d)
    if rays_only:
        t_ = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(device)
        return {'ray_origins': t_(ray_o[None]), 'ray_directions': t_(ray_d[None]), 'near': t_(near[None, :, None]), 'far': t_(far[None, :, None]),
                'mask_at_box': t_(hit), 'camera': {'K': K, 'R': R, 'T': T, 'bounds': wb}, 'spec': spec}

    out_sh, lv_shapes, sp_bounds = volume_shapes(t_vertices)
    chans = (32, 64, 96)
    volumes = []
    for (D_, H_, W_), C in zip(lv_shapes, chans):
        occ = np.zeros((D_, H_, W_), bool)
        vox = 0.005 * (out_sh[0] / D_)
        idx = np.floor((t_vertices[:, [2, 1, 0]] - sp_bounds[0][[2, 1, 0]]) / vox).astype(int)
        idx = np.clip(idx, 0, np.array([D_, H_, W_]) - 1)
        occ[idx[:, 0], idx[:, 1], idx[:, 2]] = True
        for _ in range(3):                                    # 3-voxel dilation
            g = occ.copy()
            g[1:] |= occ[:-1]; g[:-1] |= occ[1:]
            g[:, 1:] |= occ[:, :-1]; g[:, :-1] |= occ[:, 1:]
            g[:, :, 1:] |= occ[:, :, :-1]; g[:, :, :-1] |= occ[:, :, 1:]
            occ = g
        vol = np.zeros((1, C, D_, H_, W_), np.float32)
        n_on = int(occ.sum())
        vol[0][:, occ] = rng.standard_normal((C, n_on), dtype=np.float32)
        volumes.append(vol)

    planes = rng.standard_normal((1, 3, 32, 256, 256), dtype=np.float32)
    obs_feat = rng.standard_normal((1, 64, H // 2, W // 2), dtype=np.float32)
    obs_img = rng.random((1, 3, H, W), dtype=np.float32)

    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(device)

    def collate(d):                    # DataLoader batch-1 collation: prepend a batch dim to every array
        return {k: tt(np.asarray(v)[None]) for k, v in d.items()}

    input_data = {
        't_params': collate(t_params), 't_vertices': tt(t_vertices[None]), 't_world_bounds': tt(t_world_bounds[None]),
        'params': collate(params), 'vertices': tt(vertices[None]),
        'ray_o_all': tt(ray_o[None, None]), 'ray_d_all': tt(ray_d[None, None]),
        'near_all': tt(near[None, None, :, None]), 'far_all': tt(far[None, None, :, None]),
        'mask_at_box_all': tt(hit[None, None]),
        'obs_params': collate(obs_params), 'obs_vertices': tt(obs_vertices[None]),
        'obs_img_all': tt(obs_img[None]),
        'obs_K_all': tt(oK.astype(np.float32)[None, None]), 'obs_R_all': tt(oR.astype(np.float32)[None, None]),
        'obs_T_all': tt(oT.astype(np.float32)[None, None]),
    }
    rendering_options = {
        'depth_resolution': spec.samples, 'depth_resolution_importance': 0, 'clamp_mode': 'relu',
        'white_back': spec.white_back, 'density_noise': 0, 'disparity_space_sampling': False,
    }
    return {
        'input_data': input_data,
        'planes': tt(planes), 'obs_input_img': tt(obs_img), 'obs_input_feature': tt(obs_feat),
        'volumes': [tt(v) for v in volumes],
        'obs_sp_input': {'bounds': tt(sp_bounds[None]), 'out_sh': out_sh, 'batch_size': 1},
        'ray_origins': input_data['ray_o_all'][:, 0], 'ray_directions': input_data['ray_d_all'][:, 0],
        'near': input_data['near_all'][:, 0], 'far': input_data['far_all'][:, 0],
        'rendering_options': rendering_options, 'mask_at_box': tt(hit), 'spec': spec,
        # the target camera and body box the rays were made from (numpy float64 / float32, host): input of sherf_b200.rays.generate_rays
        'camera': {'K': K, 'R': R, 'T': T, 'bounds': wb},
    }
