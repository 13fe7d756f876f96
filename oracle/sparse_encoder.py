"""TEST INFRASTRUCTURE -- not product code.  CPU restatement of the reference's sparse 3-D encoder:
SparseConvNet (renderer.py:708-797: conv0, down0, conv1, down1, conv2, down2, conv3 with num_layers = 4) built from
double_conv / triple_conv / stride_conv (renderer.py:814-871) on the SparseConvTensor triplane.py:137 assembles from
prepare_sp_input (triplane.py:174-217), and the three `.dense()` volumes it samples (renderer.py:762,771,780).

PARITY STATUS: "parity unpinned".  The arithmetic lives in spconv-cu113==2.3.3 (requirement.txt:24), which is not vendored and
cannot be installed here, and the reference has no tests or fixtures for it.  This file states the published semantics of the
three spconv ops the reference uses and is pinned two ways (tests/test_sparse_encoder.py): `encode_sparse` (gather form, what
the CUDA kernels mirror) == `encode_dense` (the same network written with torch.nn.functional.conv3d on densified volumes +
activity masks), and both follow the reference's call sites / layer list line by line.

Semantics stated here (all bias-free, renderer.py:820,838,856,869):
  * SparseConvTensor(features[n,C], indices[n,4] = (batch, z, y, x), spatial_shape): one feature row per index.  The reference can
    hand it DUPLICATE indices (two SMPL vertices rounding to one 5 mm voxel, triplane.py:193); spconv does not merge them and its
    hash table keeps one of them -- which one is unspecified.  OUR CONVENTION: the vertex with the smallest index represents the
    voxel, the other rows are dropped.
  * SubMConv3d(k=3): output sites = input sites; out[p] = sum_{d in {-1,0,1}^3} W[:, d] . in[p + d] over ACTIVE neighbours.
  * SparseConv3d(k=3, stride=2, padding=1): output shape floor((D + 2 - 3) / 2) + 1 per axis; output site o is active iff some
    active input p satisfies p = 2 o - 1 + k, k in {0,1,2}^3; out[o] = sum_k W[:, k] . in[2 o - 1 + k] over active inputs.
  * weights: [out, kz, ky, kx, in] (spconv 2.x "KRSC" layout of SubMConv3d.weight / SparseConv3d.weight).
  * BatchNorm1d(eps=1e-3) over the active rows, evaluation mode (running statistics), then ReLU.
  * .dense(): zeros [1, C, D, H, W] with the active sites filled.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# (name, kind, c_in, c_out, number of conv+BN+ReLU triples) in execution order; a dense level is emitted after conv1 / conv2 / conv3
LAYERS = [('conv0', 'subm', 32, 32, 2), ('down0', 'down', 32, 32, 1), ('conv1', 'subm', 32, 32, 2), ('down1', 'down', 32, 64, 1),
          ('conv2', 'subm', 64, 64, 3), ('down2', 'down', 64, 96, 1), ('conv3', 'subm', 96, 96, 3)]
EMIT_AFTER = ('conv1', 'conv2', 'conv3')


def conv_list():
    """[(state-dict prefix of the conv, prefix of its BatchNorm, kind, c_in, c_out)] in execution order (13 convs)."""
    out = []
    for name, kind, cin, cout, n in LAYERS:
        for i in range(n):
            out.append((f'{name}.{3 * i}', f'{name}.{3 * i + 1}', kind, cin if i == 0 else cout, cout))
    return out


def bn_relu(x, sd, bn):
    scale = sd[bn + '.weight'] / torch.sqrt(sd[bn + '.running_var'] + 1e-3)
    return F.relu(x * scale + (sd[bn + '.bias'] - sd[bn + '.running_mean'] * scale))


def unique_voxels(coord, feat):
    """Our duplicate convention: first (smallest-index) vertex of a voxel wins.  coord [n,3] (z,y,x) int -> coords[m,3], feats[m,C]."""
    seen, keep = {}, []
    for i, c in enumerate(map(tuple, coord.tolist())):
        if c not in seen:
            seen[c] = i
            keep.append(i)
    keep = torch.tensor(keep, dtype=torch.long)
    return coord[keep].long(), feat[keep]


@torch.no_grad()
def encode_sparse(sd: dict, coord, feat, out_sh):
    """Gather-form evaluation.  sd: 'encoder_3d'-relative state dict; coord [n,3] int (z,y,x); feat [n,32]; out_sh (D,H,W).
    Returns the three dense levels [1,32,D/2,H/2,W/2], [1,64,D/4,..], [1,96,D/8,..] (renderer.py:762,771,780)."""
    coords, x = unique_voxels(coord, feat)
    shape = [int(s) for s in out_sh]
    dense = []
    convs = iter(conv_list())
    for name, kind, cin, cout, n in LAYERS:
        for _ in range(n):
            cprefix, bprefix, kind_, ci, co = next(convs)
            W = sd[cprefix + '.weight']                                             # [co, 3, 3, 3, ci]
            index = {tuple(c): r for r, c in enumerate(coords.tolist())}
            if kind_ == 'subm':
                out = torch.zeros(coords.shape[0], co)
                for r, (z, y, xx) in enumerate(coords.tolist()):
                    for kz in range(3):
                        for ky in range(3):
                            for kx in range(3):
                                j = index.get((z + kz - 1, y + ky - 1, xx + kx - 1))
                                if j is not None:
                                    out[r] += W[:, kz, ky, kx] @ x[j]
            else:
                oshape = [(s + 2 - 3) // 2 + 1 for s in shape]
                ocoords = {}
                for (z, y, xx) in coords.tolist():
                    for kz in range(3):
                        for ky in range(3):
                            for kx in range(3):
                                oz, oy, ox = z + 1 - kz, y + 1 - ky, xx + 1 - kx
                                if oz % 2 or oy % 2 or ox % 2:
                                    continue
                                o = (oz // 2, oy // 2, ox // 2)
                                if all(0 <= o[a] < oshape[a] for a in range(3)):
                                    ocoords.setdefault(o, len(ocoords))
                oc = torch.tensor(sorted(ocoords, key=ocoords.get), dtype=torch.long).reshape(-1, 3)
                out = torch.zeros(oc.shape[0], co)
                for r, (z, y, xx) in enumerate(oc.tolist()):
                    for kz in range(3):
                        for ky in range(3):
                            for kx in range(3):
                                j = index.get((2 * z - 1 + kz, 2 * y - 1 + ky, 2 * xx - 1 + kx))
                                if j is not None:
                                    out[r] += W[:, kz, ky, kx] @ x[j]
                coords, shape = oc, oshape
            x = bn_relu(out, sd, bprefix)
        if name in EMIT_AFTER:
            vol = torch.zeros(1, x.shape[1], *shape)
            vol[0, :, coords[:, 0], coords[:, 1], coords[:, 2]] = x.t()
            dense.append(vol)
    return dense


@torch.no_grad()
def encode_dense(sd: dict, coord, feat, out_sh):
    """The same network on densified volumes with activity masks (an independent formulation used to pin encode_sparse)."""
    coords, x = unique_voxels(coord, feat)
    shape = [int(s) for s in out_sh]
    vol = torch.zeros(1, x.shape[1], *shape)
    act = torch.zeros(1, 1, *shape)
    vol[0, :, coords[:, 0], coords[:, 1], coords[:, 2]] = x.t()
    act[0, 0, coords[:, 0], coords[:, 1], coords[:, 2]] = 1
    dense = []
    convs = iter(conv_list())
    for name, kind, cin, cout, n in LAYERS:
        for _ in range(n):
            cprefix, bprefix, kind_, ci, co = next(convs)
            W = sd[cprefix + '.weight'].permute(0, 4, 1, 2, 3).contiguous()          # [co, ci, kz, ky, kx]
            if kind_ == 'subm':
                vol = F.conv3d(vol, W, padding=1)
            else:
                vol = F.conv3d(vol, W, stride=2, padding=1)
                act = (F.conv3d(act, torch.ones(1, 1, 3, 3, 3), stride=2, padding=1) > 0).float()
            scale = sd[bprefix + '.weight'] / torch.sqrt(sd[bprefix + '.running_var'] + 1e-3)
            shift = sd[bprefix + '.bias'] - sd[bprefix + '.running_mean'] * scale
            vol = F.relu(vol * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)) * act
        if name in EMIT_AFTER:
            dense.append(vol.clone())
    return dense


def random_state_dict(encoder_module, seed=0):
    """The module's state dict with non-trivial BatchNorm statistics (the default running_mean 0 / running_var 1 would hide errors)."""
    g = torch.Generator().manual_seed(seed)
    sd = {k: v.clone().float() for k, v in encoder_module.state_dict().items()}
    for k in sd:
        if k.endswith('running_mean'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.5
        elif k.endswith('.bias'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith('.weight') and sd[k].dim() == 1:
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
        elif k.endswith('.weight'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * (2.0 / (27 * sd[k].shape[-1])) ** 0.5 * 2.0
    return sd
