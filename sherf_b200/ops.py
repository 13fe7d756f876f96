"""Functional access to the auxiliary C-ABI entry points: the per-pose SMPL transforms (`sherf_lbs_transforms`), the global depth
range of a view (`sherf_depth_range`, what ray shards need) and the single-layer diagnostic (`sherf_debug_linear`)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_ACT = {None: 0, 'none': 0, 'relu': 1, 'gelu': 2}


def linear(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor | None = None, act: str | None = None, precision: str = 'fp32'):
    """Y = act(A @ W.T + bias) through libsherf_b200.so (`sherf_debug_linear`); A[M,K], W[N,K] fp32 CUDA tensors."""
    from .renderer import PRECISIONS
    lib = _lib.load()
    assert A.is_cuda and W.is_cuda and A.dtype == torch.float32 and W.dtype == torch.float32
    A, W = A.contiguous(), W.contiguous()
    M, K = A.shape
    N = W.shape[0]
    assert K % 4 == 0 or True
    lda = (K + 3) // 4 * 4
    if lda != K:
        Ap = A.new_zeros(M, lda)
        Ap[:, :K] = A
        A = Ap
    Y = torch.empty(M, N, device=A.device, dtype=torch.float32)
    scratch = torch.empty(8 * 272 * 272 * 4 + 1024, dtype=torch.uint8, device=A.device)
    b = bias.contiguous() if bias is not None else None
    rc = lib.sherf_debug_linear(PRECISIONS[precision], A.data_ptr(), lda, W.data_ptr(), b.data_ptr() if b is not None else None,
                                Y.data_ptr(), N, M, N, K, _ACT[act], scratch.data_ptr(), scratch.numel(),
                                torch.cuda.current_stream(A.device).cuda_stream)
    _lib.check(rc)
    return Y


def lbs_transforms(renderer, params: dict) -> torch.Tensor:
    """The 24 rigid joint transforms A[24,4,4] of one pose through `sherf_lbs_transforms` (replaces get_transform_params_torch,
    renderer.py:129-157).  `renderer` supplies the SMPL model (ImportanceRenderer.set_smpl_model / assets/SMPL_NEUTRAL.pkl);
    `params` is an `input_data['params']`-style dict of CUDA tensors (poses [..,72], shapes [..,10], R, Th)."""
    lib = _lib.load()
    device = params['poses'].device
    if device.type != 'cuda':
        raise RuntimeError('sherf_b200.ops.lbs_transforms runs on CUDA tensors only (no CPU fallback)')
    with torch.cuda.device(device):
        keep = []
        smpl = renderer._smpl_struct(device)
        pose = renderer._pose_struct(params, device, keep)
        A = torch.empty(24, 4, 4, device=device, dtype=torch.float32)
        scratch = torch.empty(4096, dtype=torch.uint8, device=device)
        _lib.check(lib.sherf_lbs_transforms(C.byref(smpl), C.byref(pose), A.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                            torch.cuda.current_stream(device).cuda_stream))
        del keep
    return A


def smpl_vertices(renderer, params: dict, world: bool = True) -> torch.Tensor:
    """Posed SMPL vertices [1,V,3] of one frame through `sherf_smpl_vertices` (replaces the host-side sherf/smpl/smpl_numpy.py:46-98 and,
    with `world`, the `xyz @ R.T + Th` of RenderPeople_dataset.py:210, i.e. `input_data['vertices']`).  `params`: CUDA tensors poses [..,72],
    shapes [..,10], R [..,3,3], Th [..,3]."""
    lib = _lib.load()
    device = params['poses'].device
    if device.type != 'cuda':
        raise RuntimeError('sherf_b200.ops.smpl_vertices runs on CUDA tensors only (no CPU fallback)')
    with torch.cuda.device(device):
        keep = []
        smpl = renderer._smpl_struct(device)
        pose = renderer._pose_struct(params, device, keep)
        out = torch.empty(1, smpl.n_verts, 3, device=device, dtype=torch.float32)
        scratch = torch.empty(8192, dtype=torch.uint8, device=device)
        _lib.check(lib.sherf_smpl_vertices(C.byref(smpl), C.byref(pose), None if world else out.data_ptr(), out.data_ptr() if world else None,
                                           scratch.data_ptr(), scratch.numel(), torch.cuda.current_stream(device).cuda_stream))
        del keep
    return out


def depth_range(near: torch.Tensor, far: torch.Tensor, n_samples: int):
    """(min, max) of all stratified sample depths of a view through `sherf_depth_range` (torch.min / torch.max(depths),
    ray_marcher.py:57); near / far: CUDA tensors with one value per ray."""
    lib = _lib.load()
    device = near.device
    if device.type != 'cuda':
        raise RuntimeError('sherf_b200.ops.depth_range runs on CUDA tensors only (no CPU fallback)')
    nr, fa = near.detach().float().contiguous(), far.detach().float().contiguous()
    rays = _lib.SherfRays()
    rays.near_, rays.far_, rays.n_rays, rays.n_samples = nr.data_ptr(), fa.data_ptr(), nr.numel(), int(n_samples)
    lo, hi = C.c_float(0), C.c_float(0)
    with torch.cuda.device(device):
        scratch = torch.empty(4096, dtype=torch.uint8, device=device)
        _lib.check(lib.sherf_depth_range(C.byref(rays), C.byref(lo), C.byref(hi), scratch.data_ptr(), scratch.numel(),
                                         torch.cuda.current_stream(device).cuda_stream))
    return float(lo.value), float(hi.value)
