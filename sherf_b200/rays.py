"""Dataset-side ray setup on the GPU: the CUDA counterpart of `get_rays` / `get_near_far` and the near/far fill of
`sample_ray_RenderPeople_batch` (/root/reference/sherf/training/RenderPeople_dataset.py:14-27, 68-101, 129-134).

    rays = generate_rays(H, W, K, R, T, bounds, device)   # -> dict like the dataset's ray entries of `input_data`

SURVEY.md 8(f) rank 3: a streamed novel-pose / novel-view sequence then uploads only the camera and the pose per frame
instead of 3 x N floats of host-generated rays.  No CPU fallback: the arithmetic runs in libsherf_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _dbl(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    assert a.size == n, (a.size, n)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def generate_rays(H: int, W: int, K, R, T, bounds, device) -> dict:
    """K [3,3], R [3,3], T [3] or [3,1] (world -> camera), bounds [2,3] (min / max corner of the body box, before the
    dataset's own 1 cm padding).  Returns device tensors shaped like the collated dataset entries:
    ray_o_all / ray_d_all [1,1,N,3], near_all / far_all [1,1,N,1], mask_at_box_all [1,1,N] (bool)."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('sherf_b200.rays.generate_rays runs on CUDA only (no CPU fallback)')
    lib = _lib.load()
    N = H * W
    with torch.cuda.device(device):
        o = torch.empty(1, 1, N, 3, device=device)
        d = torch.empty(1, 1, N, 3, device=device)
        nr = torch.empty(1, 1, N, 1, device=device)
        fr = torch.empty(1, 1, N, 1, device=device)
        m = torch.empty(1, 1, N, dtype=torch.uint8, device=device)
        (Ka, Kp), (Ra, Rp), (Ta, Tp), (Ba, Bp) = _dbl(K, 9), _dbl(R, 9), _dbl(T, 3), _dbl(bounds, 6)
        _lib.check(lib.sherf_generate_rays(Kp, Rp, Tp, H, W, Bp, o.data_ptr(), d.data_ptr(), nr.data_ptr(), fr.data_ptr(), m.data_ptr(),
                                           torch.cuda.current_stream(device).cuda_stream))
    return {'ray_o_all': o, 'ray_d_all': d, 'near_all': nr, 'far_all': fr, 'mask_at_box_all': m.bool()}
