"""Host-side mirror of the decoder half of the reference's `training.triplane` that sits on the render hot path.

`NeRFDecoder` (/root/reference/sherf/training/triplane.py:267-316) is a pure parameter container here: same
constructor, same parameter names (`pts_linears.{0..7}`, `views_linear`, `feature_linear`, `alpha_linear`,
`rgb_linear`) so `copy_params_and_buffers(require_all=True)` resumes by name.  Its arithmetic runs inside
libsherf_b200.so (ImportanceRenderer.forward passes the module as the `decoder` argument, exactly like
triplane.py:156-157); calling it directly raises, because there is deliberately no PyTorch path.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class NeRFDecoder(nn.Module):
    def __init__(self, n_features):
        super().__init__()
        W = 128
        self.with_viewdirs = True
        self.skips = [4]
        nerf_input_ch = n_features + 39
        self.pts_linears = nn.ModuleList(
            [nn.Linear(nerf_input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + nerf_input_ch, W) for i in range(7)])
        nerf_input_ch_2 = n_features + W
        self.views_linear = nn.Linear(nerf_input_ch_2 + 27, W // 2)
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        if n_features != 32:
            raise NotImplementedError('the CUDA decoder is built for n_features=32 (triplane.py:64)')

    def forward(self, ray_points, sampled_features, ray_directions):
        raise RuntimeError('sherf_b200.NeRFDecoder is evaluated inside the fused CUDA renderer; '
                           'pass it as the `decoder` argument of ImportanceRenderer.forward')


def hot_path_modules(smpl_model: dict | None = None, seed: int = 0, mlp_precision: str = 'bf16x3', dense_sigma: bool = False):
    """(ImportanceRenderer, NeRFDecoder) in the configuration every shipped SHERF script uses (train.py:310-318,
    *.sh: use_trans / use_nerf_decoder True), default PyTorch init under `seed`.  `dense_sigma` rescales the density
    head so that a randomly initialised network renders an opaque body (used by parity tests and the bench so that
    the composite is exercised with non-trivial weights)."""
    from .renderer import ImportanceRenderer
    torch.manual_seed(seed)
    ren = ImportanceRenderer(use_1d_feature=True, use_2d_feature=True, use_3d_feature=True, use_trans=True,
                             use_NeRF_decoder=True, smpl_model=smpl_model, mlp_precision=mlp_precision)
    dec = NeRFDecoder(32)
    if dense_sigma:
        with torch.no_grad():
            dec.alpha_linear.weight *= 30
            dec.alpha_linear.bias += 2.0
    return ren.eval().requires_grad_(False), dec.eval().requires_grad_(False)
