"""TEST INFRASTRUCTURE -- not product code.  Functional stand-ins for the three spconv classes and the tensor record the reference's
sparse 3-D encoder is written against (renderer.py:26 `import spconv.pytorch as spconv`; SparseConvNet renderer.py:707-797, its block
builders :799-871; triplane.py:137 builds the `spconv.core.SparseConvTensor`).

Why: spconv-cu113==2.3.3 (requirement.txt:24) is not vendored and cannot be installed here, so the reference's `SparseConvNet.forward`
could only be constructed, never run.  With these stand-ins the reference's OWN forward runs on the CPU -- its layer order, its four
`.dense()` taps, its `grid_sample` calls and the concatenation are then the reference's code, and only the arithmetic of one sparse
convolution is ours.  oracle/sparse_encoder.py (the checker of csrc/sparse_encoder.cu) is compared against that in
tests/test_sparse_encoder.py.  PARITY STATUS stays "unpinned" for the per-layer arithmetic: the rules below are restated from spconv's
documentation and source as published for the 2.x line, not checked against a running spconv.

spconv 2.3.3 rules restated (file paths inside the spconv repository, v2.3.3):
  * `spconv/pytorch/core.py: SparseConvTensor(features, indices, spatial_shape, batch_size)`: features [n, C]; indices [n, 4] int32 =
    (batch, z, y, x); `.dense()` = `scatter_nd(indices, features, [batch, *spatial_shape, C])` permuted to [batch, C, *spatial_shape];
    `.replace_feature(f)` returns a tensor sharing indices / shape / the indice_dict.
  * `spconv/pytorch/modules.py: SparseSequential.forward`: a sparse module receives the SparseConvTensor; any other module (BatchNorm1d,
    ReLU) is applied to `.features` and the result put back with `replace_feature`.
  * `spconv/pytorch/conv.py: SparseConvolution`: weight layout KRSC = [out, kz, ky, kx, in] for every algorithm since 2.2
    (`docs/USAGE.md`, "all weights are KRSC"); no bias here (renderer.py:820,838,856,869 pass bias=False).
      - `SubMConv3d` (subm=True): output indices = input indices; out[p] = sum over kernel offsets d of W[:, d] . in[p + d - 1] for the
        offsets whose neighbour is an ACTIVE input site (`ops.get_indice_pairs`, subm branch).
      - `SparseConv3d(k=3, stride=2, padding=1)`: output spatial shape floor((D + 2 pad - k) / stride) + 1 per axis
        (`ops.get_conv_output_size`); output site o is active iff at least one active input p and offset k satisfy p = stride o - pad + k;
        out[o] = sum of W[:, k] . in[stride o - pad + k] over those.
  * Duplicate indices (two SMPL vertices in one 5 mm voxel, triplane.py:193): spconv neither merges nor rejects them; its hash table keeps
    one row per coordinate and which one is implementation-defined (GPU insertion order).  CONVENTION here and in csrc/sparse_encoder.cu:
    the row with the smallest index represents the voxel.

The arithmetic below is the dense formulation (torch.nn.functional.conv3d on the densified volume times an activity mask): independent of
the gather form in oracle/sparse_encoder.encode_sparse.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _first_rows(indices: torch.Tensor):
    """Row numbers of the first occurrence of every distinct (batch, z, y, x), in input order (the duplicate convention)."""
    seen, keep = set(), []
    for r, c in enumerate(map(tuple, indices.tolist())):
        if c not in seen:
            seen.add(c)
            keep.append(r)
    return torch.tensor(keep, dtype=torch.long, device=indices.device)


class SparseConvTensor:
    """spconv.core.SparseConvTensor stand-in (batch_size 1: the reference renders one subject per call, triplane.py:213)."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None, benchmark=False):
        assert int(batch_size) == 1, 'the stand-in covers batch_size == 1'
        self.features = features
        self.indices = indices.to(torch.int64) if torch.is_tensor(indices) else torch.as_tensor(indices, dtype=torch.int64)
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size)

    def _scatter(self, values):
        """values [n, C] -> [1, C, D, H, W], first row of a voxel wins."""
        keep = _first_rows(self.indices)
        idx = self.indices[keep]
        vol = torch.zeros(1, values.shape[1], *self.spatial_shape, dtype=values.dtype, device=values.device)
        vol[0, :, idx[:, 1], idx[:, 2], idx[:, 3]] = values[keep].t()
        return vol

    def dense(self, channels_first=True):
        vol = self._scatter(self.features)
        return vol if channels_first else vol.permute(0, 2, 3, 4, 1).contiguous()

    def activity(self):
        return self._scatter(torch.ones(self.features.shape[0], 1, dtype=self.features.dtype, device=self.features.device))


class SparseConvolution(nn.Module):
    """Common part of the SubMConv3d / SparseConv3d stand-ins: constructor signature of spconv.pytorch.conv.SparseConvolution's 3-D
    subclasses, `weight` in the KRSC layout (zero-filled without touching torch's RNG stream, like the constructor-only stub it replaces:
    state-dict names and shapes of the reference module can be compared; tests load their own values)."""
    subm = False

    def __init__(self, in_channels=None, out_channels=None, kernel_size=3, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, *a, **k):
        super().__init__()
        ks = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        self.kernel_size, self.stride, self.padding = ks, (stride if isinstance(stride, int) else stride[0]), \
            (padding if isinstance(padding, int) else padding[0])
        assert dilation == 1 and groups == 1 and not bias, 'the reference uses dilation 1, groups 1, bias=False only'
        if in_channels is not None and out_channels is not None:
            self.weight = nn.Parameter(torch.zeros(out_channels, ks, ks, ks, in_channels))

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        w = self.weight.permute(0, 4, 1, 2, 3).contiguous()                      # KRSC -> [out, in, kz, ky, kx]
        vol, act = x._scatter(x.features), x.activity()           # (not .dense(): that name stays the reference's own call)
        if self.subm:
            # output sites = input sites (duplicates included: each row reads the value at its own coordinate)
            out = F.conv3d(vol, w, padding=self.kernel_size // 2)
            idx = x.indices
            return SparseConvTensor(out[0, :, idx[:, 1], idx[:, 2], idx[:, 3]].t().contiguous(), idx, x.spatial_shape, x.batch_size)
        out = F.conv3d(vol, w, stride=self.stride, padding=self.padding)
        oact = F.conv3d(act, torch.ones(1, 1, *([self.kernel_size] * 3), dtype=act.dtype), stride=self.stride, padding=self.padding) > 0
        oz, oy, ox = torch.nonzero(oact[0, 0], as_tuple=True)                      # row order of the outputs is not observable (.dense())
        idx = torch.stack([torch.zeros_like(oz), oz, oy, ox], 1)
        return SparseConvTensor(out[0, :, oz, oy, ox].t().contiguous(), idx, list(out.shape[2:]), x.batch_size)


class SubMConv3d(SparseConvolution):
    subm = True


class SparseConv3d(SparseConvolution):
    subm = False


class SparseSequential(nn.Sequential):
    def forward(self, x):
        if not isinstance(x, SparseConvTensor):
            return super().forward(x)
        for m in self:
            x = m(x) if isinstance(m, SparseConvolution) else x.replace_feature(m(x.features))
        return x
