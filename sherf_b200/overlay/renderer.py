"""Resolved as `training.volumetric_rendering.renderer` by sherf_b200.overlay: the import surface `triplane.py:19` uses
(`ImportanceRenderer, read_pickle, SMPL_to_tensor`) plus the module-level class names a pickled reference renderer refers to
(`PositionalEncoding`, `Transformer`, `SparseConvNet`, ...), all backed by sherf_b200.renderer (hand-written sm_100a CUDA behind the
C ABI; no torch arithmetic, no fallback)."""
from sherf_b200.renderer import (ImportanceRenderer, PositionalEncoding, SMPL_to_tensor, SparseConvNet, SparseConvTensor,  # noqa: F401
                                 Transformer, read_pickle)
from sherf_b200.renderer import _Attention as Attention, _FeedForward as FeedForward, _PreNorm as PreNorm, _Fn as Residual  # noqa: F401

__all__ = ['ImportanceRenderer', 'read_pickle', 'SMPL_to_tensor', 'PositionalEncoding', 'Transformer', 'SparseConvNet', 'SparseConvTensor',
           'Attention', 'FeedForward', 'PreNorm', 'Residual']
