"""GPU unit tests of the linear-layer kernels (CUDA-core fp32, tcgen05 TF32, tcgen05 3xTF32) against an fp64 torch reference,
on every (N, K) shape the fusion / transformer / decoder stack uses plus ragged row counts."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(96, 192), (32, 96), (144, 32), (32, 48), (32, 32), (128, 71), (128, 128), (128, 199), (64, 187)]
TOL = {'fp32': 2e-6, 'tf32x3': 4e-6, 'tf32': 2e-3}          # max |err| / (|A| @ |W|.T row scale)


@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'tf32x3'])
@pytest.mark.parametrize('N,K', SHAPES)
def test_linear_matches_fp64(precision, N, K):
    from sherf_b200 import ops
    torch.manual_seed(N * 1000 + K)
    dev = torch.device('cuda:0')
    for M, act in ((1, None), (127, 'relu'), (128, 'gelu'), (1000, 'relu'), (4099, None)):
        A = torch.randn(M, K, device=dev)
        W = torch.randn(N, K, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev)
        Y = ops.linear(A, W, b, act, precision)
        ref = A.double() @ W.double().t() + b.double()
        if act == 'relu':
            ref = torch.relu(ref)
        elif act == 'gelu':
            ref = torch.nn.functional.gelu(ref)
        scale = (A.abs().double() @ W.abs().double().t()).clamp_min(1.0)
        err = float(((Y.double() - ref).abs() / scale).max())
        assert torch.isfinite(Y).all()
        assert err <= TOL[precision], f'{precision} N={N} K={K} M={M} act={act}: scaled err {err:.3e}'


@pytest.mark.parametrize('N,K', [(128, 128), (128, 71), (64, 187), (32, 48)])
def test_linear_a_operand_from_tensor_memory(N, K):
    """The fused decoder keeps the error-compensation (lo) activations in TMEM: validates tcgen05.mma with A from tensor memory."""
    from sherf_b200 import ops
    torch.manual_seed(7)
    dev = torch.device('cuda:0')
    A = torch.randn(777, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    Y = ops.linear(A, W, b, None, '_tf32x3_tmem_a')
    ref = A.double() @ W.double().t() + b.double()
    scale = (A.abs().double() @ W.abs().double().t()).clamp_min(1.0)
    err = float(((Y.double() - ref).abs() / scale).max())
    assert err <= TOL['tf32x3'], f'scaled err {err:.3e}'
