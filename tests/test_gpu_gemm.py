"""GPU: the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) against an fp64 product; fp32-level accuracy required."""
import pytest
import torch

from helpers import rel_err
from robust_e2e_gan_b200 import _lib
from robust_e2e_gan_b200.linear import gemm_tf32x3

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,N,K", [(128, 160, 32), (300, 200, 96), (6400, 320, 320), (777, 4233, 320), (1000, 320, 4236)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_matches_fp64(M, N, K, a_mn, b_mn):
    g = torch.Generator().manual_seed(M + N + K)
    pad = lambda n: (n + 3) // 4 * 4
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ Bm.double().t() + bias.double()
    # storage with padded leading dimensions
    if a_mn:
        As = torch.zeros(K, pad(M)); As[:, :M] = A.t()
    else:
        As = torch.zeros(M, pad(K)); As[:, :K] = A
    if b_mn:
        Bs = torch.zeros(K, pad(N)); Bs[:, :N] = Bm.t()
    else:
        Bs = torch.zeros(N, pad(K)); Bs[:, :K] = Bm
    C = torch.full((M, N + 3), 7.0, device=DEV)
    n0 = _lib.launch_count()
    gemm_tf32x3(As.to(DEV), bool(a_mn), Bs.to(DEV), bool(b_mn), C, M, N, K, bias=bias.to(DEV))
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0 + 1
    assert torch.all(C[:, N:] == 7.0)                       # nothing written outside the N columns
    e = rel_err(C[:, :N], ref)
    e32 = rel_err((A.to(DEV) @ Bm.to(DEV).t() + bias.to(DEV)), ref)
    assert e < 1e-5, (e, e32)                               # 3xTF32: ~2^-20 relative to sum |a||b|
    # accumulate
    gemm_tf32x3(As.to(DEV), bool(a_mn), Bs.to(DEV), bool(b_mn), C, M, N, K, accumulate=True)
    assert rel_err(C[:, :N], 2 * ref - bias.double()) < 2e-5


@pytest.mark.parametrize("rows,cols,pad", [(6400, 4233, 3), (1, 5, 0), (257, 320, 0), (130, 129, 7)])
def test_colsum_matches_fp64_and_is_deterministic(rows, cols, pad):
    from robust_e2e_gan_b200.linear import colsum
    g = torch.Generator().manual_seed(rows + cols)
    X = torch.randn(rows, cols + pad, generator=g).to(DEV)
    out = colsum(X[:, :cols])
    assert rel_err(out, X[:, :cols].double().sum(0)) < 1e-5
    assert torch.equal(out, colsum(X[:, :cols]))
