"""Dense layers of the hot path on the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu).

``linear(x, weight, bias)`` is a drop-in for ``torch.nn.functional.linear`` on 2-D/3-D fp32 CUDA
inputs with fp32-level accuracy (the north star's 1e-4 parity rules out single-pass TF32/BF16):
forward  Y = X W^T + b,  backward  dX = dY W  and  dW = dY^T X  all run on the tensor cores.
The output (and dY) may have a padded row pitch ``ld_out`` (multiple of 4 floats) so that matrices
whose logical width is odd (V = 4233) stay TMA-addressable.
"""
import torch

from . import _lib


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "need a row-major 2-D view"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm_tf32x3(A, a_mn, B, b_mn, C, M, N, K, bias=None, accumulate=False):
    """C[M,N] (+)= Aop[M,K] Bop[N,K]^T (+bias).  A: [M][lda] (a_mn False) or [K][lda] (a_mn True); same for B."""
    L = _lib.lib()
    lda, ldb, ldc = _ld(A), _ld(B), _ld(C)
    with torch.cuda.device(C.device):
        _lib.check(L.re2e_gemm_tf32x3(_lib.ptr(A), lda, int(a_mn), _lib.ptr(B), ldb, int(b_mn), _lib.ptr(C), ldc,
                                      _lib.ptr(bias), M, N, K, int(accumulate), _lib.stream_ptr()), "re2e_gemm_tf32x3")
    return C


def _pad4(n):
    return (n + 3) // 4 * 4


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, weight, bias):
        M, K = x2.shape
        N = weight.shape[0]
        x2 = _lib.f32c(x2)
        w = _lib.f32c(weight.detach())
        out = torch.empty(M, _pad4(N), device=x2.device, dtype=torch.float32)[:, :N]
        gemm_tf32x3(x2, False, w, False, out, M, N, K, bias=_lib.f32c(bias.detach()) if bias is not None else None)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        if g.stride(1) != 1 or g.stride(0) % 4 != 0 or g.data_ptr() % 16 != 0:
            gp = torch.empty(M, _pad4(N), device=g.device, dtype=torch.float32)[:, :N]
            gp.copy_(g)
            g = gp
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=g.device, dtype=torch.float32)
            gemm_tf32x3(g, False, w, True, dx, M, K, N)               # dX = g W      (B = W stored [N][K]: MN-major)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(N, K, device=g.device, dtype=torch.float32)
            gemm_tf32x3(g, True, x2, True, dw, N, K, M)               # dW = g^T X    (both MN-major)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = g.sum(0)
        return dx, dw, db


def linear(x, weight, bias=None):
    """F.linear on the tcgen05 path; x (..., K) fp32 CUDA, weight (N, K), K % 4 == 0."""
    K = x.shape[-1]
    N = weight.shape[0]
    y = _LinearTC.apply(x.reshape(-1, K), weight, bias)
    lead = tuple(x.shape[:-1])
    if y.is_contiguous():
        return y.view(*lead, N)
    # row-padded output: expose (..., N) with the padded pitch, no copy
    ld = y.stride(0)
    strides, acc = [], ld
    for n in reversed(lead):
        strides.append(acc)
        acc *= n
    return y.as_strided(lead + (N,), tuple(reversed(strides)) + (1,))
