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
    with _lib.on(C.device):
        _lib.check(L.re2e_gemm_tf32x3(_lib.ptr(A), lda, int(a_mn), _lib.ptr(B), ldb, int(b_mn), _lib.ptr(C), ldc,
                                      _lib.ptr(bias), M, N, K, int(accumulate), _lib.stream_ptr()), "re2e_gemm_tf32x3")
    return C


def colsum(g):
    """Column sums of a row-major 2-D view (unit inner stride, any row pitch): db of a dense layer, deterministic."""
    L = _lib.lib()
    rows, cols = g.shape
    assert g.stride(1) == 1
    ld = g.stride(0) if rows > 1 else max(g.stride(0), cols)
    nrb = int(L.re2e_colsum_blocks(rows))
    partial = torch.empty(nrb, cols, device=g.device, dtype=torch.float32)
    out = torch.empty(cols, device=g.device, dtype=torch.float32)
    with _lib.on(g.device):
        _lib.check(L.re2e_colsum(_lib.ptr(g), ld, rows, cols, _lib.ptr(partial), _lib.ptr(out), _lib.stream_ptr()),
                   "re2e_colsum")
    return out


def _pad4(n):
    return (n + 3) // 4 * 4


def _padded_view(out2, lead, N):
    """(M, ld) row-padded 2-D buffer -> (*lead, N) view with the padded pitch (no copy)."""
    ld = out2.stride(0)
    strides, acc = [], ld
    for n in reversed(lead):
        strides.append(acc)
        acc *= n
    return out2.as_strided(tuple(lead) + (N,), tuple(reversed(strides)) + (1,))


def _as_rows(g, M, N):
    """View an incoming gradient (*lead, N) as TMA-addressable rows (M, N) with a uniform pitch; copies only
    when the layout really is not row-regular (never on the hot path: the CTC backward writes its gradient in
    the padded layout of the logits)."""
    if g.dim() >= 2 and g.stride(-1) == 1:
        ld = g.stride(-2) if g.shape[-2] > 1 else max(g.stride(-2), N)
        ok = ld % 4 == 0 and g.data_ptr() % 16 == 0
        acc = ld * g.shape[-2]
        for d in range(g.dim() - 3, -1, -1):            # outer dims must continue the same row pitch
            ok = ok and (g.shape[d] == 1 or g.stride(d) == acc)
            acc *= g.shape[d]
        if ok:
            return g.as_strided((M, N), (ld, 1))
    gp = torch.empty(M, _pad4(N), device=g.device, dtype=torch.float32)[:, :N]
    gp.copy_(g.reshape(M, N))
    return gp


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        lead = tuple(x.shape[:-1])
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = _lib.f32c(x.detach()).reshape(-1, K)
        M = x2.shape[0]
        w = _lib.f32c(weight.detach())
        out = torch.empty(M, _pad4(N), device=x2.device, dtype=torch.float32)[:, :N]
        gemm_tf32x3(x2, False, w, False, out, M, N, K, bias=_lib.f32c(bias.detach()) if bias is not None else None)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.lead = lead
        # the (possibly row-padded) N-D view is created HERE: an as_strided outside the Function would make autograd
        # materialise a zero-filled copy of the whole gradient (108 MB for the CTC logits) in its backward
        return out if (len(lead) == 1 and out.is_contiguous()) else _padded_view(out, lead, N)

    @staticmethod
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        g = _as_rows(g, M, N)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=g.device, dtype=torch.float32)
            gemm_tf32x3(g, False, w, True, dx, M, K, N)               # dX = g W      (B = W stored [N][K]: MN-major)
            dx = dx.view(*ctx.lead, K)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(N, K, device=g.device, dtype=torch.float32)
            gemm_tf32x3(g, True, x2, True, dw, N, K, M)               # dW = g^T X    (both MN-major)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(g)
        return dx, dw, db


def linear(x, weight, bias=None):
    """F.linear on the tcgen05 path; x (..., K) fp32 CUDA, weight (N, K), K % 4 == 0.  The result may be a view with
    a row pitch padded to a multiple of 4 floats (TMA-addressable rows); values are exactly (..., N)."""
    return _LinearTC.apply(x, weight, bias)
