"""Device time of the decoder's per-position products: re2e_batch_nt against the generic batch-sized kernels / GEMM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robust_e2e_gan_b200 import _lib
from robust_e2e_gan_b200.linear import gemm_tf32x3
dev = torch.device("cuda:0")
L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def timeit_warm(fn, n=50):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n



def graph_time(fn, n=40):
    """Per-launch device time from a replayed CUDA graph of n back-to-back launches (no host launch cost)."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
            for _ in range(n):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


sp = lambda: _lib.stream_ptr()
for M, N, K in [(32, 1200, 320), (32, 1200, 300), (32, 320, 1200), (32, 300, 1200)]:
    X, W, O = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.zeros(M, N, device=dev)
    WT = W.t().contiguous()
    row = {"shape": (M, N, K),
           "batch_nt": graph_time(lambda: L.re2e_batch_nt(_lib.ptr(X), _lib.ptr(W), None, _lib.ptr(O), M, N, K, 0, sp())),
           "gemm": graph_time(lambda: gemm_tf32x3(X, False, WT, True, O, M, N, K))}
    if K <= 320:
        row["skinny"] = graph_time(lambda: L.re2e_skinny_nt(_lib.ptr(X), _lib.ptr(W), _lib.ptr(O), M, N, K, 0, sp()))
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)

B, D, Z = 32, 320, 300
xc, hp, cp = torch.randn(B, D, device=dev), torch.randn(B, Z, device=dev), torch.randn(B, Z, device=dev)
Wcat = torch.randn(4 * Z, D + Z, device=dev) * 0.05
WcatT = Wcat.t().contiguous()
eg, act = torch.randn(B, 4 * Z, device=dev), torch.empty(B, 4 * Z, device=dev)
c, h = torch.empty(B, Z, device=dev), torch.empty(B, Z, device=dev)
dg, dctx, dhp, dcp = torch.randn(B, 4 * Z, device=dev), torch.empty(B, D, device=dev), torch.empty(B, Z, device=dev), torch.empty(B, Z, device=dev)
print({"lstm_step_fwd": round(graph_time(lambda: L.re2e_lstm_step_fwd(
    _lib.ptr(xc), _lib.ptr(hp), _lib.ptr(cp), _lib.ptr(Wcat), _lib.ptr(eg), None, _lib.ptr(act), _lib.ptr(c), _lib.ptr(h), B, D, Z, sp())), 2),
    "lstm_step_bwd": round(graph_time(lambda: L.re2e_lstm_step_bwd(
        _lib.ptr(dg), _lib.ptr(WcatT), _lib.ptr(dctx), _lib.ptr(dhp), B, D, Z, sp())), 2),
    "pointwise_fwd": round(graph_time(lambda: L.re2e_lstm_pointwise_fwd(
        _lib.ptr(act), _lib.ptr(eg), _lib.ptr(cp), _lib.ptr(c), _lib.ptr(h), B, Z, sp())), 2),
    "pointwise_bwd": round(graph_time(lambda: L.re2e_lstm_pointwise_bwd(
        _lib.ptr(act), _lib.ptr(cp), _lib.ptr(c), _lib.ptr(h), _lib.ptr(c), _lib.ptr(dg), _lib.ptr(dcp), B, Z, sp())), 2)})
