"""Per-CTA timeline of the tcgen05 3xTF32 GEMM (needs `make -C robust_e2e_gan_b200/csrc gemmdebug`)."""
import ctypes, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from robust_e2e_gan_b200 import _lib
_lib.LIB_PATH = os.environ.get("RE2E_DEBUG_LIB", os.path.join(ROOT, "robust_e2e_gan_b200", "libre2e_b200_gemmdbg.so"))
from robust_e2e_gan_b200.hotpath import DEFAULT_CFG, HotPath, make_batch
dev = torch.device("cuda:0")
cfg = dict(DEFAULT_CFG)
hp = HotPath(cfg, seed=4000).to(dev)
db = make_batch(cfg, seed=4000).to(dev)
L = _lib.lib()
L.re2e_gemm_debug_read.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_longlong * (8 * 2048))()
for name, fn, nbytes, reps in bench.kernel_specs(hp, db, cfg, dev):
    if not name.startswith("gemm"):
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    L.re2e_gemm_debug_read(buf)        # discard the warm-up records
    fn()
    torch.cuda.synchronize()
    L.re2e_gemm_debug_read(buf)
    rows = [[buf[i * 8 + k] for k in range(5)] for i in range(2048) if buf[i * 8] > 0]
    t0 = min(r[0] for r in rows)
    ends = sorted(r[4] - t0 for r in rows)
    med = lambda k: statistics.median(r[k] - r[0] for r in rows)
    print("%s: %d CTAs recorded; kernel span %.1f us" % (name, len(rows), ends[-1] / 1e3))
    print("   per CTA (median ns): setup %d | first stage converted %d | accumulator complete %d | epilogue done %d"
          % (med(1), med(2), med(3), med(4)))
