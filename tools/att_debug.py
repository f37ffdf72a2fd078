"""Phase timeline of the AttLoc kernels (needs a library built with EXTRA=-DRE2E_ATT_DEBUG)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from robust_e2e_gan_b200 import _lib
_lib.LIB_PATH = os.environ.get("RE2E_DEBUG_LIB", os.path.join(ROOT, "robust_e2e_gan_b200", "libre2e_b200_dbg.so"))
from robust_e2e_gan_b200.hotpath import DEFAULT_CFG, HotPath, make_batch
dev = torch.device("cuda:0")
cfg = dict(DEFAULT_CFG)
hp = HotPath(cfg, seed=4000).to(dev)
db = make_batch(cfg, seed=4000).to(dev)
L = _lib.lib()
L.re2e_att_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_longlong * (16 * 512))()
names = {0: ["start", "sync1", "decproj", "sync2", "sync3", "sync4(main)", "pushed", "end", "m_wait", "m_energy",
             "m_pairbar", "m_ctx"],
         1: ["start", "sync1", "sync2(p1)", "sync3(de)", "sync4(p2)", "postA", "sync5", "sync6", "fin", "bulkwait",
             "A_wzissued", "A_pushed", "A_pushed(T)", "A_dWconv(T)", "A_xbar2"]}
for spec in bench.kernel_specs(hp, db, cfg, dev):
    name, fn = spec[0], spec[1]
    if not name.startswith("attloc_step"):
        continue
    which = 0 if "fwd" in name else 1
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(8):
            fn()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    L.re2e_att_debug_read(buf, which)
    n = len(names[which])
    t0 = min(buf[i * 16] for i in range(128))
    print(name, "(last launch of a graph of 8; ns relative to the first CTA's start)")
    print("  cta " + " ".join("%11s" % x for x in names[which]))
    for i in (0, 1, 2, 3, 64, 65, 126, 127):
        print("  %3d " % i + " ".join("%11d" % (buf[i * 16 + k] - t0) for k in range(n)))
    import statistics
    print("  med " + " ".join("%11d" % statistics.median(buf[i * 16 + k] - buf[i * 16] for i in range(128)) for k in range(n)), "(per-CTA elapsed)")
