"""Timeline of the tcgen05 front-end kernel (needs a library built with -DRE2E_FB_DEBUG)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from robust_e2e_gan_b200 import _lib
_lib.LIB_PATH = os.environ.get("RE2E_DEBUG_LIB", os.path.join(ROOT, "robust_e2e_gan_b200", "libre2e_b200_fbdbg.so"))
from robust_e2e_gan_b200.hotpath import DEFAULT_CFG, HotPath, make_batch
dev = torch.device("cuda:0")
cfg = dict(DEFAULT_CFG)
hp = HotPath(cfg, seed=4000).to(dev)
db = make_batch(cfg, seed=4000).to(dev)
L = _lib.lib()
buf = (ctypes.c_longlong * (16 * 256))()
for name, fn, nbytes, reps in bench.kernel_specs(hp, db, cfg, dev):
    if not name.startswith("fbank_fwd"):
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    L.re2e_fb_debug_read.argtypes = [ctypes.c_void_p]
    print(name, "rc", L.re2e_fb_debug_read(buf))
    t0 = min(buf[i * 16] for i in range(148))
    print(" cta  start  setup | cvt_end cvt_wait | epi_end epi_wait | mma_end mma_wait | total   (ns)")
    for i in list(range(0, 148, 21)) + [147]:
        r = [buf[i * 16 + k] for k in range(14)]
        print(" %3d %6d %6d | %7d %8d | %7d %8d | %7d %8d | %6d" % (i, r[0] - t0, r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]))
    print(" converter warp 0, SM clocks summed over its items:  wait-empty   convert+STS   fence.proxy   syncwarp+arrive   issue next loads")
    for i in list(range(0, 148, 21)) + [147]:
        r = [buf[i * 16 + k] for k in range(9, 14)]
        print(" %3d %48d %13d %13d %17d %18d" % (i, r[0], r[1], r[2], r[3], r[4]))
