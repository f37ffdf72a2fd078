"""Where a step of the persistent decoder-loop kernels goes: per-phase SM cycles of thread 0, summed over all steps
(needs the instrumented library: make -C robust_e2e_gan_b200/csrc debug).  Prints microseconds PER STEP."""
import ctypes, os, statistics, sys
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
L.re2e_loop_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_longlong * (16 * 512))()
names = {0: ["prologue", "conv", "conv_red", "sweep(e+ctx)", "bar3", "pushes", "xwait", "softmax+out"],
         1: ["prologue", "loads", "pass1", "smx_exch", "pass2", "tma+bar4", "pushes", "param_grads", "xwait2",
             "datt_part", "datt_red"]}
mhz = 1965.0   # SM clock under load on this pool (clock64 counts SM cycles)
for spec in bench.kernel_specs(hp, db, cfg, dev):
    name, fn = spec[0], spec[1]
    if not name.startswith("attloc_loop"):
        continue
    which = 0 if "fwd" in name else 1
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    L.re2e_loop_debug_read(buf, which)
    n = len(names[which])
    steps = cfg["steps"]
    ncta = 128
    print("%s: %.1f us per launch (instrumented), %.2f us per step; phase times in us PER STEP at %.0f MHz"
          % (name, e0.elapsed_time(e1) * 1e3, e0.elapsed_time(e1) * 1e3 / steps, mhz))
    print("  cta  " + " ".join("%12s" % x for x in names[which]) + "   total")
    for i in (0, 1, 2, 3, 64, 127):
        row = [buf[i * 16 + k] / mhz / (1 if k == 0 else steps) for k in range(n)]
        print("  %3d  " % i + " ".join("%12.2f" % v for v in row) + "   %.2f" % sum(row[1:]))
    med = [statistics.median(buf[i * 16 + k] for i in range(ncta)) / mhz / (1 if k == 0 else steps) for k in range(n)]
    print("  med  " + " ".join("%12.2f" % v for v in med) + "   %.2f  (prologue in us per launch)" % sum(med[1:]))
