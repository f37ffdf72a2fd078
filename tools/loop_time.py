"""CUDA-event time of the two persistent decoder-loop kernels at the bench shape (release library, graph of 4
launches, median of 7 replays) -- the quick A/B number while working on csrc/attloc_loop.cu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from robust_e2e_gan_b200.hotpath import DEFAULT_CFG, HotPath, make_batch
dev = torch.device("cuda:0")
cfg = dict(DEFAULT_CFG)
hp = HotPath(cfg, seed=4000).to(dev)
db = make_batch(cfg, seed=4000).to(dev)
for spec in bench.kernel_specs(hp, db, cfg, dev):
    name, fn = spec[0], spec[1]
    if not name.startswith("attloc_loop"):
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            fn()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 4)
    ts.sort()
    print("%s: %.1f us per launch, %.2f us per step" % (name, ts[3], ts[3] / cfg["steps"]), flush=True)
