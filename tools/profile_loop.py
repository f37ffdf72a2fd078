"""Runs the two persistent decoder-loop kernels (and nothing else of ours) a few times at the bench shape: the target
of `ncu --set full --import-source on -k regex:attloc_loop` captures (profiles/README.md)."""
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
    if spec[0].startswith("attloc_loop"):
        for _ in range(3):
            spec[1]()
        torch.cuda.synchronize()
print("done")
