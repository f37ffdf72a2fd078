"""Run each hot-path kernel a few times on the bench tensors (for `ncu -k regex:... --set full`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from robust_e2e_gan_b200.hotpath import DEFAULT_CFG, HotPath, make_batch  # noqa: E402

dev = torch.device("cuda:0")
cfg = dict(DEFAULT_CFG)
hp = HotPath(cfg, seed=4000).to(dev)
db = make_batch(cfg, seed=4000).to(dev)
only = sys.argv[1:] 
for name, fn, nbytes, reps in bench.kernel_specs(hp, db, cfg, dev):
    if only and not any(o in name for o in only):
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    print("ran", name, flush=True)
