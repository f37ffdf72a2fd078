"""Run each hot-path kernel once inside a cudaProfilerStart/Stop range on the bench tensors, for
    ncu --set full --profile-from-start off --import-source on -o REP python tools/profile_kernels.py [name substrings]
(two untimed warm-up launches of every kernel happen outside the range)."""
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
specs = [s for s in bench.kernel_specs(hp, db, cfg, dev) if not only or any(o in s[0] for o in only)]
for spec in specs:
    for _ in range(2):
        spec[1]()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for spec in specs:
    spec[1]()
    torch.cuda.synchronize()
    print("ran", spec[0], flush=True)
torch.cuda.cudart().cudaProfilerStop()
