"""Where the end-to-end step time goes: copies only / replays only / both (StepRunner internals)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robust_e2e_gan_b200.hotpath import DEFAULT_CFG, HotPath, make_batch, StepRunner, Batch
dev = torch.device("cuda:0")
cfg = dict(DEFAULT_CFG)
hp = HotPath(cfg, seed=4000).to(dev)
hb = make_batch(cfg, seed=4000).pin()
r = StepRunner(hp, hb, slots=3)
N = 30
def timeit(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): fn()
    torch.cuda.synchronize(); print("%-40s %.3f ms/step" % (name, (time.perf_counter() - t0) / N * 1e3), flush=True)
k = [0]
def copies():
    r._stage(k[0] % 3, hb); k[0] += 1
timeit(copies, "copies only (_stage)")
def copies_nohost():
    s = k[0] % 3; k[0] += 1
    db = r.slots[s]["batch"]
    with torch.cuda.stream(r.copy_stream):
        for f in Batch.FIELDS: getattr(db, f).copy_(getattr(hb, f), non_blocking=True)
timeit(copies_nohost, "copies only (no label staging)")
def big3():
    s = k[0] % 3; k[0] += 1
    db = r.slots[s]["batch"]
    with torch.cuda.stream(r.copy_stream):
        for f in ("mix", "clean", "mask_logits"): getattr(db, f).copy_(getattr(hb, f), non_blocking=True)
timeit(big3, "3 big copies only (79 MB)")
timeit(lambda: r.replay_resident(0), "replays only")
def both():
    copies_nohost(); r.replay_resident(0)
timeit(both, "copies || replay (no dependency)")
t0 = time.perf_counter()
for _ in range(N): r._stage(0, hb)
print("host time of _stage: %.3f ms" % ((time.perf_counter() - t0) / N * 1e3)); torch.cuda.synchronize()
for f in Batch.FIELDS: print(f, tuple(getattr(hb, f).shape), getattr(hb, f).numel() * 4 / 1e6, "MB", getattr(hb, f).is_pinned())
