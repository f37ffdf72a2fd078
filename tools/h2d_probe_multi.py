#!/usr/bin/env python
"""Host-link ceiling of the end-to-end step at N ranks: every rank copies the step's host-resident inputs (mix, clean:
2 x 26.3 MB pinned fp32) to ITS GPU with cudaMemcpyAsync in a loop, nothing else running.  Aggregate GB/s over the ranks =
what the box's host side can feed; ms per 52.7 MB = the floor of the e2e step time at that rank count.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe_multi.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from robust_e2e_gan_b200.parallel import bind_host_to_gpu, init_distributed  # noqa: E402


def main():
    rank, world = init_distributed()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bound = bind_host_to_gpu(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = 32 * 800 * 257
    host = [torch.randn(n).pin_memory() for _ in range(2)]
    devb = [torch.empty(n, device=dev) for _ in range(2)]
    nbytes = 2 * n * 4
    s = torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s):
            for h, d in zip(host, devb):
                d.copy_(h, non_blocking=True)

    for _ in range(5):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) * 1e3
        print(json.dumps({"ranks": world, "bytes_per_rank_per_step": nbytes, "ms_per_step_max_over_ranks": round(ms, 3),
                          "GBps_per_rank": round(nbytes / ms / 1e6, 1), "GBps_aggregate": round(world * nbytes / ms / 1e6, 1),
                          "host_cpus": os.cpu_count(), "cpus_bound_per_rank": bound}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
