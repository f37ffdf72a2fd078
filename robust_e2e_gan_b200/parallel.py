"""Utterance-sharded data parallelism for the hot path.

The reference is single-process / single-GPU (only ``gpu_ids[0]`` is used, joint_train.py:55).
Every op on the hot path is independent across utterances, so the step shards by utterance with
ONE exchange: the parameter gradients.  One process per GPU; gradients live as views into a few
flat buckets and each bucket is all-reduced (NCCL over NVLink 5 / NVSwitch; gloo in the CPU tests)
as soon as its last gradient has been accumulated, overlapping the rest of the backward.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / MASTER_* from the env). Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def bind_host_to_gpu(local_rank):
    """Pin this process to the CPUs that are closest to GPU ``local_rank`` (same socket / NUMA node as its PCIe root).

    One process per GPU: with eight ranks on a two-socket host, pinned staging buffers that are first touched on the
    far socket make every H2D copy cross the inter-socket link.  Call this BEFORE allocating pinned memory.  Returns
    the number of CPUs bound to, or 0 when the topology cannot be read (restricted container, no NVML) -- the process
    then simply keeps its inherited affinity.  ``RE2E_NUMA_BIND=0`` disables it.
    """
    if os.environ.get("RE2E_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return 0
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(int(local_rank))
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = ideal & allowed
        if not cpus or cpus == allowed:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of ``n_items`` utterances for ``rank`` (remainder to the low ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradBuckets(object):
    """Flat gradient buckets with asynchronous all-reduce overlapped with the backward pass.

    ``params`` keep being ordinary nn.Parameters (the reference's optimisers and
    ``clip_grad_norm_`` see them unchanged, joint_train.py:127-140,188); their ``.grad`` tensors are
    views into the buckets.  Usage per step::

        buckets.zero()            # instead of optimizer.zero_grad()
        loss.backward()           # hooks launch all-reduces bucket by bucket
        buckets.finish()          # wait + average; .grad now holds the mean over ranks

    Ordering rules (the same ones DDP keeps):
      * buckets are reduced in FIXED index order on every rank -- a bucket whose gradients are complete is held back
        until every lower-index bucket has been launched, and ``finish()`` launches whatever is left (parameters the
        loss did not touch) in index order, so two ranks with different sets of unused parameters still issue the same
        sequence of collectives;
      * on CUDA the collectives run on ONE dedicated communication stream.  Each parameter's hook records an event on
        the stream its gradient was accumulated on (the three-stream hot path accumulates gradients of one bucket on
        different streams); the communication stream waits for every event of a bucket before reducing it, and
        ``finish()`` makes the caller's stream wait for the communication stream.
    """

    def __init__(self, params, bucket_mb=25.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        cap = int(bucket_mb * 1024 * 1024)
        # reverse order: gradients become ready roughly last-parameter-first
        order = list(reversed(self.params))
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in order:
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > cap or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self.flat, self._bucket_of, self._pending = [], {}, []
        self._hooks = []
        self._comm = {}                   # device -> communication stream
        for bi, bucket in enumerate(self.buckets):
            total = sum(p.numel() for p in bucket)
            flat = torch.zeros(total, dtype=bucket[0].dtype, device=bucket[0].device)
            o = 0
            for p in bucket:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
                self._bucket_of[p] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(p)))
            self.flat.append(flat)
            self._pending.append(len(bucket))
        self._reset_step()

    def _reset_step(self):
        self._next = 0                                        # next bucket index to launch (fixed order)
        self._events = [[] for _ in self.buckets]             # CUDA events of the accumulations, per bucket
        self._handles = []
        for bi, bucket in enumerate(self.buckets):
            self._pending[bi] = len(bucket)

    def _comm_stream(self, dev):
        if dev not in self._comm:
            self._comm[dev] = torch.cuda.Stream(dev)
        return self._comm[dev]

    def _make_hook(self, p):
        def hook(param):
            bi = self._bucket_of[p]
            if param.is_cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(param.device))     # the stream this gradient was accumulated on
                self._events[bi].append(ev)
            self._pending[bi] -= 1
            self._launch_ready()
        return hook

    def _launch_ready(self):
        while self._next < len(self.buckets) and self._pending[self._next] <= 0:
            self._launch(self._next)
            self._next += 1

    def _launch(self, bi):
        flat = self.flat[bi]
        if self.world > 1:
            if flat.is_cuda:
                cs = self._comm_stream(flat.device)
                for ev in self._events[bi]:
                    cs.wait_event(ev)
                # a bucket launched from finish() (unused parameters): order it after the caller's stream as well
                cs.wait_stream(torch.cuda.current_stream(flat.device))
                with torch.cuda.stream(cs):
                    self._handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                self._handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero(self):
        for flat in self.flat:
            flat.zero_()
        self._reset_step()
        # re-attach views (an optimizer.zero_grad(set_to_none=True) would have dropped them)
        for bi, bucket in enumerate(self.buckets):
            o = 0
            for p in bucket:
                if p.grad is None or p.grad.data_ptr() != self.flat[bi][o:o + 1].data_ptr():
                    p.grad = self.flat[bi][o:o + p.numel()].view_as(p)
                o += p.numel()

    def finish(self):
        """Reduce (in index order) every bucket not launched yet -- buckets held back by the fixed order, or whose
        hooks did not all fire (unused parameters) -- then wait and average."""
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        if self.world > 1:
            cuda_devs = {flat.device for flat in self.flat if flat.is_cuda}
            for dev in cuda_devs:
                # wait + average ON the communication stream (Work.wait() orders the stream that is current when it is
                # called after the collective -- the division must not run ahead of it), then order the caller after it
                cs = self._comm_stream(dev)
                with torch.cuda.stream(cs):
                    for h in self._handles:
                        h.wait()
                    for flat in self.flat:
                        if flat.device == dev:
                            flat.div_(self.world)
                torch.cuda.current_stream(dev).wait_stream(cs)
            if not cuda_devs:
                for h in self._handles:
                    h.wait()
                for flat in self.flat:
                    flat.div_(self.world)
        self._handles = []

    def nbytes(self):
        return sum(f.numel() * f.element_size() for f in self.flat)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
