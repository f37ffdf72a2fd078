"""CPU: host-side logic -- target preparation, filter banks, sharding, gradient buckets over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from robust_e2e_gan_b200 import synth
from robust_e2e_gan_b200.e2e_ctc import prepare_targets
from robust_e2e_gan_b200.feat_model import generic_mel_banks, kaldi_mel_banks
from robust_e2e_gan_b200.parallel import GradBuckets, shard_range


def test_prepare_targets_padded_and_list_agree():
    ys = [torch.tensor([3, 3, 5]), torch.tensor([], dtype=torch.long), torch.tensor([7, 2, 9, 4])]
    pad = torch.full((3, 5), -1, dtype=torch.long)
    for i, y in enumerate(ys):
        pad[i, :len(y)] = y
    a = prepare_targets(ys, "cpu")
    b = prepare_targets(pad, "cpu")
    for t in (a, b):
        assert t.lens.tolist() == [3, 0, 4] and t.offs.tolist() == [0, 3, 3] and t.umax == 4 and t.nutt == 3
        assert t.labels.tolist() == [3, 3, 5, 7, 2, 9, 4] and t.labels.dtype == torch.int32


def test_mel_banks_shapes_and_support():
    k = kaldi_mel_banks(80)
    assert k.shape == (80, 257) and (k >= 0).all() and (k > 0).sum() == 501
    assert k[:, 0].sum() == 0 and k[:, 256].sum() == 0          # bins 0 and 256 unused
    g = generic_mel_banks(40)
    assert g.shape == (40, 257) and ((g > 0).sum(0) <= 2).all()
    assert synth.mel_fc(257, 40).shape == (257, 40)


def test_default_80_filter_bank_is_bit_identical_to_the_reference_table():
    """A freshly constructed FbankModel(fbank_dim=80) carries the reference's constants (model/feat_model.py:15-33),
    not a regenerated bank: bit-equal to the ``fc`` of the unmodified reference module held in the golden file."""
    import types
    from conftest import golden
    from robust_e2e_gan_b200.feat_model import FbankModel, reference_fbank80
    args = types.SimpleNamespace(idim=257, fbank_dim=80, enhance_type="blstm", fbank_opti_type="frozen",
                                 train_dataset_len=10, num_utt_cmvn=5)
    m = FbankModel(args)
    ref_fc = golden("fbank")["fc"]
    assert m.fc.dtype == torch.float32 and tuple(m.fc.shape) == (257, 80) and not m.fc.requires_grad
    assert np.array_equal(m.fc.detach().numpy(), ref_fc)
    assert np.array_equal(reference_fbank80().T.astype(np.float32), ref_fc)
    args.enhance_type = "unet_256"                                  # model/feat_model.py:102-104: bank cut to 256 bins
    assert np.array_equal(FbankModel(args).fc.detach().numpy(), ref_fc[:256])


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 32, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synth_targets_are_ctc_feasible():
    h, hl = synth.encoder_batch(B=6, Th=20, D=8, seed=3)
    ys = synth.targets(B=6, V=30, hlens=hl, umin=8, umax=24, seed=3)
    for y, l in zip(ys, hl):
        rep = int((y[1:] == y[:-1]).sum()) if len(y) > 1 else 0
        assert len(y) + rep <= l


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical init on every rank
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    unused = torch.nn.Parameter(torch.ones(5))             # never touched by the loss
    params = list(model.parameters()) + [unused]
    buckets = GradBuckets(params, bucket_mb=0.001)         # tiny buckets -> several all-reduces
    g = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
    lo, hi = shard_range(8, rank, world)
    out = []
    for _ in range(2):                                     # two steps: views survive zero()
        buckets.zero()
        loss = ((model(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean()
        loss.backward()
        buckets.finish()
        out.append([p.grad.clone() for p in params])
    # single-process truth: mean over the equal-sized shards == grad of the mean over all rows
    model.zero_grad()
    for p in params:
        p.grad = None
    ((model(x_all) - y_all) ** 2).mean().backward()
    ok = all(torch.allclose(a, p.grad, atol=1e-6) for a, p in zip(out[1][:-1], params[:-1]))
    ok = ok and torch.equal(out[1][-1], torch.zeros(5)) and len(buckets.buckets) > 1
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_buckets_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bind_host_to_gpu_is_a_no_op_without_topology(monkeypatch):
    """No NVML / no GPU (this container), or RE2E_NUMA_BIND=0: the process keeps its affinity and 0 is returned."""
    import os
    from robust_e2e_gan_b200.parallel import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    assert bind_host_to_gpu(0) >= 0
    monkeypatch.setenv("RE2E_NUMA_BIND", "0")
    assert bind_host_to_gpu(0) == 0
    if not __import__("torch").cuda.is_available():
        assert os.sched_getaffinity(0) == before


def _cmvn_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robust_e2e_gan_b200.feat_model import global_mean_var
    g = torch.Generator().manual_seed(5)
    y = torch.randn(40, 6, generator=g).double() * 3 + 1.5          # 40 frames of 6 mel bins, the whole set
    lo, hi = (0, 13) if rank == 0 else (13, 40)                      # unequal shards
    mine = y[lo:hi]
    mean, var = global_mean_var(torch.stack([mine.sum(0), (mine * mine).sum(0)]), hi - lo)
    ok = np.allclose(mean, y.mean(0).numpy()) and np.allclose(var, y.var(0, unbiased=False).numpy())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_cmvn_statistics_are_reduced_over_ranks_world2_gloo():
    """compute_cmvn's final step under data parallelism: per-rank (sum, sumsq, count) all-reduced, every
    rank gets the statistics of the WHOLE set."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cmvn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_beam_bookkeeping_matches_the_reference_formulation():
    """Decoder._host_merge (linked hypotheses, token lists only for finished ones) against the reference's bookkeeping
    written out literally (model/e2e_decoder.py:296-333: copy yseq per candidate, append <eos> at the last position, move
    <eos> hypotheses longer than minlen to the ended list with the length penalty) on random candidate streams."""
    from robust_e2e_gan_b200.e2e_decoder import Decoder
    dec = Decoder.__new__(Decoder)          # bookkeeping only: no parameters needed
    eos, sos, beam = 7, 7, 4
    dec.eos, dec.sos = eos, sos
    rng = np.random.RandomState(0)
    for maxlen, minlen, penalty in ((9, 0, 0.0), (6, 3, 0.25), (12, 5, -0.1)):
        hyps = [(np.float32(0.0), sos, None, 1)]
        ref_hyps = [{'score': np.float32(0.0), 'yseq': [sos]}]
        ended, ref_ended = [], []
        for i in range(maxlen):
            n = len(hyps)
            entries = [(np.float32(rng.randn()), int(rng.randint(n)), int(rng.randint(5, 9)), int(rng.randint(6)))
                       for _ in range(beam)]
            hyps = dec._host_merge(hyps, ended, entries, i, maxlen, minlen, penalty)
            new = [{'score': np.float32(s), 'yseq': ref_hyps[r]['yseq'] + [t]} for s, r, t, _ in entries]
            if i == maxlen - 1:
                for h in new:
                    h['yseq'].append(eos)
            ref_rem = []
            for h in new:
                if h['yseq'][-1] == eos:
                    if len(h['yseq']) > minlen:
                        h['score'] = np.float32(h['score'] + np.float32((i + 1) * penalty))
                        ref_ended.append(h)
                else:
                    ref_rem.append(h)
            ref_hyps = ref_rem
            assert len(hyps) == len(ref_hyps)
            for a, b in zip(hyps, ref_hyps):
                seq, node = [], a
                while node is not None:
                    seq.append(node[1])
                    node = node[2]
                assert seq[::-1] == b['yseq'] and a[0] == b['score'] and a[3] == len(b['yseq'])
            assert ended == ref_ended
            if not hyps:
                break
