"""GPU: the composed hot-path step (what bench.py times and smoke() runs) against the oracle."""
import pytest
import torch

from helpers import assert_close
from oracle.hotpath_oracle import oracle_step
from robust_e2e_gan_b200 import _lib
from robust_e2e_gan_b200.hotpath import HotPath, make_batch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_config1_step_matches_oracle():
    """BASELINE config 1: B=8, T=400, 257 bins, 40 mel, vocab 4233, Th=100."""
    cfg = dict(B=8, T=400, F=257, M=40, Th=100, D=320, A=320, Z=300, C=10, filts=100, V=4233, U=16, steps=9)
    batch = make_batch(cfg, seed=1234)
    hp = HotPath(cfg, seed=1234).to(DEV)
    n0 = _lib.launch_count()
    out = hp.step(batch.to(DEV))
    torch.cuda.synchronize()
    n_launch = _lib.launch_count() - n0
    assert 10 <= n_launch < 2 * cfg["steps"] + 12      # joint front-end + CTC + ONE loop kernel per direction + dense layers
    ref = oracle_step(cfg, batch, hp.state_dict_cpu())
    ref64 = oracle_step(cfg, batch, hp.state_dict_cpu(), dtype=torch.float64)
    for k in ref:
        if k == "d_att.gvec.bias":
            continue
        got = torch.stack(out[k]) if isinstance(out[k], (list, tuple)) else out[k]
        assert_close(got, ref[k], truth=ref64[k], what=k)


def test_step_runner_matches_eager_and_oracle():
    """StepRunner (CUDA-graph replay over static buffers, pipelined H2D) returns what the eager step returns,
    also for a second, different host batch fed through the same captured graphs."""
    from robust_e2e_gan_b200.hotpath import StepRunner
    cfg = dict(B=4, T=64, F=257, M=40, Th=16, D=320, A=320, Z=300, C=10, filts=100, V=97, U=5, steps=4)
    hp = HotPath(cfg, seed=11).to(DEV)
    b0 = make_batch(cfg, seed=11).pin()
    b1 = make_batch(cfg, seed=12).pin()
    runner = StepRunner(hp, b0, slots=2)
    runner.submit(b0)
    runner.submit(b1)
    outs = []
    for _ in range(2):
        o = runner.result()
        outs.append({k: (torch.stack(v) if isinstance(v, (list, tuple)) else v).detach().float().cpu().clone()
                     for k, v in o.items()})
    for hb, got in zip((b0, b1), outs):
        ref = oracle_step(cfg, hb, hp.state_dict_cpu())
        ref64 = oracle_step(cfg, hb, hp.state_dict_cpu(), dtype=torch.float64)
        for k in ref:
            if k == "d_att.gvec.bias":
                continue
            assert_close(got[k], ref[k], truth=ref64[k], what="runner " + k)
    # replaying again on the same inputs is bit-identical except for the atomically accumulated sums
    again = runner(b1)
    assert torch.equal(again["enhance_feat"].cpu(), outs[1]["enhance_feat"])
    assert torch.equal(again["att_c"].cpu(), outs[1]["att_c"])


def test_bench_shape_step_through_graph_replay_matches_oracle():
    """The EXACT thing bench.py times: DEFAULT_CFG (B=32, T=800, Th=200, U=40, 41 decoder steps, V=4233), three
    branch streams, CUDA-graph replay through StepRunner from a pinned host batch -- every output and gradient
    against the fp32 oracle with the fp64 oracle as tie-breaker, tolerance 1e-4 (max-norm and element-wise)."""
    from robust_e2e_gan_b200.hotpath import DEFAULT_CFG, StepRunner
    cfg = dict(DEFAULT_CFG)
    hp = HotPath(cfg, seed=4000).to(DEV)
    hb = make_batch(cfg, seed=4001).pin()
    runner = StepRunner(hp, make_batch(cfg, seed=4000).pin(), slots=2)
    n0 = _lib.launch_count()
    out = runner(hb)                                  # a batch the graphs were NOT captured on
    got = {k: (torch.stack(v) if isinstance(v, (list, tuple)) else v).detach().float().cpu().clone()
           for k, v in out.items()}
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0                  # replay only: no eager launches of ours
    sd = hp.state_dict_cpu()
    ref = oracle_step(cfg, hb, sd)
    ref64 = oracle_step(cfg, hb, sd, dtype=torch.float64)
    worst = 0.0
    for k in ref:
        if k == "d_att.gvec.bias":                    # analytically zero (softmax is shift invariant)
            continue
        worst = max(worst, assert_close(got[k], ref[k], truth=ref64[k], what="bench-shape " + k))
    print("bench-shape step: worst error metric vs oracle %.2e" % worst)
    runner.close()


def test_step_runner_three_slots_two_batches_ahead():
    """The bench's end-to-end loop: three input slots, two host batches in flight ahead of the compute.  Results come
    back in submission order and match the single-step results of the same batches."""
    from robust_e2e_gan_b200.hotpath import StepRunner
    cfg = dict(B=4, T=64, F=257, M=40, Th=16, D=320, A=320, Z=300, C=10, filts=100, V=97, U=5, steps=4)
    hp = HotPath(cfg, seed=21).to(DEV)
    batches = [make_batch(cfg, seed=30 + i).pin() for i in range(3)]
    runner = StepRunner(hp, batches[0], slots=3)
    single = []
    for hb in batches:
        o = runner(hb)
        single.append((float(o["loss_ctc"].detach().cpu()), o["att_c"].detach().cpu().clone()))
    order = [0, 1, 2, 2, 0, 1, 1, 0]
    runner.submit(batches[order[0]])
    runner.submit(batches[order[1]])
    got = []
    for i in range(len(order)):
        if i + 2 < len(order):
            runner.submit(batches[order[i + 2]])
        o = runner.result()
        got.append((float(o["loss_ctc"].detach().cpu()), o["att_c"].detach().cpu().clone()))
    for k, (loss, c) in zip(order, got):
        assert loss == single[k][0]
        assert torch.equal(c, single[k][1])
    with pytest.raises(RuntimeError):          # a fourth outstanding submit would overwrite an unread slot
        for _ in range(4):
            runner.submit(batches[0])


def test_after_backward_hook_runs_inside_the_captured_step_and_close_drops_the_graphs():
    """`HotPath.after_backward` is where a data-parallel job joins its gradient exchange: it must run once per step,
    after the gradients exist, and whatever it enqueues must be part of the captured graph (it is replayed with it)."""
    from robust_e2e_gan_b200.hotpath import StepRunner
    cfg = dict(B=4, T=64, F=257, M=40, Th=16, D=320, A=320, Z=300, C=10, filts=100, V=97, U=5, steps=4)
    hp = HotPath(cfg, seed=22).to(DEV)
    hb = make_batch(cfg, seed=40).pin()
    calls = []

    def joined(out):
        calls.append(1)
        out["_gsum"] = sum(out[k].sum() for k in out if k.startswith("d_ctc.") or k.startswith("d_att."))

    hp.after_backward = joined
    ref = hp.step(hb.to(DEV), hlens_for_att=hb.to(DEV).hlens)
    assert len(calls) == 1 and torch.isfinite(ref["_gsum"])
    g_ref = float(ref["_gsum"].cpu())
    del ref        # an eager graph kept alive would pin the parameters' grad accumulators to the (legacy) stream it ran on
    runner = StepRunner(hp, hb, slots=2)
    n_capture = len(calls)
    o1 = runner(hb)
    v1 = float(o1["_gsum"].cpu())
    o2 = runner(make_batch(cfg, seed=41).pin())
    v2 = float(o2["_gsum"].cpu())
    assert len(calls) == n_capture                      # replays do not call back into Python ...
    assert abs(v1 - g_ref) <= 1e-3 * max(1.0, abs(v1))
    assert v1 != v2                                     # ... but the hook's kernels ran again on the new batch
    runner.close()
    assert all(s["graph"] is None and s["out"] is None for s in runner.slots)


def test_collate_writes_pinned_batches():
    from robust_e2e_gan_b200 import kaldi_feats as kf
    g = torch.Generator().manual_seed(1)
    samples = [("u%d" % i, "s", *[torch.rand(5 + i, 7, generator=g) for _ in range(5)], torch.tensor([1, 2])) for i in range(3)]
    out = kf.collate(samples)
    assert all(out[i].is_pinned() for i in range(2, 7)) and out[8].is_pinned()
    assert out[0] == ["u2", "u1", "u0"] and out[8].tolist() == [7, 6, 5]
    assert torch.equal(out[4][0], samples[2][4]) and float(out[4][2, 5:].abs().sum()) == 0.0
