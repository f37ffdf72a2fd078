"""GPU parity: CTC loss / gradient / log-softmax / best path / prefix scores vs oracle and goldens."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from helpers import assert_close, rel_err
from oracle import ctc as o_ctc
from robust_e2e_gan_b200 import CTC, CTCPrefixScore, ctc_loss, ctc_prefix_score_batch, log_softmax_rows, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_golden_module_loss_and_grads():
    g = golden("ctc")
    B, Th, D = g["hs"].shape
    V = g["W"].shape[0]
    ctc = CTC(V, D, 0.0).to(DEV)
    ctc.ctc_lo.weight.data.copy_(T(g["W"]))
    ctc.ctc_lo.bias.data.copy_(T(g["b"]))
    hs = T(g["hs"]).to(DEV).requires_grad_(True)
    loss = ctc(hs, list(g["hlens"]), T(g["ys_pad"]).to(DEV))
    assert loss.shape == (1,)
    (0.5 * loss).sum().backward()
    assert_close(loss, g["loss"], what="loss")
    assert_close(hs.grad, g["d_hs"], what="d hs")
    assert_close(ctc.ctc_lo.weight.grad, g["d_W"], what="d W")
    assert_close(ctc.ctc_lo.bias.grad, g["d_b"], what="d b")
    assert_close(ctc.log_softmax(T(g["hs"]).to(DEV)), g["log_softmax"], what="log_softmax")
    assert np.array_equal(ctc.best_path(T(g["hs"]).to(DEV)).cpu().numpy(), g["best"])
    # list-of-tensors targets (what E2E.forward passes) give the same loss
    ys = [T(r[r != -1]).to(DEV) for r in g["ys_pad"]]
    assert_close(ctc(T(g["hs"]).to(DEV), list(g["hlens"]), ys), g["loss"], what="loss (list targets)")


def test_golden_logit_gradient():
    g = golden("ctc")
    x = T(g["logits"]).to(DEV).requires_grad_(True)
    ys = [T(r[r != -1]) for r in g["ys_pad"]]
    loss, nll = ctc_loss(x, list(g["hlens"]), ys)
    loss.sum().backward()
    assert_close(loss, g["loss_logits"], what="loss")
    assert_close(x.grad, g["d_logits"], what="d logits")
    for b, Tb in enumerate(g["hlens"]):
        assert torch.all(x.grad[b, int(Tb):] == 0)


@pytest.mark.parametrize("B,Th,V,umin,umax,seed", [(8, 100, 4233, 8, 24, 1234), (5, 33, 50, 0, 12, 3),
                                                    (3, 40, 7, 10, 19, 4), (2, 300, 129, 100, 140, 5),
                                                    # lattice widths of every register-column instantiation
                                                    # (states per lane 3, 4, 6, 8) and of the CTA-wide fallback
                                                    (4, 120, 31, 33, 47, 6), (3, 150, 40, 50, 63, 7),
                                                    (3, 220, 50, 70, 95, 8), (2, 280, 60, 100, 127, 9)])
def test_oracle_parity(B, Th, V, umin, umax, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Th, V, generator=g) * 3
    hl = sorted([int(v) for v in torch.randint(max(1, Th // 2), Th + 1, (B,), generator=g)], reverse=True)
    ys = synth.targets(B=B, V=max(V, 3), hlens=hl, umin=umin, umax=umax, seed=seed)
    res = {}
    for dt in (torch.float32, torch.float64):
        xx = x.detach().clone().to(dt).requires_grad_(True)
        loss, nll = o_ctc.ctc_loss_torch(xx, hl, ys)
        loss.sum().backward()
        res[dt] = (loss.detach(), nll.detach(), xx.grad)
    xd = x.detach().clone().to(DEV).requires_grad_(True)
    loss, nll = ctc_loss(xd, hl, ys)
    loss.sum().backward()
    assert_close(loss, res[torch.float32][0], truth=res[torch.float64][0], what="loss")
    assert_close(nll, res[torch.float32][1], truth=res[torch.float64][1], what="nll")
    assert_close(xd.grad, res[torch.float32][2], truth=res[torch.float64][2], what="d logits")
    # against fp64 truth the renormalised recursion is at least as good as the fp32 reference (whose own
    # log-domain rounding grows with Th: ~1e-4 at Th=300, S=281)
    assert rel_err(xd.grad, res[torch.float64][2]) < 5e-5 + 2 * rel_err(res[torch.float32][2], res[torch.float64][2])
    # best path == argmax of the oracle's log_softmax
    lsm, best = log_softmax_rows(x.to(DEV), want_best=True)
    assert_close(lsm, F.log_softmax(x, dim=2), what="log_softmax")
    assert torch.equal(best.cpu().long(), F.log_softmax(x, dim=2).argmax(2))


def test_full_size_properties():
    """Config 3 size (Th=200, B=32, V=4233, U=40): per-frame gradient rows sum to zero (softmax minus
    a distribution), padded frames are zero, repeated evaluation is bit-identical."""
    g = torch.Generator().manual_seed(3)
    B, Th, V, U = 32, 200, 4233, 40
    x = (torch.randn(B, Th, V, generator=g) * 2).to(DEV).requires_grad_(True)
    hl = sorted([int(v) for v in torch.randint(120, Th + 1, (B,), generator=g)], reverse=True)
    ys = synth.targets(B=B, V=V, hlens=hl, seed=3, fixed_U=U)
    loss, nll = ctc_loss(x, hl, ys)
    loss.sum().backward()
    g1 = x.grad.clone()
    rows = g1.sum(-1)
    assert float(rows.abs().max()) < 1e-5
    for b in range(B):
        assert torch.all(g1[b, hl[b]:] == 0)
    assert torch.isfinite(nll).all() and float(loss) == pytest.approx(float(nll.sum()) / B, rel=1e-6)
    x.grad = None
    loss2, _ = ctc_loss(x, hl, ys)
    loss2.sum().backward()
    assert torch.equal(loss, loss2)
    # spot-check 3 utterances against the fp64 numpy alpha/beta restatement
    for b in (0, 17, 31):
        r = o_ctc.ctc_alpha_beta(x.detach()[b, :hl[b]].cpu().numpy(), ys[b].numpy())
        assert abs(float(nll[b]) - r["nll"]) < 1e-5 * r["nll"]
        assert rel_err(g1[b, :hl[b]] * B, r["grad"]) < 5e-5


def test_golden_prefix_score_dropin_and_batch():
    g = golden("prefix")
    lpz = g["lpz"]
    eos = lpz.shape[1] - 1
    sc = CTCPrefixScore(lpz, 0, eos, np)
    assert np.array_equal(sc.initial_state(), g["r0"])
    for step in range(4):
        start = int(g["start_%d" % step])
        psi, r = sc(list(g["y_%d" % step]), torch.from_numpy(g["cs_%d" % step]), g["rprev_%d" % step])
        assert psi.dtype == np.float32 and r.shape == (len(g["cs_%d" % step]), lpz.shape[0], 2)
        assert_close(psi, g["psi_%d" % step], what="log_psi step %d" % step)
        assert_close(r[:, start - 1:], g["r_%d" % step], what="r step %d" % step)
    # batched: the four calls as four hypotheses of one launch (pad candidate sets to 4)
    H, C, Tn = 4, 4, lpz.shape[0]
    cs = np.zeros((H, C), np.int32)
    for h in range(H):
        c = g["cs_%d" % h]
        cs[h, :len(c)] = c
        cs[h, len(c):] = c[0]
    rp = np.stack([g["rprev_%d" % h] for h in range(H)])
    last = np.array([g["y_%d" % h][-1] for h in range(H)], np.int32)
    ol = np.array([len(g["y_%d" % h]) - 1 for h in range(H)], np.int32)
    psi, r = ctc_prefix_score_batch(T(lpz).to(DEV), T(rp).to(DEV), T(cs).to(DEV), T(last).to(DEV), T(ol).to(DEV), 0, eos)
    for h in range(H):
        n = len(g["cs_%d" % h])
        assert_close(psi[h, :n], g["psi_%d" % h], what="batched psi %d" % h)
