"""GPU parity: fused front-end (csrc/fbank.cu through the C ABI) vs the oracle and the reference goldens.
Tolerance: 1e-4 relative to each tensor's scale (north star), fp64 oracle as tie-breaker."""
import numpy as np
import pytest
import torch

from conftest import golden
from helpers import assert_close, rel_err, TOL
from oracle import frontend as o_fe
from robust_e2e_gan_b200 import FbankModel, apply_mask, fbank, masked_fbank, synth
from robust_e2e_gan_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.asarray(a))


class Args:
    def __init__(self, **kw):
        self.__dict__.update(dict(idim=257, fbank_dim=80, enhance_type="blstm", fbank_opti_type="frozen",
                                  train_dataset_len=1000, num_utt_cmvn=100))
        self.__dict__.update(kw)


def test_golden_single_input_and_dfc():
    g = golden("fbank")
    m = FbankModel(Args(fbank_opti_type="train")).to(DEV)
    m.fc.data.copy_(T(g["fc"]))
    x = T(g["mag"]).to(DEV).requires_grad_(True)
    n0 = _lib.launch_count()
    y = m(x, T(g["cmvn"]))
    y.backward(T(g["dY"]).to(DEV))
    assert _lib.launch_count() - n0 >= 3           # fwd + bwd + dfc kernels really ran
    assert_close(y, g["y_cmvn"], what="Y")
    assert_close(x.grad, g["dmag"], what="d mag")
    assert_close(m.fc.grad, g["dfc"], what="d fc")
    assert_close(m(T(g["mag"]).to(DEV)), g["y_plain"], what="Y no cmvn")
    assert torch.all(x.grad[0, 3] == 0)            # clamped frame: exactly zero gradient
    # CPU input is moved like the reference's to_cuda does
    assert_close(m(T(g["mag"]), T(g["cmvn"])), g["y_cmvn"], what="Y from CPU input")


def test_fresh_default_model_is_the_reference_model():
    """No ``fc`` injection: FbankModel(fbank_dim=80) as constructed must already BE the reference's model
    (bit-identical filter table, model/feat_model.py:15-33) and reproduce the reference's outputs."""
    g = golden("fbank")
    m = FbankModel(Args()).to(DEV)
    assert np.array_equal(m.fc.detach().cpu().numpy(), g["fc"])
    assert_close(m(T(g["mag"]).to(DEV), T(g["cmvn"])), g["y_cmvn"], what="Y (fresh default model)")
    assert_close(m(T(g["mag"]).to(DEV)), g["y_plain"], what="Y no cmvn (fresh default model)")


def test_golden_masked_fused_and_standalone_tail():
    g = golden("fbank")
    fc = T(g["fc"]).to(DEV)
    lo = T(g["logits"]).to(DEV).requires_grad_(True)
    mag = T(g["mag"]).to(DEV)
    y, enh = masked_fbank(lo, mag, T(g["lens"]), fc, T(g["cmvn"]).to(DEV), return_enhanced=True)
    y.backward(T(g["dY"]).to(DEV))
    assert_close(y, g["y_masked"], what="Y masked")
    assert_close(enh, g["enh"], what="enhance_out")
    assert_close(lo.grad, g["dlogits"], what="d logits (fused)")
    assert torch.all(lo.grad[1, 10:] == 0)
    lo2 = T(g["logits"]).to(DEV).requires_grad_(True)
    e2 = apply_mask(lo2, mag, T(g["lens"]))
    y2 = fbank(e2, fc, T(g["cmvn"]).to(DEV))
    y2.backward(T(g["dY"]).to(DEV))
    assert_close(e2, g["enh"], what="apply_mask")
    assert_close(lo2.grad, g["dlogits"], what="d logits (two-stage)")


@pytest.mark.parametrize("B,T_,M,seed", [(8, 400, 40, 1234), (3, 101, 80, 5), (2, 7, 23, 6), (1, 1, 40, 8), (2, 33, 96, 9)])
def test_oracle_parity_shapes(B, T_, M, seed):
    d = synth.frontend_batch(B=B, T=T_, seed=seed, zeros=min(16, B * T_))
    g = torch.Generator().manual_seed(seed)
    fc = torch.rand(257, M, generator=g) * (torch.rand(257, M, generator=g) < 0.2) if M == 23 else synth.mel_fc(257, M)
    cm = synth.cmvn(M, seed)
    dY = torch.randn(B, T_, M, generator=g)
    outs = {}
    for dt in (torch.float32, torch.float64):
        lo = d["mask_logits"].detach().clone().to(dt).requires_grad_(True)
        y = o_fe.masked_fbank_forward(lo, d["mix"].to(dt), d["lens"], fc.to(dt), cm.to(dt))
        y.backward(dY.to(dt))
        outs[dt] = (y.detach(), lo.grad)
    lo = d["mask_logits"].detach().clone().to(DEV).requires_grad_(True)
    y = masked_fbank(lo, d["mix"].to(DEV), d["lens"], fc.to(DEV), cm.to(DEV))
    y.backward(dY.to(DEV))
    assert_close(y, outs[torch.float32][0], truth=outs[torch.float64][0], what="Y")
    assert_close(lo.grad, outs[torch.float32][1], truth=outs[torch.float64][1], what="d logits")
    # single-input form on the clean channel
    x32 = d["clean"].clone().requires_grad_(True)
    y32 = o_fe.fbank_forward(x32, fc, None)
    y32.backward(dY)
    x = d["clean"].to(DEV).requires_grad_(True)
    yy = fbank(x, fc.to(DEV))
    yy.backward(dY.to(DEV))
    assert_close(yy, y32, what="Y single")
    assert_close(x.grad, x32.grad, what="d mag single")


def test_full_size_properties():
    """BASELINE config 2 size (B=32, T=800): properties that do not need the oracle."""
    d = synth.frontend_batch(B=32, T=800, seed=2)
    fc = synth.mel_fc(257, 40).to(DEV)
    mix, lo, lens = d["mix"].to(DEV), d["mask_logits"].to(DEV), d["lens"]
    y = masked_fbank(lo, mix, lens, fc)
    # (1) fused == two-stage
    y2 = fbank(apply_mask(lo, mix, lens), fc)
    assert rel_err(y, y2) < 1e-6
    # (2) padded frames are exactly log(1e-7)
    pad = torch.arange(800, device=DEV)[None, :] >= lens.to(DEV)[:, None]
    assert torch.all(y[pad] == float(np.log(np.float32(1e-7))))
    # (3) scaling the magnitudes by a shifts un-clamped log-mels by 2 log a
    a = 3.0
    ys = fbank(mix * a, fc)
    y1 = fbank(mix, fc)
    ok = y1 > -10
    assert torch.allclose((ys - y1)[ok], torch.full_like(ys[ok], 2 * np.log(a)), atol=2e-5)
    # (4) CMVN is affine
    cm = synth.cmvn(40).to(DEV)
    assert rel_err(fbank(mix, fc, cm), (y1 + cm[0]) * cm[1]) < 1e-6


def test_compute_cmvn_matches_reference_golden():
    g = golden("fbank")
    m = FbankModel(Args(train_dataset_len=2, num_utt_cmvn=2)).to(DEV)
    m.fc.data.copy_(T(g["fc"]))
    assert m.compute_cmvn(T(g["mag"]), g["lens"]) is None
    est = m.compute_cmvn(T(g["mag"]), g["lens"])
    assert est.shape == (2, 80) and est.dtype == np.float32
    assert_close(est, g["cmvn_est"], what="cmvn")
    assert m.frame_count == 23 and m.cmvn_processed_num == 2


def test_errors_are_loud():
    fc = synth.mel_fc(257, 40).to(DEV)
    with pytest.raises((RuntimeError, AssertionError)):
        fbank(torch.rand(1, 4, 100, device=DEV), fc)          # wrong idim


# ---- banded (mel) bank: the streaming kernels of csrc/fbank_band.cu ------------------------------------------------
@pytest.mark.parametrize("B,T_,M,seed", [(8, 400, 40, 1234), (32, 800, 40, 2), (4, 200, 80, 3), (2, 12, 40, 4)])
def test_banded_kernels_match_oracle_and_dense_path(B, T_, M, seed):
    """The banded kernels are what a frozen mel bank runs on; same results as the oracle AND as the dense tcgen05/SIMT
    path (forced by perturbing nothing but the dispatch), masked and single-input forms, forward and backward."""
    from robust_e2e_gan_b200 import feat_model as fm
    d = synth.frontend_batch(B=B, T=T_, seed=seed, zeros=min(16, B * T_))
    g = torch.Generator().manual_seed(seed)
    fc = (synth.mel_fc(257, M) if M == 40 else torch.from_numpy(fm.reference_fbank80().T.astype(np.float32))).to(DEV)
    assert fm._band_for(fc, B, T_) is not None, "expected the banded path for this bank / shape"
    cm = synth.cmvn(M, seed)
    dY = torch.randn(B, T_, M, generator=g)
    outs = {}
    for dt in (torch.float32, torch.float64):
        lo = d["mask_logits"].detach().clone().to(dt).requires_grad_(True)
        y = o_fe.masked_fbank_forward(lo, d["mix"].to(dt), d["lens"], fc.cpu().to(dt), cm.to(dt))
        y.backward(dY.to(dt))
        mg = d["clean"].detach().clone().to(dt).requires_grad_(True)
        y1 = o_fe.fbank_forward(mg, fc.cpu().to(dt), cm.to(dt))
        y1.backward(dY.to(dt))
        outs[dt] = (y.detach(), lo.grad, y1.detach(), mg.grad)
    n0 = _lib.launch_count()
    lo = d["mask_logits"].to(DEV).requires_grad_(True)
    y = masked_fbank(lo, d["mix"].to(DEV), d["lens"], fc, cm.to(DEV))
    y.backward(dY.to(DEV))
    mg = d["clean"].to(DEV).requires_grad_(True)
    y1 = fbank(mg, fc, cm.to(DEV))
    y1.backward(dY.to(DEV))
    assert _lib.launch_count() - n0 == 4          # one banded launch per direction and form
    r32, r64 = outs[torch.float32], outs[torch.float64]
    for got, i, what in ((y, 0, "Y masked"), (lo.grad, 1, "d logits"), (y1, 2, "Y plain"), (mg.grad, 3, "d mag")):
        assert_close(got, r32[i], truth=r64[i], what=what + " (banded)")
    # the dense kernels on the same inputs (a dense-looking copy of the bank: tiny weights everywhere keep it off the band path)
    fcd = fc.clone()
    fcd[0, :] += 1e-30
    assert fm._band_for(fcd, B, T_) is None
    lo2 = d["mask_logits"].to(DEV).requires_grad_(True)
    y2 = masked_fbank(lo2, d["mix"].to(DEV), d["lens"], fcd, cm.to(DEV))
    y2.backward(dY.to(DEV))
    assert_close(y, y2, what="banded vs dense Y")
    assert_close(lo.grad, lo2.grad, what="banded vs dense d logits")
    assert torch.all(lo.grad[B - 1, int(d["lens"][B - 1]):] == 0)      # padded frames: exactly zero gradient


def test_joint_three_output_forward_is_one_launch():
    """FbankModel.forward_joint = the three front-end calls of joint_train.py:158-161 in one launch (mix read once)."""
    B, T_, M = 8, 400, 40
    d = synth.frontend_batch(B=B, T=T_, seed=21)
    m = FbankModel(Args(fbank_dim=M)).to(DEV)
    m.fc.data.copy_(synth.mel_fc(257, M))
    cm = synth.cmvn(M, 5).to(DEV)
    dY = torch.randn(B, T_, M, generator=torch.Generator().manual_seed(3)).to(DEV)
    lo = d["mask_logits"].to(DEV).requires_grad_(True)
    m.forward_joint(lo, d["mix"].to(DEV), d["clean"].to(DEV), d["lens"], cm)          # warm-up: band tables
    n0 = _lib.launch_count()
    enh, mixf, cleanf = m.forward_joint(lo, d["mix"].to(DEV), d["clean"].to(DEV), d["lens"], cm)
    assert _lib.launch_count() - n0 == 1
    enh.backward(dY)
    assert _lib.launch_count() - n0 == 2
    lo2 = d["mask_logits"].to(DEV).requires_grad_(True)
    e2 = m.forward_masked(lo2, d["mix"].to(DEV), d["lens"], cm)
    e2.backward(dY)
    assert torch.equal(enh, e2) and torch.equal(lo.grad, lo2.grad)
    assert torch.equal(mixf, m(d["mix"].to(DEV), cm)) and torch.equal(cleanf, m(d["clean"].to(DEV), cm))
    ref = o_fe.fbank_forward(d["clean"], m.fc.detach().cpu(), cm.cpu())
    assert_close(cleanf, ref, what="clean_feat (joint)")
    assert not mixf.requires_grad and not cleanf.requires_grad
