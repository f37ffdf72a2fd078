"""GPU parity: AttLoc cluster kernels (fwd + bwd through the drop-in nn.Module) vs oracle and goldens."""
import numpy as np
import pytest
import torch

from conftest import golden
from helpers import assert_close, rel_err
from oracle import attloc as o_att
from oracle.gen_golden import make_attloc_inputs
from robust_e2e_gan_b200 import AttLoc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_module(dims, params, enc, hlens, zs, gc, gw, scaling=2.0):
    e, d, a, c, f = dims
    att = AttLoc(e, d, a, c, f, "softmax").to(DEV)
    att.load_state_dict(params)
    enc = enc.detach().clone().to(DEV).requires_grad_(True)
    zs = [None] + [z.detach().clone().to(DEV).requires_grad_(True) for z in zs[1:]]
    att.reset()
    w = None
    cs, ws = [], []
    for z in zs:
        cvec, w = att(enc, hlens, z, w, scaling)
        cs.append(cvec)
        ws.append(w)
    loss = sum((ci * gi.to(DEV)).sum() for ci, gi in zip(cs, gc)) + (ws[-1] * gw.to(DEV)).sum()
    loss.backward()
    grads = {k: p.grad for k, p in att.named_parameters()}
    return cs, ws, enc.grad, [z.grad for z in zs[1:]], grads


def run_oracle(params, enc, hlens, zs, gc, gw, dt, scaling=2.0):
    p = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in params.items()}
    enc = enc.detach().clone().to(dt).requires_grad_(True)
    zs = [None] + [z.detach().clone().to(dt).requires_grad_(True) for z in zs[1:]]
    cs, ws = o_att.run_steps(p, enc, hlens, zs, scaling)
    loss = sum((ci * gi.to(dt)).sum() for ci, gi in zip(cs, gc)) + (ws[-1] * gw.to(dt)).sum()
    loss.backward()
    return cs, ws, enc.grad, [z.grad for z in zs[1:]], {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("name", ["attloc_small", "attloc_default"])
def test_golden(name):
    g = golden(name)
    e, d, a, c, f, B, Th, steps, seed = [int(v) for v in g["dims"]]
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(e, d, a, c, f, B, Th, steps, seed)
    cs, ws, d_enc, d_zs, grads = run_module((e, d, a, c, f), params, enc, hlens, zs, gc, gw)
    assert_close(torch.stack(cs), g["c"], what="c")
    assert_close(torch.stack(ws), g["w"], what="w")
    assert_close(torch.stack(d_zs), g["d_dec_z"], what="d dec_z")
    assert_close(grads["mlp_att.weight"], g["d_mlp_att_weight"], what="d mlp_att")
    assert_close(grads["loc_conv.weight"], g["d_loc_conv_weight"], what="d loc_conv")
    assert_close(grads["gvec.weight"], g["d_gvec_weight"], what="d gvec")
    assert float(grads["gvec.bias"].abs().max()) < 1e-5          # softmax is shift invariant
    if "d_enc" in g:
        assert_close(d_enc, g["d_enc"], what="d enc")
        assert_close(grads["mlp_enc.weight"], g["d_mlp_enc_weight"], what="d mlp_enc.w")
        assert_close(grads["mlp_enc.bias"], g["d_mlp_enc_bias"], what="d mlp_enc.b")
        assert_close(grads["mlp_dec.weight"], g["d_mlp_dec_weight"], what="d mlp_dec")
    else:
        assert_close(d_enc[:, ::7, ::9], g["d_enc_slice"], what="d enc slice")
        assert_close(grads["mlp_enc.weight"][::11, ::13], g["d_mlp_enc_weight_slice"], what="d mlp_enc slice")
        assert_close(grads["mlp_dec.weight"][::11, ::13], g["d_mlp_dec_weight_slice"], what="d mlp_dec slice")


@pytest.mark.parametrize("dims,B,Th,steps,seed", [
    ((320, 300, 320, 10, 100), 8, 100, 5, 1234),      # BASELINE config 1 shape
    ((320, 300, 320, 10, 100), 32, 200, 3, 3),        # config 3 shape (cluster of 4 per utterance)
    ((320, 300, 320, 10, 100), 1, 163, 3, 5),         # decode-like: B=1, odd Th, cluster of 8
    ((64, 40, 128, 7, 3), 5, 19, 4, 6),               # odd sizes: C != 10 path, short filters
    ((512, 128, 512, 10, 20), 2, 400, 2, 7),          # wide + long: fwd ring wraps, bwd enc_h ring wraps
    ((320, 300, 320, 10, 100), 2, 700, 2, 8),         # long Th: backward splits over a cluster of 16
])
def test_oracle_parity(dims, B, Th, steps, seed):
    e, d, a, c, f = dims
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(e, d, a, c, f, B, Th, steps, seed)
    got = run_module(dims, params, enc, hlens, zs, gc, gw)
    r32 = run_oracle(params, enc, hlens, zs, gc, gw, torch.float32)
    r64 = run_oracle(params, enc, hlens, zs, gc, gw, torch.float64)
    assert_close(torch.stack(got[0]), torch.stack(r32[0]), truth=torch.stack(r64[0]), what="c")
    assert_close(torch.stack(got[1]), torch.stack(r32[1]), truth=torch.stack(r64[1]), what="w")
    assert_close(got[2], r32[2], truth=r64[2], what="d enc")
    assert_close(torch.stack(got[3]), torch.stack(r32[3]), truth=torch.stack(r64[3]), what="d dec_z")
    for k in r32[4]:
        if k == "gvec.bias":
            assert float(got[4][k].abs().max()) < 1e-4 * float(r32[4]["gvec.weight"].abs().max())
            continue
        assert_close(got[4][k], r32[4][k], truth=r64[4][k], what="d " + k)
    # rows of w are distributions over ALL Th frames (no length mask)
    assert torch.allclose(torch.stack(got[1]).sum(-1), torch.ones(steps, B, device=DEV), atol=1e-5)


def test_state_cache_and_reset_and_nograd():
    dims = (320, 300, 320, 10, 100)
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(*dims, 2, 50, 2, 11)
    att = AttLoc(*dims, "softmax").to(DEV)
    att.load_state_dict(params)
    with torch.no_grad():
        c1, w1 = att(enc.to(DEV), hlens, None, None)
        assert att.pre_compute_enc_h is not None and att.h_length == 50
        pre_id = att.pre_compute_enc_h.data_ptr()
        c2, w2 = att(enc.to(DEV), hlens, zs[1].to(DEV), w1)
        assert att.pre_compute_enc_h.data_ptr() == pre_id       # cached until reset()
        att.reset()
        assert att.pre_compute_enc_h is None and att.enc_h is None and att.h_length is None
        c1b, w1b = att(enc.to(DEV), hlens, None, None)
    assert torch.equal(c1, c1b) and torch.equal(w1, w1b)
    p = {k: v for k, v in params.items()}
    pre = o_att.precompute(p, enc)
    co, wo = o_att.step(p, enc, pre, hlens, None, None)
    assert_close(c1, co, what="c (no grad)")
    assert_close(w1, wo, what="w (no grad)")
    assert sorted(att.state_dict().keys()) == sorted(params.keys())
    with pytest.raises(NotImplementedError):
        AttLoc(*dims, "sigmoid").to(DEV)(enc.to(DEV), hlens, None, None)


def test_backward_twice_and_partial_use():
    """Only some step outputs feed the loss; a second forward/backward through the same module
    gives the same gradients (per-reset accumulators are cleared)."""
    dims = (320, 300, 320, 10, 100)
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(*dims, 3, 40, 3, 12)
    att = AttLoc(*dims, "softmax").to(DEV)
    att.load_state_dict(params)
    res = []
    for _ in range(2):
        att.zero_grad()
        att.reset()
        e = enc.to(DEV).requires_grad_(True)
        c0, w0 = att(e, hlens, None, None)
        c1, w1 = att(e, hlens, zs[1].to(DEV), w0)
        c2, w2 = att(e, hlens, zs[2].to(DEV), w1)          # c2/w2 unused by the loss
        (c1 * gc[1].to(DEV)).sum().backward()
        res.append((e.grad.clone(), {k: p.grad.clone() for k, p in att.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0]) or rel_err(res[0][0], res[1][0]) < 1e-6
    for k in res[0][1]:
        if k == "gvec.bias":
            continue
        assert rel_err(res[0][1][k], res[1][1][k]) < 1e-5 or float(res[0][1][k].abs().max()) < 1e-6
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    eo = enc.clone().requires_grad_(True)
    cs, ws = o_att.run_steps(p, eo, hlens, [None, zs[1], zs[2]])
    (cs[1] * gc[1]).sum().backward()
    assert_close(res[0][0], eo.grad, what="d enc (partial use)")
    assert_close(res[0][1]["mlp_dec.weight"], p["mlp_dec.weight"].grad, what="d mlp_dec (partial use)")
    assert_close(res[0][1]["loc_conv.weight"], p["loc_conv.weight"].grad, what="d loc_conv (partial use)")


# ---- the whole decoder loop in one persistent kernel per direction (AttLoc.forward_loop, csrc/attloc_loop.cu) -------
def run_loop(dims, params, enc, hlens, zs, gc, gw_all, scaling=2.0):
    """gw_all: (steps, B, Th) gradient of EVERY step's alignment (the per-step tests only feed the last one)."""
    e, d, a, c, f = dims
    att = AttLoc(e, d, a, c, f, "softmax").to(DEV)
    att.load_state_dict(params)
    enc = enc.detach().clone().to(DEV).requires_grad_(True)
    dz = torch.stack(zs[1:]).detach().clone().to(DEV).requires_grad_(True)
    c_all, w_all = att.forward_loop(enc, hlens, dz, first_none=True, scaling=scaling)
    loss = (c_all * torch.stack(gc).to(DEV)).sum() + (w_all * gw_all.to(DEV)).sum()
    loss.backward()
    return c_all, w_all, enc.grad, dz.grad, {k: p.grad for k, p in att.named_parameters()}, att


def run_oracle_loop(params, enc, hlens, zs, gc, gw_all, dt, scaling=2.0):
    p = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in params.items()}
    enc = enc.detach().clone().to(dt).requires_grad_(True)
    zs = [None] + [z.detach().clone().to(dt).requires_grad_(True) for z in zs[1:]]
    cs, ws = o_att.run_steps(p, enc, hlens, zs, scaling)
    loss = (torch.stack(cs) * torch.stack(gc).to(dt)).sum() + (torch.stack(ws) * gw_all.to(dt)).sum()
    loss.backward()
    return torch.stack(cs), torch.stack(ws), enc.grad, torch.stack([z.grad for z in zs[1:]]), {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("dims,B,Th,steps,seed,every_w", [
    ((320, 300, 320, 10, 100), 32, 200, 41, 3, False),    # the bench shape: clusters of 4, 41 steps
    ((320, 300, 320, 10, 100), 8, 100, 9, 1234, True),    # BASELINE config 1 shape: clusters of 8
    ((320, 300, 320, 10, 100), 4, 16, 6, 7, True),        # smoke shape: clusters of 2, 8 frames per CTA
    ((320, 300, 320, 10, 100), 1, 163, 4, 5, False),      # B = 1, odd Th, clusters of 8 with a short last CTA
    ((64, 40, 128, 7, 3), 5, 19, 4, 6, True),             # C != 10 instantiation, D != A, short filters
    ((320, 300, 320, 10, 100), 3, 37, 2, 9, True),        # two steps only
])
def test_decoder_loop_kernel_matches_oracle(dims, B, Th, steps, seed, every_w):
    e, d, a, c, f = dims
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(e, d, a, c, f, B, Th, steps, seed)
    g = torch.Generator().manual_seed(seed + 77)
    gw_all = torch.randn(steps, B, Th, generator=g) / B ** 0.5 if every_w else torch.zeros(steps, B, Th)
    gw_all[-1] = gw
    att0 = AttLoc(e, d, a, c, f, "softmax")
    assert att0.loop_supported(steps, B, Th), "shape expected to run in the persistent loop kernels"
    from robust_e2e_gan_b200 import _lib
    n0 = _lib.launch_count()
    got = run_loop(dims, params, enc, hlens, zs, gc, gw_all)
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 < 30, "the loop must not fall back to one launch per step"
    r32 = run_oracle_loop(params, enc, hlens, zs, gc, gw_all, torch.float32)
    r64 = run_oracle_loop(params, enc, hlens, zs, gc, gw_all, torch.float64)
    for i, what in enumerate(("c_all", "w_all", "d enc", "d dec_z")):
        assert_close(got[i], r32[i], truth=r64[i], what=what)
    for k in r32[4]:
        if k == "gvec.bias":
            assert float(got[4][k].abs().max()) < 1e-4 * float(r32[4]["gvec.weight"].abs().max())
            continue
        assert_close(got[4][k], r32[4][k], truth=r64[4][k], what="d " + k)
    assert torch.allclose(got[1].sum(-1), torch.ones(steps, B, device=DEV), atol=1e-5)


def test_decoder_loop_falls_back_per_step_for_long_utterances():
    """Th = 700 does not fit on chip for a whole loop: forward_loop must give the same API through the per-step kernels."""
    dims = (320, 300, 320, 10, 100)
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(*dims, 2, 700, 3, 8)
    gw_all = torch.zeros(3, 2, 700)
    gw_all[-1] = gw
    assert not AttLoc(*dims, "softmax").loop_supported(3, 2, 700)
    got = run_loop(dims, params, enc, hlens, zs, gc, gw_all)
    r32 = run_oracle_loop(params, enc, hlens, zs, gc, gw_all, torch.float32)
    r64 = run_oracle_loop(params, enc, hlens, zs, gc, gw_all, torch.float64)
    for i, what in enumerate(("c_all", "w_all", "d enc", "d dec_z")):
        assert_close(got[i], r32[i], truth=r64[i], what=what + " (fallback)")


def test_decoder_loop_replays_identically():
    """Two runs of the loop kernels on the same inputs are bit-identical (fixed summation orders, no atomics)."""
    dims = (320, 300, 320, 10, 100)
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(*dims, 6, 120, 5, 21)
    gw_all = torch.zeros(5, 6, 120)
    gw_all[-1] = gw
    a = run_loop(dims, params, enc, hlens, zs, gc, gw_all)
    b = run_loop(dims, params, enc, hlens, zs, gc, gw_all)
    for i in range(4):
        assert torch.equal(a[i], b[i])
    for k in a[4]:
        assert torch.equal(a[4][k], b[4][k]), k
