"""Shared comparison helpers.  Tolerance of the north star: 1e-4 RELATIVE on fp32 outputs and grads.

Two metrics, both must hold:
``rel_err(a, ref)``  = max|a - ref| / max|ref|: error relative to the tensor's scale (max norm).
``elem_err(a, ref)`` = max_i |a_i - ref_i| / (|ref_i| + FLOOR * max|ref|): ELEMENT-WISE relative error with an
absolute floor of 10 % of the tensor's scale, so small entries are held to (up to 10x) tighter absolute error than
the max-norm alone would (entries below the floor are sums with cancellation: their own relative error is
unbounded for ANY fp32 evaluation order, the reference's included).
``assert_close`` additionally accepts an fp64 "truth": a CUDA result also passes if, in the same metric, it is
at least as close to the fp64 truth as the fp32 reference is (plus tol) -- the fp32 reference's own rounding must
not be held against the kernel.
"""
import numpy as np
import torch

TOL = 1e-4
FLOOR = 0.1


def to_np(x):
    if torch.is_tensor(x):
        return x.detach().cpu().double().numpy()
    return np.asarray(x, dtype=np.float64)


def rel_err(a, ref):
    a, ref = to_np(a), to_np(ref)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = max(np.abs(ref).max(), 1e-30)
    return float(np.abs(a - ref).max() / scale)


def elem_err(a, ref, floor=FLOOR):
    a, ref = to_np(a), to_np(ref)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = max(np.abs(ref).max(), 1e-30)
    return float((np.abs(a - ref) / (np.abs(ref) + floor * scale)).max())


def assert_close(a, ref, tol=TOL, truth=None, what=""):
    worst = 0.0
    for name, metric in (("max-norm", rel_err), ("element-wise", elem_err)):
        e = metric(a, ref)
        if e > tol:
            if truth is None:
                raise AssertionError("%s: %s rel err %.3e > tol %.1e" % (what, name, e, tol))
            e_t, e_ref = metric(a, truth), metric(ref, truth)
            if e_t > e_ref + tol:
                raise AssertionError("%s: %s rel err %.3e vs fp32 ref, %.3e vs fp64 truth (ref itself %.3e), tol %.1e"
                                     % (what, name, e, e_t, e_ref, tol))
            e = e_t
        worst = max(worst, e)
    return worst


# ---- beam-search cases (config 5): seeded decoder / attention / CTC parameters and encoder output ------------------
BEAM_CASES = {
    # name: dims + search parameters.  "eos_bias" lifts the <eos> logit so that hypotheses end and end_detect fires.
    "beam_small": dict(V=30, D=64, Z=48, A=64, C=4, filts=5, Th=25, beam=4, ctc_weight=0.3, nbest=2, penalty=0.0,
                       maxlenratio=0.0, minlenratio=0.0, eos_bias=0.0, seed=101),
    "beam_eos": dict(V=30, D=64, Z=48, A=64, C=4, filts=5, Th=31, beam=5, ctc_weight=0.5, nbest=3, penalty=0.1,
                     maxlenratio=0.0, minlenratio=0.0, eos_bias=2.5, seed=102),
    "beam_att_only": dict(V=40, D=64, Z=48, A=64, C=4, filts=5, Th=20, beam=3, ctc_weight=0.0, nbest=1, penalty=0.0,
                          maxlenratio=0.5, minlenratio=0.1, eos_bias=1.0, seed=103),
    "beam_default_dims": dict(V=120, D=320, Z=300, A=320, C=10, filts=100, Th=40, beam=10, ctc_weight=0.3, nbest=1,
                              penalty=0.0, maxlenratio=0.0, minlenratio=0.0, eos_bias=1.5, seed=104),
}


def beam_case(name):
    """Seeded parameters (reference state_dict names of Decoder / AttLoc / CTC) and encoder output of a case.
    CPU torch.Generator streams are deterministic, so the fixture stores only a checksum of these tensors."""
    c = dict(BEAM_CASES[name])
    V, D, Z, A, C, filts, Th = (c[k] for k in ("V", "D", "Z", "A", "C", "filts", "Th"))
    K = 2 * filts + 1
    g = torch.Generator().manual_seed(c["seed"])

    def n(*shape, fan):
        return torch.randn(*shape, generator=g) / fan ** 0.5

    sd = {
        "att.mlp_enc.weight": n(A, D, fan=D), "att.mlp_enc.bias": n(A, fan=4.0),
        "att.mlp_dec.weight": n(A, Z, fan=Z), "att.mlp_att.weight": n(A, C, fan=C),
        "att.loc_conv.weight": n(C, 1, 1, K, fan=K), "att.gvec.weight": 3.0 * n(1, A, fan=A),
        "att.gvec.bias": n(1, fan=4.0),
        "embed.weight": n(V, Z, fan=1.0),
        "decoder.0.weight_ih": n(4 * Z, Z + D, fan=Z + D), "decoder.0.weight_hh": n(4 * Z, Z, fan=Z),
        "decoder.0.bias_ih": n(4 * Z, fan=4.0), "decoder.0.bias_hh": n(4 * Z, fan=4.0),
        "output.weight": 4.0 * n(V, Z, fan=Z), "output.bias": n(V, fan=4.0),
        "ctc_lo.weight": 4.0 * n(V, D, fan=D), "ctc_lo.bias": n(V, fan=4.0),
    }
    sd["output.bias"][V - 1] += c["eos_bias"]
    h = torch.tanh(torch.randn(Th, D, generator=g))
    c.update(K=K, sos=V - 1, eos=V - 1)
    checksum = float(sum(v.double().abs().sum() for v in sd.values()) + h.double().abs().sum())
    return c, sd, h, checksum
