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


# ---- beam-search cases (config 5): defined next to the other synthetic data so that tools / bench do not import tests
from robust_e2e_gan_b200.synth import BEAM_CASES, beam_case  # noqa: E402,F401
