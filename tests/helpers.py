"""Shared comparison helpers.  Tolerance of the north star: 1e-4 RELATIVE on fp32 outputs and grads.

``rel_err(a, ref)`` = max|a - ref| / max(|ref|_inf, tiny): error relative to the tensor's scale.
``assert_close`` additionally accepts an fp64 "truth": a CUDA result passes if it is within tol of
the fp32 reference OR at least as close to the fp64 truth as the fp32 reference is (plus tol) --
the fp32 reference's own rounding must not be held against the kernel.
"""
import numpy as np
import torch

TOL = 1e-4


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


def assert_close(a, ref, tol=TOL, truth=None, what=""):
    e = rel_err(a, ref)
    if e <= tol:
        return e
    if truth is not None:
        e_t = rel_err(a, truth)
        e_ref = rel_err(ref, truth)
        if e_t <= e_ref + tol:
            return e_t
        raise AssertionError("%s: rel err %.3e vs fp32 ref, %.3e vs fp64 truth (ref itself %.3e), tol %.1e"
                             % (what, e, e_t, e_ref, tol))
    raise AssertionError("%s: rel err %.3e > tol %.1e" % (what, e, tol))
