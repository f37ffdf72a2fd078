"""CPU oracle for the Robust_e2e_gan hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``robust_e2e_gan_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU arm.

The oracle restates, in plain CPU PyTorch / numpy (fp32, with an fp64 switch
for error attribution), the algorithm of the reference's three hot modules:

* ``oracle.frontend``  <- model/enhance_model.py:157-164, model/feat_model.py:62-135
* ``oracle.attloc``    <- model/e2e_attention.py:199-299, model/e2e_common.py:178-217
* ``oracle.ctc``       <- model/e2e_ctc.py:17-155 (+ textbook CTC alpha/beta for the
  arithmetic the reference delegates to the un-vendored ``warpctc_pytorch``)

Pinning: the reference ships NO tests, golden vectors or fixtures for this
path.  The restatement is therefore pinned against the
reference modules themselves, imported from /root/reference in the build
container by ``oracle/refshim.py`` and dumped by ``oracle/gen_golden.py`` into
``tests/golden/*.npz`` (committed).  The CTC arithmetic lives in third-party
``warpctc_pytorch`` (not pinned, not vendored): for that piece parity is
UNPINNED by the reference and anchored on ``torch.nn.functional.ctc_loss``
plus an independent numpy alpha/beta restatement.
"""
