"""Import the UNMODIFIED reference modules from /root/reference (build container only).

Used by ``oracle/gen_golden.py`` and by the CPU tests that pin the oracle
restatement against the reference itself.  /root/reference does not exist on
the GPU box, so nothing that runs there may call :func:`load`.

Shims -- all test-harness only:
  * ``progressbar``  stub  (model/feat_model.py:4 imports it, not installed)
  * ``jiwer``        stub  (model/e2e_decoder.py:5)
  * ``kenlm``        stub  (model/extlm.py:10)
  * ``np.int = int``       (model/e2e_model.py:39, removed from numpy >= 1.24)
  * ``warpctc_pytorch.CTCLoss`` -> F.ctc_loss(log_softmax(acts), reduction='sum') / B
    (the real package is absent and cannot be installed offline)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


class _WarpCTCLoss(torch.nn.Module):
    """Stand-in with warp-ctc's call shape: acts (T,B,V) raw logits, flat int labels,
    int act_lens / label_lens on host; returns a (1,) tensor = sum_b nll_b / B."""

    def __init__(self, size_average=False, length_average=False):
        super().__init__()
        self.size_average = size_average

    def forward(self, acts, labels, act_lens, label_lens):
        lp = F.log_softmax(acts, dim=2)
        loss = F.ctc_loss(lp, labels.long(), act_lens.long(), label_lens.long(),
                          blank=0, reduction="sum", zero_infinity=False)
        if self.size_average:
            loss = loss / acts.size(1)
        return loss.view(1)


def _install_stubs():
    if "progressbar" not in sys.modules:
        m = types.ModuleType("progressbar")

        class ProgressBar(object):
            def start(self):
                return self

            def update(self, *_a, **_k):
                pass

            def finish(self):
                pass

        m.ProgressBar = ProgressBar
        sys.modules["progressbar"] = m
    if "jiwer" not in sys.modules:
        m = types.ModuleType("jiwer")
        m.wer = lambda *a, **k: 0.0
        sys.modules["jiwer"] = m
    if "kenlm" not in sys.modules:
        sys.modules["kenlm"] = types.ModuleType("kenlm")
    if "warpctc_pytorch" not in sys.modules:
        m = types.ModuleType("warpctc_pytorch")
        m.CTCLoss = _WarpCTCLoss
        sys.modules["warpctc_pytorch"] = m
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001  (reference era alias)


_cache = {}


def load():
    """Return a namespace with the reference's hot-path classes."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    feat_model = importlib.import_module("model.feat_model")
    e2e_attention = importlib.import_module("model.e2e_attention")
    e2e_ctc = importlib.import_module("model.e2e_ctc")
    e2e_common = importlib.import_module("model.e2e_common")
    ns = types.SimpleNamespace(
        FbankModel=feat_model.FbankModel,
        AttLoc=e2e_attention.AttLoc,
        CTC=e2e_ctc.CTC,
        CTCPrefixScore=e2e_ctc.CTCPrefixScore,
        get_filterbanks_generic=e2e_common.get_filterbanks,
        get_filterbanks_80=feat_model.get_filterbanks,
        pad_list=e2e_common.pad_list,
        feat_model=feat_model,
        e2e_common=e2e_common,
    )
    _cache["ns"] = ns
    return ns


def fbank_args(idim=257, fbank_dim=80, enhance_type="blstm", fbank_opti_type="frozen",
               train_dataset_len=1000, num_utt_cmvn=100):
    return types.SimpleNamespace(idim=idim, fbank_dim=fbank_dim, enhance_type=enhance_type,
                                 fbank_opti_type=fbank_opti_type,
                                 train_dataset_len=train_dataset_len, num_utt_cmvn=num_utt_cmvn,
                                 gpu_ids=[])
