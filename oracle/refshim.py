"""Import the UNMODIFIED reference modules: from /root/reference in the build container, else from the
byte-identical copy ``oracle/_ref`` that ``oracle/make_ref.py`` materialises there (git-ignored, shipped to the
GPU box with the snapshot; its SHA-256 manifest is verified before import).

Used by ``oracle/gen_golden.py``, by the tests that pin the oracle restatement and the drop-in classes against
the reference itself, and by ``bench.py --impl reference`` / ``cpu_baseline`` (the reference's own CPU path).
TEST / BENCH INFRASTRUCTURE ONLY: nothing under ``robust_e2e_gan_b200/`` imports it.

Shims -- all test-harness only:
  * ``progressbar``  stub  (model/feat_model.py:4 imports it, not installed)
  * ``jiwer``        stub  (model/e2e_decoder.py:5)
  * ``kenlm``        stub  (model/extlm.py:10)
  * ``np.int = int``       (model/e2e_model.py:39, removed from numpy >= 1.24)
  * ``warpctc_pytorch.CTCLoss`` -> F.ctc_loss(log_softmax(acts), reduction='sum') / B
    (the real package is absent and cannot be installed offline)
  * ``model.e2e_encoder.pack_padded_sequence`` gets its ``lengths`` moved to the CPU first: E2E.forward sends
    ``input_sizes`` through ``to_cuda`` (model/e2e_model.py:176) and torch >= 1.7 rejects CUDA lengths there
    (only matters when the reference's encoder itself runs on a GPU, i.e. in the import-swap tests)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"
REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """Directory to import the reference's ``model`` package from, or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "model")):
        return REFERENCE_ROOT
    if os.path.isdir(os.path.join(REF_COPY, "model")):
        from . import make_ref
        if not make_ref.verify(REF_COPY):
            raise RuntimeError("oracle/_ref does not match its manifest (re-run python -m oracle.make_ref)")
        return REF_COPY
    return None


def available():
    return reference_root() is not None


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
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present at %s nor at %s" % (REFERENCE_ROOT, REF_COPY))
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib

    feat_model = importlib.import_module("model.feat_model")
    e2e_attention = importlib.import_module("model.e2e_attention")
    e2e_ctc = importlib.import_module("model.e2e_ctc")
    e2e_common = importlib.import_module("model.e2e_common")
    e2e_encoder = importlib.import_module("model.e2e_encoder")
    if not getattr(e2e_encoder.pack_padded_sequence, "_re2e_cpu_lengths", False):
        _pps = e2e_encoder.pack_padded_sequence

        def pack_padded_sequence(x, lengths, *a, **k):
            if torch.is_tensor(lengths):
                lengths = lengths.cpu()
            return _pps(x, lengths, *a, **k)

        pack_padded_sequence._re2e_cpu_lengths = True
        e2e_encoder.pack_padded_sequence = pack_padded_sequence
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
        e2e_model=importlib.import_module("model.e2e_model"),
        e2e_decoder=importlib.import_module("model.e2e_decoder"),
    )
    _cache["ns"] = ns
    return ns


def fbank_args(idim=257, fbank_dim=80, enhance_type="blstm", fbank_opti_type="frozen",
               train_dataset_len=1000, num_utt_cmvn=100):
    return types.SimpleNamespace(idim=idim, fbank_dim=fbank_dim, enhance_type=enhance_type,
                                 fbank_opti_type=fbank_opti_type,
                                 train_dataset_len=train_dataset_len, num_utt_cmvn=num_utt_cmvn,
                                 gpu_ids=[])


def e2e_args(odim=30, fbank_dim=40, etype="blstmp", elayers=2, eunits=48, eprojs=64, subsample="1_2_2",
             adim=64, aconv_chans=4, aconv_filts=5, dlayers=1, dunits=48, mtlalpha=0.5, gpu_ids=()):
    """Option namespace with exactly the fields model/e2e_model.py:20-137 reads (options/base_options.py names)."""
    return types.SimpleNamespace(
        fbank_dim=fbank_dim, odim=odim, etype=etype, verbose=0, char_list=[str(i) for i in range(odim)],
        mtlalpha=mtlalpha, elayers=elayers, eunits=eunits, eprojs=eprojs, subsample=subsample,
        subsample_type="skip", dropout_rate=0.0, atype="location", adim=adim, aconv_chans=aconv_chans,
        aconv_filts=aconv_filts, dlayers=dlayers, dunits=dunits, lsm_type="", lsm_weight=0.0, fusion=None,
        gpu_ids=list(gpu_ids))


class swapped(object):
    """Context manager: the reference's modules with the import lines of INTEGRATION.md swapped --
    ``model.e2e_model.AttLoc / CTC`` (model/e2e_model.py:14-16), ``model.e2e_decoder.CTCPrefixScore``
    (model/e2e_decoder.py:14) and, optionally, ``model.e2e_model.Decoder`` -- bound to the given replacements.
    Nothing in the reference's files is edited; on exit the original bindings are restored."""

    def __init__(self, AttLoc=None, CTC=None, CTCPrefixScore=None, Decoder=None):
        self.repl = {("model.e2e_model", "AttLoc"): AttLoc, ("model.e2e_model", "CTC"): CTC,
                     ("model.e2e_decoder", "CTCPrefixScore"): CTCPrefixScore,
                     ("model.e2e_model", "Decoder"): Decoder}
        self.saved = {}

    def __enter__(self):
        import importlib
        load()
        for (mod, name), new in self.repl.items():
            if new is None:
                continue
            m = importlib.import_module(mod)
            self.saved[(mod, name)] = getattr(m, name)
            setattr(m, name, new)
        return importlib.import_module("model.e2e_model")

    def __exit__(self, *exc):
        import importlib
        for (mod, name), old in self.saved.items():
            setattr(importlib.import_module(mod), name, old)
        self.saved = {}
        return False
