"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (build container only).

    python -m oracle.gen_golden

Imports model/feat_model.py, model/e2e_attention.py and model/e2e_ctc.py from /root/reference through
oracle/refshim.py (stubs listed there), feeds them small seeded inputs and stores inputs, outputs and
autograd gradients.  The committed fixtures pin (a) the oracle restatement (tests/test_oracle.py, CPU)
and (b) the CUDA path (tests/test_gpu_golden.py).  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def npy(t):
    return t.detach().cpu().numpy()


def gen_fbank(ns):
    torch.manual_seed(11)
    B, T, F = 2, 13, 257
    model = ns.FbankModel(refshim.fbank_args(fbank_dim=80, fbank_opti_type='train'))
    g = torch.Generator().manual_seed(12)
    mag = (torch.randn(B, T, F, generator=g).abs() * torch.exp(1.5 * torch.randn(B, T, F, generator=g)) * 300).clamp(0, 3e5)
    mag[0, 3] = 0.0                      # a silent frame: exercises the <= 1e-7 clamp (zero gradient)
    mag[1, 10:] = 0.0                    # collate padding
    logits = 2.0 * torch.randn(B, T, F, generator=g)
    lens = torch.tensor([13, 10], dtype=torch.int32)
    cmvn = torch.stack([-(5 + 10 * torch.rand(80, generator=g)), 0.3 + 0.7 * torch.rand(80, generator=g)])
    dY = torch.randn(B, T, 80, generator=g)
    out = {}
    # (a) single-input form, with and without CMVN
    x = mag.clone().requires_grad_(True)
    y = model(x, cmvn)
    y.backward(dY)
    out.update(mag=npy(mag), cmvn=npy(cmvn), dY=npy(dY), fc=npy(model.fc), y_cmvn=npy(y), dmag=npy(x.grad),
               dfc=npy(model.fc.grad))
    out["y_plain"] = npy(model(mag))
    # (b) mask tail of EnhanceModel.forward (enhance_model.py:157-164; torch.bool mask, see refshim notes)
    lo = logits.clone().requires_grad_(True)
    s = torch.sigmoid(lo)
    pad = torch.zeros(B, T, 1, dtype=torch.bool)
    for i, l in enumerate(lens):
        pad[i, int(l):] = True
    enh = s.masked_fill(pad, 0) * mag
    model.fc.grad = None
    y2 = model(enh, cmvn)
    y2.backward(dY)
    out.update(logits=npy(logits), lens=npy(lens), enh=npy(enh), y_masked=npy(y2), dlogits=npy(lo.grad))
    # (c) compute_cmvn over two batches
    m2 = ns.FbankModel(refshim.fbank_args(fbank_dim=80, train_dataset_len=2, num_utt_cmvn=2))
    r1 = m2.compute_cmvn(mag, lens)
    r2 = m2.compute_cmvn(mag, lens)
    assert r1 is None
    out["cmvn_est"] = np.asarray(r2).copy()
    np.savez_compressed(os.path.join(OUT, "fbank.npz"), **out)


ATT_PARAM_SHAPES = lambda e, d, a, c, f: [  # registration order of AttLoc.__init__ (e2e_attention.py:212-219)
    ("mlp_enc.weight", (a, e)), ("mlp_enc.bias", (a,)), ("mlp_dec.weight", (a, d)), ("mlp_att.weight", (a, c)),
    ("loc_conv.weight", (c, 1, 1, 2 * f + 1)), ("gvec.weight", (1, a)), ("gvec.bias", (1,))]


def make_attloc_inputs(eprojs, dunits, att_dim, chans, filts, B, Th, steps, seed):
    """Seeded parameters and inputs for an AttLoc case; shared by the generator and the tests."""
    g = torch.Generator().manual_seed(seed + 1)
    params = {}
    for name, shape in ATT_PARAM_SHAPES(eprojs, dunits, att_dim, chans, filts):
        fan = 1
        for v in shape[1:]:
            fan *= v
        params[name] = torch.randn(shape, generator=g) * (1.5 / max(1, fan) ** 0.5)
    enc = torch.tanh(torch.randn(B, Th, eprojs, generator=g))
    hlens = sorted([int(v) for v in torch.randint(Th // 2, Th + 1, (B,), generator=g)], reverse=True)
    hlens[0] = Th
    for b in range(B):
        enc[b, hlens[b]:] = 0
    zs = [None] + [0.5 * torch.randn(B, dunits, generator=g) for _ in range(steps - 1)]
    gc = [torch.randn(B, eprojs, generator=g) for _ in range(steps)]
    gw = torch.randn(B, Th, generator=g)
    return params, enc, hlens, zs, gc, gw


def gen_attloc(ns, name, eprojs, dunits, att_dim, chans, filts, B, Th, steps, seed, full):
    att = ns.AttLoc(eprojs, dunits, att_dim, chans, filts, 'softmax')
    params, enc, hlens, zs, gc, gw = make_attloc_inputs(eprojs, dunits, att_dim, chans, filts, B, Th, steps, seed)
    att.load_state_dict(params)
    enc.requires_grad_(True)
    zs = [None] + [z.requires_grad_(True) for z in zs[1:]]
    att.reset()
    a = None
    cs, ws = [], []
    for i in range(steps):
        c, a = att(enc, hlens, zs[i], a)
        cs.append(c)
        ws.append(a)
    loss = sum((c * gci).sum() for c, gci in zip(cs, gc)) + (ws[-1] * gw).sum()
    loss.backward()
    out = dict(hlens=np.array(hlens, dtype=np.int32), dims=np.array([eprojs, dunits, att_dim, chans, filts, B, Th, steps, seed]),
               c=np.stack([npy(c) for c in cs]), w=np.stack([npy(w) for w in ws]))
    small = {"d_" + k.replace('.', '_'): npy(p.grad) for k, p in att.named_parameters()
             if full or p.numel() <= 4096}
    out.update(small)
    out["d_dec_z"] = np.stack([npy(z.grad) for z in zs[1:]])
    if full:
        out.update(enc=npy(enc), gc=np.stack([npy(t) for t in gc]), gw=npy(gw), d_enc=npy(enc.grad),
                   dec_z=np.stack([npy(z) for z in zs[1:]]))
        out.update({"p_" + k.replace('.', '_'): npy(p) for k, p in att.named_parameters()})
    else:
        # inputs are regenerated from the seed by the test; keep checksums + a slice of the big grads
        out.update(enc_sum=np.float64(enc.double().sum().item()), d_enc_slice=npy(enc.grad[:, ::7, ::9]),
                   d_mlp_enc_weight_slice=npy(att.mlp_enc.weight.grad[::11, ::13]),
                   d_mlp_dec_weight_slice=npy(att.mlp_dec.weight.grad[::11, ::13]),
                   d_mlp_att_weight=npy(att.mlp_att.weight.grad), d_loc_conv_weight=npy(att.loc_conv.weight.grad))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def gen_ctc(ns):
    torch.manual_seed(21)
    B, Th, D, V = 4, 17, 12, 13
    ctc = ns.CTC(V, D, 0.0)
    g = torch.Generator().manual_seed(22)
    hs = torch.tanh(torch.randn(B, Th, D, generator=g)).requires_grad_(True)
    hlens = [17, 15, 12, 9]
    ys = [torch.tensor([3, 3, 5, 1]), torch.tensor([7, 2, 2, 2, 9, 4]), torch.tensor([11]), torch.tensor([6, 6, 1])]
    ys_pad = torch.full((B, 6), -1, dtype=torch.long)
    for i, y in enumerate(ys):
        ys_pad[i, :len(y)] = y
    loss = ctc(hs, hlens, ys_pad)
    (0.5 * loss).sum().backward()          # mtlalpha-style scaling: checks grad_output is honoured
    lsm = ctc.log_softmax(hs)
    out = dict(hs=npy(hs), hlens=np.array(hlens, np.int32), ys_pad=npy(ys_pad), W=npy(ctc.ctc_lo.weight),
               b=npy(ctc.ctc_lo.bias), loss=npy(loss), d_hs=npy(hs.grad), d_W=npy(ctc.ctc_lo.weight.grad),
               d_b=npy(ctc.ctc_lo.bias.grad), log_softmax=npy(lsm), best=npy(lsm.argmax(2)).astype(np.int32))
    # gradient w.r.t. the logits themselves (what the kernel emits)
    logits = (torch.randn(B, Th, V, generator=g) * 2).requires_grad_(True)
    l2 = ctc.loss_fn(logits.transpose(0, 1), torch.cat(ys).int(), torch.tensor(hlens, dtype=torch.int32),
                     torch.tensor([len(y) for y in ys], dtype=torch.int32))
    l2.sum().backward()
    out.update(logits=npy(logits), loss_logits=npy(l2), d_logits=npy(logits.grad))
    np.savez_compressed(os.path.join(OUT, "ctc.npz"), **out)


def gen_prefix(ns):
    g = torch.Generator().manual_seed(31)
    T, V = 14, 9
    lpz = torch.log_softmax(torch.randn(T, V, generator=g) * 2, dim=1).numpy()
    eos = V - 1
    sc = ns.CTCPrefixScore(lpz, 0, eos, np)
    r0 = sc.initial_state()
    out = dict(lpz=lpz, r0=r0)
    y = [eos]
    r = r0
    # walk a hypothesis 4 labels deep, scoring candidate sets that include y[-1] and eos
    for step, (cs, pick) in enumerate([([1, 2, 3, eos], 1), ([2, 5, 1, 0 + 4], 2), ([1, eos, 6, 7], 0),
                                       ([1, 3, eos], 0)]):
        cs = np.array(cs)
        psi, rs = sc(y, torch.from_numpy(cs), r)
        start = max(len(y) - 1, 1)
        out["y_%d" % step] = np.array(y)
        out["cs_%d" % step] = cs
        out["rprev_%d" % step] = r.copy()
        out["psi_%d" % step] = psi.copy()
        out["r_%d" % step] = rs[:, start - 1:].copy()      # rows below start-1 are uninitialised in the reference
        out["start_%d" % step] = np.int64(start)
        y = y + [int(cs[pick])]
        r = rs[pick]
        if start - 1 > 0:
            r = r.copy()
            r[:start - 1] = sc.logzero
    np.savez_compressed(os.path.join(OUT, "prefix.npz"), **out)


def gen_beam(ns):
    """Decoder.recognize_beam (model/e2e_decoder.py:170-369) with the reference AttLoc / CTC / CTCPrefixScore on
    seeded small models (tests/helpers.py:beam_case): expected n-best token sequences and scores."""
    import importlib
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    e2e_decoder = importlib.import_module("model.e2e_decoder")
    out = {}
    for name in helpers.BEAM_CASES:
        c, sd, h, checksum = helpers.beam_case(name)
        att = ns.AttLoc(c["D"], c["Z"], c["A"], c["C"], c["filts"], "softmax")
        dec = e2e_decoder.Decoder(c["D"], c["V"], 1, c["Z"], c["sos"], c["eos"], att)
        ctc = ns.CTC(c["V"], c["D"], 0.0)
        dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("ctc_lo")}, strict=True)
        ctc.load_state_dict({k: v for k, v in sd.items() if k.startswith("ctc_lo")}, strict=True)
        dec.eval()
        args = types.SimpleNamespace(beam_size=c["beam"], penalty=c["penalty"], ctc_weight=c["ctc_weight"],
                                     maxlenratio=c["maxlenratio"], minlenratio=c["minlenratio"], nbest=c["nbest"],
                                     lm_weight=0.0)
        with torch.no_grad():
            lpz = ctc.log_softmax(h.unsqueeze(0)).data[0] if c["ctc_weight"] > 0.0 else None
            nbest = dec.recognize_beam(h, lpz, args, [str(i) for i in range(c["V"])], None, None)
            best_path = ctc.log_softmax(h.unsqueeze(0)).data[0].argmax(1)
        out[name + ".n"] = np.array(len(nbest))
        out[name + ".checksum"] = np.array(checksum)
        out[name + ".best_path"] = npy(best_path).astype(np.int32)
        for i, hyp in enumerate(nbest):
            out["%s.yseq%d" % (name, i)] = np.array([int(t) for t in hyp["yseq"]], dtype=np.int32)
            out["%s.score%d" % (name, i)] = np.array(float(hyp["score"]), dtype=np.float64)
        print(name, [(len(hh["yseq"]), round(float(hh["score"]), 4)) for hh in nbest])
    np.savez_compressed(os.path.join(OUT, "beam.npz"), **out)


def gen_kaldi(ns):
    """Archive written / decoded by the reference's data/kaldi_io.py, transform of mix_data_loader.__getitem__ and
    _collate_fn outputs on seeded samples (inputs: tests/test_kaldi_feats.py helpers)."""
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_kaldi_feats as tk
    import types as _types
    # data/audioparse.py imports audio front-ends that are not installed and not used by _collate_fn / kaldi_io
    for missing, attrs in (("librosa", ()), ("python_speech_features", ("fbank", "delta")), ("torchaudio", ()),
                           ("soundfile", ()), ("sox", ()), ("data.extract_fbanks_module", ())):   # last: prebuilt .so for an old numpy ABI
        if missing not in sys.modules:
            try:
                importlib.import_module(missing)
            except Exception:
                m = _types.ModuleType(missing)
                for a in attrs:
                    setattr(m, a, None)
                sys.modules[missing] = m
    kaldi_io = importlib.import_module("data.kaldi_io")
    loader = importlib.import_module("data.mix_data_loader")
    rng = np.random.RandomState(7)
    f32 = rng.randn(9, 257).astype(np.float32)
    f64 = rng.rand(4, 40)
    cm_src = (rng.randn(33, 23) * 3).astype(np.float32)
    ark = os.path.join(OUT, "kaldi_ref.ark")
    out = {}
    with open(ark, "wb") as f:
        for key, m in (("f32", f32), ("f64", f64)):
            f.write((key + " ").encode())
            out["off_" + key] = np.array(f.tell())
            kaldi_io.write_mat(f, m)
        f.write(b"cm ")
        out["off_cm"] = np.array(f.tell())
        f.write(tk.encode_cm(cm_src))
    ref = {k: m for k, m in kaldi_io.read_mat_ark(ark)}
    out.update(f32=ref["f32"], f64=ref["f64"], cm_decoded=ref["cm"])
    # __getitem__ transform (data/mix_data_loader.py:203-207 + audioparse.transform_feat, delta 0, no splice)
    spect = np.abs(rng.randn(11, 257)).astype(np.float32)
    spect[2, :5] = 0.0
    cmvn = np.stack([-rng.rand(257), 0.5 + rng.rand(257)]).astype(np.float32)
    out["spect"] = spect.copy()
    sp = spect.copy()
    sp[sp <= 1e-7] = 1e-7
    log_spect = 10 * np.log10(sp)
    log_spect = (log_spect + cmvn[0, :]) * cmvn[1, :]
    out.update(spect_clamped=sp, cmvn=cmvn, log_spect=log_spect)
    res = loader._collate_fn(tk.make_samples())
    out["c_utt_ids"] = np.array(res[0])
    for i, name in zip(range(2, 7), ("clean", "clean_log", "mix", "mix_log", "cos")):
        out["c_" + name] = npy(res[i])
    out.update(c_targets=npy(res[7]), c_input_sizes=npy(res[8]), c_target_sizes=npy(res[9]))
    np.savez_compressed(os.path.join(OUT, "kaldi.npz"), **out)


def main(only=None):
    if only == "beam":
        gen_beam(refshim.load())
        return
    if only == "kaldi":
        gen_kaldi(refshim.load())
        return
    if not refshim.available():
        raise SystemExit("reference tree not available; fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    ns = refshim.load()
    gen_fbank(ns)
    gen_attloc(ns, "attloc_small", eprojs=32, dunits=24, att_dim=64, chans=10, filts=5, B=3, Th=21, steps=4,
               seed=41, full=True)
    gen_attloc(ns, "attloc_default", eprojs=320, dunits=300, att_dim=320, chans=10, filts=100, B=2, Th=45,
               steps=3, seed=43, full=False)
    gen_ctc(ns)
    gen_prefix(ns)
    gen_beam(ns)
    gen_kaldi(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
