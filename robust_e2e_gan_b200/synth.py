"""Seeded synthetic AISHELL-shaped batches for the hot path.

Everything is generated on the CPU with ``torch.Generator().manual_seed`` so the
same tensors feed the CUDA path, the oracle and the golden-vector generator.
Shapes follow the reference's collate (data/mix_data_loader.py:264-302: batch
sorted longest first, zero padded) and options (options/base_options.py).
"""
import math

import numpy as np
import torch


def frontend_batch(B=8, T=400, F=257, seed=1234, zeros=16):
    """mix / clean magnitudes (B,T,F) >= 0 at int16-STFT scale, mask logits, lengths."""
    g = torch.Generator().manual_seed(seed)

    def mag():
        m = torch.randn(B, T, F, generator=g).abs() * torch.exp(1.5 * torch.randn(B, T, F, generator=g)) * 300.0
        return m.clamp_(0.0, 3.0e5)

    mix, clean = mag(), mag()
    logits = 2.0 * torch.randn(B, T, F, generator=g)
    lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
    lens[0] = T
    lens, _ = torch.sort(lens, descending=True)
    lens = lens.to(torch.int32)
    for b in range(B):                       # collate zero-pads beyond the length
        mix[b, int(lens[b]):] = 0
        clean[b, int(lens[b]):] = 0
    if zeros:                                # a few all-zero frames / bins exercise the 1e-7 clamp
        idx = torch.randint(0, B * T, (zeros,), generator=g)
        mix.view(B * T, F)[idx] = 0
    return {"mix": mix, "clean": clean, "mask_logits": logits, "lens": lens}


def mel_fc(F=257, M=40):
    """(F,M) non-negative triangular mel weights, formula of model/e2e_common.py:104-132."""
    from .feat_model import generic_mel_banks
    return torch.from_numpy(generic_mel_banks(M, nfft=2 * (F - 1)).T.astype(np.float32)).contiguous()


def cmvn(M=40, seed=7):
    g = torch.Generator().manual_seed(seed)
    c = torch.empty(2, M)
    c[0] = -(5.0 + 10.0 * torch.rand(M, generator=g))
    c[1] = 0.3 + 0.7 * torch.rand(M, generator=g)
    return c


def encoder_batch(B=8, Th=100, D=320, lens_T=None, seed=1234):
    """Encoder-output stand-in: tanh(N(0,1)) (BLSTMP ends in tanh, model/e2e_encoder.py:147),
    zeroed beyond hlens as Decoder.forward does (model/e2e_decoder.py:85)."""
    g = torch.Generator().manual_seed(seed + 1)
    h = torch.tanh(torch.randn(B, Th, D, generator=g))
    if lens_T is None:
        hl = torch.randint(int(0.6 * Th), Th + 1, (B,), generator=g)
        hl[0] = Th
        hl, _ = torch.sort(hl, descending=True)
    else:
        hl = torch.tensor([min(Th, int(math.ceil(int(l) / 4.0))) for l in lens_T])
    for b in range(B):
        h[b, int(hl[b]):] = 0
    return h, [int(v) for v in hl]


def targets(B=8, V=4233, hlens=None, umin=8, umax=24, seed=1234, fixed_U=None):
    """List of 1-D int64 label tensors with ids in [1, V-2]; kept CTC-feasible."""
    g = torch.Generator().manual_seed(seed + 2)
    ys = []
    for b in range(B):
        U = fixed_U if fixed_U is not None else int(torch.randint(umin, umax + 1, (1,), generator=g))
        y = torch.randint(1, V - 1, (U,), generator=g)
        if hlens is not None:
            rep = int((y[1:] == y[:-1]).sum())
            while U + rep > hlens[b] and U > 0:
                U -= 1
                y = y[:U]
                rep = int((y[1:] == y[:-1]).sum()) if U > 1 else 0
        ys.append(y.long())
    return ys


def dec_states(B=8, Z=300, steps=25, seed=1234):
    g = torch.Generator().manual_seed(seed + 3)
    return [None] + [0.3 * torch.randn(B, Z, generator=g) for _ in range(steps - 1)]


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
