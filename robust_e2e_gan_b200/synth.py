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
