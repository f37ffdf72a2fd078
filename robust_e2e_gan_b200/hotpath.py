"""The joint-step hot path as one callable: fused front-end x3, CTC, AttLoc decoder-loop, fwd + bwd.

Mirrors the in-scope lines of one iteration of joint_train.py (:158-173, :184-186):
  enhance_feat = feat_model(mask-tail(enhance net output) )      <- grad flows back to the mask logits
  clean_feat   = feat_model(clean_inputs)   mix_feat = feat_model(mix_inputs)      (no grad)
  loss_ctc     = ctc(hpad, hlens, ys)                                              (model/e2e_model.py:192)
  for i in range(olength): att_c, att_w = att(hpad, hlen, z_list[0], att_w)        (model/e2e_decoder.py:121-122)
  backward
The networks around the path (enhancement BLSTM, encoder, LSTMCell decoder, discriminator; cuDNN /
cuBLAS library code, out of scope per SURVEY.md section 8) are replaced by seeded stand-ins: the
mask logits, the encoder output ``hpad``, the decoder states ``dec_z[i]`` and the upstream
gradients arriving at enhance_feat, att_c[i] and the last att_w are synthetic tensors.
"""
import numpy as np
import torch

from . import synth
from .e2e_attention import AttLoc
from .e2e_ctc import CTC, prepare_targets
from .feat_model import FbankModel

DEFAULT_CFG = dict(B=32, T=800, F=257, M=40, Th=200, D=320, A=320, Z=300, C=10, filts=100, V=4233, U=40,
                   steps=41)


class Batch(object):
    """Host (pinned) or device copy of one synthetic AISHELL-shaped batch."""
    FIELDS = ("mix", "clean", "mask_logits", "cmvn", "hpad", "dec_z", "g_feat", "g_c", "g_w", "lens", "hlens")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device, non_blocking=True):
        d = {k: getattr(self, k).to(device, non_blocking=non_blocking) for k in self.FIELDS}
        d["ys"] = self.ys
        d["hlens_list"] = self.hlens_list
        d["targets"] = prepare_targets(self.ys, device) if torch.device(device).type == "cuda" else None
        return Batch(**d)

    def pin(self):
        for k in self.FIELDS:
            setattr(self, k, getattr(self, k).pin_memory())
        return self

    def h2d_bytes(self):
        return int(sum(getattr(self, k).numel() * getattr(self, k).element_size() for k in self.FIELDS)
                   + sum(y.numel() for y in self.ys) * 4)


def make_batch(cfg, seed=1234):
    B, T, F, M, Th, D, Z, V, U, steps = (cfg[k] for k in ("B", "T", "F", "M", "Th", "D", "Z", "V", "U", "steps"))
    fe = synth.frontend_batch(B=B, T=T, F=F, seed=seed)
    hpad, hl = synth.encoder_batch(B=B, Th=Th, D=D, lens_T=None if Th * 4 != T else fe["lens"], seed=seed)
    ys = synth.targets(B=B, V=V, hlens=hl, seed=seed, fixed_U=U)
    g = torch.Generator().manual_seed(seed + 9)
    dec_z = 0.3 * torch.randn(max(steps - 1, 1), B, Z, generator=g)
    return Batch(mix=fe["mix"], clean=fe["clean"], mask_logits=fe["mask_logits"], lens=fe["lens"],
                 cmvn=synth.cmvn(M, seed), hpad=hpad, hlens=torch.tensor(hl, dtype=torch.int32), hlens_list=hl,
                 ys=ys, dec_z=dec_z,
                 g_feat=torch.randn(B, T, M, generator=g) / (B * T * M) ** 0.5,
                 g_c=torch.randn(steps, B, D, generator=g) / (B * D) ** 0.5,
                 g_w=torch.randn(B, Th, generator=g) / B ** 0.5, targets=None)


class _Opt(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


class HotPath(torch.nn.Module):
    """FbankModel + CTC + AttLoc with seeded parameters, and ``step(batch)`` = fwd + bwd."""

    def __init__(self, cfg, seed=1234, mtlalpha=0.5):
        super().__init__()
        self.cfg = dict(cfg)
        c = self.cfg
        self.feat = FbankModel(_Opt(idim=c["F"], fbank_dim=c["M"], enhance_type="blstm", fbank_opti_type="frozen",
                                    train_dataset_len=1000, num_utt_cmvn=100))
        self.feat.fc.data.copy_(synth.mel_fc(c["F"], c["M"]))          # injected (257, M) bank
        self.att = AttLoc(c["D"], c["Z"], c["A"], c["C"], c["filts"], "softmax")
        self.ctc = CTC(c["V"], c["D"], 0.0)
        g = torch.Generator().manual_seed(seed + 17)
        for p in list(self.att.parameters()) + list(self.ctc.parameters()):
            fan = max(1, p[0].numel()) if p.dim() > 1 else 4
            p.data.copy_(torch.randn(p.shape, generator=g) / fan ** 0.5)
        self.mtlalpha = mtlalpha

    def trainable(self):
        return [p for p in self.parameters() if p.requires_grad]

    def state_dict_cpu(self):
        return {k: v.detach().cpu().clone() for k, v in self.state_dict().items()}

    def step(self, b, backward=True, hlens_for_att=None):
        """One fwd+bwd of the hot path on a device Batch.  Returns a dict of outputs and gradients."""
        steps = self.cfg["steps"]
        mask_logits = b.mask_logits.detach().requires_grad_(backward)
        hpad = b.hpad.detach().requires_grad_(backward)
        # one leaf per decoder step (as the LSTMCell outputs are separate tensors in Decoder.forward)
        dec_zs = [b.dec_z[i].detach().requires_grad_(backward) for i in range(steps - 1)]
        # -- front-end (joint_train.py:158-161)
        enhance_feat = self.feat.forward_masked(mask_logits, b.mix, b.lens, b.cmvn)
        with torch.no_grad():
            clean_feat = self.feat(b.clean, b.cmvn)
            mix_feat = self.feat(b.mix, b.cmvn)
        # -- CTC branch (model/e2e_model.py:192)
        loss_ctc = self.ctc(hpad, b.hlens, b.targets if b.targets is not None else b.ys)
        # -- attention decoder loop (model/e2e_decoder.py:114-122)
        self.att.reset()
        att_w = None
        cs = []
        hl = hlens_for_att if hlens_for_att is not None else b.hlens_list
        for i in range(steps):
            z = None if i == 0 else dec_zs[i - 1]
            att_c, att_w = self.att(hpad, hl, z, att_w)
            cs.append(att_c)
        out = {"enhance_feat": enhance_feat, "clean_feat": clean_feat, "mix_feat": mix_feat, "loss_ctc": loss_ctc,
               "att_c": torch.stack(cs), "att_w": att_w}
        if backward:
            outs = [enhance_feat, loss_ctc, att_w] + cs
            grads = [b.g_feat, torch.full_like(loss_ctc, self.mtlalpha), b.g_w] + [b.g_c[i] for i in range(steps)]
            torch.autograd.backward(outs, grads)
            out.update(d_mask_logits=mask_logits.grad, d_hpad=hpad.grad, d_dec_z=[z.grad for z in dec_zs])
            for k, p in self.named_parameters():
                if p.requires_grad and p.grad is not None:
                    out["d_" + k] = p.grad
        return out
