"""Oracle (CPU restatement) of robust_e2e_gan_b200.hotpath.HotPath.step.  TEST INFRASTRUCTURE ONLY.

Composes oracle.frontend / oracle.ctc / oracle.attloc in the order of joint_train.py:158-173 with the
same synthetic stand-ins; used by smoke(), the parity tests and bench.py's cpu_baseline /
--impl reference legs (the reference's own CPU path, restated; torch CPU threads = host cores).
"""
import torch

from . import attloc as o_att
from . import ctc as o_ctc
from . import frontend as o_fe


def oracle_step(cfg, b, sd, backward=True, mtlalpha=0.5, dtype=torch.float32):
    """b: host Batch (robust_e2e_gan_b200.hotpath.Batch); sd: state dict (cpu) of HotPath.
    Returns the same keys as HotPath.step, as CPU tensors."""
    cv = lambda t: t.detach().to(dtype)
    steps = cfg["steps"]
    fc = cv(sd["feat.fc"])
    p_att = {k[len("att."):]: cv(v).requires_grad_(backward) for k, v in sd.items() if k.startswith("att.")}
    W = cv(sd["ctc.ctc_lo.weight"]).requires_grad_(backward)
    bias = cv(sd["ctc.ctc_lo.bias"]).requires_grad_(backward)
    mask_logits = cv(b.mask_logits).requires_grad_(backward)
    hpad = cv(b.hpad).requires_grad_(backward)
    dec_z = cv(b.dec_z).requires_grad_(backward)
    cm = cv(b.cmvn)
    enhance_feat = o_fe.masked_fbank_forward(mask_logits, cv(b.mix), b.lens, fc, cm)
    with torch.no_grad():
        clean_feat = o_fe.fbank_forward(cv(b.clean), fc, cm)
        mix_feat = o_fe.fbank_forward(cv(b.mix), fc, cm)
    loss_ctc = o_ctc.ctc_module_forward(W, bias, hpad, b.hlens_list, b.ys)
    zs = [None] + [dec_z[i] for i in range(steps - 1)]
    cs, ws = o_att.run_steps(p_att, hpad, b.hlens_list, zs)
    out = {"enhance_feat": enhance_feat, "clean_feat": clean_feat, "mix_feat": mix_feat, "loss_ctc": loss_ctc,
           "att_c": torch.stack(cs), "att_w": ws[-1]}
    if backward:
        outs = [enhance_feat, loss_ctc, ws[-1]] + cs
        grads = [cv(b.g_feat), torch.full_like(loss_ctc, mtlalpha), cv(b.g_w)] + [cv(b.g_c[i]) for i in range(steps)]
        torch.autograd.backward(outs, grads)
        out.update(d_mask_logits=mask_logits.grad, d_hpad=hpad.grad, d_dec_z=dec_z.grad)
        for k, v in p_att.items():
            out["d_att." + k] = v.grad
        out["d_ctc.ctc_lo.weight"] = W.grad
        out["d_ctc.ctc_lo.bias"] = bias.grad
    return {k: v.detach() for k, v in out.items()}
