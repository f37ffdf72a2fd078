"""The hot-path step on the REFERENCE's own modules (CPU).  TEST / BENCH INFRASTRUCTURE ONLY.

Same computation as ``robust_e2e_gan_b200.hotpath.HotPath.step`` and ``oracle.hotpath_oracle.oracle_step`` -- the
in-scope lines of one joint_train.py iteration (:158-173, :184-186) -- but executed by the unmodified reference
classes imported through ``oracle/refshim.py`` (from /root/reference, or from the byte-identical ``oracle/_ref`` copy
on the GPU box): ``FbankModel.forward`` (model/feat_model.py:118-135), ``CTC.forward`` (model/e2e_ctc.py:33-66, with
the warp-ctc stand-in of refshim) and ``AttLoc.forward`` x steps (model/e2e_attention.py:236-299) fed back exactly as
Decoder.forward does (model/e2e_decoder.py:114-122).  The one restated line is the mask tail of EnhanceModel.forward
(model/enhance_model.py:157-164): torch 2.x rejects its ByteTensor mask, so a bool mask is used.

Used by ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference") and by tests that pin ``oracle_step``.
"""
import torch

from . import refshim

_models = {}


def _build(cfg, sd):
    """Reference modules with the HotPath parameters loaded (cached per shape)."""
    key = (cfg["F"], cfg["M"], cfg["D"], cfg["Z"], cfg["A"], cfg["C"], cfg["filts"], cfg["V"])
    ns = refshim.load()
    if key not in _models:
        feat = ns.FbankModel(refshim.fbank_args(idim=cfg["F"], fbank_dim=80))      # the reference only builds 80 filters
        feat.fc = torch.nn.Parameter(torch.zeros(cfg["F"], cfg["M"]), requires_grad=False)   # (257, M) bank injected
        att = ns.AttLoc(cfg["D"], cfg["Z"], cfg["A"], cfg["C"], cfg["filts"], "softmax")
        ctc = ns.CTC(cfg["V"], cfg["D"], 0.0)
        _models[key] = (feat, att, ctc)
    feat, att, ctc = _models[key]
    feat.load_state_dict({"fc": sd["feat.fc"]})
    att.load_state_dict({k[len("att."):]: v for k, v in sd.items() if k.startswith("att.")})
    ctc.load_state_dict({k[len("ctc."):]: v for k, v in sd.items() if k.startswith("ctc.")})
    return feat, att, ctc


def mask_tail(linear_out, mix_inputs, ilens):
    """model/enhance_model.py:157-164 with a bool mask (the reference's ByteTensor mask is rejected by torch 2.x)."""
    out = torch.sigmoid(linear_out)
    mask = torch.zeros(out.size(), dtype=torch.bool)
    for i, length in enumerate(ilens):
        length = int(length)
        if (mask[i].size(0) - length) > 0:
            mask[i].narrow(0, length, mask[i].size(0) - length).fill_(True)
    out = out.masked_fill(mask, 0)
    return out * mix_inputs


def reference_step(cfg, b, sd, backward=True, mtlalpha=0.5):
    """b: host Batch; sd: cpu state dict of HotPath.  Returns the keys of HotPath.step as CPU tensors."""
    feat, att, ctc = _build(cfg, sd)
    steps = cfg["steps"]
    for m in (att, ctc):
        m.zero_grad()
    mask_logits = b.mask_logits.detach().clone().requires_grad_(backward)
    hpad = b.hpad.detach().clone().requires_grad_(backward)
    dec_z = [b.dec_z[i].detach().clone().requires_grad_(backward) for i in range(steps - 1)]
    enhance_feat = feat(mask_tail(mask_logits, b.mix, b.lens), b.cmvn)
    with torch.no_grad():
        clean_feat = feat(b.clean, b.cmvn)
        mix_feat = feat(b.mix, b.cmvn)
    loss_ctc = ctc(hpad, b.hlens_list, b.ys)
    att.reset()
    att_w, cs = None, []
    for i in range(steps):
        att_c, att_w = att(hpad, b.hlens_list, None if i == 0 else dec_z[i - 1], att_w)
        cs.append(att_c)
    out = {"enhance_feat": enhance_feat, "clean_feat": clean_feat, "mix_feat": mix_feat, "loss_ctc": loss_ctc,
           "att_c": torch.stack(cs), "att_w": att_w}
    if backward:
        outs = [enhance_feat, loss_ctc, att_w] + cs
        grads = [b.g_feat, torch.full_like(loss_ctc, mtlalpha), b.g_w] + [b.g_c[i] for i in range(steps)]
        torch.autograd.backward(outs, grads)
        out.update(d_mask_logits=mask_logits.grad, d_hpad=hpad.grad,
                   d_dec_z=torch.stack([z.grad for z in dec_z]) if dec_z else torch.zeros(0))
        for k, p in att.named_parameters():
            out["d_att." + k] = p.grad
        for k, p in ctc.named_parameters():
            out["d_ctc." + k] = p.grad
    return {k: v.detach().clone() for k, v in out.items()}
