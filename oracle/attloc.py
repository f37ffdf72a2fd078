"""Oracle (CPU restatement) of location-aware attention.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/model/e2e_attention.py:199-299 (class AttLoc) with the
helpers it uses: linear_tensor (model/e2e_common.py:178-187) and pad_list
(model/e2e_common.py:208-217).  Parameter names and shapes are the reference's:
  mlp_enc.weight (A,D)  mlp_enc.bias (A)   mlp_dec.weight (A,Z)
  mlp_att.weight (A,C)  loc_conv.weight (C,1,1,2*filts+1)   gvec.weight (1,A)  gvec.bias (1)
Quirks reproduced deliberately: softmax over ALL Th frames (no length mask,
e2e_attention.py:282-288); scaling=2.0; the initial alignment is uniform over
enc_hs_len[b] and zero-padded (e2e_attention.py:264-268).
"""
import torch
import torch.nn.functional as F


def init_params(eprojs=320, dunits=300, att_dim=320, aconv_chans=10, aconv_filts=100,
                seed=0, dtype=torch.float32, scale=1.0):
    """Seeded LeCun-normal-like parameters with the reference's names and shapes."""
    g = torch.Generator().manual_seed(seed)

    def n(*shape, fan_in):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale / fan_in ** 0.5).to(dtype)

    K = 2 * aconv_filts + 1
    return {
        "mlp_enc.weight": n(att_dim, eprojs, fan_in=eprojs),
        "mlp_enc.bias": n(att_dim, fan_in=4.0),
        "mlp_dec.weight": n(att_dim, dunits, fan_in=dunits),
        "mlp_att.weight": n(att_dim, aconv_chans, fan_in=aconv_chans),
        "loc_conv.weight": n(aconv_chans, 1, 1, K, fan_in=K),
        "gvec.weight": n(1, att_dim, fan_in=att_dim),
        "gvec.bias": n(1, fan_in=4.0),
    }


def precompute(p, enc_hs_pad):
    """e2e_attention.py:252-256: pre_compute_enc_h = linear_tensor(mlp_enc, enc_h)."""
    B, Th, D = enc_hs_pad.shape
    y = F.linear(enc_hs_pad.contiguous().view(-1, D), p["mlp_enc.weight"], p["mlp_enc.bias"])
    return y.view(B, Th, -1)


def uniform_att_prev(enc_hs_pad, enc_hs_len):
    """e2e_attention.py:264-268 with pad_list(.., 0)."""
    B, Th = enc_hs_pad.shape[0], enc_hs_pad.shape[1]
    a = enc_hs_pad.new_zeros(B, Th)
    for b, l in enumerate(enc_hs_len):
        l = int(l)
        a[b, :l] = 1.0 / l
    return a


def step(p, enc_hs_pad, pre, enc_hs_len, dec_z, att_prev, scaling=2.0, aact_fuc="softmax"):
    """One AttLoc.forward call after the pre-compute (e2e_attention.py:258-299)."""
    B, Th, _ = enc_hs_pad.shape
    A = p["mlp_enc.weight"].shape[0]
    Z = p["mlp_dec.weight"].shape[1]
    C = p["mlp_att.weight"].shape[1]
    filts = (p["loc_conv.weight"].shape[3] - 1) // 2
    if dec_z is None:
        dec_z = enc_hs_pad.new_zeros(B, Z)
    else:
        dec_z = dec_z.view(B, Z)
    if att_prev is None:
        att_prev = uniform_att_prev(enc_hs_pad, enc_hs_len)
    att_conv = F.conv2d(att_prev.view(B, 1, 1, Th), p["loc_conv.weight"], padding=(0, filts))
    att_conv = att_conv.squeeze(2).transpose(1, 2)                      # (B,Th,C)
    att_conv = F.linear(att_conv.contiguous().view(-1, C), p["mlp_att.weight"]).view(B, Th, A)
    dec_z_tiled = F.linear(dec_z, p["mlp_dec.weight"]).view(B, 1, A)
    t = torch.tanh(att_conv + pre + dec_z_tiled)
    e = F.linear(t.view(-1, A), p["gvec.weight"], p["gvec.bias"]).view(B, Th)
    if aact_fuc == "softmax":
        w = F.softmax(scaling * e, dim=1)
    elif aact_fuc == "sigmoid":
        w = torch.sigmoid(scaling * e)
    elif aact_fuc == "sigmoid_softmax":
        w = F.softmax(scaling * torch.sigmoid(e), dim=1)
    else:
        raise ValueError(aact_fuc)
    c = torch.sum(enc_hs_pad * w.view(B, Th, 1), dim=1)
    return c, w


def run_steps(p, enc_hs_pad, enc_hs_len, dec_zs, scaling=2.0):
    """The Decoder's use of AttLoc (model/e2e_decoder.py:114-122): reset, then feed the
    previous alignment back, un-detached.  ``dec_zs`` is a list of (B,Z) tensors or None."""
    pre = precompute(p, enc_hs_pad)
    att_w = None
    cs, ws = [], []
    for z in dec_zs:
        c, att_w = step(p, enc_hs_pad, pre, enc_hs_len, z, att_w, scaling)
        cs.append(c)
        ws.append(att_w)
    return cs, ws
