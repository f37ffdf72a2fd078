"""Drop-in ``AttLoc`` (reference: model/e2e_attention.py:199-299) on the sm_100a cluster kernels.

Same constructor, submodule / parameter names (``mlp_enc``, ``mlp_dec``, ``mlp_att``, ``loc_conv``,
``gvec``), stateful encoder-projection cache with ``reset()``, and
``forward(enc_hs_pad, enc_hs_len, dec_z, att_prev, scaling=2.0) -> (c, w)``; ``w`` can be fed back as
the next ``att_prev`` with its graph intact, exactly as Decoder.forward does
(model/e2e_decoder.py:122).

Autograd layout (what makes the backward cheap):
  * ``_Precompute`` (once per reset) owns ALL parameters and ``enc_hs_pad``; it returns ``pre`` and a
    1-element *anchor*.
  * every ``_Step`` takes the anchor (so the engine runs ``_Precompute.backward`` after the last
    step backward), ``dec_z`` and ``att_prev``.  Its backward launches one cluster kernel that adds
    d pre into a per-reset buffer through the TMA reduce-add path and adds the small parameter
    gradients into per-reset accumulators; it returns only ``d dec_z`` and ``d att_prev``.
  * ``_Precompute.backward`` then emits every parameter gradient once, and forms
    d enc_h = d_pre @ W_enc + sum_i w_i (x) dc_i  (a rank-#steps update, re2e_attloc_enc_grad)
    instead of #steps read-modify-write passes over (B,Th,D).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .linear import colsum, gemm_tf32x3


class _State(object):
    """Per-reset scratch shared by the autograd nodes of one decoder loop."""

    def __init__(self):
        self.enc = None          # (B,Th,D) fp32 contiguous, detached
        self.pre = None          # (B,Th,A) detached
        self.hlens_dev = None
        self.weights = None      # detached contiguous views: W_dec, W_att, W_conv(C,K), gvec(A), gvec_b(1)
        self.W_decT = None       # (Z,A) transposed copy of W_dec for the backward kernel
        self.dims = None         # B,Th,D,A,Z,C,K
        self.clear_grads()

    def clear_grads(self):
        self.d_pre = None
        self.acc = None          # dict of parameter-gradient accumulators
        self.bwd_w, self.bwd_dc, self.bwd_ddp, self.bwd_decz = [], [], [], []

    def ensure_acc(self, dev):
        """Per-CTA private accumulators of the small parameter gradients (no atomics; summed once per loop)."""
        if self.acc is None:
            L = _lib.lib()
            B, Th, D, A, Z, C, K = self.dims
            self.acc_stride = int(L.re2e_attloc_acc_floats(A, C, K))
            self.acc_nslots = int(L.re2e_attloc_acc_slots(B, Th, D, A, Z, C, K))
            _lib.check(min(self.acc_nslots, 0), "re2e_attloc_acc_slots")
            self.acc = torch.zeros(self.acc_nslots, self.acc_stride, device=dev, dtype=torch.float32)


class _Precompute(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc_hs_pad, W_enc, b_enc, W_dec, W_att, W_conv, gvec_w, gvec_b, state):
        _lib.lib()
        dev = W_enc.device
        enc = _lib.f32c(enc_hs_pad.detach(), dev)
        B, Th, D = enc.shape
        A, Z, C, K = W_enc.shape[0], W_dec.shape[1], W_att.shape[1], W_conv.shape[-1]
        # e2e_attention.py:252-256 linear_tensor(mlp_enc, enc_h): a plain dense layer
        pre = torch.empty(B, Th, A, device=dev, dtype=torch.float32)
        gemm_tf32x3(enc.view(B * Th, D), False, _lib.f32c(W_enc.detach()), False, pre.view(B * Th, A), B * Th, A, D,
                    bias=_lib.f32c(b_enc.detach()))
        state.enc, state.pre = enc, pre
        state.dims = (B, Th, D, A, Z, C, K)
        state.weights = (_lib.f32c(W_dec.detach()), _lib.f32c(W_att.detach()),
                         _lib.f32c(W_conv.detach()).view(C, K), _lib.f32c(gvec_w.detach()).view(A),
                         _lib.f32c(gvec_b.detach()).view(1))
        state.W_decT = None      # (Z, A) copy for the backward, built by the first backward step of the loop
        state.clear_grads()
        ctx.state = state
        ctx.save_for_backward(W_enc)
        ctx.set_materialize_grads(False)
        anchor = torch.zeros(1, device=dev, dtype=torch.float32)
        return pre, anchor

    @staticmethod
    def backward(ctx, d_pre_direct, _d_anchor):
        L = _lib.lib()
        st = ctx.state
        (W_enc,) = ctx.saved_tensors
        B, Th, D, A, Z, C, K = st.dims
        dev = st.enc.device
        d_pre = st.d_pre
        if d_pre_direct is not None:
            d_pre = d_pre_direct.contiguous() if d_pre is None else d_pre + d_pre_direct
        need = ctx.needs_input_grad
        d_enc = dW_enc = db_enc = dW_dec = dW_att = dW_conv = dgw = dgb = None
        if d_pre is not None:
            dp2 = d_pre.view(B * Th, A)
            We = _lib.f32c(W_enc.detach())
            if need[0]:
                d_enc = torch.empty(B, Th, D, device=dev, dtype=torch.float32)
                gemm_tf32x3(dp2, False, We, True, d_enc.view(B * Th, D), B * Th, D, A)      # d_pre @ W_enc
            if need[1]:
                dW_enc = torch.empty(A, D, device=dev, dtype=torch.float32)
                gemm_tf32x3(dp2, True, st.enc.view(B * Th, D), True, dW_enc, A, D, B * Th)  # d_pre^T @ enc
            if need[2]:
                db_enc = colsum(dp2)
        if need[0] and st.bwd_w:
            if d_enc is None:
                d_enc = torch.zeros(B, Th, D, device=dev, dtype=torch.float32)
            w_all = torch.stack(st.bwd_w, 0).contiguous()
            dc_all = torch.stack(st.bwd_dc, 0).contiguous()
            with _lib.on(dev):
                _lib.check(L.re2e_attloc_enc_grad(_lib.ptr(w_all), _lib.ptr(dc_all), _lib.ptr(d_enc),
                                                  w_all.shape[0], B, Th, D, 1, _lib.stream_ptr()),
                           "re2e_attloc_enc_grad")
        if need[3] and st.bwd_ddp:
            ddp_all = torch.cat(st.bwd_ddp, 0)
            dz_all = torch.cat(st.bwd_decz, 0)
            dW_dec = torch.empty(A, Z, device=dev, dtype=torch.float32)
            gemm_tf32x3(ddp_all, True, dz_all, True, dW_dec, A, Z, ddp_all.shape[0])
        elif need[3]:
            dW_dec = torch.zeros(A, Z, device=dev, dtype=torch.float32)
        if st.acc is not None:
            tot = torch.empty(A * C + C * K + A + 1, device=dev, dtype=torch.float32)
            with _lib.on(dev):
                _lib.check(L.re2e_attloc_acc_reduce(_lib.ptr(st.acc), st.acc_nslots, _lib.ptr(tot), A, C, K,
                                                    _lib.stream_ptr()), "re2e_attloc_acc_reduce")
            dW_att = tot[:A * C].view(A, C)
            dW_conv = tot[A * C:A * C + C * K].view(C, 1, 1, K)
            dgw = tot[A * C + C * K:A * C + C * K + A].view(1, A)
            dgb = tot[A * C + C * K + A:].view(1)
        st.clear_grads()
        return d_enc, dW_enc, db_enc, dW_dec, dW_att, dW_conv, dgw, dgb, None


class _Step(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, dec_z, att_prev, state, scaling, att_prev_is_init):
        L = _lib.lib()
        st = state
        B, Th, D, A, Z, C, K = st.dims
        dev = st.enc.device
        W_dec, W_att, W_conv, gvec, gvec_b = st.weights
        dz = _lib.f32c(dec_z.detach(), dev).view(B, Z) if dec_z is not None else None
        ap = _lib.f32c(att_prev.detach(), dev).view(B, Th)
        c = torch.empty(B, D, device=dev, dtype=torch.float32)
        w = torch.empty(B, Th, device=dev, dtype=torch.float32)
        need_bwd = any(ctx.needs_input_grad[:3])
        conv = torch.empty(B, Th, C, device=dev, dtype=torch.float32) if need_bwd else None
        xsave = torch.empty(B, Th, A, device=dev, dtype=torch.float32) if need_bwd else None
        with _lib.on(dev):
            _lib.check(L.re2e_attloc_step_fwd(_lib.ptr(st.pre), _lib.ptr(st.enc), _lib.ptr(dz), _lib.ptr(ap),
                                              _lib.ptr(W_dec), _lib.ptr(W_att), _lib.ptr(W_conv), _lib.ptr(gvec),
                                              _lib.ptr(gvec_b), float(scaling), _lib.ptr(c), _lib.ptr(w),
                                              None, _lib.ptr(conv), _lib.ptr(xsave), B, Th, D, A, Z, C, K,
                                              _lib.stream_ptr()), "re2e_attloc_step_fwd")
        if need_bwd:
            ctx.state = st
            ctx.scaling = float(scaling)
            ctx.has_dz = dz is not None
            ctx.skip_dprev = bool(att_prev_is_init)
            ctx.save_for_backward(ap, w, xsave, conv, dz if dz is not None else ap)
            ctx.set_materialize_grads(False)
        return c, w

    @staticmethod
    def backward(ctx, dc, dw):
        L = _lib.lib()
        st = ctx.state
        ap, w, xsave, conv, dz = ctx.saved_tensors
        B, Th, D, A, Z, C, K = st.dims
        dev = st.enc.device
        W_dec, W_att, W_conv, gvec, gvec_b = st.weights
        if dc is None and dw is None:
            return None, None, None, None, None, None
        dc = _lib.f32c(dc, dev) if dc is not None else None
        dw = _lib.f32c(dw, dev) if dw is not None else None
        st.ensure_acc(dev)
        if st.W_decT is None:    # d dec_z = d dec_proj @ W_dec reads W_dec^T rows (coalesced); once per decoder loop
            st.W_decT = W_dec.t().contiguous()
        first = st.d_pre is None
        if first:
            st.d_pre = torch.empty(B, Th, A, device=dev, dtype=torch.float32)
        d_decproj = torch.empty(B, A, device=dev, dtype=torch.float32)
        want_dprev = ctx.needs_input_grad[2] and not ctx.skip_dprev
        d_prev = torch.empty(B, Th, device=dev, dtype=torch.float32) if want_dprev else None
        d_dz = (torch.empty(B, Z, device=dev, dtype=torch.float32)
                if (ctx.has_dz and ctx.needs_input_grad[1]) else None)
        with _lib.on(dev):
            _lib.check(L.re2e_attloc_step_bwd(
                _lib.ptr(dc), _lib.ptr(dw), _lib.ptr(xsave), _lib.ptr(st.enc), _lib.ptr(ap), _lib.ptr(w),
                _lib.ptr(conv), _lib.ptr(W_dec), _lib.ptr(st.W_decT), _lib.ptr(W_att), _lib.ptr(W_conv),
                _lib.ptr(gvec), ctx.scaling, _lib.ptr(st.d_pre), 0 if first else 1, _lib.ptr(d_decproj), _lib.ptr(d_dz),
                _lib.ptr(d_prev), _lib.ptr(st.acc), st.acc_nslots, B, Th, D, A, Z, C, K, _lib.stream_ptr()),
                "re2e_attloc_step_bwd")
        if ctx.has_dz:
            st.bwd_ddp.append(d_decproj)
            st.bwd_decz.append(dz)
        if dc is not None:
            st.bwd_w.append(w)
            st.bwd_dc.append(dc)
        return None, d_dz, d_prev, None, None, None


class _Loop(torch.autograd.Function):
    """All S attention steps of one decoder loop in ONE persistent cluster kernel per direction (csrc/attloc_loop.cu).

    inputs : enc_hs_pad (B,Th,D), dec_z_all (S-1,B,Z) [first_none: step 0 has dec_z = None] or (S,B,Z), parameters
    outputs: c_all (S,B,D), w_all (S,B,Th) -- step i's context / alignment; the alignment is fed back inside the kernel
    """

    @staticmethod
    def forward(ctx, enc_hs_pad, dec_z_all, W_enc, b_enc, W_dec, W_att, W_conv, gvec_w, gvec_b, att_init, scaling,
                first_none):
        L = _lib.lib()
        dev = W_enc.device
        enc = _lib.f32c(enc_hs_pad.detach(), dev)
        dz = _lib.f32c(dec_z_all.detach(), dev)
        B, Th, D = enc.shape
        A, Z, C, K = W_enc.shape[0], W_dec.shape[1], W_att.shape[1], W_conv.shape[-1]
        off = 1 if first_none else 0
        S = dz.shape[0] + off
        We, Wd = _lib.f32c(W_enc.detach()), _lib.f32c(W_dec.detach())
        Wa, Wc = _lib.f32c(W_att.detach()), _lib.f32c(W_conv.detach()).view(C, K)
        gw, gb = _lib.f32c(gvec_w.detach()).view(A), _lib.f32c(gvec_b.detach()).view(1)
        pre = torch.empty(B, Th, A, device=dev, dtype=torch.float32)
        gemm_tf32x3(enc.view(B * Th, D), False, We, False, pre.view(B * Th, A), B * Th, A, D,
                    bias=_lib.f32c(b_enc.detach()))
        # mlp_dec of EVERY step as one dense product (e2e_attention.py:278 x S)
        dec_proj = torch.empty(S, B, A, device=dev, dtype=torch.float32)
        if off:
            dec_proj[0].zero_()
        if S - off > 0:
            gemm_tf32x3(dz.view(-1, Z), False, Wd, False, dec_proj[off:].view(-1, A), (S - off) * B, A, Z)
        att0 = _lib.f32c(att_init.detach(), dev).view(B, Th)
        need_bwd = any(ctx.needs_input_grad[:9])
        c_all = torch.empty(S, B, D, device=dev, dtype=torch.float32)
        w_all = torch.empty(S, B, Th, device=dev, dtype=torch.float32)
        conv_all = torch.empty(S, B, Th, C, device=dev, dtype=torch.float32) if need_bwd else None
        with _lib.on(dev):
            _lib.check(L.re2e_attloc_loop_fwd(_lib.ptr(pre), _lib.ptr(enc), _lib.ptr(dec_proj), _lib.ptr(att0),
                                              _lib.ptr(Wa), _lib.ptr(Wc), _lib.ptr(gw), _lib.ptr(gb), float(scaling),
                                              _lib.ptr(c_all), _lib.ptr(w_all), _lib.ptr(conv_all), S, B, Th, D, A, C, K,
                                              _lib.stream_ptr()), "re2e_attloc_loop_fwd")
        if need_bwd:
            ctx.save_for_backward(enc, pre, dec_proj, att0, w_all, conv_all, dz, We, Wd, Wa, Wc, gw)
            ctx.dims = (S, B, Th, D, A, Z, C, K, off)
            ctx.scaling = float(scaling)
            ctx.set_materialize_grads(False)
        return c_all, w_all

    @staticmethod
    def backward(ctx, dc_all, dw_all):
        L = _lib.lib()
        enc, pre, dec_proj, att0, w_all, conv_all, dz, We, Wd, Wa, Wc, gw = ctx.saved_tensors
        S, B, Th, D, A, Z, C, K, off = ctx.dims
        dev = enc.device
        need = ctx.needs_input_grad
        if dc_all is None and dw_all is None:
            return (None,) * 12
        dc_all = _lib.f32c(dc_all, dev) if dc_all is not None else None
        dw_all = _lib.f32c(dw_all, dev) if dw_all is not None else None
        nslots = int(L.re2e_attloc_loop_slots(S, B, Th, D, A, C, K))
        _lib.check(min(nslots, 0), "re2e_attloc_loop_slots")
        stride = int(L.re2e_attloc_acc_floats(A, C, K))
        acc = torch.empty(nslots, stride, device=dev, dtype=torch.float32)
        d_pre = torch.empty(B, Th, A, device=dev, dtype=torch.float32)
        d_decproj = torch.empty(S, B, A, device=dev, dtype=torch.float32)
        with _lib.on(dev):
            _lib.check(L.re2e_attloc_loop_bwd(_lib.ptr(pre), _lib.ptr(enc), _lib.ptr(dec_proj), _lib.ptr(att0),
                                              _lib.ptr(w_all), _lib.ptr(conv_all), _lib.ptr(dc_all), _lib.ptr(dw_all),
                                              _lib.ptr(Wa), _lib.ptr(Wc), _lib.ptr(gw), ctx.scaling, _lib.ptr(d_pre),
                                              _lib.ptr(d_decproj), _lib.ptr(acc), nslots, S, B, Th, D, A, C, K,
                                              _lib.stream_ptr()), "re2e_attloc_loop_bwd")
        d_enc = d_dz = dW_enc = db_enc = dW_dec = dW_att = dW_conv = dgw = dgb = None
        dp2 = d_pre.view(B * Th, A)
        if need[0]:
            d_enc = torch.empty(B, Th, D, device=dev, dtype=torch.float32)
            gemm_tf32x3(dp2, False, We, True, d_enc.view(B * Th, D), B * Th, D, A)         # d_pre @ W_enc
            if dc_all is not None:                                                          # + sum_i w_i (x) dc_i
                with _lib.on(dev):
                    _lib.check(L.re2e_attloc_enc_grad(_lib.ptr(w_all), _lib.ptr(dc_all), _lib.ptr(d_enc), S, B, Th, D, 1,
                                                      _lib.stream_ptr()), "re2e_attloc_enc_grad")
        if need[2]:
            dW_enc = torch.empty(A, D, device=dev, dtype=torch.float32)
            gemm_tf32x3(dp2, True, enc.view(B * Th, D), True, dW_enc, A, D, B * Th)        # d_pre^T @ enc
        if need[3]:
            db_enc = colsum(dp2)
        rows = (S - off) * B
        ddp2 = d_decproj[off:].view(rows, A)
        if need[1] and rows > 0:
            d_dz = torch.empty(S - off, B, Z, device=dev, dtype=torch.float32)
            gemm_tf32x3(ddp2, False, Wd, True, d_dz.view(rows, Z), rows, Z, A)             # d dec_z = d dec_proj @ W_dec
        if need[4]:
            dW_dec = torch.empty(A, Z, device=dev, dtype=torch.float32)
            if rows > 0:
                gemm_tf32x3(ddp2, True, dz.view(rows, Z), True, dW_dec, A, Z, rows)        # d dec_proj^T @ dec_z
            else:
                dW_dec.zero_()
        tot = torch.empty(A * C + C * K + A + 1, device=dev, dtype=torch.float32)
        with _lib.on(dev):
            _lib.check(L.re2e_attloc_acc_reduce(_lib.ptr(acc), nslots, _lib.ptr(tot), A, C, K, _lib.stream_ptr()),
                       "re2e_attloc_acc_reduce")
        dW_att = tot[:A * C].view(A, C)
        dW_conv = tot[A * C:A * C + C * K].view(C, 1, 1, K)
        dgw = tot[A * C + C * K:A * C + C * K + A].view(1, A)
        dgb = tot[A * C + C * K + A:].view(1)
        return d_enc, d_dz, dW_enc, db_enc, dW_dec, dW_att, dW_conv, dgw, dgb, None, None, None


class AttLoc(torch.nn.Module):
    """location-aware attention (model/e2e_attention.py:199-299).

    :param int eprojs: # projection-units of encoder
    :param int dunits: # units of decoder
    :param int att_dim: attention dimension
    :param int aconv_chans: # channels of attention convolution
    :param int aconv_filts: filter size of attention convolution
    :param str aact_fuc: 'softmax' (the only value E2E ever passes, model/e2e_model.py:71-72)
    """

    def __init__(self, eprojs, dunits, att_dim, aconv_chans, aconv_filts, aact_fuc='softmax'):
        super(AttLoc, self).__init__()
        self.mlp_enc = torch.nn.Linear(eprojs, att_dim)
        self.mlp_dec = torch.nn.Linear(dunits, att_dim, bias=False)
        self.mlp_att = torch.nn.Linear(aconv_chans, att_dim, bias=False)
        self.loc_conv = torch.nn.Conv2d(
            1, aconv_chans, (1, 2 * aconv_filts + 1), padding=(0, aconv_filts), bias=False)
        self.gvec = torch.nn.Linear(att_dim, 1)

        self.dunits = dunits
        self.eprojs = eprojs
        self.att_dim = att_dim
        self.h_length = None
        self.enc_h = None
        self.pre_compute_enc_h = None
        self.aconv_chans = aconv_chans
        self.aact_fuc = aact_fuc
        self._state = None
        self._anchor = None

    def reset(self):
        '''reset states'''
        self.h_length = None
        self.enc_h = None
        self.pre_compute_enc_h = None
        self._state = None
        self._anchor = None

    def loop_supported(self, steps, batch, h_length):
        """True when ``forward_loop`` can run this shape in the persistent loop kernels (the frame range a CTA owns of
        one utterance must stay resident on chip for the whole loop); otherwise it falls back to ``forward`` per step."""
        K = self.loc_conv.weight.shape[-1]
        return bool(_lib.load().re2e_attloc_loop_supported(int(steps), int(batch), int(h_length), self.eprojs,
                                                           self.att_dim, self.aconv_chans, K))

    def forward_loop(self, enc_hs_pad, enc_hs_len, dec_z_all, first_none=True, att_prev=None, scaling=2.0):
        """The whole attention loop of Decoder.forward (model/e2e_decoder.py:114-122) for decoder states that are all
        known up front: step i computes ``att_c_i, att_w_i = forward(enc_hs_pad, enc_hs_len, z_i, att_w_{i-1})``.

        :param dec_z_all: (S-1, B, dunits) with ``first_none`` (step 0 gets dec_z = None, i.e. zeros), else (S, B, dunits)
        :param att_prev: alignment fed to step 0 (None = uniform over each length, e2e_attention.py:264-268)
        :return: c_all (S, B, eprojs), w_all (S, B, T_max)
        One persistent cluster kernel per direction when the shape fits (see csrc/attloc_loop.cu); identical results
        (to rounding) from the per-step kernels otherwise.
        """
        if self.aact_fuc != 'softmax':
            raise NotImplementedError("AttLoc sm_100a kernels implement aact_fuc='softmax' only")
        dev = self.mlp_enc.weight.device
        if dev.type != 'cuda':
            raise RuntimeError("AttLoc parameters must live on a CUDA device (no CPU fallback)")
        batch, h_length = enc_hs_pad.size(0), enc_hs_pad.size(1)
        steps = dec_z_all.size(0) + (1 if first_none else 0)
        if not self.loop_supported(steps, batch, h_length):
            self.reset()
            cs, ws, w = [], [], att_prev
            for i in range(steps):
                z = None if (first_none and i == 0) else dec_z_all[i - (1 if first_none else 0)]
                c, w = self.forward(enc_hs_pad, enc_hs_len, z, w, scaling)
                cs.append(c)
                ws.append(w)
            return torch.stack(cs), torch.stack(ws)
        if att_prev is None:
            L = _lib.lib()
            if torch.is_tensor(enc_hs_len) and enc_hs_len.is_cuda:
                hl = enc_hs_len.to(torch.int32).contiguous()
            else:
                hl = torch.from_numpy(np.fromiter((int(l) for l in enc_hs_len), dtype=np.int32)).to(dev, non_blocking=True)
            att_prev = torch.empty(batch, h_length, device=dev, dtype=torch.float32)
            with _lib.on(dev):
                _lib.check(L.re2e_attloc_init_att(_lib.ptr(hl), _lib.ptr(att_prev), batch, h_length, _lib.stream_ptr()),
                           "re2e_attloc_init_att")
        return _Loop.apply(enc_hs_pad, dec_z_all, self.mlp_enc.weight, self.mlp_enc.bias, self.mlp_dec.weight,
                           self.mlp_att.weight, self.loc_conv.weight, self.gvec.weight, self.gvec.bias, att_prev,
                           scaling, bool(first_none))

    def precompute(self, enc_hs_pad):
        """mlp_enc(enc_h) of the utterance batch, once per reset() (e2e_attention.py:252-256); returns the per-batch
        state (``pre``, ``enc``, contiguous weight views, dims) the step kernels read."""
        if self.pre_compute_enc_h is None:
            self._state = _State()
            self.enc_h = enc_hs_pad
            self.h_length = self.enc_h.size(1)
            self.pre_compute_enc_h, self._anchor = _Precompute.apply(
                enc_hs_pad, self.mlp_enc.weight, self.mlp_enc.bias, self.mlp_dec.weight, self.mlp_att.weight,
                self.loc_conv.weight, self.gvec.weight, self.gvec.bias, self._state)
        return self._state

    def forward(self, enc_hs_pad, enc_hs_len, dec_z, att_prev, scaling=2.0):
        '''AttLoc forward

        :param enc_hs_pad: padded encoder hidden state (B x T_max x D_enc)
        :param list enc_hs_len: encoder hidden state lengths (B)
        :param dec_z: decoder hidden state (B x D_dec) or None (= zeros)
        :param att_prev: previous attention weight (B x T_max) or None (= uniform over each length)
        :param float scaling: scaling parameter before applying softmax
        :return: attention weighted encoder state (B, D_enc), attention weights (B x T_max)
        '''
        if self.aact_fuc != 'softmax':
            raise NotImplementedError("AttLoc sm_100a kernels implement aact_fuc='softmax' only "
                                      "(the only mode reachable from E2E, model/e2e_model.py:71-72)")
        dev = self.mlp_enc.weight.device
        if dev.type != 'cuda':
            raise RuntimeError("AttLoc parameters must live on a CUDA device (no CPU fallback)")
        batch = len(enc_hs_pad)
        st = self.precompute(enc_hs_pad)
        is_init = att_prev is None
        if is_init:
            # e2e_attention.py:264-268: uniform over enc_hs_len[b], zero padded
            L = _lib.lib()
            if st.hlens_dev is None:
                if torch.is_tensor(enc_hs_len) and enc_hs_len.is_cuda:
                    st.hlens_dev = enc_hs_len.to(torch.int32).contiguous()
                else:
                    st.hlens_dev = torch.from_numpy(
                        np.fromiter((int(l) for l in enc_hs_len), dtype=np.int32)).to(dev, non_blocking=True)
            att_prev = torch.empty(batch, self.h_length, device=dev, dtype=torch.float32)
            with _lib.on(dev):
                _lib.check(L.re2e_attloc_init_att(_lib.ptr(st.hlens_dev), _lib.ptr(att_prev), batch,
                                                  self.h_length, _lib.stream_ptr()), "re2e_attloc_init_att")
        if dec_z is not None:
            dec_z = dec_z.view(batch, self.dunits)
        c, w = _Step.apply(self._anchor, dec_z, att_prev, st, scaling, is_init)
        return c, w
