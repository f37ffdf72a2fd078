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
    # What is HOST-resident in the reference flow (the collated batch of data/mix_data_loader.py:264-302: spectra and
    # lengths; labels are staged separately) ...
    HOST_FIELDS = ("mix", "clean", "lens", "hlens")
    # ... and what the networks upstream / downstream of the path produce ON THE DEVICE in the reference flow
    # (enhancement net output, model/enhance_model.py:131-156; encoder output; LSTMCell states; gradients arriving
    # from the losses; the CMVN constants of the model): stand-ins here, they never cross PCIe in the real job.
    UPSTREAM_FIELDS = ("mask_logits", "hpad", "dec_z", "g_feat", "g_c", "g_w", "cmvn")

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

    def __init__(self, cfg, seed=1234, mtlalpha=0.5, overlap=True, fused_loop=True):
        super().__init__()
        self.cfg = dict(cfg)
        # fused_loop=True: the attention decoder loop runs as ONE persistent cluster kernel per direction
        # (AttLoc.forward_loop; the decoder states of all steps are inputs of the step, so mlp_dec of every step is one
        # dense product).  False: one AttLoc.forward per output position, as Decoder.forward calls it (A/B timing).
        self.fused_loop = fused_loop
        self.joint_frontend = True     # the three front-end calls of joint_train.py:158-161 as one launch
        # overlap=True: the three independent branches of the step (front-end, CTC, attention decoder loop) run on
        # three CUDA streams, fork/join inside step(); autograd replays each branch's backward on its own stream.
        # The decoder loop is a serial chain of latency-bound cluster kernels that leaves SMs and issue slots idle;
        # the bandwidth/tensor-bound front-end and CTC kernels fill them.
        self.overlap = overlap
        self._side = None
        # optional callable(out) invoked at the end of step() after the backward pass, on the caller's stream -- inside
        # a CUDA-graph capture of the step when there is one.  A data-parallel job joins its gradient all-reduces here.
        self.after_backward = None
        if overlap and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            # the leaves (mask logits, encoder output) are created on the caller's stream while two of the three branches
            # produce their gradients on side streams: intentional, the engine inserts the needed waits
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
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

    def _branch_streams(self, dev):
        """(front-end stream, CTC stream): low priority, so the decoder loop's kernels on the caller's stream are
        scheduled first whenever both have CTAs pending."""
        if self._side is None or self._side[0].device != dev:
            self._side = (torch.cuda.Stream(dev, priority=0), torch.cuda.Stream(dev, priority=0))
        return self._side

    def step(self, b, backward=True, hlens_for_att=None):
        """One fwd+bwd of the hot path on a device Batch.  Returns a dict of outputs and gradients."""
        steps = self.cfg["steps"]
        mask_logits = b.mask_logits.detach().requires_grad_(backward)
        hpad = b.hpad.detach().requires_grad_(backward)
        fused = self.fused_loop and steps > 1
        if fused:
            dec_z_all = b.dec_z[:steps - 1].detach().requires_grad_(backward)
        else:   # one leaf per decoder step (as the LSTMCell outputs are separate tensors in Decoder.forward)
            dec_zs = [b.dec_z[i].detach().requires_grad_(backward) for i in range(steps - 1)]
        main = torch.cuda.current_stream(hpad.device)
        if self.overlap:
            s_fe, s_ctc = self._branch_streams(hpad.device)
            s_fe.wait_stream(main)
            s_ctc.wait_stream(main)
        else:
            s_fe = s_ctc = main
        # -- front-end (joint_train.py:158-161)
        with torch.cuda.stream(s_fe):
            if self.joint_frontend:      # one launch: mask, mix and clean each read once
                enhance_feat, mix_feat, clean_feat = self.feat.forward_joint(mask_logits, b.mix, b.clean, b.lens, b.cmvn)
            else:
                enhance_feat = self.feat.forward_masked(mask_logits, b.mix, b.lens, b.cmvn)
                with torch.no_grad():
                    clean_feat = self.feat(b.clean, b.cmvn)
                    mix_feat = self.feat(b.mix, b.cmvn)
        # -- CTC branch (model/e2e_model.py:192)
        with torch.cuda.stream(s_ctc):
            loss_ctc = self.ctc(hpad, b.hlens, b.targets if b.targets is not None else b.ys)
        # -- attention decoder loop (model/e2e_decoder.py:114-122)
        self.att.reset()
        hl = hlens_for_att if hlens_for_att is not None else b.hlens_list
        if fused:
            c_all, w_all = self.att.forward_loop(hpad, hl, dec_z_all, first_none=True)
            att_w, cs = w_all[steps - 1], [c_all]
        else:
            att_w = None
            cs = []
            for i in range(steps):
                z = None if i == 0 else dec_zs[i - 1]
                att_c, att_w = self.att(hpad, hl, z, att_w)
                cs.append(att_c)
            c_all = torch.stack(cs)
        if self.overlap:
            main.wait_stream(s_fe)
            main.wait_stream(s_ctc)
        out = {"enhance_feat": enhance_feat, "clean_feat": clean_feat, "mix_feat": mix_feat, "loss_ctc": loss_ctc,
               "att_c": c_all, "att_w": att_w}
        if backward:
            outs = [enhance_feat, loss_ctc, att_w] + cs
            grads = [b.g_feat, torch.full_like(loss_ctc, self.mtlalpha), b.g_w] + \
                ([b.g_c[:steps]] if fused else [b.g_c[i] for i in range(steps)])
            torch.autograd.backward(outs, grads)
            out.update(d_mask_logits=mask_logits.grad, d_hpad=hpad.grad,
                       d_dec_z=dec_z_all.grad if fused else [z.grad for z in dec_zs])
            for k, p in self.named_parameters():
                if p.requires_grad and p.grad is not None:
                    out["d_" + k] = p.grad
            if self.after_backward is not None:
                self.after_backward(out)
        return out


def _targets_host(ys, ignore_id=-1):
    """Flat int32 labels / offsets / lengths on the host (numpy), as prepare_targets builds them."""
    seqs = [np.asarray(y.detach().cpu().numpy() if torch.is_tensor(y) else y).reshape(-1) for y in ys]
    seqs = [s[s != ignore_id] for s in seqs]
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    offs = np.zeros(len(seqs), dtype=np.int32)
    if len(seqs) > 1:
        offs[1:] = np.cumsum(lens[:-1])
    flat = np.concatenate(seqs).astype(np.int32) if lens.sum() else np.zeros(1, np.int32)
    return flat, offs, lens


class StepRunner(object):
    """The hot-path step as a user drives it from host batches: ``runner(host_batch) -> outputs``.

    The whole fwd+bwd (``HotPath.step``) is captured once per input slot into a CUDA graph over static
    device buffers.  ``submit(host_batch)`` copies a (pinned) host batch into the free slot on a copy
    stream -- overlapping the previous step's kernels -- and enqueues the replay; ``result()`` returns
    the oldest outstanding step's output dict (device tensors, valid until that slot is reused two
    submits later).  The host batch's own (pinned) tensors are read by asynchronous copies: keep them unchanged until
    the slot's ``copied`` event has completed (``result()`` of that step implies it).  ``__call__`` = submit + result.  Outputs of EAGER steps taken on the default stream must not be
    kept alive across the construction: a live autograd graph pins the parameters' gradient accumulators to the stream
    it ran on, and the capture then fails with cudaErrorStreamCaptureImplicit.  Shapes are fixed at construction (the reference's
    bucketing sampler yields fixed-shape batches, data/mix_data_loader.py:314-346); label sequences may
    vary up to ``umax`` labels each.
    """

    def __init__(self, hp, example_host_batch, slots=2, umax=None, upstream="host"):
        """upstream="host": every field of the host batch is copied in per step (the stand-ins included -- what the
        parity tests feed).  upstream="device": only what is host-resident in the reference flow is copied per step
        (``Batch.HOST_FIELDS`` + labels); the stand-ins for tensors that upstream networks produce on the device
        (``Batch.UPSTREAM_FIELDS``) stay resident in each slot as set at construction / by ``set_upstream``."""
        if upstream not in ("host", "device"):
            raise ValueError("upstream must be 'host' or 'device'")
        self.copy_fields = Batch.FIELDS if upstream == "host" else Batch.HOST_FIELDS
        self.hp = hp
        self.dev = next(hp.parameters()).device
        self.copy_stream = torch.cuda.Stream(self.dev)
        # high priority: the decoder loop (a serial chain of latency-bound kernels on this stream) gets SMs before the
        # bulk front-end / CTC kernels that HotPath.step forks onto its (default-priority) branch streams
        self.run_stream = torch.cuda.Stream(self.dev, priority=-1)
        B = hp.cfg["B"]
        flat, offs, lens = _targets_host(example_host_batch.ys)
        self.umax = int(umax if umax is not None else max(int(lens.max()) if len(lens) else 1, 1))
        self.slots = []
        self._pending = []
        self._next = 0
        self.post = None          # optional callable(out) run on the step's stream right after the replay
                                  # (e.g. the gradient all-reduce of a data-parallel job)
        self._lab_pin = [torch.empty(B * self.umax + 2 * B, dtype=torch.int32).pin_memory() for _ in range(slots)]
        for s in range(slots):
            db = example_host_batch.to(self.dev, non_blocking=False)
            labels = torch.zeros(B * self.umax + 2 * B, device=self.dev, dtype=torch.int32)
            from .e2e_ctc import PreparedTargets
            db.targets = PreparedTargets(labels[:B * self.umax], labels[B * self.umax:B * self.umax + B],
                                         labels[B * self.umax + B:], self.umax, B)
            db._labels_all = labels
            self.slots.append({"batch": db, "graph": None, "out": None,
                               "copied": torch.cuda.Event(), "done": torch.cuda.Event()})
        torch.cuda.synchronize(self.dev)
        for s in range(slots):
            self._stage(s, example_host_batch)
        torch.cuda.synchronize(self.dev)
        self._capture()

    # -- graph capture -----------------------------------------------------------------------------
    def _drop_refs(self):
        import gc
        self.hp.att.reset()
        self.hp.ctc.loss = None
        self.hp.ctc.nll = None
        for p in self.hp.parameters():
            p.grad = None
        gc.collect()

    def _capture(self):
        hp, st = self.hp, self.run_stream
        self._drop_refs()
        st.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(st):
            for _ in range(2):                                  # warm-up: sizes the caching allocator, smem attrs
                hp.step(self.slots[0]["batch"], hlens_for_att=self.slots[0]["batch"].hlens)
                self._drop_refs()
        torch.cuda.synchronize(self.dev)
        for slot in self.slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                out = hp.step(slot["batch"], hlens_for_att=slot["batch"].hlens)
            slot["graph"], slot["out"] = g, out
            self._drop_refs()
        torch.cuda.synchronize(self.dev)

    # -- per step ----------------------------------------------------------------------------------
    def _stage(self, s, hb):
        """H2D of one host batch into slot s (on the copy stream)."""
        slot = self.slots[s]
        db = slot["batch"]
        B = self.hp.cfg["B"]
        flat, offs, lens = _targets_host(hb.ys)
        if len(lens) != B or (len(lens) and int(lens.max()) > self.umax):
            raise ValueError("StepRunner: batch has %d utterances / %d labels max, captured for %d / %d"
                             % (len(lens), int(lens.max()) if len(lens) else 0, B, self.umax))
        pin = self._lab_pin[s]
        # the previous asynchronous H2D copy out of this slot's pinned label block must have finished before the host
        # rewrites the block (a loop that never reads a result back can run several submits ahead of the device)
        slot["copied"].synchronize()
        pin.zero_()
        pin[:len(flat)] = torch.from_numpy(flat)
        pin[B * self.umax:B * self.umax + B] = torch.from_numpy(offs)
        pin[B * self.umax + B:] = torch.from_numpy(lens)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["done"])           # the slot's previous replay has finished
            for k in self.copy_fields:
                getattr(db, k).copy_(getattr(hb, k), non_blocking=True)
            db._labels_all.copy_(pin, non_blocking=True)
            slot["copied"].record(self.copy_stream)

    def h2d_bytes(self, hb):
        return int(sum(self.h2d_tensors(hb).values()))

    def h2d_tensors(self, hb):
        """Bytes per tensor of what ``submit`` copies host -> device for one step."""
        d = {k: int(getattr(hb, k).numel() * getattr(hb, k).element_size()) for k in self.copy_fields}
        d["labels+offsets+lengths"] = int(self._lab_pin[0].numel() * 4)
        return d

    def resident_tensors(self, hb):
        """Bytes per tensor of the device-resident stand-ins (not copied per step)."""
        return {k: int(getattr(hb, k).numel() * getattr(hb, k).element_size())
                for k in Batch.FIELDS if k not in self.copy_fields}

    def set_upstream(self, hb):
        """(Re)load the device-resident stand-ins of every slot from a host batch (outside the timed loop)."""
        torch.cuda.synchronize(self.dev)
        for slot in self.slots:
            for k in Batch.UPSTREAM_FIELDS:
                getattr(slot["batch"], k).copy_(getattr(hb, k))
        torch.cuda.synchronize(self.dev)

    def submit(self, host_batch):
        s = self._next
        self._next = (s + 1) % len(self.slots)
        if any(p == s for p in self._pending):
            raise RuntimeError("StepRunner: slot %d still has an unread result; call result() first" % s)
        self._stage(s, host_batch)
        slot = self.slots[s]
        with torch.cuda.stream(self.run_stream):
            self.run_stream.wait_event(slot["copied"])
            slot["graph"].replay()
            if self.post is not None:
                self.post(slot["out"])
            slot["done"].record(self.run_stream)
        self._pending.append(s)
        return s

    def replay_resident(self, s=0):
        """Replay slot s on its current device-resident inputs (no host copy); returns the output dict.
        Runs on ``run_stream``; the caller synchronises (events on that stream or device sync)."""
        slot = self.slots[s]
        with torch.cuda.stream(self.run_stream):
            slot["graph"].replay()
            if self.post is not None:
                self.post(slot["out"])
        return slot["out"]

    def close(self):
        """Drop the captured graphs and their static buffers.  A job whose step captured NCCL collectives must call
        this before ``destroy_process_group()``: a communicator cannot be torn down while a live CUDA graph still
        references its kernels."""
        torch.cuda.synchronize(self.dev)
        for slot in self.slots:
            slot["graph"] = None
            slot["out"] = None
        self._pending = []
        self._drop_refs()
        torch.cuda.synchronize(self.dev)

    def result(self):
        s = self._pending.pop(0)
        torch.cuda.current_stream(self.dev).wait_event(self.slots[s]["done"])
        return self.slots[s]["out"]

    def __call__(self, host_batch):
        self.submit(host_batch)
        return self.result()
