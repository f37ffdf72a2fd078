"""Oracle (CPU restatement) of the hybrid CTC/attention beam search.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/model/e2e_decoder.py:170-369 (Decoder.recognize_beam, no LM / no fusion, dlayers = 1)
with E2E.recognize's CTC side (model/e2e_model.py:222-233): one hypothesis at a time, exactly the reference's
order of operations and tie-breaking (Python's stable sort), on top of the oracle AttLoc step
(oracle/attloc.py) and the oracle prefix scorer (oracle/ctc.py).  ``end_detect`` restates
model/e2e_common.py:226-254.

Pinned by tests/golden/beam.npz (generated from the unmodified reference by oracle/gen_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import attloc
from .ctc import CTCPrefixScoreOracle

CTC_SCORING_RATIO = 1.5   # model/e2e_decoder.py:20


def end_detect(ended_hyps, i, M=3, D_end=np.log(1 * np.exp(-10))):
    """model/e2e_common.py:226-254."""
    if len(ended_hyps) == 0:
        return False
    count = 0
    best_hyp = sorted(ended_hyps, key=lambda x: x['score'], reverse=True)[0]
    for m in range(M):
        hyp_length = i - m
        same = [x for x in ended_hyps if len(x['yseq']) == hyp_length]
        if len(same) > 0:
            best_same = sorted(same, key=lambda x: x['score'], reverse=True)[0]
            if best_same['score'] - best_hyp['score'] < D_end:
                count += 1
    return count == M


def lstm_cell(sd, prefix, x, hc):
    """torch.nn.LSTMCell arithmetic (gate order i, f, g, o)."""
    h, c = hc
    gates = F.linear(x, sd[prefix + "weight_ih"], sd[prefix + "bias_ih"]) + \
        F.linear(h, sd[prefix + "weight_hh"], sd[prefix + "bias_hh"])
    i, f, g, o = gates.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def recognize_beam(sd, h, cfg):
    """sd: state dict with the reference's key names (att.*, embed.weight, decoder.0.*, output.*, ctc_lo.*);
    h (Th, D) encoder output of one utterance; cfg: beam, penalty, ctc_weight, maxlenratio, minlenratio, nbest,
    sos, eos.  Returns the n-best list of {'yseq', 'score'}."""
    att_p = {k[len("att."):]: v for k, v in sd.items() if k.startswith("att.")}
    Th = h.shape[0]
    Z = sd["embed.weight"].shape[1]
    beam, penalty, ctc_weight = cfg["beam"], cfg["penalty"], cfg["ctc_weight"]
    sos, eos = cfg["sos"], cfg["eos"]
    hb = h.unsqueeze(0)
    pre = attloc.precompute(att_p, hb)
    lpz = None
    if ctc_weight > 0.0:   # model/e2e_model.py:222-226
        lpz = F.log_softmax(F.linear(h, sd["ctc_lo.weight"], sd["ctc_lo.bias"]), dim=1)
    maxlen = Th if cfg["maxlenratio"] == 0 else max(1, int(cfg["maxlenratio"] * Th))
    minlen = int(cfg["minlenratio"] * Th)
    zero = h.new_zeros(1, Z)
    hyp = {'score': 0.0, 'yseq': [sos], 'c_prev': zero, 'z_prev': zero, 'a_prev': None}
    if lpz is not None:
        scorer = CTCPrefixScoreOracle(lpz.numpy(), 0, eos)
        hyp['ctc_state_prev'] = scorer.initial_state()
        hyp['ctc_score_prev'] = 0.0
        ctc_beam = min(lpz.shape[-1], int(beam * CTC_SCORING_RATIO)) if ctc_weight != 1.0 else lpz.shape[-1]
    hyps, ended = [hyp], []
    for i in range(maxlen):
        kept = []
        for hyp in hyps:
            ey = sd["embed.weight"][hyp['yseq'][i]].view(1, -1)
            att_c, att_w = attloc.step(att_p, hb, pre, [Th], hyp['z_prev'], hyp['a_prev'])
            z, c = lstm_cell(sd, "decoder.0.", torch.cat((ey, att_c), dim=1), (hyp['z_prev'], hyp['c_prev']))
            local_att = F.log_softmax(F.linear(z, sd["output.weight"], sd["output.bias"]), dim=1)
            if lpz is not None:
                _, ids = torch.topk(local_att, ctc_beam, dim=1)
                ctc_scores, ctc_states, _ = scorer(hyp['yseq'], ids[0].numpy(), hyp['ctc_state_prev'])
                local = (1.0 - ctc_weight) * local_att[:, ids[0]] + \
                    ctc_weight * torch.from_numpy(ctc_scores - hyp['ctc_score_prev'])
                best_scores, joint = torch.topk(local, beam, dim=1)
                best_ids = ids[:, joint[0]]
            else:
                best_scores, best_ids = torch.topk(local_att, beam, dim=1)
            for j in range(beam):
                new = {'z_prev': z, 'c_prev': c, 'a_prev': att_w, 'score': hyp['score'] + best_scores[0, j],
                       'yseq': list(hyp['yseq']) + [int(best_ids[0, j])]}
                if lpz is not None:
                    new['ctc_state_prev'] = ctc_states[joint[0, j]]
                    new['ctc_score_prev'] = ctc_scores[joint[0, j]]
                kept.append(new)
            kept = sorted(kept, key=lambda x: x['score'], reverse=True)[:beam]
        hyps = kept
        if i == maxlen - 1:
            for hyp in hyps:
                hyp['yseq'].append(eos)
        remained = []
        for hyp in hyps:
            if hyp['yseq'][-1] == eos:
                if len(hyp['yseq']) > minlen:
                    hyp['score'] += (i + 1) * penalty
                    ended.append(hyp)
            else:
                remained.append(hyp)
        if end_detect(ended, i) and cfg["maxlenratio"] == 0.0:
            break
        hyps = remained
        if len(hyps) == 0:
            break
    nbest = sorted(ended, key=lambda x: x['score'], reverse=True)[:min(len(ended), cfg["nbest"])]
    return [{'yseq': [int(t) for t in x['yseq']], 'score': float(x['score'])} for x in nbest]


def decoder_forward(sd, hpad, hlen, ys, sos, eos, ignore_id=-1):
    """Decoder.forward (model/e2e_decoder.py:79-167) with teacher forcing (scheduled_sampling_rate = 0), dlayers = 1,
    no label smoothing: returns (loss, acc).  hpad (B,Th,D) may require grad; sd values may require grad."""
    att_p = {k[len("att."):]: v for k, v in sd.items() if k.startswith("att.")}
    hlen = [int(l) for l in hlen]
    mask = torch.zeros_like(hpad)
    for b, l in enumerate(hlen):
        mask[b, :l] = 1.0
    hpad = hpad * mask                                     # mask_by_length(hpad, hlen, 0)
    B = hpad.shape[0]
    Z = sd["embed.weight"].shape[1]
    ys_in = [torch.cat([y.new_tensor([sos]), y]) for y in ys]
    ys_out = [torch.cat([y, y.new_tensor([eos])]) for y in ys]
    L = max(len(y) for y in ys_in)
    pad_in = torch.full((B, L), eos, dtype=torch.long)
    pad_out = torch.full((B, L), ignore_id, dtype=torch.long)
    for b in range(B):
        pad_in[b, :len(ys_in[b])] = ys_in[b]
        pad_out[b, :len(ys_out[b])] = ys_out[b]
    z = hpad.new_zeros(B, Z)
    c = hpad.new_zeros(B, Z)
    pre = attloc.precompute(att_p, hpad)
    att_w = None
    eys = sd["embed.weight"][pad_in]
    y_all = []
    for i in range(L):
        att_c, att_w = attloc.step(att_p, hpad, pre, hlen, z, att_w)
        z, c = lstm_cell(sd, "decoder.0.", torch.cat((eys[:, i, :], att_c), dim=1), (z, c))
        y_all.append(F.linear(z, sd["output.weight"], sd["output.bias"]))
    y_all = torch.stack(y_all, dim=0).transpose(0, 1).contiguous().view(B * L, -1)
    loss = F.cross_entropy(y_all, pad_out.view(-1), ignore_index=ignore_id, reduction='mean')
    loss = loss * (np.mean([len(x) for x in ys_in]) - 1)
    pred = y_all.detach().view(B, L, -1).argmax(2)
    m = pad_out != ignore_id
    acc = float((pred[m] == pad_out[m]).sum()) / float(m.sum())
    return loss, acc
