#!/usr/bin/env python
"""BASELINE config 5: hybrid CTC/attention beam search (beam 10, ctc_weight 0.3) over synthetic utterances.

    python tools/bench_recog.py [--utts 64] [--cpu-utts 2]            (torchrun for N > 1: utterances sharded by rank)

Model: default AttLoc / decoder / CTC dimensions (D = A = 320, Z = 300, C = 10, K = 201, V = 4233) with seeded
weights; encoder outputs h (Th x 320), Th ~ U(75, 200) (T/4 of 300..800 STFT frames), are stand-ins (the encoder is
outside the hot path).  Output length is bounded by maxlenratio 0.15 (AISHELL: ~14 characters per utterance).
GPU arm: robust_e2e_gan_b200.Decoder.recognize_beam (whole beam per launch, device-resident CTC prefix scores).
CPU arm: oracle/beam.py (the reference's one-hypothesis-at-a-time search) on a few utterances of the same set; the
n-best token sequences of both arms must be identical.  Prints one JSON line (secondary metric; the driver's
headline stays bench.py).
"""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=64)
    ap.add_argument("--cpu-utts", type=int, default=2)
    ap.add_argument("--beam", type=int, default=10)
    args = ap.parse_args()
    from robust_e2e_gan_b200 import synth as helpers
    from robust_e2e_gan_b200 import CTC, AttLoc, Decoder
    from robust_e2e_gan_b200.parallel import init_distributed, shard_range
    rank, world = init_distributed()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    helpers.BEAM_CASES["_recog"] = dict(V=4233, D=320, Z=300, A=320, C=10, filts=100, Th=200, beam=args.beam,
                                        ctc_weight=0.3, nbest=1, penalty=0.0, maxlenratio=0.15, minlenratio=0.0,
                                        eos_bias=2.0, seed=5000)
    c, sd, _, _ = helpers.beam_case("_recog")
    att = AttLoc(c["D"], c["Z"], c["A"], c["C"], c["filts"], "softmax")
    dec = Decoder(c["D"], c["V"], 1, c["Z"], c["sos"], c["eos"], att)
    ctc = CTC(c["V"], c["D"], 0.0)
    dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("ctc_lo")})
    ctc.load_state_dict({k: v for k, v in sd.items() if k.startswith("ctc_lo")})
    dec, ctc = dec.to(dev).eval(), ctc.to(dev).eval()
    g = torch.Generator().manual_seed(5001)
    lens = torch.randint(75, 201, (args.utts,), generator=g).tolist()
    hs = [torch.tanh(torch.randn(l, c["D"], generator=g)).pin_memory() for l in lens]
    ra = types.SimpleNamespace(beam_size=c["beam"], penalty=c["penalty"], ctc_weight=c["ctc_weight"],
                               maxlenratio=c["maxlenratio"], minlenratio=c["minlenratio"], nbest=c["nbest"], lm_weight=0.0)
    lo, hi = shard_range(args.utts, rank, world)

    def decode(i):
        h = hs[i].to(dev, non_blocking=True)
        lpz = ctc.log_softmax(h.unsqueeze(0))[0]
        return dec.recognize_beam(h, lpz, ra, None)

    decode(lo)                                   # warm-up (allocator, smem attributes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = {i: decode(i) for i in range(lo, hi)}
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    gpu_s = float(dt.item())
    line = {"metric": "utterances/sec (beam search, beam=%d, ctc_weight=0.3)" % c["beam"], "unit": "utt/s",
            "value": args.utts / gpu_s, "n_gpus": world, "utts": args.utts, "ms_per_utt": gpu_s / (hi - lo) * 1e3,
            "tokens_per_utt": sum(len(out[i][0]["yseq"]) - 1 for i in out) / max(1, len(out)),
            "data": "synthetic", "config": {"workload": "BASELINE configs[4]: beam search over synthetic utterances, "
                                            "Th~U(75,200), V=4233, D=A=320, Z=300, maxlenratio=0.15"}}
    if rank == 0 and args.cpu_utts > 0:
        from oracle import beam as obeam
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n = min(args.cpu_utts, hi - lo)
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = [obeam.recognize_beam(sd, hs[lo + k], c) for k in range(n)]
        cpu_s = time.perf_counter() - t0
        same = all(ref[k][0]["yseq"] == out[lo + k][0]["yseq"] for k in range(n))
        line["cpu_baseline"] = {"value": n / cpu_s, "unit": "utt/s", "cores": cores, "kind": "port",
                                "sample": "%d utterances of the same set" % n, "tokens_identical": bool(same)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
