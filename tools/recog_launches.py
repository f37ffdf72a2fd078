"""One beam-search utterance of the config-5 workload inside a cudaProfilerStart/Stop range (for an ncu launch list),
plus the wall time per output position without the profiler.
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file OUT python tools/recog_launches.py"""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robust_e2e_gan_b200 import CTC, AttLoc, Decoder, synth

dev = torch.device("cuda:0")
synth.BEAM_CASES["_recog"] = dict(V=4233, D=320, Z=300, A=320, C=10, filts=100, Th=200, beam=10, ctc_weight=0.3,
                                  nbest=1, penalty=0.0, maxlenratio=0.0, minlenratio=0.0, eos_bias=2.0, seed=5000)
c, sd, _, _ = synth.beam_case("_recog")
att = AttLoc(c["D"], c["Z"], c["A"], c["C"], c["filts"], "softmax")
dec = Decoder(c["D"], c["V"], 1, c["Z"], c["sos"], c["eos"], att)
ctc = CTC(c["V"], c["D"], 0.0)
dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("ctc_lo")})
ctc.load_state_dict({k: v for k, v in sd.items() if k.startswith("ctc_lo")})
dec, ctc = dec.to(dev).eval(), ctc.to(dev).eval()
ra = types.SimpleNamespace(cuda_graph=os.environ.get('RE2E_RECOG_GRAPH', '1') == '1', beam_size=10, penalty=0.0, ctc_weight=0.3, maxlenratio=0.0, minlenratio=0.0, nbest=1, lm_weight=0.0)
g = torch.Generator().manual_seed(5001)
Th = int(sys.argv[1]) if len(sys.argv) > 1 else 137
hs = [torch.tanh(torch.randn(Th, c["D"], generator=g)).to(dev) for _ in range(4)]


def decode(h):
    lpz = ctc.log_softmax(h.unsqueeze(0))[0]
    return dec.recognize_beam(h, lpz, ra, None)


with torch.no_grad():
    decode(hs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for h in hs[1:3]:
        n += len(decode(h)[0]["yseq"]) - 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ra.profile = {}
    for h in hs[1:3]:
        decode(h)
    torch.cuda.synchronize()
    print("host seconds per phase over the same two utterances:", {k: round(v, 5) for k, v in ra.profile.items()})
    del ra.profile
    print("wall: %.1f us per output position (%d positions, Th=%d)" % (dt * 1e6 / n, n, Th), flush=True)
    hs8 = [torch.tanh(torch.randn(Th, c["D"], generator=g)).to(dev) for _ in range(16)]
    lp8 = [ctc.log_softmax(h.unsqueeze(0))[0] for h in hs8]
    for conc in (1, 2, 4, 6):
        dec.recognize_beam_batch(hs8[:conc], lp8[:conc], ra, None, concurrency=conc)
        torch.cuda.synchronize()
        ra.profile = {}
        t0 = time.perf_counter()
        out = dec.recognize_beam_batch(hs8, lp8, ra, None, concurrency=conc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("batch of 16, %d concurrent searches: %.2f ms per utterance (%.1f utt/s); host ms per utterance by phase: %s"
              % (conc, dt * 1e3 / 16, 16 / dt, {k: round(v * 1e3 / 16, 3) for k, v in ra.profile.items() if k != "positions"}),
              flush=True)
        del ra.profile
    torch.cuda.cudart().cudaProfilerStart()
    decode(hs[3])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
