"""One eager Decoder.forward + backward at the bench shape (for an ncu launch list of the real decoder loop)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robust_e2e_gan_b200 import AttLoc, Decoder, synth
from robust_e2e_gan_b200.hotpath import DEFAULT_CFG
cfg = dict(DEFAULT_CFG)
dev = torch.device("cuda:0")
B, Th, D, A, Z, C, V, U = (cfg[k] for k in ("B", "Th", "D", "A", "Z", "C", "V", "U"))
torch.manual_seed(7)
att = AttLoc(D, Z, A, C, cfg["filts"], "softmax")
dec = Decoder(D, V, 1, Z, V - 1, V - 1, att).to(dev).train()
hpad, hl = synth.encoder_batch(B=B, Th=Th, D=D, seed=77)
ys = [y.to(dev) for y in synth.targets(B=B, V=V, hlens=hl, seed=77, fixed_U=U)]
hpad = hpad.to(dev).requires_grad_(True)
for it in range(3):
    dec.zero_grad(); hpad.grad = None
    if it == 2:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    loss, acc = dec(hpad, hl, ys, 0.0)
    loss.backward()
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
