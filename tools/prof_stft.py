"""Tiny driver for ncu: a few launches of the headline STFT at BASELINE config 2."""
import sys, torch
sys.path.insert(0, ".")
import diffsptk_b200 as D
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
x = torch.randn(B, 160000, device=dev)
m = D.STFT(400, 80, 512).to(dev)
with torch.no_grad():
    for _ in range(n):
        y = m(x)
torch.cuda.synchronize()
print(float(y.sum()))
