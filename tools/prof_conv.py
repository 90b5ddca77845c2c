"""Tiny driver for ncu: a few launches of the section-8f converter kernels at config-3 size (1 024 000 rows, M = 24)."""
import sys, torch
sys.path.insert(0, ".")
import diffsptk_b200.functional as F
dev = torch.device("cuda", 0)
k = torch.empty(1024, 1000, 25, device=dev).uniform_(-0.6, 0.6)
k[..., 0] = k[..., 0].abs() + 0.5
with torch.no_grad():
    a = F.par2lpc(k)
    for _ in range(3):
        p = F.lpc2par(a)
        w = F.lpc2lsp(a)
        c = F.mgc2mgc(k * 0.3, 24, in_gamma=0.0, out_gamma=-0.5)
    torch.cuda.synchronize()
    from diffsptk_b200 import ops
    kk = k * 0.3
    for name, fn in (("lpc2par", lambda: F.lpc2par(a)), ("lpc2lsp", lambda: F.lpc2lsp(a)),
                     ("gc2gc 24->24 n=512", lambda: ops.gc2gc(kk, 24, 0.0, -0.5, 512)),
                     ("gnorm", lambda: F.gnorm(kk, -0.5))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"TIMING {name}: {e0.elapsed_time(e1) / 5:.3f} ms per 1 024 000 rows")
print(float(p.sum()), float(w.sum()), float(c.sum()))
