"""Why does the overlapped step run at 16.2-16.8 ms right after start-up and at 14.9-15.2 ms a few seconds
later (tools/ab_step.py)?  Separates elapsed time under load from training progress:
  A: fresh networks, 5 warm-up steps, 10 timed steps
  B: fresh networks, 4 s of unrelated tensor load first (bf16 matmuls), then as A
  C: the networks of A after 200 more steps, 10 timed steps
  D: fresh networks again (no pre-load), as A -- the GPU has now been busy for a while
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from tests.util import build_segan, load_opts             # noqa: E402

B = 300
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
clean = (0.3 * torch.randn(B, 1, 16384, generator=g)).clamp_(-1, 1).to(dev)
noisy = (clean.cpu() + 0.1 * torch.randn(B, 1, 16384, generator=g)).clamp_(-1, 1).to(dev)


def fresh():
    s = build_segan(seed=111, batch_size=B, z_device="cuda").to(dev)
    s.G.train()
    s.D.train()
    Gopt, Dopt = s.build_optimizers(load_opts(batch_size=B, z_device="cuda"))
    return s, Gopt, Dopt


def timed(s, Gopt, Dopt, warm, n=10):
    losses = torch.zeros(4, device=dev)
    for _ in range(warm):
        s.train_step(clean, noisy, Gopt, Dopt, 100.0, losses=losses)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        s.train_step(clean, noisy, Gopt, Dopt, 100.0, losses=losses)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t0 = time.time()
s, Gopt, Dopt = fresh()
print("A fresh, 5 warm-up:            %.2f ms/step  (t=%.1fs)" % (timed(s, Gopt, Dopt, 5), time.time() - t0))
s2, G2, D2 = fresh()
x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
t1 = time.time()
while time.time() - t1 < 4.0:
    for _ in range(20):
        y = x @ x
    torch.cuda.synchronize()
print("B fresh after 4 s of matmuls:  %.2f ms/step  (t=%.1fs)" % (timed(s2, G2, D2, 5), time.time() - t0))
print("C networks of A, +200 steps:   %.2f ms/step  (t=%.1fs)" % (timed(s, Gopt, Dopt, 200), time.time() - t0))
s3, G3, D3 = fresh()
print("D fresh again, 5 warm-up:      %.2f ms/step  (t=%.1fs)" % (timed(s3, G3, D3, 5), time.time() - t0))
