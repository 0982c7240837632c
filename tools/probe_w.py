"""Probe: which operand-format combinations does the MN-major tcgen05 weight-gradient kernel accept?"""
import subprocess
import sys

import torch

CASES = {"bf16_f16": (2, 1), "bf16_bf16": (2, 2), "f16_f16": (1, 1), "f16_bf16": (1, 2)}


def run(name):
    import torch.nn.functional as F
    from segan_pytorch_b200 import engine as E
    gd, ad = CASES[name]
    tdt = {1: torch.float16, 2: torch.bfloat16}
    g = torch.Generator().manual_seed(2)
    B, cin, cout, R, halo = 3, 64, 128, 128, 4
    kc, nc = 4 * cin, cout
    taps = E.tap_ranges("conv_fwd", cin, kc, nc)
    a0 = torch.randn(B, R + 2 * halo, kc, generator=g).to(tdt[ad]).cuda()
    gg = (torch.randn(B, R, nc, generator=g) * 0.1).to(tdt[gd]).cuda()
    dw = torch.zeros(9, nc, kc, dtype=torch.float32, device="cuda")
    E.run_w(gg, R, gd, a0, None, R, halo, ad, kc, nc, taps, dw, B, ksplit=3, backend=1)
    torch.cuda.synchronize()
    ap = F.pad(a0.float(), (0, 0, 16, 16))
    worst = 0.0
    for d in range(-4, 5):
        rows = ap[:, 16 + halo + d: 16 + halo + d + R, :]
        ref = torch.einsum("bmn,bmk->nk", gg.float(), rows)
        mask = torch.zeros(nc, kc, device="cuda")
        mask[taps[2][d + 4]:taps[3][d + 4], taps[0][d + 4]:taps[1][d + 4]] = 1
        worst = max(worst, float(((dw[d + 4] - ref) * mask).abs().max()) / float(ref.abs().max()))
    print(name, "OK rel-max-err %.3e" % worst)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for n in CASES:
            r = subprocess.run([sys.executable, __file__, n], capture_output=True, text=True)
            print(n, "rc", r.returncode, (r.stdout.strip().splitlines() or ["-"])[-1], (r.stderr.strip().splitlines() or ["-"])[-1][:160])
