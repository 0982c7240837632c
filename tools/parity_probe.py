"""Gradient-precision probe (GPU box): runs the reference's golden train step (B = 4) and a B = 16 oracle step with
fp16 (loss-scaled) and with bf16 gradient tensors, prints per-mode loss / gradient errors against the reference and
the dynamic range of every gradient tensor (so the loss scale can be chosen with evidence).
    python tools/parity_probe.py [--scales 256,1024,4096] > profiles/r2_parity_probe.txt"""
import argparse
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import segan_oracle as O                                   # noqa: E402
from segan_pytorch_b200 import engine as E                             # noqa: E402
from tests.util import build_segan, cpu_state, golden, load_opts, rel_err  # noqa: E402

DEV = "cuda"


def grad_ranges(s):
    out = []
    for eng, tag in ((s.D.engine, "D"), (s.G.engine, "G")):
        for name, t in sorted(eng.buf.t.items()):
            short = name.split(".")[-1]
            if t.dtype != E.GT or not (short.startswith(("ga", "gh", "gin", "gad", "gz", "colg", "P2")) or short.startswith("ghp")):
                continue
            x = t.float().abs()
            nz = x[x > 0]
            if nz.numel() == 0:
                continue
            amax = float(nz.max())
            med = float(nz.median())
            sub = float((nz < 6.1e-5).float().mean()) if E.GT == torch.float16 else 0.0
            out.append("%s %-10s amax %.3e median %.3e  (unscaled amax %.3e)  fp16-subnormal %.1f%%"
                       % (tag, name, amax, med, amax / E.LOSS_SCALE, 100 * sub))
    return out


def golden_step():
    t = golden("train_step_b4.npz")
    B = t["clean"].shape[0]
    s = build_segan(batch_size=B).to(DEV)
    s.G.train()
    s.D.train()
    Gopt, Dopt = s.build_optimizers(load_opts(batch_size=B))
    random.seed(int(t["py_random_seed"]))
    torch.manual_seed(int(t["torch_seed_z"]))
    clean = torch.from_numpy(t["clean"]).unsqueeze(1).to(DEV)
    noisy = torch.from_numpy(t["noisy"]).unsqueeze(1).to(DEV)
    lv = s.train_step(clean, noisy, Gopt, Dopt, 100.0).tolist()
    torch.cuda.synchronize()
    ref = [float(t[k]) for k in ("d_real_loss", "d_fake_loss", "g_adv_loss", "g_l1_loss")]
    worst = {}
    for tag, eng in (("gD.", s.D.engine), ("gG.", s.G.engine)):
        for name, p in eng.module.named_parameters():
            if tag == "gD." and name.startswith("enc_blocks") and name.endswith("conv.bias"):
                continue
            idx = torch.from_numpy(t["idx." + tag + name])
            r = torch.from_numpy(t["val." + tag + name])
            g = eng.grad_of(name).detach().float().cpu().reshape(-1)[idx]
            rms = float(t["norm." + tag + name]) / max(1.0, p.numel()) ** 0.5
            worst[tag + name] = float((g - r).norm()) / (float(r.norm()) + rms + 1e-12)
    return lv, ref, worst, s


def oracle_step(B):
    s = build_segan(batch_size=B)
    sdG, sdD = cpu_state(s.G), cpu_state(s.D)
    s = s.to(DEV)
    s.G.train()
    s.D.train()
    g = torch.Generator().manual_seed(113)
    clean = (0.3 * torch.randn(B, 1, 16384, generator=g)).clamp(-1, 1)
    noisy = (clean + 0.1 * torch.randn(B, 1, 16384, generator=g)).clamp(-1, 1)
    z = torch.randn(B, 1024, 16, generator=g)
    Gopt, Dopt = s.build_optimizers(load_opts(batch_size=B))
    random.seed(3)
    shifts3 = [O.draw_phase_shifts(5, 5) for _ in range(3)]
    lv = s.train_step(clean.to(DEV), noisy.to(DEV), Gopt, Dopt, 100.0, z=z.to(DEV), shifts3=shifts3).tolist()
    gD = {k: s.D.engine.grad_of(k).cpu() for k, _ in s.D.named_parameters()}
    gG = {k: s.G.engine.grad_of(k).cpu() for k, _ in s.G.named_parameters()}
    return lv, gD, gG, (sdG, sdD, clean, noisy, z, shifts3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scales", default="256,1024,4096")
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args()
    modes = [("bf16", 1.0)] + [("f16", float(v)) for v in a.scales.split(",")]
    ref_cache = None
    for kind, scale in modes:
        E.set_grad_dtype(kind, scale)
        print("=" * 100)
        print("gradient tensors: %s, loss scale %g" % (kind, scale))
        lv, ref, worst, s = golden_step()
        print("golden step (B=4): losses", ["%.5f" % v for v in lv], "reference", ["%.5f" % v for v in ref])
        wD = {k: v for k, v in worst.items() if k.startswith("gD.")}
        wG = {k: v for k, v in worst.items() if k.startswith("gG.")}
        print("  D-step grads rel-L2 (256 sampled entries): max %.3e (%s) median %.3e"
              % (max(wD.values()), max(wD, key=wD.get), float(np.median(list(wD.values())))))
        print("  G grads through the updated D:             max %.3e (%s) median %.3e"
              % (max(wG.values()), max(wG, key=wG.get), float(np.median(list(wG.values())))))
        for line in grad_ranges(s):
            print("   ", line)
        del s
        lv, gD, gG, inputs = oracle_step(a.batch)
        if ref_cache is None:
            sdG, sdD, clean, noisy, z, shifts3 = inputs
            sqG = {k: torch.zeros_like(sdG[k]) for k in O._trainable(sdG)}
            sqD = {k: torch.zeros_like(sdD[k]) for k in O._trainable(sdD)}
            ref_cache = O.segan_train_step(sdG, sdD, sqG, sqD, clean, noisy, z, shifts3, l1_weight=100.0)
        r = ref_cache
        refl = [r[k] for k in ("d_real_loss", "d_fake_loss", "g_adv_loss", "g_l1_loss")]
        eD = {k: rel_err(gD[k], g) for k, g in r["gradsD"].items() if not (k.startswith("enc_blocks") and k.endswith("conv.bias"))}
        eG = {k: rel_err(gG[k], g) for k, g in r["gradsG"].items()}
        print("oracle step (B=%d): losses %s oracle %s" % (a.batch, ["%.5f" % v for v in lv], ["%.5f" % v for v in refl]))
        print("  D grads full-tensor rel-L2: max %.3e (%s) median %.3e" % (max(eD.values()), max(eD, key=eD.get),
                                                                            float(np.median(list(eD.values())))))
        for k in sorted(eD, key=eD.get, reverse=True)[:6]:
            print("      %-36s %.3e" % (k, eD[k]))
        print("  G grads (through updated D) rel-L2: max %.3e (%s) median %.3e" % (max(eG.values()), max(eG, key=eG.get),
                                                                                    float(np.median(list(eG.values())))))
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
