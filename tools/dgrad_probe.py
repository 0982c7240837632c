"""Discriminator gradient probe (GPU box): where does the D-step gradient error come from when the activation
derivative is continuous (PReLU slopes = 1)?  Runs single D passes and the two-pass D step through the engine API at
several loss scales, with and without the side-stream schedule, against the oracle; prints logit errors, per-tensor
gradient errors and the gradient tensors' dynamic range.
    PYTHONPATH=. python tools/dgrad_probe.py > profiles/r2_dgrad_probe.txt"""
import ctypes as C
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import segan_oracle as O                                   # noqa: E402
from segan_pytorch_b200 import engine as E                             # noqa: E402
from tests.util import build_segan, cpu_state, rel_err                 # noqa: E402

DEV = "cuda"
SKIP = lambda k: k.startswith("enc_blocks") and (k.endswith("conv.bias") or (k.endswith("norm.bias") and not k.startswith("enc_blocks.4")))


def setup(slope, B, seed=117):
    s = build_segan(batch_size=B)
    with torch.no_grad():
        for net in (s.G, s.D):
            for n, p in net.named_parameters():
                if n.endswith("act.weight"):
                    p.fill_(slope)
    sdD = cpu_state(s.D)
    g = torch.Generator().manual_seed(seed)
    clean = (0.3 * torch.randn(B, 1, 16384, generator=g)).clamp(-1, 1)
    noisy = (clean + 0.1 * torch.randn(B, 1, 16384, generator=g)).clamp(-1, 1)
    fake = (clean + 0.05 * torch.randn(B, 1, 16384, generator=g)).clamp(-1, 1)
    return s.to(DEV), sdD, clean, noisy, fake


def oracle_grads(sdD, passes, shifts):
    pD = {k: sdD[k].clone().requires_grad_(True) for k in O._trainable(sdD)}
    with O.oracle_mode():
        loss, logits = 0, []
        for (x0, x1, tgt), sh in zip(passes, shifts):
            l = O.discriminator_forward({**{k: v.clone() for k, v in sdD.items()}, **pD}, torch.cat((x0, x1), 1), sh)
            logits.append(l.detach())
            loss = loss + torch.nn.functional.mse_loss(l.view(-1), torch.full((x0.shape[0],), tgt))
        g = dict(zip(pD.keys(), torch.autograd.grad(loss, list(pD.values()))))
    return float(loss), logits, g


def engine_grads(s, passes, shifts, lanes):
    de = s.D.engine
    de.bind()
    de.grad.zero_()
    de._alpha_fixed = False
    losses = torch.zeros(len(passes), device=DEV)
    logits = []
    streams = [E.side_stream(DEV, 2), None] if lanes else [None, None]
    for i, ((x0, x1, tgt), sh) in enumerate(zip(passes, shifts)):
        side = streams[i % 2]
        with E.on_side(side):
            lg, cx = de.forward(x0.to(DEV), x1.to(DEV), sh, training=True, lane=(1 if side is not None else 0))
            de.backward(cx, tgt, 1.0, param_grads=True, loss_out=C.c_void_p(losses.data_ptr() + 4 * i))
            logits.append(lg)
    for st in streams:
        E.join_side(st)
    torch.cuda.synchronize()
    g = {k: de.grad_of(k).cpu() for k, _ in s.D.named_parameters()}
    rng = []
    for name, t in sorted(de.buf.t.items()):
        short = name.split(".")[-1]
        if t.dtype == E.GT and short.startswith(("ga", "gh", "gz")):
            rng.append((name, float(t.float().abs().max())))
    return float(losses.sum()), [l.cpu() for l in logits], g, rng


def report(tag, g, go, lg, lo, loss, losso, rng):
    e = {k: rel_err(g[k], v) for k, v in go.items() if not SKIP(k)}
    top = sorted(e.items(), key=lambda kv: -kv[1])[:4]
    le = max(float((a - b).abs().max()) for a, b in zip(lg, lo))
    print("%-46s loss %.5f/%.5f logits max-abs %.2e | grads median %.3e max %.3e  top %s | amax %s"
          % (tag, loss, losso, le, float(np.median(list(e.values()))), max(e.values()),
             ["%s %.1e" % (k, v) for k, v in top], ["%s %.0f" % (n.split(".")[-1], v) for n, v in rng[:6]]))


def main():
    B = 8
    random.seed(3)
    shifts = [O.draw_phase_shifts(5, 5) for _ in range(2)]
    for slope in (1.0, 0.0):
        for scale in (64.0, 1024.0):
            E.set_grad_dtype("f16", scale)
            s, sdD, clean, noisy, fake = setup(slope, B)
            real = (clean, noisy, 1.0)
            fk = (fake, noisy, 0.0)
            for tag, passes, lanes in (("real only", [real], False), ("fake only", [fk], False),
                                       ("real+fake serial", [real, fk], False), ("real+fake lanes", [real, fk], True)):
                losso, lo, go = oracle_grads(sdD, passes, shifts)
                loss, lg, g, rng = engine_grads(s, passes, shifts, lanes)
                report("slope %g scale %g %s" % (slope, scale, tag), g, go, lg, lo, loss, losso, rng)
            del s
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
