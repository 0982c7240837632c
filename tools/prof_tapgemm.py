"""Launches representative SEGAN+ layer shapes of the two tcgen05 tap-GEMMs at batch 300 (for ncu /
timing): conv fwd enc2 (Cin 128 -> 256), deconv fwd dec1 (1024 -> 256, two K sources), conv dgrad enc3,
wgrad enc2, wgrad dec1.  Prints CUDA-event times and TFLOP/s."""
import os
import sys

import torch

from segan_pytorch_b200 import engine as E
from segan_pytorch_b200._lib import SG_BF16, SG_F16

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 5
COMPARE = sys.argv[3] if len(sys.argv) > 3 else ""
MODEL_ATOMIC = float(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[3] == "streamk" else 4.5
dev = "cuda"
h = lambda *s: (torch.randn(*s, device=dev) * 0.5).half()
b = lambda *s: (torch.randn(*s, device=dev) * 0.5).bfloat16()


def timeit(name, fn, flops):
    """COMPARE = "compare": wave split on / off;  "areuse": sg_set_cta_pair(2) (A reuse) vs (1)."""
    from segan_pytorch_b200 import _lib
    lib = _lib.load()
    if COMPARE == "areuse":
        settings = [("areuse", lambda: lib.sg_set_cta_pair(2)), ("pair", lambda: lib.sg_set_cta_pair(1))]
    elif COMPARE == "splitk":
        settings = [("splitk", lambda: setattr(E, "SPLITK_TAIL", True)), ("plain", lambda: setattr(E, "SPLITK_TAIL", False))]
    elif COMPARE == "streamk":
        # split factor of the leftover tiles of the last wave: off, forced 2..37 (cost constant ~0), then the model
        settings = [("off", lambda: lib.sg_set_stream_k(0, -1.0))] + \
                   [("S<=%d" % S, (lambda S=S: lib.sg_set_stream_k(S, 1e-6))) for S in (2, 4, 8, 16, 37)] + \
                   [("model", lambda: lib.sg_set_stream_k(16, MODEL_ATOMIC))]
    elif COMPARE == "compare":
        settings = [("split", lambda: setattr(E, "SPLIT_WAVES", True)), ("unsplit", lambda: setattr(E, "SPLIT_WAVES", False))]
    else:
        settings = [("", lambda: None)]
    res = []
    for _, setup in settings:
        setup()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(REP):
            fn()
        e.record()
        torch.cuda.synchronize()
        res.append(s.elapsed_time(e) / REP)
    if COMPARE == "streamk":
        print("%-28s %s" % (name, " | ".join("%s %.3f" % (settings[i][0], ms) for i, ms in enumerate(res))))
        return
    print("%-28s %s" % (name, "   | ".join("%-7s %8.3f ms  %7.1f TFLOP/s" % (settings[i][0], ms, flops / ms / 1e9)
                                          for i, ms in enumerate(res))))


def conv_fwd(cin, cout, R):
    a = h(B, R + 8, 4 * cin)
    w = h(9, cout, 4 * cin)
    out = torch.empty(B, R, cout, device=dev, dtype=torch.float16)
    taps = E.tap_ranges("conv_fwd", cin, 4 * cin, cout)
    fl = E._tap_flops(taps, -4, 4, 0, cout, R * B)
    timeit("conv_fwd %d->%d R=%d" % (cin, cout, R),
           lambda: E.run_f(a, None, R, 4, SG_F16, w, SG_F16, 4 * cin, cout, taps, out, SG_F16, R, 0, 0, R, B, backend=1), fl)


def deconv_fwd(cin, cout, R):
    a0, a1 = h(B, R, cin // 2), h(B, R, cin // 2)
    w = h(9, 4 * cout, cin)
    out = torch.empty(B, R, 4 * cout, device=dev, dtype=torch.float16)
    taps = E.tap_ranges("deconv_fwd", cout, cin, 4 * cout)
    fl = E._tap_flops(taps, -4, 4, 0, 4 * cout, R * B)
    timeit("deconv_fwd %d->%d R=%d" % (cin, cout, R),
           lambda: E.run_f(a0, a1, R, 0, SG_F16, w, SG_F16, cin, 4 * cout, taps, out, SG_F16, R, 0, 0, R, B,
                           a0_c=cin // 2, a1_c=cin // 2, backend=1), fl)


def conv_dgrad(cin, cout, R):
    g = b(B, R, cout)
    w = b(9, 4 * cin, cout)
    out = torch.empty(B, R + 8, 4 * cin, device=dev, dtype=torch.bfloat16)
    taps = E.tap_ranges("conv_dgrad", cin, cout, 4 * cin)
    fl = E._tap_flops(taps, -4, 4, 0, 4 * cin, (R + 8) * B)
    timeit("conv_dgrad %d<-%d R=%d" % (cin, cout, R),
           lambda: E.run_f(g, None, R, 0, SG_BF16, w, SG_BF16, cout, 4 * cin, taps, out, SG_BF16, R, 4, -4, R + 4, B,
                           backend=1), fl)


def conv_wgrad(cin, cout, R):
    g = b(B, R, cout)
    a = b(B, R + 8, 4 * cin)
    dw = torch.zeros(9, cout, 4 * cin, device=dev)
    taps = E.tap_ranges("conv_fwd", cin, 4 * cin, cout)
    fl = E._tap_flops(taps, -4, 4, 0, cout, R * B)
    nt = 9 * (cout // 128) * max(1, 4 * cin // 256)
    ks = E.wgrad_ksplit(B * R, nt)
    timeit("conv_wgrad %d->%d R=%d ks=%d" % (cin, cout, R, ks),
           lambda: E.run_w(g, R, SG_BF16, a, None, R, 4, SG_BF16, 4 * cin, cout, taps, dw, B, ksplit=ks, backend=1), fl)


def wave0():
    """The waveform-end layer as it runs in the step: single-tap GEMM over the 64-column im2col (K = 64, N = 64)."""
    col = h(B, 4096, 64)
    w = h(1, 64, 64)
    bias = torch.randn(64, device=dev)
    out = torch.empty(B, 4096, 64, device=dev, dtype=torch.float16)
    taps = E.tap_ranges("full", 0, 64, 64)
    fl = 2.0 * B * 4096 * 64 * 64
    timeit("wave0 im2col-GEMM 64x64 R=4096", lambda: E.run_f(col, None, 4096, 0, SG_F16, w, SG_F16, 64, 64, taps, out,
                                                              SG_F16, 4096, 0, 0, 4096, B, bias=bias, bias_mod=64,
                                                              d_lo=0, d_hi=0, w_tap0=4, backend=1), fl)


def dump_timeline(name):
    """SEGAN_B200_DEBUG bit 20: per-CTA phase stamps of the LAST tapgemm_f_tc2 launch (sg_debug_timeline)."""
    import ctypes as C
    from segan_pytorch_b200 import _lib
    lib = _lib.load()
    torch.cuda.synchronize()
    buf = (C.c_uint64 * (160 * 32))()
    lib.sg_debug_timeline.argtypes = [C.c_void_p, C.c_int]
    n = lib.sg_debug_timeline(buf, 160 * 32)
    rows = []
    for cta in range(148):
        w = [buf[cta * 32 + i] for i in range(32)]
        w = [x for x in w if x]
        if w:
            rows.append((cta, w))
    if not rows:
        print("no timeline recorded")
        return
    t0 = min(w[0] & ~(1 << 63) for _, w in rows)
    print("# %s timeline (us since the earliest CTA start): cta: start | per piece (acc ready, epilogue done, *finisher done) | exit" % name)
    ends = []
    for cta, w in rows:
        vals = ["%s%.1f" % ("*" if x >> 63 else "", ((x & ~(1 << 63)) - t0) / 1e3) for x in w]
        ends.append(((w[-1] & ~(1 << 63)) - t0) / 1e3)
        if cta < 32 or cta % 16 == 0:
            print("cta %3d: %s" % (cta, " ".join(vals)))
    ends.sort()
    print("exit times us: min %.1f median %.1f max %.1f" % (ends[0], ends[len(ends) // 2], ends[-1]))


ONE = {"wave0": wave0, "enc1": lambda: conv_fwd(64, 128, 1024), "enc3": lambda: conv_fwd(256, 512, 64),
       "enc4": lambda: conv_fwd(512, 1024, 16), "dec1": lambda: deconv_fwd(1024, 256, 64),
       "dgrad3": lambda: conv_dgrad(256, 512, 64), "wgrad3": lambda: conv_wgrad(256, 512, 64),
       "wgrad4": lambda: conv_wgrad(512, 1024, 16)}

if __name__ == "__main__":
    if COMPARE == "one":                # a single shape (ncu captures): python tools/prof_tapgemm.py 300 2 one enc3
        COMPARE = ""
        for name in sys.argv[4:]:
            ONE[name]()
            if int(os.environ.get("SEGAN_B200_DEBUG", "0")) & (1 << 20):
                dump_timeline(name)
        sys.exit(0)
    if COMPARE == "streamk":
        for fn, args in ((conv_fwd, (64, 128, 1024)), (conv_fwd, (128, 256, 256)), (conv_fwd, (256, 512, 64)),
                         (conv_fwd, (512, 1024, 16)), (deconv_fwd, (2048, 512, 16)), (deconv_fwd, (1024, 256, 64)),
                         (deconv_fwd, (512, 128, 256)), (deconv_fwd, (256, 64, 1024)), (conv_dgrad, (512, 1024, 16)),
                         (conv_dgrad, (256, 512, 64)), (conv_dgrad, (128, 256, 256)), (conv_dgrad, (64, 128, 1024))):
            fn(*args)
        sys.exit(0)
    if COMPARE == "areuse3":          # three representative shapes only (timing experiments)
        COMPARE = "areuse"
        conv_fwd(64, 128, 1024)
        deconv_fwd(512, 128, 256)
        conv_dgrad(128, 256, 256)
        sys.exit(0)
    conv_fwd(64, 128, 1024)
    conv_fwd(128, 256, 256)
    conv_fwd(256, 512, 64)
    conv_fwd(512, 1024, 16)
    deconv_fwd(2048, 512, 16)
    deconv_fwd(1024, 256, 64)
    deconv_fwd(512, 128, 256)
    deconv_fwd(256, 64, 1024)
    conv_dgrad(512, 1024, 16)
    conv_dgrad(256, 512, 64)
    conv_dgrad(128, 256, 256)
    conv_dgrad(64, 128, 1024)
    conv_wgrad(64, 128, 1024)
    conv_wgrad(128, 256, 256)
    conv_wgrad(256, 512, 64)
    conv_wgrad(512, 1024, 16)
