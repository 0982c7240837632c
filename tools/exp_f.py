"""Timing experiments on the 1-CTA forward-form kernel (results are WRONG under SEGAN_B200_DEBUG)."""
import os, sys, torch
from segan_pytorch_b200 import engine as E, _lib
from segan_pytorch_b200._lib import SG_F16
B = 300
dev = "cuda"
def run(cin, cout, R, tag):
    a = (torch.randn(B, R + 8, 4 * cin, device=dev) * 0.5).half()
    w = (torch.randn(9, cout, 4 * cin, device=dev) * 0.5).half()
    out = torch.empty(B, R, cout, device=dev, dtype=torch.float16)
    taps = E.tap_ranges("conv_fwd", cin, 4 * cin, cout)
    fl = E._tap_flops(taps, -4, 4, 0, cout, R * B)
    fn = lambda: E.run_f(a, None, R, 4, SG_F16, w, SG_F16, 4 * cin, cout, taps, out, SG_F16, R, 0, 0, R, B, backend=1)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print("%s conv_fwd %d->%d R=%d: %.3f ms %.0f TFLOP/s" % (tag, cin, cout, R, ms, fl / ms / 1e9))
pair = int(sys.argv[1]); _lib.load().sg_set_cta_pair(pair)
tag = "pair=%d dbg=%s" % (pair, os.environ.get("SEGAN_B200_DEBUG", "0"))
run(128, 256, 256, tag)   # 600 tiles, 62 k-blocks
run(64, 256, 1024, tag)   # 2400 tiles of N=256, 31 k-blocks (synthetic)
