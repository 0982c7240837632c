"""Streaming-kernel microbenchmarks on the GPU box (CUDA events): the optimiser kernels on a bucket of the step's size,
with and without clear-on-read, operand emission and the alpha gradient on the largest layers.
    PYTHONPATH=. python tools/microbench.py > profiles/r2_microbench.txt"""
import ctypes as C
import sys

import torch

from segan_pytorch_b200 import _lib
from segan_pytorch_b200._lib import SG_F16

dev = "cuda"
p_ = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(name, fn, nbytes, rep=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(rep):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / rep
    print("%-44s %8.3f ms  %7.1f GB/s" % (name, ms, nbytes / ms / 1e6))


def main():
    for n in (75_215_000, 30_000_000):
        p, g, s1, s2 = (torch.randn(n, device=dev) for _ in range(4))
        s1.abs_()
        s2.abs_()
        for clear in (0, 1):
            timeit("rmsprop n=%dM clear=%d" % (n // 1_000_000, clear),
                   lambda: _lib.call("sg_rmsprop_step", p_(p), p_(g), p_(s1), n, 5e-5, 0.99, 1e-8, 1.0, clear, st()),
                   n * (20 + 4 * clear))
            timeit("adam    n=%dM clear=%d" % (n // 1_000_000, clear),
                   lambda: _lib.call("sg_adam_step", p_(p), p_(g), p_(s1), p_(s2), n, 5e-5, 0.0, 0.9, 1e-8, 3, 1.0, clear, st()),
                   n * (28 + 4 * clear))
        del p, g, s1, s2
    for (T, nc, kc, name) in ((9, 2048, 2048, "dec0"), (9, 1024, 2048, "enc4"), (1, 256, 16384, "fc0")):
        m = torch.randn(T * nc * kc, device=dev)
        f = torch.empty(T * nc * kc, dtype=torch.float16, device=dev)
        d = torch.empty_like(f)
        timeit("emit %s [%d][%d][%d]" % (name, T, nc, kc),
               lambda: _lib.call("sg_emit_operands", p_(m), T, nc, kc, None, 0, p_(f), p_(d), SG_F16, SG_F16, None, st()),
               T * nc * kc * 8)
        alpha = torch.rand(kc // 2, device=dev)
        da = torch.zeros(kc // 2, device=dev)
        g = torch.randn(T * nc * kc, device=dev)
        timeit("alpha_grad %s" % name,
               lambda: _lib.call("sg_alpha_grad", p_(g), p_(m), T, nc, kc, p_(alpha), kc // 2, p_(da), st()),
               T * nc * kc * 6)
        del m, f, d, g


if __name__ == "__main__":
    main()
