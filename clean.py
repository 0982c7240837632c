#!/usr/bin/env python
"""clean.py -- drop-in for the reference entry point (clean.py:28-110): load train.opts + a G
checkpoint, enhance every wav under --test_files, write 16 kHz wavs to --synthesis_path.
SEGAN: the windows of ALL files are batched across files and streamed (SEGAN.clean_files: wav decode, upload, int16 ->
float + pre-emphasis on the device, G, segmented de-emphasis, download and wav writing overlap); --per_file keeps the
reference's loop (one file per SEGAN.generate call).  WSEGAN: per file, un-chunked (model.py:755-766)."""
import argparse
import glob
import json
import os
import random
import timeit

import numpy as np
import torch
from scipy.io import wavfile

from segan_pytorch_b200.segan.models import SEGAN, WSEGAN
from segan_pytorch_b200.segan.datasets import normalize_wave_minmax, pre_emphasize
from segan_pytorch_b200.hostbind import bind_host_to_gpu


class ArgParser(object):
    def __init__(self, args):
        for k, v in args.items():
            setattr(self, k, v)


def main(opts):
    assert opts.cfg_file is not None and opts.test_files is not None and opts.g_pretrained_ckpt is not None
    with open(opts.cfg_file, "r") as cfg_f:
        args = ArgParser(json.load(cfg_f))
    args.cuda = True
    segan = WSEGAN(args) if getattr(args, "wsegan", False) else SEGAN(args)
    segan.G.load_pretrained(opts.g_pretrained_ckpt, True)
    bind_host_to_gpu(torch.device('cuda', torch.cuda.current_device()))
    segan.cuda()
    segan.G.eval()
    twavs = glob.glob(os.path.join(opts.test_files[0], "*.wav")) if len(opts.test_files) == 1 else opts.test_files
    print("Cleaning {} wavs".format(len(twavs)))
    beg_t = timeit.default_timer()
    if not getattr(args, "wsegan", False) and not opts.per_file:
        # SEGAN: windows of all files batched on the GPU, wav decode / upload / G / download / wav write overlapped
        # (SEGAN.clean_files); WSEGAN enhances each utterance un-chunked (model.py:755-766): the per-file loop below
        done = []

        def on_done(path, n):
            done.append(path)
            print("Cleaned {}/{}: {} ({} samples)".format(len(done), len(twavs), path, n))
        nwin = segan.clean_files(twavs, opts.synthesis_path, batch=opts.batch_windows, on_done=on_done)
        dt = timeit.default_timer() - beg_t
        print("Cleaned {} wavs / {} windows in {:.3f} s ({:.0f} windows/s)".format(len(twavs), nwin, dt, nwin / dt))
        return
    for t_i, twav in enumerate(twavs, start=1):
        rate, wav = wavfile.read(twav)
        wav = pre_emphasize(normalize_wave_minmax(wav), args.preemph)
        pwav = torch.FloatTensor(wav).view(1, 1, -1)
        g_wav, g_c = segan.generate(pwav)
        wavfile.write(os.path.join(opts.synthesis_path, os.path.basename(twav)), 16000, g_wav)
        end_t = timeit.default_timer()
        print("Cleaned {}/{}: {} in {} s".format(t_i, len(twavs), twav, end_t - beg_t))
        beg_t = timeit.default_timer()


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--g_pretrained_ckpt", type=str, default=None)
    parser.add_argument("--test_files", type=str, nargs="+", default=None)
    parser.add_argument("--h5", action="store_true", default=False)
    parser.add_argument("--seed", type=int, default=111)
    parser.add_argument("--synthesis_path", type=str, default="segan_samples")
    parser.add_argument("--cuda", action="store_true", default=False)
    parser.add_argument("--soundfile", action="store_true", default=False)
    parser.add_argument("--cfg_file", type=str, default=None)
    # additive flags
    parser.add_argument("--batch_windows", type=int, default=256, help="windows per Generator launch (streaming path)")
    parser.add_argument("--per_file", action="store_true", default=False, help="the reference's one-file-at-a-time loop")
    opts = parser.parse_args()
    os.makedirs(opts.synthesis_path, exist_ok=True)
    random.seed(opts.seed)
    np.random.seed(opts.seed)
    torch.manual_seed(opts.seed)
    torch.cuda.manual_seed_all(opts.seed)
    main(opts)
