#!/bin/bash
# End-to-end smoke of the two drop-in entry points on a GPU box: train.py on synthetic pairs (8 steps: eager
# warm-up, graph capture, replays, checkpoint at the end of the epoch), then clean.py on a synthetic wav with
# the checkpoint and the train.opts it wrote.
set -e
export PYTHONPATH=.
OUT=${1:-/tmp/segan_cli_smoke}
rm -rf "$OUT"; mkdir -p "$OUT/wavs"
python train.py --save_path "$OUT/ckpt" --synthetic 512 --batch_size 64 --epoch 1 --save_freq 2 --z_device cuda --num_workers 0
ls "$OUT/ckpt" | head
python - "$OUT" <<'PY'
import sys, numpy as np
from scipy.io import wavfile
rng = np.random.RandomState(0)
wavfile.write(sys.argv[1] + "/wavs/a.wav", 16000, (rng.randn(40000) * 3000).astype(np.int16))
wavfile.write(sys.argv[1] + "/wavs/b.wav", 16000, (rng.randn(16384) * 3000).astype(np.int16))
PY
# the same entry point on wav directories (SEDataset, int16 windows preprocessed on the device)
python - "$OUT" <<'PY'
import os, sys, numpy as np
from scipy.io import wavfile
rng = np.random.RandomState(1)
for d in ("clean_trainset", "noisy_trainset"):
    os.makedirs(os.path.join(sys.argv[1], d))
for i in range(6):
    c = (rng.randn(60000) * 3000).astype(np.int16)
    wavfile.write(os.path.join(sys.argv[1], "clean_trainset", "u%d.wav" % i), 16000, c)
    wavfile.write(os.path.join(sys.argv[1], "noisy_trainset", "u%d.wav" % i), 16000,
                  (c + rng.randn(60000) * 500).clip(-32768, 32767).astype(np.int16))
PY
python train.py --save_path "$OUT/ckpt_wav" --clean_trainset "$OUT/clean_trainset" --noisy_trainset "$OUT/noisy_trainset" \
    --batch_size 8 --epoch 1 --save_freq 2 --num_workers 0
G=$(ls "$OUT"/ckpt/*G*.ckpt | head -1)
python clean.py --g_pretrained_ckpt "$G" --cfg_file "$OUT/ckpt/train.opts" --test_files "$OUT/wavs" --synthesis_path "$OUT/clean"
python - "$OUT" <<'PY'
import sys, numpy as np
from scipy.io import wavfile
for n, T in (("a.wav", 40000), ("b.wav", 16384)):
    r, w = wavfile.read(sys.argv[1] + "/clean/" + n)
    assert r == 16000 and w.shape[0] == T and np.isfinite(w).all(), (n, r, w.shape)
print("cli smoke OK")
PY
