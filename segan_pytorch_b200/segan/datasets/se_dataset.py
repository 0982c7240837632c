"""Input contract of the hot path (reference: segan/datasets/se_dataset.py:21-29,108-126,355-368):
int16 PCM -> normalize_wave_minmax -> pre_emphasize(0.95) -> 16384-sample (clean, noisy) windows,
collated as [names, clean(B,16384), noisy(B,16384), slice_idx(B)].

SEDataset is the wav-directory dataset: the reference's slicing and preprocessing
(se_dataset.py:66-94,128-368) without its per-item pickle + double WAV read -- decoded files are kept in a
small cache -- and with an int16 mode whose windows are normalised + pre-emphasised on the GPU.
BASELINE.json's configs use synthetic pairs, provided by SyntheticSEDataset."""
import glob
import os
import random
from collections import OrderedDict
import numpy as np
import torch
from torch.utils.data.dataset import Dataset
from torch.utils.data.dataloader import default_collate


def collate_fn(batch):
    data_batch = []
    uttname_batch = []
    for sample in batch:
        uttname_batch.append(sample[0])
        data_batch.append(sample[1:])
    data_batch = default_collate(data_batch)
    return [uttname_batch] + data_batch


def normalize_wave_minmax(x):
    return (2. / 65535.) * (x - 32767.) + 1.


def pre_emphasize(x, coef=0.95):
    if coef <= 0:
        return x
    x0 = np.reshape(x[0], (1,))
    diff = x[1:] - coef * x[:-1]
    return np.concatenate((x0, diff), axis=0)


def de_emphasize(y, coef=0.95):
    """Host version (the GPU scan `sg_deemphasis` is what SEGAN.generate uses).  Same recurrence as
    se_dataset.py:119-126, evaluated with scipy's direct-form IIR instead of a Python loop."""
    if coef <= 0:
        return y
    from scipy.signal import lfilter
    return lfilter([1.0], [1.0, -coef], np.asarray(y, dtype=np.float64)).astype(np.float32)


class SyntheticSEDataset(Dataset):
    """Synthetic (clean, noisy) 16384-sample pairs of SURVEY.md 8(d): clean = 0.3*randn,
    noisy = clean + 0.1*randn, clamped to [-1, 1]; items shaped like SEDataset.__getitem__."""

    def __init__(self, n_items, slice_size=16384, seed=111):
        g = torch.Generator().manual_seed(seed)
        self.clean = (0.3 * torch.randn(n_items, slice_size, generator=g)).clamp_(-1, 1)
        self.noisy = (self.clean + 0.1 * torch.randn(n_items, slice_size, generator=g)).clamp_(-1, 1)

    def __len__(self):
        return self.clean.shape[0]

    def __getitem__(self, i):
        return ['synthetic_%d' % i, self.clean[i], self.noisy[i], 0]


class DevicePrefetcher(object):
    """Wraps a loader of collated batches [names, clean(B,L), noisy(B,L), slice_idx] and stages the
    windows of batch n+1 on the GPU (a dedicated copy stream, double-buffered device slots) while
    batch n trains, replacing the blocking `.to(device)` at the top of the reference's step
    (segan/models/model.py:288-290).  Yields [names, clean(B,1,L), noisy(B,1,L), slice_idx] with the
    two tensors resident on `device` and the consumer's stream ordered after their copies.
    Pinned source tensors (DataLoader(pin_memory=True)) make the copies truly asynchronous.
    Batches may also arrive as int16 PCM (what the wav files hold): they cross the link at 2 bytes per sample
    and are normalised + pre-emphasised per window on the device (`sg_pcm16_to_wave`)."""

    def __init__(self, loader, device, depth=2, preemph=0.95):
        """preemph: pre-emphasis coefficient applied ON THE DEVICE to batches that arrive as int16 PCM (see
        `_stage`); float batches are taken as already normalised + pre-emphasised, like SEDataset's."""
        self.loader = loader
        self.device = torch.device(device)
        self.depth = max(2, int(depth))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.preemph = float(preemph)
        self.h2d_bytes = 0

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch, slot):
        names, clean, noisy, slice_idx = batch[:4]
        prevs = batch[4] if len(batch) > 4 else None          # (B, 2) int32: sample before each clean / noisy window
        clean, noisy = torch.as_tensor(clean), torch.as_tensor(noisy)
        pcm = clean.dtype == torch.int16
        if not pcm and clean.dtype != torch.float32:
            clean, noisy = clean.float(), noisy.float()
        shape = (clean.shape[0], 1, clean.shape[-1])
        if slot.get("clean") is None or tuple(slot["clean"].shape) != shape:
            slot["clean"] = torch.empty(shape, dtype=torch.float32, device=self.device)
            slot["noisy"] = torch.empty(shape, dtype=torch.float32, device=self.device)
        if pcm and (slot.get("pcm") is None or tuple(slot["pcm"].shape) != (2,) + shape):
            slot["pcm"] = torch.empty((2,) + shape, dtype=torch.int16, device=self.device)
        with torch.cuda.stream(self.copy_stream):
            if slot.get("free") is not None:
                self.copy_stream.wait_event(slot["free"])      # the step that read this slot has finished
            if pcm:
                # int16 PCM over the link (2 B/sample), normalisation + per-window pre-emphasis on the device
                # (se_dataset.py:108-117 done by sg_pcm16_to_wave instead of the host)
                from ... import _lib
                import ctypes as C
                slot["pcm"][0].copy_(clean.reshape(shape), non_blocking=True)
                slot["pcm"][1].copy_(noisy.reshape(shape), non_blocking=True)
                pv = None
                if prevs is not None:
                    # (2, B) int32 on the device: the PCM sample preceding every window in its file
                    pv = torch.as_tensor(prevs, dtype=torch.int32).t().contiguous().to(self.device, non_blocking=True)
                    slot["prev"] = pv
                st = C.c_void_p(self.copy_stream.cuda_stream)
                for i, dst in enumerate((slot["clean"], slot["noisy"])):
                    _lib.call("sg_pcm16_to_wave", C.c_void_p(slot["pcm"][i].data_ptr()),
                              C.c_void_p(pv[i].data_ptr()) if pv is not None else None, shape[0], shape[2],
                              self.preemph, C.c_void_p(dst.data_ptr()), None, st)
            else:
                slot["clean"].copy_(clean.reshape(shape), non_blocking=True)
                slot["noisy"].copy_(noisy.reshape(shape), non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        self.h2d_bytes += 2 * clean.numel() * (2 if pcm else 4)
        return names, slot, slice_idx, ready

    def __iter__(self):
        it = iter(self.loader)
        slots = [dict() for _ in range(self.depth)]
        n = 0
        try:
            staged = self._stage(next(it), slots[0])
        except StopIteration:
            return
        while staged is not None:
            names, slot, slice_idx, ready = staged
            n += 1
            try:
                staged = self._stage(next(it), slots[n % self.depth])      # batch n+1 copies while batch n trains
            except StopIteration:
                staged = None
            torch.cuda.current_stream(self.device).wait_event(ready)
            yield [names, slot["clean"], slot["noisy"], slice_idx]
            free = torch.cuda.Event()
            free.record(torch.cuda.current_stream(self.device))            # everything enqueued for this batch
            slot["free"] = free


class SEDataset(Dataset):
    """Wav-directory speech-enhancement dataset with the reference's semantics (se_dataset.py:128-368):
    pairs of 16 kHz wavs from `clean_dir` / `noisy_dir` (matched by sorted file name), every file normalised with
    normalize_wave_minmax and pre-emphasised AS A WHOLE (read_wav_file, :191-199; `preemph_norm` swaps the
    order), sliced into `slice_size` windows every int(slice_size * stride) samples (slice_signal_index, :66-94:
    no partial tail window), one random scale per item.  Items: [basename, clean, noisy, slice_idx].

    pcm16=True returns the windows as int16 PCM plus the sample preceding each window ([..., prev(2,)]):
    DevicePrefetcher ships them at 2 bytes per sample and sg_pcm16_to_wave applies the same normalisation and
    whole-file pre-emphasis on the GPU (random_scale must be [1] and preemph_norm False in that mode).
    Differences from the reference: no pickle cache on disk (`cache_dir` is accepted and ignored), decoded files
    are kept in a per-process LRU, file pairs are matched by sorted name instead of glob order."""

    NO_PREV = 0x7fffffff

    def __init__(self, clean_dir, noisy_dir, preemph, cache_dir='.', split='train', slice_size=2 ** 14, stride=0.5,
                 max_samples=None, do_cache=False, verbose=False, slice_workers=2, preemph_norm=False,
                 random_scale=[1], pcm16=False, cache_files=64):
        super(SEDataset, self).__init__()
        from scipy.io import wavfile
        self._wavfile = wavfile
        self.clean_names = sorted(glob.glob(os.path.join(clean_dir, '*.wav')))
        self.noisy_names = sorted(glob.glob(os.path.join(noisy_dir, '*.wav')))
        if len(self.clean_names) != len(self.noisy_names) or len(self.clean_names) == 0:
            raise ValueError('No wav data found! Check your data path please')
        if max_samples is not None:
            self.clean_names = self.clean_names[:max_samples]
            self.noisy_names = self.noisy_names[:max_samples]
        self.slice_size, self.stride, self.split = int(slice_size), stride, split
        self.preemph, self.preemph_norm, self.random_scale = preemph, preemph_norm, list(random_scale)
        self.pcm16 = bool(pcm16)
        if self.pcm16 and (preemph_norm or any(r != 1 for r in self.random_scale)):
            raise ValueError('pcm16=True needs preemph_norm=False and random_scale=[1]')
        assert 0 < stride <= 1, stride
        self._cache, self._cache_files = OrderedDict(), int(cache_files)
        # slice_signal_index (se_dataset.py:66-94): windows [beg, beg + slice_size), beg += int(slice_size * stride)
        offset = int(self.slice_size * stride)
        self.idx2slice = []
        for w_i, (c_path, n_path) in enumerate(zip(self.clean_names, self.noisy_names)):
            n_samples = self._raw(c_path).shape[0]
            if self.pcm16 and self._raw(n_path).shape[0] < n_samples:
                # a noisy file shorter than its clean twin gets trimmed + zero-padded windows (_common_window); the
                # padding lives in the pre-emphasised domain, which raw PCM cannot express: serve float windows
                print('SEDataset: %s is shorter than %s -> float windows instead of int16 PCM'
                      % (os.path.basename(n_path), os.path.basename(c_path)))
                self.pcm16 = False
            for t_i, beg in enumerate(range(0, n_samples - self.slice_size + 1, offset)):
                self.idx2slice.append((w_i, t_i, beg))

    def __len__(self):
        return len(self.idx2slice)

    def _raw(self, path):
        wav = self._cache.get(path)
        if wav is None:
            rate, wav = self._wavfile.read(path)
            if wav.ndim != 1:
                raise ValueError('mono wavs expected: %s' % path)
            if wav.dtype != np.int16:
                raise ValueError('16-bit PCM wavs expected: %s (%s)' % (path, wav.dtype))
            self._cache[path] = wav
            while len(self._cache) > self._cache_files:
                self._cache.popitem(last=False)
        else:
            self._cache.move_to_end(path)
        return wav

    def read_wav_file(self, path):
        """se_dataset.py:191-199 on the decoded int16 samples."""
        wav = self._raw(path)
        if self.preemph_norm:
            return normalize_wave_minmax(pre_emphasize(wav, self.preemph))
        return pre_emphasize(normalize_wave_minmax(wav), self.preemph)

    def _common_window(self, c_slice, n_slice, pad_value):
        """extract_slice, se_dataset.py:338-347: the pair is cut to the shorter of the two (a noisy file may be
        shorter than its clean twin) and zero-padded up to slice_size."""
        m = min(len(c_slice), len(n_slice))
        c_slice, n_slice = c_slice[:m], n_slice[:m]
        if m < self.slice_size:
            pad = np.full((self.slice_size - m,), pad_value, dtype=c_slice.dtype)
            c_slice, n_slice = np.concatenate((c_slice, pad)), np.concatenate((n_slice, pad))
        return c_slice, n_slice

    def __getitem__(self, index):
        w_i, t_i, beg = self.idx2slice[index]
        c_path, n_path = self.clean_names[w_i], self.noisy_names[w_i]
        bname = os.path.splitext(os.path.basename(n_path))[0]
        end = beg + self.slice_size
        if self.pcm16:
            c, n = self._raw(c_path), self._raw(n_path)
            prev = np.array([int(c[beg - 1]) if beg > 0 else self.NO_PREV,
                             int(n[beg - 1]) if beg > 0 else self.NO_PREV], dtype=np.int32)
            return [bname, torch.from_numpy(np.ascontiguousarray(c[beg:end])),
                    torch.from_numpy(np.ascontiguousarray(n[beg:end])), t_i, torch.from_numpy(prev)]
        c_slice, n_slice = self._common_window(self.read_wav_file(c_path)[beg:end],
                                               self.read_wav_file(n_path)[beg:end], 0.0)
        rscale = random.choice(self.random_scale)
        if rscale != 1:
            c_slice, n_slice = rscale * c_slice, rscale * n_slice
        return [bname, torch.FloatTensor(c_slice), torch.FloatTensor(n_slice), t_i]
