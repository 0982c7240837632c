"""Input contract of the hot path (reference: segan/datasets/se_dataset.py:21-29,108-126,355-368):
int16 PCM -> normalize_wave_minmax -> pre_emphasize(0.95) -> 16384-sample (clean, noisy) windows,
collated as [names, clean(B,16384), noisy(B,16384), slice_idx(B)].

The wav-directory dataset itself (pickle cache + per-sample WAV reads) is a SURVEY.md 8(f)-N3
"next" row; BASELINE.json's configs use synthetic pairs, provided here by SyntheticSEDataset."""
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
        names, clean, noisy, slice_idx = batch
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
                st = C.c_void_p(self.copy_stream.cuda_stream)
                for i, dst in enumerate((slot["clean"], slot["noisy"])):
                    _lib.call("sg_pcm16_to_wave", C.c_void_p(slot["pcm"][i].data_ptr()), shape[0], shape[2],
                              self.preemph, C.c_void_p(dst.data_ptr()), st)
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
