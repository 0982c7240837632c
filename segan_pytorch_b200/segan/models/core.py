"""Checkpointing and the `Model` base class of the drop-in.

Written for this package; what is shared with the reference (segan/models/core.py:11-207) is the ON-DISK FORMAT
and the public method names, so that checkpoints are interchangeable in both directions:

  <save_path>/<prefix>checkpoints                 JSON index {"latest": [file, ...], "current": file}
  <save_path>/weights_<prefix><Name>-<step>.ckpt  torch.save({"step", "state_dict"[, "optimizer"]})
  a bare state dict (legacy reference checkpoints) is accepted on load

and the retention rule (when the index already lists more than `max_ckpts` files, the oldest is deleted before the
new one is appended).  State-dict key names are those of SURVEY.md App. B.
"""
import json
import os

import torch
import torch.nn as nn


class _CkptIndex(object):
    """The JSON index file next to the checkpoints."""

    def __init__(self, path):
        self.path = path
        self.files, self.current = [], None
        if os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            self.files = list(d.get('latest', []))
            cur = d.get('current')
            self.current = cur if isinstance(cur, str) else None

    def push(self, name, keep, directory):
        """Registers `name` as the newest checkpoint; drops the oldest one from disk and index when the index
        holds more than `keep` entries (keep=None: unlimited)."""
        if keep is not None and len(self.files) > keep:
            victim = os.path.join(directory, 'weights_' + self.files[0])
            if os.path.exists(victim):
                os.remove(victim)
                self.files.pop(0)
            else:
                print('ERROR: ckpt is not there?')
        self.files.append(name)
        self.current = name
        with open(self.path, 'w') as f:
            json.dump({'latest': self.files, 'current': self.current}, f, indent=2)


def _payload(blob):
    """(state_dict, optimizer_state | None) of a loaded checkpoint file, new or legacy format."""
    if isinstance(blob, dict) and 'state_dict' in blob:
        return blob['state_dict'], blob.get('optimizer')
    return blob, None


class Saver(object):

    def __init__(self, model, save_path, max_ckpts=5, optimizer=None, prefix=''):
        self.model, self.optimizer = model, optimizer
        self.save_path, self.prefix, self.max_ckpts = save_path, prefix, max_ckpts
        self.ckpt_path = os.path.join(save_path, prefix + 'checkpoints')

    def _file(self, name):
        return os.path.join(self.save_path, 'weights_' + name)

    def save(self, model_name, step, best_val=False):
        os.makedirs(self.save_path, exist_ok=True)
        name = '%s%s%s-%s.ckpt' % (self.prefix, 'best_' if best_val else '', model_name, step)
        _CkptIndex(self.ckpt_path).push(name, self.max_ckpts, self.save_path)
        blob = {'step': step, 'state_dict': {k: v.detach().to('cpu', copy=True)
                                             for k, v in self.model.state_dict().items()}}
        if self.optimizer is not None:
            blob['optimizer'] = self.optimizer.state_dict()
        torch.save(blob, self._file(name))

    def read_latest_checkpoint(self):
        cur = _CkptIndex(self.ckpt_path).current
        if cur is None:
            print('[!] No checkpoint found in {}'.format(self.save_path))
            return False
        return cur

    def load_weights(self):
        cur = self.read_latest_checkpoint()
        if cur is False:
            print('[!] No weights to be loaded')
            return False
        state, opt_state = _payload(torch.load(self._file(cur), map_location='cpu'))
        self.model.load_state_dict(state)
        if self.optimizer is not None and opt_state is not None:
            self.optimizer.load_state_dict(opt_state)
        print('[*] Loaded weights')
        return True

    def load_pretrained_ckpt(self, ckpt_file, load_last=False, load_opt=True):
        """Partial load: keys that exist in the model with the same shape; unless `load_last`, the final two
        entries of the file (the D output layer, in file order) are skipped -- the reference's convention."""
        state, opt_state = _payload(torch.load(ckpt_file, map_location='cpu'))
        own = self.model.state_dict()
        order = list(state.keys())
        usable = set(order if load_last else order[:-2])
        picked = {k: v for k, v in state.items()
                  if k in usable and k in own and tuple(v.shape) == tuple(own[k].shape)}
        print('Current Model keys: ', len(own))
        print('Loading Pt Model keys: ', len(picked))
        if len(picked) != len(own):
            print('WARNING: LOADING DIFFERENT NUM OF KEYS')
        for k in own:
            if k not in usable:
                print('WARNING: {} weights not loaded from pt ckpt'.format(k))
        own.update(picked)
        self.model.load_state_dict(own)
        if load_opt and self.optimizer is not None and opt_state is not None:
            self.optimizer.load_state_dict(opt_state)


class Model(nn.Module):

    def __init__(self, name='BaseModel'):
        super().__init__()
        self.name = name
        self.optim = None

    def _own_saver(self, save_path):
        if not hasattr(self, 'saver'):
            self.saver = Saver(self, save_path, optimizer=self.optim, prefix=self.name + '-')
        return self.saver

    def save(self, save_path, step, best_val=False, saver=None):
        (saver if saver is not None else self._own_saver(save_path)).save(self.name, step, best_val=best_val)

    def load(self, save_path):
        if os.path.isdir(save_path):
            self._own_saver(save_path).load_weights()
        else:
            print('Loading ckpt from ckpt: ', save_path)
            self.load_pretrained(save_path)

    def load_pretrained(self, ckpt_path, load_last=False):
        Saver(self, '.', optimizer=self.optim).load_pretrained_ckpt(ckpt_path, load_last)

    def activation(self, name):
        return getattr(nn, name)()

    def parameters(self):
        """Trainable parameters only (matters for skip_type='constant': its alphas are frozen)."""
        return (p for p in super().parameters() if p.requires_grad)

    def get_n_params(self):
        return sum(p.numel() for p in self.parameters())
