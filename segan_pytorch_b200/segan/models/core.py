"""Checkpoint format and Model base, compatible with the reference (segan/models/core.py:11-207):
`weights_<prefix><Name>-<step>.ckpt` = torch.save({'step','state_dict','optimizer'}), a JSON index
`<prefix>checkpoints` with 'latest' / 'current', rolling window of `max_ckpts`, and the
shape-filtered partial load of `load_pretrained_ckpt` (load_last=False drops the last two keys).
State-dict key names are those of SURVEY.md App. B, so reference checkpoints load unchanged."""
import json
import os

import torch
import torch.nn as nn


class Saver(object):

    def __init__(self, model, save_path, max_ckpts=5, optimizer=None, prefix=''):
        self.model = model
        self.save_path = save_path
        self.ckpt_path = os.path.join(save_path, '{}checkpoints'.format(prefix))
        self.max_ckpts = max_ckpts
        self.optimizer = optimizer
        self.prefix = prefix

    def save(self, model_name, step, best_val=False):
        save_path = self.save_path
        os.makedirs(save_path, exist_ok=True)
        if os.path.exists(self.ckpt_path):
            with open(self.ckpt_path, 'r') as f:
                ckpts = json.load(f)
        else:
            ckpts = {'latest': [], 'current': []}
        model_path = '{}-{}.ckpt'.format(model_name, step)
        if best_val:
            model_path = 'best_' + model_path
        model_path = '{}{}'.format(self.prefix, model_path)
        latest = ckpts['latest']
        if len(latest) > 0 and self.max_ckpts is not None and len(latest) > self.max_ckpts:
            todel = latest[0]
            try:
                os.remove(os.path.join(save_path, 'weights_' + todel))
                latest = latest[1:]
            except FileNotFoundError:
                print('ERROR: ckpt is not there?')
        latest += [model_path]
        ckpts['latest'] = latest
        ckpts['current'] = model_path
        with open(self.ckpt_path, 'w') as f:
            f.write(json.dumps(ckpts, indent=2))
        st_dict = {'step': step,
                   'state_dict': {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}}
        if self.optimizer is not None:
            st_dict['optimizer'] = self.optimizer.state_dict()
        torch.save(st_dict, os.path.join(save_path, 'weights_' + model_path))

    def read_latest_checkpoint(self):
        if not os.path.exists(self.ckpt_path):
            print('[!] No checkpoint found in {}'.format(self.save_path))
            return False
        with open(self.ckpt_path, 'r') as f:
            ckpts = json.load(f)
        return ckpts['current']

    def load_weights(self):
        curr_ckpt = self.read_latest_checkpoint()
        if curr_ckpt is False:
            print('[!] No weights to be loaded')
            return False
        st_dict = torch.load(os.path.join(self.save_path, 'weights_' + curr_ckpt), map_location='cpu')
        if 'state_dict' in st_dict:
            self.model.load_state_dict(st_dict['state_dict'])
            if self.optimizer is not None and 'optimizer' in st_dict:
                self.optimizer.load_state_dict(st_dict['optimizer'])
        else:
            self.model.load_state_dict(st_dict)   # legacy: bare state dict
        print('[*] Loaded weights')
        return True

    def load_pretrained_ckpt(self, ckpt_file, load_last=False, load_opt=True):
        model_dict = self.model.state_dict()
        st_dict = torch.load(ckpt_file, map_location=lambda storage, loc: storage)
        pt_dict = st_dict['state_dict'] if 'state_dict' in st_dict else st_dict
        all_pt_keys = list(pt_dict.keys())
        allowed_keys = all_pt_keys[:] if load_last else all_pt_keys[:-2]
        pt_dict = {k: v for k, v in pt_dict.items() if k in model_dict and
                   k in allowed_keys and v.size() == model_dict[k].size()}
        print('Current Model keys: ', len(list(model_dict.keys())))
        print('Loading Pt Model keys: ', len(list(pt_dict.keys())))
        if len(pt_dict.keys()) != len(model_dict.keys()):
            print('WARNING: LOADING DIFFERENT NUM OF KEYS')
        model_dict.update(pt_dict)
        self.model.load_state_dict(model_dict)
        for k in model_dict.keys():
            if k not in allowed_keys:
                print('WARNING: {} weights not loaded from pt ckpt'.format(k))
        if self.optimizer is not None and 'optimizer' in st_dict and load_opt:
            self.optimizer.load_state_dict(st_dict['optimizer'])


class Model(nn.Module):

    def __init__(self, name='BaseModel'):
        super().__init__()
        self.name = name
        self.optim = None

    def save(self, save_path, step, best_val=False, saver=None):
        model_name = self.name
        if not hasattr(self, 'saver') and saver is None:
            self.saver = Saver(self, save_path, optimizer=self.optim, prefix=model_name + '-')
        if saver is None:
            self.saver.save(model_name, step, best_val=best_val)
        else:
            saver.save(model_name, step, best_val=best_val)

    def load(self, save_path):
        if os.path.isdir(save_path):
            if not hasattr(self, 'saver'):
                self.saver = Saver(self, save_path, optimizer=self.optim, prefix=self.name + '-')
            self.saver.load_weights()
        else:
            print('Loading ckpt from ckpt: ', save_path)
            self.load_pretrained(save_path)

    def load_pretrained(self, ckpt_path, load_last=False):
        saver = Saver(self, '.', optimizer=self.optim)
        saver.load_pretrained_ckpt(ckpt_path, load_last)

    def activation(self, name):
        return getattr(nn, name)()

    def parameters(self):
        return filter(lambda p: p.requires_grad, super().parameters())

    def get_n_params(self):
        return sum(p.numel() for p in self.parameters())
