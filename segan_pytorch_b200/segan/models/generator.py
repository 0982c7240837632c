"""Drop-in `Generator` (reference: segan/models/generator.py:18-230).

Same constructor signature, same `forward(x, z=None, ret_hid=False)` contract, same state-dict
keys; the arithmetic runs in libsegan_b200.so (segan_pytorch_b200.engine.GeneratorEngine)."""
import torch
import torch.nn as nn

from .core import Model
from .modules import GConv1DBlock, GDeconv1DBlock
from ... import engine as _engine


class GSkip(nn.Module):
    """Learnable per-channel skip scale (generator.py:18-78).  skip_type 'alpha' | 'constant' with merge 'concat'
    (ckpt_segan+/train.opts) or 'sum' (generator.py:72-74) is served by the kernels."""

    def __init__(self, skip_type, size, skip_init, skip_dropout=0, merge_mode='sum', kwidth=11, bias=True):
        super().__init__()
        self.merge_mode = merge_mode
        if skip_type in ('alpha', 'constant'):
            if skip_init == 'zero':
                alpha_ = torch.zeros(size)
            elif skip_init == 'randn':
                alpha_ = torch.randn(size)
            elif skip_init == 'one':
                alpha_ = torch.ones(size)
            else:
                raise TypeError('Unrecognized alpha init scheme: ', skip_init)
            self.skip_k = nn.Parameter(alpha_.view(1, -1, 1))
            if skip_type == 'constant':
                self.skip_k.requires_grad = False
        elif skip_type == 'conv':
            raise NotImplementedError("skip_type='conv' is a SURVEY.md 8(f)-N4 'next' row; not built yet")
        else:
            raise TypeError('Unrecognized GSkip scheme: ', skip_type)
        self.skip_type = skip_type
        if skip_dropout > 0:
            raise NotImplementedError("skip_dropout > 0 is not built yet (non-default)")

    def __repr__(self):
        return self._get_name() + ('(Alpha(1))' if self.skip_type == 'alpha' else '(Constant(1))')


class _GeneratorFn(torch.autograd.Function):
    """Whole-network autograd node: forward / backward are the fused kernel pipelines."""

    @staticmethod
    def forward(ctx, eng, x, z, *params):
        y, ectx = eng.forward(x, z, fresh=True)
        ctx.eng, ctx.ectx = eng, ectx
        eng._last_ctx = ectx
        ctx.names = [n for n, p in eng.module.named_parameters()]
        return y

    @staticmethod
    def backward(ctx, gy):
        eng = ctx.eng
        scale = _engine.LOSS_SCALE                # the engine's gradient tensors carry the fp16 loss scale
        eng.backward(ctx.ectx, gy * scale if scale != 1.0 else gy)
        grads = []
        for n, p in eng.module.named_parameters():
            grads.append(eng.grad_of(n) if p.requires_grad else None)      # reference layout, true units
        return (None, None, None) + tuple(grads)


class Generator(Model):

    def __init__(self, ninputs, fmaps, kwidth, poolings, dec_fmaps=None, dec_kwidth=None, dec_poolings=None,
                 z_dim=None, no_z=False, skip=True, bias=False, skip_init='one', skip_dropout=0,
                 skip_type='alpha', norm_type=None, skip_merge='sum', skip_kwidth=11, name='Generator'):
        super().__init__(name=name)
        self.skip = skip
        self.bias = bias
        self.no_z = no_z
        self.z_dim = z_dim
        self.enc_blocks = nn.ModuleList()
        assert isinstance(fmaps, list), type(fmaps)
        assert isinstance(poolings, list), type(poolings)
        if isinstance(kwidth, int):
            kwidth = [kwidth] * len(fmaps)
        assert isinstance(kwidth, list), type(kwidth)
        skips = {}
        ninp = ninputs
        for pi, (fmap, pool, kw) in enumerate(zip(fmaps, poolings, kwidth), start=1):
            if skip and pi < len(fmaps):
                gskip = GSkip(skip_type, fmap, skip_init, skip_dropout, merge_mode=skip_merge,
                              kwidth=skip_kwidth, bias=bias)
                l_i = pi - 1
                skips[l_i] = {'alpha': gskip}
                setattr(self, 'alpha_{}'.format(l_i), skips[l_i]['alpha'])
            self.enc_blocks.append(GConv1DBlock(ninp, fmap, kw, stride=pool, bias=bias, norm_type=norm_type))
            ninp = fmap
        self.skips = skips
        if not no_z and z_dim is None:
            z_dim = fmaps[-1]
        if not no_z:
            ninp += z_dim
        if dec_fmaps is None:
            dec_fmaps = fmaps[::-1][1:] + [1]
        else:
            assert isinstance(dec_fmaps, list), type(dec_fmaps)
        if dec_poolings is None:
            dec_poolings = poolings[:]
        else:
            assert isinstance(dec_poolings, list), type(dec_poolings)
        self.dec_poolings = dec_poolings
        if dec_kwidth is None:
            dec_kwidth = kwidth[:]
        elif isinstance(dec_kwidth, int):
            dec_kwidth = [dec_kwidth] * len(dec_fmaps)
        assert isinstance(dec_kwidth, list), type(dec_kwidth)
        self.dec_blocks = nn.ModuleList()
        for pi, (fmap, pool, kw) in enumerate(zip(dec_fmaps, dec_poolings, dec_kwidth), start=1):
            if skip and pi > 1 and pool > 1:
                if skip_merge == 'concat':
                    ninp *= 2
            act = 'Tanh' if pi >= len(dec_fmaps) else None
            if pool > 1:
                dec_block = GDeconv1DBlock(ninp, fmap, kw, stride=pool, norm_type=norm_type, bias=bias, act=act)
            else:
                dec_block = GConv1DBlock(ninp, fmap, kw, stride=1, bias=bias, norm_type=norm_type)
            self.dec_blocks.append(dec_block)
            ninp = fmap
        # ---- what the kernels serve (everything else is a "next" row)
        self.enc_fmaps = list(fmaps)
        self.skip_merge = skip_merge
        self._served = (ninputs == 1 and skip and not no_z and skip_merge in ('concat', 'sum')
                        and skip_type in ('alpha', 'constant') and norm_type is None
                        and all(k == 31 for k in kwidth) and all(k == 31 for k in dec_kwidth)
                        and all(p == 4 for p in poolings) and all(p == 4 for p in dec_poolings)
                        and list(dec_fmaps) == fmaps[::-1][1:] + [1]
                        and (self.z_dim if self.z_dim is not None else fmaps[-1]) == fmaps[-1]
                        and all(f % 64 == 0 for f in fmaps) and fmaps[0] == 64)
        if self.z_dim is None and not no_z:
            self.z_dim = fmaps[-1]
        self.z_device = 'cpu'   # 'cpu' = draw z on the CPU generator like generator.py:197-199; 'cuda' = on device
        self._engine = None

    # -- engine ---------------------------------------------------------------------------------
    @property
    def engine(self):
        if not self._served:
            raise NotImplementedError("this Generator configuration is outside the built hot path "
                                      "(SEGAN+ defaults: alpha skips, concat merge, k=31, stride 4, no norm)")
        if self._engine is None:
            self._engine = _engine.GeneratorEngine(self)
        return self._engine

    # The big weights' nn.Parameters are reference-layout mirrors of the engine's packed fp32 masters: anything that
    # reads or moves the parameters wholesale first refreshes them (a no-op unless an optimiser step intervened).
    def state_dict(self, *args, **kwargs):
        if self._engine is not None:
            self._engine.sync_to_reference()
        return super().state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        if getattr(self, '_engine', None) is not None:
            self._engine.sync_to_reference()
        return super()._apply(fn, *args, **kwargs)

    def forward(self, x, z=None, ret_hid=False):
        eng = self.engine
        _engine._require_cuda(x)
        B, _, L = x.shape
        if z is None:
            code_len = L // (4 ** len(self.enc_blocks))
            if self.z_device == 'cpu':
                z = torch.randn(B, self.z_dim, code_len).to(x.device)      # generator.py:197-199
            else:
                z = torch.randn(B, self.z_dim, code_len, device=x.device)
        if len(z.size()) != 3:
            raise ValueError('len(z.size) {} != len(hi.size) {}'.format(len(z.size()), 3))
        if not hasattr(self, 'z'):
            self.z = z                                                   # generator.py:203-204
        eng.bind()
        if torch.is_grad_enabled() and any(p.requires_grad for p in super().parameters()):
            params = [p for _, p in self.named_parameters()]
            y = _GeneratorFn.apply(eng, x, z, *params)
            ectx = eng._last_ctx          # hidden activations are exposed detached (inspection only)
            eng._last_ctx = None
        else:
            y, ectx = eng.forward(x, z, twins=False)      # inference: no bf16 twins for weight gradients
        if ret_hid:
            # ret_hid may be an iterable of keys (additive): only those activations are converted to NCL
            only = None if ret_hid is True else set(ret_hid)
            return y, eng.hidden_ncl(ectx, only)
        return y
