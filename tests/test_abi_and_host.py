"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the
ctypes structs mirror the C structs, and the host-side logic of the drop-in (state-dict keys,
checkpoint format, RNG order, tap tables, flat buckets) behaves like the reference."""
import ctypes
import os
import random
import re
import subprocess
import tempfile

import pytest
import torch

from tests.util import build_segan, load_opts, sd_sha, golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "segan_b200.h")


@pytest.fixture(scope="module")
def lib():
    from segan_pytorch_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    src = open(HEADER).read()
    names = sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    from segan_pytorch_b200 import _lib
    assert sorted(set(_lib.EXPORTS)) == names
    assert lib.sg_abi_version() == 2


def test_ctypes_structs_match_c_layout(lib):
    from segan_pytorch_b200._lib import TapGemmF, TapGemmW
    code = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "segan_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(sg_tapgemm_f), offsetof(sg_tapgemm_f, out), offsetof(sg_tapgemm_f, backend),
             sizeof(sg_tapgemm_w), offsetof(sg_tapgemm_w, dw), offsetof(sg_tapgemm_w, backend));
      return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    got = [ctypes.sizeof(TapGemmF), TapGemmF.out.offset, TapGemmF.backend.offset,
           ctypes.sizeof(TapGemmW), TapGemmW.dw.offset, TapGemmW.backend.offset]
    assert [int(v) for v in out] == got


def test_state_dict_keys_and_seed_parity():
    s = build_segan()
    g = golden("g_forward_cfg1.npz")
    assert sd_sha(s.G.state_dict()) == str(g["sha_G"])
    keys = list(s.G.state_dict().keys())
    assert keys[:3] == ["enc_blocks.0.conv.weight", "enc_blocks.0.conv.bias", "enc_blocks.0.act.weight"]
    assert "alpha_3.skip_k" in keys and "dec_blocks.4.deconv.bias" in keys and "dec_blocks.4.act.weight" not in keys
    dk = list(s.D.state_dict().keys())
    assert "enc_blocks.0.norm.running_var" in dk and "fc.4.bias" in dk
    assert s.G.get_n_params() == 64770561 and s.D.get_n_params() == 25825793
    nb = build_segan(bias=False)
    assert "enc_blocks.0.conv.bias" not in nb.G.state_dict() and "dec_blocks.0.deconv.bias" in nb.G.state_dict()


def test_missing_reg_loss_tolerated():
    o = load_opts()
    del o.reg_loss
    from segan_pytorch_b200.segan.models import SEGAN
    assert SEGAN(o).reg_loss_name == "l1_loss"


def test_phase_shift_draw_order_matches_oracle():
    from oracle import segan_oracle as O
    from segan_pytorch_b200.segan.models.discriminator import draw_phase_shifts
    random.seed(99)
    a = [draw_phase_shifts(5, 5) for _ in range(3)]
    random.seed(99)
    b = [O.draw_phase_shifts(5, 5) for _ in range(3)]
    assert a == b


def test_flat_buckets_and_checkpoint_roundtrip(tmp_path):
    s = build_segan()
    eng = s.G.engine.bind()
    assert eng.flat.numel() == 64770561
    w = dict(s.G.named_parameters())["enc_blocks.1.conv.weight"]
    assert w.data_ptr() == eng.pview("enc_blocks.1.conv.weight").data_ptr()
    sha = sd_sha(s.G.state_dict())
    s.G.save(str(tmp_path), 7)
    files = os.listdir(str(tmp_path))
    assert "weights_Generator-Generator-7.ckpt" in files and "Generator-checkpoints" in files
    st = torch.load(os.path.join(str(tmp_path), "weights_Generator-Generator-7.ckpt"))
    assert set(st.keys()) >= {"step", "state_dict"}
    s2 = build_segan(seed=5)
    assert sd_sha(s2.G.state_dict()) != sha
    s2.G.load_pretrained(os.path.join(str(tmp_path), "weights_Generator-Generator-7.ckpt"), True)
    assert sd_sha(s2.G.state_dict()) == sha


def test_tap_tables_cover_31_taps():
    from segan_pytorch_b200.engine import tap_ranges
    for kind, c, kc, nc in (("conv_fwd", 64, 256, 128), ("conv_dgrad", 64, 128, 256),
                            ("deconv_fwd", 64, 256, 256), ("deconv_dgrad", 64, 256, 256)):
        k_lo, k_hi, n_lo, n_hi = tap_ranges(kind, c, kc, nc)
        blocks = 0
        for i in range(9):
            blocks += ((k_hi[i] - k_lo[i]) // c if "conv_fwd" == kind or kind == "deconv_dgrad" else 4) * \
                      ((n_hi[i] - n_lo[i]) // c if kind in ("conv_dgrad", "deconv_fwd") else 4) // 4
        assert blocks == 31, (kind, blocks)


def test_cpu_forward_fails_loudly():
    s = build_segan()
    with pytest.raises(RuntimeError):
        s.G(torch.zeros(1, 1, 16384))
    with pytest.raises(RuntimeError):
        s.D(torch.zeros(1, 2, 16384))
