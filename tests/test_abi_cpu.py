"""CPU-only checks of the C-ABI library and the host-side mirror (no compute calls)."""
import ctypes as C
import os
import re

import pytest
import torch

import aasist_b200
from aasist_b200 import _lib
from tests.util import ROOT, load_sd


def _header_functions():
    text = open(os.path.join(ROOT, "include", "aasist_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aasist_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aasist_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.aasist_abi_version() == _lib.ABI_VERSION == 2


def test_config_struct_layout_matches_header():
    # 4 + 12 + 2 int32, 8 doubles, 1 + 6 + 1 int32 (the v2 fields took six of v1's seven reserved slots)
    assert C.sizeof(_lib.AasistConfig) == 4 * (4 + 12 + 2) + 8 * 8 + 4 * 8
    # pointer + 2 + 4 int32
    assert C.sizeof(_lib.ForwardOpts) == 8 + 4 * 6


def _cls(name):
    return {"RawGAT-ST": aasist_b200.RawGATSTModel, "AASIST-Robust": aasist_b200.RobustModel}.get(name, aasist_b200.Model)


@pytest.mark.parametrize("name", ["AASIST", "AASIST-L", "RawGAT-ST", "AASIST2", "AASIST2-small", "AASIST-Robust"])
def test_expected_param_list_equals_checkpoint_keys(name):
    lib = _lib.load()
    m = _cls(name)(aasist_b200.CONFIGS[name])
    cfg, h = m._config(), C.c_void_p()
    assert lib.aasist_create(C.byref(cfg), C.byref(h)) == 0
    try:
        got = set()
        for i in range(lib.aasist_num_params(h)):
            ne = C.c_int64()
            got.add((lib.aasist_param_name(h, i, C.byref(ne)).decode(), ne.value))
        sd = load_sd(name)
        exp = {(k, v.numel()) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
        assert got == exp
        # strict semantics: unknown key / wrong size are rejected, num_batches_tracked ignored
        buf = (C.c_float * 4)()
        assert lib.aasist_set_param(h, b"no.such.key", buf, 4) == -3
        assert b"unexpected key" in lib.aasist_last_error()
        assert lib.aasist_set_param(h, b"first_bn.weight", buf, 4) == -3
        assert b"size mismatch" in lib.aasist_last_error()
        assert lib.aasist_set_param(h, b"first_bn.num_batches_tracked", buf, 1) == 0
        assert lib.aasist_set_param(h, b"first_bn.weight", buf, 1) == 0
        if not torch.cuda.is_available():
            # no device: compute entry points fail loudly, nothing falls back to the CPU
            assert lib.aasist_finalize(h) < 0
            assert lib.aasist_forward(h, None, 1, 64600, None, None, None, None, None, 0, None) < 0
    finally:
        lib.aasist_destroy(h)


def test_create_rejects_bad_configs():
    lib = _lib.load()
    m = aasist_b200.Model(aasist_b200.CONFIGS["AASIST"])
    h = C.c_void_p()
    cfg = m._config()
    cfg.n_filters = 60                                   # pos_S fixes 23 spectral nodes
    assert lib.aasist_create(C.byref(cfg), C.byref(h)) == -1
    cfg = m._config()
    cfg.enc_channels[1][0] = 16                          # channel chain broken
    assert lib.aasist_create(C.byref(cfg), C.byref(h)) == -1
    cfg = m._config()
    cfg.precision = 7
    assert lib.aasist_create(C.byref(cfg), C.byref(h)) == -1


@pytest.mark.parametrize("name,count", [("AASIST", 297866), ("AASIST-L", 85306), ("RawGAT-ST", 437034),
                                        ("AASIST2", 259079), ("AASIST2-small", None), ("AASIST-Robust", 96556)])
def test_model_mirror_loads_shipped_checkpoints_strictly(name, count):
    """The fork models' state_dicts are the REFERENCE classes' own (oracle/make_golden*.py saved them): same
    keys, same order, same shapes -- `Model(d_args)` here and in the reference build the same network."""
    m = _cls(name)(aasist_b200.CONFIGS[name])
    sd = load_sd(name)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # parameter count exactly as reference main.py:256 computes it (README.md:63: 85,306)
    if count is not None:
        assert sum(p.view(-1).size()[0] for p in m.parameters()) == count
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert all(m.state_dict()[k].shape == v.shape for k, v in sd.items())


def test_encoder_type_follows_the_fork_d_args_keys():
    base = aasist_b200.CONFIGS["AASIST"]
    assert aasist_b200.Model(base)._encoder_kind == _lib.ENC_RESIDUAL23          # shipped checkpoints
    assert aasist_b200.Model(dict(base, res2net_width=14))._encoder_kind == _lib.ENC_RES2NET
    assert aasist_b200.Model(dict(base, encoder="res2net"))._encoder_kind == _lib.ENC_RES2NET
    m = aasist_b200.Model(aasist_b200.CONFIGS["AASIST2"])
    assert m.use_speaker_conditioning and m.spk_emb_dim == 256 and m.conditioning_level == "frame"
    keys = set(m.state_dict().keys())
    assert "encoder.1.0.convs.13.weight" in keys and "encoder.2.0.se.fc.2.weight" in keys
    assert "spk_cond_gat.attention.2.bias" in keys and "encoder.1.0.conv1.weight" not in keys


def test_freq_aug_draw_consumes_the_reference_rngs():
    import random
    import numpy as np
    from aasist_b200.model import draw_freq_mask
    np.random.seed(3)
    random.seed(3)
    a = int(np.random.uniform(0, 20))
    a0 = random.randint(0, 70 - a)
    np.random.seed(3)
    random.seed(3)
    assert draw_freq_mask(70) == (a0, a)


def test_plugin_modules_follow_reference_architecture_names():
    from importlib import import_module
    for arch, name in (("AASIST", "AASIST"), ("RawNetGatSpoofST", "RawGAT-ST"), ("AASIST_Robust", "AASIST-Robust")):
        mod = import_module(f"aasist_b200.models.{arch}")          # reference main.py:253
        model = getattr(mod, "Model")(aasist_b200.CONFIGS[name])
        assert hasattr(model, "forward") and hasattr(model, "load_state_dict")


def test_no_cpu_fallback_in_python_shim():
    m = aasist_b200.Model(aasist_b200.CONFIGS["AASIST-L"]).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 64600))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 64600), Freq_aug=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aasist_b200.RobustModel(aasist_b200.CONFIGS["AASIST-Robust"]).eval()(torch.zeros(1, 600000))
    with pytest.raises(RuntimeError):
        m.score_host(torch.zeros(64600))                   # shape / device validated before the C call


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "aasist_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_load_model_config_reads_reference_format(tmp_path):
    import json
    p = tmp_path / "x.conf"
    p.write_text(json.dumps({"model_path": "a.pth", "model_config": aasist_b200.CONFIGS["AASIST-L"]}))
    assert aasist_b200.load_model_config(str(p)) == aasist_b200.CONFIGS["AASIST-L"]
