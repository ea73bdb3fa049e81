"""The C-ABI library loads, exports every symbol include/timegan_b200.h declares, and refuses to compute without
a CUDA device (no CPU fallback).  No kernel is launched here."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "timegan_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from timegan_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported by libtimegan_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.ABI_VERSION == int(re.search(r"#define TG_ABI_VERSION (\d+)", (ROOT / "include" / "timegan_b200.h").read_text()).group(1))


def test_set_option_knobs():
    from timegan_b200 import _lib
    before = _lib.lib.tg_wgrad_gru_workspace_bytes(256, 768, 64, 64)
    assert _lib.lib.tg_set_option(b"wgrad_ctas", 37) == 0
    capped = _lib.lib.tg_wgrad_gru_workspace_bytes(256, 768, 64, 64)
    assert _lib.lib.tg_set_option(b"wgrad_ctas", 0) == 0
    assert _lib.lib.tg_wgrad_gru_workspace_bytes(256, 768, 64, 64) == before
    assert 0 < capped <= before                      # one partial per CTA: fewer CTAs, smaller workspace
    assert _lib.lib.tg_set_option(b"no_such_knob", 1) < 0 and "unknown key" in _lib.last_error()


def test_argument_errors_are_reported_not_crashed():
    from timegan_b200 import _lib
    rc = _lib.lib.tg_gru_fwd(None, None, None, None, None, None, 1, 1, 1, 0)
    assert rc < 0 and "null" in _lib.last_error()
    rc = _lib.lib.tg_gru_fwd(None, 16, 16, 16, 16, None, 1, 1, 3000, 0)   # pointers are never dereferenced on the host
    assert rc < 0 and "hidden size" in _lib.last_error()
    with pytest.raises(RuntimeError, match="argument error"):
        _lib.check(rc, "tg_gru_fwd")
    assert _lib.lib.tg_wgrad_workspace_bytes(196608, 192, 64) > 0
    assert _lib.launch_count() == 0 or torch.cuda.is_available()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import timegan_b200 as tg
    from timegan_b200 import train_timegan as tt
    m = tg.TimeGAN(14, 8, 8, 1, 0.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.reconstruct(torch.rand(2, 4, 14))
    with pytest.raises(RuntimeError, match="no CPU path"):
        tt.device_autoselect()
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="only exists as CUDA kernels"):
        tg.FusedAdam([p]).clip_and_step(0.5)


def test_model_state_dict_schema_matches_reference_checkpoints():
    """Key names / shapes of the shipped checkpoints' `model` dict: z=28, h=56, L=1."""
    import timegan_b200 as tg
    m = tg.TimeGAN(14, 28, 56, 1, 0.2)
    sd = m.state_dict()
    expect = {
        "embedder.rnn.rnn.weight_ih_l0": (84, 14), "embedder.rnn.rnn.weight_hh_l0": (84, 28),
        "recovery.rnn.rnn.weight_ih_l0": (168, 28), "recovery.out.weight": (14, 56), "recovery.out.bias": (14,),
        "generator.proj.weight": (28, 56), "supervisor.proj.bias": (28,),
        "discriminator.fc.bias": (1,), "discriminator.fc.weight_orig": (1, 56), "discriminator.fc.weight_u": (1,),
        "discriminator.fc.weight_v": (56,),
    }
    for k, shp in expect.items():
        assert tuple(sd[k].shape) == shp, k
    assert sum(p.numel() for p in m.parameters()) == 65535      # SURVEY.md App. C.1, reference default
    assert m.embedder.rnn.rnn.hidden_size == 28                  # attribute path read at train_timegan.py:179,235
