"""The C-ABI library loads and exports every symbol include/titok_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "titok_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ttk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "ttk_attn_varlen_fwd" in syms and "ttk_fsq_fwd" in syms and "ttk_vq_argmin" in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    from titok_video_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in titok_b200.h but not exported: {missing}"


def test_python_binding_covers_header():
    from titok_video_b200 import _lib

    assert sorted(_lib.SIGNATURES.keys()) == declared_symbols()


def test_status_strings():
    from titok_video_b200 import _lib

    assert _lib.strerror(0) == "ok"
    assert "10.x" in _lib.strerror(-4)
    assert _lib.version() >= 100
    with pytest.raises(_lib.TitokB200Error):
        _lib.check(-2, "unit")


def test_missing_extension_fails_loudly(tmp_path, monkeypatch):
    """The product must not degrade silently when the .so is absent."""
    import importlib

    from titok_video_b200 import _lib

    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.TitokB200Error):
        _lib._load()


def test_cpu_tensors_are_rejected():
    import torch

    import titok_video_b200 as T

    q = T.FSQ([7, 5, 5, 5, 5])
    with pytest.raises(T._lib.TitokB200Error):
        q(torch.zeros(4, 5))


def test_sequencers_reject_null_arguments_without_a_gpu():
    """Argument validation happens before any CUDA call: callable (and checkable) on a machine without a GPU."""
    from titok_video_b200 import _lib

    z = ctypes.c_void_p(0)
    assert _lib.fn("ttk_layers_fwd")(z, z, z, z, z, z, z, z) == -1
    assert _lib.fn("ttk_layers_fwd_train")(z, z, z, z, 0, z, z, z) == -1
    assert _lib.fn("ttk_layers_bwd")(z, z, z, z, 0, z, z, z, z, z, z, z, z, z) == -1
    d = _lib.LayersDesc()
    assert ctypes.sizeof(d) == 8 * 4 + 2 * 4 + 6 * 8  # layout of ttk_layers_desc in include/titok_b200.h
    with pytest.raises(_lib.TitokB200Error):
        _lib.call("ttk_layers_fwd", ctypes.byref(d), z, z, z, z, z, z, z)
