"""The C-ABI library loads and exports every symbol include/szb200.h declares (no compute, no GPU)."""
import os
import re

from tests import util


def test_header_symbols_exported():
    from sigma_zero_b200 import _lib
    header = open(os.path.join(util.ROOT, "include", "szb200.h")).read()
    declared = set(re.findall(r"\b(szb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.szb_version().startswith(b"szb200")


def test_struct_layouts():
    import ctypes
    from sigma_zero_b200 import _lib
    assert ctypes.sizeof(_lib.Pos) == 112
    assert ctypes.sizeof(_lib.Config) == 20
    assert ctypes.sizeof(_lib.Stats) == 48
    assert ctypes.sizeof(_lib.TowerSpans) == 56 and ctypes.sizeof(_lib.PhaseTimes) == 80
    assert ctypes.sizeof(_lib.TrainConfig) == 56 and _lib.TrainConfig.step0.offset == 40


def test_no_device_fails_loudly():
    import torch
    from sigma_zero_b200.engine import Engine, SzbError
    if torch.cuda.is_available():
        return
    try:
        Engine(4, 10)
    except SzbError as e:
        assert e.code == -2
    else:
        raise AssertionError("Engine() must not succeed without a CUDA device")
