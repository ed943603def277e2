"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/qilcuda.h declares, and fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "qilcuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qil_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree(q):
    assert _header_symbols() == q.declared_symbols()


def test_library_exports_every_declared_symbol(q):
    lib = ctypes.CDLL(q.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in q.load_library().qil_version()


def test_no_cpu_fallback_without_gpu(q):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(q.CudaError):
        q.Context(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "qilaplace.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "qil_oracle" not in src and "oracle/" not in src, os.path.join(dirpath, f)


def test_config_parsers(q):
    from qilaplace_b200 import api
    assert api._parse_config_string("0101") == [0, 1, 0, 1]
    assert api._parse_config_string("[1, 0, 1]") == [1, 0, 1]
    assert api._parse_config_string("(1 0 1)") == [1, 0, 1]
    with pytest.raises(q.ArgumentError):
        api._parse_config_string("  ")
    with pytest.raises(q.ArgumentError):
        api._parse_config_string("012")
    assert api._bits_from_integer(0b101, 3) == [1, 0, 1]
    with pytest.raises(q.ArgumentError):
        api._bits_from_integer(0b1000, 3)
    with pytest.raises(q.ArgumentError):
        api._bits_from_integer(-1, 3)


def test_generate_signal_matches_reference_formula(q):
    import numpy as np
    n = 10
    x = q.generate_signal(n, kind="sin_decay", freq=[1.0, 2.5], decay_rate=[0.08, 0.03])
    dt = 1.0 / (2.5 * 2**n)
    j = np.arange(2**n)
    ref = np.sin(1.0 * dt * j) * np.exp(-0.08 * dt * j) + np.sin(2.5 * dt * j) * np.exp(-0.03 * dt * j)
    assert np.allclose(x, ref, rtol=0, atol=1e-15)
    with pytest.raises(q.ArgumentError):
        q.generate_signal(4, kind="nope")
