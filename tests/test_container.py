"""QILTN001 container: the oracle's writer/reader on the CPU, and -- on the GPU -- the library's
qil_mps_save / qil_mps_load / qil_mpo_save / qil_mpo_load against it (same bytes, same tensors)."""
import os

import numpy as np
import pytest

import qil_container as QC
import qil_oracle as O


def test_oracle_container_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    cores, c = O.tt_svd(x, cutoff=1e-14)
    p = tmp_path / "a.qiltn"
    QC.save(p, cores, c)
    got, amp, kind = QC.load(p)
    assert kind == 0 and amp == c and len(got) == len(cores)
    for a, b in zip(got, cores):
        assert np.array_equal(a, b)
    W = O.build_qft_mpo(4)
    QC.save(tmp_path / "w.qiltn", W)
    gw, _, kind = QC.load(tmp_path / "w.qiltn")
    assert kind == 1 and all(np.array_equal(a, b) for a, b in zip(gw, W))
    # header layout is part of the contract
    raw = open(p, "rb").read()
    assert raw[:8] == b"QILTN001" and int.from_bytes(raw[8:12], "little") == 0 and int.from_bytes(raw[12:16], "little") == 1


@pytest.mark.gpu
def test_library_container_matches_oracle(q, tmp_path):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(256)
    psi = q.signal_mps(x, cutoff=1e-13)
    p = str(tmp_path / "psi.qiltn")
    q.save(psi, p)
    cores, amp, kind = QC.load(p)
    assert kind == 0 and amp == psi.amplitude
    for a, b in zip(cores, psi.cores()):
        assert np.array_equal(a, b)
    back = q.load_mps(p)
    assert back.bonds == psi.bonds and back.amplitude == psi.amplitude
    assert np.array_equal(q.mps_to_vector(back), q.mps_to_vector(psi))
    # oracle-written file read by the library (complex MPS, complex MPO)
    xo = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    co, c = O.tt_svd(xo, cutoff=1e-14)
    QC.save(tmp_path / "o.qiltn", co, c)
    lib = q.load_mps(str(tmp_path / "o.qiltn"))
    assert np.abs(q.mps_to_vector(lib) - O.mps_to_vector(co, c)).max() <= 1e-14 * np.abs(xo).max()
    W = q.build_qft_mpo(6, cutoff=1e-14)
    q.save(W, str(tmp_path / "w.qiltn"))
    W2 = q.load_mpo(str(tmp_path / "w.qiltn"))
    assert W2.bonds == W.bonds and all(np.array_equal(a, b) for a, b in zip(W2.cores(), W.cores()))
    gw, _, kind = QC.load(tmp_path / "w.qiltn")
    assert kind == 1 and all(np.array_equal(a, b) for a, b in zip(gw, W.cores()))
    with pytest.raises(q.ArgumentError):
        q.load_mpo(p)
