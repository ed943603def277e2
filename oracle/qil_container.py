"""QILTN001 container (include/qilcuda.h): dims + raw cores, the wire format shared by libqilcuda, this numpy oracle and
the Julia reader julia/QILContainer.jl.  TEST INFRASTRUCTURE like the rest of oracle/ (only tests/ and bench.py use it).

  "QILTN001" | u32 kind (0 MPS, 1 MPO) | u32 is_complex | u32 n | u32 0 | f64 amplitude | i64 bond[n+1] | cores (C order)
"""
import struct

import numpy as np

MAGIC = b"QILTN001"


def save(path, cores, amplitude=1.0, kind=None):
    cores = [np.asarray(c) for c in cores]
    if kind is None:
        kind = 0 if cores[0].ndim == 3 else 1
    is_complex = any(np.iscomplexobj(c) for c in cores)
    dt = np.complex128 if is_complex else np.float64
    bond = [c.shape[0] for c in cores] + [cores[-1].shape[-1]]
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<4I", kind, int(is_complex), len(cores), 0))
        f.write(struct.pack("<d", float(amplitude)))
        f.write(np.asarray(bond, dtype="<i8").tobytes())
        for c in cores:
            f.write(np.ascontiguousarray(c, dtype=dt).tobytes())


def load(path):
    """-> (cores, amplitude, kind)"""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path} is not a QILTN001 container")
        kind, is_complex, n, _ = struct.unpack("<4I", f.read(16))
        (amplitude,) = struct.unpack("<d", f.read(8))
        bond = np.frombuffer(f.read(8 * (n + 1)), dtype="<i8")
        dt = np.complex128 if is_complex else np.float64
        mid = (2,) if kind == 0 else (2, 2)
        cores = []
        for i in range(n):
            shape = (int(bond[i]),) + mid + (int(bond[i + 1]),)
            cnt = int(np.prod(shape))
            cores.append(np.frombuffer(f.read(cnt * np.dtype(dt).itemsize), dtype=dt).reshape(shape).copy())
    return cores, amplitude, kind
