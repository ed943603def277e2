"""Host-side mirror of the QILaplace.jl public API over the C ABI of libqilcuda.so.

The reference keeps ITensor `Index` bookkeeping on the host; here the host objects only hold an
opaque device handle plus dimensions.  Names, argument meaning and error behaviour follow the
reference (file:line cited per function); Julia's `f!` is spelled `f_` / `f` returning the mutated
object, `W * psi` is `W * psi` (or `apply(W, psi)`).
"""
from __future__ import annotations

import ctypes as C
import math
import re

import numpy as np

from . import _lib
from ._lib import ArgumentError, DomainError, ErrorException, UnsupportedError, CudaError, call

RSVD_ADAPTIVE = 1   # include/qilcuda.h: QIL_RSVD_ADAPTIVE (rank-adaptive sketch width at the top split; opt-in)

BIG = 2**62


# ------------------------------------------------------------------------------------------
# context
# ------------------------------------------------------------------------------------------
class Context:
    """One CUDA device + stream + memory pool (qil_create)."""

    def __init__(self, device=0, stream=None):
        self.handle = _lib.c_ctx()
        if stream is None:
            call("qil_create", int(device), C.byref(self.handle))
        else:
            call("qil_create_on_stream", int(device), C.c_void_p(int(stream)), C.byref(self.handle))
        self.device = int(device)
        self._live = 0                 # SignalMPS / MPO objects that still hold a handle into this context
        self._close_pending = False

    def sync(self):
        call("qil_sync", self.handle)

    def launch_count(self):
        v = C.c_uint64(0)
        call("qil_launch_count", self.handle, C.byref(v))
        return int(v.value)

    def truncation_margin(self, reset=True):
        """Closest cutoff decision since the last reset (qil_truncation_margin); inf if none was taken."""
        v = C.c_double()
        call("qil_truncation_margin", self.handle, 1 if reset else 0, C.byref(v))
        return float("inf") if v.value >= 1e299 else float(v.value)

    def profile_enable(self, on=True):
        call("qil_profile_enable", self.handle, 1 if on else 0)

    def profile_reset(self):
        call("qil_profile_reset", self.handle)

    def profile_read(self, kernel_class):
        """(total_ms, launches) of one kernel class: 0 stream GEMM, 1 coefficient, 2 apply."""
        t, c = C.c_double(), C.c_int64()
        call("qil_profile_read", self.handle, int(kernel_class), C.byref(t), C.byref(c))
        return float(t.value), int(c.value)

    def profile_read_work(self, kernel_class):
        """(total_ms, launches, algorithmic bytes, algorithmic flops) of one kernel class."""
        t, c, b, f = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
        call("qil_profile_read_work", self.handle, int(kernel_class), C.byref(t), C.byref(c), C.byref(b), C.byref(f))
        return float(t.value), int(c.value), float(b.value), float(f.value)

    def close(self):
        """Destroy the context -- now if nothing hangs off it, otherwise as soon as the last SignalMPS / MPO that holds a
        handle into it is released (their handles point into the qil_ctx, so it must outlive them).  A plain counter:
        creating a chain costs one integer increment (a batch step creates hundreds)."""
        if not self.handle:
            return
        for dev, c in list(_default_ctx.items()):
            if c is self:
                del _default_ctx[dev]
        if self._live > 0:
            self._close_pending = True
            return
        self._destroy()

    def _destroy(self):
        if self.handle:
            _lib.load().qil_destroy(self.handle)
            self.handle = None
        self._close_pending = False

    @property
    def closed(self):
        return self.handle is None


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _np_dtype(is_complex):
    return np.complex128 if is_complex else np.float64


def _core_ptrs(cores, is_complex, ndim):
    dt = _np_dtype(is_complex)
    keep = [np.ascontiguousarray(c, dtype=dt) for c in cores]
    for c in keep:
        if c.ndim != ndim:
            raise ArgumentError(f"core must have {ndim} legs, got shape {c.shape}")
    bond = [int(keep[0].shape[0])] + [int(c.shape[-1]) for c in keep]
    for i, c in enumerate(keep):
        if c.shape[0] != bond[i] or any(d != 2 for d in c.shape[1:-1]):
            raise ArgumentError(f"core {i} has inconsistent shape {c.shape}")
    arr = (C.c_void_p * len(keep))(*[c.ctypes.data for c in keep])
    b = (C.c_int64 * len(bond))(*bond)
    return keep, arr, b


# ------------------------------------------------------------------------------------------
# containers (src/mps.jl:37-130, src/mpo.jl:26-99)
# ------------------------------------------------------------------------------------------
class _DeviceChain:
    _free = None
    _dims = None
    _get = None
    _legs = 3

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.handle = handle
        ctx._live += 1

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                getattr(_lib.load(), self._free)(h)
            except Exception:
                pass
            self.handle = None
            ctx = self.ctx
            ctx._live -= 1
            if ctx._close_pending and ctx._live <= 0:
                try:
                    ctx._destroy()
                except Exception:
                    pass

    def _bond_dims(self):
        n = self.nsites_flat
        b = (C.c_int64 * (n + 1))()
        call(self._dims, self.handle, b)
        return [int(v) for v in b]

    def cores(self):
        """Download every core as a numpy array ([l,2,r] or [l,2,2,r])."""
        b = self._bond_dims()
        out = []
        dt = _np_dtype(self.is_complex)
        for i in range(self.nsites_flat):
            shape = (b[i],) + (2,) * (self._legs - 2) + (b[i + 1],)
            a = np.empty(shape, dtype=dt)
            call(self._get, self.handle, i, C.c_void_p(a.ctypes.data))
            out.append(a)
        return out

    def cores_into(self, host_buffer):
        """Download every core into one caller-owned host buffer (any object exposing the buffer protocol or a
        `data_ptr()` / `nbytes`-like pair, e.g. a pinned torch uint8 tensor) with a single synchronisation; returns
        numpy views into that buffer.  MPS chains only."""
        b = self._bond_dims()
        dt = np.dtype(_np_dtype(self.is_complex))
        if hasattr(host_buffer, "data_ptr"):
            ptr, nbytes = int(host_buffer.data_ptr()), int(host_buffer.numel() * host_buffer.element_size())
            raw = (C.c_ubyte * nbytes).from_address(ptr)
        else:
            raw = host_buffer
            ptr, nbytes = C.addressof(C.c_ubyte.from_buffer(raw)), memoryview(raw).nbytes
        call("qil_mps_get_cores", self.handle, C.c_void_p(ptr), C.c_int64(nbytes))
        out, off = [], 0
        for i in range(self.nsites_flat):
            shape = (b[i], 2, b[i + 1])
            cnt = b[i] * 2 * b[i + 1]
            out.append(np.frombuffer(raw, dtype=dt, count=cnt, offset=off).reshape(shape))
            off += cnt * dt.itemsize
        return out


class SignalMPS(_DeviceChain):
    """SignalMPS (src/mps.jl:70-81): n cores [chi_l, 2, chi_r] on the device + `amplitude`."""
    _free, _dims, _get, _legs = "qil_mps_free", "qil_mps_dims", "qil_mps_get_core", 3

    @classmethod
    def from_cores(cls, cores, amplitude=1.0, ctx=None):
        ctx = ctx or default_context()
        is_complex = any(np.iscomplexobj(c) for c in cores)
        keep, arr, b = _core_ptrs(cores, is_complex, 3)
        h = _lib.c_mps()
        call("qil_mps_from_host", ctx.handle, len(keep), int(is_complex), b, arr, float(amplitude), C.byref(h))
        return cls(ctx, h)

    def _info(self):
        n, c, a = C.c_int(), C.c_int(), C.c_double()
        call("qil_mps_info", self.handle, C.byref(n), C.byref(c), C.byref(a))
        return n.value, c.value, a.value

    @property
    def nsites_flat(self):
        return self._info()[0]

    def __len__(self):
        return self.nsites_flat

    @property
    def is_complex(self):
        return bool(self._info()[1])

    @property
    def amplitude(self):
        return self._info()[2]

    @amplitude.setter
    def amplitude(self, v):
        call("qil_mps_set_amplitude", self.handle, float(v))

    @property
    def bonds(self):
        """Inner bond dimensions (length n-1), like `dim.(psi.bonds)`."""
        return self._bond_dims()[1:-1]

    def copy(self):
        h = _lib.c_mps()
        call("qil_mps_clone", self.handle, C.byref(h))
        return type(self)(self.ctx, h)

    def __getitem__(self, config):
        if not isinstance(config, tuple):
            config = (config,)
        return coefficient(self, list(config))


class ZTMPS(SignalMPS):
    """ZTMPS (src/mps.jl:98-121) stored as its 2n-site chain main1, copy1, main2, ... (mps.jl:421-445)."""

    def __len__(self):
        return self.nsites_flat // 2

    @property
    def bonds_main(self):
        return self._bond_dims()[2:-1:2]

    @property
    def bonds_copy(self):
        return self._bond_dims()[1:-1:2]


class SingleSiteMPO(_DeviceChain):
    """SingleSiteMPO (src/mpo.jl:26-46): n cores [D_l, 2(in), 2(out), D_r] on the device."""
    _free, _dims, _get, _legs = "qil_mpo_free", "qil_mpo_dims", "qil_mpo_get_core", 4

    @classmethod
    def from_cores(cls, cores, ctx=None):
        ctx = ctx or default_context()
        is_complex = any(np.iscomplexobj(c) for c in cores)
        keep, arr, b = _core_ptrs(cores, is_complex, 4)
        h = _lib.c_mpo()
        call("qil_mpo_from_host", ctx.handle, len(keep), int(is_complex), b, arr, C.byref(h))
        return cls(ctx, h)

    def _info(self):
        n, c = C.c_int(), C.c_int()
        call("qil_mpo_info", self.handle, C.byref(n), C.byref(c))
        return n.value, c.value

    @property
    def nsites_flat(self):
        return self._info()[0]

    def __len__(self):
        return self.nsites_flat

    @property
    def is_complex(self):
        return bool(self._info()[1])

    @property
    def bonds(self):
        return self._bond_dims()[1:-1]

    def __mul__(self, other):
        return apply(self, other)


class PairedSiteMPO(SingleSiteMPO):
    """PairedSiteMPO (src/mpo.jl:54-75) stored as its 2n-site chain (apply.jl:16-33)."""

    def __len__(self):
        return self.nsites_flat // 2

    @property
    def bonds_main(self):
        return self._bond_dims()[2:-1:2]

    @property
    def bonds_copy(self):
        return self._bond_dims()[1:-1:2]


# ------------------------------------------------------------------------------------------
# coefficient (src/mps.jl:609-693)
# ------------------------------------------------------------------------------------------
def _parse_config_string(spec):
    """_parse_config_string (mps.jl:616-631)."""
    s = spec.strip().strip("[](){}")
    if not s:
        raise ArgumentError("coefficient: configuration string is empty")
    if re.search(r"[,\s]", s):
        toks = [t for t in re.split(r"[,\s]+", s) if t]
        if not toks:
            raise ArgumentError("coefficient: configuration string did not contain any entries")
        return [int(t) for t in toks]
    if any(ch not in "01" for ch in s):
        raise ArgumentError("coefficient: bit strings may contain only '0' or '1'")
    return [1 if ch == "1" else 0 for ch in s]


def _bits_from_integer(value, n):
    """_bits_from_integer (mps.jl:633-645): n-bit big-endian pattern."""
    if value < 0:
        raise ArgumentError("coefficient: integer configuration must be non-negative")
    if value >> n:
        raise ArgumentError(f"coefficient: integer {value} requires more than {n} bits")
    return [(value >> (n - 1 - i)) & 1 for i in range(n)]


def coefficients(psi, bits):
    """Batched `coefficient`: bits is a (B, n) array of 0/1 (2n interleaved columns for a ZTMPS)."""
    b = np.ascontiguousarray(bits)
    if b.ndim != 2:
        raise ArgumentError("coefficients: expected a (B, n) array of bits")
    N = psi.nsites_flat
    if b.shape[1] != N:
        raise ArgumentError(f"coefficient: expected {N} entries, got {b.shape[1]}")
    if b.size and (b.min() < 0 or b.max() > 1):
        bad = int(b.max() if b.max() > 1 else b.min())
        raise ArgumentError(f"coefficient: bit value {bad} outside [0,1]")
    b8 = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(b8.shape[0], dtype=_np_dtype(psi.is_complex))
    call("qil_coefficient_batch", psi.ctx.handle, psi.handle, C.c_void_p(b8.ctypes.data),
         C.c_int64(b8.shape[0]), C.c_void_p(out.ctypes.data))
    return out


def coefficient(psi, *config):
    """coefficient(psi, config) (mps.jl:669-693): vector / tuple / varargs / string / integer."""
    N = psi.nsites_flat
    if len(config) == 1:
        c = config[0]
        if isinstance(c, str):
            bits = _parse_config_string(c)
        elif isinstance(c, (int, np.integer)):
            bits = _bits_from_integer(int(c), N)
        else:
            bits = [int(v) for v in c]
    else:
        bits = [int(v) for v in config]
    if len(bits) != N:
        raise ArgumentError(f"coefficient: expected {N} entries, got {len(bits)}")
    for v in bits:
        if not 0 <= v <= 1:
            raise ArgumentError(f"coefficient: bit value {v} outside [0,1]")
    v = coefficients(psi, np.asarray([bits], dtype=np.uint8))[0]
    return complex(v) if psi.is_complex else float(v)


# ------------------------------------------------------------------------------------------
# apply (src/linalg/apply.jl)
# ------------------------------------------------------------------------------------------
def apply(W, other, **kwargs):
    """apply(W, psi) / apply(W1, W2) (apply.jl:75-236).  kwargs are accepted and ignored, as in the
    reference (`apply` never truncates)."""
    if isinstance(other, SignalMPS):
        paired = isinstance(W, PairedSiteMPO)
        if paired != isinstance(other, ZTMPS):
            raise ArgumentError("apply: PairedSiteMPO needs a ZTMPS and SingleSiteMPO a SignalMPS")
        if paired and W.nsites_flat != 2 * len(other):
            raise ArgumentError("apply: MPO and MPS must have compatible sizes.")
        h = _lib.c_mps()
        call("qil_apply_mpo_mps", W.ctx.handle, W.handle, other.handle, C.byref(h))
        return type(other)(W.ctx, h)
    if isinstance(other, SingleSiteMPO):
        h = _lib.c_mpo()
        s1 = int(kwargs.get("start1", 0))
        s2 = int(kwargs.get("start2", 0))
        call("qil_apply_mpo_mpo", W.ctx.handle, W.handle, other.handle, s1, s2, C.byref(h))
        cls = PairedSiteMPO if isinstance(W, PairedSiteMPO) and isinstance(other, PairedSiteMPO) else SingleSiteMPO
        return cls(W.ctx, h)
    raise ArgumentError("apply: unsupported operand types")


# ------------------------------------------------------------------------------------------
# signals (src/signals/Signals.jl:188-235) -- trivial elementwise host code, out of kernel scope
# ------------------------------------------------------------------------------------------
def apply_zipup(W, psi, cutoff=1e-12, maxdim=None):
    """Truncating W * psi (zip-up sweep with the path's truncated SVD); see qil_apply_mpo_mps_zipup."""
    h = _lib.c_mps()
    call("qil_apply_mpo_mps_zipup", W.ctx.handle, W.handle, psi.handle, float(cutoff), C.c_int64(_maxdim_arg(maxdim)),
         C.byref(h))
    return (ZTMPS if isinstance(psi, ZTMPS) else SignalMPS)(W.ctx, h)


def apply_batch(W, psis):
    """[W * psi for psi in psis] in one launch (qil_apply_mpo_mps_batch); psis: SignalMPS / ZTMPS of one context."""
    psis = list(psis)
    if not psis:
        return []
    count = len(psis)
    hin = (_lib.c_mps * count)(*[p.handle for p in psis])
    hout = (_lib.c_mps * count)()
    call("qil_apply_mpo_mps_batch", W.ctx.handle, W.handle, hin, C.c_int64(count), hout)
    cls = ZTMPS if isinstance(psis[0], ZTMPS) else SignalMPS
    return [cls(W.ctx, _lib.c_mps(h)) for h in hout]


def generate_signal(n, kind="sin", dt=None, freq=None, **kw):
    """generate_signal(n; kind, dt, freq, kwargs...) for the deterministic kinds of the reference
    (:sin, :sin_decay, :abs_cos_power_p8) plus :random via numpy's generator (the reference's
    Xoshiro stream is Julia-specific)."""
    kind = str(kind).lstrip(":")
    N = 2**n
    if kind == "random":
        return np.random.default_rng(kw.get("seed", 1234)).standard_normal(N)
    f = 2 * math.pi if freq is None else freq
    is_vec = isinstance(f, (list, tuple, np.ndarray))
    if dt is None:
        fmax = max(abs(float(v)) for v in f) if is_vec else abs(float(f))
        dt = 1.0 if fmax == 0 else 1.0 / (fmax * N)
    j = np.arange(N, dtype=np.float64)
    if kind == "sin":
        if is_vec:
            ph = kw.get("phase", [0.0] * len(f))
            if len(ph) != len(f):
                raise ArgumentError("Frequency and phase vectors must be of the same length.")
            x = sum(np.sin(w * dt * j + p) for w, p in zip(f, ph))
        else:
            x = np.sin(f * dt * j + kw.get("phase", 0.0))
        nl = kw.get("noise_level", 0.0)
        if nl:
            x = x + nl * np.random.default_rng(kw.get("seed")).standard_normal(N)
        return x
    if kind == "sin_decay":
        dr = kw["decay_rate"]
        if is_vec:
            if len(dr) != len(f):
                raise ArgumentError("Frequency and decay_rate vectors must be of the same length.")
            ph = kw.get("phase")
            if ph is None:
                ph = [0.0] * len(f)
            elif len(ph) != len(f):
                raise ArgumentError("Frequency and phase vectors must be of the same length.")
            return sum(np.sin(w * dt * j + p) * np.exp(-l * dt * j) for w, l, p in zip(f, dr, ph))
        return np.sin(f * dt * j + kw.get("phase", 0.0)) * np.exp(-dr * dt * j)
    if kind == "abs_cos_power_p8":
        return np.abs(np.cos(2 * math.pi * dt * j)) ** kw.get("power", 0.8)
    if kind in ("multi_sin", "multi_sin_exp"):
        # Signals.jl:23-85: n_terms (10) random tones.  The reference draws amplitudes / frequencies / decays from
        # Xoshiro(seed_amp / seed_freq / seed_decay), a Julia-specific stream; this is the same construction on numpy's
        # generator with the same seeds (a deterministic SURROGATE of the same family, not the same numbers).
        nt = int(kw.get("n_terms", 10))
        ak = np.random.default_rng(kw.get("seed_amp", 1001)).random(nt)
        ak = ak / np.linalg.norm(ak)
        wk = (kw.get("ω_scale", kw.get("omega_scale", 40.0)) * dt) * (np.random.default_rng(kw.get("seed_freq", 2002)).random(nt) - 0.5)
        if kind == "multi_sin":
            ph = 2 * math.pi * np.random.default_rng(kw.get("seed_phase", 3003)).random(nt)
            return sum(a * np.sin(w * j + p) for a, w, p in zip(ak, wk, ph))
        lk = -(kw.get("λ_scale", kw.get("lambda_scale", 2.0)) * dt) * np.random.default_rng(kw.get("seed_decay", 4004)).random(nt)
        return sum(a * np.sin(w * j) * np.exp(l * j) for a, w, l in zip(ak, wk, lk))
    raise ArgumentError(
        f"Unsupported signal kind: {kind}. Supported kinds are :sin, :multi_sin, :sin_decay, "
        ":multi_sin_exp, :abs_cos_power_p8, :random.")


# ------------------------------------------------------------------------------------------
# signal -> MPS (src/signals/SignalConverters.jl:228-283)
# ------------------------------------------------------------------------------------------
def _maxdim_arg(maxdim):
    return 0 if maxdim is None or maxdim >= BIG else int(maxdim)


def _signal_array(x):
    x = np.asarray(x)
    if x.ndim != 1:
        raise ArgumentError("signal must be a vector")
    is_complex = np.iscomplexobj(x)
    return np.ascontiguousarray(x, dtype=_np_dtype(is_complex)), is_complex


def signal_mps(x, method="svd", ctx=None, **kwargs):
    """signal_mps(x; method=:svd, kwargs...) (SignalConverters.jl:228-233).

    method "svd": cutoff=1e-15, maxdim.  method "rsvd": additionally k=20, p=10, q=0, random_seed=1234,
    mindim=1 (rsvd.jl:38-50); cutoff/maxdim always override (SignalConverters.jl:133)."""
    ctx = ctx or default_context()
    method = str(method).lstrip(":")
    if method not in ("svd", "rsvd"):
        raise ArgumentError(f"tensor_to_mps: unknown method {method}. Use :svd or :rsvd.")
    xa, is_complex = _signal_array(x)
    cutoff = float(kwargs.pop("cutoff", 1e-15))
    maxdim = _maxdim_arg(kwargs.pop("maxdim", None))
    h = _lib.c_mps()
    if method == "svd":
        if kwargs:
            raise TypeError(f"signal_mps(method=:svd): unexpected keyword(s) {sorted(kwargs)}")
        call("qil_encode_svd", ctx.handle, int(is_complex), C.c_void_p(xa.ctypes.data), C.c_int64(xa.size),
             cutoff, C.c_int64(maxdim), C.byref(h))
    else:
        k = int(kwargs.pop("k", 20)); p = int(kwargs.pop("p", 10)); q = int(kwargs.pop("q", 0))
        seed = int(kwargs.pop("random_seed", 1234)); mindim = int(kwargs.pop("mindim", 1))
        kwargs.pop("verbose", None); kwargs.pop("bondtag", None)
        stream = kwargs.pop("normal_stream", None)
        flags = RSVD_ADAPTIVE if kwargs.pop("adaptive", False) else 0
        if kwargs:
            raise TypeError(f"signal_mps(method=:rsvd): unexpected keyword(s) {sorted(kwargs)}")
        if stream is not None:
            st = np.ascontiguousarray(stream, dtype=_np_dtype(is_complex)).reshape(-1)
            sp, slen = C.c_void_p(st.ctypes.data), st.size
        else:
            sp, slen = None, 0
        call("qil_encode_rsvd", ctx.handle, int(is_complex), C.c_void_p(xa.ctypes.data), C.c_int64(xa.size),
             k, p, q, C.c_int64(seed), cutoff, C.c_int64(maxdim), C.c_int64(mindim), sp, C.c_int64(slen),
             C.c_int64(flags), C.byref(h))
    return SignalMPS(ctx, h)


def signal_ztmps(x, cutoff=1e-10, maxdim=None, ctx=None, **kwargs):
    """signal_ztmps(x; cutoff=1e-10, maxdim, kwargs...) (SignalConverters.jl:247-283)."""
    psi = signal_mps(x, ctx=ctx, cutoff=cutoff, maxdim=maxdim, **kwargs)
    h = _lib.c_mps()
    call("qil_ztmps_split", psi.ctx.handle, psi.handle, float(cutoff), C.c_int64(_maxdim_arg(maxdim)), C.byref(h))
    return ZTMPS(psi.ctx, h)


# ------------------------------------------------------------------------------------------
# canonicalize! / compress! / norm (src/mps.jl:754-999)
# ------------------------------------------------------------------------------------------
def canonicalize(psi, direction, center=None, cutoff=1e-12, maxdim=None):
    """canonicalize!(psi, direction; center, cutoff=1e-12, maxdim) -- in place, returns psi."""
    d = str(direction).lstrip(":")
    if d not in ("right", "left"):
        raise ArgumentError("Direction must be :right or :left")
    n = psi.nsites_flat
    c = 0 if center is None else int(center)
    if center is not None and not 1 <= c <= n:
        raise DomainError(f"Center out of range [1,{n}]")
    call("qil_canonicalize", psi.ctx.handle, psi.handle, 1 if d == "right" else 0, c, float(cutoff),
         C.c_int64(_maxdim_arg(maxdim)))
    return psi


def compress(psi, maxdim=None, tol=1e-12, sweeps=1):
    """compress!(psi; maxdim, tol=1e-12, sweeps=1) -- in place, returns psi."""
    call("qil_compress", psi.ctx.handle, psi.handle, C.c_int64(_maxdim_arg(maxdim)), float(tol), int(sweeps))
    return psi


canonicalize_ = canonicalize
compress_ = compress


def norm(psi):
    """norm(psi) (mps.jl:754-771): sqrt(|<psi|psi>|); ignores `amplitude`."""
    v = C.c_double()
    call("qil_norm", psi.ctx.handle, psi.handle, C.byref(v))
    return float(v.value)


def _grid_args(psi, site_mode, out_bit):
    n = psi.nsites_flat
    mode = np.ascontiguousarray(site_mode, dtype=np.int64)
    if mode.ndim != 1 or mode.shape[0] != n:
        raise ArgumentError(f"coefficient_grid: expected {n} site modes, got {mode.shape}")
    if mode.size and (mode.min() < 0 or mode.max() > 2):
        raise ArgumentError("coefficient_grid: site modes must be 0, 1 (fixed bit) or 2 (free)")
    mode8 = mode.astype(np.uint8)
    F = int((mode8 == 2).sum())
    ob = None
    if out_bit is not None:
        ob = np.ascontiguousarray(out_bit, dtype=np.int32)
        if ob.shape != (F,) or sorted(ob.tolist()) != list(range(F)):
            raise ArgumentError(f"coefficient_grid: out_bit must be a permutation of 0..{F - 1}")
    return mode8, F, ob


def coefficient_grid(psi, site_mode, out_bit=None):
    """All 2^F coefficients over the free sites (site_mode[i] == 2; 0/1 = fixed bit) in one call -- the grids the
    reference evaluates by looping `coefficient` (docs/src/tutorials/zt.jl:152-157, 283-411).  The j-th free
    site (in site order) lands on bit out_bit[j] of the output index; default big-endian (mps.jl:633-645)."""
    mode8, F, ob = _grid_args(psi, site_mode, out_bit)
    if F > 30:
        raise UnsupportedError("coefficient_grid: refusing to materialise more than 2^30 amplitudes on the host")
    out = np.empty(2**F, dtype=_np_dtype(psi.is_complex))
    call("qil_coefficient_grid", psi.ctx.handle, psi.handle, C.c_void_p(mode8.ctypes.data),
         C.c_void_p(ob.ctypes.data) if ob is not None else None, C.c_void_p(out.ctypes.data))
    return out


def coefficient_grid_dev(psi, site_mode, d_out, out_bit=None):
    """Same with a device output buffer (2^F scalars); stream-ordered on the context's stream."""
    mode8, F, ob = _grid_args(psi, site_mode, out_bit)
    call("qil_coefficient_grid_dev", psi.ctx.handle, psi.handle, C.c_void_p(mode8.ctypes.data),
         C.c_void_p(ob.ctypes.data) if ob is not None else None, C.c_void_p(int(d_out)))
    return F


def mps_to_vector(psi, reverse=False):
    """mps_to_vector(psi; reverse=false) (mps.jl:716-743): every coefficient, MSB-first by default,
    bit-reversed ordering with reverse=true.  One dense-grid evaluation with all sites free."""
    n = psi.nsites_flat
    return coefficient_grid(psi, np.full(n, 2), out_bit=np.arange(n) if reverse else None)


def pole_scan_modes(n, k0, l0, log2_k, log2_l, stride_log2_k=0, stride_log2_l=0):
    """Site modes / output bits of a (k, l) block on a paired-register chain of 2n sites (main_j = bit j-1 of k,
    copy_j = bit j-1 of l, LSB first; test/test_zt_transformer.jl:20-62, docs/src/tutorials/zt.jl:152-157):
    k = k0 + a * 2^stride_log2_k, a in [0, 2^log2_k); l likewise.  k0 / l0 must have zero bits on the free range.
    The output index is a * 2^log2_l + b, i.e. a row-major (k, l) table."""
    mode = np.zeros(2 * n, dtype=np.uint8)
    out_bit = []
    for j in range(n):
        for reg, (base, lg, st) in enumerate(((k0, log2_k, stride_log2_k), (l0, log2_l, stride_log2_l))):
            site = 2 * j + reg
            if st <= j < st + lg:
                if (base >> j) & 1:
                    raise ArgumentError("pole_scan: the block origin must be aligned to the block size")
                mode[site] = 2
                out_bit.append((j - st) + (log2_l if reg == 0 else 0))
            else:
                mode[site] = (base >> j) & 1
    return mode, np.asarray(out_bit, dtype=np.int32)


def pole_scan(psi, k0=0, l0=0, log2_k=None, log2_l=None, stride_log2_k=0, stride_log2_l=0):
    """chi[a, b] = coefficient(psi, interleave(lsb(k), lsb(l))) on the block k = k0 + a * 2^stride_log2_k,
    l = l0 + b * 2^stride_log2_l (the coarse / fine / superfine scans of docs/src/tutorials/zt.jl:283-411)."""
    n = psi.nsites_flat // 2
    log2_k = n - stride_log2_k if log2_k is None else log2_k
    log2_l = n - stride_log2_l if log2_l is None else log2_l
    mode, ob = pole_scan_modes(n, k0, l0, log2_k, log2_l, stride_log2_k, stride_log2_l)
    return coefficient_grid(psi, mode, ob).reshape(2**log2_k, 2**log2_l)


class Uploader:
    """qil_uploader_* (include/qilcuda.h): ring of device buffers filled on the library's own copy stream.

        up = Uploader(ctx, nbytes)
        up.submit(ptr0, nbytes)                    # host address (pinned memory -> truly asynchronous)
        for ...:
            d_x = up.acquire()
            up.submit(ptr_next, nbytes)            # overlaps the work below
            psi = signal_mps_dev(ctx, d_x, N, ...)
            up.release()
    """

    def __init__(self, ctx, nbytes, depth=2):
        self.ctx = ctx
        self.handle = C.c_void_p()
        self.inflight = 0
        call("qil_uploader_create", ctx.handle, C.c_int64(int(nbytes)), int(depth), C.byref(self.handle))

    def submit(self, host_ptr, nbytes):
        call("qil_uploader_submit", self.handle, C.c_void_p(int(host_ptr)), C.c_int64(int(nbytes)))
        self.inflight += 1

    def acquire(self):
        p = C.c_void_p()
        call("qil_uploader_acquire", self.handle, C.byref(p))
        return p.value

    def release(self):
        call("qil_uploader_release", self.handle)
        self.inflight -= 1

    def close(self):
        if self.handle:
            _lib.load().qil_uploader_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def save(obj, path):
    """Write an MPS / ZTMPS / MPO to the QILTN001 container (include/qilcuda.h)."""
    fn = "qil_mpo_save" if isinstance(obj, SingleSiteMPO) else "qil_mps_save"
    call(fn, obj.handle, str(path).encode())


def load_mps(path, ctx=None, paired=False):
    ctx = ctx or default_context()
    h = _lib.c_mps()
    call("qil_mps_load", ctx.handle, str(path).encode(), C.byref(h))
    return (ZTMPS if paired else SignalMPS)(ctx, h)


def load_mpo(path, ctx=None, paired=False):
    ctx = ctx or default_context()
    h = _lib.c_mpo()
    call("qil_mpo_load", ctx.handle, str(path).encode(), C.byref(h))
    return (PairedSiteMPO if paired else SingleSiteMPO)(ctx, h)


def _argmax_result(psi, idx, av, val):
    v = complex(val[0], val[1]) if psi.is_complex else float(val[0])
    return int(idx.value), float(av.value), v


def coefficient_grid_argmax(psi, site_mode, out_bit=None):
    """(flat index, |value|, value) of the largest coefficient of a dense grid, reduced on the device."""
    mode8, F, ob = _grid_args(psi, site_mode, out_bit)
    idx, av, val = C.c_int64(), C.c_double(), (C.c_double * 2)()
    call("qil_coefficient_grid_argmax", psi.ctx.handle, psi.handle, C.c_void_p(mode8.ctypes.data),
         C.c_void_p(ob.ctypes.data) if ob is not None else None, C.byref(idx), C.byref(av), val)
    return _argmax_result(psi, idx, av, val)


def coefficients_argmax(psi, bits):
    """(row index, |value|, value) of the largest coefficient of a batch of bitstrings, reduced on the device."""
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    if b.ndim != 2 or b.shape[1] != psi.nsites_flat:
        raise ArgumentError(f"coefficient: expected B x {psi.nsites_flat} bits, got {b.shape}")
    idx, av, val = C.c_int64(), C.c_double(), (C.c_double * 2)()
    call("qil_coefficient_batch_argmax", psi.ctx.handle, psi.handle, C.c_void_p(b.ctypes.data), C.c_int64(b.shape[0]),
         C.byref(idx), C.byref(av), val)
    return _argmax_result(psi, idx, av, val)


def sum_sites(psi, mask):
    """Sum over every configuration of the masked sites (contract them with the all-ones vector); returns an MPS over
    the remaining sites."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    if m.shape != (psi.nsites_flat,):
        raise ArgumentError(f"sum_sites: expected {psi.nsites_flat} mask entries")
    h = _lib.c_mps()
    call("qil_mps_sum_sites", psi.ctx.handle, psi.handle, C.c_void_p(m.ctypes.data), C.byref(h))
    return SignalMPS(psi.ctx, h)


def laplace_coefficients(psi_out, dt):
    """L(s_k) ~ dt * sqrt(N) * sum_j <k_LSB, j | psi_out> for EVERY k (docs/src/tutorials/dt.jl:187-197 calls
    `coefficient` N times per value): the copy register is summed once on the device and the main register read out as
    a dense vector (main sites are LSB first, dt.jl:176-178)."""
    n = psi_out.nsites_flat // 2
    mask = np.zeros(2 * n, dtype=np.uint8)
    mask[1::2] = 1
    red = sum_sites(psi_out, mask)
    return dt * math.sqrt(2.0**n) * mps_to_vector(red, reverse=True)


def z_from_kl(k, l, n, omega_r, omega_i=2 * math.pi):
    """z = exp(-(omega_r k + i omega_i l) / N) (docs/src/tutorials/zt.jl:140-150)."""
    N = 2.0**n
    return np.exp(-(omega_r * np.asarray(k, dtype=np.float64) + 1j * omega_i * np.asarray(l, dtype=np.float64)) / N)


def kl_bits(ks, ls, n):
    """Interleaved LSB-first bitstrings of (k, l) pairs (zt.jl:152-157): uint8 [len, 2n]."""
    ks = np.asarray(ks, dtype=np.int64).reshape(-1)
    ls = np.asarray(ls, dtype=np.int64).reshape(-1)
    bits = np.zeros((ks.size, 2 * n), dtype=np.uint8)
    for j in range(n):
        bits[:, 2 * j] = (ks >> j) & 1
        bits[:, 2 * j + 1] = (ls >> j) & 1
    return bits


def pole_scan_argmax(psi, k0=0, l0=0, log2_k=None, log2_l=None, stride_log2_k=0, stride_log2_l=0):
    """(k, l, |chi|, chi) of the peak of an aligned strided (k, l) block, grid and arg-max both on the device."""
    n = psi.nsites_flat // 2
    log2_k = n - stride_log2_k if log2_k is None else log2_k
    log2_l = n - stride_log2_l if log2_l is None else log2_l
    mode, ob = pole_scan_modes(n, k0, l0, log2_k, log2_l, stride_log2_k, stride_log2_l)
    idx, av, v = coefficient_grid_argmax(psi, mode, ob)
    a, b = divmod(idx, 2**log2_l)
    return k0 + (a << stride_log2_k), l0 + (b << stride_log2_l), av, v


def pole_scan_list_argmax(psi, ks, ls):
    """(k, l, |chi|, chi) of the peak over the Cartesian product ks x ls (arbitrary index lists: the fine and superfine
    stages of the tutorial, zt.jl:345-415); first maximum in Julia's column-major order (k fastest)."""
    n = psi.nsites_flat // 2
    ks = np.asarray(ks, dtype=np.int64)
    ls = np.asarray(ls, dtype=np.int64)
    K, L = np.meshgrid(ks, ls, indexing="ij")
    order_k = K.T.reshape(-1)           # column-major: k varies fastest
    order_l = L.T.reshape(-1)
    idx, av, v = coefficients_argmax(psi, kl_bits(order_k, order_l, n))
    return int(order_k[idx]), int(order_l[idx]), av, v


def pole_scan_driver(psi_coarse, psi_fine, omega_r_coarse, omega_r_fine, omega_i=2 * math.pi, step_coarse_log2=12,
                     r_window=(1 - 1.6e-4, 1.0), theta_window=(-5e-3, 9e-3), n_fine=128, z_target=None, half=24):
    """The three-stage pole search of docs/src/tutorials/zt.jl:296-415 on two transformed ZTMPS (coarse omega_r and fine
    omega_r): strided coarse grid -> polar window near the unit circle -> full-resolution block around `z_target`
    (default: the fine-stage peak).  Every stage evaluates its grid and reduces |chi| on the device."""
    n = psi_coarse.nsites_flat // 2
    N = 2**n
    out = {}
    k, l, av, v = pole_scan_argmax(psi_coarse, 0, 0, n - step_coarse_log2, n - step_coarse_log2, step_coarse_log2,
                                   step_coarse_log2)
    out["coarse"] = dict(k=k, l=l, abs=av, z=complex(z_from_kl(k, l, n, omega_r_coarse, omega_i)))
    r_t = np.linspace(r_window[0], r_window[1], n_fine)
    ks = np.clip(np.round((-N / omega_r_fine) * np.log(r_t)).astype(np.int64), 0, N - 1)
    th = np.mod(np.linspace(theta_window[0], theta_window[1], n_fine), 2 * math.pi)
    ls = np.mod(np.round((N / omega_i) * th).astype(np.int64), N)
    k, l, av, v = pole_scan_list_argmax(psi_fine, ks, ls)
    out["fine"] = dict(k=k, l=l, abs=av, z=complex(z_from_kl(k, l, n, omega_r_fine, omega_i)))
    zt = out["fine"]["z"] if z_target is None else complex(z_target)
    th_t = (-np.angle(zt)) % (2 * math.pi)
    kc = int(np.clip(round((-N / omega_r_fine) * math.log(abs(zt))), 0, N - 1))
    lc = int(round((N / omega_i) * th_t)) % N
    ks2 = np.arange(kc - half, kc + half + 1)
    ks2 = ks2[(ks2 >= 0) & (ks2 < N)]
    ls2 = np.mod(np.arange(lc - half, lc + half + 1), N)
    k, l, av, v = pole_scan_list_argmax(psi_fine, ks2, ls2)
    out["superfine"] = dict(k=k, l=l, abs=av, z=complex(z_from_kl(k, l, n, omega_r_fine, omega_i)))
    return out


# ------------------------------------------------------------------------------------------
# transform MPOs (src/transforms/*.jl)
# ------------------------------------------------------------------------------------------
def _n_from(arg):
    return len(arg) if isinstance(arg, (SignalMPS, ZTMPS)) else int(arg)


def build_qft_mpo(n_or_psi, cutoff=1e-14, maxdim=1000, ctx=None):
    """build_qft_mpo(n, sites; cutoff=1e-14, maxdim=1000) / build_qft_mpo(psi; ...) (qft_transformer.jl:121-165)."""
    ctx = ctx or (n_or_psi.ctx if isinstance(n_or_psi, SignalMPS) else default_context())
    h = _lib.c_mpo()
    call("qil_build_qft_mpo", ctx.handle, _n_from(n_or_psi), float(cutoff), C.c_int64(_maxdim_arg(maxdim)), C.byref(h))
    return SingleSiteMPO(ctx, h)


def build_dt_mpo(n_or_psi, omega_r, cutoff=1e-14, maxdim=1000, ctx=None):
    """build_dt_mpo(n, wr, sites_main, sites_copy; ...) / build_dt_mpo(psi::ZTMPS, wr; ...) (dt_transformer.jl:312-412)."""
    ctx = ctx or (n_or_psi.ctx if isinstance(n_or_psi, SignalMPS) else default_context())
    h = _lib.c_mpo()
    call("qil_build_dt_mpo", ctx.handle, _n_from(n_or_psi), float(omega_r), float(cutoff),
         C.c_int64(_maxdim_arg(maxdim)), C.byref(h))
    return PairedSiteMPO(ctx, h)


def build_zt_mpo(n_or_psi, omega_r, cutoff=1e-14, maxdim=1000, ctx=None):
    """build_zt_mpo(n, wr, sites_main, sites_copy; ...) / build_zt_mpo(psi::ZTMPS, wr; ...) (zt_transformer.jl:41-112)."""
    ctx = ctx or (n_or_psi.ctx if isinstance(n_or_psi, SignalMPS) else default_context())
    h = _lib.c_mpo()
    call("qil_build_zt_mpo", ctx.handle, _n_from(n_or_psi), float(omega_r), float(cutoff),
         C.c_int64(_maxdim_arg(maxdim)), C.byref(h))
    return PairedSiteMPO(ctx, h)


# ------------------------------------------------------------------------------------------
# dense factorizations (the ITensors calls of the path)
# ------------------------------------------------------------------------------------------
def qr(A, positive=False, ctx=None):
    ctx = ctx or default_context()
    A = np.asarray(A)
    ic = np.iscomplexobj(A)
    A = np.ascontiguousarray(A, dtype=_np_dtype(ic))
    m, n = A.shape
    k = min(m, n)
    Q = np.empty((m, k), dtype=A.dtype)
    R = np.empty((k, n), dtype=A.dtype)
    call("qil_qr", ctx.handle, int(ic), C.c_int64(m), C.c_int64(n), C.c_void_p(A.ctypes.data), int(positive),
         C.c_void_p(Q.ctypes.data), C.c_void_p(R.ctypes.data))
    return Q, R


def svd_trunc(A, cutoff=0.0, maxdim=None, mindim=1, ctx=None):
    ctx = ctx or default_context()
    A = np.asarray(A)
    ic = np.iscomplexobj(A)
    A = np.ascontiguousarray(A, dtype=_np_dtype(ic))
    m, n = A.shape
    k = min(m, n)
    U = np.empty(m * k, dtype=A.dtype)
    S = np.empty(k, dtype=np.float64)
    Vh = np.empty(k * n, dtype=A.dtype)
    r = C.c_int64(0)
    call("qil_svd_trunc", ctx.handle, int(ic), C.c_int64(m), C.c_int64(n), C.c_void_p(A.ctypes.data), float(cutoff),
         C.c_int64(_maxdim_arg(maxdim)), C.c_int64(mindim), C.byref(r), C.c_void_p(U.ctypes.data),
         C.c_void_p(S.ctypes.data), C.c_void_p(Vh.ctypes.data))
    r = int(r.value)
    return U[: m * r].reshape(m, r), S[:r], Vh[: r * n].reshape(r, n)


def rsvd(A, k=20, p=10, q=0, random_seed=1234, cutoff=1e-15, maxdim=None, mindim=1, normal_stream=None, ctx=None,
         **_ignored):
    """rsvd(A, Linds...; k=20, p=10, q=0, random_seed=1234, cutoff=1e-15, maxdim=k, mindim=1) on a matrix
    (src/linalg/rsvd.jl:38-121) -> (U, S, Vh)."""
    ctx = ctx or default_context()
    A = np.asarray(A)
    if A.ndim != 2 or A.shape[0] == 0 or A.shape[1] == 0:
        raise ErrorException("In `rsvd`, left or right index set is empty.")
    ic = np.iscomplexobj(A)
    A = np.ascontiguousarray(A, dtype=_np_dtype(ic))
    m, n = A.shape
    l = min(k + p, m, n)
    U = np.empty(m * l, dtype=A.dtype)
    S = np.empty(l, dtype=np.float64)
    Vh = np.empty(l * n, dtype=A.dtype)
    r = C.c_int64(0)
    if normal_stream is not None:
        st = np.ascontiguousarray(normal_stream, dtype=A.dtype).reshape(-1)
        sp, slen = C.c_void_p(st.ctypes.data), st.size
    else:
        sp, slen = None, 0
    call("qil_rsvd", ctx.handle, int(ic), C.c_int64(m), C.c_int64(n), C.c_void_p(A.ctypes.data), int(k), int(p), int(q),
         C.c_int64(random_seed), float(cutoff),
         C.c_int64(k if maxdim is None else (k + p if maxdim >= BIG else int(maxdim))),   # typemax(Int): cap is k+p
         C.c_int64(mindim), sp, C.c_int64(slen), C.byref(r), C.c_void_p(U.ctypes.data), C.c_void_p(S.ctypes.data),
         C.c_void_p(Vh.ctypes.data))
    r = int(r.value)
    return U[: m * r].reshape(m, r), S[:r], Vh[: r * n].reshape(r, n)


# ------------------------------------------------------------------------------------------
# device-resident entry points (inputs already in HBM; pointers are raw CUDA device addresses)
# ------------------------------------------------------------------------------------------
def signal_mps_dev(ctx, d_x, N, is_complex, method="rsvd", cutoff=1e-15, maxdim=None, k=20, p=10, q=0,
                   random_seed=1234, mindim=1, adaptive=False, normal_stream_dev=None, stream_len=0):
    """signal_mps on a signal that already lives on the device (stream-ordered, no host copies).
    normal_stream_dev / stream_len: optional device pointer to the host-drawn normal stream (Omega) and its length."""
    h = _lib.c_mps()
    if method == "svd":
        call("qil_encode_svd_dev", ctx.handle, int(is_complex), C.c_void_p(int(d_x)), C.c_int64(N), float(cutoff),
             C.c_int64(_maxdim_arg(maxdim)), C.byref(h))
    else:
        call("qil_encode_rsvd_dev", ctx.handle, int(is_complex), C.c_void_p(int(d_x)), C.c_int64(N), int(k), int(p),
             int(q), C.c_int64(random_seed), float(cutoff), C.c_int64(_maxdim_arg(maxdim)), C.c_int64(mindim),
             C.c_void_p(int(normal_stream_dev)) if normal_stream_dev else None, C.c_int64(int(stream_len)),
             C.c_int64(RSVD_ADAPTIVE if adaptive else 0), C.byref(h))
    return SignalMPS(ctx, h)


def signal_mps_batch_dev(ctx, d_x, N, count, is_complex, cutoff=1e-15, maxdim=None, k=20, p=10, q=0, random_seed=1234,
                         mindim=1, workers=16, adaptive=False, normal_stream_dev=None, stream_len=0):
    """signal_mps(x_b; method=:rsvd) for `count` signals of N samples stored back to back on the device; the
    independent encodes run concurrently on `workers` streams.  Returns a list of SignalMPS."""
    hs = (_lib.c_mps * int(count))()
    call("qil_encode_rsvd_batch_dev", ctx.handle, int(is_complex), C.c_void_p(int(d_x)), C.c_int64(N), C.c_int64(count),
         int(k), int(p), int(q), C.c_int64(random_seed), float(cutoff), C.c_int64(_maxdim_arg(maxdim)),
         C.c_int64(mindim), int(workers), C.c_void_p(int(normal_stream_dev)) if normal_stream_dev else None,
         C.c_int64(int(stream_len)), C.c_int64(RSVD_ADAPTIVE if adaptive else 0), hs)
    return [SignalMPS(ctx, _lib.c_mps(h)) for h in hs]


def ztmps_from_mps(psi, cutoff=1e-10, maxdim=None):
    """The copy-tensor split of signal_ztmps (SignalConverters.jl:258-277) applied to an encoded SignalMPS."""
    h = _lib.c_mps()
    call("qil_ztmps_split", psi.ctx.handle, psi.handle, float(cutoff), C.c_int64(_maxdim_arg(maxdim)), C.byref(h))
    return ZTMPS(psi.ctx, h)


def coefficients_dev(psi, d_bits, B, d_out):
    """Batched coefficient with device-resident bits (uint8[B][n]) and output (B scalars)."""
    call("qil_coefficient_batch_dev", psi.ctx.handle, psi.handle, C.c_void_p(int(d_bits)), C.c_int64(B),
         C.c_void_p(int(d_out)))
