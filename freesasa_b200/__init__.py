"""freesasa_b200 — B200-native engine for the FreeSASA hot path (neighbour search + Lee-Richards /
Shrake-Rupley surface integration), behind the reference's own C entry points.

Layers (bottom up):

* ``csrc/libfsb200.so``             hand-written sm_100a CUDA kernels + the C ABI of include/fsb200.h
* ``csrc/libfreesasa_b200_host.so`` C host layer exporting the reference's hot-path functions under
                                    their own names (freesasa_calc_coord, freesasa_calc,
                                    freesasa_lee_richards, freesasa_shrake_rupley, ...), see
                                    include/freesasa_b200_host.h
* ``structure.py``                  ctypes binding of the structure / classifier / result-tree / selection API of the
                                    host layer (PDB text -> structure -> SASA -> areas); binds this library or, in
                                    the tests, the compiled reference through the same code
* this module                       a thin ctypes mirror of that C interface for Python callers, tests and
                                    bench.py.  ``calc_coord`` goes through the C host layer exactly as a C
                                    program would; ``Engine`` talks to the C ABI directly (explicit
                                    context, device-resident buffers, sharding, statistics).

There is no CPU implementation here.  If the CUDA library has not been built, or no B200 is visible,
every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import workloads  # noqa: F401  (numpy-only synthetic inputs)

__all__ = [
    "LEE_RICHARDS", "SHRAKE_RUPLEY", "FP32", "FP64", "Parameters", "Result", "Engine", "Stats",
    "available", "calc_batch", "calc_coord", "calc_coord_batch", "calc_multi", "default_parameters", "device_count",
    "library_paths", "multi_stats", "trim", "workloads", "IpcBuffer", "MultiStats",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
# FSB200_ENGINE_LIB: experiment hook of tests/tools/ab_variants.py (a variant build of the same sources); the host layer
# always links the product library.
_LIB_PATH = os.environ.get("FSB200_ENGINE_LIB") or os.path.join(_HERE, "csrc", "libfsb200.so")
_HOST_PATH = os.path.join(_HERE, "csrc", "libfreesasa_b200_host.so")

LEE_RICHARDS = 0  # enum freesasa_algorithm (reference src/freesasa.h:89-92)
SHRAKE_RUPLEY = 1
FP32 = 0
FP64 = 1

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_dpp = ctypes.POINTER(_dp)


class Parameters(ctypes.Structure):
    """struct freesasa_parameters (reference src/freesasa.h:232-238)."""

    _fields_ = [
        ("alg", ctypes.c_int),
        ("probe_radius", ctypes.c_double),
        ("shrake_rupley_n_points", ctypes.c_int),
        ("lee_richards_n_slices", ctypes.c_int),
        ("n_threads", ctypes.c_int),
    ]

    def resolution(self) -> int:
        return self.lee_richards_n_slices if self.alg == LEE_RICHARDS else self.shrake_rupley_n_points


class _CResult(ctypes.Structure):
    """struct freesasa_result (reference src/freesasa.h:267-272)."""

    _fields_ = [("total", ctypes.c_double), ("sasa", _dp), ("n_atoms", ctypes.c_int), ("parameters", Parameters)]


class Stats(ctypes.Structure):
    """struct fsb200_stats (include/fsb200.h)."""

    _fields_ = [
        ("n_atoms", ctypes.c_int), ("n_structures", ctypes.c_int), ("n_items", ctypes.c_int),
        ("n_overflow", ctypes.c_int), ("max_neighbours", ctypes.c_int), ("n_certified", ctypes.c_int),
        ("kernel_launches", ctypes.c_int),
        ("device_ms", ctypes.c_float), ("integrate_ms", ctypes.c_float),
        ("host_stage_ms", ctypes.c_float), ("host_total_ms", ctypes.c_float), ("n_marginal", ctypes.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


MAX_DEVICES = 16  # FSB200_MAX_DEVICES
SECOND_PASS = 1   # FSB200_SECOND_PASS


class MultiStats(ctypes.Structure):
    """struct fsb200_multi_stats (include/fsb200.h)."""

    _fields_ = [
        ("n_devices", ctypes.c_int), ("n_atoms", ctypes.c_int), ("n_structures", ctypes.c_int), ("n_certified", ctypes.c_int),
        ("upload_ms", ctypes.c_float), ("compute_ms", ctypes.c_float), ("download_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
        ("integrate_ms", ctypes.c_float * MAX_DEVICES), ("device_ms", ctypes.c_float * MAX_DEVICES),
    ]

    def as_dict(self):
        n = self.n_devices
        return {"n_devices": n, "n_atoms": self.n_atoms, "n_structures": self.n_structures, "upload_ms": self.upload_ms,
                "compute_ms": self.compute_ms, "download_ms": self.download_ms, "total_ms": self.total_ms,
                "integrate_ms": list(self.integrate_ms)[:n], "device_ms": list(self.device_ms)[:n]}


class Result:
    """Python view of a freesasa_result: per-atom SASA (Å²), total, parameters used."""

    def __init__(self, sasa: np.ndarray, total: float, parameters: Parameters):
        self.sasa = sasa
        self.total = total
        self.n_atoms = int(sasa.shape[0])
        self.parameters = parameters


def default_parameters() -> Parameters:
    """freesasa_default_parameters (reference src/freesasa.c:38-43): L&R, 1.4 Å, 100 points, 20 slices, 2 threads."""
    return Parameters(LEE_RICHARDS, 1.4, 100, 20, 2)


def library_paths():
    return {"engine": _LIB_PATH, "host": _HOST_PATH}


_lib = None
_host = None


def _missing(path):
    return RuntimeError(
        f"{path} has not been built (run `python -c 'import __graft_entry__ as g; g.build()'` or "
        f"`python freesasa_b200/build.py`); freesasa_b200 has no CPU fallback"
    )


def _engine_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise _missing(_LIB_PATH)
        L = ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_GLOBAL)  # the host layer links against it
        vp = ctypes.c_void_p
        L.fsb200_last_error.restype = ctypes.c_char_p
        L.fsb200_version.restype = ctypes.c_char_p
        L.fsb200_launch_count.restype = ctypes.c_ulonglong
        L.fsb200_ctx_create.restype = vp
        L.fsb200_ctx_create.argtypes = [ctypes.c_int]
        L.fsb200_ctx_destroy.restype = None
        L.fsb200_ctx_destroy.argtypes = [vp]
        L.fsb200_ctx_set_precision.argtypes = [vp, ctypes.c_int]
        L.fsb200_ctx_set_certificate.argtypes = [vp, ctypes.c_int]
        L.fsb200_ctx_stats.argtypes = [vp, ctypes.POINTER(Stats)]
        L.fsb200_ctx_calc.argtypes = [vp, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.fsb200_ctx_calc_batch.argtypes = [vp, ctypes.c_int, ctypes.c_int, _ip, _dpp, _dpp, _dpp, ctypes.c_double, ctypes.c_int]
        L.fsb200_calc_batch.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _dpp, _dpp, _dpp, ctypes.c_double, ctypes.c_int]
        L.fsb200_lr.argtypes = [_dp, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.fsb200_sr.argtypes = [_dp, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.fsb200_ctx_calc_device.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, _ip, ctypes.c_double,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
        L.fsb200_ctx_calc_device_async.argtypes = L.fsb200_ctx_calc_device.argtypes
        L.fsb200_ctx_finish.argtypes = [vp]
        L.fsb200_ctx_set_peer_outputs.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp)]
        L.fsb200_ctx_set_peer_zero_skipping.argtypes = [vp, ctypes.c_int]
        L.fsb200_ctx_peer_barrier.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp), vp]
        L.fsb200_ctx_peer_barrier_status.argtypes = [vp]
        L.fsb200_ctx_generation.restype = ctypes.c_ulonglong
        L.fsb200_ctx_generation.argtypes = [vp]
        L.fsb200_ipc_alloc.argtypes = [ctypes.c_int, ctypes.c_ulonglong, ctypes.POINTER(vp), ctypes.c_char_p]
        L.fsb200_ipc_open.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(vp)]
        L.fsb200_ipc_close.argtypes = [ctypes.c_int, vp]
        L.fsb200_ipc_free.argtypes = [ctypes.c_int, vp]
        L.fsb200_calc_multi.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _dpp, _dpp, _dpp, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.fsb200_get_multi_stats.argtypes = [ctypes.POINTER(MultiStats)]
        L.fsb200_trim.argtypes = []
        L.fsb200_ctx_unpermute.argtypes = [vp, vp, vp, ctypes.c_int, vp]
        L.fsb200_ctx_neighbour_counts.argtypes = [vp, _ip, _dp, _dp, ctypes.c_int, ctypes.c_double]
        for f in ("fsb200_shard_begin", "fsb200_shard_end"):
            getattr(L, f).argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def _host_lib():
    global _host
    if _host is None:
        _engine_lib()
        if not os.path.exists(_HOST_PATH):
            raise _missing(_HOST_PATH)
        H = ctypes.CDLL(_HOST_PATH)
        H.freesasa_calc_coord.restype = ctypes.POINTER(_CResult)
        H.freesasa_calc_coord.argtypes = [_dp, _dp, ctypes.c_int, ctypes.POINTER(Parameters)]
        H.freesasa_result_free.restype = None
        H.freesasa_result_free.argtypes = [ctypes.POINTER(_CResult)]
        H.freesasa_calc_coord_batch.argtypes = [ctypes.c_int, _dpp, _dpp, _ip, ctypes.POINTER(Parameters),
                                                ctypes.POINTER(ctypes.POINTER(_CResult))]
        H.freesasa_set_verbosity.argtypes = [ctypes.c_int]
        _host = H
    return _host


def _last_error() -> str:
    return _engine_lib().fsb200_last_error().decode(errors="replace")


def available() -> bool:
    """True iff the CUDA library is built and a compute-capability-10.x device is visible."""
    if not os.path.exists(_LIB_PATH):
        return False
    return bool(_engine_lib().fsb200_available())


def launch_count() -> int:
    return int(_engine_lib().fsb200_launch_count())


def _f64(a, n3: Optional[int] = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if n3 is not None and a.shape[0] != n3:
        raise ValueError(f"expected {n3} values, got {a.shape[0]}")
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _ptr_array(arrays: Sequence[np.ndarray]):
    return (_dp * len(arrays))(*[_ptr(a) for a in arrays])


# ------------------------------------------------------------------------------------------------------
# reference-facing calls (through the C host layer)
# ------------------------------------------------------------------------------------------------------
def calc_coord(xyz, radii, parameters: Optional[Parameters] = None) -> Result:
    """freesasa_calc_coord(xyz, radii, n, parameters) — reference src/freesasa.c:122-142.

    ``xyz``: n×3 (or flat 3n) coordinates, ``radii``: n van der Waals radii without the probe.
    Returns a Result; raises RuntimeError where the C function returns NULL."""
    H = _host_lib()
    radii = _f64(radii)
    n = int(radii.shape[0])
    if n <= 0:
        raise ValueError("freesasa_calc_coord requires n > 0")  # the C function asserts (src/freesasa.c:133)
    xyz = _f64(xyz, 3 * n)
    p = parameters if parameters is not None else default_parameters()
    res = H.freesasa_calc_coord(_ptr(xyz), _ptr(radii), n, ctypes.byref(p))
    if not res:
        raise RuntimeError("freesasa_calc_coord failed: " + _last_error())
    try:
        sasa = np.ctypeslib.as_array(res.contents.sasa, shape=(n,)).copy()
        out = Result(sasa, float(res.contents.total), res.contents.parameters)
        out.parameters = Parameters.from_buffer_copy(res.contents.parameters)
    finally:
        H.freesasa_result_free(res)
    return out


def calc_coord_batch(structures: Sequence, parameters: Optional[Parameters] = None):
    """Additive freesasa_calc_coord_batch(): a list of (xyz, radii) -> list of Result, one device pass."""
    H = _host_lib()
    p = parameters if parameters is not None else default_parameters()
    rad = [_f64(r) for _, r in structures]
    xyz = [_f64(x, 3 * r.shape[0]) for (x, _), r in zip(structures, rad)]
    n = len(rad)
    counts = (ctypes.c_int * n)(*[int(r.shape[0]) for r in rad])
    results = (ctypes.POINTER(_CResult) * n)()
    rc = H.freesasa_calc_coord_batch(n, _ptr_array(xyz), _ptr_array(rad), counts, ctypes.byref(p), results)
    if rc != 0:
        raise RuntimeError("freesasa_calc_coord_batch failed: " + _last_error())
    out = []
    for k in range(n):
        r = results[k].contents
        out.append(Result(np.ctypeslib.as_array(r.sasa, shape=(int(counts[k]),)).copy(), float(r.total),
                          Parameters.from_buffer_copy(r.parameters)))
        H.freesasa_result_free(results[k])
    return out


def calc_batch(alg: int, structures: Sequence, probe: float = 1.4, resolution: int = 20):
    """fsb200_calc_batch() of the C ABI (context-free): list of (xyz, radii) -> list of per-atom SASA arrays.
    Batches of >= 400k atoms are worked through as overlapped sub-batches on two pooled contexts."""
    L = _engine_lib()
    rad = [_f64(r) for _, r in structures]
    xyz = [_f64(x, 3 * r.shape[0]) for (x, _), r in zip(structures, rad)]
    outs = [np.empty(r.shape[0], dtype=np.float64) for r in rad]
    counts = (ctypes.c_int * len(rad))(*[int(r.shape[0]) for r in rad])
    if L.fsb200_calc_batch(int(alg), len(rad), counts, _ptr_array(xyz), _ptr_array(rad), _ptr_array(outs), float(probe),
                           int(resolution)) != 0:
        raise RuntimeError("fsb200_calc_batch failed: " + _last_error())
    return outs


def calc_multi(alg: int, structures: Sequence, probe: float = 1.4, resolution: int = 20, n_devices: int = 0):
    """fsb200_calc_multi() of the C ABI: one process, several GPUs.  One structure: inputs replicated, outputs partitioned
    over the devices; several structures: dealt to the devices by size.  Returns the per-atom SASA arrays."""
    L = _engine_lib()
    rad = [_f64(r) for _, r in structures]
    xyz = [_f64(x, 3 * r.shape[0]) for (x, _), r in zip(structures, rad)]
    outs = [np.empty(r.shape[0], dtype=np.float64) for r in rad]
    counts = (ctypes.c_int * len(rad))(*[int(r.shape[0]) for r in rad])
    if L.fsb200_calc_multi(int(alg), len(rad), counts, _ptr_array(xyz), _ptr_array(rad), _ptr_array(outs), float(probe),
                           int(resolution), int(n_devices)) != 0:
        raise RuntimeError("fsb200_calc_multi failed: " + _last_error())
    return outs


def multi_stats() -> dict:
    s = MultiStats()
    _engine_lib().fsb200_get_multi_stats(ctypes.byref(s))
    return s.as_dict()


def device_count() -> int:
    return int(_engine_lib().fsb200_device_count())


def trim() -> int:
    """fsb200_trim(): destroy every idle pooled context (device scratch + pinned staging)."""
    return int(_engine_lib().fsb200_trim())


def _stream_handle(torch_stream) -> int:
    """cudaStream_t of a torch stream for the C ABI, where NULL means "the context's own stream": torch's legacy default
    stream has the handle 0, so it is passed as cudaStreamLegacy (1) — otherwise work queued through torch on the default
    stream and the engine's kernels would run on two unrelated streams."""
    return int(torch_stream.cuda_stream) or 1


class IpcBuffer:
    """A device buffer other processes can map (CUDA IPC): the symmetric output / flag buffers of the fused all-gather.
    ``handle`` (64 bytes) is what a peer passes to ``IpcBuffer.open``."""

    def __init__(self, device: int, nbytes: int, _ptr=None, _handle=None, _owner=True):
        L = _engine_lib()
        self.device, self.nbytes, self._owner = int(device), int(nbytes), _owner
        if _ptr is None:
            p, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            if L.fsb200_ipc_alloc(self.device, self.nbytes, ctypes.byref(p), h) != 0:
                raise RuntimeError("fsb200_ipc_alloc failed: " + _last_error())
            self.ptr, self.handle = p.value, h.raw
        else:
            self.ptr, self.handle = _ptr, _handle

    @classmethod
    def open(cls, device: int, handle: bytes, nbytes: int):
        p = ctypes.c_void_p()
        if _engine_lib().fsb200_ipc_open(int(device), handle, ctypes.byref(p)) != 0:
            raise RuntimeError("fsb200_ipc_open failed: " + _last_error())
        return cls(device, nbytes, _ptr=p.value, _handle=handle, _owner=False)

    def tensor(self, dtype, count):
        """torch view of the buffer (no copy) on this process's device."""
        import torch

        class _Iface:
            pass

        it = _Iface()
        itemsize = torch.tensor([], dtype=dtype).element_size()
        typestr = {torch.float64: "<f8", torch.int32: "<i4"}[dtype]
        it.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(self.ptr), False), "version": 2}
        assert count * itemsize <= self.nbytes
        t = torch.as_tensor(it, device=torch.device("cuda", self.device))
        if t.data_ptr() != int(self.ptr):
            raise RuntimeError("IpcBuffer.tensor: torch copied the buffer instead of aliasing it")
        t._fsb200_keepalive = self
        return t

    def close(self):
        if self.ptr:
            L = _engine_lib()
            (L.fsb200_ipc_free if self._owner else L.fsb200_ipc_close)(self.device, ctypes.c_void_p(self.ptr))
            self.ptr = None


# ------------------------------------------------------------------------------------------------------
# engine-level access (explicit context)
# ------------------------------------------------------------------------------------------------------
class Engine:
    """One fsb200 context = one device + one stream + cached device scratch."""

    def __init__(self, device: int = 0, precision: int = FP32):
        L = _engine_lib()
        self._L = L
        self._ctx = L.fsb200_ctx_create(int(device))
        if not self._ctx:
            raise RuntimeError("fsb200_ctx_create failed: " + _last_error())
        self.device = int(device)
        self.set_precision(precision)

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.fsb200_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {_last_error()}")

    def set_precision(self, precision: int):
        self._check(self._L.fsb200_ctx_set_precision(self._ctx, int(precision)), "fsb200_ctx_set_precision")
        self.precision = int(precision)

    def set_certificate(self, on: bool):
        """Buried-atom certificate on/off (default on; never changes a result, only the time)."""
        self._check(self._L.fsb200_ctx_set_certificate(self._ctx, 1 if on else 0), "fsb200_ctx_set_certificate")

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.fsb200_ctx_stats(self._ctx, ctypes.byref(s)), "fsb200_ctx_stats")
        return s.as_dict()

    def calc(self, alg: int, xyz, radii, probe: float = 1.4, resolution: int = 20) -> np.ndarray:
        radii = _f64(radii)
        n = int(radii.shape[0])
        xyz = _f64(xyz, 3 * n)
        out = np.empty(n, dtype=np.float64)
        self._check(self._L.fsb200_ctx_calc(self._ctx, int(alg), _ptr(out), _ptr(xyz), _ptr(radii), n, float(probe),
                                            int(resolution)), "fsb200_ctx_calc")
        return out

    def calc_batch(self, alg: int, structures: Sequence, probe: float = 1.4, resolution: int = 20):
        rad = [_f64(r) for _, r in structures]
        xyz = [_f64(x, 3 * r.shape[0]) for (x, _), r in zip(structures, rad)]
        outs = [np.empty(r.shape[0], dtype=np.float64) for r in rad]
        counts = (ctypes.c_int * len(rad))(*[int(r.shape[0]) for r in rad])
        self._check(self._L.fsb200_ctx_calc_batch(self._ctx, int(alg), len(rad), counts, _ptr_array(xyz), _ptr_array(rad),
                                                  _ptr_array(outs), float(probe), int(resolution)), "fsb200_ctx_calc_batch")
        return outs

    def neighbour_counts(self, xyz, radii, probe: float = 1.4) -> np.ndarray:
        radii = _f64(radii)
        n = int(radii.shape[0])
        xyz = _f64(xyz, 3 * n)
        out = np.empty(n, dtype=np.int32)
        self._check(self._L.fsb200_ctx_neighbour_counts(self._ctx, out.ctypes.data_as(_ip), _ptr(xyz), _ptr(radii), n,
                                                        float(probe)), "fsb200_ctx_neighbour_counts")
        self.last_certified = (out >> 30) & 1  # atoms settled by the buried-atom certificate in that pass
        return out & 0x3FFFFFFF

    # ---- device-resident path: torch CUDA tensors (float64) in, torch tensor out -------------------------
    def calc_device(self, alg: int, d_xyz, d_radii, probe: float = 1.4, resolution: int = 20, offsets=None,
                    shard=(0, 1), out=None, stream=None):
        """d_xyz [n,3] / d_radii [n] float64 CUDA tensors on this engine's device.  Returns the per-atom
        SASA tensor (caller order; with shard=(i, k>1) the owned slice of the SORTED order is filled,
        see include/fsb200.h).  Runs on ``stream`` (default: torch's current stream)."""
        import torch

        assert d_xyz.is_cuda and d_radii.is_cuda and d_xyz.dtype == torch.float64 and d_radii.dtype == torch.float64
        assert d_xyz.is_contiguous() and d_radii.is_contiguous()
        n = int(d_radii.shape[0])
        if out is None:
            out = torch.zeros(n, dtype=torch.float64, device=d_radii.device)
        if stream is None:
            stream = _stream_handle(torch.cuda.current_stream(d_radii.device))
        n_struct, off_p = 1, None
        if offsets is not None:
            off = np.ascontiguousarray(offsets, dtype=np.int32)
            n_struct, off_p = int(off.shape[0]) - 1, off.ctypes.data_as(_ip)
        self._check(self._L.fsb200_ctx_calc_device(self._ctx, int(alg), d_xyz.data_ptr(), d_radii.data_ptr(), n, n_struct,
                                                   off_p, float(probe), int(resolution), int(shard[0]), int(shard[1]),
                                                   out.data_ptr(), ctypes.c_void_p(stream)), "fsb200_ctx_calc_device")
        return out

    def calc_device_async(self, alg: int, d_xyz, d_radii, probe: float = 1.4, resolution: int = 20, offsets=None,
                          shard=(0, 1), out=None, stream=None):
        """First half of calc_device: validates and enqueues, does not wait.  Queue a collective on the same stream, then
        call finish()."""
        import torch

        assert d_xyz.is_cuda and d_radii.is_cuda and d_xyz.dtype == torch.float64 and d_radii.dtype == torch.float64
        assert d_xyz.is_contiguous() and d_radii.is_contiguous()
        n = int(d_radii.shape[0])
        if out is None:
            out = torch.zeros(n, dtype=torch.float64, device=d_radii.device)
        if stream is None:
            stream = _stream_handle(torch.cuda.current_stream(d_radii.device))
        n_struct, off_p = 1, None
        if offsets is not None:
            off = np.ascontiguousarray(offsets, dtype=np.int32)
            n_struct, off_p = int(off.shape[0]) - 1, off.ctypes.data_as(_ip)
        self._check(self._L.fsb200_ctx_calc_device_async(self._ctx, int(alg), d_xyz.data_ptr(), d_radii.data_ptr(), n, n_struct,
                                                         off_p, float(probe), int(resolution), int(shard[0]), int(shard[1]),
                                                         out.data_ptr(), ctypes.c_void_p(stream)), "fsb200_ctx_calc_device_async")
        return out

    def finish(self) -> int:
        """Second half: the one synchronisation.  Returns 0, or SECOND_PASS (work queued in between must be redone)."""
        rc = self._L.fsb200_ctx_finish(self._ctx)
        if rc not in (0, SECOND_PASS):
            raise RuntimeError("fsb200_ctx_finish failed: " + _last_error())
        return rc

    def set_peer_outputs(self, pointers: Sequence[int]):
        """Mirror every area into these device buffers (usually on other GPUs) from the kernel epilogue."""
        arr = (ctypes.c_void_p * max(1, len(pointers)))(*[ctypes.c_void_p(int(p)) for p in pointers])
        self._check(self._L.fsb200_ctx_set_peer_outputs(self._ctx, len(pointers), arr), "fsb200_ctx_set_peer_outputs")

    def set_peer_zero_skipping(self, on: bool):
        self._check(self._L.fsb200_ctx_set_peer_zero_skipping(self._ctx, 1 if on else 0), "fsb200_ctx_set_peer_zero_skipping")

    def peer_barrier(self, rank: int, world: int, flag_pointers: Sequence[int], stream=None):
        import torch

        if stream is None:
            stream = _stream_handle(torch.cuda.current_stream(self.device))
        arr = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p)) for p in flag_pointers])
        self._check(self._L.fsb200_ctx_peer_barrier(self._ctx, int(rank), int(world), arr, ctypes.c_void_p(stream)),
                    "fsb200_ctx_peer_barrier")

    def peer_barrier_status(self):
        self._check(self._L.fsb200_ctx_peer_barrier_status(self._ctx), "fsb200_ctx_peer_barrier_status")

    def generation(self) -> int:
        return int(self._L.fsb200_ctx_generation(self._ctx))

    def unpermute(self, d_sorted, out=None, stream=None):
        import torch

        n = int(d_sorted.shape[0])
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=d_sorted.device)
        if stream is None:
            stream = _stream_handle(torch.cuda.current_stream(d_sorted.device))
        self._check(self._L.fsb200_ctx_unpermute(self._ctx, d_sorted.data_ptr(), out.data_ptr(), n, ctypes.c_void_p(stream)),
                    "fsb200_ctx_unpermute")
        return out

    def shard_range(self, n: int, index: int, count: int):
        return int(self._L.fsb200_shard_begin(n, index, count)), int(self._L.fsb200_shard_end(n, index, count))
