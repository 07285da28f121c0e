"""ctypes binding of libtidk_cuda.so -- the same C ABI a Rust `tidk-cuda` crate
would bind (include/tidk_cuda.h)."""
from __future__ import annotations

import ctypes as C
import weakref
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TIDK_OK, TIDK_EARG, TIDK_ECUDA, TIDK_ENOMEM, TIDK_ENONASCII = 0, -1, -2, -3, -4

EXPORTS = [
    "tidk_cuda_abi_version", "tidk_cuda_create", "tidk_cuda_destroy", "tidk_cuda_last_error",
    "tidk_cuda_host_alloc", "tidk_cuda_host_free", "tidk_cuda_window_counts",
    "tidk_cuda_seqbuf_upload", "tidk_cuda_seqbuf_free", "tidk_cuda_window_counts_resident",
    "tidk_cuda_explore_runs", "tidk_cuda_explore_runs_resident", "tidk_cuda_free_runs",
    "tidk_cuda_last_explore_timings", "tidk_cuda_last_transfer", "tidk_cuda_last_kernel_time",
    "tidk_cuda_window_counts_resident_async", "tidk_cuda_wait", "tidk_cuda_seqbuf_planes_info",
]


class TidkCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tidk_cuda error {code}: {msg}")
        self.code = code


class Run(C.Structure):  # tidk_run
    _fields_ = [("record", C.c_uint32), ("slice", C.c_uint8), ("k", C.c_uint8),
                ("first_in_slice", C.c_uint8), ("pad", C.c_uint8),
                ("start", C.c_uint64), ("end", C.c_uint64), ("label_pos", C.c_uint64)]


RUN_DTYPE = np.dtype([("record", "<u4"), ("slice", "u1"), ("k", "u1"), ("first", "u1"), ("pad", "u1"),
                      ("start", "<u8"), ("end", "<u8"), ("label_pos", "<u8")])


def library_path() -> str:
    # TIDK_CUDA_LIB: kernel-tuning experiments load an alternative build of the same library
    return os.environ.get("TIDK_CUDA_LIB") or os.path.join(_HERE, "libtidk_cuda.so")


def load_library():
    """Load libtidk_cuda.so.  Raises if it has not been built: there is no
    Python or CPU implementation to fall back to."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise TidkCudaError(TIDK_ECUDA, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                                        " (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.tidk_cuda_abi_version.restype = C.c_uint32
    L.tidk_cuda_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.tidk_cuda_create.restype = C.c_int
    L.tidk_cuda_destroy.argtypes = [vp]
    L.tidk_cuda_destroy.restype = None
    L.tidk_cuda_last_error.argtypes = [vp]
    L.tidk_cuda_last_error.restype = C.c_char_p
    L.tidk_cuda_host_alloc.argtypes = [vp, C.c_size_t]
    L.tidk_cuda_host_alloc.restype = vp
    L.tidk_cuda_host_free.argtypes = [vp, vp]
    L.tidk_cuda_host_free.restype = None
    L.tidk_cuda_window_counts.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, C.POINTER(C.c_char_p), u32p,
                                          C.c_uint32, vp, vp]
    L.tidk_cuda_window_counts.restype = C.c_int
    L.tidk_cuda_seqbuf_upload.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(vp)]
    L.tidk_cuda_seqbuf_upload.restype = C.c_int
    L.tidk_cuda_seqbuf_free.argtypes = [vp, vp]
    L.tidk_cuda_seqbuf_free.restype = None
    L.tidk_cuda_window_counts_resident.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_char_p), u32p, C.c_uint32,
                                                   vp, vp, C.POINTER(C.c_float), u32p]
    L.tidk_cuda_window_counts_resident.restype = C.c_int
    L.tidk_cuda_window_counts_resident_async.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_char_p), u32p, C.c_uint32, vp, vp]
    L.tidk_cuda_window_counts_resident_async.restype = C.c_int
    L.tidk_cuda_seqbuf_planes_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_float), u64p]
    L.tidk_cuda_seqbuf_planes_info.restype = C.c_int
    L.tidk_cuda_wait.argtypes = [vp, C.POINTER(C.c_float), u32p]
    L.tidk_cuda_wait.restype = C.c_int
    L.tidk_cuda_explore_runs.argtypes = [vp, vp, vp, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, C.c_uint64,
                                         C.POINTER(C.POINTER(Run)), u64p]
    L.tidk_cuda_explore_runs.restype = C.c_int
    L.tidk_cuda_explore_runs_resident.argtypes = [vp, vp, C.c_double, C.c_uint32, C.c_uint32, C.c_uint64,
                                                  C.POINTER(C.POINTER(Run)), u64p, C.POINTER(C.c_float), u32p]
    L.tidk_cuda_explore_runs_resident.restype = C.c_int
    L.tidk_cuda_free_runs.argtypes = [C.POINTER(Run)]
    L.tidk_cuda_free_runs.restype = None
    L.tidk_cuda_last_transfer.argtypes = [vp, u64p, u64p, C.POINTER(C.c_int)]
    L.tidk_cuda_last_transfer.restype = C.c_int
    L.tidk_cuda_last_kernel_time.argtypes = [vp, C.POINTER(C.c_float), u32p]
    L.tidk_cuda_last_kernel_time.restype = C.c_int
    L.tidk_cuda_last_explore_timings.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    L.tidk_cuda_last_explore_timings.restype = C.c_int
    _LIB = L
    return L


def _qargs(queries: Sequence[bytes]):
    qs = [bytes(q) for q in queries]
    arr = (C.c_char_p * max(len(qs), 1))(*qs)
    lens = (C.c_uint32 * max(len(qs), 1))(*[len(q) for q in qs])
    return arr, lens, len(qs)


def n_windows(offsets: np.ndarray, window: int) -> np.ndarray:
    lens = (offsets[1:] - offsets[:-1]).astype(np.uint64)
    return (lens // np.uint64(window) + (lens % np.uint64(window) != 0)).astype(np.uint64)


class SeqBatch:
    """A batch of records resident in HBM (tidk_seqbuf)."""

    def __init__(self, ctx: "Context", handle, offsets: np.ndarray):
        self._ctx, self._h, self.offsets = ctx, handle, offsets
        ctx._batches.add(self)

    def planes_info(self):
        """(ready, build_ms, bytes) of the batch's resident bit-planes (K1, built once behind the upload)."""
        r, ms, b = C.c_int(0), C.c_float(0), C.c_uint64(0)
        self._ctx._L.tidk_cuda_seqbuf_planes_info(self._h, C.byref(r), C.byref(ms), C.byref(b))
        return bool(r.value), ms.value, b.value

    def free(self):
        if self._h and self._ctx._h:  # (a closed context has freed the batch already)
            self._ctx._L.tidk_cuda_seqbuf_free(self._ctx._h, self._h)
        self._h = None
        self._ctx._batches.discard(self)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """tidk_ctx: one CUDA device, one stream -- or, with a list of device ids, one context over several devices
    (tidk_cuda_create with n_devices > 1: every call is fanned out, include/tidk_cuda.h)."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        devs = [int(device)] if np.isscalar(device) else [int(d) for d in device]
        ids = (C.c_int * len(devs))(*devs)
        rc = self._L.tidk_cuda_create(ids, len(devs), C.byref(h))
        if rc != 0:
            raise TidkCudaError(rc, (self._L.tidk_cuda_last_error(None) or b"").decode())
        self._h = h
        self.last_kernel_ms = 0.0
        self.last_kernel_launches = 0
        self._pinned = []          # blocks handed out by pinned_empty
        self._batches = weakref.WeakSet()  # live SeqBatch handles

    def close(self):
        """Frees the resident batches and the pinned blocks of this context, then the context.  Arrays returned by
        pinned_empty must not be touched afterwards."""
        if self._h:
            for b in list(self._batches):
                b.free()
            for p in self._pinned:
                self._L.tidk_cuda_host_free(self._h, p)
            self._pinned = []
            self._L.tidk_cuda_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise TidkCudaError(rc, (self._L.tidk_cuda_last_error(self._h) or b"").decode())

    # -- pinned host memory -------------------------------------------------
    def pinned_empty(self, nbytes: int) -> np.ndarray:
        p = self._L.tidk_cuda_host_alloc(self._h, max(nbytes, 1))
        if not p:
            raise TidkCudaError(TIDK_ENOMEM, "tidk_cuda_host_alloc failed")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
        arr = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
        arr.flags.writeable = True
        self._pinned.append(p)
        return arr

    # -- search / find ------------------------------------------------------
    def upload(self, seq: np.ndarray, offsets: np.ndarray) -> SeqBatch:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        h = C.c_void_p()
        self._check(self._L.tidk_cuda_seqbuf_upload(self._h, seq.ctypes.data, offsets.ctypes.data,
                                                    len(offsets) - 1, C.byref(h)))
        return SeqBatch(self, h, offsets.copy())

    def window_counts(self, seq: np.ndarray, offsets: np.ndarray, window: int,
                      queries: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
        """tidk_cuda_window_counts: host buffers in, counts [record][query][window] out."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        qa, ql, nq = _qargs(queries)
        total = int(n_windows(offsets, window).sum()) * nq if window > 0 else 0
        fwd = np.zeros(max(total, 1), dtype=np.uint64)
        rev = np.zeros(max(total, 1), dtype=np.uint64)
        self._check(self._L.tidk_cuda_window_counts(self._h, seq.ctypes.data, offsets.ctypes.data,
                                                    len(offsets) - 1, window, qa, ql, nq,
                                                    fwd.ctypes.data, rev.ctypes.data))
        ms, nl = C.c_float(0), C.c_uint32(0)
        self._L.tidk_cuda_last_kernel_time(self._h, C.byref(ms), C.byref(nl))
        self.last_kernel_ms, self.last_kernel_launches = ms.value, nl.value
        return fwd[:total], rev[:total]

    def last_transfer(self):
        """(h2d_bytes, d2h_bytes, host_packed) of the last window_counts call."""
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        self._L.tidk_cuda_last_transfer(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, bool(c.value)

    def window_counts_resident(self, batch: SeqBatch, window: int, queries: Sequence[bytes],
                               out: Optional[Tuple[np.ndarray, np.ndarray]] = None):
        key = tuple(queries)
        cached = getattr(self, "_qcache", None)
        if cached is None or cached[0] != key:  # the ctypes argument arrays are rebuilt only when the queries change
            self._qcache = cached = (key, _qargs(queries))
        qa, ql, nq = cached[1]
        nwc = batch.__dict__.setdefault("_nwin", {})  # windows per batch, computed once per window size
        if window not in nwc:
            nwc[window] = int(n_windows(batch.offsets, window).sum()) if window > 0 else 0
        total = nwc[window] * nq
        if out is None:
            fwd = np.zeros(max(total, 1), dtype=np.uint64)
            rev = np.zeros(max(total, 1), dtype=np.uint64)
        else:
            fwd, rev = out
        ms, nl = C.c_float(0), C.c_uint32(0)
        self._check(self._L.tidk_cuda_window_counts_resident(self._h, batch._h, window, qa, ql, nq,
                                                             fwd.ctypes.data, rev.ctypes.data,
                                                             C.byref(ms), C.byref(nl)))
        self.last_kernel_ms, self.last_kernel_launches = ms.value, nl.value
        return fwd[:total], rev[:total]

    def window_counts_resident_async(self, batch: SeqBatch, window: int, queries: Sequence[bytes],
                                     out: Tuple[np.ndarray, np.ndarray]) -> None:
        """Enqueue one counting pass; `out` (pinned arrays from pinned_empty) is filled by the time wait() returns."""
        key = tuple(queries)
        cached = getattr(self, "_qcache", None)
        if cached is None or cached[0] != key:
            self._qcache = cached = (key, _qargs(queries))
        qa, ql, nq = cached[1]
        fwd, rev = out
        self._check(self._L.tidk_cuda_window_counts_resident_async(self._h, batch._h, window, qa, ql, nq,
                                                                   fwd.ctypes.data, rev.ctypes.data))

    def wait(self) -> Tuple[float, int]:
        """Collect every asynchronous call: (summed kernel ms, launches) since the previous wait."""
        ms, nl = C.c_float(0), C.c_uint32(0)
        self._check(self._L.tidk_cuda_wait(self._h, C.byref(ms), C.byref(nl)))
        self.last_kernel_ms, self.last_kernel_launches = ms.value, nl.value
        return ms.value, nl.value

    # -- explore ------------------------------------------------------------
    def _runs_out(self, runs, n) -> np.ndarray:
        if n.value:
            arr = np.empty(n.value, dtype=RUN_DTYPE)
            C.memmove(arr.ctypes.data, runs, n.value * C.sizeof(Run))  # one copy out of the library's buffer
            self._L.tidk_cuda_free_runs(runs)
            return arr
        return np.zeros(0, dtype=RUN_DTYPE)

    def explore_runs(self, seq: np.ndarray, offsets: np.ndarray, distance: float, kmin: int, kmax: int,
                     threshold: int) -> np.ndarray:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        runs, n = C.POINTER(Run)(), C.c_uint64(0)
        self._check(self._L.tidk_cuda_explore_runs(self._h, seq.ctypes.data, offsets.ctypes.data,
                                                   len(offsets) - 1, distance, kmin, kmax, threshold,
                                                   C.byref(runs), C.byref(n)))
        return self._runs_out(runs, n)

    def explore_timings(self):
        """(scan_ms, post_ms) of the last explore call (device time)."""
        out = (C.c_float * 2)()
        self._L.tidk_cuda_last_explore_timings(self._h, out, 2)
        return float(out[0]), float(out[1])

    def explore_host_timings(self):
        """(prep_ms, post_wall_ms): host wall clock before the scan / of the post-processing incl. the runs copy."""
        out = (C.c_float * 4)()
        self._L.tidk_cuda_last_explore_timings(self._h, out, 4)
        return float(out[2]), float(out[3])

    def explore_runs_resident(self, batch: SeqBatch, distance: float, kmin: int, kmax: int,
                              threshold: int) -> np.ndarray:
        runs, n = C.POINTER(Run)(), C.c_uint64(0)
        ms, nl = C.c_float(0), C.c_uint32(0)
        self._check(self._L.tidk_cuda_explore_runs_resident(self._h, batch._h, distance, kmin, kmax, threshold,
                                                            C.byref(runs), C.byref(n), C.byref(ms), C.byref(nl)))
        self.last_kernel_ms, self.last_kernel_launches = ms.value, nl.value
        return self._runs_out(runs, n)
