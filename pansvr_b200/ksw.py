"""Host-side mirror of the reference's ksw call boundary, over the C ABI of libpansvr_b200.so.

`KswContext.extd2_batch(batch)` is the batched form of the reference's only ksw call site on the
aln path (KSW_ALN_handler::align_non_splice -> ksw_extd2_sse, read_realignment.cpp:872-891);
`ksw_extd2_sse(...)` below is the one-call drop-in with the reference's argument order
(ksw2.h:63-64).  There is no CPU fallback: if the CUDA library is missing or no B200 is present
these raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .synth import KswBatch, KswParams

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpansvr_b200.so")
RES_WORDS = 12
RES_COLS = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end", "status")
EXPORTS = ("pansvr_ksw_create", "pansvr_ksw_destroy", "pansvr_last_error", "pansvr_host_alloc", "pansvr_host_free",
           "pansvr_ksw_extd2_batch", "pansvr_ksw_extd2_batch_device", "pansvr_ksw_last_stats", "pansvr_ksw_band_cells",
           "pansvr_ksw_extd2", "ksw_extd2_sse", "pansvr_int_alu_peak", "pansvr_int_pipe_peaks",
           "pansvr_aln_create", "pansvr_aln_create_multi", "pansvr_aln_destroy", "pansvr_aln_header_text", "pansvr_aln_last_error", "pansvr_aln_block",
           "pansvr_aln_block_bam", "pansvr_bam_open", "pansvr_bam_write", "pansvr_bam_close", "pansvr_aln_last_stats", "pansvr_aln_reset", "pansvr_free", "pansvr_fc_aln_main",
           "pansvr_aln_prime_read_stats", "pansvr_aln_await_state", "pansvr_aln_publish_state", "pansvr_aln_pieces")


class KswParamsC(C.Structure):
    _fields_ = [("m", C.c_int32), ("mat", C.c_void_p), ("gapo", C.c_int8), ("gape", C.c_int8), ("gapo2", C.c_int8),
                ("gape2", C.c_int8), ("w", C.c_int32), ("zdrop", C.c_int32), ("end_bonus", C.c_int32), ("flag", C.c_int32)]


class KswStatsC(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("tasks_fast_wrap", C.c_int64), ("tasks_fast_nowrap", C.c_int64),
                ("tasks_generic", C.c_int64), ("tasks_trivial", C.c_int64), ("kernel_ms", C.c_double),
                ("total_ms", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("tb_bytes_per_warp", C.c_int64), ("resident_warps", C.c_int64)]


class KswExtz(C.Structure):
    """ksw_extz_t (ksw2.h:26-35)."""
    _fields_ = [("max_zdropped", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int), ("mqe_t", C.c_int),
                ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int), ("m_cigar", C.c_int), ("n_cigar", C.c_int),
                ("reach_end", C.c_int), ("cigar", C.POINTER(C.c_uint32))]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m pansvr_b200.build` "
                           "(the ksw path has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.pansvr_last_error.restype = C.c_char_p
    lib.pansvr_ksw_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.pansvr_ksw_destroy.argtypes = [C.c_void_p]
    lib.pansvr_host_alloc.restype = C.c_void_p
    lib.pansvr_host_alloc.argtypes = [C.c_size_t]
    lib.pansvr_host_free.argtypes = [C.c_void_p]
    lib.pansvr_ksw_extd2_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(KswParamsC),
                                           C.c_void_p, C.c_void_p, C.c_int32]
    lib.pansvr_ksw_extd2_batch_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(KswParamsC),
                                                  C.c_void_p, C.c_void_p, C.c_int32]
    lib.pansvr_ksw_last_stats.argtypes = [C.c_void_p, C.POINTER(KswStatsC)]
    lib.pansvr_int_alu_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.pansvr_int_pipe_peaks.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.pansvr_ksw_band_cells.restype = C.c_int64
    lib.pansvr_ksw_band_cells.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    sse_args = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int8, C.c_void_p, C.c_int8, C.c_int8, C.c_int8,
                C.c_int8, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(KswExtz)]
    lib.pansvr_ksw_extd2.argtypes = sse_args
    lib.pansvr_ksw_extd2.restype = None
    lib.ksw_extd2_sse.argtypes = sse_args
    lib.ksw_extd2_sse.restype = None
    _lib = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib.pansvr_last_error().decode()}")


def _params_c(p: KswParams):
    mat = np.ascontiguousarray(p.mat, dtype=np.int8)
    pc = KswParamsC(p.m, mat.ctypes.data, p.q, p.e, p.q2, p.e2, p.w, p.zdrop, p.end_bonus, p.flag)
    return pc, mat


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (the batcher's pinned staging buffers)."""

    def __init__(self, shape, dtype):
        lib = load_library()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = lib.pansvr_host_alloc(max(nbytes, 1))
        if not self._ptr:
            raise RuntimeError("pansvr_host_alloc failed: " + lib.pansvr_last_error().decode())
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            load_library().pansvr_host_free(self._ptr)
            self._ptr = None


class KswContext:
    """One GPU + one stream + reusable scratch (pansvr_ksw_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib, self.lib.pansvr_ksw_create(device, C.byref(h)), "pansvr_ksw_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.pansvr_ksw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> dict:
        st = KswStatsC()
        _check(self.lib, self.lib.pansvr_ksw_last_stats(self.h, C.byref(st)), "pansvr_ksw_last_stats")
        return {k: getattr(st, k) for k, _ in KswStatsC._fields_}

    def int_alu_peak_gops(self) -> float:
        g = C.c_double()
        _check(self.lib, self.lib.pansvr_int_alu_peak(self.h, C.byref(g)), "pansvr_int_alu_peak")
        return g.value

    def int_pipe_peaks_gops(self) -> dict:
        """Integer throughput per pipe (Gop/s): the mixed yardstick, ALU pipe only, FMA pipe only, both pipes at once."""
        g = (C.c_double * 4)()
        _check(self.lib, self.lib.pansvr_int_pipe_peaks(self.h, g), "pansvr_int_pipe_peaks")
        return {"mixed": g[0], "alu_pipe": g[1], "fma_pipe": g[2], "both_pipes": g[3]}

    def extd2_batch(self, b: KswBatch, cigar_cap: int = 64, out=None):
        """Host buffers in, host buffers out: results[n,12] int32, cigar[n,cigar_cap] uint32."""
        n = b.n
        res, cig = out if out is not None else (np.zeros((n, RES_WORDS), np.int32), np.zeros((n, max(cigar_cap, 1)), np.uint32))
        pc, _mat = _params_c(b.params)
        qseq = np.ascontiguousarray(b.qseq, np.uint8); tseq = np.ascontiguousarray(b.tseq, np.uint8)
        qoff = np.ascontiguousarray(b.qoff, np.int64); toff = np.ascontiguousarray(b.toff, np.int64)
        qlen = np.ascontiguousarray(b.qlen, np.int32); tlen = np.ascontiguousarray(b.tlen, np.int32)
        rc = self.lib.pansvr_ksw_extd2_batch(self.h, n, qseq.ctypes.data, qseq.size, qoff.ctypes.data, qlen.ctypes.data,
                                             tseq.ctypes.data, tseq.size, toff.ctypes.data, tlen.ctypes.data, C.byref(pc),
                                             res.ctypes.data, cig.ctypes.data, cigar_cap)
        _check(self.lib, rc, "pansvr_ksw_extd2_batch")
        return res, cig

    def extd2_batch_device(self, n, d_qseq, d_qoff, d_qlen, d_tseq, d_toff, d_tlen, h_qlen, h_tlen, params: KswParams,
                           d_res, d_cigar, cigar_cap: int):
        """All arrays already in HBM (raw device pointers, e.g. torch .data_ptr()); results stay there."""
        pc, _mat = _params_c(params)
        h_qlen = np.ascontiguousarray(h_qlen, np.int32); h_tlen = np.ascontiguousarray(h_tlen, np.int32)
        rc = self.lib.pansvr_ksw_extd2_batch_device(self.h, n, d_qseq, d_qoff, d_qlen, d_tseq, d_toff, d_tlen,
                                                    h_qlen.ctypes.data, h_tlen.ctypes.data, C.byref(pc), d_res, d_cigar, cigar_cap)
        _check(self.lib, rc, "pansvr_ksw_extd2_batch_device")


def band_cells(qlen: int, tlen: int, w: int) -> int:
    return int(load_library().pansvr_ksw_band_cells(qlen, tlen, w))


def ksw_extd2_sse(query: np.ndarray, target: np.ndarray, p: KswParams, ez: KswExtz | None = None, symbol: str = "ksw_extd2_sse"):
    """One call through the drop-in symbol (same argument order as ksw2.h:63-64); returns the ksw_extz_t."""
    lib = load_library()
    ez = ez if ez is not None else KswExtz()
    q = np.ascontiguousarray(query, np.uint8); t = np.ascontiguousarray(target, np.uint8)
    mat = np.ascontiguousarray(p.mat, np.int8)
    getattr(lib, symbol)(None, q.size, q.ctypes.data, t.size, t.ctypes.data, p.m, mat.ctypes.data, p.q, p.e, p.q2, p.e2,
                         p.w, p.zdrop, p.end_bonus, p.flag, C.byref(ez))
    return ez
