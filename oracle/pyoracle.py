"""ctypes front end of oracle/libksw_oracle.so -- TEST / BASELINE INFRASTRUCTURE ONLY.

`run(batch, impl="oracle")` executes every task of a KswBatch on host threads through either
  * impl="oracle": the scalar C restatement oracle/ksw_extd2_oracle.c, or
  * impl="ref":    the reference's own ksw2_extd2_sse.c built unmodified into oracle/_ref/libksw_ref.so
and returns (results[n,12] int32, cigar[n,cap] uint32, seconds).

Result columns: max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar reach_end overflow
(the fields of ksw_extz_t, /root/reference/src/kswlib/ksw2.h:26-35).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libksw_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libksw_ref.so")
RES_COLS = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end",
            "overflow")
_lib = None


def build(force: bool = False) -> None:
    """Compile the checker(s); a no-op when the objects are newer than their sources."""
    src = [os.path.join(HERE, f) for f in ("ksw_extd2_oracle.c", "batch_driver.c")]
    stale = force or not os.path.exists(ORACLE_SO) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src)
    need_ref = os.path.exists("/root/reference/src/kswlib/ksw2_extd2_sse.c") and (force or not os.path.exists(REF_SO))
    if stale:
        subprocess.check_call(["make", "-s", "-C", HERE, os.path.join(HERE, "libksw_oracle.so")])
    if need_ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.ksw_batch_run.restype = C.c_double
        _lib.ksw_batch_run.argtypes = [C.c_char_p, C.c_char_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_int]
        _lib.ksw_extd2_oracle_cells.restype = C.c_int64
        _lib.ksw_extd2_oracle_cells.argtypes = [C.c_int, C.c_int, C.c_int]
    return _lib


def cells(qlen: int, tlen: int, w: int) -> int:
    return int(lib().ksw_extd2_oracle_cells(qlen, tlen, w))


def run(batch, impl: str = "oracle", threads: int = 1, cigar_cap: int = 64):
    L = lib()
    p = batch.params
    n = batch.n
    res = np.zeros((n, 12), dtype=np.int32)
    cig = np.zeros((n, cigar_cap), dtype=np.uint32)
    qseq = np.ascontiguousarray(batch.qseq, dtype=np.uint8)
    tseq = np.ascontiguousarray(batch.tseq, dtype=np.uint8)
    qoff = np.ascontiguousarray(batch.qoff, dtype=np.int64)
    toff = np.ascontiguousarray(batch.toff, dtype=np.int64)
    qlen = np.ascontiguousarray(batch.qlen, dtype=np.int32)
    tlen = np.ascontiguousarray(batch.tlen, dtype=np.int32)
    mat = np.ascontiguousarray(p.mat, dtype=np.int8)
    if impl == "oracle":
        path, sym = b"", b""
    elif impl == "ref":
        if not have_ref():
            raise FileNotFoundError(REF_SO)
        path, sym = REF_SO.encode(), b"ksw_extd2_sse"
    else:
        raise ValueError(impl)
    secs = L.ksw_batch_run(path, sym, n, qseq.ctypes.data, qoff.ctypes.data, qlen.ctypes.data,
                           tseq.ctypes.data, toff.ctypes.data, tlen.ctypes.data,
                           p.m, mat.ctypes.data, p.q, p.e, p.q2, p.e2, p.w, p.zdrop, p.end_bonus, p.flag,
                           threads, res.ctypes.data, cig.ctypes.data, cigar_cap)
    if secs < 0:
        raise RuntimeError(f"ksw_batch_run failed ({secs})")
    return res, cig, secs
