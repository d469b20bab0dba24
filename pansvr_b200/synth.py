"""Seeded synthetic ksw workloads (BASELINE.json configs) and fuzz cases.

A workload is a :class:`KswBatch`: two byte pools (one base per byte, 0..3 = ACGT, 4 = N, the
encoding the reference hands to ``ksw_extd2_sse`` -- read_realignment.cpp:646-654, deBGA_index.cpp:307)
plus per-task (offset, length) windows into them, and one set of scoring parameters for the whole
batch (the reference uses one parameter set per run: read_realignment.cpp:817-827,889).
Target windows may overlap: in the pipeline they are slices of the same anchor reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# flags of ksw2.h:9-15
KSW_EZ_SCORE_ONLY = 0x01
KSW_EZ_RIGHT = 0x02
KSW_EZ_GENERIC_SC = 0x04
KSW_EZ_APPROX_MAX = 0x08
KSW_EZ_APPROX_DROP = 0x10
KSW_EZ_EXTZ_ONLY = 0x40
KSW_EZ_REV_CIGAR = 0x80
KSW_NEG_INF = -0x40000000


def dna_matrix(match: int = 2, mismatch: int = 12, sc_ambi: int = 0) -> np.ndarray:
    """5x5 matrix as KSW_ALN_handler::ksw_gen_mat_D builds it (read_realignment.cpp:829-844)."""
    mat = np.zeros(25, dtype=np.int8)
    for i in range(4):
        for j in range(4):
            mat[i * 5 + j] = match if i == j else -mismatch
        mat[i * 5 + 4] = sc_ambi
    mat[20:25] = sc_ambi
    return mat


@dataclass
class KswParams:
    """Scoring/band parameters of one batch; defaults are fc_aln's (read_realignment.hpp:33-41, .cpp:817-827)."""
    m: int = 5
    mat: np.ndarray = field(default_factory=dna_matrix)
    q: int = 16
    e: int = 1
    q2: int = 32
    e2: int = 0
    w: int = 200
    zdrop: int = 400
    end_bonus: int = -1
    flag: int = 0


@dataclass
class KswBatch:
    qseq: np.ndarray  # uint8 pool
    qoff: np.ndarray  # int64 [n]
    qlen: np.ndarray  # int32 [n]
    tseq: np.ndarray  # uint8 pool
    toff: np.ndarray  # int64 [n]
    tlen: np.ndarray  # int32 [n]
    params: KswParams
    name: str = ""

    @property
    def n(self) -> int:
        return int(self.qlen.shape[0])

    def take(self, idx) -> "KswBatch":
        idx = np.asarray(idx)
        return KswBatch(self.qseq, self.qoff[idx].copy(), self.qlen[idx].copy(), self.tseq,
                        self.toff[idx].copy(), self.tlen[idx].copy(), self.params, self.name)

    def head(self, n: int) -> "KswBatch":
        return self.take(np.arange(min(n, self.n)))


def band_cells(qlen: int, tlen: int, w: int) -> int:
    """In-band DP cells the reference visits before 16-lane rounding (KSW:131-138)."""
    if qlen <= 0 or tlen <= 0:
        return 0
    if w < 0:
        w = max(qlen, tlen)
    r = np.arange(qlen + tlen - 1, dtype=np.int64)
    lo = np.maximum(np.maximum(0, r - qlen + 1), (r - w + 1) >> 1)
    hi = np.minimum(np.minimum(tlen - 1, r), (r + w) >> 1)
    bad = np.nonzero(lo > hi)[0]
    if bad.size:
        lo, hi = lo[: bad[0]], hi[: bad[0]]
    return int((hi - lo + 1).sum())


def batch_cells(b: KswBatch) -> int:
    """Sum of band_cells over a batch, memoised over distinct shapes."""
    shapes, counts = np.unique(np.stack([b.qlen, b.tlen], 1), axis=0, return_counts=True)
    return int(sum(band_cells(int(ql), int(tl), b.params.w) * int(c) for (ql, tl), c in zip(shapes, counts)))


def _mutated_rows(rng: np.random.Generator, pool: np.ndarray, toff: np.ndarray, out_len: int,
                  sub: float, ins: float, dele: float) -> np.ndarray:
    """Row i = the first `out_len` bases of an error-injected copy of pool[toff[i]:...]."""
    n = toff.shape[0]
    steps = out_len + max(16, int(out_len * (ins + dele) * 8) + 16)
    u = rng.random((n, steps), dtype=np.float32)
    is_ins = u < ins
    is_del = (u >= ins) & (u < ins + dele)
    consume = ~is_ins                                  # step reads one target base
    emit = ~is_del                                     # step writes one query base
    tpos = np.cumsum(consume, axis=1, dtype=np.int32) - consume
    base = pool[(toff[:, None] + tpos).astype(np.int64)]
    do_sub = (rng.random((n, steps), dtype=np.float32) < sub) & ~is_ins & ~is_del
    shift = rng.integers(1, 4, size=(n, steps), dtype=np.uint8)
    base = np.where(do_sub, (base + shift) & 3, base).astype(np.uint8)
    rnd = rng.integers(0, 4, size=(n, steps), dtype=np.uint8)
    base = np.where(is_ins, rnd, base)
    order = np.argsort(~emit, axis=1, kind="stable")[:, :out_len]
    return np.take_along_axis(base, order, axis=1)


def config2_batch(n: int, seed: int = 11, pool_bases: int = 1 << 24, qlen: int = 150, tlen: int = 1100,
                  w: int = 100, chunk: int = 1 << 16) -> KswBatch:
    """BASELINE.json configs[1]: n x 150 bp reads vs 1.1 kb anchor windows, band 100 (SURVEY.md 8d "Config 2").

    qlen = the first 150 target bases with 0.8 % substitutions, 0.15 % insertions, 0.15 % deletions;
    zdrop 400, end_bonus -1, flag 0, 2/-12, gaps min(16+1k, 32+0k).  Every task is band-clipped and
    ends with zdropped=1 after 399 anti-diagonals (25 100 in-band cells).
    """
    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 4, size=pool_bases + tlen + 64, dtype=np.uint8)
    toff = rng.integers(0, pool_bases, size=n, dtype=np.int64)
    q = np.empty((n, qlen), dtype=np.uint8)
    for s in range(0, n, chunk):
        q[s:s + chunk] = _mutated_rows(rng, pool, toff[s:s + chunk], qlen, 0.008, 0.0015, 0.0015)
    return KswBatch(q.reshape(-1), np.arange(n, dtype=np.int64) * qlen, np.full(n, qlen, np.int32),
                    pool, toff, np.full(n, tlen, np.int32), KswParams(w=w), name=f"config2_{qlen}x{tlen}_w{w}")


def config4_batch(n: int, kind: str = "ext", seed: int = 13, pool_bases: int = 1 << 24, w: int = 500) -> KswBatch:
    """BASELINE.json configs[3]: 250 bp reads, wide band (SURVEY.md 8d "Config 4").

    kind "ext": qlen 250 vs tlen 280 (70 000 cells, never clipped); "window": tlen 1500 (156 375 cells);
    "global": qlen in [200,250] end-to-end against a target that differs by 1-3 indels of 1-40 bp.
    """
    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 4, size=pool_bases + 2048, dtype=np.uint8)
    toff = rng.integers(0, pool_bases, size=n, dtype=np.int64)
    if kind in ("ext", "window"):
        qlen, tlen = 250, (280 if kind == "ext" else 1500)
        q = np.empty((n, qlen), dtype=np.uint8)
        for s in range(0, n, 1 << 15):
            q[s:s + (1 << 15)] = _mutated_rows(rng, pool, toff[s:s + (1 << 15)], qlen, 0.008, 0.0015, 0.0015)
        return KswBatch(q.reshape(-1), np.arange(n, dtype=np.int64) * qlen, np.full(n, qlen, np.int32), pool, toff,
                        np.full(n, tlen, np.int32), KswParams(w=w), name=f"config4_{kind}_w{w}")
    if kind != "global":
        raise ValueError(kind)
    qs, qlens, tlens = [], np.empty(n, np.int32), np.empty(n, np.int32)
    for i in range(n):
        tl = int(rng.integers(200, 251))
        t = pool[toff[i]:toff[i] + tl]
        qv = t.copy()
        for _ in range(int(rng.integers(1, 4))):
            L = int(rng.integers(1, 41))
            p = int(rng.integers(10, max(11, qv.size - 10)))
            if rng.random() < 0.5 and qv.size - L > 60:
                qv = np.concatenate([qv[:p], qv[p + L:]])
            else:
                qv = np.concatenate([qv[:p], rng.integers(0, 4, L, dtype=np.uint8), qv[p:]])
        sub = rng.random(qv.size) < 0.008
        qv = np.where(sub, (qv + rng.integers(1, 4, qv.size, dtype=np.uint8)) & 3, qv).astype(np.uint8)
        qs.append(qv)
        qlens[i], tlens[i] = qv.size, tl
    qoff = np.zeros(n, np.int64)
    qoff[1:] = np.cumsum(qlens[:-1])
    return KswBatch(np.concatenate(qs), qoff, qlens, pool, toff, tlens, KswParams(w=w), name=f"config4_global_w{w}")


def fcsv_batch(n: int, seed: int = 19, pool_bases: int = 1 << 23) -> KswBatch:
    """the `fc_sv` contig alignments (SignalAssembly.hpp:411-421,459-464; SignalAssembly.cpp:822-831).

    An assembled contig of 200-1500 bp against the anchor window it was assembled over, the window reaching 60 bp past
    the contig's expected end; scoring 2/-10, gaps min(24+2k, 32+1k), w = zdrop = 132, flag 0.  Half of the contigs
    carry one SV-sized indel (30-300 bp) relative to the window, all carry ~0.5 % substitutions."""
    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 4, size=pool_bases + 4096, dtype=np.uint8)
    toff = rng.integers(0, pool_bases, size=n, dtype=np.int64)
    qs, qlens, tlens = [], np.empty(n, np.int32), np.empty(n, np.int32)
    for i in range(n):
        L = int(rng.integers(200, 1501))
        t = pool[toff[i]:toff[i] + L]
        qv = t.copy()
        if rng.random() < 0.5:
            sv = int(rng.integers(30, 301))
            p = int(rng.integers(60, max(61, qv.size - 60)))
            if rng.random() < 0.5 and qv.size - sv > 120:
                qv = np.concatenate([qv[:p], qv[p + sv:]])
            else:
                qv = np.concatenate([qv[:p], rng.integers(0, 4, sv, dtype=np.uint8), qv[p:]])
        sub = rng.random(qv.size) < 0.005
        qv = np.where(sub, (qv + rng.integers(1, 4, qv.size, dtype=np.uint8)) & 3, qv).astype(np.uint8)
        qs.append(qv)
        qlens[i], tlens[i] = qv.size, L + 60
    qoff = np.zeros(n, np.int64)
    qoff[1:] = np.cumsum(qlens[:-1])
    return KswBatch(np.concatenate(qs), qoff, qlens, pool, toff, tlens,
                    KswParams(mat=dna_matrix(2, 10), q=24, e=2, q2=32, e2=1, w=132, zdrop=132), name="fc_sv_contigs")


def pipeline_like_batch(n: int, seed: int = 17, pool_bases: int = 1 << 22) -> KswBatch:
    """Task shapes fc_aln really emits: extensions with qlen p50 26 / p90 90 / max 121 and
    tlen = qlen + 30, plus short end-to-end gaps (qlen, tlen <= 50), w=200, zdrop=400, flag=0."""
    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 4, size=pool_bases + 512, dtype=np.uint8)
    toff = rng.integers(0, pool_bases, size=n, dtype=np.int64)
    qs, qlens, tlens = [], np.empty(n, np.int32), np.empty(n, np.int32)
    for i in range(n):
        if rng.random() < 0.8:
            ql = int(min(121, max(8, rng.gamma(2.0, 20.0))))
            tl = ql + 30
            qv = _mutated_rows(rng, pool, toff[i:i + 1], ql, 0.03, 0.01, 0.01)[0]
        else:
            tl = int(rng.integers(1, 51))
            ql = int(max(1, tl + rng.integers(-10, 11)))
            qv = _mutated_rows(rng, pool, toff[i:i + 1], ql, 0.05, 0.03, 0.03)[0]
        qs.append(qv)
        qlens[i], tlens[i] = ql, tl
    qoff = np.zeros(n, np.int64)
    qoff[1:] = np.cumsum(qlens[:-1])
    return KswBatch(np.concatenate(qs), qoff, qlens, pool, toff, tlens, KswParams(), name="pipeline_like")


def fuzz_batch(n: int, seed: int, max_len: int = 260, params: KswParams | None = None, n_frac: float = 0.02,
               related: float = 0.8) -> KswBatch:
    """Ragged random tasks: lengths 1..max_len, related or unrelated pairs, occasional N, big indels."""
    rng = np.random.default_rng(seed)
    params = params or KswParams()
    qs, ts = [], []
    for _ in range(n):
        tl = int(rng.integers(1, max_len + 1))
        t = rng.integers(0, 4, tl, dtype=np.uint8)
        if rng.random() < related:
            qv = t.copy()
            for _ in range(int(rng.integers(0, 4))):
                L = int(rng.integers(1, 1 + max(1, min(120, qv.size // 2))))
                p = int(rng.integers(0, qv.size + 1))
                if rng.random() < 0.5:
                    qv = np.concatenate([qv[:p], qv[p + L:]])
                else:
                    qv = np.concatenate([qv[:p], rng.integers(0, 4, L, dtype=np.uint8), qv[p:]])
            if qv.size == 0:
                qv = rng.integers(0, 4, 1, dtype=np.uint8)
            sub = rng.random(qv.size) < rng.choice([0.0, 0.01, 0.05, 0.3])
            qv = np.where(sub, (qv + rng.integers(1, 4, qv.size, dtype=np.uint8)) & 3, qv).astype(np.uint8)
            if rng.random() < 0.3:  # extension-like: truncate the query
                qv = qv[: max(1, int(rng.integers(1, qv.size + 1)))]
        else:
            qv = rng.integers(0, 4, int(rng.integers(1, max_len + 1)), dtype=np.uint8)
        if n_frac > 0 and rng.random() < 0.2:
            qv = np.where(rng.random(qv.size) < n_frac, 4, qv).astype(np.uint8)
            t = np.where(rng.random(t.size) < n_frac, 4, t).astype(np.uint8)
        qs.append(qv)
        ts.append(t)
    qlen = np.array([x.size for x in qs], np.int32)
    tlen = np.array([x.size for x in ts], np.int32)
    qoff = np.zeros(n, np.int64)
    toff = np.zeros(n, np.int64)
    qoff[1:] = np.cumsum(qlen[:-1])
    toff[1:] = np.cumsum(tlen[:-1])
    return KswBatch(np.concatenate(qs), qoff, qlen, np.concatenate(ts), toff, tlen, params, name=f"fuzz{seed}")
