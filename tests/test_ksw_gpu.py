"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and against the
fixtures recorded from the reference's own ksw2_extd2_sse.c.  Bit-exact (integer work)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle
from pansvr_b200 import ksw, synth
from tests.kswtest_util import assert_matches_golden, assert_same, cigar_lengths, golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_reference_fixture(ksw_ctx, name):
    b, gres, gcigs = load_golden(name)
    res, cig = ksw_ctx.extd2_batch(b, cigar_cap=96)
    assert not (res[:, 11] & 1).any()
    assert_matches_golden(res, cig, gres, gcigs, name)


FUZZ = [(41, 200, 400, 0, 260), (42, 100, 400, 0, 260), (43, 50, 100, 0, 260), (44, 20, 50, 0, 260), (45, 8, 30, 0, 200),
        (46, -1, -1, 0, 200), (47, 100, 400, 0x40, 260), (48, 30, 100, 0x40, 260), (49, 64, 200, 0x80, 200),
        (50, 100, 400, 0x01, 260), (51, 3, 400, 0, 200), (52, 1, 400, 0, 100), (53, 0, 400, 0, 100), (54, 16, 400, 0, 300),
        (55, 500, 400, 0, 520), (56, 64, 200, 0x02, 200), (57, 33, 400, 0x04, 200), (58, 100, 100, 0x18, 200),
        (59, 100, 100, 0x08, 200), (60, 200, 400, 0xC0, 260)]


@pytest.mark.parametrize("seed,w,zdrop,flag,max_len", FUZZ)
def test_fuzz_against_oracle(ksw_ctx, seed, w, zdrop, flag, max_len):
    b = synth.fuzz_batch(3000, seed, max_len=max_len, params=synth.KswParams(w=w, zdrop=zdrop, flag=flag))
    res, cig = ksw_ctx.extd2_batch(b, cigar_cap=128)
    r0, c0, _ = pyoracle.run(b, "oracle", threads=8, cigar_cap=128)
    assert_same(r0, c0, res, cig, f"fuzz seed {seed}")


def test_other_scoring_sets(ksw_ctx):
    sets = [synth.KswParams(mat=synth.dna_matrix(2, 10), q=24, e=2, q2=32, e2=1, w=132, zdrop=132),      # fc_sv (SignalAssembly.hpp:411-421)
            synth.KswParams(mat=synth.dna_matrix(1, 4, sc_ambi=-1), q=24, e=1, q2=6, e2=2, w=60, zdrop=80),  # swapped pieces
            synth.KswParams(mat=synth.dna_matrix(4, 24), q=60, e=8, q2=100, e2=20, w=50, zdrop=300),        # gap costs near the int8 edge
            synth.KswParams(mat=synth.dna_matrix(1, 1), q=1, e=1, q2=2, e2=1, w=40, zdrop=20),
            # first piece with the smaller extension (e < e2 after the swap): inconsistent boundary, wraps inside unclipped bands
            synth.KswParams(mat=synth.dna_matrix(2, 11), q=22, e=3, q2=14, e2=0, w=200, zdrop=384),
            synth.KswParams(q=16, e=0, q2=32, e2=1, w=200, zdrop=400),
            synth.KswParams(mat=synth.dna_matrix(2, 11), q=14, e=1, q2=22, e2=3, w=300, zdrop=400)]
    for k, p in enumerate(sets):
        b = synth.fuzz_batch(2000, 70 + k, max_len=300, params=p)
        res, cig = ksw_ctx.extd2_batch(b, cigar_cap=160)
        r0, c0, _ = pyoracle.run(b, "oracle", threads=8, cigar_cap=160)
        assert_same(r0, c0, res, cig, f"scoring set {k}")


def test_pipeline_like_and_config4(ksw_ctx):
    for b in (synth.pipeline_like_batch(6000), synth.config4_batch(1500, "ext"), synth.config4_batch(600, "window"),
              synth.config4_batch(1500, "global")):
        res, cig = ksw_ctx.extd2_batch(b, cigar_cap=64)
        r0, c0, _ = pyoracle.run(b, "oracle", threads=8, cigar_cap=64)
        assert_same(r0, c0, res, cig, b.name)
    st = ksw_ctx.stats()
    assert st["kernel_launches"] >= 1


def test_fc_sv_contig_tasks(ksw_ctx):
    """contig-vs-anchor-window alignments of fc_sv (clipped band, thousands of anti-diagonals)."""
    b = synth.fcsv_batch(400)
    res, cig = ksw_ctx.extd2_batch(b, cigar_cap=96)
    r0, c0, _ = pyoracle.run(b, "oracle", threads=8, cigar_cap=96)
    assert_same(r0, c0, res, cig, b.name)


def test_empty_ragged_and_trivial(ksw_ctx):
    p = synth.KswParams()
    res, cig = ksw_ctx.extd2_batch(synth.KswBatch(np.zeros(0, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int32),
                                                  np.zeros(0, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int32), p))
    assert res.shape == (0, 12)
    seq = np.array([0, 1, 2, 3, 4, 0, 1], np.uint8)
    b = synth.KswBatch(seq, np.array([0, 0, 0, 2], np.int64), np.array([0, 1, 7, 5], np.int32), seq,
                       np.array([0, 0, 0, 0], np.int64), np.array([3, 0, 7, 1], np.int32), p)
    res, cig = ksw_ctx.extd2_batch(b)
    r0, c0, _ = pyoracle.run(b, "oracle")
    assert_same(r0, c0, res, cig, "ragged")
    # mismatch score too large for the gap costs: the reference returns right after the reset (KSW:93)
    b2 = synth.fuzz_batch(50, 5, params=synth.KswParams(mat=synth.dna_matrix(2, 60), q=4, e=1, q2=5, e2=1))
    res, cig = ksw_ctx.extd2_batch(b2)
    r0, c0, _ = pyoracle.run(b2, "oracle")
    assert_same(r0, c0, res, cig, "trivial")
    assert (res[:, 8] == synth.KSW_NEG_INF).all() and (res[:, 9] == 0).all()


def test_cigar_capacity_overflow_is_reported(ksw_ctx):
    b = synth.fuzz_batch(300, 90, params=synth.KswParams(w=100, zdrop=400), related=0.3)
    r0, c0, _ = pyoracle.run(b, "oracle", cigar_cap=128)
    res, cig = ksw_ctx.extd2_batch(b, cigar_cap=4)
    assert np.array_equal(res[:, :11], r0[:, :11])          # n_cigar stays exact
    assert np.array_equal((res[:, 11] & 1) == 1, r0[:, 9] > 4)
    ok = r0[:, 9] <= 4
    assert_same(r0[ok], c0[ok][:, :4], res[ok], cig[ok], "within cap")


def test_drop_in_symbol_matches_reference_semantics(ksw_ctx):
    """ksw_extd2_sse-compatible entry: same ez contents, ez reused across calls, m_cigar grows by doubling."""
    b = synth.fuzz_batch(40, 91, params=synth.KswParams(w=100, zdrop=400))
    r0, c0, _ = pyoracle.run(b, "oracle", cigar_cap=128)
    ez = ksw.KswExtz()
    for sym in ("ksw_extd2_sse", "pansvr_ksw_extd2"):
        for i in range(b.n):
            q = b.qseq[b.qoff[i]:b.qoff[i] + b.qlen[i]]
            t = b.tseq[b.toff[i]:b.toff[i] + b.tlen[i]]
            ksw.ksw_extd2_sse(q, t, b.params, ez, symbol=sym)
            got = [ez.max_zdropped & 0x7fffffff, ez.max_zdropped >> 31, ez.max_q, ez.max_t, ez.mqe, ez.mqe_t, ez.mte, ez.mte_q,
                   ez.score, ez.n_cigar, ez.reach_end]
            assert got == r0[i, :11].tolist(), (sym, i)
            assert [ez.cigar[k] for k in range(ez.n_cigar)] == c0[i, :ez.n_cigar].tolist()
            assert ez.m_cigar >= ez.n_cigar and (ez.m_cigar == 0 or (ez.m_cigar & (ez.m_cigar - 1)) == 0)
    C.CDLL(None).free(ez.cigar)


def test_device_resident_entry_matches_host_entry(ksw_ctx):
    import torch
    b = synth.config2_batch(4096, pool_bases=1 << 18)
    res_h, cig_h = ksw_ctx.extd2_batch(b, cigar_cap=16)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(np.ascontiguousarray(getattr(b, k))).to(dev) for k in ("qseq", "qoff", "qlen", "tseq", "toff", "tlen")}
    d_res = torch.zeros((b.n, 12), dtype=torch.int32, device=dev)
    d_cig = torch.zeros((b.n, 16), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ksw_ctx.extd2_batch_device(b.n, d["qseq"].data_ptr(), d["qoff"].data_ptr(), d["qlen"].data_ptr(), d["tseq"].data_ptr(),
                               d["toff"].data_ptr(), d["tlen"].data_ptr(), b.qlen, b.tlen, b.params, d_res.data_ptr(),
                               d_cig.data_ptr(), 16)
    assert np.array_equal(d_res.cpu().numpy(), res_h)
    assert np.array_equal(d_cig.cpu().numpy().view(np.uint32), cig_h)


def test_config2_full_size_properties(ksw_ctx):
    """BASELINE configs[1] at its full size (1 M tasks): size-independent properties + an oracle-checked sample."""
    n = 1_000_000
    b = synth.config2_batch(n)
    res, cig = ksw_ctx.extd2_batch(b, cigar_cap=16)
    assert (res[:, 1] == 1).all() and (res[:, 8] == synth.KSW_NEG_INF).all()      # band closes: zdropped, no global score
    assert (res[:, 0] > 0).all() and (res[:, 4] > 0).all() and not (res[:, 11] & 1).any()
    # the CIGAR is the path to the max cell: it consumes max_q+1 query and max_t+1 target bases
    idx = np.random.default_rng(5).choice(n, 3000, replace=False)
    for i in idx[:600]:
        ql, tl = cigar_lengths(cig[i], res[i, 9])
        assert (ql, tl) == (res[i, 2] + 1, res[i, 3] + 1)
    sub = b.take(idx)
    r0, c0, _ = pyoracle.run(sub, "oracle", threads=8, cigar_cap=16)
    assert_same(r0, c0, res[idx], cig[idx], "config2 sample")
    # determinism: a second pass gives the same bytes
    res2, cig2 = ksw_ctx.extd2_batch(b, cigar_cap=16)
    assert np.array_equal(res, res2) and np.array_equal(cig, cig2)


def test_long_query_narrow_band_goes_generic(ksw_ctx):
    """a long query against a tiny target / narrow band picks a narrow team, whose shared memory (one query copy per
    alignment of the warp) would exceed the per-CTA limit; the planner must route such tasks to the generic kernel, not fail."""
    rng = np.random.default_rng(91)
    n = 24
    qlen = rng.integers(4000, 8001, n).astype(np.int32)
    tlen = rng.integers(1, 17, n).astype(np.int32)
    qoff = np.concatenate([[0], np.cumsum(qlen[:-1])]).astype(np.int64)
    toff = np.concatenate([[0], np.cumsum(tlen[:-1])]).astype(np.int64)
    qseq = rng.integers(0, 4, int(qlen.sum()), dtype=np.uint8)
    tseq = rng.integers(0, 4, int(tlen.sum()), dtype=np.uint8)
    for w in (3, 15, 47):
        b = synth.KswBatch(qseq, qoff, qlen, tseq, toff, tlen, synth.KswParams(w=w, zdrop=400, flag=0), f"long query w={w}")
        res, cig = ksw_ctx.extd2_batch(b, cigar_cap=64)
        r0, c0, _ = pyoracle.run(b, "oracle", threads=4, cigar_cap=64)
        assert_same(r0, c0, res, cig, b.name)


def test_contexts_in_flight_with_different_query_lengths():
    """Several contexts launch the same kernel variants at the same time with different shared-memory sizes (the sub-blocks of
    fc_aln in flight, each with its own longest query).  The limit on a kernel's dynamic shared memory is a property of the
    function, so a context must not lower it under another one's launch (round 2: intermittent "invalid argument")."""
    import threading
    from pansvr_b200 import ksw
    p = synth.KswParams(w=30)                                              # every task wider than the band: one TEAM variant for all lengths
    batches = [synth.fuzz_batch(1500, 77, max_len=100, params=p), synth.fuzz_batch(1500, 78, max_len=700, params=p)]
    want = [pyoracle.run(b, "oracle", threads=4, cigar_cap=64) for b in batches]
    errors = []

    def work(k):
        try:
            ctx = ksw.KswContext(0)
            for _ in range(40):
                res, cig = ctx.extd2_batch(batches[k], cigar_cap=64)
                assert_same(want[k][0], want[k][1], res, cig, f"context {k}")
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in (0, 1, 0, 1)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors[:2]
