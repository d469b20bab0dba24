"""Shared helpers of the ksw tests: golden-fixture loader and result comparison."""
import os

import numpy as np

from pansvr_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ksw_golden.npz")


def golden_names():
    with np.load(GOLDEN) as z:
        return [str(n) for n in z["names"]]


def load_golden(name):
    """-> (KswBatch, res[n,11], cigar list) recorded from the reference's ksw2_extd2_sse.c."""
    with np.load(GOLDEN) as z:
        q, t = z[name + "/q"], z[name + "/t"]
        qlen, tlen = z[name + "/qlen"], z[name + "/tlen"]
        pv, mat = z[name + "/params"], z[name + "/mat"]
        res, cig = z[name + "/res"], z[name + "/cigar"]
    n = qlen.size
    qoff = np.zeros(n, np.int64); toff = np.zeros(n, np.int64)
    qoff[1:] = np.cumsum(qlen[:-1]); toff[1:] = np.cumsum(tlen[:-1])
    p = synth.KswParams(m=int(pv[0]), mat=mat, q=int(pv[1]), e=int(pv[2]), q2=int(pv[3]), e2=int(pv[4]), w=int(pv[5]),
                        zdrop=int(pv[6]), end_bonus=int(pv[7]), flag=int(pv[8]))
    b = synth.KswBatch(q, qoff, qlen, t, toff, tlen, p, name)
    coff = np.zeros(n + 1, np.int64); coff[1:] = np.cumsum(res[:, 9])
    cigs = [cig[coff[i]:coff[i + 1]] for i in range(n)]
    return b, res, cigs


def assert_same(res_a, cig_a, res_b, cig_b, what=""):
    """Bit-exact equality of the 11 ksw_extz_t fields and of the CIGAR words (cig_* are [n,cap] arrays)."""
    ra, rb = np.asarray(res_a)[:, :11], np.asarray(res_b)[:, :11]
    bad = np.nonzero((ra != rb).any(1))[0]
    assert bad.size == 0, f"{what}: {bad.size} tasks differ in ksw_extz_t, first {bad[0]}: {ra[bad[0]]} vs {rb[bad[0]]}"
    n_cig = ra[:, 9]
    cap = min(cig_a.shape[1], cig_b.shape[1])
    assert n_cig.max(initial=0) <= cap, f"{what}: cigar_cap {cap} too small for the comparison"
    mask = np.arange(cap)[None, :] < n_cig[:, None]
    diff = np.nonzero(((cig_a[:, :cap] != cig_b[:, :cap]) & mask).any(1))[0]
    assert diff.size == 0, f"{what}: {diff.size} tasks differ in CIGAR, first {diff[0]}"


def assert_matches_golden(res, cig, gres, gcigs, what=""):
    res = np.asarray(res)[:, :11]
    bad = np.nonzero((res != gres).any(1))[0]
    assert bad.size == 0, f"{what}: {bad.size} tasks differ from the reference, first {bad[0]}: {res[bad[0]]} vs {gres[bad[0]]}"
    for i, g in enumerate(gcigs):
        assert np.array_equal(cig[i, :g.size], g), f"{what}: CIGAR of task {i} differs from the reference"


def cigar_lengths(cig_row, n):
    """(query bases, target bases) consumed by a CIGAR."""
    ops = cig_row[:n] & 0xf
    lens = cig_row[:n] >> 4
    return int(lens[(ops == 0) | (ops == 1)].sum()), int(lens[(ops == 0) | (ops == 2)].sum())
