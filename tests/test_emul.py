"""CPU test of the warp program itself: pansvr_b200/csrc/ksw_team.cuh compiled for the host and
stepped on the 32-fibre lock-step warp of tests/emul (no GPU needed), against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle
from pansvr_b200 import synth
from tests.kswtest_util import assert_same

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emul")])
    lib = C.CDLL(os.path.join(HERE, "emul", "libksw_emul.so"))
    lib.emul_ksw_team_batch.restype = C.c_int

    def run(b, cap=96, force_team=0, force_wrap=0):
        p = b.params
        res = np.zeros((b.n, 12), np.int32); cig = np.zeros((b.n, cap), np.uint32)
        mat = np.ascontiguousarray(p.mat, np.int8)
        arrs = [np.ascontiguousarray(x) for x in (b.qseq, b.qoff, b.qlen, b.tseq, b.toff, b.tlen)]
        vp = lambda a: C.c_void_p(a.ctypes.data)
        rc = lib.emul_ksw_team_batch(b.n, vp(arrs[0]), vp(arrs[1]), vp(arrs[2]), vp(arrs[3]), vp(arrs[4]), vp(arrs[5]), p.m,
                                     vp(mat), p.q, p.e, p.q2, p.e2, p.w, p.zdrop, p.end_bonus, p.flag, vp(res), vp(cig), cap,
                                     force_team, force_wrap)
        assert rc == 0
        return res, cig
    return run


CASES = [
    ("config2", lambda: synth.config2_batch(6, pool_bases=1 << 16), 0, 0),
    ("pipeline", lambda: synth.pipeline_like_batch(60, pool_bases=1 << 16), 0, 0),
    ("pipeline_forced_wrap", lambda: synth.pipeline_like_batch(40, seed=5, pool_bases=1 << 16), 0, 1),
    ("clipped_w50", lambda: synth.fuzz_batch(40, 31, params=synth.KswParams(w=50, zdrop=100)), 0, 0),
    ("clipped_w8", lambda: synth.fuzz_batch(40, 32, params=synth.KswParams(w=8, zdrop=30)), 0, 0),
    ("extz", lambda: synth.fuzz_batch(40, 33, params=synth.KswParams(w=100, zdrop=400, flag=0x40)), 0, 0),
    ("rev_cigar_team32", lambda: synth.fuzz_batch(30, 34, max_len=80, params=synth.KswParams(w=64, zdrop=200, flag=0x80)), 32, 0),
    ("zdrop_tight", lambda: synth.fuzz_batch(60, 36, params=synth.KswParams(w=30, zdrop=20)), 0, 0),
    ("tiny_bands", lambda: synth.fuzz_batch(60, 37, max_len=120, params=synth.KswParams(w=2, zdrop=400)), 0, 0),
    ("fc_sv_contigs", lambda: synth.fcsv_batch(3, pool_bases=1 << 16), 0, 0),
    ("swapped_e_lt_e2", lambda: synth.fuzz_batch(60, 77, max_len=200, params=synth.KswParams(mat=synth.dna_matrix(2, 11), q=22, e=3, q2=14, e2=0, w=200, zdrop=384)), 0, 0),
    ("e_lt_e2", lambda: synth.fuzz_batch(60, 78, max_len=200, params=synth.KswParams(q=16, e=0, q2=32, e2=1, w=200, zdrop=400)), 0, 0),
    ("wide_w500", lambda: synth.fuzz_batch(6, 35, max_len=460, params=synth.KswParams(w=500)), 0, 0),
]


@pytest.mark.parametrize("name,make,force_team,force_wrap", CASES, ids=[c[0] for c in CASES])
def test_warp_program_matches_oracle(emul, name, make, force_team, force_wrap):
    b = make()
    r1, c1 = emul(b, force_team=force_team, force_wrap=force_wrap)
    r0, c0, _ = pyoracle.run(b, "oracle", cigar_cap=96)
    assert_same(r0, c0, r1, c1, name)
