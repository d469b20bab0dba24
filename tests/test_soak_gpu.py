"""Bounded versions of the randomized soaks (tests/soak_ksw.py, tests/soak_aln.py) as `-m gpu` tests, against the REFERENCE's own
code at run time: ksw tasks against oracle/_ref/libksw_ref.so (the reference's ksw2_extd2_sse.c compiled unmodified), whole
data sets against oracle/_ref/panSVR fc_aln -t 1 (SAM and BAM files, byte for byte).  Fixed seeds; a few minutes in all."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle
from pansvr_b200 import synth
from tests.alntest_util import need_ref_tools

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCORES = [(2, 12, 16, 1, 32, 0), (2, 10, 24, 2, 32, 1), (1, 4, 6, 2, 24, 1), (4, 24, 60, 8, 100, 20), (1, 1, 1, 1, 2, 1), (2, 4, 4, 2, 13, 1),
          (2, 11, 22, 3, 14, 0), (2, 12, 16, 0, 32, 1), (2, 11, 14, 1, 22, 3), (3, 5, 4, 3, 19, 2), (1, 9, 30, 1, 13, 4)]


def _ref_impl():
    pyoracle.build()
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/libksw_ref.so not built (make -C oracle ref, where /root/reference exists)")
    return "ref"


@pytest.mark.parametrize("exotic", [False, True], ids=["regular", "exotic"])
def test_ksw_soak_against_reference_object(ksw_ctx, exotic):
    """Random band / z-drop / flags / scoring / end bonus / lengths; 'exotic' = bands and z-drops of 0-5, the generic kernel's flags,
    3 ... 1500 bp tasks.  ~100 k ragged tasks per variant, every field and every CIGAR word against the reference's object."""
    impl = _ref_impl()
    rng = np.random.default_rng(4 if exotic else 2026)
    total = 0
    for i in range(8):
        if exotic:
            w = int(rng.choice([0, 1, 2, 3, 5, 17, 600, -1])); zd = int(rng.choice([0, 1, 5, 50, 1000, -1]))
            flag = int(rng.choice([0, 0x02, 0x04, 0x08, 0x10, 0x18, 0x42, 0x82, 0x44, 0x01 | 0x02, 0x40, 0xC0]))
            ml = int(rng.choice([3, 12, 40, 260, 1500]))
        else:
            w = int(rng.choice([2, 8, 30, 50, 100, 132, 200, 500, -1])); zd = int(rng.choice([20, 100, 132, 400, -1]))
            flag = int(rng.choice([0, 0, 0, 0x40, 0x80, 0x01, 0xC0]))
            ml = int(rng.choice([60, 160, 260, 400, 700]))
        sc = SCORES[int(rng.integers(0, len(SCORES)))]
        p = synth.KswParams(mat=synth.dna_matrix(sc[0], sc[1], sc_ambi=int(rng.choice([0, -1]))), q=sc[2], e=sc[3], q2=sc[4], e2=sc[5], w=w, zdrop=zd,
                            flag=flag, end_bonus=int(rng.choice([-1, 0, 5])))
        n = 15000 if ml <= 260 else (4000 if ml <= 700 else 800)
        b = synth.fuzz_batch(n, 3000 + i + (100 if exotic else 0), max_len=ml, params=p, related=float(rng.choice([0.3, 0.8, 0.95])))
        cap = 256 if ml <= 700 else 1024
        res, cig = ksw_ctx.extd2_batch(b, cigar_cap=cap)
        r0, c0, _ = pyoracle.run(b, impl, threads=min(16, os.cpu_count() or 1), cigar_cap=cap)
        d = (res[:, :11] != r0[:, :11]).any(1)
        if not (flag & 1):
            mask = np.arange(cap)[None, :] < r0[:, 9][:, None]
            d |= ((cig != c0) & mask).any(1)
        assert not d.any(), f"set {i}: w={w} zdrop={zd} flag={flag:#x} sc={sc} max_len={ml}: {int(d.sum())} of {n} tasks differ from the reference's ksw_extd2_sse"
        total += n
    assert total > 30000


def test_config4_extension_at_size(ksw_ctx):
    """250 bp reads, w=500, extension with traceback, 1 M tasks (70 000 cells each) in one batch; a random
    sample against the reference's object, and every task's score bounded by the read length (a size-independent property)."""
    impl = _ref_impl()
    b = synth.config4_batch(1_000_000, "ext")
    res, cig = ksw_ctx.extd2_batch(b, cigar_cap=32)
    idx = np.random.default_rng(5).choice(b.n, 3000, replace=False)
    r0, c0, _ = pyoracle.run(b.take(idx), impl, threads=min(16, os.cpu_count() or 1), cigar_cap=32)
    assert np.array_equal(res[idx][:, :11], r0[:, :11])
    mask = np.arange(32)[None, :] < r0[:, 9][:, None]
    assert not ((cig[idx] != c0) & mask).any()
    assert (res[:, 4] <= 2 * 250).all() and (res[:, 4] > 0).all()          # mqe: at most match * qlen, and these reads do align


@pytest.mark.parametrize("mode,seed,n_sets", [("", 4242, 3), ("harsh", 99, 2), ("options", 7, 2)], ids=["plain", "harsh", "options"])
def test_aln_soak_against_reference(mode, seed, n_sets):
    """Random data sets (alleles per locus, N rate, tandem repeats, read length, chromosomes, lower-case / IUPAC bases, random
    helper-thread counts and sub-block cuts, random scoring options) through the product's command line in SAM and BAM mode,
    byte for byte against `panSVR fc_aln -t 1`."""
    need_ref_tools()
    cmd = [sys.executable, os.path.join(ROOT, "tests", "soak_aln.py"), str(n_sets), str(seed)] + ([mode] if mode else [])
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420)          # (about 8 s per set on a B200 box)
    assert p.returncode == 0 and f"sets {n_sets} failures 0" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
