"""CPU tests of the oracle: the scalar restatement against the reference's own outputs."""
import numpy as np
import pytest

from oracle import pyoracle
from pansvr_b200 import synth
from tests.kswtest_util import assert_matches_golden, assert_same, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_fixture(name):
    b, gres, gcigs = load_golden(name)
    res, cig, _ = pyoracle.run(b, "oracle", threads=2, cigar_cap=96)
    assert not res[:, 11].any()
    assert_matches_golden(res, cig, gres, gcigs, name)


@pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("seed,w,zdrop,flag", [(21, 200, 400, 0), (22, 100, 400, 0), (23, 40, 60, 0), (24, 100, 400, 0x40),
                                                (25, 64, 200, 0x02), (26, 100, 100, 0x18), (27, 33, 400, 0x04), (28, 5, 400, 0)])
def test_oracle_matches_live_reference(seed, w, zdrop, flag):
    b = synth.fuzz_batch(400, seed, params=synth.KswParams(w=w, zdrop=zdrop, flag=flag))
    r0, c0, _ = pyoracle.run(b, "ref", threads=2, cigar_cap=128)
    r1, c1, _ = pyoracle.run(b, "oracle", threads=2, cigar_cap=128)
    assert_same(r0, c0, r1, c1, f"seed {seed}")


@pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built (reference sources absent)")
def test_oracle_matches_live_reference_on_fc_sv_contigs():
    b = synth.fcsv_batch(60)                                   # SignalAssembly.hpp:411-421,459-464
    r0, c0, _ = pyoracle.run(b, "ref", threads=4, cigar_cap=128)
    r1, c1, _ = pyoracle.run(b, "oracle", threads=4, cigar_cap=128)
    assert_same(r0, c0, r1, c1, b.name)


def test_cells_definition():
    # SURVEY.md section 8d table
    assert pyoracle.cells(150, 180, 200) == 27000
    assert pyoracle.cells(150, 180, 100) == 22615
    assert pyoracle.cells(150, 1100, 100) == 25100
    assert pyoracle.cells(250, 280, 500) == 70000
    assert pyoracle.cells(250, 280, 200) == 65615
    assert pyoracle.cells(250, 1500, 500) == 156375
    assert synth.band_cells(150, 1100, 100) == 25100
    assert synth.band_cells(250, 1500, 500) == 156375


def test_config2_shape_and_behaviour():
    b = synth.config2_batch(64, pool_bases=1 << 16)
    assert b.n == 64 and (b.qlen == 150).all() and (b.tlen == 1100).all() and b.params.w == 100
    res, cig, _ = pyoracle.run(b, "oracle")
    # the band closes after 399 diagonals: zdropped, no end-to-end score, mqe valid, CIGAR from the max cell
    assert (res[:, 1] == 1).all() and (res[:, 8] == synth.KSW_NEG_INF).all()
    assert (res[:, 4] > 200).all() and (res[:, 9] > 0).all()


def test_empty_and_degenerate_inputs():
    p = synth.KswParams()
    one = np.array([1], np.uint8)
    b = synth.KswBatch(one, np.zeros(3, np.int64), np.array([0, 1, 1], np.int32), one, np.zeros(3, np.int64),
                       np.array([1, 0, 1], np.int32), p)
    res, cig, _ = pyoracle.run(b, "oracle")
    reset = [0, 0, -1, -1, synth.KSW_NEG_INF, -1, synth.KSW_NEG_INF, -1, synth.KSW_NEG_INF, 0, 0]
    assert res[0, :11].tolist() == reset and res[1, :11].tolist() == reset
    assert res[2, 8] == 2 and res[2, 9] == 1 and cig[2, 0] == (1 << 4 | 0)
