"""Generates tests/golden/ksw_golden.npz from the REFERENCE's own ksw2_extd2_sse.c.

Run in the build container (needs /root/reference): the reference file is compiled unmodified
into oracle/_ref/libksw_ref.so (oracle/Makefile) and every case below is executed through it.
The fixture stores inputs and the reference's outputs, so the GPU box (which has no
/root/reference) can check both the oracle restatement and the CUDA path against the real thing.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from pansvr_b200 import synth  # noqa: E402

CASES = [
    # name, batch factory
    ("config2", lambda: synth.config2_batch(48, seed=11, pool_bases=1 << 16)),
    ("config4_ext", lambda: synth.config4_batch(16, "ext", pool_bases=1 << 16)),
    ("config4_window", lambda: synth.config4_batch(8, "window", pool_bases=1 << 16)),
    ("config4_global", lambda: synth.config4_batch(24, "global", pool_bases=1 << 16)),
    ("pipeline_like", lambda: synth.pipeline_like_batch(160, pool_bases=1 << 16)),
    ("fuzz_w200", lambda: synth.fuzz_batch(120, 1, params=synth.KswParams(w=200, zdrop=400))),
    ("fuzz_w100", lambda: synth.fuzz_batch(120, 2, params=synth.KswParams(w=100, zdrop=400))),
    ("fuzz_w50_z100", lambda: synth.fuzz_batch(120, 3, params=synth.KswParams(w=50, zdrop=100))),
    ("fuzz_w20_z50", lambda: synth.fuzz_batch(120, 4, params=synth.KswParams(w=20, zdrop=50))),
    ("fuzz_w8", lambda: synth.fuzz_batch(80, 5, params=synth.KswParams(w=8, zdrop=30))),
    ("fuzz_w3", lambda: synth.fuzz_batch(80, 6, params=synth.KswParams(w=3, zdrop=400))),
    ("fuzz_unbanded", lambda: synth.fuzz_batch(80, 7, params=synth.KswParams(w=-1, zdrop=-1))),
    ("fuzz_extz", lambda: synth.fuzz_batch(100, 8, params=synth.KswParams(w=100, zdrop=400, flag=0x40))),
    ("fuzz_extz_bonus", lambda: synth.fuzz_batch(100, 9, params=synth.KswParams(w=100, zdrop=400, flag=0x40, end_bonus=5))),
    ("fuzz_rev", lambda: synth.fuzz_batch(80, 10, params=synth.KswParams(w=64, zdrop=200, flag=0x80))),
    ("fuzz_score_only", lambda: synth.fuzz_batch(80, 11, params=synth.KswParams(w=100, zdrop=400, flag=0x01))),
    ("fuzz_right", lambda: synth.fuzz_batch(80, 12, params=synth.KswParams(w=64, zdrop=200, flag=0x02))),
    ("fuzz_generic_sc", lambda: synth.fuzz_batch(80, 13, params=synth.KswParams(w=33, zdrop=400, flag=0x04))),
    ("fuzz_approx", lambda: synth.fuzz_batch(80, 14, params=synth.KswParams(w=100, zdrop=100, flag=0x18))),
    ("fc_sv_params", lambda: synth.fuzz_batch(100, 15, max_len=400, params=synth.KswParams(
        mat=synth.dna_matrix(2, 10), q=24, e=2, q2=32, e2=1, w=132, zdrop=132))),
    ("swapped_gaps", lambda: synth.fuzz_batch(100, 16, params=synth.KswParams(
        mat=synth.dna_matrix(1, 4, sc_ambi=-1), q=24, e=1, q2=6, e2=2, w=60, zdrop=80))),
    ("wide_w500", lambda: synth.fuzz_batch(40, 17, max_len=520, params=synth.KswParams(w=500, zdrop=400))),
]


def main():
    assert pyoracle.have_ref() or (pyoracle.build() or pyoracle.have_ref()), "reference library not built"
    out = {}
    names = []
    for name, make in CASES:
        b = make()
        res, cig, _ = pyoracle.run(b, "ref", threads=4, cigar_cap=96)
        assert not res[:, 11].any(), name
        # compact the pools to what the tasks touch
        q = np.concatenate([b.qseq[o:o + l] for o, l in zip(b.qoff, b.qlen)]) if b.n else np.zeros(0, np.uint8)
        t = np.concatenate([b.tseq[o:o + l] for o, l in zip(b.toff, b.tlen)]) if b.n else np.zeros(0, np.uint8)
        p = b.params
        out[name + "/q"] = q.astype(np.uint8)
        out[name + "/t"] = t.astype(np.uint8)
        out[name + "/qlen"] = b.qlen.astype(np.int32)
        out[name + "/tlen"] = b.tlen.astype(np.int32)
        out[name + "/params"] = np.array([p.m, p.q, p.e, p.q2, p.e2, p.w, p.zdrop, p.end_bonus, p.flag], np.int32)
        out[name + "/mat"] = np.asarray(p.mat, np.int8)
        out[name + "/res"] = res[:, :11].astype(np.int32)
        ncig = res[:, 9]
        out[name + "/cigar"] = np.concatenate([cig[i, :ncig[i]] for i in range(b.n)]).astype(np.uint32) if b.n else np.zeros(0, np.uint32)
        names.append(name)
    out["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ksw_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", sum(len(out[n + "/qlen"]) for n in names), "tasks")


if __name__ == "__main__":
    main()
