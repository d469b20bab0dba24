"""One-off randomized soak of the ksw path (not collected by pytest): 24 random parameter sets (band, z-drop, flags, scoring,
end bonus, lengths) x 15-60 k ragged tasks each through the C ABI on the GPU, against the reference's own compiled
ksw_extd2_sse (oracle/_ref) when present, else the oracle port.  `python tests/soak_ksw.py [exotic [seed]]` on a GPU box ("exotic": bands and z-drops of 0-5, the generic kernel's flags, 3 ... 1500 bp tasks); results in
profiles/r1t_soak.md."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from pansvr_b200 import ksw, synth
from oracle import pyoracle
ctx = ksw.KswContext(0)
rng = np.random.default_rng(2026)
tot = bad = 0
t0 = time.time()
sets = []
exotic = len(sys.argv) > 1 and sys.argv[1] == "exotic"      # tiny bands and z-drops, the generic kernel's flags, very short / long tasks
if exotic:
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
for i in range(24):
    w = int(rng.choice([2, 8, 30, 50, 100, 132, 200, 500, -1]))
    zd = int(rng.choice([20, 100, 132, 400, -1]))
    flag = int(rng.choice([0, 0, 0, 0x40, 0x80, 0x01, 0xC0]))
    if exotic:
        w = int(rng.choice([0, 1, 2, 3, 5, 17, 600, -1]))
        zd = int(rng.choice([0, 1, 5, 50, 1000, -1]))
        flag = int(rng.choice([0, 0x02, 0x04, 0x08, 0x10, 0x18, 0x42, 0x82, 0x44, 0x01 | 0x02, 0x40, 0xC0]))
    sc = [(2, 12, 16, 1, 32, 0), (2, 10, 24, 2, 32, 1), (1, 4, 6, 2, 24, 1), (4, 24, 60, 8, 100, 20), (1, 1, 1, 1, 2, 1), (2, 4, 4, 2, 13, 1),
              (2, 11, 22, 3, 14, 0), (2, 12, 16, 0, 32, 1), (2, 11, 14, 1, 22, 3), (3, 5, 4, 3, 19, 2), (1, 9, 30, 1, 13, 4)][int(rng.integers(0, 11))]
    p = synth.KswParams(mat=synth.dna_matrix(sc[0], sc[1], sc_ambi=int(rng.choice([0, -1]))), q=sc[2], e=sc[3], q2=sc[4], e2=sc[5], w=w, zdrop=zd, flag=flag,
                        end_bonus=int(rng.choice([-1, 0, 5])))
    ml = int(rng.choice([60, 160, 260, 400, 700]))
    if exotic:
        ml = int(rng.choice([3, 12, 40, 260, 1500]))
    n = 60000 if ml <= 260 else (15000 if ml <= 700 else 3000)
    b = synth.fuzz_batch(n, 1000 + i, max_len=ml, params=p, related=float(rng.choice([0.3, 0.8, 0.95])))
    cap = 256 if ml <= 700 else 1024
    res, cig = ctx.extd2_batch(b, cigar_cap=cap)
    r0, c0, _ = pyoracle.run(b, "ref" if pyoracle.have_ref() else "oracle", threads=16, cigar_cap=cap)
    score_only = bool(flag & 1)
    d = (res[:, :11] != r0[:, :11]).any(1)
    if not score_only:
        nc = r0[:, 9]
        mask = np.arange(cap)[None, :] < nc[:, None]
        d |= ((cig != c0) & mask).any(1)
    st = ctx.stats()
    tot += n; bad += int(d.sum())
    print(f"set {i}: w={w} zdrop={zd} flag={flag:#x} sc={sc} max_len={ml} n={n} mismatches={int(d.sum())} generic={st['tasks_generic']} wrap={st['tasks_fast_wrap']} nowrap={st['tasks_fast_nowrap']}", flush=True)
print("total", tot, "mismatches", bad, "%.1f s" % (time.time() - t0))
