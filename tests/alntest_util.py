"""Helpers of the aln (pipeline-level) tests: seeded demo data sets and the reference's SAM for them."""
import gzip
import os
import shutil
import tempfile

import pytest

from oracle import synth_pipeline as sp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DATASETS = {
    # name: make_demo kwargs (config 1; config-3-like shared flanks; reads with N)
    "demo": dict(),
    "multi_allele": dict(seed=21, genome_len=300_000, n_sv=40, alleles_per_locus=3, pairs_per_sv=30),
    "n_bases": dict(seed=22, n_sv=20, pairs_per_sv=30, n_frac=0.004),
    # three chromosomes, a fifth of the original mates elsewhere (RNAME / RNEXT ids), 250 bp reads (BASELINE configs[3] shape)
    "chroms_250bp": dict(seed=23, genome_len=120_000, n_sv=24, alleles_per_locus=2, pairs_per_sv=25, read_len=250, frag=(500, 700),
                         n_chrom=3, mate_elsewhere=0.2, n_frac=0.0005),
    # inserted alleles that are short tandem repeats: the STR branch of the seeding loop and of the chain parameters
    "tandem_repeats": dict(seed=24, genome_len=160_000, n_sv=40, sv_lens=(120, 200, 300, 500), str_every=2, pairs_per_sv=30),
    # 3 kb tandem-repeat alleles (one unipath with > 500 positions -> random_r) together with N bases: the random_r streams must be
    # consumed in input order across batched reads, N variants and reads prepared during the replay (found by tests/soak_aln.py)
    "repeat_with_n": dict(seed=481, n_sv=16, pairs_per_sv=24, sub_rate=0.01, n_frac=0.001, str_every=1, sv_lens=(3000, 3000),
                          genome_len=5000 + 7000 * 16 + 4000, frag=(300, 464)),
    # one 150 bp element shared by 520 inserted alleles: its unipath has > 500 positions, expand_seed samples them with random_r
    "shared_element": dict(seed=25, genome_len=5000 + 7000 * 520 + 4000, n_sv=520, alleles_per_locus=2, pairs_per_sv=2,
                           shared_insert=150, sv_lens=(200, 300, 400)),
}


def need_ref_tools():
    if not sp.have_reference_tools():
        pytest.skip("oracle/_ref/panSVR and deBGA not built (oracle/build_ref_pipeline.sh)")


class Demo:
    """A data set on disk plus the reference's output for it (run once per session)."""

    def __init__(self, name):
        self.name = name
        self.wd = tempfile.mkdtemp(prefix=f"pansvr_{name}_")
        self.data = sp.make_demo(self.wd, **DATASETS[name])
        self.ref_sam = os.path.join(self.wd, "ref_out.sam")
        self.ref_ori = os.path.join(self.wd, "ref_ori.sam")
        sp.run_reference_aln(self.data, self.ref_sam, self.ref_ori, threads=1)

    def cleanup(self):
        shutil.rmtree(self.wd, ignore_errors=True)


_DEMOS = {}


def get_demo(name):
    """One shared Demo per data set and test session (index build + reference run are the expensive part)."""
    import atexit
    if name not in _DEMOS:
        _DEMOS[name] = Demo(name)
        atexit.register(_DEMOS[name].cleanup)
    return _DEMOS[name]


def read(path):
    with open(path, "rb") as f:
        return f.read()


def golden(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name), "rb") as f:
        return f.read()


def first_diff(a: bytes, b: bytes):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return i, x[:300], y[:300]
    return (min(len(la), len(lb)), b"<eof>", b"<eof>") if len(la) != len(lb) else None
