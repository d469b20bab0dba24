"""CPU tests of the aln pipeline's host logic (stages A, C, D, F of pansvr_b200/csrc/aln/pipeline.cpp): the pipeline is built
with host stand-ins for its two device services (tests/emul/aln_host_services.cpp: seed_core.cuh stepped on the host, ksw
through the oracle) and must reproduce the reference's `fc_aln -t 1 -S` output byte for byte."""
import ctypes as C
import os
import subprocess

import pytest

from tests.alntest_util import DATASETS, Demo, first_diff, golden, need_ref_tools, read

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def fc_aln_emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emul"), os.path.join(HERE, "emul", "fc_aln_emul")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "libksw_oracle.so")])
    env = dict(os.environ, PANSVR_ORACLE_SO=os.path.join(ROOT, "oracle", "libksw_oracle.so"))

    def run(d, out, ori, extra=()):
        subprocess.check_call([os.path.join(HERE, "emul", "fc_aln_emul"), "-t", "1", "-S", "-o", out, "-p", ori, *extra,
                               d.index_dir, d.reads_fq, d.header_sam], env=env, stderr=subprocess.DEVNULL)
    return run


def test_glibc_random_replica_matches_libc():
    """refrand.hpp must be the libc stream: compile the checker against the real rand()/random_r()."""
    src = os.path.join(HERE, "emul", "refrand_check.cpp")
    exe = os.path.join(HERE, "emul", "refrand_check")
    subprocess.check_call(["g++", "-O1", "-o", exe, src])
    assert subprocess.run([exe]).returncode == 0
    os.unlink(exe)


@pytest.mark.parametrize("name", list(DATASETS))
def test_host_pipeline_matches_reference_sam(fc_aln_emul, name):
    need_ref_tools()
    demo = Demo(name)
    try:
        out, ori = os.path.join(demo.wd, "my.sam"), os.path.join(demo.wd, "my_ori.sam")
        fc_aln_emul(demo.data, out, ori)
        assert first_diff(read(out), read(demo.ref_sam)) is None
        assert first_diff(read(out.replace("my.sam", "my_ori.sam")), read(demo.ref_ori)) is None
        assert read(out).count(b"\n") > 1000 or name != "demo"
        if name == "demo":      # and the recorded fixture of the reference's output for this seed
            assert read(demo.ref_sam) == golden("aln_demo.sam.gz")
            assert read(demo.ref_ori) == golden("aln_demo_ori.sam.gz")
    finally:
        demo.cleanup()


def test_anchor_zero_one_collapse_is_reproduced():
    """the reference's calloc'ed loader merges anchors 0 and 1; the fixture shows it, and so must we."""
    g = golden("aln_demo.sam.gz")
    assert b"SV:Z:0_" not in g and g.count(b"SV:Z:1_") > 60
