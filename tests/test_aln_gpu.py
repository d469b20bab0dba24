"""GPU parity tests of the aln stage: the product library (CUDA seeding against the device-resident index + CUDA ksw batch +
host replay), called through the C ABI, must write the same SAM as the reference's `panSVR fc_aln -t 1 -S` run on the same
box on the same synthetic inputs (byte for byte, main and `-p` outputs)."""
import os
import subprocess

import pytest

from pansvr_b200 import aln
from oracle import synth_pipeline as sp
from tests.alntest_util import DATASETS, get_demo, first_diff, golden, need_ref_tools, read

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", list(DATASETS))
def test_block_api_matches_reference_sam(name):
    need_ref_tools()
    demo = get_demo(name)
    try:
        ctx = aln.AlnContext(demo.data.index_dir, demo.data.header_sam)
        sam, ori = ctx.align_fastq(read(demo.data.reads_fq))
        hdr = ctx.header_text().encode()
        st = ctx.stats()
        ctx.close()
        assert first_diff(hdr + sam, read(demo.ref_sam)) is None
        assert first_diff(hdr + ori, read(demo.ref_ori)) is None
        assert st["reads"] == 2 * demo.data.n_pairs and st["ksw_tasks"] > 0 and st["mems"] > 0
        if name == "demo":
            assert hdr + sam == golden("aln_demo.sam.gz")
    finally:
        pass


def test_command_line_and_block_boundaries():
    """fc_aln command line of the product; and two half-blocks must give the same bytes as one block (the rand() replay and
    the per-handler random_r streams carry over between blocks)."""
    need_ref_tools()
    demo = get_demo("multi_allele")
    try:
        out, ori = os.path.join(demo.wd, "cli.sam"), os.path.join(demo.wd, "cli_ori.sam")
        rc = aln.fc_aln_main(["-t", "1", "-S", "-o", out, "-p", ori, demo.data.index_dir, demo.data.reads_fq, demo.data.header_sam])
        assert rc == 0
        assert first_diff(read(out), read(demo.ref_sam)) is None
        assert first_diff(read(ori), read(demo.ref_ori)) is None
        fq = read(demo.data.reads_fq).split(b"\n")
        half = (len(fq) // 8 // 2) * 8
        ctx = aln.AlnContext(demo.data.index_dir, demo.data.header_sam)
        s1, o1 = ctx.align_fastq(b"\n".join(fq[:half]) + b"\n")
        s2, o2 = ctx.align_fastq(b"\n".join(fq[half:]))
        hdr = ctx.header_text().encode()
        ctx.close()
        assert hdr + s1 + s2 == read(demo.ref_sam)
        assert hdr + o1 + o2 == read(demo.ref_ori)
        # BAM mode of the command line (the reference's default output) and the BAM record API written in two calls
        rb, rbo = os.path.join(demo.wd, "ref.bam"), os.path.join(demo.wd, "ref_ori.bam")
        sp.run_reference_aln(demo.data, rb, rbo, threads=1, bam=True)
        mb, mbo = os.path.join(demo.wd, "cli.bam"), os.path.join(demo.wd, "cli_ori.bam")
        assert aln.fc_aln_main(["-t", "4", "-o", mb, "-p", mbo, demo.data.index_dir, demo.data.reads_fq, demo.data.header_sam]) == 0
        assert read(mb) == read(rb) and read(mbo) == read(rbo)
        ctx = aln.AlnContext(demo.data.index_dir, demo.data.header_sam)
        b1 = ctx.align_fastq_bam(b"\n".join(fq[:half]) + b"\n")
        b2 = ctx.align_fastq_bam(b"\n".join(fq[half:]))
        ctx.write_bam(mb, [b1[0], b2[0]])
        ctx.close()
        assert read(mb) == read(rb)
    finally:
        pass


@pytest.mark.parametrize("name", ["demo", "multi_allele", "tandem_repeats", "shared_element", "chroms_250bp"])
def test_device_stages_equal_host_stepped_stages(name, tmp_path):
    """Stage-level differential (SURVEY.md 4 ii): what the CUDA stages A..F1 hand back for a block -- STR flags, MEM counts, the
    sorted seeds and chain tables of every strand, every candidate with its score and final CIGAR -- equals, byte for byte, what the
    same stage functions produce when stepped on the host (tests/emul), whose SAM in turn equals the reference's (CPU tests)."""
    need_ref_tools()
    demo = get_demo(name)
    emul = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-s", "-C", emul, os.path.join(emul, "fc_aln_emul")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "libksw_oracle.so")])
    env = dict(os.environ, PANSVR_DUMP_STAGES=str(tmp_path / "emul"), PANSVR_ORACLE_SO=os.path.join(ROOT, "oracle", "libksw_oracle.so"))
    subprocess.check_call([os.path.join(emul, "fc_aln_emul"), "-t", "2", "-S", "-o", str(tmp_path / "e.sam"), "-p", str(tmp_path / "eo.sam"),
                           demo.data.index_dir, demo.data.reads_fq, demo.data.header_sam], env=env, stderr=subprocess.DEVNULL)
    os.environ["PANSVR_DUMP_STAGES"] = str(tmp_path / "cuda")
    try:
        ctx = aln.AlnContext(demo.data.index_dir, demo.data.header_sam, threads=2)
        ctx.align_fastq(read(demo.data.reads_fq))
        st = ctx.stats()
        ctx.close()
    finally:
        del os.environ["PANSVR_DUMP_STAGES"]
    a, b = read(str(tmp_path / "cuda.0")), read(str(tmp_path / "emul.0"))
    assert len(a) > 64 and a == b
    assert st["stage_kernel_ms"] > 0 and st["seed_kernel_ms"] > 0 and st["seed_probes"] > 0


def test_device_list_one_process_same_bytes():
    """SURVEY.md 8e in one process: a context over several GPUs deals the sub-blocks of a block to the devices round robin; the
    output does not depend on it (needs a box with at least two GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    need_ref_tools()
    demo = get_demo("multi_allele")
    os.environ["PANSVR_SUB_PAIRS"] = "101"
    try:
        ctx = aln.AlnContext(demo.data.index_dir, demo.data.header_sam, device=list(range(min(4, torch.cuda.device_count()))), threads=4)
        sam, ori = ctx.align_fastq(read(demo.data.reads_fq))
        hdr = ctx.header_text().encode()
        ctx.close()
    finally:
        del os.environ["PANSVR_SUB_PAIRS"]
    assert first_diff(hdr + sam, read(demo.ref_sam)) is None
    assert hdr + ori == read(demo.ref_ori)


@pytest.mark.parametrize("name,piece_pairs", [("multi_allele", 41), ("n_bases", 64)])
def test_pieces_dealt_to_two_contexts_same_bytes(name, piece_pairs, tmp_path):
    """SURVEY.md 8e on one GPU: the input is dealt piece by piece to two contexts (what two ranks of `bench.py --gpus 2` are), each
    realigns its pieces in one pansvr_aln_pieces call from its own thread, and the in-order passes follow each other from context
    to context through the stream-state files.  Joined in piece order the output is the reference's -- with rand() ties at almost
    every pair, and with N bases (pairs the host path finishes in their turn)."""
    import ctypes
    import threading
    need_ref_tools()
    demo = get_demo(name)
    lines = read(demo.data.reads_fq).split(b"\n")
    n_pairs = len([1 for x in lines if x]) // 8
    cuts = list(range(0, n_pairs, piece_pairs)) + [n_pairs]
    texts = [b"\n".join(lines[8 * b:8 * e]) + b"\n" for b, e in zip(cuts[:-1], cuts[1:])]
    head = b"\n".join(lines[:4]) + b"\n"
    assert len(texts) >= 4
    world = 2
    ctxs = [aln.AlnContext(demo.data.index_dir, demo.data.header_sam, threads=2) for _ in range(world)]
    results, errors = {}, []

    def run(rank):
        try:
            ctx = ctxs[rank]
            ctx.prime_read_stats(head)
            mine = [b for b in range(len(texts)) if b % world == rank]
            bufs = [ctypes.create_string_buffer(texts[b], len(texts[b])) for b in mine]
            pieces = [(ctypes.addressof(buf), len(texts[b]), str(tmp_path / f"b{b}") if b > 0 else None,
                       str(tmp_path / f"b{b + 1}") if b + 1 < len(texts) else None) for b, buf in zip(mine, bufs)]
            (sa, sn), (oa, on), per, release = ctx.align_pieces(pieces)
            sam, ori = ctypes.string_at(sa, sn), ctypes.string_at(oa, on)
            release()
            so = oo = 0
            for b, (ps, po) in zip(mine, per):
                results[b] = (sam[so:so + ps], ori[oo:oo + po]); so += ps; oo += po
            assert so == sn and oo == on
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    hdr = ctxs[0].header_text().encode()
    for c in ctxs:
        c.close()
    assert not errors, errors
    assert first_diff(hdr + b"".join(results[b][0] for b in range(len(texts))), read(demo.ref_sam)) is None
    assert hdr + b"".join(results[b][1] for b in range(len(texts))) == read(demo.ref_ori)
