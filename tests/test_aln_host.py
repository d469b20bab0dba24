"""CPU tests of the aln pipeline's host logic (stages A, C, D, F of pansvr_b200/csrc/aln/pipeline.cpp): the pipeline is built
with host stand-ins for its two device services (tests/emul/aln_host_services.cpp: seed_core.cuh stepped on the host, ksw
through the oracle) and must reproduce the reference's `fc_aln -t 1 -S` output byte for byte."""
import ctypes as C
import gzip
import os
import subprocess

import pytest

from oracle import synth_pipeline as sp
from tests.alntest_util import DATASETS, get_demo, first_diff, golden, need_ref_tools, read

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def fc_aln_emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emul"), os.path.join(HERE, "emul", "fc_aln_emul")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "libksw_oracle.so")])
    env = dict(os.environ, PANSVR_ORACLE_SO=os.path.join(ROOT, "oracle", "libksw_oracle.so"))

    def run(d, out, ori, extra=("-S",), threads=1, sub_pairs=0):
        e = dict(os.environ, **{k: env[k] for k in ("PANSVR_ORACLE_SO",)})      # (the environment as it is at the call)
        if sub_pairs:
            e["PANSVR_SUB_PAIRS"] = str(sub_pairs)
        subprocess.check_call([os.path.join(HERE, "emul", "fc_aln_emul"), "-t", str(threads), "-o", out, "-p", ori, *extra,
                               d.index_dir, d.reads_fq, d.header_sam], env=e, stderr=subprocess.DEVNULL)
    return run


def test_glibc_random_replica_matches_libc():
    """refrand.hpp must be the libc stream: compile the checker against the real rand()/random_r()."""
    src = os.path.join(HERE, "emul", "refrand_check.cpp")
    exe = os.path.join(HERE, "emul", "refrand_check")
    subprocess.check_call(["g++", "-O1", "-o", exe, src])
    assert subprocess.run([exe]).returncode == 0
    os.unlink(exe)


def test_wordwise_mem_extension_matches_base_by_base():
    """seed_core.cuh extends MEMs with 32-base XOR windows; the reference walks base by base (deBGA_index.cpp:118-131)."""
    src = os.path.join(HERE, "emul", "seed_check.cpp")
    exe = os.path.join(HERE, "emul", "seed_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", src, "-o", exe])
    try:
        assert subprocess.run([exe], stdout=subprocess.DEVNULL).returncode == 0
    finally:
        os.unlink(exe)


@pytest.mark.parametrize("name", list(DATASETS))
def test_host_pipeline_matches_reference_sam(fc_aln_emul, name):
    need_ref_tools()
    demo = get_demo(name)
    try:
        out, ori = os.path.join(demo.wd, "my.sam"), os.path.join(demo.wd, "my_ori.sam")
        fc_aln_emul(demo.data, out, ori)
        assert first_diff(read(out), read(demo.ref_sam)) is None
        assert first_diff(read(out.replace("my.sam", "my_ori.sam")), read(demo.ref_ori)) is None
        assert read(out).count(b"\n") > 1000 or name != "demo"
        if name == "demo":      # and the recorded fixture of the reference's output for this seed
            assert read(demo.ref_sam) == golden("aln_demo.sam.gz")
            assert read(demo.ref_ori) == golden("aln_demo_ori.sam.gz")
        # the block cut into overlapping sub-blocks (two in flight, in-order sections passed on by sequence number)
        fc_aln_emul(demo.data, out, ori, threads=4, sub_pairs=97)
        assert read(out) == read(demo.ref_sam) and read(ori) == read(demo.ref_ori)
        # BAM mode (the reference's default): whole files byte-identical, BGZF blocks and deflate streams included
        rb, rbo = os.path.join(demo.wd, "ref.bam"), os.path.join(demo.wd, "ref_ori.bam")
        sp.run_reference_aln(demo.data, rb, rbo, threads=1, bam=True)
        mb, mbo = os.path.join(demo.wd, "my.bam"), os.path.join(demo.wd, "my_ori.bam")
        fc_aln_emul(demo.data, mb, mbo, extra=(), threads=3)
        assert read(mb) == read(rb) and read(mbo) == read(rbo)
        assert gzip.decompress(read(mb))[:4] == b"BAM\x01"
    finally:
        pass


def test_device_list_round_robin(fc_aln_emul):
    """`-d 0,1,2`: the sub-blocks of a block are dealt to the devices of the list round robin, each with its own stage service
    (on the host-stepped build the device numbers are only labels); same bytes as one device."""
    need_ref_tools()
    demo = get_demo("multi_allele")
    out, ori = os.path.join(demo.wd, "dl.sam"), os.path.join(demo.wd, "dl_ori.sam")
    fc_aln_emul(demo.data, out, ori, extra=("-S", "-d", "0,1,2"), threads=4, sub_pairs=97)
    assert first_diff(read(out), read(demo.ref_sam)) is None
    assert read(ori) == read(demo.ref_ori)


def test_output_room_spill(fc_aln_emul):
    """The device path writes every sub-block's SAM text straight behind the previous one's in one output buffer; when that
    buffer is too small (forced here) the remaining text is assembled by copying -- same bytes either way."""
    need_ref_tools()
    demo = get_demo("multi_allele")
    out, ori = os.path.join(demo.wd, "spill.sam"), os.path.join(demo.wd, "spill_ori.sam")
    for room in ("150000", "1"):
        os.environ["PANSVR_ROOM_BYTES"] = room
        try:
            fc_aln_emul(demo.data, out, ori, threads=4, sub_pairs=97)
        finally:
            del os.environ["PANSVR_ROOM_BYTES"]
        assert first_diff(read(out), read(demo.ref_sam)) is None
        assert read(ori) == read(demo.ref_ori)


def test_edge_inputs_match_reference(fc_aln_emul):
    """Empty input, a single pair, a dangling unpaired read and a missing final newline: same files as the reference."""
    need_ref_tools()
    demo = get_demo("demo")
    try:
        fq = read(demo.data.reads_fq).decode().split("\n")
        cases = {"empty": "", "one_pair": "\n".join(fq[:8]) + "\n", "three_reads": "\n".join(fq[:12]) + "\n",
                 "no_final_newline": "\n".join(fq[:16])}
        for name, text in cases.items():
            path = os.path.join(demo.wd, name + ".fq")
            with open(path, "w") as f:
                f.write(text)
            d = sp.PipelineData(demo.data.workdir, demo.data.ref_fa, demo.data.vcf, demo.data.anchors_fa, demo.data.index_dir, path,
                                demo.data.header_sam, 0, 0)
            r, ro = os.path.join(demo.wd, name + "_ref.sam"), os.path.join(demo.wd, name + "_ref_ori.sam")
            m, mo = os.path.join(demo.wd, name + "_my.sam"), os.path.join(demo.wd, name + "_my_ori.sam")
            sp.run_reference_aln(d, r, ro, threads=1)
            fc_aln_emul(d, m, mo, threads=2)
            assert read(m) == read(r) and read(mo) == read(ro), name
        # lower-case bases (incl. 'n', which the reference encodes as 4 and lets spill into the neighbouring base), IUPAC codes
        # and a junk character: SEQ comes back through htslib's 4-bit round trip
        import random
        rnd = random.Random(3)
        lc = list(fq)
        for i in range(1, len(lc), 4):
            lc[i] = "".join((c.lower() if rnd.random() < 0.05 else ("NRy."[rnd.randrange(4)] if rnd.random() < 0.004 else c)) for c in lc[i])
        path = os.path.join(demo.wd, "lower.fq")
        with open(path, "w") as f:
            f.write("\n".join(lc))
        d = sp.PipelineData(demo.data.workdir, demo.data.ref_fa, demo.data.vcf, demo.data.anchors_fa, demo.data.index_dir, path,
                            demo.data.header_sam, 0, 0)
        r, ro = os.path.join(demo.wd, "lc_ref.sam"), os.path.join(demo.wd, "lc_ref_ori.sam")
        m, mo = os.path.join(demo.wd, "lc_my.sam"), os.path.join(demo.wd, "lc_my_ori.sam")
        sp.run_reference_aln(d, r, ro, threads=1)
        fc_aln_emul(d, m, mo, threads=2)
        assert read(m) == read(r) and read(mo) == read(ro)
        # gzip-compressed input, and -R (max_use_read) cutting the input after 100 pairs
        gz = os.path.join(demo.wd, "reads.fq.gz")
        with open(demo.data.reads_fq, "rb") as f, gzip.open(gz, "wb", compresslevel=1) as g:
            g.write(f.read())
        d = sp.PipelineData(demo.data.workdir, demo.data.ref_fa, demo.data.vcf, demo.data.anchors_fa, demo.data.index_dir, gz,
                            demo.data.header_sam, 0, 0)
        m, mo = os.path.join(demo.wd, "gz_my.sam"), os.path.join(demo.wd, "gz_my_ori.sam")
        fc_aln_emul(d, m, mo, threads=2)
        assert read(m) == read(demo.ref_sam) and read(mo) == read(demo.ref_ori)
        r, ro = os.path.join(demo.wd, "r100_ref.sam"), os.path.join(demo.wd, "r100_ref_ori.sam")
        sp.run_reference_aln(demo.data, r, ro, threads=1, extra=("-R", "100"))
        fc_aln_emul(demo.data, m, mo, extra=("-S", "-R", "100"), threads=2)
        assert read(m) == read(r) and read(mo) == read(ro)
    finally:
        pass


@pytest.mark.parametrize("opts", [("-Q",), ("-M", "1", "-m", "4", "-O", "6", "-E", "2", "-P", "24", "-F", "1", "-z", "200"),
                                  ("-M", "3", "-m", "9", "-O", "20", "-E", "3", "-P", "40", "-F", "1", "-z", "100", "-Q", "-w", "50"),
                                  ("-E", "0", "-O", "12"), ("-R", "0")],
                         ids=["not_ori", "soft_scoring", "hard_scoring", "explicit_zero", "no_reads"])
def test_command_line_options_match_reference(fc_aln_emul, opts):
    """MAP_PARA::get_option (read_realignment.hpp:82-128): scoring, z-drop, -Q and the ignored -w."""
    need_ref_tools()
    demo = get_demo("n_bases")
    try:
        r, ro = os.path.join(demo.wd, "opt_ref.sam"), os.path.join(demo.wd, "opt_ref_ori.sam")
        m, mo = os.path.join(demo.wd, "opt_my.sam"), os.path.join(demo.wd, "opt_my_ori.sam")
        sp.run_reference_aln(demo.data, r, ro, threads=1, extra=opts)
        fc_aln_emul(demo.data, m, mo, extra=("-S",) + tuple(opts), threads=3)
        assert read(m) == read(r) and read(mo) == read(ro)
    finally:
        pass


def test_iupac_bases_after_in_place_reverse_complement(fc_aln_emul):
    """The reference reverse-complements its kseq_t in place around the write of a reverse-strand record and back again, which
    turns every non-ACGT base into N for whatever is written later for that read (the -p record).  Found by tests/soak_aln.py."""
    need_ref_tools()
    import random
    demo = get_demo("tandem_repeats")
    rnd = random.Random(11)
    lines = read(demo.data.reads_fq).decode().split("\n")
    for i in range(1, len(lines), 4):
        lines[i] = "".join(("RYKMSWn."[rnd.randrange(8)] if rnd.random() < 0.02 else (c.lower() if rnd.random() < 0.05 else c)) for c in lines[i])
    path = os.path.join(demo.wd, "iupac.fq")
    with open(path, "w") as f:
        f.write("\n".join(lines))
    d = sp.PipelineData(demo.data.workdir, demo.data.ref_fa, demo.data.vcf, demo.data.anchors_fa, demo.data.index_dir, path,
                        demo.data.header_sam, 0, 0)
    r, ro = os.path.join(demo.wd, "iu_ref.sam"), os.path.join(demo.wd, "iu_ref_ori.sam")
    m, mo = os.path.join(demo.wd, "iu_my.sam"), os.path.join(demo.wd, "iu_my_ori.sam")
    sp.run_reference_aln(d, r, ro, threads=1)
    fc_aln_emul(d, m, mo, threads=3)
    assert read(m) == read(r) and read(mo) == read(ro)
    assert read(ro).count(b"\n") > 20


def test_short_cigar_records_are_left_out(fc_aln_emul):
    """With a z-drop far below the default an extension can stop early and leave a CIGAR shorter than the read.  The reference
    logs an error, htslib rejects the record and the reference writes the half-parsed record with stale buffer bytes; we leave
    such records out.  Every well-formed record of the reference and its whole -p file must still be ours, and BAM mode must
    not fail on them."""
    need_ref_tools()
    demo = get_demo("chroms_250bp")
    opts = ("-z", "60", "-O", "10", "-E", "2")
    r, ro = os.path.join(demo.wd, "z_ref.sam"), os.path.join(demo.wd, "z_ref_ori.sam")
    m, mo = os.path.join(demo.wd, "z_my.sam"), os.path.join(demo.wd, "z_my_ori.sam")
    sp.run_reference_aln(demo.data, r, ro, threads=1, extra=opts)
    n_bad = read(os.path.join(demo.wd, "fc_aln.log")).count(b"different length")
    assert n_bad > 0                                     # the data set does hit the case
    fc_aln_emul(demo.data, m, mo, extra=("-S",) + opts, threads=3)
    well_formed = [ln for ln in read(r).split(b"\n") if ln and not ln.startswith(b"@") and len(ln.split(b"\t")) > 11
                   and ln.split(b"\t")[11].startswith(b"AS:i:")]
    mine = [ln for ln in read(m).split(b"\n") if ln and not ln.startswith(b"@")]
    assert mine == well_formed
    assert read(mo) == read(ro)
    fc_aln_emul(demo.data, os.path.join(demo.wd, "z_my.bam"), os.path.join(demo.wd, "z_my_ori.bam"), extra=opts, threads=3)
    assert gzip.decompress(read(os.path.join(demo.wd, "z_my.bam")))[:4] == b"BAM\x01"


def test_bam_records_and_writer_across_calls(fc_aln_emul):
    """pansvr_aln_block_bam + pansvr_bam_open/write/close through the C ABI of the host build: records written in several
    calls (open BGZF block carried over) give the reference's BAM file; every record equals its SAM line re-encoded."""
    need_ref_tools()
    from pansvr_b200 import aln
    os.environ["PANSVR_ORACLE_SO"] = os.path.join(ROOT, "oracle", "libksw_oracle.so")
    lib = C.CDLL(os.path.join(HERE, "emul", "libaln_emul.so"))
    demo = get_demo("multi_allele")
    try:
        rb, rbo = os.path.join(demo.wd, "ref.bam"), os.path.join(demo.wd, "ref_ori.bam")
        sp.run_reference_aln(demo.data, rb, rbo, threads=1, bam=True)
        fq = read(demo.data.reads_fq).split(b"\n")
        cuts = [0, (len(fq) // 8 // 3) * 8, (len(fq) // 8 // 2) * 8, len(fq)]
        ctx = aln.AlnContext(demo.data.index_dir, demo.data.header_sam, lib=lib, threads=4)
        chunks = [ctx.align_fastq_bam(b"\n".join(fq[a:b]) + b"\n") for a, b in zip(cuts[:-1], cuts[1:])]
        mb, mbo = os.path.join(demo.wd, "my.bam"), os.path.join(demo.wd, "my_ori.bam")
        ctx.write_bam(mb, [c[0] for c in chunks])
        ctx.write_bam(mbo, [c[1] for c in chunks])
        ctx.close()
        assert read(mb) == read(rb) and read(mbo) == read(rbo)
        # the uncompressed stream is header + exactly our records
        raw = gzip.decompress(read(rb))
        assert raw.endswith(b"".join(c[0] for c in chunks))
    finally:
        pass


def test_anchor_zero_one_collapse_is_reproduced():
    """the reference's calloc'ed loader merges anchors 0 and 1; the fixture shows it, and so must we."""
    g = golden("aln_demo.sam.gz")
    assert b"SV:Z:0_" not in g and g.count(b"SV:Z:1_") > 60


def _bgzf(data: bytes, block=60000) -> bytes:
    """BGZF as htslib's bgzip writes it: independent gzip members with a 'BC' extra field holding the member size, then the
    empty end-of-file member (SAM specification 4.1)."""
    import struct, zlib
    out = []
    for at in list(range(0, len(data), block)) + [None]:
        chunk = b"" if at is None else data[at:at + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 18 + len(body) + 8 - 1
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + body + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return b"".join(out)


def test_bgzf_gzip_and_pipe_inputs(fc_aln_emul):
    """the command line takes plain text, gzip or BGZF, from a file or from standard input; BGZF members are
    inflated on all helper threads (text_in.hpp).  Same files as the reference's for each."""
    need_ref_tools()
    demo = get_demo("multi_allele")
    text = read(demo.data.reads_fq)
    exe = os.path.join(HERE, "emul", "fc_aln_emul")
    env = dict(os.environ, PANSVR_ORACLE_SO=os.path.join(ROOT, "oracle", "libksw_oracle.so"))
    forms = {"bgzf": _bgzf(text, 7001), "gzip": gzip.compress(text, 1), "gzip2": gzip.compress(text[:len(text) // 2]) + gzip.compress(text[len(text) // 2:]),
             "plain": text}
    for name, blob in forms.items():
        path = os.path.join(demo.wd, "in_" + name)
        with open(path, "wb") as f:
            f.write(blob)
        for how in ("file", "pipe"):
            out, ori = os.path.join(demo.wd, f"tin_{name}_{how}.sam"), os.path.join(demo.wd, f"tin_{name}_{how}_ori.sam")
            argv = [exe, "-t", "3", "-S", "-o", out, "-p", ori, demo.data.index_dir, path if how == "file" else "-", demo.data.header_sam]
            with open(path, "rb") as f:
                subprocess.check_call(argv, env=env, stdin=f if how == "pipe" else subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert first_diff(read(out), read(demo.ref_sam)) is None, (name, how)
            assert read(ori) == read(demo.ref_ori), (name, how)
    # a damaged BGZF member is an error, not silently shorter input
    bad = bytearray(forms["bgzf"]); bad[len(bad) // 2] ^= 0x55
    path = os.path.join(demo.wd, "in_bad")
    with open(path, "wb") as f:
        f.write(bytes(bad))
    p = subprocess.run([exe, "-t", "2", "-S", "-o", os.path.join(demo.wd, "bad.sam"), "-p", os.path.join(demo.wd, "bad_ori.sam"), demo.data.index_dir, path,
                        demo.data.header_sam], env=env, capture_output=True)
    assert p.returncode != 0 and b"BGZF" in p.stderr


def test_tie_pairs_in_order_equal_host_path(fc_aln_emul):
    """Pairs whose rand() ties decide their outcome are finished by the in-order pass from the device stages' results; with
    PANSVR_TIES_ON_HOST_PATH=1 the host path re-does them from scratch (the earlier way).  Same files either way, and the
    reference's, on the data set where a tenth of the pairs are of that kind (tandem-repeat alleles)."""
    need_ref_tools()
    demo = get_demo("tandem_repeats")
    outs = {}
    for mode in ("in_order", "host_path"):
        out, ori = os.path.join(demo.wd, f"tie_{mode}.sam"), os.path.join(demo.wd, f"tie_{mode}_ori.sam")
        if mode == "host_path":
            os.environ["PANSVR_TIES_ON_HOST_PATH"] = "1"
        try:
            fc_aln_emul(demo.data, out, ori, threads=3, sub_pairs=200)
        finally:
            os.environ.pop("PANSVR_TIES_ON_HOST_PATH", None)
        outs[mode] = (read(out), read(ori))
    assert outs["in_order"] == outs["host_path"]
    assert first_diff(outs["in_order"][0], read(demo.ref_sam)) is None and outs["in_order"][1] == read(demo.ref_ori)


def test_failed_sub_block_is_an_error_not_a_hang():
    """A sub-block that fails on its way (here: injected into the first trip of sub-block 1 of 5+) must still let the sub-blocks
    behind it have their in-order turn: the call returns the error (round 2: a failed launch used to leave them waiting)."""
    need_ref_tools()
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emul"), os.path.join(HERE, "emul", "fc_aln_emul")])
    demo = get_demo("multi_allele")
    env = dict(os.environ, PANSVR_ORACLE_SO=os.path.join(ROOT, "oracle", "libksw_oracle.so"), PANSVR_SUB_PAIRS="100", PANSVR_TEST_FAIL_SEQ="1")
    p = subprocess.run([os.path.join(HERE, "emul", "fc_aln_emul"), "-t", "3", "-S", "-o", os.path.join(demo.wd, "fail.sam"), "-p",
                        os.path.join(demo.wd, "fail_ori.sam"), demo.data.index_dir, demo.data.reads_fq, demo.data.header_sam],
                       env=env, capture_output=True, timeout=120)
    assert p.returncode != 0 and b"injected failure" in p.stderr


def test_pieces_api_edges_two_contexts(fc_aln_emul, tmp_path):
    """pansvr_aln_pieces on the host-stepped build, two contexts in two threads: pieces of uneven size, an EMPTY piece in the middle
    of the chain (it still takes the stream state over and hands it on), a context that owns consecutive pieces; and one piece
    without files = pansvr_aln_block."""
    import ctypes
    import threading
    from pansvr_b200 import aln
    need_ref_tools()
    os.environ["PANSVR_ORACLE_SO"] = os.path.join(ROOT, "oracle", "libksw_oracle.so")
    lib = ctypes.CDLL(os.path.join(HERE, "emul", "libaln_emul.so"))
    demo = get_demo("multi_allele")
    lines = read(demo.data.reads_fq).split(b"\n")
    n_pairs = len([1 for x in lines if x]) // 8
    cuts = [0, 13, 13, 200, 201, 640, n_pairs]                        # piece 1 is empty, piece 3 is a single pair
    texts = [b"\n".join(lines[8 * b:8 * e]) + (b"\n" if e > b else b"") for b, e in zip(cuts[:-1], cuts[1:])]
    owner = [0, 1, 1, 0, 1, 0]                                        # (context 1 owns pieces 1 and 2: consecutive)
    head = b"\n".join(lines[:4]) + b"\n"
    ctxs = [aln.AlnContext(demo.data.index_dir, demo.data.header_sam, lib=lib, threads=2) for _ in range(2)]
    results, errors = {}, []

    def run(rank):
        try:
            ctx = ctxs[rank]
            ctx.prime_read_stats(head)
            mine = [b for b in range(len(texts)) if owner[b] == rank]
            bufs = [ctypes.create_string_buffer(texts[b], max(len(texts[b]), 1)) for b in mine]
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
            errors.append(repr(e))

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    hdr = ctxs[0].header_text().encode()
    assert not errors, errors
    assert results[1] == (b"", b"")
    assert first_diff(hdr + b"".join(results[b][0] for b in range(len(texts))), read(demo.ref_sam)) is None
    assert hdr + b"".join(results[b][1] for b in range(len(texts))) == read(demo.ref_ori)
    # one piece, no files: the block call
    ctxs[0].reset()
    whole = read(demo.data.reads_fq)
    buf = ctypes.create_string_buffer(whole, len(whole))
    (sa, sn), (oa, on), per, release = ctxs[0].align_pieces([(ctypes.addressof(buf), len(whole), None, None)])
    assert hdr + ctypes.string_at(sa, sn) == read(demo.ref_sam) and per == [(sn, on)]
    release()
    for c in ctxs:
        c.close()
