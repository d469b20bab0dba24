"""CPU tests of the N > 1 host path: world_size-2 `gloo` process groups exercise the sharding helpers, the
max-over-ranks of the timings and the rank-0-only printing of bench.py's reference arm."""
import json
import os
import subprocess
import sys

import pytest

from pansvr_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from pansvr_b200 import shard
rank, local, world = shard.rank_env()
dist.init_process_group("gloo")
assert dist.get_world_size() == world == 2
# every rank times its own shard; the job's time is the slowest rank's
mine = [10.0 + 5.0 * rank, 3.0 - rank]
mx = shard.max_over_ranks(mine, dist)
b, e = shard.shard_range(1001, rank, world)
# results are merged by pair index: gather the ranges and check they tile [0, n)
got = [None, None]
dist.all_gather_object(got, (b, e))
seed = shard.shard_seed(11, rank)
dist.barrier()
if rank == 0:
    print(json.dumps({"max": mx, "ranges": got, "seed": seed, "rate": shard.whole_job_rate(1000, 5, world, mx[0])}))
dist.destroy_process_group()
"""


def _torchrun(args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533"] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_shard_range_tiles_the_block():
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            parts = [shard.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_two_ranks_gloo_max_and_ranges(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    p = _torchrun([str(w), ROOT])
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["max"] == [15.0, 3.0]                      # max over ranks, element-wise
    assert line["ranges"] == [[0, 501], [501, 1001]]
    assert line["seed"] == 11
    assert abs(line["rate"] - 2 * 1000 * 5 / 15e-3) < 1e-6


def test_reference_arm_two_ranks_prints_once(tmp_path):
    """bench.py --impl reference under torchrun: rank 0 alone runs the CPU arm (the reference's own `panSVR fc_aln`) and prints one
    JSON line, rank 1 exits 0."""
    from tests.alntest_util import need_ref_tools
    need_ref_tools()
    os.environ["PANSVR_BENCH_CACHE"] = str(tmp_path)
    try:
        p = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--loci", "24", "--pairs", "4096",
                       "--ref-sample-pairs", "1500"], timeout=600)
    finally:
        del os.environ["PANSVR_BENCH_CACHE"]
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and "fc_aln" in d["metric"]


SHARD_WORKER = r"""
# one input, two processes: rank r realigns its contiguous range of pairs on the host-emulated pipeline (tests/emul), takes the
# reference's rand()/random_r() streams where rank r-1 left them (pansvr_aln_await_state / publish_state) and rank 0 joins the
# outputs in rank order = input order
import ctypes, json, os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from pansvr_b200 import aln, shard
root, index_dir, header_sam, reads_fq, state_dir, out_path = sys.argv[1:7]
rank, local, world = shard.rank_env()
dist.init_process_group("gloo")
os.environ["PANSVR_ORACLE_SO"] = os.path.join(root, "oracle", "libksw_oracle.so")
lib = ctypes.CDLL(os.path.join(root, "tests", "emul", "libaln_emul.so"))
fq = open(reads_fq, "rb").read()
lines = fq.split(b"\n")
n_pairs = len([1 for x in lines if x]) // 8
pb, pe = shard.shard_range(n_pairs, rank, world)
mine = b"\n".join(lines[8 * pb:8 * pe]) + b"\n"
head = b"\n".join(lines[:4]) + b"\n"
ctx = aln.AlnContext(index_dir, header_sam, lib=lib, threads=2)
outs = []
for step in range(2):                                 # two passes: reset() must give the same bytes again
    ctx.reset()
    ctx.prime_read_stats(head)
    if rank > 0:
        ctx.await_state(os.path.join(state_dir, f"s{step}_r{rank}"))
    sam, ori = ctx.align_fastq(mine)
    if rank + 1 < world:
        ctx.publish_state(os.path.join(state_dir, f"s{step}_r{rank + 1}"))
    outs.append((sam, ori))
assert outs[0] == outs[1]
got = [None] * world
dist.all_gather_object(got, outs[0])
if rank == 0:
    hdr = ctx.header_text().encode()
    open(out_path + ".sam", "wb").write(hdr + b"".join(g[0] for g in got))
    open(out_path + "_ori.sam", "wb").write(hdr + b"".join(g[1] for g in got))
    print(json.dumps({"pairs": n_pairs, "ranges": [shard.shard_range(n_pairs, r, world) for r in range(world)]}))
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("name", ["multi_allele", "n_bases"])
def test_sharded_aln_two_ranks_equals_reference_t1(tmp_path, name):
    """one input cut into two contiguous pair ranges, one process each; the joined SAM equals `fc_aln -t 1` --
    with exact score ties (rand() draws at almost every pair) and with N bases (draws whose number depends on the stream)."""
    from tests.alntest_util import get_demo, need_ref_tools, read, first_diff
    need_ref_tools()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul"), os.path.join(ROOT, "tests", "emul", "fc_aln_emul")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "libksw_oracle.so")])
    demo = get_demo(name)
    w = tmp_path / "shard_worker.py"
    w.write_text(SHARD_WORKER)
    state = tmp_path / "state"
    state.mkdir()
    out = str(tmp_path / "joined")
    p = _torchrun([str(w), ROOT, demo.data.index_dir, demo.data.header_sam, demo.data.reads_fq, str(state), out], timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    assert first_diff(read(out + ".sam"), read(demo.ref_sam)) is None
    assert read(out + "_ori.sam") == read(demo.ref_ori)


PIECES_WORKER = r"""
# one input dealt to two processes piece by piece (piece b goes to rank b mod 2): every rank realigns its pieces in ONE
# pansvr_aln_pieces call; the random streams go from piece to piece through files, rank 0 joins the outputs in piece order
import ctypes, json, os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from pansvr_b200 import aln, shard
root, index_dir, header_sam, reads_fq, state_dir, out_path, piece_pairs = sys.argv[1:8]
piece_pairs = int(piece_pairs)
rank, local, world = shard.rank_env()
dist.init_process_group("gloo")
os.environ["PANSVR_ORACLE_SO"] = os.path.join(root, "oracle", "libksw_oracle.so")
lib = ctypes.CDLL(os.path.join(root, "tests", "emul", "libaln_emul.so"))
fq = open(reads_fq, "rb").read()
lines = fq.split(b"\n")
n_pairs = len([1 for x in lines if x]) // 8
cuts = list(range(0, n_pairs, piece_pairs)) + [n_pairs]
texts = [b"\n".join(lines[8 * b:8 * e]) + b"\n" for b, e in zip(cuts[:-1], cuts[1:])]
head = b"\n".join(lines[:4]) + b"\n"
ctx = aln.AlnContext(index_dir, header_sam, lib=lib, threads=2)
outs = []
for step in range(2):                                 # two passes: reset() must give the same bytes again
    ctx.reset()
    ctx.prime_read_stats(head)
    mine = [b for b in range(len(texts)) if b % world == rank]
    bufs = [ctypes.create_string_buffer(texts[b], len(texts[b])) for b in mine]
    pieces = [(ctypes.addressof(buf), len(texts[b]),
               os.path.join(state_dir, f"s{step}_b{b}") if b > 0 else None,
               os.path.join(state_dir, f"s{step}_b{b + 1}") if b + 1 < len(texts) else None) for b, buf in zip(mine, bufs)]
    if pieces:
        (sa, sn), (oa, on), per, release = ctx.align_pieces(pieces)
        sam, ori = ctypes.string_at(sa, sn), ctypes.string_at(oa, on)
        release()
        assert sum(p[0] for p in per) == sn and sum(p[1] for p in per) == on
        parts, so, oo = [], 0, 0
        for b, (ps, po) in zip(mine, per):
            parts.append((b, sam[so:so + ps], ori[oo:oo + po])); so += ps; oo += po
    else:
        parts = []
    outs.append(parts)
assert outs[0] == outs[1]
got = [None] * world
dist.all_gather_object(got, outs[0])
if rank == 0:
    allp = sorted(p for g in got for p in g)
    hdr = ctx.header_text().encode()
    open(out_path + ".sam", "wb").write(hdr + b"".join(p[1] for p in allp))
    open(out_path + "_ori.sam", "wb").write(hdr + b"".join(p[2] for p in allp))
    print(json.dumps({"pairs": n_pairs, "pieces": len(texts)}))
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("name,piece_pairs", [("multi_allele", 37), ("n_bases", 64)])
def test_pieces_dealt_to_two_ranks_equal_reference_t1(tmp_path, name, piece_pairs):
    """SURVEY.md 8e, block-cyclic: the input is dealt to two processes piece by piece, each realigns its pieces in one
    pansvr_aln_pieces call and the in-order passes follow each other from process to process through the state files; the SAM
    joined in piece order equals `fc_aln -t 1`."""
    from tests.alntest_util import get_demo, need_ref_tools, read, first_diff
    need_ref_tools()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul"), os.path.join(ROOT, "tests", "emul", "fc_aln_emul")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "libksw_oracle.so")])
    demo = get_demo(name)
    w = tmp_path / "pieces_worker.py"
    w.write_text(PIECES_WORKER)
    state = tmp_path / "state"
    state.mkdir()
    out = str(tmp_path / "joined")
    p = _torchrun([str(w), ROOT, demo.data.index_dir, demo.data.header_sam, demo.data.reads_fq, str(state), out, str(piece_pairs)], timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    assert json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])["pieces"] >= 4
    assert first_diff(read(out + ".sam"), read(demo.ref_sam)) is None
    assert read(out + "_ori.sam") == read(demo.ref_ori)
