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


def test_reference_arm_two_ranks_prints_once():
    """bench.py --impl reference under torchrun: rank 0 alone runs the CPU arm and prints one JSON line, rank 1 exits 0."""
    p = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-sample", "3000"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
