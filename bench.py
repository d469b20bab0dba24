#!/usr/bin/env python
"""bench.py -- the hot-path benchmark of pansvr_b200: the whole `fc_aln` stage (BASELINE.json metric: realigned reads/s and
GCUPS at 1/2/4/8 B200 vs the CPU `panSVR aln`), on SURVEY.md section 8d "Config 3" = BASELINE.json configs[2]:
10 M x 150 bp signal reads vs 10 500 anchors with 50 bp - 10 kb INS/DEL alleles, 2-4 alleles per locus (multi-candidate reads,
exact score ties), FASTQ text in -> SAM text out through the C ABI (pansvr_aln_block, include/pansvr_b200.h).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pairs P]

step    : one pass of the stage over the whole input (all ranks together).  ctx.reset() between steps puts the rand() replay
          back to the state of a freshly started fc_aln, so every step produces the same bytes.
value   : whole-job realigned reads/s with the FASTQ text of the rank's shard already staged in pinned host memory
          (what step 0 of the reference's pipeline hands to step 1) -- every stage of the path runs, host and device.
e2e     : the same through the plain host-buffer call: pageable FASTQ text in, malloc'ed SAM text out; h2d/d2h bytes are what
          the library copied (counted where the copies are issued).
N > 1   : one process per GPU (torchrun); ONE input, dealt to the ranks piece by piece (piece b of --piece-pairs pairs goes to
          rank b mod N: strong scaling); no collective on the data path.  Every rank realigns its pieces in one
          pansvr_aln_pieces call.  The only state that flows between pairs is the reference's process-wide rand() stream: the
          in-order pass of piece b takes the stream state from piece b-1's through a 400-byte file in /dev/shm and hands it on the
          moment it is done; all other stages of all pieces overlap.  time = max over ranks.
roofline: the DP kernel (ksw_team, integer ALU: cells x 55 ops / kernel time / measured integer peak) as it runs inside the
          stage, and "roofline_seed": the seeding kernels against HBM (SURVEY.md section 8d byte model / kernel time / hbm_gbs).
ksw_config2: the kernel-level line of BASELINE.json configs[1] (1 M x 150 bp vs 1.1 kb windows, w=100), from bench_ksw.py.
--impl reference : the reference's own `panSVR fc_aln -t <cores> -S` (oracle/_ref, built from the unmodified reference) on a
          bounded prefix of the same input per step, start-up (`-R 1`: the 2 GiB index read) subtracted.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# eight sub-blocks in flight with a handful of streams each: give them hardware queues of their own (read when CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import bench_ksw  # noqa: E402
from bench_ksw import ClockSampler, env_int, measured_peaks  # noqa: E402

OPS_PER_CELL = 55               # int ops per DP cell with traceback
SEED_BYTES_MISS, SEED_BYTES_HIT = 64, 290      # SURVEY.md 8d seeding byte model: per miss probe / per hit
KSW_TRAFFIC_BYTES = 384_961_280   # profiles/r2_ncu_stage_kernels.md (r2f): DRAM bytes of one ksw_team_kernel<4,0,1> launch, 262144-pair sub-block
METRIC = "realigned reads/s (fc_aln stage: FASTQ text in, SAM text out; config 3)"


def workload_name(d):
    return (f"config3: {2 * d.n_pairs} x {d.read_len} bp signal reads vs {d.n_anchors} anchors, 50 bp-10 kb INS/DEL alleles, 2-4 alleles "
            "per locus sharing flanks (BASELINE.json configs[2], SURVEY.md 8d)")


def ref_threads():
    return min(48, os.cpu_count() or 1)          # the reference caps -t at 48 (read_realignment.hpp:121)


def reference_startup(d, cache=True):
    """Wall seconds of `fc_aln -R 1` (index load etc.), min of 2."""
    from benchdata import config3
    return min(config3.run_reference_aln(d, 1, max_pairs=1) for _ in range(2))


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from benchdata import config3
    d = config3.prepare(pairs=args.pairs, loci=args.loci)
    threads = ref_threads()
    sample = min(args.ref_sample_pairs, d.n_pairs)
    t_start = reference_startup(d)
    for _ in range(min(args.warmup, 2)):
        config3.run_reference_aln(d, threads, max_pairs=min(sample, 50_000))
    secs = []
    for _ in range(args.steps):
        secs.append(max(config3.run_reference_aln(d, threads, max_pairs=sample) - t_start, 1e-6))
    tot = sum(secs)
    rate = 2 * sample * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int8 differences / int32 H (SSE2 ksw), host C++", "data": "synthetic",
        "config": {"workload": workload_name(d), "reads_per_step": 2 * sample,
                   "sample": f"first {sample} pairs of the same input per step"},
        "cpu_baseline": {"value": rate, "unit": "reads/s", "cores": threads, "kind": "reference",
                         "sample": f"`panSVR fc_aln -t {threads} -S -R {sample}` (oracle/_ref: the unmodified reference) on the first {sample} pairs "
                                   f"of the same input, wall minus the {t_start:.2f} s start-up measured with -R 1; SAM written to /dev/null"},
        "e2e": {"value": rate, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "startup_seconds": t_start,
    }
    print(json.dumps(line), flush=True)


def reference_t1_md5(d, pairs):
    """md5 of the SAM / -p files of `fc_aln -t 1 -S -R pairs` (the deterministic mode = the output oracle), cached per box."""
    from benchdata import config3
    cache = os.path.join(d.workdir, f"ref_t1_{pairs}.json")
    if os.path.exists(cache):
        with open(cache) as f:
            return json.load(f)
    out, ori = os.path.join(d.workdir, f"ref_t1_{pairs}.sam"), os.path.join(d.workdir, f"ref_t1_{pairs}_ori.sam")
    t0 = time.time()
    config3.run_reference_aln(d, 1, max_pairs=pairs, out_sam=out, ori_sam=ori)
    secs = time.time() - t0
    res = {"seconds": secs}
    for k, p in (("sam", out), ("ori", ori)):
        h = hashlib.md5()
        with open(p, "rb") as f:
            for chunk in iter(lambda: f.read(64 << 20), b""):
                h.update(chunk)
        res[k] = h.hexdigest()
        res[k + "_bytes"] = os.path.getsize(p)
        os.unlink(p)
    with open(cache, "w") as f:
        json.dump(res, f)
    return res


def md5_of(header: bytes, addr: int, n: int) -> str:
    h = hashlib.md5(header)
    if n:
        h.update((C.c_char * n).from_address(addr))
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=5_000_000, help="read pairs of the input (config 3: 5 M pairs = 10 M reads)")
    ap.add_argument("--loci", type=int, default=5250, help="SV loci of the anchor set (config 3: 5250 loci = 10 500 anchors)")
    ap.add_argument("--block-pairs", type=int, default=0, help="pairs per pansvr_aln_block call at N=1 (multiple of 4096; 0 = the whole input in one call)")
    ap.add_argument("--piece-pairs", type=int, default=0, help="N>1: pairs per piece of the block-cyclic deal (multiple of 4096; 0 = the largest of "
                    "262144 ... 65536 that gives every rank at least four pieces and keeps the ranks within 8 %% of each other)")
    ap.add_argument("--ref-sample-pairs", type=int, default=250_000, help="pairs per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--parity-pairs", type=int, default=491_520, help="prefix checked against `fc_aln -t 1` in the run (0 = off)")
    ap.add_argument("--threads", type=int, default=0, help="host helper threads per rank (0 = cores / ranks)")
    ap.add_argument("--no-ksw", action="store_true", help="skip the config-2 kernel line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    from benchdata import config3
    from pansvr_b200 import aln, ksw, shard
    # PANSVR_BENCH_EMUL=1 (tests only): dry run of this script's flow on the host-emulated pipeline of tests/emul, gloo instead of
    # nccl; the line it prints is marked "emulated" and its numbers mean nothing.  The product path needs a B200.
    emul = bool(os.environ.get("PANSVR_BENCH_EMUL"))
    emul_lib = None
    if emul:
        os.environ.setdefault("PANSVR_ORACLE_SO", os.path.join(ROOT, "oracle", "libksw_oracle.so"))
        emul_lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "libaln_emul.so"))
        args.no_ksw = args.no_cpu_baseline = True
        dev = torch.device("cpu")
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if emul:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=dev)

    d = config3.prepare(pairs=args.pairs, loci=args.loci)
    # ---- this rank's pieces of the input, cut at multiples of IDX_STRIDE pairs (results are merged by piece index = input order)
    S = config3.IDX_STRIDE
    if world == 1:
        bp = max(S, args.block_pairs // S * S) if args.block_pairs > 0 else d.n_pairs
        cuts = list(range(0, d.n_pairs, bp)) + [d.n_pairs]
        mine = list(range(len(cuts) - 1))
    else:
        pp = max(S, args.piece_pairs // S * S) if args.piece_pairs > 0 else 65536
        if args.piece_pairs <= 0:
            for cand in (262144, 196608, 131072, 98304, 65536):      # larger pieces = fuller kernels; smaller = better balance
                n_pc = -(-d.n_pairs // cand)
                if n_pc >= 4 * world and -(-n_pc // world) * cand <= 1.08 * d.n_pairs / world + cand * 0.0:
                    pp = cand
                    break
        cuts = list(range(0, d.n_pairs, pp)) + [d.n_pairs]
        mine = [b for b in range(len(cuts) - 1) if b % world == rank]
    n_pieces_total = len(cuts) - 1
    spans = [d.byte_range(cuts[b], cuts[b + 1]) for b in mine]
    nbytes = sum(e - b for b, e in spans)
    my_pairs = sum(cuts[b + 1] - cuts[b] for b in mine)
    pin = ksw.PinnedArray((max(nbytes, 1),), np.uint8) if not emul else argparse.Namespace(array=np.empty(max(nbytes, 1), np.uint8))
    offs = []
    with open(d.reads_fq, "rb") as f:
        at = 0
        for b0, b1 in spans:
            f.seek(b0)
            got = f.readinto(memoryview(pin.array)[at:at + (b1 - b0)])
            assert got == b1 - b0
            offs.append((at, b1 - b0))
            at += b1 - b0
    pageable = bytes(memoryview(pin.array)[:nbytes])                     # the e2e leg's input: ordinary host memory
    with open(d.reads_fq, "rb") as f:
        head = f.read(4096)                                              # first record of the input (STAT_ fields, RR:134-148)
    parity_pairs = min(args.parity_pairs // S * S, cuts[1]) if args.parity_pairs > 0 else 0
    threads = args.threads or max(1, min(48, (os.cpu_count() or 1) // world))
    ctx = aln.AlnContext(d.index_dir, d.header_sam, device=local, threads=threads, lib=emul_lib)
    ctx.prime_read_stats(head)
    header = ctx.header_text().encode()
    run_id = [f"{os.getpid()}_{int(time.time())}"]
    if dist is not None:
        dist.broadcast_object_list(run_id, src=0)      # stream-state files of this run only
    run_id = run_id[0]
    state_dir = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else d.workdir, f"pansvr_state_{run_id}")
    os.makedirs(state_dir, exist_ok=True)
    step_no = [0]

    def barrier():
        if dist is not None:
            dist.barrier()
        if not emul:
            torch.cuda.synchronize()

    def one_step(kind, prefix_pairs=0):
        """One pass over this rank's pieces (prefix_pairs: only the first pairs of the input, on rank 0 -- the parity check).
        Returns (stats, (md5 sam, md5 ori, bytes, bytes) of the output when prefix_pairs)."""
        k = step_no[0]; step_no[0] += 1
        ctx.reset()
        ctx.prime_read_stats(head)
        base = pin.array.ctypes.data if kind == "resident" else C.cast(C.c_char_p(pageable), C.c_void_p).value
        first = None
        if prefix_pairs:
            if rank == 0:
                o0, o1 = d.byte_range(0, prefix_pairs)
                sam, ori, release = ctx.align_ptr(base, o1 - o0)
                first = (md5_of(header, *sam), md5_of(header, *ori), sam[1], ori[1])
                release()
            return ctx.stats(), first
        if world == 1:
            for off, n in offs:
                sam, ori, release = ctx.align_ptr(base + off, n)
                release()
        elif mine:
            pieces = [(base + off, n, os.path.join(state_dir, f"s{k}_b{b}") if b > 0 else None,
                       os.path.join(state_dir, f"s{k}_b{b + 1}") if b + 1 < n_pieces_total else None) for b, (off, n) in zip(mine, offs)]
            sam, ori, _, release = ctx.align_pieces(pieces)
            release()
        return ctx.stats(), first

    def timed(kind, k_steps):
        acc = None
        barrier()
        if not emul:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        t0 = time.perf_counter()
        for _ in range(k_steps):
            st, _ = one_step(kind)
            if acc is None:
                acc = {k: (list(v) if isinstance(v, list) else v) for k, v in st.items()}
            else:
                for k, v in st.items():
                    acc[k] = [a + b for a, b in zip(acc[k], v)] if isinstance(v, list) else acc[k] + v
        if not emul:
            ev1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ev_ms = ev0.elapsed_time(ev1) if not emul else wall_ms
        wall_ms, ev_ms = shard.max_over_ranks([wall_ms, ev_ms], dist, dev)
        return wall_ms, ev_ms, acc

    for _ in range(args.warmup):
        one_step("resident")
    for _ in range(min(args.warmup, 1)):
        one_step("e2e")
    sampler = ClockSampler(local)
    if rank == 0 and not emul:
        sampler.start()
    wall_ms, ev_ms, st = timed("resident", args.steps)
    e_wall_ms, e_ev_ms, e_st = timed("e2e", args.steps)
    clocks = sampler.stop() if rank == 0 and not emul else None

    # ---- parity in the run: the first call's output against `panSVR fc_aln -t 1 -S -R <pairs>` (md5 of both files)
    parity = None
    if parity_pairs > 0:
        _, first = one_step("resident", prefix_pairs=parity_pairs)
        if rank == 0 and os.access(os.path.join(config3.REF_BIN, "panSVR"), os.X_OK):
            pp = parity_pairs
            ref = reference_t1_md5(d, pp)
            parity = {"pairs": pp, "sam_md5_equal": first[0] == ref["sam"], "ori_md5_equal": first[1] == ref["ori"],
                      "sam_bytes": first[2], "ori_bytes": first[3], "reference": f"oracle/_ref/panSVR fc_aln -t 1 -S -R {pp}",
                      "reference_t1_seconds": ref.get("seconds")}
    barrier()
    # ---- one more pass with a single sub-block in flight: every kernel then runs alone on the device, so its CUDA-event time is its
    # own duration (in the timed passes the kernels of three sub-blocks overlap on different streams and stretch each other)
    os.environ["PANSVR_FLIGHT"] = "1"
    pst, _ = one_step("resident")
    del os.environ["PANSVR_FLIGHT"]
    barrier()

    # ---- gather the per-rank device counters (sums) on rank 0
    keys = ("reads", "mems", "ksw_tasks", "ksw_cells", "kernel_launches", "h2d_bytes", "d2h_bytes", "seed_probes",
            "in_order_seconds", "in_order_pairs", "in_order_draws", "host_pairs", "tie_pairs")
    vec = [float(st[k]) for k in keys] + [float(e_st[k]) for k in keys]
    mx = [st["seed_kernel_ms"], st["ksw_kernel_ms"], st["stage_kernel_ms"]] + list(st["stage_seconds"])
    pmx = [pst["seed_kernel_ms"], pst["ksw_kernel_ms"], pst["stage_kernel_ms"], float(pst["ksw_cells"]), float(pst["seed_probes"]), float(pst["mems"])] + list(pst["stage_kernel_ms_by"])
    if dist is not None:
        t = torch.tensor(vec, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        vec = [float(x) for x in t.cpu()]
        mx = shard.max_over_ranks(mx, dist, dev)
        pmx = shard.max_over_ranks(pmx, dist, dev)
    tot = dict(zip(keys, vec[:len(keys)]))
    e_tot = dict(zip(keys, vec[len(keys):]))
    seed_ms, ksw_ms, stage_ms = mx[0], mx[1], mx[2]
    stage_seconds = mx[3:]

    ksw_line = None
    if not args.no_ksw:
        ctx.close()
        ka = bench_ksw.parse_args(["--steps", "3", "--warmup", "3", "--no-cpu-baseline"])
        ksw_line = bench_ksw.measure(ka, rank, world, local, dist)
        ctx = None
    pipes = {}
    if rank == 0 and ksw_line is not None:
        pipes = ksw_line["roofline"].get("pipe_peaks_gops", {})
    elif rank == 0 and not emul:
        kc = ksw.KswContext(local)
        pipes = kc.int_pipe_peaks_gops()
        kc.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    n_reads = 2 * d.n_pairs
    reads_s = n_reads * args.steps / (ev_ms * 1e-3)
    e_reads_s = n_reads * args.steps / (e_ev_ms * 1e-3)
    cells_s = tot["ksw_cells"] / (ev_ms * 1e-3)
    # DP kernel as it runs inside the stage, from the single-flight pass (slowest rank): cells of the emitted tasks / CUDA-event time
    # of the ksw kernels alone on the device
    int_peak = pipes.get("mixed", 0.0)
    p_seed_ms, p_ksw_ms, p_stage_ms, p_cells, p_probes, p_hits = pmx[:6]
    ksw_gcups = p_cells / (p_ksw_ms * 1e-3) / 1e9 if p_ksw_ms > 0 else 0.0
    ksw_gops = ksw_gcups * OPS_PER_CELL
    # seeding: SURVEY 8d byte model (every probe made once: the counting pass keeps the MEMs it finds), plus the copy of each MEM
    # (32 B read + 32 B written) to its place in the dense list
    hits = tot["mems"]
    seed_bytes = (p_probes - min(p_hits, p_probes)) * SEED_BYTES_MISS + p_hits * (SEED_BYTES_HIT + 64)
    seed_gbs = seed_bytes / (p_seed_ms * 1e-3) / 1e9 if p_seed_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": reads_s, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "u16x2 (exact int8 emulation) / int32 H in the DP kernel; uint64 2-bit words in seeding", "data": "synthetic",
        "gcups": cells_s / 1e9,
        "config": {"workload": workload_name(d), "reads_per_step": n_reads, "pairs_per_rank": my_pairs, "calls_per_step_per_rank": len(offs) if world == 1 else 1, "pieces": n_pieces_total,
                   "host_threads_per_rank": threads, "host_cores": os.cpu_count(),
                   "l2": "inputs exceed L2: every call streams hundreds of MB of packed reads, MEMs, ksw tasks and traceback (no reuse between steps)",
                   "parallelism": (f"one input dealt to {world} processes (one per GPU) in {n_pieces_total} pieces of {cuts[1] - cuts[0]} pairs, round robin; no collective "
                                   "on the data path; the rand() stream state goes from piece to piece through files in /dev/shm") if world > 1 else "1 GPU",
                   "timing": "CUDA events on the rank's current stream around the K steps after a barrier + synchronize, max over ranks (wall clock agrees: see wall_ms_per_step)"},
        "e2e": {"value": e_reads_s, "unit": "reads/s", "h2d_bytes_per_step": int(e_tot["h2d_bytes"] / args.steps),
                "d2h_bytes_per_step": int(e_tot["d2h_bytes"] / args.steps), "ms_per_step": e_ev_ms / args.steps,
                "host_in_bytes_per_step": int(d.pair_offsets[-1]), "note": "pageable FASTQ text in, malloc'ed SAM text out through pansvr_aln_block"},
        "gpu_launches": int(tot["kernel_launches"] + e_tot["kernel_launches"]),
        "roofline": {"bound": "int_alu", "kernel": "ksw_team_kernel (stage E, all variants of the fc_aln task mix, w=200)",
                     "achieved": ksw_gops, "peak": int_peak, "unit": "Gop/s", "frac": ksw_gops / int_peak if int_peak else None,
                     "gcups": ksw_gcups, "ops_per_cell": OPS_PER_CELL, "kernel_ms_per_step": p_ksw_ms,
                     "timing": "CUDA events around the kernels in one extra pass with a single sub-block in flight (kernels alone on the device)",
                     "cells_per_step": tot["ksw_cells"] / args.steps, "tasks_per_step": tot["ksw_tasks"] / args.steps,
                     "pipe_peaks_gops": pipes, "frac_of": {k: (ksw_gops / v if v else None) for k, v in pipes.items()},
                     "peak_source": "pansvr_int_pipe_peaks measured live on this GPU ('mixed' = IADD3/LOP3/VIMNMX chains; ALU pipe, FMA pipe and both together alongside)",
                     "traffic": KSW_TRAFFIC_BYTES,
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the largest variant, ksw_team_kernel<4,0,1>, on a 262144-pair "
                                     "sub-block (1.34 ms; 90 MB read, 295 MB written = the traceback bytes), from the ncu --set full capture profiles/r2_ncu_stage_kernels.md "
                                     "(r2f) -- not measured in this run; the kernel is bound by the integer pipes, HBM sees 0.29 TB/s"},
        "roofline_seed": {"bound": "hbm", "kernel": "for_each_kernel<FnSeed> + for_each_kernel<FnSeedPlace> (stage B)", "achieved": seed_gbs,
                          "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": seed_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                          "kernel_ms_per_step": p_seed_ms, "probes_per_step": tot["seed_probes"] / args.steps,
                          "hits_per_step": hits / args.steps, "bytes_model": f"{SEED_BYTES_MISS} B per miss probe, {SEED_BYTES_HIT} B per hit (SURVEY.md 8d), 64 B per MEM moved to the dense list",
                          "peak_source": peak_src, "traffic": None},
        "clocks": clocks,
        "wall_ms_per_step": {"resident": wall_ms / args.steps, "e2e": e_wall_ms / args.steps},
        "stage_seconds_per_step": {k: v / args.steps for k, v in zip(("A_encode_census", "B_seeding_gpu", "C_merge_chain", "D_plan", "E_ksw_gpu",
                                                                       "F_replay_text", "fastq_parse", "output_join"), stage_seconds)},
        "device_busy_ms_per_step": {"seed_kernels": seed_ms / args.steps, "ksw_kernels": ksw_ms / args.steps, "stage_kernels": stage_ms / args.steps,
                                    "note": "summed CUDA-event times of the timed passes; kernels of the sub-blocks in flight overlap, so these exceed the busy time"},
        "kernel_ms_alone_per_step": {"seeding": p_seed_ms, "ksw": p_ksw_ms,
                                     **dict(zip(("records_ori_encode", "seeding_", "merge_chain", "ksw_plan", "resolve", "cell_count", "pair_probe_finalize", "sam_text"), pmx[6:14]))},
        "in_order_chain": {"seconds_per_step": tot["in_order_seconds"] / args.steps, "pairs_with_draws_per_step": tot["in_order_pairs"] / args.steps,
                           "draws_per_step": tot["in_order_draws"] / args.steps, "host_path_pairs_per_step": tot["host_pairs"] / args.steps,
                           "pairs_finished_in_order_from_device_results_per_step": tot["tie_pairs"] / args.steps,
                           "note": "the in-order passes over the reference's rand() stream, summed over all ranks: the one part of the stage that is sequential "
                                   "from pair to pair (one piece after the other, across ranks); everything else overlaps -- the lower bound of a step at any N"},
        "parity": parity,
    }
    if emul:
        line["emulated"] = "PANSVR_BENCH_EMUL dry run on the CPU emulation of tests/emul: NOT a measurement"
    if ksw_line is not None:
        line["ksw_config2"] = ksw_line
    if world == 1 and not args.no_cpu_baseline:
        thr = ref_threads()
        sample = min(args.ref_sample_pairs, d.n_pairs)
        t_start = reference_startup(d)
        t = config3.run_reference_aln(d, thr, max_pairs=sample) - t_start
        line["cpu_baseline"] = {"value": 2 * sample / max(t, 1e-6), "unit": "reads/s", "cores": thr, "kind": "reference",
                                "sample": f"`panSVR fc_aln -t {thr} -S -R {sample}` (oracle/_ref) on the first {sample} pairs of the same input: "
                                          f"{t:.2f} s after subtracting the {t_start:.2f} s start-up (-R 1)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    import shutil
    shutil.rmtree(state_dir, ignore_errors=True)                          # (the hand-over files are consumed; the directory is this run's)


if __name__ == "__main__":
    main()
