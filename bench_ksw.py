#!/usr/bin/env python
"""bench_ksw.py -- kernel-level benchmark of the ksw extension stage (round 1's bench.py; bench.py, the driver's benchmark,
measures the whole fc_aln stage and embeds this one's line under "ksw_config2").

One "step" = one pass of the ksw extension stage over one batch of synthetic tasks of
BASELINE.json configs[1] ("1 M x 150 bp signal reads vs ~1.1 kb anchor windows, band 100", SURVEY.md
section 8d "Config 2"): every read is one ksw_extd2 task (qlen 150, tlen 1100, w 100, zdrop 400,
flag 0, 2/-12, gaps min(16+k, 32)) = 25 100 in-band DP cells, with traceback and CIGAR.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--tasks T] [--impl reference]

value   : whole-job reads/s with the batch already resident in HBM (pansvr_ksw_extd2_batch_device),
          CUDA-event time of the call on its own stream (host planning gap + plan upload + kernels).
e2e     : the same metric through the host-buffer C-ABI call (pansvr_ksw_extd2_batch): pinned host
          buffers in, H2D + kernels + D2H of results and CIGARs inside the timed region.
N > 1   : one process per GPU (torchrun), every rank runs its own T tasks (weak scaling), no
          collective on the data path; time = max over ranks.
--impl reference : the reference's own ksw2_extd2_sse.c (oracle/_ref, else the oracle port) on all
          host cores, one bounded sample of the same workload per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CELLS_PER_TASK = 25100          # band_cells(150, 1100, 100), SURVEY.md 8d
OPS_PER_CELL = 55               # int ops per DP cell with traceback
CIGAR_CAP = 16


NCU_DRAM_BYTES_PER_TASK = 56400       # profiles/r1h_ksw_team_full.md (r1o addendum: 20.68 GB read + 35.72 GB written per 1 M tasks)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        sm, smax, reasons = [], [], set()
        with open(self.f.name) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); smax.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(batch, sample_tasks, threads):
    """reads/s of the reference's CPU ksw on `threads` host threads over a bounded sample."""
    from oracle import pyoracle
    pyoracle.build()
    kind, impl = ("reference", "ref") if pyoracle.have_ref() else ("port", "oracle")
    sub = batch.head(sample_tasks)
    pyoracle.run(sub.head(min(2000, sub.n)), impl, threads=threads, cigar_cap=CIGAR_CAP)      # page in / warm
    _, _, secs = pyoracle.run(sub, impl, threads=threads, cigar_cap=CIGAR_CAP)
    return sub.n / secs, kind, secs, sub.n


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from pansvr_b200 import synth
    threads = min(48, os.cpu_count() or 1)          # the reference caps -t at 48 (read_realignment.hpp:121)
    sample = args.ref_sample
    batch = synth.config2_batch(sample, seed=11)
    from oracle import pyoracle
    pyoracle.build()
    kind, impl = ("reference", "ref") if pyoracle.have_ref() else ("port", "oracle")
    for _ in range(args.warmup):
        pyoracle.run(batch.head(max(1000, sample // 10)), impl, threads=threads, cigar_cap=CIGAR_CAP)
    secs = []
    for _ in range(args.steps):
        _, _, s = pyoracle.run(batch, impl, threads=threads, cigar_cap=CIGAR_CAP)
        secs.append(s)
    tot = sum(secs)
    rate = sample * args.steps / tot
    line = {
        "impl": "reference", "metric": "ksw tasks/s (ksw extension stage, config 2: one task per read)", "value": rate, "unit": "tasks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 differences / int32 H (SSE2)",
        "data": "synthetic", "gcups": rate * CELLS_PER_TASK / 1e9,
        "config": {"workload": "config2: 150 bp reads vs 1.1 kb anchor windows, w=100, zdrop=400, flag=0, with CIGAR",
                   "tasks_per_step": sample, "cells_per_task": CELLS_PER_TASK},
        "cpu_baseline": {"value": rate, "unit": "tasks/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} tasks of the same workload per step, ksw_extd2_sse from "
                                   f"{'oracle/_ref/libksw_ref.so (reference source, -O3, SSE2 as shipped)' if kind == 'reference' else 'oracle port'}"
                                   f" on a pthread pool, one ksw_extz_t per thread"},
        "e2e": {"value": rate, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tasks", type=int, default=1_000_000, help="tasks (reads) per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=200_000, help="tasks per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=600_000, help="tasks of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args(argv)


def main():
    args = parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_ksw.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(args, rank, world, local, dist)
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def measure(args, rank, world, local, dist):
    """The config-2 measurement on an initialised process (group); returns the JSON line as a dict on rank 0, None elsewhere."""
    import torch
    from pansvr_b200 import ksw, shard, synth
    dev = torch.device("cuda", local)
    ctx = ksw.KswContext(local)

    # ---- synthetic workload (every rank its own seed: weak scaling, reads shard trivially)
    n = args.tasks
    batch = synth.config2_batch(n, seed=shard.shard_seed(11, rank))
    p = batch.params
    cells = CELLS_PER_TASK * n
    # pinned host staging (the batcher's buffers) and the device-resident copy
    pins = {}
    for k, dt in (("qseq", np.uint8), ("tseq", np.uint8), ("qoff", np.int64), ("toff", np.int64), ("qlen", np.int32), ("tlen", np.int32)):
        a = np.ascontiguousarray(getattr(batch, k), dt)
        pa = ksw.PinnedArray(a.shape, dt)
        pa.array[...] = a
        pins[k] = pa
    hb = synth.KswBatch(pins["qseq"].array, pins["qoff"].array, pins["qlen"].array, pins["tseq"].array, pins["toff"].array,
                        pins["tlen"].array, p, batch.name)
    out_res = ksw.PinnedArray((n, ksw.RES_WORDS), np.int32)
    out_cig = ksw.PinnedArray((n, CIGAR_CAP), np.uint32)
    d = {k: torch.from_numpy(pins[k].array).to(dev) for k in pins}
    d_res = torch.zeros((n, ksw.RES_WORDS), dtype=torch.int32, device=dev)
    d_cig = torch.zeros((n, CIGAR_CAP), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    torch.cuda.synchronize()

    def step_resident():
        ctx.extd2_batch_device(n, d["qseq"].data_ptr(), d["qoff"].data_ptr(), d["qlen"].data_ptr(), d["tseq"].data_ptr(),
                               d["toff"].data_ptr(), d["tlen"].data_ptr(), hb.qlen, hb.tlen, p, d_res.data_ptr(), d_cig.data_ptr(),
                               CIGAR_CAP)
        return ctx.stats()

    def step_e2e():
        ctx.extd2_batch(hb, cigar_cap=CIGAR_CAP, out=(out_res.array, out_cig.array))
        return ctx.stats()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, k_steps):
        tot_ms, kern_ms, launches, last = 0.0, 0.0, 0, None
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_steps):
            flush.zero_()                      # L2 flush between timed iterations (outside the event brackets)
            torch.cuda.synchronize()
            last = step()
            tot_ms += last["total_ms"]; kern_ms += last["kernel_ms"]; launches += last["kernel_launches"]
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        tot_ms, kern_ms, wall_ms = shard.max_over_ranks([tot_ms, kern_ms, wall_ms], dist, dev)
        return tot_ms, kern_ms, launches, wall_ms, last

    for _ in range(args.warmup):
        step_resident()
    for _ in range(min(args.warmup, 2)):
        step_e2e()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tot_ms, kern_ms, launches, wall_ms, last = timed(step_resident, args.steps)
    e_tot_ms, e_kern_ms, e_launches, e_wall_ms, e_last = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # parity spot check of what was just timed (result of the e2e pass, sample vs the oracle) -- rank 0
    parity = None
    if rank == 0:
        from oracle import pyoracle
        idx = np.random.default_rng(1).choice(n, min(n, 2000), replace=False)
        r0, c0, _ = pyoracle.run(hb.take(idx), "oracle", threads=min(16, os.cpu_count() or 1), cigar_cap=CIGAR_CAP)
        parity = bool(np.array_equal(r0[:, :11], out_res.array[idx][:, :11]) and np.array_equal(c0, out_cig.array[idx])
                      and np.array_equal(d_res.cpu().numpy()[idx][:, :11], r0[:, :11]))

    pipes = ctx.int_pipe_peaks_gops() if rank == 0 else {}
    int_peak = pipes.get("mixed", 0.0)
    ctx.close()
    for pa in list(pins.values()) + [out_res, out_cig]:
        pa.free()
    del d, d_res, d_cig, flush
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peaks, peak_src = measured_peaks()
    reads_s = shard.whole_job_rate(n, args.steps, world, tot_ms)
    e2e_reads_s = shard.whole_job_rate(n, args.steps, world, e_tot_ms)
    kernel_gcups = cells * args.steps / (kern_ms * 1e-3) / 1e9                  # per GPU, dominant kernel only
    achieved_gops = kernel_gcups * OPS_PER_CELL
    # HBM view of the same kernel: bytes it must move per task (query + touched target + traceback written
    # + results), SURVEY 8d: 1 B per computed cell of traceback dominates
    tb_bytes = 0
    for r in range(150 + 1100 - 1):
        lo = max(0, r - 149, (r - 100 + 1) >> 1); hi = min(1099, r, (r + 100) >> 1)
        if lo > hi:
            break
        tb_bytes += (hi | 15) - (lo & ~15) + 1
    bytes_per_task = 150 + 272 + tb_bytes + 48 + 4 * CIGAR_CAP
    hbm_gbs = bytes_per_task * n * args.steps / (kern_ms * 1e-3) / 1e9

    line = {
        "metric": "ksw tasks/s (ksw extension stage, config 2: one task per read)",
        "value": reads_s, "unit": "tasks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16x2 (exact int8 emulation) / int32 H", "data": "synthetic",
        "gcups": reads_s * CELLS_PER_TASK / 1e9, "kernel_gcups_per_gpu": kernel_gcups,
        "config": {"workload": "config2: 1 M x 150 bp reads vs 1.1 kb anchor windows, w=100, zdrop=400, flag=0, with traceback/CIGAR "
                               "(BASELINE.json configs[1])",
                   "tasks_per_gpu_per_step": n, "cells_per_task": CELLS_PER_TASK, "cigar_cap": CIGAR_CAP,
                   "l2": "explicit 256 MB L2 flush between timed iterations; inputs 167 MB + traceback scratch also exceed L2",
                   "parallelism": f"reads sharded over {world} GPU(s), no collective on the data path"},
        "e2e": {"value": e2e_reads_s, "unit": "tasks/s", "h2d_bytes_per_step": int(e_last["h2d_bytes"]) * world,
                "d2h_bytes_per_step": int(e_last["d2h_bytes"]) * world, "ms_per_step": e_tot_ms / args.steps,
                "gcups": e2e_reads_s * CELLS_PER_TASK / 1e9},
        "gpu_launches": int(launches + e_launches),
        "roofline": {"bound": "int_alu", "achieved": achieved_gops, "peak": int_peak, "unit": "Gop/s",
                     "frac": achieved_gops / int_peak if int_peak else None,
                     # DRAM bytes per launch of this kernel from ncu (profiles/r1h_ksw_team_full.md, r1o addendum:
                     # 56.40 GB at 1 M tasks = dram__bytes_read 20.68 GB + dram__bytes_write 35.72 GB), scaled to this launch
                     "traffic": NCU_DRAM_BYTES_PER_TASK * n, "traffic_unit": "bytes/launch",
                     "algorithmic_bytes": bytes_per_task * n,
                     "kernel": "ksw_team_kernel<8,true>", "kernel_ms_per_launch": kern_ms / args.steps,
                     "ops_per_cell": OPS_PER_CELL, "gcups": kernel_gcups,
                     "peak_source": "pansvr_int_alu_peak measured live on this GPU (IADD3/LOP3/VIMNMX chains)",
                     "pipe_peaks_gops": pipes,
                     "frac_of": {k: (achieved_gops / v if v else None) for k, v in pipes.items()},
                     "traffic_source": "profiles/r1h_ksw_team_full.md (one ncu --set full capture, r1o addendum), scaled to this launch; not measured in this run",
                     "hbm": {"achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                             "frac": hbm_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                             "bytes_per_task": bytes_per_task, "peak_source": peak_src}},
        "clocks": clocks,
        "wall_ms_per_step": {"resident": wall_ms / args.steps, "e2e": e_wall_ms / args.steps},
        "parity_sample_ok": parity,
        "resident_warps": int(last["resident_warps"]), "tb_bytes_per_warp": int(last["tb_bytes_per_warp"]),
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = min(48, os.cpu_count() or 1)
        rate, kind, secs, ns = cpu_reference_rate(batch, args.cpu_sample, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "tasks/s", "cores": threads, "kind": kind,
                                "sample": f"first {ns} tasks of the same workload, {secs:.1f} s wall on {threads} threads",
                                "gcups": rate * CELLS_PER_TASK / 1e9}
    return line


if __name__ == "__main__":
    main()
