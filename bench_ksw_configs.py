#!/usr/bin/env python
"""bench_ksw_configs.py -- the ksw stage on the other kernel-level workloads of SURVEY.md section 8d / 8f
(bench.py stays the driver's single line on BASELINE.json configs[1]):

  config4_ext      250 bp reads vs 280 bp windows, w=500, never clipped (70 000 cells)
  config4_window   250 bp reads vs 1.5 kb windows, w=500 (156 375 cells)
  config4_global   end-to-end, qlen,tlen in [200,250] with 1-3 indels of 1-40 bp, w=500
  pipeline_like    the task shapes fc_aln really emits (short extensions + tiny end-to-end gaps), w=200
  fc_sv_contigs    contig vs anchor window of fc_sv, scoring 2/-10, gaps 24/2 + 32/1, w=zdrop=132

One JSON line per workload: reads/s and GCUPS through the C ABI with host buffers (H2D/D2H inside, "e2e"), the dominant
kernel's GCUPS from CUDA events ("kernel"), the fraction of the live-measured integer-ALU peak at 55 ops/cell, the same
tasks on the reference's own ksw_extd2_sse (oracle/_ref) on all host threads, and a parity check of a sample of exactly
the timed outputs against the oracle.

  python bench_ksw_configs.py [--scale 1.0] [--steps 3] [--out profiles/...jsonl]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="multiplies the task counts")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import torch
    from oracle import pyoracle
    from pansvr_b200 import ksw, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench_ksw_configs.py: no CUDA device (there is no CPU fallback)")
    ctx = ksw.KswContext(0)
    threads = min(48, os.cpu_count() or 1)
    int_peak = ctx.int_alu_peak_gops()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    work = [("config4_ext", lambda n: synth.config4_batch(n, "ext"), 400_000, 64),
            ("config4_window", lambda n: synth.config4_batch(n, "window"), 200_000, 64),
            ("config4_global", lambda n: synth.config4_batch(n, "global"), 200_000, 64),
            ("pipeline_like", lambda n: synth.pipeline_like_batch(n), 500_000, 64),
            ("fc_sv_contigs", lambda n: synth.fcsv_batch(n), 40_000, 96)]
    lines = []
    for name, make, n0, cap in work:
        if a.only and a.only != name:
            continue
        n = max(1000, int(n0 * a.scale))
        b = make(n)
        cells = int(synth.batch_cells(b))
        out_res = np.zeros((n, ksw.RES_WORDS), np.int32)
        out_cig = np.zeros((n, cap), np.uint32)
        for _ in range(a.warmup):
            ctx.extd2_batch(b, cigar_cap=cap, out=(out_res, out_cig))
        tot_ms = kern_ms = 0.0
        launches = 0
        for _ in range(a.steps):
            flush.zero_()
            torch.cuda.synchronize()
            ctx.extd2_batch(b, cigar_cap=cap, out=(out_res, out_cig))
            st = ctx.stats()
            tot_ms += st["total_ms"]; kern_ms += st["kernel_ms"]; launches += st["kernel_launches"]
        assert not (out_res[:, 11] & 1).any(), f"{name}: cigar_cap {cap} too small"
        # CPU reference on a bounded sample of the same tasks
        ns = min(n, max(2000, int(2.0e9 / max(1, cells // n))))
        sample = b.head(ns)
        t0 = time.time()
        r_ref, c_ref, _ = pyoracle.run(sample, "ref" if pyoracle.have_ref() else "oracle", threads=threads, cigar_cap=cap)
        cpu_s = time.time() - t0
        parity = bool(np.array_equal(r_ref[:, :11], out_res[:ns, :11]))
        ncig = r_ref[:, 9]
        mask = np.arange(cap)[None, :] < ncig[:, None]
        parity = parity and bool(((c_ref == out_cig[:ns]) | ~mask).all())
        gcups_kernel = cells * a.steps / (kern_ms * 1e-3) / 1e9
        line = {"workload": name, "tasks": n, "cells_per_task_mean": cells / n, "w": b.params.w, "zdrop": b.params.zdrop,
                "e2e": {"reads_per_s": n * a.steps / (tot_ms * 1e-3), "gcups": cells * a.steps / (tot_ms * 1e-3) / 1e9},
                "kernel": {"gcups": gcups_kernel, "launches_per_step": launches / a.steps,
                           "int_alu_frac": gcups_kernel * 55 / int_peak, "int_alu_peak_gops": int_peak},
                "cpu_reference": {"kind": "reference" if pyoracle.have_ref() else "port", "threads": threads, "tasks": ns,
                                  "reads_per_s": ns / cpu_s, "gcups": int(synth.batch_cells(sample)) / cpu_s / 1e9},
                "parity_sample_ok": parity, "steps": a.steps}
        print(json.dumps(line), flush=True)
        lines.append(line)
    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
