#!/usr/bin/env python
"""bench_pipeline.py -- pipeline-level measurement of the aln stage ("Config 3"-like data):
realigned reads/s of `pansvr_b200` (CUDA seeding + CUDA ksw + host replay, through the C ABI) next to the reference's own
`panSVR fc_aln -t N -S` on the same box, same inputs, in the same run, with the SAM of both compared byte for byte.

Not the driver's bench (that is bench.py, the ksw stage on BASELINE configs[1]); this one documents the whole path.

  python bench_pipeline.py [--loci 2000] [--alleles 2] [--pairs-per-sv 100] [--threads N] [--out profiles/...json]

Index load (a fixed 2 GiB table read, 2-3 s) is excluded on both sides: the reference's start-up is measured with `-R 1`
(SURVEY.md section 8d), ours is the pansvr_aln_create call.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=2000)
    ap.add_argument("--alleles", type=int, default=2)
    ap.add_argument("--pairs-per-sv", type=int, default=100)
    ap.add_argument("--threads", type=int, default=0, help="reference -t and our host helper threads (0 = min(48, cores))")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    from pansvr_b200 import aln
    from oracle import synth_pipeline as sp
    threads = a.threads or min(48, os.cpu_count() or 1)
    wd = tempfile.mkdtemp(prefix="pansvr_pipe_")
    try:
        t0 = time.time()
        d = sp.make_demo(wd, seed=31, genome_len=5000 + 7000 * a.loci + 4000, n_sv=a.loci, alleles_per_locus=a.alleles,
                         pairs_per_sv=a.pairs_per_sv, sv_lens=(50, 80, 150, 300, 600, 1000, 3000))
        t_data = time.time() - t0
        n_reads = 2 * d.n_pairs
        # ---- reference: start-up (-R 1), -t 1 (the SAM oracle), -t N (the CPU baseline)
        t_start = min(sp.run_reference_aln(d, os.path.join(wd, "s.sam"), os.path.join(wd, "s_ori.sam"), threads=1, extra=("-R", "1"))
                      for _ in range(2))
        t_ref1 = sp.run_reference_aln(d, os.path.join(wd, "ref1.sam"), os.path.join(wd, "ref1_ori.sam"), threads=1)
        t_refN = min(sp.run_reference_aln(d, os.path.join(wd, "refN.sam"), os.path.join(wd, "refN_ori.sam"), threads=min(threads, 48))
                     for _ in range(2))
        # ---- ours
        t0 = time.time()
        ctx = aln.AlnContext(d.index_dir, d.header_sam, threads=threads)
        t_create = time.time() - t0
        fq = open(d.reads_fq, "rb").read()
        times, last = [], None
        hdr = ctx.header_text().encode()
        for _ in range(a.repeat + 1):                                          # first pass warms the device buffers
            ctx.reset()                                                        # fresh replay state: same bytes every run
            t0 = time.time()
            sam_v, ori_v, release = ctx.align_fastq_view(fq)                    # the C ABI's own host buffers, no Python copy
            times.append(time.time() - t0)
            sam, ori = bytes(sam_v), bytes(ori_v)
            release()
            last = ctx.stats()
        times = times[1:]
        ctx.close()
        # ---- whole command lines, wall clock (start-up, FASTQ file in, SAM files out): what a user of `fc_aln` sees
        import subprocess
        cli = [sys.executable, "-c", "import sys; from pansvr_b200 import aln; sys.exit(aln.fc_aln_main(sys.argv[1:]))",
               "-t", str(threads), "-S", "-o", os.path.join(wd, "cli.sam"), "-p", os.path.join(wd, "cli_ori.sam"),
               d.index_dir, d.reads_fq, d.header_sam]
        t_cli = []
        for _ in range(2):
            t0 = time.time()
            subprocess.check_call(cli, cwd=ROOT, stderr=subprocess.DEVNULL)
            t_cli.append(time.time() - t0)
        cli_same = open(os.path.join(wd, "cli.sam"), "rb").read() == open(os.path.join(wd, "ref1.sam"), "rb").read()
        same = (hdr + sam == open(os.path.join(wd, "ref1.sam"), "rb").read()) and (hdr + ori == open(os.path.join(wd, "ref1_ori.sam"), "rb").read())
        best = min(times)
        line = {
            "what": "aln stage end to end (FASTQ text in, SAM text out), synthetic multi-allele anchors",
            "reads": n_reads, "anchors": d.n_sv, "index_load_excluded": True,
            "pansvr_b200": {"reads_per_s": n_reads / best, "seconds": best, "all_runs": times, "host_threads": threads,
                            "create_seconds": t_create, "stage_seconds": last["stage_seconds"], "ksw_tasks": last["ksw_tasks"],
                            "ksw_cells": last["ksw_cells"], "mems": last["mems"]},
            "reference": {"startup_seconds": t_start,
                          "t1": {"seconds": t_ref1, "reads_per_s": n_reads / max(t_ref1 - t_start, 1e-9)},
                          "tN": {"threads": min(threads, 48), "seconds": t_refN, "reads_per_s": n_reads / max(t_refN - t_start, 1e-9)}},
            "command_line_wall_seconds": {"pansvr_b200": min(t_cli), "reference_tN": t_refN, "reference_t1": t_ref1,
                                          "pansvr_b200_sam_identical": bool(cli_same)},
            "sam_identical_to_reference_t1": bool(same),
            "data_seconds": t_data,
        }
        txt = json.dumps(line)
        print(txt)
        if a.out:
            with open(a.out, "w") as f:
                f.write(txt + "\n")
    finally:
        if not a.keep:
            shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
