"""Generates tests/golden/aln_demo.sam.gz / aln_demo_ori.sam.gz: the REFERENCE's own `panSVR fc_aln -t 1 -S` output on the
seeded synthetic demo (oracle/synth_pipeline.make_demo defaults).  Needs oracle/_ref/panSVR and
oracle/_ref/deBGA (oracle/build_ref_pipeline.sh, build container only).  The inputs are regenerated from the seed by the
tests; only the reference's output is stored."""
import gzip
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth_pipeline as sp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    wd = tempfile.mkdtemp(prefix="aln_golden_")
    try:
        d = sp.make_demo(wd)
        sp.run_reference_aln(d, os.path.join(wd, "ref.sam"), os.path.join(wd, "ref_ori.sam"))
        for src, dst in (("ref.sam", "aln_demo.sam.gz"), ("ref_ori.sam", "aln_demo_ori.sam.gz")):
            with open(os.path.join(wd, src), "rb") as f, gzip.GzipFile(os.path.join(HERE, dst), "wb", mtime=0) as g:
                g.write(f.read())
            print("wrote", dst, os.path.getsize(os.path.join(HERE, dst)))
    finally:
        shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
