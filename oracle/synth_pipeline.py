"""Seeded synthetic inputs for the pipeline-level path (`panSVR fc_aln`), SURVEY.md section 8d "Config 1/3".

The reference ships no demo data, so a demo is synthesised: a random genome, a VCF of
INS/DEL structural variants, and interleaved "signal" read pairs sampled from the ALT haplotype whose FASTQ
comment carries the original alignment in the wire format `fc_signal` writes
(src/PanSVgenerateVCF/getSignalRead.cpp:158-249):

  tid_pos_softL_score_mapq_matemapq_XA_mateXA_isize_FLAGS_MATEFLAGS_[STAT_len_min_mid_max_]FLAG_f_q_CIGAR_c_MATE_tid_pos_isize_TAG_NM:i:n_

Anchor FASTA and the deBGA index are produced by the reference's own tools (oracle/_ref/panSVR fc_anchor_ref,
oracle/_ref/deBGA index) -- they are input-preparation stages outside the hot path -- and the SAM oracle by
`oracle/_ref/panSVR fc_aln -t 1 -S` (deterministic only single-threaded).
Input constraints of the reference: no '_' in chromosome names and VCF ids, tid <= 24.
"""
from __future__ import annotations

import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
COMP[list(b"ACGTN")] = list(b"TGCAN")


def have_reference_tools() -> bool:
    return all(os.access(os.path.join(REF_BIN, b), os.X_OK) for b in ("panSVR", "deBGA"))


@dataclass
class PipelineData:
    workdir: str
    ref_fa: str
    vcf: str
    anchors_fa: str
    index_dir: str
    reads_fq: str
    header_sam: str
    n_pairs: int
    n_sv: int


def _wrap(seq: bytes, width: int = 70) -> str:
    return "\n".join(seq[i:i + width].decode() for i in range(0, len(seq), width))


def _revcomp(a: np.ndarray) -> np.ndarray:
    return COMP[a[::-1]]


def make_demo(workdir: str, seed: int = 7, genome_len: int = 200_000, n_sv: int = 25, sv_step: int = 7000,
              sv_lens=(60, 150, 400, 1000), alleles_per_locus: int = 1, pairs_per_sv: int = 40, read_len: int = 150,
              sub_rate: float = 0.01, frag=(300, 500), edge_len: int = 500, build_index: bool = True,
              n_frac: float = 0.0, n_chrom: int = 1, mate_elsewhere: float = 0.0, str_every: int = 0,
              shared_insert: int = 0) -> PipelineData:
    """SURVEY.md 8d config 1 with the defaults; alleles_per_locus > 1 gives config-3-like shared flanks
    (several INS alleles at one locus => multi-candidate reads and exact score ties).  n_chrom > 1 spreads the loci
    round-robin over chromosomes "1", "2", ...; mate_elsewhere is the fraction of pairs whose ORIGINAL alignment had the
    mate on another chromosome (RNEXT / MATE_ fields of the -p output).  str_every = k makes every k-th inserted allele a
    short tandem repeat (unit 2-6 bp) so that reads from it take the STR branch of the seeding loop (RR:553-598);
    shared_insert = L puts one common L-bp element into every inserted allele, so its unipath has as many reference
    positions as there are INS alleles (> 500 of them switch expand_seed to its random_r sampling, IDX:219-258)."""
    os.makedirs(workdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    genomes = [ACGT[rng.integers(0, 4, genome_len)] for _ in range(n_chrom)]
    chroms = [str(c + 1) for c in range(n_chrom)]
    ref_fa = os.path.join(workdir, "ref.fa")
    with open(ref_fa, "w") as f:
        for chrom, genome in zip(chroms, genomes):
            f.write(f">{chrom}\n{_wrap(genome.tobytes())}\n")
    # ---- variants
    common = ACGT[np.random.default_rng(seed + 1000).integers(0, 4, shared_insert)] if shared_insert else None
    svs = []   # (pos1 (1-based, anchor base), type, ref_bytes, alt_bytes, id, chromosome index)
    for k in range(n_sv):
        ci = k % n_chrom
        genome = genomes[ci]
        pos = 5000 + (k // n_chrom) * sv_step
        if pos + 2000 > genome_len:
            break
        L = int(sv_lens[k % len(sv_lens)])
        kind = "INS" if k % 2 == 0 else "DEL"
        base = genome[pos - 1:pos]
        for a in range(alleles_per_locus if kind == "INS" else 1):
            if kind == "INS":
                ins = ACGT[rng.integers(0, 4, L + 37 * a)]
                if str_every and (k // 2) % str_every == 0:
                    unit = ACGT[rng.integers(0, 4, int(rng.integers(2, 7)))]
                    ins = np.tile(unit, (L + 37 * a) // unit.size + 1)[:L + 37 * a]
                if shared_insert:
                    ins = np.concatenate([ins[:ins.size // 2], common, ins[ins.size // 2:]])
                svs.append((pos, "INS", base.tobytes(), base.tobytes() + ins.tobytes(), f"sv{k}a{a}", ci))
            else:
                svs.append((pos, "DEL", genome[pos - 1:pos + L].tobytes(), base.tobytes(), f"sv{k}a{a}", ci))
    svs.sort(key=lambda v: v[5])        # the VCF is grouped by chromosome (stable: positions stay ascending)
    vcf = os.path.join(workdir, "sv.vcf")
    with open(vcf, "w") as f:
        f.write("##fileformat=VCFv4.2\n")
        for chrom in chroms:
            f.write(f"##contig=<ID={chrom},length={genome_len}>\n")
        f.write('##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">\n')
        f.write('##INFO=<ID=SVLEN,Number=1,Type=Integer,Description="Length of structural variant">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for pos, kind, r, a, vid, ci in svs:
            svlen = len(a) - len(r)
            f.write(f"{chroms[ci]}\t{pos}\t{vid}\t{r.decode()}\t{a.decode()}\t.\tPASS\tSVTYPE={kind};SVLEN={svlen}\n")
    header_sam = os.path.join(workdir, "header.sam")
    with open(header_sam, "w") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{chrom}\tLN:{genome_len}\n" for chrom in chroms))
    # ---- read pairs from the ALT haplotype of each allele
    reads_fq = os.path.join(workdir, "reads.fq")
    n_pairs = 0
    with open(reads_fq, "w") as f:
        first = True
        for si, (pos, kind, r, a, vid, ci) in enumerate(svs):
            genome = genomes[ci]
            lo = max(0, pos - 1 - 700)
            left = genome[lo:pos - 1]
            alt = np.frombuffer(a, dtype=np.uint8)
            right = genome[pos - 1 + len(r):pos - 1 + len(r) + 700]
            hap = np.concatenate([left, alt, right])
            bp = len(left)                                   # haplotype offset of the variant
            for p in range(pairs_per_sv):
                flen = int(rng.integers(frag[0], frag[1] + 1))
                centre = bp + int(rng.integers(-flen // 2, len(alt) + flen // 2))
                st = int(np.clip(centre - flen // 2, 0, len(hap) - flen))
                fr = hap[st:st + flen].copy()
                r1, r2 = fr[:read_len].copy(), _revcomp(fr[-read_len:])
                for rd in (r1, r2):
                    m = rng.random(read_len) < sub_rate
                    rd[m] = ACGT[(np.searchsorted(ACGT, rd[m]) + rng.integers(1, 4, int(m.sum()))) % 4]
                    if n_frac > 0:
                        rd[rng.random(read_len) < n_frac] = ord("N")
                # the "original" alignment: what a linear-reference aligner would have reported
                ref_pos1 = max(0, lo + st)
                ref_pos2 = max(0, lo + st + flen - read_len - (len(a) - len(r)))
                isize = flen
                soft1 = int(rng.integers(0, 70)) if rng.random() < 0.5 else 0
                soft2 = int(rng.integers(0, 70)) if rng.random() < 0.5 else 0
                soft1, soft2 = min(soft1, read_len - 25), min(soft2, read_len - 25)   # a clip never exceeds the read (fc_signal copies a real CIGAR)
                sc1 = 2 * (read_len - soft1) - int(rng.integers(10, 90))
                sc2 = 2 * (read_len - soft2) - int(rng.integers(10, 90))
                name = f"r{si}x{p}"
                mate_tid = ci
                if mate_elsewhere > 0 and rng.random() < mate_elsewhere:
                    mate_tid = (ci + 1) % n_chrom
                for mate, (rd, rp, sl, sc, fl, mfl, flag, mp) in enumerate((
                        (r1, ref_pos1, soft1, sc1, "FNNY", "RNNY", 99, ref_pos2),
                        (r2, ref_pos2, soft2, sc2, "RNNY", "FNNY", 147, ref_pos1))):
                    stat = f"STAT_{read_len}_{frag[0]}_{(frag[0] + frag[1]) // 2}_{frag[1]}_" if first and mate == 0 else ""
                    cig = (f"{sl}S{read_len - sl}M" if sl else f"{read_len}M")
                    comment = (f"{ci}_{rp}_{sl}_{sc}_60_60_0_0_{isize}_{fl}_{mfl}_{stat}FLAG_{flag}_60_CIGAR_{cig}_"
                               f"MATE_{mate_tid}_{mp}_{isize if mate == 0 else -isize}_TAG_NM:i:{int(rng.integers(0, 6))}_")
                    qual = "I" * read_len
                    f.write(f"@{name} {comment}\n{rd.tobytes().decode()}\n+\n{qual}\n")
                first = False
                n_pairs += 1
    anchors_fa = os.path.join(workdir, "anchors.fa")
    index_dir = os.path.join(workdir, "idx") + "/"
    data = PipelineData(workdir, ref_fa, vcf, anchors_fa, index_dir, reads_fq, header_sam, n_pairs, len(svs))
    if build_index:
        build_anchor_index(data, edge_len)
    return data


def build_anchor_index(d: PipelineData, edge_len: int = 500) -> None:
    """S1 + S2 of the reference pipeline with the reference's own tools (input preparation, not the hot path)."""
    if not have_reference_tools():
        raise FileNotFoundError("oracle/_ref/panSVR and oracle/_ref/deBGA are needed to prepare anchors and the index "
                                "(oracle/build_ref_pipeline.sh)")
    with open(d.anchors_fa, "w") as out, open(os.path.join(d.workdir, "anchor.log"), "w") as log:
        subprocess.check_call([os.path.join(REF_BIN, "panSVR"), "fc_anchor_ref", "-e", str(edge_len), d.ref_fa, d.vcf],
                              stdout=out, stderr=log)
    os.makedirs(d.index_dir, exist_ok=True)
    with open(os.path.join(d.workdir, "index.log"), "w") as log:
        subprocess.check_call([os.path.join(REF_BIN, "deBGA"), "index", "-k", "22", d.anchors_fa, d.index_dir],
                              stdout=log, stderr=log)


def run_reference_aln(d: PipelineData, out_sam: str, ori_sam: str, threads: int = 1, extra=(), bam: bool = False) -> float:
    """`panSVR fc_aln -t N -S` of the reference (BAM files without -S when bam=True); returns wall seconds.
    threads=1 is the output oracle."""
    import time
    t0 = time.time()
    with open(os.path.join(d.workdir, "fc_aln.log"), "w") as log:
        subprocess.check_call([os.path.join(REF_BIN, "panSVR"), "fc_aln", "-t", str(threads), *(() if bam else ("-S",)),
                               "-o", out_sam, "-p", ori_sam, *extra, d.index_dir, d.reads_fq, d.header_sam], stdout=log, stderr=log)
    return time.time() - t0
