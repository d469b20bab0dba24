// gen_signal.cpp -- seeded synthetic input of the `fc_aln` stage at SURVEY.md section 8d "Config 3" / "Config 5" scale.
// BENCH / TEST TOOLING, not part of the product library.
//
// Writes into <out_dir>: ref.fa (random genome), sv.vcf (INS / DEL structural variants, several INS alleles per locus
// sharing their flanks => multi-candidate reads and exact score ties), header.sam (header of the "original" BAM) and
// reads.fq: interleaved signal read pairs sampled from the ALT haplotypes, whose FASTQ comment carries the original
// alignment in the wire format `fc_signal` writes (src/PanSVgenerateVCF/getSignalRead.cpp:158-249):
//   tid_pos_softL_score_mapq_matemapq_XA_mateXA_isize_FLAGS_MATEFLAGS_[STAT_len_min_mid_max_]FLAG_f_q_CIGAR_c_MATE_tid_pos_isize_TAG_NM:i:n_
// Same data model as oracle/synth_pipeline.py:make_demo (which stays the generator of the small parity data sets); this
// one exists because 10 M reads take minutes in numpy and a few seconds here.  Anchors and the deBGA index are then made by
// the reference's own tools (benchdata/config3.py).
//
//   gen_signal <out_dir> [--seed S] [--loci N] [--alleles-max K] [--pairs-per-sv P] [--read-len L] [--step B] [--n-frac F]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

namespace {

struct Rng {                                  // splitmix64
	uint64_t s;
	explicit Rng(uint64_t seed) : s(seed) {}
	uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
	uint32_t below(uint32_t n) { return (uint32_t)((next() >> 11) % n); }
	int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo)); }      // [lo, hi)
	double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

const char ACGT[5] = "ACGT";
inline char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }

struct Sv { long pos; bool ins; std::string ref, alt, id; };

struct Out {                                  // buffered file
	FILE *f; std::vector<char> buf; size_t n = 0;
	explicit Out(const std::string &path) : f(fopen(path.c_str(), "wb")), buf(8 << 20) { if (!f) { perror(path.c_str()); exit(1); } }
	~Out() { flush(); fclose(f); }
	void flush() { if (n) { if (fwrite(buf.data(), 1, n, f) != n) { perror("write"); exit(1); } n = 0; } }
	void put(const char *p, size_t l) { if (n + l > buf.size()) flush(); if (l > buf.size()) { fwrite(p, 1, l, f); return; } memcpy(buf.data() + n, p, l); n += l; }
	void put(const std::string &s) { put(s.data(), s.size()); }
	void put(char c) { put(&c, 1); }
	void num(long v) { char b[24]; int k = snprintf(b, sizeof b, "%ld", v); put(b, (size_t)k); }
};

} // namespace

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: gen_signal <out_dir> [--seed S] [--loci N] [--alleles-max K] [--pairs-per-sv P] [--read-len L] [--step B] [--n-frac F]\n"); return 1; }
	const std::string dir = argv[1];
	uint64_t seed = 31; long loci = 5250, pairs_per_sv = 476, step = 13000; int alleles_max = 4, read_len = 150; double n_frac = 0, sub_rate = 0.01;
	for (int i = 2; i + 1 < argc; i += 2) {
		const std::string k = argv[i]; const char *v = argv[i + 1];
		if (k == "--seed") seed = strtoull(v, 0, 10); else if (k == "--loci") loci = atol(v); else if (k == "--alleles-max") alleles_max = atoi(v);
		else if (k == "--pairs-per-sv") pairs_per_sv = atol(v); else if (k == "--read-len") read_len = atoi(v); else if (k == "--step") step = atol(v);
		else if (k == "--n-frac") n_frac = atof(v); else if (k == "--sub-rate") sub_rate = atof(v);
		else { fprintf(stderr, "gen_signal: unknown option %s\n", k.c_str()); return 1; }
	}
	static const int sv_lens[8] = {50, 80, 150, 300, 600, 1000, 3000, 10000};
	const int frag_lo = read_len == 150 ? 300 : 2 * read_len, frag_hi = frag_lo + 200;
	const long genome_len = 5000 + loci * step + 4000;
	Rng rng(seed);
	const double inv_log_sub = sub_rate > 0 ? 1.0 / log(1.0 - sub_rate) : 0, inv_log_n = n_frac > 0 ? 1.0 / log(1.0 - n_frac) : 0;
	std::string genome((size_t)genome_len, 'A');
	for (long i = 0; i < genome_len; ++i) genome[(size_t)i] = ACGT[rng.next() >> 62];
	{
		Out fa(dir + "/ref.fa");
		fa.put(">1\n");
		for (long i = 0; i < genome_len; i += 70) { fa.put(genome.data() + i, (size_t)std::min<long>(70, genome_len - i)); fa.put('\n'); }
	}
	// ---- variants: loci alternate INS / DEL; an INS locus carries 2..alleles_max alleles of different lengths at one position
	std::vector<Sv> svs;
	for (long k = 0; k < loci; ++k) {
		const long pos = 5000 + k * step;                       // 1-based anchor base
		const int L = sv_lens[k % 8];
		const bool ins = (k % 2) == 0;
		const char base = genome[(size_t)pos - 1];
		const int n_alleles = ins ? 2 + (int)((k / 2) % (alleles_max - 1)) : 1;
		for (int a = 0; a < n_alleles; ++a) {
			Sv s; s.pos = pos; s.ins = ins; s.id = "sv" + std::to_string(k) + "a" + std::to_string(a);
			if (ins) {
				s.ref.assign(1, base); s.alt.assign(1, base);
				const int len = L + 37 * a;
				for (int i = 0; i < len; ++i) s.alt += ACGT[rng.next() >> 62];
			} else { s.ref = genome.substr((size_t)pos - 1, (size_t)L + 1); s.alt.assign(1, base); }
			svs.push_back(std::move(s));
		}
	}
	{
		Out vcf(dir + "/sv.vcf");
		vcf.put("##fileformat=VCFv4.2\n##contig=<ID=1,length="); vcf.num(genome_len); vcf.put(">\n");
		vcf.put("##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">\n");
		vcf.put("##INFO=<ID=SVLEN,Number=1,Type=Integer,Description=\"Length of structural variant\">\n");
		vcf.put("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n");
		for (const Sv &s : svs) {
			vcf.put("1\t"); vcf.num(s.pos); vcf.put('\t'); vcf.put(s.id); vcf.put('\t'); vcf.put(s.ref); vcf.put('\t'); vcf.put(s.alt);
			vcf.put("\t.\tPASS\tSVTYPE="); vcf.put(s.ins ? "INS" : "DEL"); vcf.put(";SVLEN="); vcf.num((long)s.alt.size() - (long)s.ref.size()); vcf.put('\n');
		}
		Out hdr(dir + "/header.sam");
		hdr.put("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:1\tLN:"); hdr.num(genome_len); hdr.put('\n');
	}
	// ---- read pairs from the ALT haplotype of each allele
	Out fq(dir + "/reads.fq");
	const std::string qual((size_t)read_len, 'I');
	std::string hap, r1, r2;
	long n_pairs = 0;
	bool first = true;
	for (size_t si = 0; si < svs.size(); ++si) {
		const Sv &s = svs[si];
		const long lo = std::max<long>(0, s.pos - 1 - 700);
		hap.assign(genome, (size_t)lo, (size_t)(s.pos - 1 - lo));
		const long bp = (long)hap.size();
		hap += s.alt;
		const long right_b = s.pos - 1 + (long)s.ref.size();
		hap.append(genome, (size_t)right_b, (size_t)std::min<long>(700, genome_len - right_b));
		const long alt_l = (long)s.alt.size(), ref_l = (long)s.ref.size();
		for (long p = 0; p < pairs_per_sv; ++p) {
			const int flen = rng.range(frag_lo, frag_hi + 1);
			const long centre = bp + rng.range(-(flen / 2), (int)alt_l + flen / 2);
			const long st = std::min<long>(std::max<long>(centre - flen / 2, 0), (long)hap.size() - flen);
			r1.assign(hap, (size_t)st, (size_t)read_len);
			r2.resize((size_t)read_len);
			for (int i = 0; i < read_len; ++i) r2[(size_t)i] = comp(hap[(size_t)(st + flen - 1 - i)]);
			for (std::string *rd : {&r1, &r2}) {
				// substitutions (and N) at geometric gaps: one draw per event instead of one per base
				if (sub_rate > 0)
					for (double i = log(1.0 - rng.unit()) * inv_log_sub; i < read_len; i += 1.0 + log(1.0 - rng.unit()) * inv_log_sub) {
						char &c = (*rd)[(size_t)i];
						const char *q = strchr(ACGT, c);
						c = ACGT[((q ? q - ACGT : 0) + 1 + rng.below(3)) & 3];
					}
				if (n_frac > 0)
					for (double i = log(1.0 - rng.unit()) * inv_log_n; i < read_len; i += 1.0 + log(1.0 - rng.unit()) * inv_log_n) (*rd)[(size_t)i] = 'N';
			}
			const long ref_pos1 = std::max<long>(0, lo + st), ref_pos2 = std::max<long>(0, lo + st + flen - read_len - (alt_l - ref_l));
			int soft1 = rng.unit() < 0.5 ? rng.range(0, 70) : 0, soft2 = rng.unit() < 0.5 ? rng.range(0, 70) : 0;
			soft1 = std::min(soft1, read_len - 25); soft2 = std::min(soft2, read_len - 25);
			const int sc1 = 2 * (read_len - soft1) - rng.range(10, 90), sc2 = 2 * (read_len - soft2) - rng.range(10, 90);
			for (int mate = 0; mate < 2; ++mate) {
				const std::string &rd = mate ? r2 : r1;
				const long rp = mate ? ref_pos2 : ref_pos1, mp = mate ? ref_pos1 : ref_pos2;
				const int sl = mate ? soft2 : soft1, sc = mate ? sc2 : sc1;
				fq.put("@r"); fq.num((long)si); fq.put('x'); fq.num(p); fq.put(" 0_"); fq.num(rp); fq.put('_'); fq.num(sl); fq.put('_'); fq.num(sc);
				fq.put("_60_60_0_0_"); fq.num(flen); fq.put(mate ? "_RNNY_FNNY_" : "_FNNY_RNNY_");
				if (first && mate == 0) { fq.put("STAT_"); fq.num(read_len); fq.put('_'); fq.num(frag_lo); fq.put('_'); fq.num((frag_lo + frag_hi) / 2); fq.put('_'); fq.num(frag_hi); fq.put('_'); }
				fq.put("FLAG_"); fq.num(mate ? 147 : 99); fq.put("_60_CIGAR_");
				if (sl) { fq.num(sl); fq.put('S'); }
				fq.num(read_len - sl); fq.put("M_MATE_0_"); fq.num(mp); fq.put('_'); fq.num(mate ? -flen : flen); fq.put("_TAG_NM:i:"); fq.num(rng.below(6)); fq.put("_\n");
				fq.put(rd); fq.put("\n+\n"); fq.put(qual); fq.put('\n');
			}
			first = false;
			++n_pairs;
		}
	}
	printf("{\"pairs\": %ld, \"anchors\": %zu, \"loci\": %ld, \"genome_len\": %ld, \"read_len\": %d}\n", n_pairs, svs.size(), loci, genome_len, read_len);
	return 0;
}
