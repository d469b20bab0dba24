// stages_core.cuh -- the per-read stages of `fc_aln` between the FASTQ text and the candidate alignments, as host/device
// functions over flat arrays: one thread per read (or per candidate) on the GPU (stages_gpu.cu), the same functions stepped
// by a host loop in the CPU-only test build (tests/emul).  Same results as the reference's per-read code:
//
//   stage A  encode_read      binary_read_2_bit + binary_read_64_bit + the STR census      RR:646-654, 295-300, 553-601
//   stage C  merge_strand     merge_seed_in_unipath                                        IDX:151-217
//            chain_strand     expand_seed + Graph_handler::process / dynamic_programming_path   IDX:219-258, GR:53-150
//   stage D  plan_read        the chain ends get_ksw_score can be asked for, planned:       RR:308-400, 893-986, IDX:307
//                             literal CIGAR pieces, fixed score part, ksw windows (query / target bytes written out)
//   stage F1 resolve_cand     score sum + CIGAR assembly + reverseGIGAR once ksw has run    RR:967-983, RRH:159-178, 277-301
//
// (RR = src/PanSVgenerateVCF/read_realignment.cpp, RRH = .hpp, IDX = deBGA_index.cpp, GR = src/cpp_lib/graph.cpp.)
// What stays with the host pipeline (pipeline.cpp): reads whose unipaths need random_r sampling (> 500 positions), reads with
// a lower-case 'n' (code 4 spills into the packed neighbour) or more than three 'N', and everything that consumes rand().
#pragma once
#include <stdint.h>

#include "seed_core.cuh"

namespace pansvr {

enum { ST_POS_N_MAX = 500, ST_POS_N_MAX_LEVEL2 = 8000, ST_WAITING_LEN = 3 };
enum { ST_MIN_CHAIN_SCORE = 20, ST_MAX_CHAIN_SCORE_DIFF = 30, ST_MIN_CHAIN_SCORE2 = 30 };
enum { ST_ALN_LEFT = 0, ST_ALN_RIGHT = 1, ST_ALN_E2E = 2 };
enum { ST_FLAG_NEEDS_RAND = 1, ST_FLAG_STR = 2, ST_FLAG_HOST = 4, ST_FLAG_NOSEED = 8 };
const int ST_NEG_INF = -0x40000000;

struct DevRead {                  // one read state of the device path: a read, or one substitution variant of a read with 1..3 'N'
	uint32_t seq_off;             // offset of its bases in the block's text
	uint32_t len;
	uint32_t var_code;            // variant: the j-th 'N' becomes "ACGT"[(var_code >> 2j) & 3]
	uint32_t bits_off;            // word offset of the packed forward strand; the reverse strand follows (words = (len >> 5) + 2 each)
	uint32_t list_off;            // offset of the forward seed list (len - 19 bytes, STR reads only); the reverse one follows
};

struct DevSeed { uint32_t read_begin, read_end, seed_id, ref_begin, ref_end, cov; };     // UNI_SEED, graph.hpp:42-49
struct DevVertex { uint64_t uid; uint32_t read_pos, uni_pos_off, length1, length2, pos_n, cov; };   // vertex_U (same 32 bytes as Mem)

struct DevCand {                  // one (strand, chain end) of a read that get_ksw_score can be asked for
	uint32_t read, node;          // node: index in the strand's sorted seed list
	uint32_t strand;
	int32_t fixed_score;          // everything but the ksw scores
	int32_t read_begin_alignment;
	uint32_t piece_off, n_pieces;
	uint32_t cig_off, cig_cap;    // where resolve_cand writes the final CIGAR
	// filled by resolve_cand
	uint32_t align_score, n_cig, cigar_ok;
};
struct DevPiece { int32_t task; uint8_t kind, type, lit_type, pad; int16_t lit_size; int16_t pad2; };   // kind 0 = literal CIGAR entry, 1 = ksw task
struct DevCigar { uint8_t type; uint8_t pad; int16_t size; };

struct AlnScores { int match, mismatch, gap_open, gap_ex, gap_open2, gap_ex2; };

// ---------------------------------------------------------------------------------------------------------- small tools
template <class T, class Less>
SEED_HD void stable_sort(T *a, T *tmp, uint32_t n, Less less)
{
	// insertion sort on runs of 16, then bottom-up merges through tmp: stable, so the order equals glibc's merge-sort qsort for
	// comparators that are consistent orders (vertex_MEM::cmp, UNI_SEED::cmp)
	const uint32_t RUN = 16;
	for (uint32_t b = 0; b < n; b += RUN) {
		const uint32_t e = b + RUN < n ? b + RUN : n;
		for (uint32_t i = b + 1; i < e; ++i) {
			const T x = a[i];
			uint32_t j = i;
			while (j > b && less(x, a[j - 1])) { a[j] = a[j - 1]; --j; }
			a[j] = x;
		}
	}
	T *src = a, *dst = tmp;
	for (uint32_t w = RUN; w < n; w <<= 1) {
		for (uint32_t lo = 0; lo < n; lo += 2 * w) {
			const uint32_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
			uint32_t i = lo, j = mid, k = lo;
			while (i < mid && j < hi) dst[k++] = less(src[j], src[i]) ? src[j++] : src[i++];
			while (i < mid) dst[k++] = src[i++];
			while (j < hi) dst[k++] = src[j++];
		}
		T *t = src; src = dst; dst = t;
	}
	if (src != a) for (uint32_t i = 0; i < n; ++i) a[i] = src[i];
}

SEED_HD uint8_t dna5_code(uint8_t c)                               // charToDna5n, RR:180-202 (anything else, 'N' included, is 0)
{
	switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; case 'n': return 4; default: return 0; }
}
SEED_HD uint32_t packed_base(const uint64_t *bits, uint32_t i) { return (uint32_t)(bits[i >> 5] >> ((31 - (i & 31)) << 1)) & 3u; }

// ---------------------------------------------------------------------------------------------------------- stage A
// Packs both strands of read state `rd` and runs the STR census of the forward strand.  Returns ENC_STR for an STR read (whose
// seed lists are then written), ENC_HAS_N if the read holds an 'N' or a lower-case 'n' (then nothing else about the result is
// defined: the host keeps such a pair), else ENC_PLAIN.
// filter: ENC_FILTER_WORDS words of scratch, word k at filter[k * fstride] (the CUDA backend keeps them in shared memory,
// interleaved over the threads of a CTA so that the census's random accesses never conflict).
enum { ENC_PLAIN = 0, ENC_STR = 1, ENC_HAS_N = 2, ENC_FILTER_WORDS = 64 };
SEED_HD uint64_t revcomp_word(uint64_t x)                          // the 32 bases of a word in reverse order, complemented
{
	x = (x >> 32) | (x << 32);
	x = ((x >> 16) & 0x0000ffff0000ffffull) | ((x & 0x0000ffff0000ffffull) << 16);
	x = ((x >> 8) & 0x00ff00ff00ff00ffull) | ((x & 0x00ff00ff00ff00ffull) << 8);
	x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
	x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
	return ~x;
}
SEED_HD int encode_read(const uint8_t *text, const DevRead &rd, uint64_t *bits_pool, uint8_t *list_pool, uint32_t *filter, uint32_t fstride)
{
	const uint32_t L = rd.len, words = (L >> 5) + 2;
	uint64_t *fw = bits_pool + rd.bits_off, *rv = fw + words;
	for (uint32_t k = 0; k < words; ++k) { fw[k] = 0; rv[k] = 0; }
	for (uint32_t k = 0; k < ENC_FILTER_WORDS; ++k) filter[k * fstride] = 0;
	// ---- one pass over the bases (read a word of text at a time): forward strand packed, and the STR census (RR:553-598) on the
	// fly.  The read is an STR read when fewer than kn - 15 of its kn 20-mers are distinct; duplicates are counted through a
	// 2048-bit filter first: a k-mer whose bit is already set MAY be a duplicate, so fewer than 16 such events prove the read is
	// not STR (almost every read); otherwise the multiplicities are counted exactly below.
	const uint64_t kmask = (1ull << (2 * LEN_KMER)) - 1;
	const uint32_t *tw = (const uint32_t*)(text + (rd.seq_off & ~3u));
	uint32_t cur = tw[0] >> (8 * (rd.seq_off & 3u)), have = 4 - (rd.seq_off & 3u), next = 1;
	uint32_t nth = 0, maybe_dup = 0;
	bool has_n = false;
	uint64_t w = 0, roll = 0;
	for (uint32_t i = 0; i < L; ++i) {
		if (have == 0) { cur = tw[next++]; have = 4; }
		const uint8_t ch = (uint8_t)(cur & 0xffu);
		cur >>= 8; --have;
		uint32_t c;
		if (ch == 'N') { c = (rd.var_code >> (2 * nth)) & 3u; ++nth; has_n = true; }      // "ACGT"[k] encodes to k
		else { c = dna5_code(ch) & 3u; has_n |= ch == 'n'; }
		w = (w << 2) | c;
		if ((i & 31) == 31) { fw[i >> 5] = w; w = 0; }
		roll = ((roll << 2) | c) & kmask;
		if (i + 1 >= LEN_KMER) {
			const uint32_t x = (uint32_t)roll ^ (uint32_t)(roll >> 19);
			const uint32_t h = (x * 0x9E3779B1u) >> 21, bit = 1u << (h & 31);
			uint32_t &slot = filter[(h >> 5) * fstride];
			if (slot & bit) ++maybe_dup; else slot |= bit;
		}
	}
	if (L & 31) fw[L >> 5] = w << ((32 - (L & 31)) << 1);
	if (has_n && rd.var_code == 0) return ENC_HAS_N;
	{                                                              // reverse complement, a word at a time: base j of rv = 3 - base (L-1-j) of fw
		const uint32_t nw = (L + 31) >> 5, pad2 = 2 * (32 * nw - L);  // the packed string ends pad2 bits before the end of its last word
		uint64_t prev = nw ? revcomp_word(fw[nw - 1]) : 0;
		for (uint32_t k = 0; k < nw; ++k) {
			const uint64_t nxt = k + 1 < nw ? revcomp_word(fw[nw - 2 - k]) : 0;
			rv[k] = pad2 ? (prev << pad2) | (nxt >> (64 - pad2)) : prev;
			prev = nxt;
		}
	}
	const uint32_t kn = L - LEN_KMER + 1;
	if (maybe_dup <= 15) return ENC_PLAIN;
	uint32_t distinct = 0;
	for (uint32_t i = 0; i < kn; ++i) {
		const uint64_t k = get_kmer(i, fw);
		bool seen = false;
		for (uint32_t j = 0; j < i && !seen; ++j) seen = get_kmer(j, fw) == k;
		distinct += !seen;
	}
	if (!(distinct < kn - 15)) return ENC_PLAIN;
	// multiplicity mask, forced seeds at both ends (RR:575-598), reversed copy for the other strand (RR:601)
	uint8_t *sl = list_pool + rd.list_off, *sr = sl + kn;
	for (uint32_t i = 0; i < kn; ++i) {
		const uint64_t k = get_kmer(i, fw);
		uint32_t cnt = 0;
		for (uint32_t j = 0; j < kn; ++j) cnt += get_kmer(j, fw) == k;
		sl[i] = cnt >= 4 ? 0 : 1;
	}
	int bg = 0, ed = 0;
	for (uint32_t i = 0; i < SEED_STEP; ++i) {
		bg += sl[i] == 0; ed += sl[L - LEN_KMER - i] == 0;
		sl[i] += 2; sl[L - LEN_KMER - i] += 4;
	}
	if (bg < SEED_STEP && ed < SEED_STEP) {
		int n = 0;
		for (uint32_t i = 0; n < SEED_STEP && i < kn; ++i) { if (sl[i] > 0) continue; sl[i] += 8; ++n; }
	}
	for (uint32_t i = 0; i < kn; ++i) sr[i] = sl[i];
	for (uint32_t i = 0; i < (kn >> 1) + 1; ++i) { const uint32_t ri = kn - 1 - i; const uint8_t t = sr[i]; sr[i] = sr[ri]; sr[ri] = t; }   // getReverseStr_qual (sic)
	return ENC_STR;
}

// ---------------------------------------------------------------------------------------------------------- stage C
struct MemLess { SEED_HD bool operator()(const Mem &x, const Mem &y) const { return x.uid != y.uid ? x.uid < y.uid : x.read_pos < y.read_pos; } };
struct SeedLess { SEED_HD bool operator()(const DevSeed &x, const DevSeed &y) const { return x.ref_end != y.ref_end ? x.ref_end < y.ref_end : x.ref_begin < y.ref_begin; } };

// Sorts the MEMs of one strand and merges co-linear ones in place (the vertex list overwrites the front of the MEM list; both
// records are 32 bytes).  Returns the number of vertices; *n_seeds = reference positions they expand to (expand_seed stops at
// the first vertex with more than 8000 positions).
SEED_HD uint32_t merge_strand(Mem *m, Mem *tmp, uint32_t n, uint32_t *n_seeds)
{
	*n_seeds = 0;
	if (n == 0) return 0;
	if (n > 1) stable_sort(m, tmp, n, MemLess());
	DevVertex *out = (DevVertex*)m;
	uint32_t nv = 0;
	uint64_t uid_t = m[0].uid;
	uint32_t j = 0;
	while (j < n) {
		const uint32_t s1 = j;
		uint32_t cov = m[s1].length;
		++j;
		while (j < n && uid_t == m[j].uid && m[j].uni_pos_off > m[j - 1].uni_pos_off) {
			const int diff = (int)(m[j].read_pos - m[j - 1].read_pos - m[j - 1].length);
			if (diff > ST_WAITING_LEN) break;
			const int c_eindel = (int)((m[j].uni_pos_off - m[j - 1].uni_pos_off) - (m[j].read_pos - m[j - 1].read_pos));
			if (c_eindel == 0) { cov += diff > 0 ? m[j].length : (uint32_t)(diff + (int)m[j].length); ++j; }    // abs(c_eindel) < Eindel (= 1)
			else break;
		}
		const uint32_t e1 = j - 1;
		DevVertex u;
		u.uid = m[s1].uid; u.read_pos = m[s1].read_pos; u.uni_pos_off = m[s1].uni_pos_off; u.pos_n = m[s1].pos_n; u.cov = cov;
		if (s1 == e1) u.length1 = u.length2 = m[s1].length;
		else {
			u.length1 = m[e1].read_pos + m[e1].length - m[s1].read_pos;
			u.length2 = m[e1].uni_pos_off + m[e1].length - m[s1].uni_pos_off;
		}
		if (j < n) uid_t = m[j].uid;
		out[nv++] = u;                                            // nv <= s1 + 1 <= j: everything at or after j is still a MEM
	}
	uint32_t total = 0;
	for (uint32_t i = 0; i < nv; ++i) {
		if (out[i].pos_n > ST_POS_N_MAX_LEVEL2) break;
		total += out[i].pos_n;
	}
	*n_seeds = total;
	return nv;
}

// Expands the vertices of one strand to reference positions, sorts the seeds and runs the chain DP.  dist/pre = dist_path.
SEED_HD void chain_strand(const DevVertex *vu, uint32_t nv, const uint64_t *pos, const uint64_t *posp, bool is_str,
                          DevSeed *v, DevSeed *tmp, uint32_t n, float *dist, int32_t *pre)
{
	uint32_t k = 0;
	for (uint32_t i = 0; i < nv && k < n; ++i) {
		const DevVertex &u = vu[i];
		if (u.pos_n > ST_POS_N_MAX_LEVEL2) break;
		for (uint32_t mpos = 0; mpos < u.pos_n; ++mpos) {
			DevSeed s;
			s.seed_id = i; s.read_begin = u.read_pos; s.read_end = u.read_pos + u.length1 - 1;
			s.ref_begin = (uint32_t)(pos[mpos + posp[u.uid]] + u.uni_pos_off - 1);
			s.ref_end = s.ref_begin + u.length2 - 1; s.cov = u.cov;
			v[k++] = s;
		}
	}
	if (n == 0) return;
	if (n > 1) stable_sort(v, tmp, n, SeedLess());
	const int max_ref_dis = is_str ? 400 : 50, max_read_dis = is_str ? 400 : 50;
	const uint32_t max_step = is_str ? 80 : 40, max_gap = is_str ? 20 : 50;
	const uint32_t step = n < max_step ? n : max_step;
	// `pre` doubles as "has an incoming edge" while the table is built: -2 = none yet
	for (uint32_t i = 0; i < n; ++i) { dist[i] = (float)v[i].cov; pre[i] = -2; }
	for (uint32_t a = 0; a + 1 < n; ++a) {
		// every edge into `a` comes from a smaller index, so its entry is final here
		const float dist_a = dist[a];
		const uint32_t read_end = v[a].read_end, ref_end = v[a].ref_end, seed_id = v[a].seed_id;
		const uint32_t stop = n < a + step ? n : a + step;
		for (uint32_t b = a + 1; b < stop; ++b) {
			if (v[b].seed_id == seed_id) continue;
			if (v[b].ref_end == ref_end) continue;
			const int32_t dis_ref = (int32_t)(v[b].ref_begin - ref_end);
			if (dis_ref > max_ref_dis) break;
			const int32_t dis_read = (int32_t)(v[b].read_begin - read_end);
			if (dis_read > max_read_dis) continue;
			const uint32_t abs_gap = dis_read > dis_ref ? (uint32_t)(dis_read - dis_ref) : (uint32_t)(dis_ref - dis_read);
			if (abs_gap > max_gap) continue;
			const float penalty = abs_gap == 0 ? 0.f : (float)((abs_gap >> 3) + 3);
			uint32_t weight;
			if (dis_read == dis_ref) weight = v[b].cov - (uint32_t)(1 - dis_read > 0 ? 1 - dis_read : 0);
			else if (dis_read > 0 && dis_ref > 0) weight = v[b].cov;
			else if (dis_read >= -5 && dis_read <= 0 && dis_ref >= -5) weight = v[b].cov + (uint32_t)(dis_read < dis_ref ? dis_read : dis_ref);
			else continue;
			// dynamic_programming_path: per node, edges in insertion order (ascending source), `cur <= temp` so the later edge wins ties
			const float temp = dist_a + (float)(int)weight - penalty;
			if (pre[b] == -2) { dist[b] = 0.f; pre[b] = -1; }
			if (dist[b] <= temp) { dist[b] = temp; pre[b] = (int32_t)a; }
		}
	}
	for (uint32_t i = 0; i < n; ++i) if (pre[i] == -2) pre[i] = -1;
}

// ---------------------------------------------------------------------------------------------------------- stage D
struct RefView { const uint64_t *ref_seq; };                       // 2-bit anchor reference, 32 bases per word (ref.seq)
SEED_HD uint32_t ref_base(const RefView &rf, uint32_t i) { return (uint32_t)(rf.ref_seq[i >> 5] >> ((31 - (i & 0x1f)) << 1)) & 3u; }

// Where plan_read puts what it finds.  COUNT only sizes the outputs (first pass), FILL writes them (second pass).
struct PlanSink {
	bool fill;
	// running counts of this read
	uint32_t n_cand, n_piece, n_task, q_bytes, t_bytes, cig_cap;
	// FILL: output arrays, positioned at this read's first slot
	DevCand *cand; DevPiece *piece; int32_t *task_qlen, *task_tlen; int64_t *task_qoff, *task_toff; uint8_t *q, *t;
	uint32_t piece_base, task_base, cig_base; int64_t q_base, t_base;    // global offsets of those first slots
	uint32_t ksw_cig_cap;
};

struct ReadPlanner {
	const AlnScores &o; const RefView &rf; const uint64_t *bits; uint32_t read_l; PlanSink &S;
	int fixed_score, total_q_len; bool last_simple;
	uint32_t cand_piece0, cand_cig_cap;
	SEED_HD ReadPlanner(const AlnScores &o_, const RefView &rf_, const uint64_t *bits_, uint32_t L, PlanSink &s)
		: o(o_), rf(rf_), bits(bits_), read_l(L), S(s), fixed_score(0), total_q_len(0), last_simple(false), cand_piece0(0), cand_cig_cap(0) {}
	SEED_HD void lit(uint8_t type, int size)
	{
		if (S.fill) { DevPiece p; p.task = -1; p.kind = 0; p.type = 0; p.lit_type = type; p.pad = 0; p.lit_size = (int16_t)(uint16_t)size; p.pad2 = 0; S.piece[S.n_piece] = p; }
		++S.n_piece; ++S.cig_cap;
	}
	SEED_HD void lit_bin(uint32_t w) { lit((uint8_t)(w & 0xf), (int)(int16_t)(w >> 4)); }
	SEED_HD int mismatch(int rs, int re, int fs, int fe)           // get_misMatch, RR:893-908
	{
		int qlen = re - rs, tlen = fe - fs;
		if (fe < fs) { tlen = 0; qlen += fs - fe; }
		int nm = 0;
		for (int i = 0; i < qlen; ++i) nm += (i < tlen ? packed_base(bits, (uint32_t)(rs + i)) != ref_base(rf, (uint32_t)(fs + i)) : 1);
		return nm > 3 ? 3 : nm;
	}
	SEED_HD void alignment(int rs, int re, int fs, int fe, int type)   // KSW_ALN_handler::alignment, RR:910-986
	{
		int qlen = re - rs, tlen = fe - fs;
		if (fe < fs) { tlen = 0; qlen += fs - fe; }
		// LEFT extensions run on both sequences reversed (RR:923-928): position i of the window is base (len-1-i)
		const bool rev = type == ST_ALN_LEFT;
		total_q_len += qlen;
		bool simple = false; uint32_t nm = 0;
		if (qlen == 0 || tlen == 0) { simple = true; nm = (uint32_t)(qlen + tlen); }
		else if (qlen == tlen || type != ST_ALN_E2E) {
			for (int i = 0; i < qlen && nm < 6; ++i) {
				if (i < tlen) {
					const uint32_t qb = packed_base(bits, (uint32_t)(rev ? rs + qlen - 1 - i : rs + i));
					const uint32_t tb = ref_base(rf, (uint32_t)(rev ? fs + tlen - 1 - i : fs + i));
					nm += qb != tb;
				} else nm += 1;
			}
			if (nm == 1 || (nm < 6 && (int)(nm << 3) < qlen)) simple = true;
		}
		last_simple = simple;
		if (simple) {
			if (qlen == 0 || tlen == 0) {
				if (nm != 0) {
					const int a = o.gap_open + ((int)nm - 1) * o.gap_ex, b = o.gap_open2 + ((int)nm - 1) * o.gap_ex2;
					fixed_score -= a < b ? a : b;
				}
			} else fixed_score += qlen * o.match - (int)nm * (o.match + o.mismatch);
			if (qlen == 0) lit(2, tlen); else if (tlen == 0) lit(1, qlen); else lit(0, qlen);
			if (fe < fs) lit(2, fe - fs);
		} else if ((int64_t)tlen * qlen > 1000000) {               // align_non_splice guard, RR:874-887
			fixed_score += type == ST_ALN_E2E ? 0 : ST_NEG_INF;
			const uint32_t c0 = (uint32_t)qlen << 4 | 1, c1 = (uint32_t)tlen << 4 | 3;
			lit_bin(type == ST_ALN_LEFT ? c0 : c1); lit_bin(type == ST_ALN_LEFT ? c1 : c0);
		} else {
			if (S.fill) {
				DevPiece p; p.task = (int32_t)(S.task_base + S.n_task); p.kind = 1; p.type = (uint8_t)type; p.lit_type = 0; p.pad = 0; p.lit_size = 0; p.pad2 = 0;
				S.piece[S.n_piece] = p;
				S.task_qlen[S.n_task] = qlen; S.task_tlen[S.n_task] = tlen;
				S.task_qoff[S.n_task] = S.q_base + S.q_bytes; S.task_toff[S.n_task] = S.t_base + S.t_bytes;
				uint8_t *q = S.q + S.q_bytes, *t = S.t + S.t_bytes;
				for (int i = 0; i < qlen; ++i) q[i] = (uint8_t)packed_base(bits, (uint32_t)(rev ? rs + qlen - 1 - i : rs + i));
				for (int i = 0; i < tlen; ++i) t[i] = (uint8_t)ref_base(rf, (uint32_t)(rev ? fs + tlen - 1 - i : fs + i));
			}
			++S.n_piece; ++S.n_task; S.q_bytes += (uint32_t)qlen; S.t_bytes += (uint32_t)tlen; S.cig_cap += S.ksw_cig_cap;
		}
	}
	// get_ksw_score from chain end `first_node` (RR:308-400)
	SEED_HD void run(const DevSeed *v, const int32_t *pre, int first_node, int *rba_out)
	{
		const int I32MAXV = 0x7fffffff;
		int aln_read_begin = (int)read_l, aln_read_end = (int)read_l, aln_ref_begin = I32MAXV, aln_ref_end = I32MAXV;
		int last_aln_begin = (int)read_l, last_ref_begin = I32MAXV, unitig_mis = 0;
		for (int node = first_node; node != -1;) {
			const DevSeed &s = v[node];
			const int m_rb = (int)s.read_begin, m_re = (int)s.read_end, m_fb = (int)s.ref_begin, m_fe = (int)s.ref_end;
			aln_read_begin = aln_read_begin < m_re ? aln_read_begin : m_re;
			aln_ref_begin = aln_ref_begin < m_fe ? aln_ref_begin : m_fe;
			if (aln_read_begin <= aln_read_end) {
				if (aln_read_end < last_aln_begin) {
					const int mem_len = last_aln_begin - aln_read_end;
					unitig_mis += mismatch(aln_read_end, aln_read_end + mem_len, last_ref_begin, last_ref_begin + mem_len);
					lit(0, mem_len);
				}
				last_aln_begin = aln_read_begin;
				if (aln_ref_end == I32MAXV) {
					aln_ref_end = aln_ref_begin + (aln_read_end - aln_read_begin) + 30;
					alignment(aln_read_begin, aln_read_end, aln_ref_begin, aln_ref_end, ST_ALN_RIGHT);
				} else alignment(aln_read_begin, aln_read_end, aln_ref_begin, aln_ref_end, ST_ALN_E2E);
			} else {
				const int d_read = aln_read_end - aln_read_begin, d_ref = aln_ref_end - aln_ref_begin;
				if (d_read != d_ref) {
					const int del = d_ref > d_read ? d_ref - d_read : d_read - d_ref;
					const int a = o.gap_open + (del - 1) * o.gap_ex, b = o.gap_open2 + (del - 1) * o.gap_ex2;
					fixed_score -= a < b ? a : b;
				}
			}
			aln_read_end = m_rb; last_ref_begin = m_fb; aln_ref_end = m_fb;
			const int next = pre[node];
			if (next == -1) break;
			node = next;
		}
		if (aln_read_end < last_aln_begin) {
			const int mem_len = last_aln_begin - aln_read_end;
			unitig_mis += mismatch(aln_read_end, aln_read_end + mem_len, last_ref_begin, last_ref_begin + mem_len);
			lit(0, mem_len);
		}
		aln_read_begin = 0; aln_ref_begin = 0;
		int rba = 0;
		if (aln_read_begin < aln_read_end) {
			aln_ref_begin = aln_ref_end - (aln_read_end - aln_read_begin) - 30;
			if (aln_ref_begin < 0) aln_ref_begin = 0;
			alignment(aln_read_begin, aln_read_end, aln_ref_begin, aln_ref_end, ST_ALN_LEFT);
			if (aln_ref_end > aln_ref_begin) rba = last_simple ? aln_ref_end - aln_ref_begin - 30 : aln_ref_end - aln_ref_begin;
		}
		fixed_score += ((int)read_l - total_q_len) * o.match;
		fixed_score -= unitig_mis * (o.match + o.mismatch);
		*rba_out = rba;
	}
};

// All candidates of one read: every chain end with dist >= 30 and dist + 30 >= the best chain of both strands (the ends
// sort_output can still pick, RR:423-429, 442), strand 0 first, nodes ascending.
SEED_HD void plan_read(const AlnScores &o, const RefView &rf, uint32_t read_index, const uint64_t *bits_fw, uint32_t read_l,
                       const DevSeed *const v[2], const float *const dist[2], const int32_t *const pre[2], const uint32_t n[2], PlanSink &S)
{
	uint32_t best = 0;
	for (int s = 0; s < 2; ++s) for (uint32_t i = 0; i < n[s]; ++i) { const uint32_t c = (uint32_t)dist[s][i]; if (c > best) best = c; }
	if (best < (uint32_t)ST_MIN_CHAIN_SCORE) return;
	const uint32_t words = (read_l >> 5) + 2;
	for (int s = 0; s < 2; ++s) {
		for (uint32_t node = 0; node < n[s]; ++node) {
			const uint32_t c = (uint32_t)dist[s][node];
			if (c < (uint32_t)ST_MIN_CHAIN_SCORE2 || c + ST_MAX_CHAIN_SCORE_DIFF < best) continue;
			ReadPlanner pl(o, rf, bits_fw + (size_t)s * words, read_l, S);
			const uint32_t piece0 = S.n_piece, cig0 = S.cig_cap;
			int rba = 0;
			pl.run(v[s], pre[s], (int)node, &rba);
			if (S.fill) {
				DevCand cd;
				cd.read = read_index; cd.node = node; cd.strand = (uint32_t)s; cd.fixed_score = pl.fixed_score; cd.read_begin_alignment = rba;
				cd.piece_off = S.piece_base + piece0; cd.n_pieces = S.n_piece - piece0;
				cd.cig_off = S.cig_base + cig0; cd.cig_cap = S.cig_cap - cig0;
				cd.align_score = 0; cd.n_cig = 0; cd.cigar_ok = 0;
				S.cand[S.n_cand] = cd;
			}
			++S.n_cand;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------- stage F1
SEED_HD bool cigar_try_merge(DevCigar &a, const DevCigar &b)       // CIGAR_PATH::try_merge, RRH:159-178
{
	if (b.size < 0) {
		if (a.type == 0) { a.size = (int16_t)(a.size + b.size); return true; }
		if (a.type == 2) { a.size = (int16_t)(a.size - b.size); return true; }
		return true;
	}
	if (a.type == b.type || b.size == 0) { a.size = (int16_t)(a.size + b.size); return true; }
	return false;
}

// Score and final CIGAR of one candidate once its ksw tasks have results (res: 12 words per task, cig: cap words per task).
// The pieces were emitted from the read's 3' end to its 5' end; reverseGIGAR walks them backwards and merges neighbours.
// Returns false when a task's CIGAR did not fit `cap` words (the batch is then run again with room).
SEED_HD bool resolve_cand(DevCand &cd, const DevPiece *pieces, const int32_t *res, const uint32_t *cig, int cap, uint32_t read_l, DevCigar *out_pool)
{
	int score = cd.fixed_score;
	DevCigar *out = out_pool + cd.cig_off;
	uint32_t n_out = 0;
	bool overflow = false;
	auto feed = [&](DevCigar c) {
		if (n_out == 0) { out[n_out++] = c; return; }
		if (!cigar_try_merge(out[n_out - 1], c)) out[n_out++] = c;
	};
	const DevPiece *pp = pieces + cd.piece_off;
	for (int k = (int)cd.n_pieces - 1; k >= 0; --k) {
		const DevPiece &p = pp[k];
		if (p.kind == 0) { DevCigar c; c.type = p.lit_type; c.pad = 0; c.size = p.lit_size; feed(c); continue; }
		const int32_t *r = res + (size_t)p.task * 12;
		const uint32_t *cg = cig + (size_t)p.task * cap;
		const int nc = r[9];
		if ((r[11] & 1) || nc > cap) { overflow = true; continue; }
		score += p.type == ST_ALN_E2E ? r[8] : r[4];
		// emission order was: E2E and RIGHT back to front, LEFT front to back; read backwards here
		if (p.type == ST_ALN_LEFT) for (int i = nc - 1; i >= 0; --i) { DevCigar c; c.type = (uint8_t)(cg[i] & 0xf); c.pad = 0; c.size = (int16_t)(cg[i] >> 4); feed(c); }
		else for (int i = 0; i < nc; ++i) { DevCigar c; c.type = (uint8_t)(cg[i] & 0xf); c.pad = 0; c.size = (int16_t)(cg[i] >> 4); feed(c); }
	}
	if (overflow) return false;
	uint32_t first = 0;
	if (n_out > 0 && out[0].size == 0) first = 1;                  // drop a leading 0-length entry
	if (first) for (uint32_t i = 1; i < n_out; ++i) out[i - 1] = out[i];
	n_out -= first;
	int total = 0;
	for (uint32_t i = 0; i < n_out; ++i) if (out[i].type == 0 || out[i].type == 1 || out[i].type == 3 || out[i].type == 4) total += out[i].size;
	cd.align_score = (uint32_t)(score > 0 ? score : 0);
	cd.n_cig = n_out;
	cd.cigar_ok = (n_out > 0 && total == (int)read_l) ? 1u : 0u;
	return true;
}

} // namespace pansvr
