// stages_run.hpp -- one block of read states through the device stages A..F1 (stages_core.cuh), written once against a small
// backend interface: the CUDA backend of the product (stages_gpu.cu: kernels on a stream, buffers in HBM, cub scans) and the
// host backend of the CPU-only test build (tests/emul: the same functors stepped by a loop).  Everything between the FASTQ
// text and the candidate alignments stays in device memory; the host gets back what stage F needs.
//
//   text + read table --H2D--> A encode/census -> B seed count | scan | fill -> (MEMs --D2H--> for reads the host keeps)
//     -> C merge | scan | expand + chain -> D plan count | scans | fill -> E ksw (ksw_team kernels) -> F1 resolve
//     --D2H--> per strand: sorted seeds + chain table; per read: candidates with score and final CIGAR
//
// Backend concept:
//   template <class T> T *buf(int slot, size_t n)        grow-only device buffer `slot`, at least n elements
//   void h2d(void *d, const void *h, size_t bytes)  /  void d2h(void *h, const void *d, size_t bytes)   (stream order)
//   void zero(void *d, size_t bytes)
//   void sync()                                           wait for everything issued so far
//   template <class F> void for_each(size_t n, const F &f, int stage)    f(i) for i in [0, n); `stage` labels the timing
//   template <class F> void for_each_in(size_t n, const uint32_t *perm, const F &f, int stage)    f(perm[k]) for k in [0, n)
//   void order_desc(const uint32_t *key, const uint32_t *idx, uint32_t *perm, size_t n)   perm = idx sorted by key (< 256), largest first
//   void encode(size_t n, const FnEncode &f)                             f.run(i, scratch, stride) for i in [0, n) (stage A; the backend owns the scratch)
//   void scan(const uint32_t *in, uint32_t *out, size_t n)               exclusive prefix sums, n elements
//   bool ksw(n, q, qoff, qlen, t, toff, tlen, res, cig, cap, err)        ksw_extd2 batch (w=200, the stage's scoring) over device arrays
#pragma once
#include <string>

#include "pipeline.hpp"
#include "stages_core.cuh"
#include "pair_core.cuh"
#include "text_core.cuh"

namespace pansvr {

enum DevSlot {
	SL_TEXT, SL_READS, SL_BITS, SL_LIST, SL_FLAGS, SL_MEM_CNT, SL_MEM_OFF, SL_MEMS, SL_MEMS_TMP, SL_NVU, SL_SEED_CNT, SL_SEED_OFF,
	SL_SEEDS, SL_SEEDS_TMP, SL_DIST, SL_PRE, SL_PLAN_CNT, SL_PLAN_OFF, SL_CANDS, SL_PIECES, SL_QLEN, SL_TLEN, SL_QOFF, SL_TOFF,
	SL_Q, SL_T, SL_RES, SL_KCIG, SL_CIGS, SL_MISC, SL_SCAN_TMP, SL_USED, SL_ORI, SL_PSTATE, SL_PROBE, SL_WIN, SL_FINAL, SL_PFINAL, SL_RECS, SL_HOSTLEN, SL_TXT_LEN, SL_TXT_OFF, SL_TXT, SL_NL_CNT, SL_NL_OFF, SL_LINES, SL_LAY_CNT, SL_LAY_OFF, SL_SEL_PAIR, SL_SEL_FIN, SL_SEL_PFIN, SL_DRAW_CNT, SL_DRAW_OFF, SL_REDO, SL_DRAWN, SL_MEMS_KEPT, SL_WORK_KEY, SL_WORK_IDX, SL_WORK_PERM, SL_SORT_KEYS, SL_SORT_TMP, SL_TIE_CNT, SL_TIE_OFF, SL_TIE, SL_TIE_SEEDS, SL_TIE_DIST, SL_TIE_PRE, SL_TIE_CANDS, SL_TIE_PAIR, SL_TIE_DONE, SL_COUNT
};

// ---- functors (one element of work each; plain data members only, so they can be passed to a kernel by value)
enum { NL_PIECE = 64 };
// bit 7 of every byte of w that is '\n' (exact: no borrow runs into a neighbour)
SEED_HD uint32_t newline_bits(uint32_t w)
{
	const uint32_t x = w ^ 0x0a0a0a0au;
	return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}
SEED_HD uint32_t popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
	return (uint32_t)__popc(x);
#else
	return (uint32_t)__builtin_popcount(x);
#endif
}
struct alignas(16) TextWords { uint32_t w[4]; };
struct FnLines {                                                   // line ends of one 64-byte piece of the text: counted, then placed
	const uint8_t *text; uint32_t bytes; uint32_t *cnt; const uint32_t *off; uint32_t *lines; uint32_t max_lines; bool fill;
	SEED_HD void operator()(size_t k) const
	{
		const uint32_t b = (uint32_t)k * NL_PIECE, e = b + NL_PIECE < bytes ? b + NL_PIECE : bytes;
		uint32_t c = 0;
		if (e - b == NL_PIECE) {                                    // a whole piece: four 16-byte loads, line ends found a word at a time
			TextWords q[4];
			for (int i = 0; i < 4; ++i) q[i] = ((const TextWords*)(text + b))[i];
			if (!fill) {
				for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c += popc32(newline_bits(q[i].w[j]));
				cnt[k] = c;
				return;
			}
			uint32_t at = off[k];
			for (int i = 0; i < 4; ++i)
				for (int j = 0; j < 4; ++j)
					for (uint32_t m = newline_bits(q[i].w[j]); m; m &= m - 1) {
						if (at < max_lines) lines[at] = b + 16 * i + 4 * j + ((uint32_t)ctz64(m) >> 3);
						++at;
					}
			return;
		}
		if (!fill) { for (uint32_t i = b; i < e; ++i) c += text[i] == '\n'; cnt[k] = c; return; }
		uint32_t at = off[k];
		for (uint32_t i = b; i < e; ++i) if (text[i] == '\n') { if (at < max_lines) lines[at] = i; ++at; }
	}
};
struct FnRecord {                                                  // record r = lines 4r .. 4r+3: "@name comment", bases, "+...", qualities
	const uint8_t *text; const uint32_t *lines; DevRec *recs; uint32_t *words, *list; uint32_t *bad;
	SEED_HD void operator()(size_t r) const
	{
		uint32_t s[4], e[4];
		for (int k = 0; k < 4; ++k) { const size_t l = 4 * r + k; s[k] = l == 0 ? 0u : lines[l - 1] + 1; e[k] = lines[l]; }
		if (e[0] == s[0] || (text[s[0]] != '@' && text[s[0]] != '>') || e[2] == s[2] || text[s[2]] != '+') *bad = 1;
		uint32_t sp = s[0] + 1;
		while (sp < e[0] && text[sp] != ' ' && text[sp] != '\t') ++sp;
		DevRec d;
		d.name_off = s[0] + 1; d.name_l = e[0] > s[0] ? sp - (s[0] + 1) : 0;
		while (sp < e[0] && (text[sp] == ' ' || text[sp] == '\t')) ++sp;
		d.comment_off = sp; d.comment_l = e[0] > sp ? e[0] - sp : 0;
		d.seq_off = s[1]; d.seq_l = e[1] - s[1]; d.qual_off = s[3]; d.qual_l = e[3] - s[3];
		recs[r] = d;
		words[r] = d.seq_l >= LEN_KMER ? 2 * ((d.seq_l >> 5) + 2) : 0;
		list[r] = d.seq_l >= LEN_KMER ? 2 * (d.seq_l - LEN_KMER + 1) : 0;
	}
};
struct FnReadTable {
	const DevRec *recs; const uint32_t *words_off, *list_off; DevRead *reads;
	SEED_HD void operator()(size_t i) const { DevRead d; d.seq_off = recs[i].seq_off; d.len = recs[i].seq_l; d.var_code = 0; d.bits_off = words_off[i]; d.list_off = list_off[i]; reads[i] = d; }
};
struct FnOri {                                                     // original alignment of every read, out of its comment
	const uint8_t *text; const DevRec *recs; int match; DevOri *ori;
	SEED_HD void operator()(size_t i) const { dev_parse_ori((const char*)text + recs[i].comment_off, recs[i].comment_l, recs[i].seq_l, match, ori[i]); }
};
struct FnEncode {
	const uint8_t *text; const DevRead *reads; const DevOri *ori; uint64_t *bits; uint8_t *list; uint8_t *flags;
	// filter: the census's scratch (stages_core.cuh: encode_read), word k at filter[k * fstride]
	SEED_HD void run(size_t i, uint32_t *filter, uint32_t fstride) const
	{
		const DevRead &rd = reads[i];
		uint8_t f = 0;
		if (rd.len < LEN_KMER || (ori && ori[i].skip)) {
			// not seeded: skipped by RR:413-414 (full-score original alignment), or shorter than a k-mer -- unless the host path keeps
			// the pair: 'N' (its substitution draws rand()) or a lower-case 'n' (code 4 spills into the packed neighbour)
			f = ST_FLAG_NOSEED;
			for (uint32_t k = 0; k < rd.len; ++k) { const uint8_t c = text[rd.seq_off + k]; if (c == 'N' || c == 'n') { f = ST_FLAG_HOST; break; } }
		} else {
			const int r = encode_read(text, rd, bits, list, filter, fstride);
			f = r == ENC_HAS_N ? (uint8_t)ST_FLAG_HOST : r == ENC_STR ? (uint8_t)ST_FLAG_STR : (uint8_t)0;
		}
		flags[i] = f;
	}
	SEED_HD void operator()(size_t i) const { uint32_t filter[ENC_FILTER_WORDS]; run(i, filter, 1); }
};
enum { SEED_TMP_CAP = 8 };                                         // MEMs per strand kept by the counting pass (most strands have fewer)
struct FnSeed {                                                    // strand j = 2 * read + strand: its MEMs counted, the first SEED_TMP_CAP of them kept in tmp
	IndexView ix; const DevRead *reads; const uint64_t *bits; const uint8_t *list; const uint8_t *flags;
	uint32_t *count; Mem *tmp; unsigned long long *probes;
	SEED_HD void operator()(size_t j) const
	{
		const DevRead &rd = reads[j >> 1];
		if (flags[j >> 1] & (ST_FLAG_HOST | ST_FLAG_NOSEED)) { count[j] = 0; return; }
		const uint32_t s = (uint32_t)(j & 1), words = (rd.len >> 5) + 2, kn = rd.len - LEN_KMER + 1;
		const uint64_t *b = bits + rd.bits_off + (size_t)s * words;
		const bool is_str = (flags[j >> 1] & ST_FLAG_STR) != 0;
		const uint8_t *sl = list + rd.list_off + (size_t)s * kn;
		uint32_t p = 0;
		count[j] = (uint32_t)seed_read_strand(ix, b, rd.len, is_str, sl, tmp + j * SEED_TMP_CAP, SEED_TMP_CAP, &p);
#if defined(__CUDA_ARCH__)
		if (p) atomicAdd(probes, (unsigned long long)p);
#else
		*probes += p;
#endif
	}
};
struct FnSeedPlace {                                               // the strand's MEMs to their place in the dense list: copied, or found again if there were more than kept
	IndexView ix; const DevRead *reads; const uint64_t *bits; const uint8_t *list; const uint8_t *flags;
	const uint32_t *off; const Mem *tmp; Mem *mems;
	SEED_HD void operator()(size_t j) const
	{
		const uint32_t n = off[j + 1] - off[j];
		if (n == 0) return;
		if (n <= (uint32_t)SEED_TMP_CAP) { for (uint32_t k = 0; k < n; ++k) mems[off[j] + k] = tmp[j * SEED_TMP_CAP + k]; return; }
		const DevRead &rd = reads[j >> 1];
		const uint32_t s = (uint32_t)(j & 1), words = (rd.len >> 5) + 2, kn = rd.len - LEN_KMER + 1;
		seed_read_strand(ix, bits + rd.bits_off + (size_t)s * words, rd.len, (flags[j >> 1] & ST_FLAG_STR) != 0, list + rd.list_off + (size_t)s * kn, mems + off[j], (int)n);
	}
};
struct FnMerge {                                                   // per read: both strands
	const uint32_t *mem_off; Mem *mems, *tmp; uint8_t *flags; uint32_t *nvu, *seed_cnt;
	SEED_HD void operator()(size_t i) const
	{
		const uint32_t b = mem_off[2 * i], e = mem_off[2 * i + 2];
		bool needs_rand = false;
		for (uint32_t k = b; k < e; ++k) needs_rand |= mems[k].pos_n > (uint32_t)ST_POS_N_MAX;
		if (needs_rand) {                                           // expand_seed would draw from random_r: the host keeps this read
			flags[i] |= (uint8_t)ST_FLAG_NEEDS_RAND;
			nvu[2 * i] = nvu[2 * i + 1] = 0; seed_cnt[2 * i] = seed_cnt[2 * i + 1] = 0;
			return;
		}
		for (int s = 0; s < 2; ++s) {
			const uint32_t mb = mem_off[2 * i + s], me = mem_off[2 * i + s + 1];
			uint32_t ns = 0;
			nvu[2 * i + s] = merge_strand(mems + mb, tmp + mb, me - mb, &ns);
			seed_cnt[2 * i + s] = ns;
		}
	}
};
struct FnChain {
	const uint32_t *mem_off, *nvu, *seed_off; const Mem *mems; const uint64_t *pos, *posp; const uint8_t *flags;
	DevSeed *seeds, *tmp; float *dist; int32_t *pre;
	SEED_HD void operator()(size_t i) const
	{
		for (int s = 0; s < 2; ++s) {
			const uint32_t sb = seed_off[2 * i + s], n = seed_off[2 * i + s + 1] - sb;
			if (n == 0) continue;
			chain_strand((const DevVertex*)(mems + mem_off[2 * i + s]), nvu[2 * i + s], pos, posp, (flags[i] & ST_FLAG_STR) != 0,
			             seeds + sb, tmp + sb, n, dist + sb, pre + sb);
		}
	}
};
struct FnWorkKey {                                                 // how much the per-read stages behind chaining have to do for read i: its seeds (at most 255)
	const uint32_t *seed_off; uint32_t *key, *idx;
	SEED_HD void operator()(size_t i) const { const uint32_t n = seed_off[2 * i + 2] - seed_off[2 * i]; key[i] = n < 255u ? n : 255u; idx[i] = (uint32_t)i; }
};
enum { PLAN_FIELDS = 6 };                                          // cands, pieces, tasks, q bytes, t bytes, CIGAR room
struct FnPlan {
	AlnScores o; RefView rf; const DevRead *reads; const uint64_t *bits; const uint32_t *seed_off; const DevSeed *seeds; const float *dist; const int32_t *pre;
	size_t n_reads; uint32_t *cnt; const uint32_t *off;              // cnt / off: PLAN_FIELDS planes of (n_reads + 1)
	DevCand *cands; DevPiece *pieces; int32_t *qlen, *tlen; int64_t *qoff, *toff; uint8_t *q, *t; uint32_t ksw_cig_cap; bool fill;
	SEED_HD void operator()(size_t i) const
	{
		const DevRead &rd = reads[i];
		const DevSeed *v[2]; const float *d[2]; const int32_t *p[2]; uint32_t n[2];
		for (int s = 0; s < 2; ++s) {
			const uint32_t sb = seed_off[2 * i + s];
			v[s] = seeds + sb; d[s] = dist + sb; p[s] = pre + sb; n[s] = seed_off[2 * i + s + 1] - sb;
		}
		PlanSink S;
		S.fill = fill; S.n_cand = S.n_piece = S.n_task = S.q_bytes = S.t_bytes = S.cig_cap = 0; S.ksw_cig_cap = ksw_cig_cap;
		S.cand = nullptr; S.piece = nullptr; S.task_qlen = S.task_tlen = nullptr; S.task_qoff = S.task_toff = nullptr; S.q = S.t = nullptr;
		S.piece_base = S.task_base = S.cig_base = 0; S.q_base = S.t_base = 0;
		const size_t P = n_reads + 1;
		if (fill) {
			const uint32_t c0 = off[i], p0 = off[P + i], k0 = off[2 * P + i], q0 = off[3 * P + i], t0 = off[4 * P + i], g0 = off[5 * P + i];
			S.cand = cands + c0; S.piece = pieces + p0; S.task_qlen = qlen + k0; S.task_tlen = tlen + k0; S.task_qoff = qoff + k0; S.task_toff = toff + k0;
			S.q = q + q0; S.t = t + t0; S.piece_base = p0; S.task_base = k0; S.cig_base = g0; S.q_base = (int64_t)q0; S.t_base = (int64_t)t0;
		}
		if (n[0] + n[1] != 0) plan_read(o, rf, (uint32_t)i, bits + rd.bits_off, rd.len, v, d, p, n, S);
		if (!fill) { cnt[i] = S.n_cand; cnt[P + i] = S.n_piece; cnt[2 * P + i] = S.n_task; cnt[3 * P + i] = S.q_bytes; cnt[4 * P + i] = S.t_bytes; cnt[5 * P + i] = S.cig_cap; }
	}
};
struct FnResolve {
	DevCand *cands; const DevPiece *pieces; const int32_t *res; const uint32_t *kcig; int cap; const DevRead *reads; DevCigar *cigs; uint32_t *overflow;
	SEED_HD void operator()(size_t c) const
	{
		if (!resolve_cand(cands[c], pieces, res, kcig, cap, reads[cands[c].read].len, cigs)) *overflow = 1;     // (benign race: every writer stores 1)
	}
};

struct FnExplore {                                                 // one read: chain selection and candidate sort under every outcome of its ties
	PairIndexView ix; const uint8_t *flags; const uint32_t *seed_off; const DevSeed *seeds; const float *dist; const int32_t *pre; uint8_t *used;
	const uint32_t *cand_off; const DevCand *cands; const DevOri *ori; DevPairState *state;
	SEED_HD void operator()(size_t i) const
	{
		const size_t p = i >> 1;
		if ((flags[2 * p] | flags[2 * p + 1]) & (ST_FLAG_NEEDS_RAND | ST_FLAG_HOST)) return;      // the host path keeps the pair
		ReadView R;
		for (int s = 0; s < 2; ++s) {
			const uint32_t sb = seed_off[2 * i + s];
			R.v[s] = seeds + sb; R.dist[s] = dist + sb; R.pre[s] = pre + sb; R.used[s] = used + sb; R.n[s] = seed_off[2 * i + s + 1] - sb;
		}
		R.cands = cands; R.cand_b = cand_off[i]; R.cand_e = cand_off[i + 1]; R.ori = ori[i];
		dev_explore_store(ix, R, state[p], (int)(i & 1));
	}
};
struct FnProbe {                                                   // one pair = reads 2p, 2p + 1 of the table: pairing
	PairIndexView ix; PairOpts o; const uint8_t *flags; const DevOri *ori; DevPairState *state; DevProbe *probe; uint32_t *draw_cnt; uint8_t *redo;
	SEED_HD void operator()(size_t p) const
	{
		DevProbe &pr = probe[p];
		if ((flags[2 * p] | flags[2 * p + 1]) & (ST_FLAG_NEEDS_RAND | ST_FLAG_HOST)) { pr.redo = PR_REDO_HOST; pr.draws0 = pr.draws1 = pr.ev_cnt = 0; pr.tie_mask = 0; }
		else dev_probe_pair(ix, o, ori + 2 * p, state[p], pr);
		draw_cnt[p] = dev_pair_draws(pr);                           // how far the pair advances the stream (0: not at all, or the host path's business)
		redo[p] = pr.redo;
	}
};
// The pairs whose ties decide their outcome (PR_REDO_TIES) are finished by the in-order pass on the host, from what the device
// has for them: a pair's seeds (with the chain tables) and its candidates are contiguous ranges, copied into packed arrays that go
// down with the rest of the first trip's results.
struct DevTie { uint32_t pair, seed_at, cand_at; uint32_t seed_off[5], cand_off[3]; DevOri ori[2]; };   // seed_off / cand_off: the pair's global offsets (4 strands / 2 reads)
enum { TIE_FIELDS = 3 };                                           // pairs, seeds, candidates
struct FnTieCount {
	const uint8_t *redo; const uint32_t *seed_off, *cand_off; size_t n_pairs; uint32_t *cnt;   // cnt: TIE_FIELDS planes of (n_pairs + 1)
	SEED_HD void operator()(size_t p) const
	{
		const bool tie = redo[p] == PR_REDO_TIES;
		const size_t P = n_pairs + 1;
		cnt[p] = tie ? 1u : 0u;
		cnt[P + p] = tie ? seed_off[4 * p + 4] - seed_off[4 * p] : 0u;
		cnt[2 * P + p] = tie ? cand_off[2 * p + 2] - cand_off[2 * p] : 0u;
	}
};
struct FnTieGather {
	const uint8_t *redo; const uint32_t *seed_off, *cand_off, *off; size_t n_pairs;
	const DevSeed *seeds; const float *dist; const int32_t *pre; const DevCand *cands; const DevOri *ori;
	DevTie *tie; DevSeed *t_seeds; float *t_dist; int32_t *t_pre; DevCand *t_cands;
	SEED_HD void operator()(size_t p) const
	{
		if (redo[p] != PR_REDO_TIES) return;
		const size_t P = n_pairs + 1;
		DevTie t;
		t.pair = (uint32_t)p; t.seed_at = off[P + p]; t.cand_at = off[2 * P + p];
		for (int k = 0; k < 5; ++k) t.seed_off[k] = seed_off[4 * p + k];
		for (int k = 0; k < 3; ++k) t.cand_off[k] = cand_off[2 * p + k];
		t.ori[0] = ori[2 * p]; t.ori[1] = ori[2 * p + 1];
		tie[off[p]] = t;
		for (uint32_t k = t.seed_off[0], d = t.seed_at; k < t.seed_off[4]; ++k, ++d) { t_seeds[d] = seeds[k]; t_dist[d] = dist[k]; t_pre[d] = pre[k]; }
		for (uint32_t k = t.cand_off[0], d = t.cand_at; k < t.cand_off[2]; ++k, ++d) t_cands[d] = cands[k];
	}
};
struct FnTieScatter {                                              // second trip: the pairs finished in order get their state, and count as decided
	const uint32_t *pair; const DevPairState *done; DevPairState *state; DevProbe *probe;
	SEED_HD void operator()(size_t k) const { state[pair[k]] = done[k]; probe[pair[k]].redo = 0; }
};
struct FnFinalize {
	PairIndexView ix; PairOpts o; const DevOri *ori; DevPairState *state; const DevProbe *probe; const uint32_t *draw_off; const int32_t *drawn; const DevCand *cands; const DevCigar *cigs;
	DevFinal *fin; DevPairFinal *pfin;
	SEED_HD void operator()(size_t p) const { dev_finalize_pair(ix, o, ori + 2 * p, state[p], probe[p], drawn + draw_off[p], cands, cigs, fin + 2 * p, pfin[p]); }
};
struct FnOriSelect {                                               // the (few) pairs whose originals may go to the `-p` output: their results go back to the host
	PairOpts o; const DevOri *ori; const DevFinal *fin; const DevPairFinal *pfin; uint32_t *count; uint32_t cap; uint32_t *sel_pair; DevFinal *sel_fin; DevPairFinal *sel_pfin;
	SEED_HD void operator()(size_t p) const
	{
		if (!pfin[p].valid || pfin[p].max_score > o.min_filter_score || ori[2 * p].chr == PR_U32MAX || ori[2 * p + 1].chr == PR_U32MAX) return;
#if defined(__CUDA_ARCH__)
		const uint32_t k = atomicAdd(count, 1u);
#else
		const uint32_t k = (*count)++;
#endif
		if (k >= cap) return;
		sel_pair[k] = (uint32_t)p; sel_fin[2 * k] = fin[2 * p]; sel_fin[2 * k + 1] = fin[2 * p + 1]; sel_pfin[k] = pfin[p];
	}
};

struct FnCells {                                                   // in-band DP cells of the tasks (the unit GCUPS is quoted in, KSW:131-138)
	const int32_t *qlen, *tlen; int w; unsigned long long *cells;
	SEED_HD void operator()(size_t k) const
	{
		const int ql = qlen[k], tl = tlen[k];
		if (ql <= 0 || tl <= 0) return;
		unsigned long long c = 0;
		for (int r = 0; r < ql + tl - 1; ++r) {
			int lo = 0, hi = tl - 1;
			if (lo < r - ql + 1) lo = r - ql + 1;
			if (hi > r) hi = r;
			if (lo < ((r - w + 1) >> 1)) lo = (r - w + 1) >> 1;
			if (hi > ((r + w) >> 1)) hi = (r + w) >> 1;
			if (lo > hi) break;
			c += (unsigned long long)(hi - lo + 1);
		}
#if defined(__CUDA_ARCH__)
		atomicAdd(cells, c);
#else
		*cells += c;
#endif
	}
};

struct DevStageIn {
	const uint8_t *text; size_t text_bytes;                        // host: the block's FASTQ text (seq_off of the read table points into it)
	const DevRead *reads; size_t n_reads;                          // host: the read table; bits_off / list_off laid out by the caller
	size_t bits_words, list_bytes;                                 // pool sizes implied by the table
	AlnScores scores;
	// stage F on the device: reads 2p, 2p + 1 of the table are the mates of pair p (n_reads even); recs = their FASTQ records
	// (offsets into `text`): the original alignments are parsed out of the comments, the SAM records written from them
	const DevRec *recs = nullptr; PairOpts pair_opts = {0, 0, 0};
	// ... or parse_text: `text` is strict 4-line FASTQ of n_reads records, and the record and read tables are made from it on the
	// device (reads / recs / bits_words / list_bytes are not read); out.recs brings the record table back, out.parse_ok says whether
	// the text was what it was taken for (if not, nothing else was done: the host parses the block)
	bool parse_text = false;
	bool want_tables = false;                                      // also bring the sorted seeds and chain tables back (tests)
};
struct DevStageOut {                                               // host side, kept across blocks (pinned in the product)
	HostVec<uint8_t> flags;                                        // per read state: ST_FLAG_*
	HostVec<uint32_t> mem_off;                                     // 2 n + 1: MEM counts (the MEMs themselves never leave the device)
	HostVec<uint32_t> seed_off;                                    // 2 n + 1: seeds of strand s of read i = [seed_off[2i+s], seed_off[2i+s+1])
	HostVec<DevSeed> seeds; HostVec<float> dist; HostVec<int32_t> pre;
	HostVec<uint32_t> cand_off;                                    // n + 1
	HostVec<DevCand> cands; HostVec<DevCigar> cigs;
	HostVec<DevProbe> pair_probe;                                  // n / 2: the probe's record of each pair (want_tables only; it stays on the device)
	HostVec<uint8_t> redo;                                         // n / 2: 0 nothing drawn, 1 / 2 advances the stream by a known count, PR_REDO_HOST: the host path's pair
	HostVec<uint32_t> draw_off;                                    // n / 2 + 1: prefix sums of the pairs' draw counts = each pair's place in the block's drawn numbers
	// the pairs the in-order pass finishes itself (PR_REDO_TIES), in input order, with their seeds / chain tables / candidates
	HostVec<DevTie> ties; HostVec<DevSeed> tie_seeds; HostVec<float> tie_dist; HostVec<int32_t> tie_pre; HostVec<DevCand> tie_cands;
	// after run_device_finalize: the results of the pairs that may go to the `-p` output (pair index, both reads, pair) -- or, if
	// sel_all, of every pair (fin: n, pfin: n / 2)
	HostVec<uint32_t> sel_pair; HostVec<DevFinal> sel_fin; HostVec<DevPairFinal> sel_pfin; bool sel_all = false;
	HostVec<DevFinal> fin; HostVec<DevPairFinal> pfin;
	HostVec<DevRec> recs; bool parse_ok = true;                     // parse_text: the record table the device made
	HostVec<uint32_t> txt_off;                                     // n + 1: place of every read's record in the block's SAM text
	uint64_t bad_records = 0;                                      // records left out because htslib would reject them (CIGAR does not span the read)
	std::function<char*(size_t)> text_dest;                        // optional: called once with the size of the block's SAM text, returns where it goes
	char *text_ptr = nullptr; size_t text_total = 0;               // where it went
	uint64_t n_tasks = 0, n_cells = 0, probes = 0;
	DevCounters dev;
};

template <class BE>
bool run_device_stages(BE &be, const IndexView &ix, const uint64_t *d_pos, const RefView &rf, const PairIndexView &pix, const DevStageIn &in, DevStageOut &out, std::string &err)
{
	const size_t n = in.n_reads;
	out.n_tasks = out.n_cells = out.probes = 0;
	out.flags.resize(n); out.mem_off.resize(2 * n + 1); out.seed_off.resize(2 * n + 1); out.cand_off.resize(n + 1);
	out.seeds.clear(); out.dist.clear(); out.pre.clear(); out.cands.clear(); out.cigs.clear(); out.redo.clear(); out.draw_off.clear(); out.ties.clear();
	if (n == 0) { out.mem_off[0] = out.seed_off[0] = out.cand_off[0] = 0; return true; }
	// ---- upload
	out.parse_ok = true;
	uint8_t *d_text = be.template buf<uint8_t>(SL_TEXT, in.text_bytes + 16);
	DevRead *d_reads = be.template buf<DevRead>(SL_READS, n);
	uint8_t *d_flags = be.template buf<uint8_t>(SL_FLAGS, n);
	uint32_t *d_mem_cnt = be.template buf<uint32_t>(SL_MEM_CNT, 2 * n + 1), *d_mem_off = be.template buf<uint32_t>(SL_MEM_OFF, 2 * n + 1);
	uint32_t *d_misc = be.template buf<uint32_t>(SL_MISC, 16);
	if (!d_text || !d_reads || !d_flags || !d_mem_cnt || !d_mem_off || !d_misc) { err = "device stages: out of device memory"; return false; }
	be.h2d(d_text, in.text, in.text_bytes);
	be.zero(d_misc, 64);
	be.zero(d_mem_cnt + 2 * n, 4);
	size_t bits_words = in.bits_words, list_bytes = in.list_bytes;
	DevOri *d_ori = nullptr;
	if (in.parse_text) {
		// ---- the record table from the text itself (kseq_read of strict 4-line FASTQ, clib/utils.c:953-990): line ends per 64-byte
		// piece, counted then placed; one thread per record cuts name / comment / sequence / quality and sizes the read's pools
		const size_t n_piece = (in.text_bytes + NL_PIECE - 1) / NL_PIECE;
		uint32_t *d_nl_cnt = be.template buf<uint32_t>(SL_NL_CNT, n_piece + 1), *d_nl_off = be.template buf<uint32_t>(SL_NL_OFF, n_piece + 1);
		uint32_t *d_lines = be.template buf<uint32_t>(SL_LINES, 4 * n + 16);
		DevRec *d_recs = be.template buf<DevRec>(SL_RECS, n);
		uint32_t *d_lay_cnt = be.template buf<uint32_t>(SL_LAY_CNT, 2 * (n + 1)), *d_lay_off = be.template buf<uint32_t>(SL_LAY_OFF, 2 * (n + 1));
		d_ori = be.template buf<DevOri>(SL_ORI, n);
		if (!d_nl_cnt || !d_nl_off || !d_lines || !d_recs || !d_lay_cnt || !d_lay_off || !d_ori) { err = "device stages: out of device memory"; return false; }
		be.zero(d_nl_cnt + n_piece, 4);
		FnLines fl{d_text, (uint32_t)in.text_bytes, d_nl_cnt, d_nl_off, d_lines, (uint32_t)(4 * n), false};
		be.for_each(n_piece, fl, 0);
		be.scan(d_nl_cnt, d_nl_off, n_piece + 1);
		uint32_t n_lines = 0;
		be.d2h(&n_lines, d_nl_off + n_piece, 4);
		be.sync();
		if (n_lines != 4 * n) { out.parse_ok = false; return true; }      // not the strict 4-line format this block was cut for: the host parses it
		fl.fill = true;
		be.for_each(n_piece, fl, 0);
		be.zero(d_lay_cnt, 2 * (n + 1) * 4);
		be.for_each(n, FnRecord{d_text, d_lines, d_recs, d_lay_cnt, d_lay_cnt + (n + 1), d_misc + 8}, 0);
		be.scan(d_lay_cnt, d_lay_off, n + 1);
		be.scan(d_lay_cnt + (n + 1), d_lay_off + (n + 1), n + 1);
		be.for_each(n, FnReadTable{d_recs, d_lay_off, d_lay_off + (n + 1), d_reads}, 0);
		uint32_t tot[2], bad = 0;
		be.d2h(&tot[0], d_lay_off + n, 4); be.d2h(&tot[1], d_lay_off + (n + 1) + n, 4); be.d2h(&bad, d_misc + 8, 4);
		out.recs.resize(n);
		be.d2h(out.recs.data(), d_recs, n * sizeof(DevRec));
		be.sync();
		if (bad) { out.parse_ok = false; return true; }
		bits_words = tot[0]; list_bytes = tot[1];
		be.for_each(n, FnOri{d_text, d_recs, in.scores.match, d_ori}, 0);
	} else {
		be.h2d(d_reads, in.reads, n * sizeof(DevRead));
		if (in.recs) {                                             // the original alignments (comment fields), when the block came with its record table
			DevRec *d_recs = be.template buf<DevRec>(SL_RECS, n);
			d_ori = be.template buf<DevOri>(SL_ORI, n);
			if (!d_recs || !d_ori) { err = "device stages: out of device memory"; return false; }
			be.h2d(d_recs, in.recs, n * sizeof(DevRec));
			be.for_each(n, FnOri{d_text, d_recs, in.scores.match, d_ori}, 0);
		}
	}
	uint64_t *d_bits = be.template buf<uint64_t>(SL_BITS, bits_words + 2);
	uint8_t *d_list = be.template buf<uint8_t>(SL_LIST, list_bytes + 16);
	if (!d_bits || !d_list) { err = "device stages: out of device memory"; return false; }
	// ---- A
	be.encode(n, FnEncode{d_text, d_reads, d_ori, d_bits, d_list, d_flags});
	// ---- B
	unsigned long long *d_probes = (unsigned long long*)(d_misc + 2);
	Mem *d_mems_kept = be.template buf<Mem>(SL_MEMS_KEPT, 2 * n * (size_t)SEED_TMP_CAP);
	if (!d_mems_kept) { err = "device stages: out of device memory"; return false; }
	be.for_each(2 * n, FnSeed{ix, d_reads, d_bits, d_list, d_flags, d_mem_cnt, d_mems_kept, d_probes}, 1);
	be.scan(d_mem_cnt, d_mem_off, 2 * n + 1);
	be.d2h(out.mem_off.data(), d_mem_off, (2 * n + 1) * 4);
	be.sync();
	const size_t n_mems = out.mem_off[2 * n];
	Mem *d_mems = be.template buf<Mem>(SL_MEMS, n_mems + 1), *d_mems_tmp = be.template buf<Mem>(SL_MEMS_TMP, n_mems + 1);
	uint32_t *d_nvu = be.template buf<uint32_t>(SL_NVU, 2 * n), *d_seed_cnt = be.template buf<uint32_t>(SL_SEED_CNT, 2 * n + 1);
	uint32_t *d_seed_off = be.template buf<uint32_t>(SL_SEED_OFF, 2 * n + 1);
	if (!d_mems || !d_mems_tmp || !d_nvu || !d_seed_cnt || !d_seed_off) { err = "device stages: out of device memory"; return false; }
	be.for_each(2 * n, FnSeedPlace{ix, d_reads, d_bits, d_list, d_flags, d_mem_off, d_mems_kept, d_mems}, 1);
	// ---- C
	be.zero(d_seed_cnt + 2 * n, 4);
	be.for_each(n, FnMerge{d_mem_off, d_mems, d_mems_tmp, d_flags, d_nvu, d_seed_cnt}, 2);
	be.scan(d_seed_cnt, d_seed_off, 2 * n + 1);
	be.d2h(out.seed_off.data(), d_seed_off, (2 * n + 1) * 4);
	be.d2h(out.flags.data(), d_flags, n);
	uint64_t probes = 0;
	be.d2h(&probes, d_probes, 8);
	be.sync();
	out.probes = probes;
	const size_t n_seeds = out.seed_off[2 * n];
	DevSeed *d_seeds = be.template buf<DevSeed>(SL_SEEDS, n_seeds + 1), *d_seeds_tmp = be.template buf<DevSeed>(SL_SEEDS_TMP, n_seeds + 1);
	float *d_dist = be.template buf<float>(SL_DIST, n_seeds + 1);
	int32_t *d_pre = be.template buf<int32_t>(SL_PRE, n_seeds + 1);
	const size_t P = n + 1;
	uint32_t *d_plan_cnt = be.template buf<uint32_t>(SL_PLAN_CNT, PLAN_FIELDS * P), *d_plan_off = be.template buf<uint32_t>(SL_PLAN_OFF, PLAN_FIELDS * P);
	if (!d_seeds || !d_seeds_tmp || !d_dist || !d_pre || !d_plan_cnt || !d_plan_off) { err = "device stages: out of device memory"; return false; }
	be.for_each(n, FnChain{d_mem_off, d_nvu, d_seed_off, d_mems, d_pos, ix.posp, d_flags, d_seeds, d_seeds_tmp, d_dist, d_pre}, 2);
	if (in.want_tables) {
		out.seeds.resize(n_seeds); out.dist.resize(n_seeds); out.pre.resize(n_seeds);
		be.d2h(out.seeds.data(), d_seeds, n_seeds * sizeof(DevSeed));
		be.d2h(out.dist.data(), d_dist, n_seeds * 4);
		be.d2h(out.pre.data(), d_pre, n_seeds * 4);
	}
	// the per-read kernels from here on (planning, chain selection) do work that grows with the read's seeds, from nothing to tens of
	// thousands of instructions: they visit the reads in the order of their seed counts, heaviest first, so that the threads of a warp
	// have about the same to do (results are written per read: the order changes nothing else)
	uint32_t *d_wkey = be.template buf<uint32_t>(SL_WORK_KEY, n), *d_widx = be.template buf<uint32_t>(SL_WORK_IDX, n), *d_perm = be.template buf<uint32_t>(SL_WORK_PERM, n);
	if (!d_wkey || !d_widx || !d_perm) { err = "device stages: out of device memory"; return false; }
	be.for_each(n, FnWorkKey{d_seed_off, d_wkey, d_widx}, 3);
	be.order_desc(d_wkey, d_widx, d_perm, n);
	// ---- D: count, offsets, fill
	int cap = 16;
	for (;;) {                                                     // (again with more room if a ksw CIGAR does not fit `cap` words)
		be.zero(d_plan_cnt, PLAN_FIELDS * P * 4);
		FnPlan fp{in.scores, rf, d_reads, d_bits, d_seed_off, d_seeds, d_dist, d_pre, n, d_plan_cnt, d_plan_off,
		          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (uint32_t)cap, false};
		be.for_each_in(n, d_perm, fp, 3);
		for (int f = 0; f < PLAN_FIELDS; ++f) be.scan(d_plan_cnt + f * P, d_plan_off + f * P, P);
		uint32_t tot[PLAN_FIELDS];
		for (int f = 0; f < PLAN_FIELDS; ++f) be.d2h(&tot[f], d_plan_off + f * P + n, 4);
		be.d2h(out.cand_off.data(), d_plan_off, P * 4);
		be.sync();
		const size_t n_cand = tot[0], n_piece = tot[1], n_task = tot[2], q_bytes = tot[3], t_bytes = tot[4], n_cig = tot[5];
		DevCand *d_cands = be.template buf<DevCand>(SL_CANDS, n_cand + 1);
		DevPiece *d_pieces = be.template buf<DevPiece>(SL_PIECES, n_piece + 1);
		int32_t *d_qlen = be.template buf<int32_t>(SL_QLEN, n_task + 1), *d_tlen = be.template buf<int32_t>(SL_TLEN, n_task + 1);
		int64_t *d_qoff = be.template buf<int64_t>(SL_QOFF, n_task + 1), *d_toff = be.template buf<int64_t>(SL_TOFF, n_task + 1);
		uint8_t *d_q = be.template buf<uint8_t>(SL_Q, q_bytes + 16), *d_t = be.template buf<uint8_t>(SL_T, t_bytes + 16);
		int32_t *d_res = be.template buf<int32_t>(SL_RES, (n_task + 1) * 12);
		uint32_t *d_kcig = be.template buf<uint32_t>(SL_KCIG, (n_task + 1) * (size_t)cap);
		DevCigar *d_cigs = be.template buf<DevCigar>(SL_CIGS, n_cig + 1);
		if (!d_cands || !d_pieces || !d_qlen || !d_tlen || !d_qoff || !d_toff || !d_q || !d_t || !d_res || !d_kcig || !d_cigs) { err = "device stages: out of device memory"; return false; }
		fp.cands = d_cands; fp.pieces = d_pieces; fp.qlen = d_qlen; fp.tlen = d_tlen; fp.qoff = d_qoff; fp.toff = d_toff; fp.q = d_q; fp.t = d_t; fp.fill = true;
		be.for_each_in(n, d_perm, fp, 3);
		unsigned long long *d_cells = (unsigned long long*)(d_misc + 4);
		be.zero(d_cells, 8);
		if (n_task) be.for_each(n_task, FnCells{d_qlen, d_tlen, 200, d_cells}, 5);
		// ---- E
		if (n_task && !be.ksw(n_task, d_q, d_qoff, d_qlen, d_t, d_toff, d_tlen, d_res, d_kcig, cap, err)) return false;
		// ---- F1
		be.zero(d_misc, 4);
		if (n_cand) be.for_each(n_cand, FnResolve{d_cands, d_pieces, d_res, d_kcig, cap, d_reads, d_cigs, d_misc}, 4);
		uint32_t overflow = 0;
		uint64_t cells = 0;
		be.d2h(&overflow, d_misc, 4);
		be.d2h(&cells, d_cells, 8);
		out.cands.resize(in.want_tables ? n_cand : 0); out.cigs.resize(in.want_tables ? n_cig : 0);
		if (in.want_tables) {
			be.d2h(out.cands.data(), d_cands, n_cand * sizeof(DevCand));
			be.d2h(out.cigs.data(), d_cigs, n_cig * sizeof(DevCigar));
		}
		be.sync();
		if (!overflow) { out.n_tasks = n_task; out.n_cells = cells; break; }
		cap *= 4;
	}
	// ---- F (probe): chain selection, candidate sort and pairing of every pair against a scripted generator
	out.pair_probe.clear();
	if (d_ori && (n & 1) == 0) {
		const size_t np = n / 2;
		uint8_t *d_used = be.template buf<uint8_t>(SL_USED, n_seeds + 16);
		DevPairState *d_state = be.template buf<DevPairState>(SL_PSTATE, np);
		DevProbe *d_probe = be.template buf<DevProbe>(SL_PROBE, np);
		if (!d_used || !d_state || !d_probe) { err = "device stages: out of device memory"; return false; }
		be.for_each_in(n, d_perm, FnExplore{pix, d_flags, d_seed_off, d_seeds, d_dist, d_pre, d_used, d_plan_off, be.template buf<DevCand>(SL_CANDS, 0), d_ori, d_state}, 6);
		uint32_t *d_dcnt = be.template buf<uint32_t>(SL_DRAW_CNT, np + 1), *d_doff = be.template buf<uint32_t>(SL_DRAW_OFF, np + 1);
		uint8_t *d_redo = be.template buf<uint8_t>(SL_REDO, np);
		if (!d_dcnt || !d_doff || !d_redo) { err = "device stages: out of device memory"; return false; }
		be.zero(d_dcnt + np, 4);
		be.for_each(np, FnProbe{pix, in.pair_opts, d_flags, d_ori, d_state, d_probe, d_dcnt, d_redo}, 6);
		be.scan(d_dcnt, d_doff, np + 1);
		out.redo.resize(np); out.draw_off.resize(np + 1);
		be.d2h(out.redo.data(), d_redo, np);
		be.d2h(out.draw_off.data(), d_doff, (np + 1) * 4);
		if (in.want_tables) { out.pair_probe.resize(np); be.d2h(out.pair_probe.data(), d_probe, np * sizeof(DevProbe)); }
		// the pairs for the in-order pass itself: counted, placed, gathered
		const size_t TP = np + 1;
		uint32_t *d_tcnt = be.template buf<uint32_t>(SL_TIE_CNT, TIE_FIELDS * TP), *d_toff = be.template buf<uint32_t>(SL_TIE_OFF, TIE_FIELDS * TP);
		if (!d_tcnt || !d_toff) { err = "device stages: out of device memory"; return false; }
		for (int f = 0; f < TIE_FIELDS; ++f) be.zero(d_tcnt + f * TP + np, 4);
		const DevCand *d_cands_all = be.template buf<DevCand>(SL_CANDS, 0);
		be.for_each(np, FnTieCount{d_redo, d_seed_off, d_plan_off, np, d_tcnt}, 6);
		for (int f = 0; f < TIE_FIELDS; ++f) be.scan(d_tcnt + f * TP, d_toff + f * TP, TP);
		uint32_t ttot[TIE_FIELDS];
		for (int f = 0; f < TIE_FIELDS; ++f) be.d2h(&ttot[f], d_toff + f * TP + np, 4);
		be.sync();
		out.ties.resize(ttot[0]); out.tie_seeds.resize(ttot[1]); out.tie_dist.resize(ttot[1]); out.tie_pre.resize(ttot[1]); out.tie_cands.resize(ttot[2]);
		if (ttot[0]) {
			DevTie *d_tie = be.template buf<DevTie>(SL_TIE, ttot[0]);
			DevSeed *d_ts = be.template buf<DevSeed>(SL_TIE_SEEDS, ttot[1] + 1);
			float *d_td = be.template buf<float>(SL_TIE_DIST, ttot[1] + 1);
			int32_t *d_tp = be.template buf<int32_t>(SL_TIE_PRE, ttot[1] + 1);
			DevCand *d_tc = be.template buf<DevCand>(SL_TIE_CANDS, ttot[2] + 1);
			if (!d_tie || !d_ts || !d_td || !d_tp || !d_tc) { err = "device stages: out of device memory"; return false; }
			be.for_each(np, FnTieGather{d_redo, d_seed_off, d_plan_off, d_toff, np, d_seeds, d_dist, d_pre, d_cands_all, d_ori, d_tie, d_ts, d_td, d_tp, d_tc}, 6);
			be.d2h(out.ties.data(), d_tie, (size_t)ttot[0] * sizeof(DevTie));
			be.d2h(out.tie_seeds.data(), d_ts, (size_t)ttot[1] * sizeof(DevSeed));
			be.d2h(out.tie_dist.data(), d_td, (size_t)ttot[1] * 4);
			be.d2h(out.tie_pre.data(), d_tp, (size_t)ttot[1] * 4);
			be.d2h(out.tie_cands.data(), d_tc, (size_t)ttot[2] * sizeof(DevCand));
		}
		be.sync();
	}
	return true;
}

struct FnText {                                                    // the SAM record of read i; len[i] = its length (count pass), or written at off[i]
	const uint8_t *text; const DevRec *recs; const DevOri *ori; const DevFinal *fin; const DevPairFinal *pfin; const DevCand *cands; const DevCigar *cigs;
	TextTables T; const uint32_t *host_len; uint32_t *len; const uint32_t *off; char *out; uint32_t *bad; bool write;
	SEED_HD void operator()(size_t i) const
	{
		const size_t p = i >> 1; const int m = (int)(i & 1);
		if (!write) {
			if (!pfin[p].valid) { len[i] = m == 0 ? host_len[p] : 0u; return; }       // a pair the host path finished: room for its text
			CountSink s; s.n = 0;
			const int b = dev_sam_record(s, (const char*)text, recs[i], ori[i], fin[i], pfin[p], m, cands, cigs, T);
			len[i] = s.n;
#if defined(__CUDA_ARCH__)
			if (b) atomicAdd(bad, 1u);
#else
			if (b) ++*bad;
#endif
		} else {
			if (!pfin[p].valid) return;
			WriteSink s; s.p = out + off[i]; s.n = 0;
			dev_sam_record(s, (const char*)text, recs[i], ori[i], fin[i], pfin[p], m, cands, cigs, T);
		}
	}
	// write pass with the long copies (name, bases, qualities, comment) set aside as jobs for the backend's text_write()
	SEED_HD void prepare(size_t i, JobSink &s) const
	{
		const size_t p = i >> 1;
		s.p = out + off[i]; s.n = 0; s.nj = 0; s.np = 0;
		if (!pfin[p].valid) return;
		dev_sam_record(s, (const char*)text, recs[i], ori[i], fin[i], pfin[p], (int)(i & 1), cands, cigs, T);
	}
};

// Second trip of a block whose pairs were probed: the numbers the in-order pass took from the stream for the block's pairs go up
// (drawn: draw_off[n_pairs] of them, pair p's at draw_off[p]), with the length of the text the host path produced for
// each of its pairs (host_len, 0 elsewhere); primary / secondary / mate of every read come back, and the block's SAM text with
// every record in its place (gaps of host_len bytes where the host's pairs go: out.txt_off[2p] is the place of pair p).
template <class BE>
bool run_device_finalize(BE &be, const PairIndexView &pix, const PairOpts &o, const TextTables &T, size_t n_pairs, const int32_t *drawn, size_t n_drawn, const uint32_t *host_len,
                         const uint32_t *tie_pair, const DevPairState *tie_done, size_t n_ties, DevStageOut &out, HostVec<char> &text_out, std::string &err)
{
	out.fin.resize(2 * n_pairs); out.pfin.resize(n_pairs); out.txt_off.resize(2 * n_pairs + 1);
	out.bad_records = 0;
	text_out.clear();
	if (n_pairs == 0) { out.txt_off[0] = 0; return true; }
	const size_t n = 2 * n_pairs;
	int32_t *d_drawn = be.template buf<int32_t>(SL_DRAWN, n_drawn + 1);
	DevFinal *d_fin = be.template buf<DevFinal>(SL_FINAL, n);
	DevPairFinal *d_pfin = be.template buf<DevPairFinal>(SL_PFINAL, n_pairs);
	uint32_t *d_hl = be.template buf<uint32_t>(SL_HOSTLEN, n_pairs), *d_len = be.template buf<uint32_t>(SL_TXT_LEN, n + 1), *d_off = be.template buf<uint32_t>(SL_TXT_OFF, n + 1);
	uint32_t *d_misc = be.template buf<uint32_t>(SL_MISC, 16);
	if (!d_drawn || !d_fin || !d_pfin || !d_hl || !d_len || !d_off) { err = "device stages: out of device memory"; return false; }
	be.h2d(d_drawn, drawn, n_drawn * 4);
	be.h2d(d_hl, host_len, n_pairs * 4);
	be.zero(d_misc, 64);
	be.zero(d_len + n, 4);
	const DevOri *d_ori = be.template buf<DevOri>(SL_ORI, 0);
	if (n_ties) {                                                  // the pairs the in-order pass finished: their state goes in place
		uint32_t *d_tpair = be.template buf<uint32_t>(SL_TIE_PAIR, n_ties);
		DevPairState *d_tdone = be.template buf<DevPairState>(SL_TIE_DONE, n_ties);
		if (!d_tpair || !d_tdone) { err = "device stages: out of device memory"; return false; }
		be.h2d(d_tpair, tie_pair, n_ties * 4);
		be.h2d(d_tdone, tie_done, n_ties * sizeof(DevPairState));
		be.for_each(n_ties, FnTieScatter{d_tpair, d_tdone, be.template buf<DevPairState>(SL_PSTATE, 0), be.template buf<DevProbe>(SL_PROBE, 0)}, 6);
	}
	be.for_each(n_pairs, FnFinalize{pix, o, d_ori, be.template buf<DevPairState>(SL_PSTATE, 0), be.template buf<DevProbe>(SL_PROBE, 0), be.template buf<uint32_t>(SL_DRAW_OFF, 0), d_drawn,
	                                be.template buf<DevCand>(SL_CANDS, 0), be.template buf<DevCigar>(SL_CIGS, 0), d_fin, d_pfin}, 6);
	// the `-p` candidates: selected on the device, a short list for the host (all of it if the list overflows its room)
	const uint32_t sel_cap = (uint32_t)(n_pairs / 8 + 1024);
	uint32_t *d_sel_pair = be.template buf<uint32_t>(SL_SEL_PAIR, sel_cap);
	DevFinal *d_sel_fin = be.template buf<DevFinal>(SL_SEL_FIN, 2 * (size_t)sel_cap);
	DevPairFinal *d_sel_pfin = be.template buf<DevPairFinal>(SL_SEL_PFIN, sel_cap);
	if (!d_sel_pair || !d_sel_fin || !d_sel_pfin) { err = "device stages: out of device memory"; return false; }
	be.for_each(n_pairs, FnOriSelect{o, d_ori, d_fin, d_pfin, d_misc + 2, sel_cap, d_sel_pair, d_sel_fin, d_sel_pfin}, 6);
	FnText ft{be.template buf<uint8_t>(SL_TEXT, 0), be.template buf<DevRec>(SL_RECS, 0), d_ori, d_fin, d_pfin, be.template buf<DevCand>(SL_CANDS, 0), be.template buf<DevCigar>(SL_CIGS, 0),
	          T, d_hl, d_len, d_off, nullptr, d_misc, false};
	be.for_each(n, ft, 7);
	be.scan(d_len, d_off, n + 1);
	be.d2h(out.txt_off.data(), d_off, (n + 1) * 4);
	uint32_t bad = 0, n_sel = 0;
	be.d2h(&bad, d_misc, 4);
	be.d2h(&n_sel, d_misc + 2, 4);
	be.sync();
	out.bad_records = bad;
	out.sel_all = n_sel > sel_cap;
	if (out.sel_all) {                                             // (more than an eighth of the pairs: everything comes back)
		be.d2h(out.fin.data(), d_fin, n * sizeof(DevFinal));
		be.d2h(out.pfin.data(), d_pfin, n_pairs * sizeof(DevPairFinal));
		out.sel_pair.clear(); out.sel_fin.clear(); out.sel_pfin.clear();
	} else {
		out.sel_pair.resize(n_sel); out.sel_fin.resize(2 * (size_t)n_sel); out.sel_pfin.resize(n_sel);
		be.d2h(out.sel_pair.data(), d_sel_pair, (size_t)n_sel * 4);
		be.d2h(out.sel_fin.data(), d_sel_fin, 2 * (size_t)n_sel * sizeof(DevFinal));
		be.d2h(out.sel_pfin.data(), d_sel_pfin, (size_t)n_sel * sizeof(DevPairFinal));
	}
	const size_t total = out.txt_off[n];
	char *d_txt = be.template buf<char>(SL_TXT, total + 16);
	if (!d_txt) { err = "device stages: out of device memory"; return false; }
	ft.out = d_txt; ft.write = true;
	be.text_write(n, ft);
	// where the text goes: the place the caller hands out for it (e.g. behind the earlier blocks' text in one output buffer), else text_out
	char *dst = out.text_dest ? out.text_dest(total) : nullptr;
	if (!dst) { text_out.resize(total); dst = text_out.data(); } else text_out.clear();
	out.text_ptr = dst; out.text_total = total;
	be.d2h(dst, d_txt, total);
	be.sync();
	return true;
}

// ---- the link-time service (CUDA in the product library, the host backend in tests/emul)
struct StageService;
StageService *stage_service_create(const DebgaIndex &idx, SeedService *seeds, void *ksw_ctx, int device, std::string &err);
void stage_service_destroy(StageService *s);
void stage_service_set_scoring(StageService *s, const AlnScores &o, int zdrop);   // ksw parameters of stage E (copy_option, RR:817-827)
// A service instance holds one block's device state from stage_service_run to stage_service_finalize (two trips with the host's
// in-order pass in between); blocks in flight at the same time use different instances.
bool stage_service_run(StageService *s, const DevStageIn &in, DevStageOut &out, std::string &err);
bool stage_service_finalize(StageService *s, const PairOpts &o, int not_ori, size_t n_pairs, const int32_t *drawn, size_t n_drawn, const uint32_t *host_len,
                            const uint32_t *tie_pair, const DevPairState *tie_done, size_t n_ties, DevStageOut &out,
                            HostVec<char> &text_out, std::string &err);

} // namespace pansvr
