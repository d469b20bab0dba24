// pair_core.cuh -- stage F of `fc_aln` for one read pair, as host/device functions over the flat results of stages A..F1
// (stages_core.cuh): chain selection, candidate sort, pairing and primary / secondary / mate assignment.
//
//   dev_sort_output        single_end_handler::sort_output                          RR:212-293
//   dev_finish_read        the rest of single_end_handler::align                    RR:416-475
//   dev_explore_read       every outcome of a read's rand() ties (pipeline.cpp: explore_read)
//   dev_pair_up            PE_score::read_get_best_pairing_results / store_pair     RRH:434-499, 536-628
//   dev_set_primary        set_primary_secondary_mate                               RRH:501-534
//
// Only exact ties consume the reference's rand() (RR:247, RRH:553).  The probe runs every pair against a scripted generator:
// a pair that never asks is final; a read whose ties all lead to the same candidates only advances the stream by a known
// number of draws; pairing ties are recorded as events and redrawn by the host's in-order pass from the real stream, after
// which dev_finalize applies the winner.  Pairs whose outcomes differ go back to the host path (pipeline.cpp).
// Sorting replicates glibc's qsort (a top-down merge sort that takes the left element when cmp <= 0), because the reference's
// comparators return 0/1 and are not orders (RRH:303-315).
#pragma once
#include <stdint.h>

#include "stages_core.cuh"
#include "refrand.hpp"

namespace pansvr {

enum { PR_FORWARD = 1, PR_REVERSE = 0, PR_MAX_OUTPUT = 6, PR_MAX_RES = 12, PR_MAX_SCRIPT = 12, PR_MAX_LEAVES = 24, PR_MAX_EVENTS = 32 };
enum { PR_MIN_ALN_SCORE = 40 };
enum { PR_REDO_HOST = 255 };                                       // the pair goes back to the host path ('N', random_r sampling)
enum { PR_REDO_TIES = 254 };                                       // its rand() ties decide the outcome: finished in its turn of the in-order pass, from what the device computed
const uint32_t PR_U32MAX = 0xffffffffu;

struct DevOri { uint32_t chr, ref_bg, read_bg, align_score; uint8_t mapq, direction, unmapped, skip; };   // parse_ori_mapping_rst + RR:413-414
struct DevSv { uint32_t chr_id, st_pos; int32_t end_offset; uint32_t pad; };                                  // SV_chr_info, per anchor id
struct PairIndexView { const uint32_t *chr_search_index, *chr_end_n; const DevSv *sv; };
struct PairOpts { int isize_max, isize_min, read_len, min_filter_score; };   // (min_filter_score: pairs scoring at most this go to the `-p` output, RR:776)

struct DevRes {                                                    // MAX_IDX_OUTPUT, RRH:232-318 (what stage F needs of it)
	uint32_t align_score, chain_score, max_index, read_bg, chr, ref_bg;
	int32_t sv, cand;                                              // anchor id (-1: none); global index of its DevCand
	uint8_t direction, mapq, cigar_ok, rst_idx;
};
struct DevPE { int32_t max_same, max_score, cur_isize; int8_t m1, m2; uint8_t proper, gain; };   // m1/m2: index into pick() (-1 none)
struct DevPairState { DevRes res[2][PR_MAX_RES]; uint8_t n[2]; uint8_t pad[2]; int16_t draws[2]; DevPE pe; };   // draws: rand() calls of each read (-1: outcomes differ)
struct DevProbe { uint8_t redo, draws0, draws1, ev_cnt; int8_t ev_i[PR_MAX_EVENTS], ev_j[PR_MAX_EVENTS]; uint32_t tie_mask; };
struct DevFinal {                                                  // what the record text of one read needs
	uint32_t flags;                                                // FIN_*
	uint32_t p_chr, p_ref_bg, p_align, p_chain, p_mapq; int32_t p_cand, p_sv, p_mate_sv; uint32_t mate_chr, mate_ref_bg;
	uint32_t s_chr, s_ref_bg, s_read_bg, s_align; int32_t s_sv;
	uint32_t p_ins, p_ncig;                                        // inserted bases / entries of the primary's CIGAR (the `-p` decision reads them, RR:788-792)
};
enum { FIN_PRIMARY = 1, FIN_P_ORI = 2, FIN_HAS_MATE = 4, FIN_SECONDARY = 8, FIN_P_FWD = 16, FIN_S_FWD = 32, FIN_P_CIGAR_OK = 64 };
struct DevPairFinal { int32_t max_score, cur_isize; uint8_t gain, proper, valid, pad; };

struct DevTap {                                                    // RandTap of pipeline.cpp: scripted (the probe), or the reference's stream itself (host only)
	uint32_t calls, script_len; bool too_deep;
	uint8_t script[PR_MAX_SCRIPT], moduli[PR_MAX_SCRIPT];
	void *real = nullptr;                                          // host: a GlibcRandom to draw from (rand() % m, RR:247, RRH:553)
	SEED_HD int32_t draw(int32_t m)
	{
#if !defined(__CUDA_ARCH__)
		if (real) { ++calls; return (int32_t)(((GlibcRandom*)real)->next() % m); }
#endif
		const uint32_t k = calls++;
		if (k >= PR_MAX_SCRIPT || m > 255) { too_deep = true; return 0; }
		moduli[k] = (uint8_t)m;
		return k < script_len ? (int32_t)script[k] : 0;
	}
	SEED_HD void restart(uint32_t len) { calls = 0; script_len = len; too_deep = false; }
};

struct ReadView {                                                  // one read of a pair, as stages A..F1 left it
	const DevSeed *v[2]; const float *dist[2]; const int32_t *pre[2]; uint8_t *used[2]; uint32_t n[2];
	const DevCand *cands; uint32_t cand_b, cand_e;                 // its candidates are cands[cand_b .. cand_e), ascending (strand, node)
	DevOri ori;
};

SEED_HD int dev_chromosome_id(const PairIndexView &ix, uint32_t position)       // deBGA_INDEX::get_chromosome_ID, IDX:369-396
{
	int file_n = 0;
	const int pos_index = (int)(position / 0x4000);
	int low = (int)ix.chr_search_index[pos_index], high = (int)ix.chr_search_index[pos_index + 1];
	const int pos = (int)position + 1;
	while (low <= high) {
		const int mid = (low + high) >> 1;
		const uint32_t e = ix.chr_end_n[mid] - 1u;
		if ((uint32_t)pos < e) high = mid - 1;
		else if ((uint32_t)pos > e) low = mid + 1;
		else return mid;
		file_n = low;
	}
	return file_n;
}

SEED_HD int dev_sort_output(const PairIndexView &ix, const ReadView &R, int s, DevRes &rst, int direction, DevTap &rnd)
{
	const int n = (int)R.n[s];
	if (n == 0) return 0;
	const float *dist = R.dist[s]; const int32_t *pre = R.pre[s]; uint8_t *used_f = R.used[s];
	for (;;) {
		uint32_t max_index = PR_U32MAX; float max_distance = 0; uint32_t same = 1;
		for (int i = n - 1; i >= 0; --i) {
			if (used_f[i]) continue;
			const float d = dist[i];
			if (max_distance < d) { max_distance = d; max_index = (uint32_t)i; same = 1; }
			else if (max_distance == d) ++same;
		}
		if (max_index == PR_U32MAX) return 0;
		if (same > 1) {                                            // rand() % same picks among the tied ends, listed from the highest index down
			int k = rnd.draw((int32_t)same);
			for (int i = n - 1; i >= 0; --i) {
				if (used_f[i] || dist[i] != max_distance) continue;
				if (k == 0) { max_index = (uint32_t)i; break; }
				--k;
			}
		}
		int used = 0, fresh = 0;
		int node = (int)max_index;
		const int first = node;
		for (; node != -1;) {
			if (used_f[node]) ++used; else ++fresh;
			used_f[node] = 1;
			const int next = pre[node];
			if (next == -1) break;
			node = next;
		}
		const int last = node;
		if (first - last > ((fresh + used + 5) << 1))
			for (int k = last; k < first; ++k) used_f[k] = 1;
		if (used >= fresh) continue;
		const int ref_begin = (int)R.v[s][node].ref_begin;
		const int chr = dev_chromosome_id(ix, (uint32_t)ref_begin);
		rst.direction = (uint8_t)direction;
		rst.max_index = max_index;
		rst.chain_score = (uint32_t)max_distance;
		rst.read_bg = R.v[s][node].read_begin;
		rst.chr = (uint32_t)chr;
		rst.ref_bg = (uint32_t)(ref_begin - (int)(chr >= 1 ? ix.chr_end_n[chr - 1] : 0u));
		return 1;
	}
}

// glibc's msort_with_tmp over n <= 12 records; after(x, y) = the comparator's return value (0 or 1).  The recursion is unrolled
// through the template depth (12 -> 6 -> 3 -> 2 -> 1), so the device code needs no call stack.
template <int DEPTH, class After>
SEED_HD void dev_msort(DevRes *b, DevRes *tmp, int n, After after)
{
	if (n <= 1) return;
	const int n1 = n / 2, n2 = n - n1;
	if (DEPTH > 0) {
		dev_msort<(DEPTH > 0 ? DEPTH - 1 : 0)>(b, tmp, n1, after);
		dev_msort<(DEPTH > 0 ? DEPTH - 1 : 0)>(b + n1, tmp, n2, after);
	}
	int i = 0, j = n1, k = 0;
	while (i < n1 && j < n) { if (after(b[i], b[j]) <= 0) tmp[k++] = b[i++]; else tmp[k++] = b[j++]; }
	while (i < n1) tmp[k++] = b[i++];
	for (int x = 0; x < k; ++x) b[x] = tmp[x];                      // (the tail of the right half is already in place)
}
struct AfterChain { SEED_HD int operator()(const DevRes &x, const DevRes &y) const { return x.chain_score != y.chain_score ? x.chain_score < y.chain_score : x.max_index > y.max_index; } };
struct AfterAlign { SEED_HD int operator()(const DevRes &x, const DevRes &y) const { return x.align_score != y.align_score ? x.align_score < y.align_score : x.max_index > y.max_index; } };

// The part of single_end_handler::align that draws (RR:416-437): up to six chain ends per strand, sorted by chain score.  Returns
// their number; everything behind it (dev_align_chains) is a function of this list.
SEED_HD int dev_select_chains(const PairIndexView &ix, const ReadView &R, DevRes *res, DevTap &rnd, uint32_t *max_chain_out)
{
	for (int s = 0; s < 2; ++s) for (uint32_t i = 0; i < R.n[s]; ++i) R.used[s][i] = 0;
	int result_num = 0;
	uint32_t max_chain = 0;
	for (int s = 0; s < 2; ++s) {
		const int direction = s == 0 ? PR_FORWARD : PR_REVERSE;
		for (int i = 0; i < PR_MAX_OUTPUT; ++i) {
			DevRes slot;
			slot.align_score = 0; slot.sv = -1; slot.cand = -1; slot.mapq = 0; slot.cigar_ok = 1; slot.rst_idx = 0;
			if (!dev_sort_output(ix, R, s, slot, direction, rnd)) break;
			const uint32_t c = slot.chain_score;
			if (c > max_chain) max_chain = c;
			if (c + ST_MAX_CHAIN_SCORE_DIFF < max_chain || c < (uint32_t)ST_MIN_CHAIN_SCORE2) break;
			res[result_num++] = slot;
		}
	}
	DevRes tmp[PR_MAX_RES];
	dev_msort<4>(res, tmp, result_num, AfterChain());
	*max_chain_out = max_chain;
	return result_num;
}

// The rest of it (RR:438-475): alignment scores of the chosen chain ends, sort, anchor coordinates, mapq.  Returns the number of
// candidates left in res[] (sorted by alignment score), or -1 if a chain end was not planned (cannot happen).
SEED_HD int dev_align_chains(const PairIndexView &ix, const ReadView &R, DevRes *res, int result_num, uint32_t max_chain)
{
	if (result_num == 0 || max_chain < (uint32_t)ST_MIN_CHAIN_SCORE) return result_num;
	for (int k = 0; k < result_num; ++k) {
		DevRes &c = res[k];
		if (c.chain_score + ST_MAX_CHAIN_SCORE_DIFF < max_chain) { result_num = k; break; }
		const uint32_t s = c.direction == PR_REVERSE ? 1u : 0u;
		uint32_t lo = R.cand_b, hi = R.cand_e;                     // the candidate planned for (strand, chain end)
		while (lo < hi) {
			const uint32_t m = (lo + hi) >> 1;
			const DevCand &cd = R.cands[m];
			if (cd.strand < s || (cd.strand == s && cd.node < c.max_index)) lo = m + 1; else hi = m;
		}
		if (lo >= R.cand_e || R.cands[lo].strand != s || R.cands[lo].node != c.max_index) return -1;
		const DevCand &cd = R.cands[lo];
		c.ref_bg -= (uint32_t)cd.read_begin_alignment;
		c.align_score = cd.align_score;
		c.cigar_ok = (uint8_t)cd.cigar_ok;
		c.cand = (int32_t)lo;
	}
	DevRes tmp[PR_MAX_RES];
	dev_msort<4>(res, tmp, result_num, AfterAlign());
	if (result_num == 0) return 0;
	if (res[0].align_score < (uint32_t)PR_MIN_ALN_SCORE) return 0;
	for (int i = 0; i < result_num; ++i) {
		DevRes &c = res[i];
		const uint32_t sv_id = c.chr;
		c.sv = (int32_t)sv_id;
		c.chr = ix.sv[sv_id].chr_id;
		c.ref_bg += ix.sv[sv_id].st_pos;
		if (c.ref_bg >= 0x7fffffffu) c.ref_bg = 5;
		c.rst_idx = (uint8_t)i; c.mapq = 0;
	}
	const int32_t d = (int32_t)(res[0].align_score - (result_num > 1 ? res[1].align_score : 0));
	res[0].mapq = (uint8_t)(d > 40 ? 40 : d);
	return result_num;
}

SEED_HD int dev_finish_read(const PairIndexView &ix, const ReadView &R, DevRes *res, DevTap &rnd)
{
	if (R.ori.skip) return 0;
	uint32_t max_chain = 0;
	const int n = dev_select_chains(ix, R, res, rnd, &max_chain);
	return dev_align_chains(ix, R, res, n, max_chain);
}

SEED_HD bool dev_same_results(const DevRes *a, int na, const DevRes *b, int nb)
{
	if (na != nb) return false;
	for (int k = 0; k < na; ++k) {
		const DevRes &x = a[k], &y = b[k];
		if (x.align_score != y.align_score || x.chain_score != y.chain_score || x.max_index != y.max_index || x.read_bg != y.read_bg ||
		    x.chr != y.chr || x.ref_bg != y.ref_bg || x.direction != y.direction || x.mapq != y.mapq || x.sv != y.sv || x.cand != y.cand) return false;
	}
	return true;
}

// explore_read of pipeline.cpp: the number of draws the read makes if every outcome of its ties chooses the same chain ends (then
// everything behind the choice is the same too), else -1.  res / *n_res = the candidates.  Only the drawing part of the read's
// finish is run again for the other outcomes.
SEED_HD int dev_explore_read(const PairIndexView &ix, const ReadView &R, DevRes *res, int *n_res, DevTap &probe)
{
	probe.restart(0);
	*n_res = 0;
	if (R.ori.skip) return 0;
	uint32_t max_chain = 0;
	const int n_sel = dev_select_chains(ix, R, res, probe, &max_chain);
	if (probe.too_deep) return -1;
	const uint32_t c0 = probe.calls;
	if (c0 != 0) {
		uint8_t choice[PR_MAX_SCRIPT], mod[PR_MAX_SCRIPT];
		for (int k = 0; k < PR_MAX_SCRIPT; ++k) { choice[k] = 0; mod[k] = k < (int)c0 ? probe.moduli[k] : 0; }
		uint32_t c = c0;
		DevRes alt[PR_MAX_RES];
		for (int leaves = 1; ; ++leaves) {
			int p = (int)c - 1;
			while (p >= 0 && choice[p] + 1 >= mod[p]) --p;
			if (p < 0) break;
			if (leaves >= PR_MAX_LEAVES) return -1;
			++choice[p];
			for (int k = p + 1; k < PR_MAX_SCRIPT; ++k) choice[k] = 0;
			probe.restart((uint32_t)p + 1);
			for (int k = 0; k <= p; ++k) probe.script[k] = choice[k];
			uint32_t mc = 0;
			const int na = dev_select_chains(ix, R, alt, probe, &mc);
			if (probe.too_deep || probe.calls != c0 || mc != max_chain) return -1;
			c = probe.calls;
			for (uint32_t k = 0; k < c; ++k) mod[k] = probe.moduli[k];
			if (!dev_same_results(res, n_sel, alt, na)) return -1;
		}
	}
	const int n0 = dev_align_chains(ix, R, res, n_sel, max_chain);
	if (n0 < 0) return -1;
	*n_res = n0;
	return (int)c0;
}

// ---- PE_score
struct PairSide { const DevRes *res; int n; DevOri ori; };        // pick(i): i < n ? res[i] : the original alignment
struct Picked { bool some, is_ori; uint32_t chr, ref_bg, align_score; int direction; int32_t sv; };
SEED_HD Picked dev_pick(const PairSide &S, int i)
{
	Picked p;
	p.some = i >= 0; p.is_ori = false; p.chr = 0; p.ref_bg = 0; p.align_score = 0; p.direction = PR_FORWARD; p.sv = -1;
	if (i < 0) return p;
	if (i < S.n) { const DevRes &r = S.res[i]; p.chr = r.chr; p.ref_bg = r.ref_bg; p.align_score = r.align_score; p.direction = r.direction; p.sv = r.sv; }
	else { p.is_ori = true; p.chr = S.ori.chr; p.ref_bg = S.ori.ref_bg; p.align_score = S.ori.align_score; p.direction = S.ori.direction; }
	return p;
}
SEED_HD int dev_get_isize(const PairOpts &o, int p1, int p2, int d1, int d2)
{
	if (d1 == d2) return 0;
	const int max_isize = o.isize_max + 200, min_isize = o.isize_min - 200 > 0 ? o.isize_min - 200 : 0;
	const int isize = o.read_len + (d1 == PR_FORWARD ? p2 - p1 : p1 - p2);
	return (isize < max_isize && isize > min_isize) ? isize : 0;
}
SEED_HD int dev_proper_mated(const PairIndexView &ix, const PairOpts &o, const Picked &a, const Picked &b)
{
	if (!a.some || !b.some || a.chr != b.chr) return 0;
	const int a1 = (int)a.ref_bg, a2 = a1 + (a.is_ori ? 0 : ix.sv[a.sv].end_offset);
	const int b1 = (int)b.ref_bg, b2 = b1 + (b.is_ori ? 0 : ix.sv[b.sv].end_offset);
	int v;
	if ((v = dev_get_isize(o, a1, b1, a.direction, b.direction)) > 0) return v;
	if ((v = dev_get_isize(o, a1, b2, a.direction, b.direction)) > 0) return v;
	if ((v = dev_get_isize(o, a2, b1, a.direction, b.direction)) > 0) return v;
	if ((v = dev_get_isize(o, a2, b2, a.direction, b.direction)) > 0) return v;
	return 0;
}
struct EventSink { DevProbe *pr; bool overflow; };
SEED_HD void dev_store_pair(const PairIndexView &ix, const PairOpts &o, DevPE &pe, const PairSide *S, int i, int j, DevTap &rnd, EventSink *ev)
{
	const Picked a = dev_pick(S[0], i), b = dev_pick(S[1], j);
	const int isize = dev_proper_mated(ix, o, a, b);
	const int basic = (a.some ? (int)a.align_score : 0) + (b.some ? (int)b.align_score : 0);
	const bool one_new = (a.some && !a.is_ori) || (b.some && !b.is_ori);
	const int fin = basic + (isize > 0 ? 0 : -60) + (one_new ? 0 : 1);
	if (fin >= pe.max_score) {
		bool store = true;
		if (ev) {
			DevProbe &p = *ev->pr;
			if (p.ev_cnt >= PR_MAX_EVENTS) ev->overflow = true;
			else { p.ev_i[p.ev_cnt] = (int8_t)i; p.ev_j[p.ev_cnt] = (int8_t)j; if (fin == pe.max_score) p.tie_mask |= 1u << p.ev_cnt; ++p.ev_cnt; }
		}
		if (fin > pe.max_score) pe.max_same = 1;
		else if (fin == pe.max_score) { ++pe.max_same; if (rnd.draw(pe.max_same) != 0) store = false; }
		if (store) { pe.m1 = (int8_t)i; pe.m2 = (int8_t)j; pe.max_score = fin; pe.cur_isize = isize; pe.proper = isize > 0; }
	}
}
SEED_HD bool dev_is_new(const PairSide &S, int i) { return i >= 0 && i < S.n; }
SEED_HD void dev_pair_up(const PairIndexView &ix, const PairOpts &o, const PairSide *S, DevPE &pe, DevTap &rnd, EventSink *ev)
{
	pe.max_same = 1; pe.max_score = 0; pe.cur_isize = 0; pe.proper = 0; pe.gain = 0; pe.m1 = pe.m2 = -1;
	int n0 = S[0].n, n1 = S[1].n;
	if (!S[0].ori.unmapped) ++n0;
	if (!S[1].ori.unmapped) ++n1;
	for (int i = 0; i < n0; ++i) dev_store_pair(ix, o, pe, S, i, -1, rnd, ev);
	for (int j = 0; j < n1; ++j) dev_store_pair(ix, o, pe, S, -1, j, rnd, ev);
	for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) dev_store_pair(ix, o, pe, S, i, j, rnd, ev);
	pe.gain = pe.max_score > 0 && (dev_is_new(S[0], pe.m1) || dev_is_new(S[1], pe.m2));
}

// The probe of one pair, in two steps so that the two reads of a pair are explored by two threads:
//   dev_explore_store   one read: its candidates and the number of rand() draws its ties make (-1: the outcomes differ)
//   dev_probe_pair      the pair: pairing against the scripted generator.  pr.redo = what the in-order pass has to do for it:
//     0 nothing; 1 advance the stream by draws0 + draws1; 2 the same, then redraw the pairing ties from the events;
//     PR_REDO_TIES: the outcomes of a read's ties differ (or the pairing has too many events): the in-order pass finishes the pair
//     against the stream itself, from the seeds, chains and candidates the device has (dev_finish_pair_in_order);
//     PR_REDO_HOST (set by the caller): the host path finishes this pair (a read the device handed back)
SEED_HD void dev_explore_store(const PairIndexView &ix, const ReadView &R, DevPairState &st, int k)
{
	DevTap probe;
	DevRes res[PR_MAX_RES];                                        // (worked on locally; only the candidates that exist go to the pair's state)
	int n = 0;
	const int c = dev_explore_read(ix, R, res, &n, probe);
	st.draws[k] = (int16_t)((c < 0 || c > 250) ? -1 : c);
	st.n[k] = (uint8_t)(c < 0 ? 0 : n);
	if (c >= 0) for (int x = 0; x < n; ++x) st.res[k][x] = res[x];
}
SEED_HD void dev_probe_pair(const PairIndexView &ix, const PairOpts &o, const DevOri *ori, DevPairState &st, DevProbe &pr)
{
	pr.redo = 0; pr.draws0 = pr.draws1 = 0; pr.ev_cnt = 0; pr.tie_mask = 0;
	const int c0 = st.draws[0], c1 = st.draws[1];
	if (c0 < 0 || c1 < 0) { pr.redo = PR_REDO_TIES; return; }
	pr.draws0 = (uint8_t)c0; pr.draws1 = (uint8_t)c1;
	PairSide S[2];
	S[0].res = st.res[0]; S[0].n = st.n[0]; S[0].ori = ori[0];
	S[1].res = st.res[1]; S[1].n = st.n[1]; S[1].ori = ori[1];
	DevTap probe;
	probe.restart(0);
	EventSink ev; ev.pr = &pr; ev.overflow = false;
	dev_pair_up(ix, o, S, st.pe, probe, &ev);
	if (probe.calls == 0) { pr.redo = (c0 + c1) ? 1 : 0; pr.ev_cnt = 0; pr.tie_mask = 0; }
	else if (!ev.overflow) pr.redo = 2;
	else pr.redo = PR_REDO_TIES;
}

// A pair whose ties decide its outcome, in its turn of the in-order pass: both reads finished and paired against the stream
// itself (tap.real), exactly the reference's calls in the reference's order (RR:416-475 for each read, then RRH:434-499).
// R[k]: the reads' seeds, chain tables and candidates as the device left them (host copies); st receives what dev_finalize_pair
// and the record text read.
SEED_HD void dev_finish_pair_in_order(const PairIndexView &ix, const PairOpts &o, const ReadView *R, DevPairState &st, DevTap &tap)
{
	PairSide S[2];
	for (int k = 0; k < 2; ++k) {
		int n = dev_finish_read(ix, R[k], st.res[k], tap);
		if (n < 0) n = 0;
		st.n[k] = (uint8_t)n; st.draws[k] = 0;
		S[k].res = st.res[k]; S[k].n = n; S[k].ori = R[k].ori;
	}
	dev_pair_up(ix, o, S, st.pe, tap, nullptr);
}

// What a pair takes from the reference's rand() stream when nothing about it depends on the numbers but the pairing winner:
// the draws of its two reads, then one draw per tied pairing event (RRH:553).
SEED_HD uint32_t dev_pair_draws(const DevProbe &pr)
{
	if (pr.redo != 1 && pr.redo != 2) return 0;
	uint32_t n = (uint32_t)pr.draws0 + pr.draws1;
	for (uint32_t m = pr.tie_mask; m; m &= m - 1) ++n;
	return n;
}

// Primary / secondary / mate of both reads once the pairing is decided.  A pair with redo == 2 redraws its pairing ties here from
// `drawn`, the numbers the in-order pass took from the stream for this pair (store_pair's rule over the recorded events).
// set_primary_secondary_mate, RRH:501-534, and what output_BAM reads of the results.
SEED_HD void dev_finalize_pair(const PairIndexView &ix, const PairOpts &o, const DevOri *ori, DevPairState &st, const DevProbe &pr, const int32_t *drawn,
                               const DevCand *cands, const DevCigar *cigs, DevFinal *fin, DevPairFinal &pf)
{
	for (int k = 0; k < 2; ++k) { DevFinal &f = fin[k]; f.flags = 0; f.p_ins = f.p_ncig = 0; f.p_chr = f.p_ref_bg = f.p_align = f.p_chain = f.p_mapq = 0; f.p_cand = f.p_sv = f.p_mate_sv = -1; f.mate_chr = f.mate_ref_bg = 0; f.s_chr = f.s_ref_bg = f.s_read_bg = f.s_align = 0; f.s_sv = -1; }
	pf.max_score = 0; pf.cur_isize = 0; pf.gain = pf.proper = 0; pf.valid = 0; pf.pad = 0;
	const uint8_t redo = pr.redo;
	if (redo == PR_REDO_HOST || redo == PR_REDO_TIES) return;         // (a pair finished in order arrives with redo 0 and its state in place)
	PairSide S[2];
	for (int k = 0; k < 2; ++k) { S[k].res = st.res[k]; S[k].n = st.n[k]; S[k].ori = ori[k]; }
	DevPE pe = st.pe;
	if (redo == 2) {                                               // the tie rule over the events, then apply_pairing
		const int32_t *r = drawn + pr.draws0 + pr.draws1;
		int max_same = 1, win_i = -1, win_j = -1;
		for (int k = 0; k < pr.ev_cnt; ++k) {
			if (!((pr.tie_mask >> k) & 1)) { max_same = 1; win_i = pr.ev_i[k]; win_j = pr.ev_j[k]; }
			else { ++max_same; if (*r++ % max_same == 0) { win_i = pr.ev_i[k]; win_j = pr.ev_j[k]; } }
		}
		pe.m1 = (int8_t)win_i; pe.m2 = (int8_t)win_j;
		const Picked a = dev_pick(S[0], win_i), b = dev_pick(S[1], win_j);
		pe.cur_isize = dev_proper_mated(ix, o, a, b);
		pe.proper = pe.cur_isize > 0;
		pe.gain = pe.max_score > 0 && (dev_is_new(S[0], pe.m1) || dev_is_new(S[1], pe.m2));
	}
	pf.max_score = pe.max_score; pf.cur_isize = pe.cur_isize; pf.gain = pe.gain; pf.proper = pe.proper; pf.valid = 1;
	// (the chosen alignments are reported even without a gain: the -p decision reads them)
	int32_t ori_sv[2] = {-1, -1};                                  // an original alignment takes its mate's anchor (RRH:527): seen by the other read
	for (int k = 0; k < 2; ++k) {
		const int mi = k == 0 ? pe.m1 : pe.m2, oi = k == 0 ? pe.m2 : pe.m1;
		if (mi < 0) continue;
		DevFinal &f = fin[k];
		const Picked c = dev_pick(S[k], mi);
		Picked m = dev_pick(S[1 - k], oi);
		if (m.some && m.is_ori) m.sv = ori_sv[1 - k];
		f.flags |= FIN_PRIMARY;
		if (c.is_ori) f.flags |= FIN_P_ORI;
		if (c.direction == PR_FORWARD) f.flags |= FIN_P_FWD;
		f.p_chr = c.chr; f.p_ref_bg = c.ref_bg; f.p_align = c.align_score; f.p_sv = c.sv;
		if (!c.is_ori) {
			const DevRes &r = S[k].res[mi];
			f.p_chain = r.chain_score; f.p_mapq = r.mapq; f.p_cand = r.cand;
			if (r.cigar_ok) f.flags |= FIN_P_CIGAR_OK;
			const DevCand &cd = cands[r.cand];
			f.p_ncig = cd.n_cig;
			int ins = 0;
			for (uint32_t x = 0; x < cd.n_cig; ++x) if (cigs[cd.cig_off + x].type == 1) ins += cigs[cd.cig_off + x].size;
			f.p_ins = (uint32_t)ins;
		}
		else { f.p_mapq = S[k].ori.mapq; f.flags |= FIN_P_CIGAR_OK; }
		const DevRes *sec = nullptr;
		if (c.is_ori && S[k].n > 0) sec = &S[k].res[0];
		else if (S[k].n > 1) sec = (!c.is_ori && S[k].res[mi].rst_idx == 0) ? &S[k].res[1] : &S[k].res[0];
		if (sec) { f.flags |= FIN_SECONDARY; if (sec->direction == PR_FORWARD) f.flags |= FIN_S_FWD; f.s_chr = sec->chr; f.s_ref_bg = sec->ref_bg; f.s_read_bg = sec->read_bg; f.s_align = sec->align_score; f.s_sv = sec->sv; }
		if (m.some && m.chr != PR_U32MAX) {
			f.flags |= FIN_HAS_MATE; f.mate_chr = m.chr; f.mate_ref_bg = m.ref_bg; f.p_mate_sv = m.sv;
			if (c.is_ori) { f.p_sv = m.sv; ori_sv[k] = m.sv; }
		}
	}
}

} // namespace pansvr
