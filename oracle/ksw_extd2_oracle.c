/*
 * ksw_extd2_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker).
 *
 * Plain scalar C restatement of the banded two-piece affine-gap global/extension DP
 * that the reference computes in
 *     /root/reference/src/kswlib/ksw2_extd2_sse.c:26-396   (ksw_extd2_sse)
 *     /root/reference/src/kswlib/ksw2.h:106-151,238-261    (push_cigar, backtrack_D, reset, zdrop)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object.  The product path (pansvr_b200/csrc) never does.
 *
 * PARITY STATUS: pinned.  The reference ships no golden vectors, so
 * this restatement is pinned by differential fuzzing against the reference's own
 * ksw2_extd2_sse.c compiled unmodified into oracle/_ref/libksw_ref.so
 * (tests/test_oracle_vs_ref.py, tests/golden/make_ksw_golden.py) and by the committed
 * fixtures under tests/golden/ that were produced by that library.
 *
 * The reference is a 16-lane int8 SSE program.  What is observable from outside is not
 * "the banded DP" but the precise machine it implements; this file models that machine
 * one byte lane at a time:
 *
 *   - seven persistent int8 rows indexed by target position t (u v x y x2 y2 s), never
 *     re-initialised between anti-diagonals            (KSW:100-109)
 *   - every anti-diagonal r updates the cells of the band [st0,en0] *rounded outwards to
 *     16-cell blocks* [st,en]; the extra cells are computed from whatever the rows hold
 *     (KSW:139-140, 221-267)
 *   - the substitution row s[] is refreshed in unaligned 16-byte stores that start at st0,
 *     so it covers [st0, st0+16*ceil((en0-st0+1)/16)) and is stale elsewhere (KSW:158-173)
 *   - s[], the zero-padded target copy and the zero-padded reversed query live back to
 *     back in one calloc'ed block, in that order, so loads that run past the target read
 *     the reversed query and stores that run past s[] land in the target copy (KSW:100-103,
 *     121-122).  glibc returns 16-byte aligned blocks, so the layout is deterministic.
 *   - all cell arithmetic wraps in int8                 (KSW:30-58)
 *   - the exact-H side row is int32, updated only inside [st0,en0], with the 4-lane argmax
 *     tie order of the SSE scan                         (KSW:316-351)
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_NEG_INF (-0x40000000)

#define ORA_SCORE_ONLY  0x01
#define ORA_RIGHT       0x02
#define ORA_GENERIC_SC  0x04
#define ORA_APPROX_MAX  0x08
#define ORA_APPROX_DROP 0x10
#define ORA_EXTZ_ONLY   0x40
#define ORA_REV_CIGAR   0x80

/* Same memory layout as the reference's ksw_extz_t (K2H:26-35): the struct is the ABI. */
typedef struct {
	uint32_t max:31, zdropped:1;
	int max_q, max_t;
	int mqe, mqe_t;
	int mte, mte_q;
	int score;
	int m_cigar, n_cigar;
	int reach_end;
	uint32_t *cigar;
} ora_extz_t;

/* Range telemetry (tests only): extremes of the stored rows and the number of int8 operations
 * that actually wrapped, accumulated over all calls since the last ksw_extd2_oracle_stats_reset(). */
static __thread int64_t g_wraps;
static __thread int g_rng[8]; /* uv_min uv_max x_max x2_max s_min s_max - - */
static inline int8_t w8(int v) { /* int8 wrap-around */
	if (v < -128 || v > 127) ++g_wraps;
	return (int8_t)(uint8_t)(v & 0xff);
}
void ksw_extd2_oracle_stats_reset(void) { g_wraps = 0; g_rng[0] = 127; g_rng[1] = -128; g_rng[2] = g_rng[3] = -128; }
void ksw_extd2_oracle_stats(int64_t *out) { int k; out[0] = g_wraps; for (k = 0; k < 4; ++k) out[1 + k] = g_rng[k]; }
#define TRACK_UV(v) do { if ((v) < g_rng[0]) g_rng[0] = (v); if ((v) > g_rng[1]) g_rng[1] = (v); } while (0)
#define TRACK_X(v, k) do { if ((v) > g_rng[k]) g_rng[k] = (v); } while (0)

static void ora_reset(ora_extz_t *ez) /* K2H:238-243 */
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0;
	ez->score = ez->mqe = ez->mte = ORA_NEG_INF;
	ez->n_cigar = 0; ez->zdropped = 0; ez->reach_end = 0;
}

/* K2H:245-261 with is_rot=1 */
static int ora_zdrop(ora_extz_t *ez, int32_t H, int r, int t, int zdrop, int8_t e)
{
	if (H > (int32_t)ez->max) {
		ez->max = (uint32_t)H; ez->max_t = t; ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q;
		int l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && (int32_t)ez->max - H > zdrop + l * e) {
			ez->zdropped = 1;
			return 1;
		}
	}
	return 0;
}

/* K2H:106-116 */
static void ora_push(ora_extz_t *ez, uint32_t op, int len)
{
	if (ez->n_cigar == 0 || op != (ez->cigar[ez->n_cigar - 1] & 0xf)) {
		if (ez->n_cigar == ez->m_cigar) {
			ez->m_cigar = ez->m_cigar ? ez->m_cigar << 1 : 4;
			ez->cigar = (uint32_t*)realloc(ez->cigar, (size_t)ez->m_cigar << 2);
		}
		ez->cigar[ez->n_cigar++] = (uint32_t)len << 4 | op;
	} else ez->cigar[ez->n_cigar - 1] += (uint32_t)len << 4;
}

/* K2H:119-151 with is_rot=1, min_intron_len=0.  dir is the per-cell byte matrix laid out
 * as the reference does: row r has `stride` bytes, column index = t - row_lo[r]. */
static void ora_backtrack(ora_extz_t *ez, int rev, const uint8_t *dir, const int *row_lo, const int *row_hi,
                          size_t stride, int i0, int j0)
{
	int i = i0, j = j0, state = 0, k;
	ez->n_cigar = 0;
	while (i >= 0 && j >= 0) {
		int r = i + j, forced = -1;
		uint32_t b;
		if (i < row_lo[r]) forced = 2;
		if (i > row_hi[r]) forced = 1;
		b = forced < 0 ? dir[(size_t)r * stride + (size_t)(i - row_lo[r])] : 0;
		if (state == 0) state = b & 7;
		else if (!((b >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = b & 7;
		if (forced >= 0) state = forced;
		if (state == 0) { ora_push(ez, 0, 1); --i; --j; }
		else if (state == 1 || state == 3) { ora_push(ez, 2, 1); --i; }
		else { ora_push(ez, 1, 1); --j; }
	}
	if (i >= 0) ora_push(ez, 2, i + 1);
	if (j >= 0) ora_push(ez, 1, j + 1);
	if (!rev)
		for (k = 0; k < ez->n_cigar >> 1; ++k) {
			uint32_t tmp = ez->cigar[k];
			ez->cigar[k] = ez->cigar[ez->n_cigar - 1 - k];
			ez->cigar[ez->n_cigar - 1 - k] = tmp;
		}
}

/* Number of DP cells inside the (un-rounded) band: the work unit of SURVEY.md section 8d. */
int64_t ksw_extd2_oracle_cells(int qlen, int tlen, int w)
{
	int64_t n = 0;
	int r;
	if (qlen <= 0 || tlen <= 0) return 0;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int lo = 0, hi = tlen - 1;
		if (lo < r - qlen + 1) lo = r - qlen + 1;
		if (hi > r) hi = r;
		if (lo < ((r - w + 1) >> 1)) lo = (r - w + 1) >> 1;
		if (hi > ((r + w) >> 1)) hi = (r + w) >> 1;
		if (lo > hi) break;
		n += hi - lo + 1;
	}
	return n;
}

void ksw_extd2_oracle(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                      int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ora_extz_t *ez)
{
	const int with_cigar = !(flag & ORA_SCORE_ONLY), approx_max = !!(flag & ORA_APPROX_MAX);
	int T16, Q16, n_blk, stride, r, t, k, max_sc, min_sc, long_thres, long_diff, prev_lo = -1, prev_hi = -1;
	int8_t *U, *V, *X, *Y, *X2, *Y2, *S, sc_mch, sc_mis, sc_N;
	uint8_t *arena, *SF, *QR, *dir = 0, wild;
	int32_t *H = 0, H0 = 0;
	int *row_lo = 0, *row_hi = 0, last_H0_t = 0;
	const int qe_as_passed = q + e;  /* sic: the reference latches q+e BEFORE it swaps the two gap
	                                    pieces (KSW:60 vs KSW:70) and uses that for H at r==0 */
	(void)km;

	ora_reset(ez);
	if (m <= 1 || qlen <= 0 || tlen <= 0) return;                       /* KSW:68 */
	if (q2 + e2 < q + e) { int8_t x; x = q, q = q2, q2 = x; x = e, e = e2, e2 = x; } /* KSW:70 */
	sc_mch = mat[0]; sc_mis = mat[1];
	sc_N = mat[m * m - 1] == 0 ? (int8_t)-e2 : mat[m * m - 1];            /* KSW:80 */
	wild = (uint8_t)(m - 1);
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	T16 = (tlen + 15) / 16 * 16;
	Q16 = (qlen + 15) / 16 * 16;
	n_blk = qlen < tlen ? qlen : tlen;
	n_blk = ((n_blk < w + 1 ? n_blk : w + 1) + 15) / 16 + 1;             /* KSW:86-87 */
	stride = n_blk * 16;
	for (k = 1, max_sc = mat[0], min_sc = mat[1]; k < m * m; ++k) {
		if (mat[k] > max_sc) max_sc = mat[k];
		if (mat[k] < min_sc) min_sc = mat[k];
	}
	(void)max_sc;
	if (-min_sc > 2 * (q + e)) return;                                   /* KSW:93 */

	long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;                  /* KSW:95-98 */
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	long_diff = long_thres * (e - e2) - (q2 - q) - e2;

	/* six difference rows, then  s | target copy | reversed query  back to back, all zeroed */
	U = (int8_t*)malloc((size_t)T16 * 6);
	V = U + T16; X = V + T16; Y = X + T16; X2 = Y + T16; Y2 = X2 + T16;
	memset(U, w8(-q - e), (size_t)T16 * 4);
	memset(X2, w8(-q2 - e2), (size_t)T16 * 2);
	arena = (uint8_t*)calloc((size_t)T16 * 2 + Q16 + 16, 1);
	S = (int8_t*)arena; SF = arena + T16; QR = SF + T16;
	if (!approx_max) {
		H = (int32_t*)malloc((size_t)T16 * sizeof(int32_t));
		for (t = 0; t < T16; ++t) H[t] = ORA_NEG_INF;
	}
	if (with_cigar) {
		dir = (uint8_t*)malloc((size_t)(qlen + tlen - 1) * stride + 16);
		row_lo = (int*)malloc(sizeof(int) * 2 * (size_t)(qlen + tlen - 1));
		row_hi = row_lo + (qlen + tlen - 1);
	}
	for (t = 0; t < qlen; ++t) QR[t] = query[qlen - 1 - t];
	memcpy(SF, target, (size_t)tlen);

	for (r = 0; r < qlen + tlen - 1; ++r) {
		int lo0, hi0, lo, hi;               /* exact band and its 16-rounded hull */
		int8_t cx, cx2, cv;                 /* values shifted in from t = lo-1 */
		const uint8_t *qrr = QR + (qlen - 1 - r);
		lo0 = 0; hi0 = tlen - 1;
		if (lo0 < r - qlen + 1) lo0 = r - qlen + 1;
		if (hi0 > r) hi0 = r;
		if (lo0 < ((r - w + 1) >> 1)) lo0 = (r - w + 1) >> 1;
		if (hi0 > ((r + w) >> 1)) hi0 = (r + w) >> 1;
		if (lo0 > hi0) { ez->zdropped = 1; break; }                      /* KSW:135-138 */
		lo = lo0 / 16 * 16; hi = (hi0 + 16) / 16 * 16 - 1;
		/* left boundary (KSW:142-152) */
		if (lo > 0) {
			if (lo - 1 >= prev_lo && lo - 1 <= prev_hi) { cx = X[lo - 1]; cx2 = X2[lo - 1]; cv = V[lo - 1]; }
			else { cx = w8(-q - e); cx2 = w8(-q2 - e2); cv = w8(-q - e); }
		} else {
			cx = w8(-q - e); cx2 = w8(-q2 - e2);
			cv = r == 0 ? w8(-q - e) : r < long_thres ? w8(-e) : r == long_thres ? w8(long_diff) : w8(-e2);
		}
		/* first-row boundary (KSW:153-156) */
		if (hi >= r) {
			Y[r] = w8(-q - e); Y2[r] = w8(-q2 - e2);
			U[r] = r == 0 ? w8(-q - e) : r < long_thres ? w8(-e) : r == long_thres ? w8(long_diff) : w8(-e2);
		}
		/* substitution scores (KSW:158-177).  16-byte chunks starting at lo0; the chunk may
		 * read past the target copy into QR and may write past S into SF, like the original. */
		if (!(flag & ORA_GENERIC_SC)) {
			for (t = lo0; t <= hi0; t += 16)
				for (k = 0; k < 16; ++k) {
					uint8_t a = SF[t + k], b = qrr[t + k];
					int8_t sc = a == b ? sc_mch : sc_mis;
					if (a == wild || b == wild) sc = sc_N;
					S[t + k] = sc;
				}
		} else {
			for (t = lo0; t <= hi0; ++t) S[t] = mat[SF[t] * m + qrr[t]];
		}
		if (with_cigar) { row_lo[r] = lo; row_hi[r] = hi; }
		/* the cells, left to right; cx/cx2/cv carry last diagonal's values at t-1 */
		for (t = lo; t <= hi; ++t) {
			int8_t z = S[t], xt1 = cx, x2t1 = cx2, vt1 = cv, ut = U[t];
			int8_t a, b, a2, b2, tmp;
			uint8_t d = 0;
			cx = X[t]; cx2 = X2[t]; cv = V[t];
			a = w8(xt1 + vt1); b = w8(Y[t] + ut); a2 = w8(x2t1 + vt1); b2 = w8(Y2[t] + ut);
			if (!(flag & ORA_RIGHT)) {          /* first maximum wins (KSW:225-236) */
				if (a  > z) { d = 1; z = a;  }
				if (b  > z) { d = 2; z = b;  }
				if (a2 > z) { d = 3; z = a2; }
				if (b2 > z) { d = 4; z = b2; }
			} else {                            /* last maximum wins (KSW:272-283) */
				if (!(z > a))  { d = 1; z = a;  }
				if (!(z > b))  { d = 2; z = b;  }
				if (!(z > a2)) { d = 3; z = a2; }
				if (!(z > b2)) { d = 4; z = b2; }
			}
			if (z > sc_mch) z = sc_mch;
			U[t] = w8(z - vt1);
			V[t] = w8(z - ut);
			tmp = w8(z - q);  a  = w8(a  - tmp); b  = w8(b  - tmp);
			tmp = w8(z - q2); a2 = w8(a2 - tmp); b2 = w8(b2 - tmp);
			if (!(flag & ORA_RIGHT)) {          /* continue a gap only if strictly better */
				X[t]  = w8((a  > 0 ? a  : 0) - (q + e));   if (a  > 0) d |= 0x08;
				Y[t]  = w8((b  > 0 ? b  : 0) - (q + e));   if (b  > 0) d |= 0x10;
				X2[t] = w8((a2 > 0 ? a2 : 0) - (q2 + e2)); if (a2 > 0) d |= 0x20;
				Y2[t] = w8((b2 > 0 ? b2 : 0) - (q2 + e2)); if (b2 > 0) d |= 0x40;
			} else {
				X[t]  = w8((a  >= 0 ? a  : 0) - (q + e));   if (a  >= 0) d |= 0x08;
				Y[t]  = w8((b  >= 0 ? b  : 0) - (q + e));   if (b  >= 0) d |= 0x10;
				X2[t] = w8((a2 >= 0 ? a2 : 0) - (q2 + e2)); if (a2 >= 0) d |= 0x20;
				Y2[t] = w8((b2 >= 0 ? b2 : 0) - (q2 + e2)); if (b2 >= 0) d |= 0x40;
			}
			TRACK_UV(U[t]); TRACK_UV(V[t]); TRACK_X(X[t], 2); TRACK_X(Y[t], 2); TRACK_X(X2[t], 3); TRACK_X(Y2[t], 3);
			if (with_cigar) dir[(size_t)r * stride + (size_t)(t - lo)] = d;
		}
		if (!approx_max) {                      /* exact H row and its argmax (KSW:316-359) */
			int32_t best, best_t;
			if (r > 0) {
				int32_t lane_best[4], lane_t[4];
				int vec_end = lo0 + (hi0 - lo0) / 4 * 4;
				best = H[hi0] = hi0 > 0 ? H[hi0 - 1] + U[hi0] : H[hi0] + V[hi0];
				best_t = hi0;
				for (k = 0; k < 4; ++k) { lane_best[k] = best; lane_t[k] = best_t; }
				for (t = lo0; t < vec_end; t += 4)
					for (k = 0; k < 4; ++k) {
						H[t + k] += V[t + k];
						if (H[t + k] > lane_best[k]) { lane_best[k] = H[t + k]; lane_t[k] = t; }
					}
				for (k = 0; k < 4; ++k)
					if (best < lane_best[k]) { best = lane_best[k]; best_t = lane_t[k] + k; }
				for (t = vec_end; t < hi0; ++t) {
					H[t] += V[t];
					if (H[t] > best) { best = H[t]; best_t = t; }
				}
			} else { H[0] = V[0] - qe_as_passed; best = H[0]; best_t = 0; }
			if (hi0 == tlen - 1 && H[hi0] > ez->mte) { ez->mte = H[hi0]; ez->mte_q = r - hi; } /* sic: rounded hi (KSW:354) */
			if (r - lo0 == qlen - 1 && H[lo0] > ez->mqe) { ez->mqe = H[lo0]; ez->mqe_t = lo0; }
			if (ora_zdrop(ez, best, r, best_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && hi0 == tlen - 1) ez->score = H[tlen - 1];
		} else {                                /* diagonal-following approximation (KSW:360-376) */
			if (r > 0) {
				if (last_H0_t >= lo0 && last_H0_t <= hi0 && last_H0_t + 1 >= lo0 && last_H0_t + 1 <= hi0) {
					int32_t d0 = V[last_H0_t], d1 = U[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else { H0 += d1; ++last_H0_t; }
				} else if (last_H0_t >= lo0 && last_H0_t <= hi0) {
					H0 += V[last_H0_t];
				} else {
					++last_H0_t; H0 += U[last_H0_t];
				}
			} else { H0 = V[0] - qe_as_passed; last_H0_t = 0; }
			if ((flag & ORA_APPROX_DROP) && ora_zdrop(ez, H0, r, last_H0_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && hi0 == tlen - 1) ez->score = H0;
		}
		prev_lo = lo; prev_hi = hi;
	}
	free(U); free(arena); free(H);
	if (with_cigar) {                           /* KSW:382-391 */
		int rev = !!(flag & ORA_REV_CIGAR);
		if (!ez->zdropped && !(flag & ORA_EXTZ_ONLY)) {
			ora_backtrack(ez, rev, dir, row_lo, row_hi, (size_t)stride, tlen - 1, qlen - 1);
		} else if (!ez->zdropped && (flag & ORA_EXTZ_ONLY) && ez->mqe + end_bonus > (int)ez->max) {
			ez->reach_end = 1;
			ora_backtrack(ez, rev, dir, row_lo, row_hi, (size_t)stride, ez->mqe_t, qlen - 1);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			ora_backtrack(ez, rev, dir, row_lo, row_hi, (size_t)stride, ez->max_t, ez->max_q);
		}
		free(dir); free(row_lo);
	}
}
