// ksw_host.hpp -- host-side planning for the ksw kernels: parameter normalisation (what the
// reference does before its main loop, ksw2_extd2_sse.c:60-98), the proof that a parameter set
// cannot wrap the reference's int8 arithmetic (which is what lets the fast kernel use 16-bit
// lanes), kernel-variant selection and scratch sizing.  Plain C++, no CUDA types.
#pragma once
#include <stdint.h>
#include <algorithm>
#include "ksw_common.cuh"

namespace kswhost {

struct Plan {
	kswfast::Params P;
	bool trivial;        // reference returns right after ksw_reset_extz (KSW:68 m<=1, KSW:93 mismatch too large)
	bool fast_params;    // flags + scoring admit the team kernel (WRAP variant, always exact)
	bool nowrap_ok;      // ... and unclipped tasks may use the variant without wrap masks
};

// In-band cells of a band that never clips the matrix form a consistent affine-gap DP, for which
//   -(q+e) <= u,v <= mch+q+e,   x,y <= -e,   x2,y2 <= -e2,   z <= mch
// hold (Suzuki-Kasahara).  This checks that under those bounds no int8 operation of
// ksw2_extd2_sse.c:30-58,221-267 can wrap, which is what lets the no-wrap variant of the fast
// kernel skip the wrap masks for such tasks.  (Out-of-band cells DO wrap in the reference, all
// the time; with an unclipped band they never reach an in-band cell.  Clipped tasks always take
// the WRAP variant, which is exact without any assumption.)
static inline bool int8_bounds_hold(int mch, int smin, int smax, int q, int e, int q2, int e2, int long_diff)
{
	const int qe = q + e, qe2 = q2 + e2;
	const int s_lo = std::min(smin, 0), s_hi = std::max(smax, 0);   // calloc'ed s[] is 0 before its first refresh
	const int uL = -qe, uH = mch + qe;
	const int special[3] = {-e, -e2, long_diff};                    // first row / column values (KSW:151,155)
	for (int k = 0; k < 3; ++k) if (special[k] < uL || special[k] > uH) return false;
	if (s_hi > mch) return false;                                   // a score above mat[0] would make the clamp routine
	const int aL = -qe + uL, aH = -e + uH, a2L = -qe2 + uL, a2H = -e2 + uH;
	const int zL = s_lo, zH = mch;
	const int lo[] = {uL, aL, a2L, zL - q, zL - q2, aL - (zH - q), a2L - (zH - q2), -qe, -qe2, zL - uH};
	const int hi[] = {uH, aH, a2H, zH - q, zH - q2, aH - (zL - q), a2H - (zL - q2), zH - uL};
	for (int v : lo) if (v < -128) return false;
	for (int v : hi) if (v > 127) return false;
	return true;
}

static inline Plan make_plan(int m, const int8_t *mat, int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag)
{
	Plan pl;
	kswfast::Params &P = pl.P;
	pl.trivial = false; pl.fast_params = false; pl.nowrap_ok = false;
	q = (int8_t)q; e = (int8_t)e; q2 = (int8_t)q2; e2 = (int8_t)e2;  // the reference takes int8_t arguments
	P.qe_as_passed = q + e;
	if (q2 + e2 < q + e) { std::swap(q, q2); std::swap(e, e2); }
	P.one = 1u; P.mone = 0xffffffffu;
	P.wild = m - 1; P.w = w; P.zdrop = zdrop; P.end_bonus = end_bonus; P.flag = flag;
	P.q = q; P.e = e; P.q2 = q2; P.e2 = e2;
	P.sc_mch = P.sc_mis = P.sc_N = 0; P.long_thres = P.long_diff = 0;
	if (m <= 1) { pl.trivial = true; return pl; }
	P.sc_mch = mat[0]; P.sc_mis = mat[1];
	P.sc_N = mat[m * m - 1] == 0 ? (int8_t)-e2 : mat[m * m - 1];
	int min_sc = mat[1];
	for (int k = 1; k < m * m; ++k) min_sc = std::min<int>(min_sc, mat[k]);
	if (-min_sc > 2 * (q + e)) { pl.trivial = true; return pl; }
	P.long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + P.long_thres * e2 > q + e + P.long_thres * e) ++P.long_thres;
	P.long_diff = (int8_t)(P.long_thres * (e - e2) - (q2 - q) - e2);
	const int allowed = kswfast::F_SCORE_ONLY | kswfast::F_EXTZ_ONLY | kswfast::F_REV_CIGAR;
	const int smin = std::min(P.sc_mch, std::min(P.sc_mis, P.sc_N)), smax = std::max(P.sc_mch, std::max(P.sc_mis, P.sc_N));
	pl.fast_params = !(flag & ~allowed) && m <= 16 && q >= 0 && e >= 0 && q2 >= 0 && e2 >= 0 && q + e <= 127 && q2 + e2 <= 127;
	// The bounds of int8_bounds_hold assume first-row / first-column values that are consistent with the recurrence.  With
	// e < e2 the reference's long_thres is negative and its boundary steps by -e2 where the recurrence extends by -e
	// (KSW:151,155): the differences then leave those bounds inside the band, the reference wraps, and only the WRAP variant
	// reproduces it (found by tests/soak_aln.py with -O 22 -E 3 -P 14 -F 0).
	pl.nowrap_ok = pl.fast_params && P.sc_mch >= 0 && e >= e2 && int8_bounds_hold(P.sc_mch, smin, smax, q, e, q2, e2, P.long_diff);
	return pl;
}

// Anti-diagonals the reference executes before its band closes (KSW:124-138); upper bound on
// traceback rows (a z-drop can only end earlier).  (This and the next three run on the host and, for the per-task plan of a
// batch, on the device.)
LANE_HD int n_diagonals(int qlen, int tlen, int w)
{
	if (qlen <= 0 || tlen <= 0) return 0;
	if (w < 0) w = qlen > tlen ? qlen : tlen;
	// lo0 > hi0 first happens where r-qlen+1 > (r+w)>>1 or (r-w+1)>>1 > tlen-1 (the other pairs cannot cross)
	int n = qlen + tlen - 1;
	for (int r = 0; r < n; ++r) {
		int lo0, hi0;
		kswfast::band(r, qlen, tlen, w, lo0, hi0);
		if (lo0 > hi0) return r;
	}
	return n;
}

// 16-cell blocks per diagonal, as the reference sizes its traceback rows (KSW:86-87)
LANE_HD int n_col_blocks(int qlen, int tlen, int w)
{
	if (w < 0) w = qlen > tlen ? qlen : tlen;
	int n = qlen < tlen ? qlen : tlen;
	if (n > w + 1) n = w + 1;
	return (n + 15) / 16 + 1;
}

// does the band ever cut the matrix?  (unclipped iff qlen,tlen <= w+1)
LANE_HD bool band_clips(int qlen, int tlen, int w) { return w >= 0 && (qlen > w + 1 || tlen > w + 1); }

// lanes per alignment of the team kernel: one lane per 16-cell block of the widest rounded band
// (n_col_blocks), rounded up to a power of two.  0 = wider than a warp.
LANE_HD int pick_team(int qlen, int tlen, int w)
{
	const int need = n_col_blocks(qlen, tlen, w);
	for (int t = 2; t <= 32; t *= 2)
		if (t >= need) return t;
	return 0;
}

} // namespace kswhost
