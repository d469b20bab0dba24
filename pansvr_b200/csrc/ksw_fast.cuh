// ksw_fast.cuh -- warp-per-alignment banded two-piece-affine DP with on-device traceback.
//
// Re-implements, bit for bit, what the reference computes in
//   /root/reference/src/kswlib/ksw2_extd2_sse.c:26-396  (ksw_extd2_sse)
//   /root/reference/src/kswlib/ksw2.h:106-151,238-261   (ksw_push_cigar, ksw_backtrack_D, zdrop)
// for flag subsets of {SCORE_ONLY, EXTZ_ONLY, REV_CIGAR} with the match/mismatch/wildcard scoring
// (everything `panSVR fc_aln` and `fc_sv` ever ask for: read_realignment.cpp:889,
// SignalAssembly.hpp:463).  Other flags and parameter sets whose int8 arithmetic could wrap go to
// the generic kernel (ksw_generic.cuh).
//
// What is reproduced is the reference's *machine*, not the textbook recurrence:
//   * anti-diagonal r updates the 16-cell blocks [st,en] that hull the band [st0,en0]; cells
//     outside the band are still computed, from stale rows, and feed later diagonals (KSW:139-267)
//   * the substitution row is refreshed only on [st0, st0+16*ceil((en0-st0+1)/16))  (KSW:158-173)
//   * first maximum wins in  H > E > F > E2 > F2                                  (KSW:225-236)
//   * exact int32 H side row with the 4-lane argmax tie order of the SSE scan       (KSW:316-351)
//
// Int8 wrap-around.  The reference's out-of-band cells run an unstable recurrence whose int8
// values routinely wrap (hundreds of wraps per 150x1100 task), and with a band that clips the
// matrix those cells feed the in-band ones.  The WRAP variant therefore reproduces the wrap
// exactly: values are kept as value*8+bias with bias = 0x400 (mod 0x800), so `& 0x07ff07ff` IS the
// int8 wrap (one LOP3 per packed pair), applied before every comparison the reference makes.
// When the band never clips (qlen,tlen <= w+1) out-of-band cells cannot reach an in-band cell,
// in-band values are bounded (ksw_host.hpp: int8_bounds_hold) and the masks are dropped.
//
// B200 mapping.  One warp owns one alignment.  The seven int8 rows of the reference (u v x y x2
// y2 s) live in registers, CPL cells per lane, two cells per 32-bit register as unsigned 16-bit
// halves holding value*8+bias; cell t belongs to lane (t/CPL)%32, i.e. the band slides through
// the warp cyclically and a lane is re-initialised for cell group g+32 when group g falls out of
// the window [st, st+32*CPL).  Per pair of cells the inner loop is
//   3 PRMT (neighbour shift) + 4 IADD3 (a b a2 b2, priority tag folded in) + 2 VIMNMX3.U16x2
//   (5-way argmax: the tag in the low 3 bits makes the first maximum win) + LOP3/VIMNMX (strip,
//   clamp) + 2 IADD3 (u v) + 4 IADD3 + 4 VIADDMNMX.U16x2 (the four gap rows, stored as x+q+e so the
//   ReLU is the whole update) + 4 VIMNMX + 4 IMAD/IADD3 (continuation bits -> traceback byte).
// The left-neighbour cell comes from lane-1 with two SHFL per diagonal; H lives in shared memory
// (4 B/cell); the traceback byte matrix (1 B/cell, row = diagonal, column = t mod W) is streamed
// to a per-warp scratch in HBM with one coalesced store per lane per diagonal and walked by the
// whole warp speculatively (32 cells of a run per round trip) to emit the CIGAR.
#pragma once
#include "lane_rt.cuh"

namespace kswfast {
using namespace lanert;

enum { F_SCORE_ONLY = 0x01, F_RIGHT = 0x02, F_GENERIC_SC = 0x04, F_APPROX_MAX = 0x08, F_APPROX_DROP = 0x10,
       F_EXTZ_ONLY = 0x40, F_REV_CIGAR = 0x80 };
enum { NEG_INF = -0x40000000 };
enum { RES_WORDS = 12 }; // max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar reach_end status
// status: bit 0 = CIGAR longer than cigar_cap

// Biases of the packed representation (per 16-bit half); a half holds true*8 + bias.
//   bU/bV/bM/bK: stored u / stored v / stored gap rows (x+q+e ...) / 5-way-max keys and z
//   pA,pB,pV,pT: bias of a sum before it is wrapped (WRAP) or used (no wrap); with WRAP every p*
//   is 0x400 mod 0x800 and large enough that the half cannot go negative.
template <bool WRAP> struct Bias;
template <> struct Bias<false> { enum { bU = 0x2000, bV = 0x2000, bM = 0x4000, bK = 0x2000, pA = 0x2000, pB = 0x2000, pV = 0x2000, pT = 0x4000 }; };
template <> struct Bias<true>  { enum { bU = 0x2400, bV = 0x0400, bM = 0x0400, bK = 0x0400, pA = 0x0c00, pB = 0x2c00, pV = 0x1400, pT = 0x0c00 }; };
enum { QS_PAD = 1 };      // QS[0] = 0 (j<0), QS[1..qlen] = query, QS[qlen+1] = 0

struct Params {           // one per batch, filled by the host (ksw_batch.cu: make_params)
	int wild;             // m-1
	int w, zdrop, end_bonus, flag;
	int q, e, q2, e2;     // after the reference's swap (KSW:70)
	int qe_as_passed;     // q+e before the swap (KSW:60)
	int long_thres, long_diff;
	int sc_mch, sc_mis, sc_N;
};

LANE_FN uint32_t k32(int v) { return (uint32_t)((int64_t)v * 65537); } // add v to both halves with a 32-bit add

// band of anti-diagonal r before rounding (KSW:131-134)
LANE_HD void band(int r, int qlen, int tlen, int w, int &lo0, int &hi0)
{
	lo0 = 0; hi0 = tlen - 1;
	if (lo0 < r - qlen + 1) lo0 = r - qlen + 1;
	if (hi0 > r) hi0 = r;
	if (lo0 < ((r - w + 1) >> 1)) lo0 = (r - w + 1) >> 1;
	if (hi0 > ((r + w) >> 1)) hi0 = (r + w) >> 1;
}

// CPL traceback bytes (already packed 4 per word) to p, which is CPL-byte aligned
template <int CPL>
LANE_FN void store_cells(uint8_t *p, const uint32_t *w)
{
#ifndef PANSVR_HOST_EMUL
	if (CPL == 2) *(uint16_t*)p = (uint16_t)w[0];
	else if (CPL == 4) *(uint32_t*)p = w[0];
	else if (CPL == 8) *(uint2*)p = make_uint2(w[0], w[1]);
	else {
#pragma unroll
		for (int i = 0; i < CPL / 4; i += 4) *(uint4*)(p + 4 * i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
	}
#else
	for (int i = 0; i < CPL; ++i) p[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
#endif
}

LANE_FN uint32_t enc_t(uint32_t b, int wild) { return (int)b == wild ? 0x10u : (b & 0xfu); }
LANE_FN uint32_t enc_q(uint32_t b, int wild) { return (int)b == wild ? 0x20u : (b & 0xfu); }

// One alignment, executed by all 32 lanes of a warp.
//   Hs : W int32 of shared memory (this warp's), QS : >= qlen+2 bytes of shared memory (this warp's)
//   tb : this warp's traceback scratch, >= n_diagonals * W bytes (unused with SCORE_ONLY)
template <int CPL, bool WRAP>
LANE_DEV void align_task(const Params &P, int qlen, const uint8_t *__restrict__ query, int tlen,
                         const uint8_t *__restrict__ target, int32_t *__restrict__ res, uint32_t *__restrict__ cigar,
                         int cigar_cap, uint8_t *__restrict__ tb, int32_t *Hs, uint8_t *QS)
{
	constexpr int NR = CPL / 2;          // registers per row per lane
	constexpr int W = 32 * CPL;          // cells resident in the warp
	const int lane = lane_id();
	const int w = P.w < 0 ? (tlen > qlen ? tlen : qlen) : P.w;
	const bool with_cigar = !(P.flag & F_SCORE_ONLY);
	const int q8 = P.q * 8, qe8 = (P.q + P.e) * 8, q28 = P.q2 * 8, qe28 = (P.q2 + P.e2) * 8;

	// packed constants (k32: added with a 32-bit add; dup16: operand of a 16x2 min/max)
	typedef Bias<WRAP> BB;
	constexpr int bU = BB::bU, bV = BB::bV, bM = BB::bM, bK = BB::bK;
	const uint32_t WM = 0x07ff07ffu;                                   // the int8 wrap (WRAP only)
	const uint32_t CA = k32(3 + BB::pA - qe8 - bM - bV), CB = k32(2 + BB::pB - qe8 - bM - bU);
	const uint32_t CA2 = k32(1 + BB::pA - qe28 - bM - bV), CB2 = k32(0 + BB::pB - qe28 - bM - bU);
	const uint32_t MCH = dup16(P.sc_mch * 8 + bK), CU = k32(bU - bK + bV), CV = k32(BB::pV - bK + bU);
	const uint32_t CNA = k32(q8 - 3 + BB::pT), CNB = k32(q8 - 2 + BB::pT), CNA2 = k32(q28 - 1 + BB::pT), CNB2 = k32(q28 + BB::pT);
	const uint32_t BMd = dup16(bM), BM8d = dup16(bM + 8), CTB = k32(-15 * bM);
	const uint32_t SBASE = k32(P.sc_mch * 8 + 4 + bK), ONE2 = 0x00010001u;
	const int D1 = (P.sc_mis - P.sc_mch) * 8, E2 = (P.sc_N - P.sc_mis) * 8;
	const uint32_t U_DEF = dup16(-qe8 + bU), V_DEF = dup16(-qe8 + bV), S_ZERO = dup16(4 + bK);

	// ---- stage the query in shared memory, wildcard-encoded, zero-padded on both sides
	for (int j = lane; j < qlen + 2; j += 32)
		QS[j] = (j >= 1 && j <= qlen) ? (uint8_t)enc_q(query[j - 1], P.wild) : (uint8_t)0;
	for (int k = lane; k < W; k += 32) Hs[k] = 0;
	wsync();

	// ---- per-lane rows
	uint32_t U[NR], V[NR], MX[NR], MY[NR], MX2[NR], MY2[NR], S[NR], TB[NR], QB[NR];
	int t0 = lane * CPL;
	auto load_group = [&](int r_for_q) __attribute__((always_inline)) {   // (re)initialise this lane for the cell group starting at t0
#pragma unroll
		for (int i = 0; i < NR; ++i) {
			U[i] = U_DEF; V[i] = V_DEF;
			MX[i] = MY[i] = MX2[i] = MY2[i] = BMd;
			S[i] = S_ZERO;
			uint32_t tl = 0, th = 0, ql, qh;
			int ta = t0 + 2 * i, tbb = ta + 1;
			if (ta < tlen) tl = enc_t(target[ta], P.wild);
			if (tbb < tlen) th = enc_t(target[tbb], P.wild);
			TB[i] = pk(tl, th);
			int ja = r_for_q - ta, jb = r_for_q - tbb;       // query index of the cell on diagonal r_for_q
			ja = ja < -1 ? -1 : (ja > qlen ? qlen : ja);
			jb = jb < -1 ? -1 : (jb > qlen ? qlen : jb);
			ql = QS[ja + QS_PAD]; qh = QS[jb + QS_PAD];
			QB[i] = pk(ql, qh);
		}
	};
	load_group(-1);

	// ---- running ez (H-like quantities are kept scaled by 8)
	int ez_max8 = 0, ez_max_t = -1, ez_max_q = -1, mqe8 = NEG_INF, mqe_t = -1, mte8 = NEG_INF, mte_q = -1, score8 = NEG_INF;
	int zdropped = 0, last_st = -1, Hprev = 0;
	const int n_diag = qlen + tlen - 1;

	for (int r = 0; r < n_diag; ++r) {
		int st0, en0;
		band(r, qlen, tlen, w, st0, en0);
		if (st0 > en0) { zdropped = 1; break; }
		const int st = st0 & ~15, en = en0 | 15;

		// -- left-neighbour exchange on the rows as they stand after diagonal r-1
		uint32_t snd1 = prmt(MX[NR - 1], V[NR - 1], 0x7632);     // lo = MX of my last cell, hi = V of it
		uint32_t rcv1 = shfl(snd1, (lane + 31) & 31);
		uint32_t rcv2 = shfl(MX2[NR - 1], (lane + 31) & 31);     // hi half = MX2 of the neighbour's last cell

		// -- window slide: a lane whose group fell left of st takes over the group W cells further right
		if (t0 + CPL <= st) {
			t0 += W;
			load_group(r);
		} else {                                                 // query window moves one cell per diagonal
			int j = r - t0;
			j = j < -1 ? -1 : (j > qlen ? qlen : j);
			uint32_t nq = QS[j + QS_PAD];
#pragma unroll
			for (int i = NR - 1; i > 0; --i) QB[i] = prmt(QB[i - 1], QB[i], 0x5432);
			QB[0] = (QB[0] << 16) | nq;
		}

		// -- boundary of the first group of the window (KSW:142-152)
		const int uval = r == 0 ? -(P.q + P.e) : r < P.long_thres ? -P.e : r == P.long_thres ? P.long_diff : -P.e2;
		if (t0 == st && !(st > 0 && st != last_st)) {
			int bv = st > 0 ? -(P.q + P.e) : uval;
			rcv1 = pk(bM, (uint32_t)(bv * 8 + bV));
			rcv2 = BMd;
		}
		// -- first-row cell t = r (KSW:153-156)
		if (en >= r && r >= t0 && r < t0 + CPL) {
			const int k = r - t0;
			const uint32_t hm = (k & 1) ? 0xffff0000u : 0x0000ffffu;
			const uint32_t uv = dup16(uval * 8 + bU);
#pragma unroll
			for (int i = 0; i < NR; ++i) {                       // masks, not indices: keeps the rows in registers
				const uint32_t mi = (k >> 1) == i ? hm : 0u;
				MY[i] = (MY[i] & ~mi) | (BMd & mi);
				MY2[i] = (MY2[i] & ~mi) | (BMd & mi);
				U[i] = (U[i] & ~mi) | (uv & mi);
			}
		}

		// -- substitution scores on [st0, st0 + 16*ceil((en0-st0+1)/16))  (KSW:158-173)
		{
			const int rs_end = st0 + (((en0 - st0) >> 4) + 1) * 16;
			int lo = st0 - t0, hi = rs_end - t0;
			lo = lo < 0 ? 0 : (lo > CPL ? CPL : lo);
			hi = hi < 0 ? 0 : (hi > CPL ? CPL : hi);
			if (hi > lo) {
				const uint32_t cellmask = ((1u << hi) - 1u) & ~((1u << lo) - 1u);   // bit k = cell k refreshed
#pragma unroll
				for (int i = 0; i < NR; ++i) {
					uint32_t x = TB[i] ^ QB[i];
					uint32_t m1 = minu(x, ONE2);
					uint32_t m2 = minu(x & 0x00300030u, ONE2);
					uint32_t sn = SBASE + m1 * (uint32_t)D1 + m2 * (uint32_t)E2;
					uint32_t b2 = (cellmask >> (2 * i)) & 3u;
					uint32_t hm = (b2 & 1u ? 0x0000ffffu : 0u) | (b2 & 2u ? 0xffff0000u : 0u);
					S[i] = (S[i] & ~hm) | (sn & hm);
				}
			}
		}

		// -- the cells of this lane, right to left so that [i-1] is still last diagonal's
		const bool active = t0 >= st && t0 <= en;
		if (active) {
			uint32_t tbw[NR];
#pragma unroll
			for (int i = NR - 1; i >= 0; --i) {
				const uint32_t mxt1 = i > 0 ? prmt(MX[i > 0 ? i - 1 : 0], MX[i], 0x5432) : prmt(rcv1, MX[0], 0x5410);
				const uint32_t vt1 = i > 0 ? prmt(V[i > 0 ? i - 1 : 0], V[i], 0x5432) : prmt(rcv1, V[0], 0x5432);
				const uint32_t mx2t1 = i > 0 ? prmt(MX2[i > 0 ? i - 1 : 0], MX2[i], 0x5432) : prmt(rcv2, MX2[0], 0x5432);
				const uint32_t ut = U[i];
				uint32_t A = mxt1 + vt1 + CA, Bv = MY[i] + ut + CB;   // a b a2 b2 with their priority tags
				uint32_t A2 = mx2t1 + vt1 + CA2, B2 = MY2[i] + ut + CB2;
				if (WRAP) { A &= WM; Bv &= WM; A2 &= WM; B2 &= WM; }
				uint32_t zk = max3u(S[i], A, Bv);
				zk = max3u(zk, A2, B2);
				const uint32_t Z = minu(zk & 0xfff8fff8u, MCH);
				U[i] = Z - vt1 + CU;
				V[i] = Z - ut + CV;
				if (WRAP) {
					V[i] &= WM;
					MX[i] = maxu((A - Z + CNA) & WM, BMd);
					MY[i] = maxu((Bv - Z + CNB) & WM, BMd);
					MX2[i] = maxu((A2 - Z + CNA2) & WM, BMd);
					MY2[i] = maxu((B2 - Z + CNB2) & WM, BMd);
				} else {
					MX[i] = addmaxu(A, CNA - Z, BMd);
					MY[i] = addmaxu(Bv, CNB - Z, BMd);
					MX2[i] = addmaxu(A2, CNA2 - Z, BMd);
					MY2[i] = addmaxu(B2, CNB2 - Z, BMd);
				}
				if (with_cigar)
					tbw[i] = (zk & 0x00070007u) + minu(MX[i], BM8d) + 2u * minu(MY[i], BM8d) + 4u * minu(MX2[i], BM8d)
					         + 8u * minu(MY2[i], BM8d) + CTB;
			}
			if (with_cigar) {                                     // one traceback byte per cell, column = t mod W
				uint32_t pk4[NR >= 2 ? NR / 2 : 1];
				if (NR == 1) pk4[0] = prmt(tbw[0], 0, 0x4420);
#pragma unroll
				for (int i = 0; i < NR / 2; ++i) pk4[i] = prmt(tbw[2 * i], tbw[2 * i + 1], 0x6420);
				store_cells<CPL>(tb + (size_t)r * W + (t0 & (W - 1)), pk4);
			}
		}

		// -- exact H row (KSW:316-351), scaled by 8, in shared memory at column t mod W
		int maxH8, max_t;
		int Hen8, Hst8;
		if (r > 0) {
			int lmax = INT32_MIN;
			int Hk[CPL];
			const int klo = st0 - t0 < 0 ? 0 : st0 - t0;            // cells [klo,khi) of this lane are in [st0,en0)
			const int khi = en0 - t0 > CPL ? CPL : en0 - t0;
			if (active) {
				int32_t *hp = Hs + (t0 & (W - 1));
				if (klo == 0 && khi == CPL) {                        // H[t] += v[t] on [st0,en0) only
#pragma unroll
					for (int i = 0; i < NR; ++i) {
						Hk[2 * i] = hp[2 * i] + (int)lo16u(V[i]) - bV;
						Hk[2 * i + 1] = hp[2 * i + 1] + (int)hi16u(V[i]) - bV;
						hp[2 * i] = Hk[2 * i];
						hp[2 * i + 1] = Hk[2 * i + 1];
						lmax = max3s(lmax, Hk[2 * i], Hk[2 * i + 1]);
					}
				} else {
#pragma unroll
					for (int k = 0; k < CPL; ++k) {
						Hk[k] = INT32_MIN;
						if (k >= klo && k < khi) {
							const uint32_t reg = V[k >> 1];
							Hk[k] = hp[k] + (int)((k & 1) ? hi16u(reg) : lo16u(reg)) - bV;
							hp[k] = Hk[k];
							if (Hk[k] > lmax) lmax = Hk[k];
						}
					}
				}
				if (en0 >= t0 && en0 < t0 + CPL) {                 // the special last element (KSW:322)
					const int k = en0 - t0;
					uint32_t reg = 0;
#pragma unroll
					for (int i = 0; i < NR; ++i) reg |= (en0 > 0 ? U[i] : V[i]) & ((k >> 1) == i ? 0xffffffffu : 0u);
					int d = (int)((k & 1) ? hi16u(reg) : lo16u(reg));
					if (en0 > 0) d = WRAP ? (d & 0x7ff) - 0x400 : d - bU;   // stored u is not wrapped yet
					else d -= bV;
					Hs[en0 & (W - 1)] = Hprev + d;
				}
			}
			wsync();
			Hen8 = Hs[en0 & (W - 1)];
			Hst8 = Hs[st0 & (W - 1)];
			maxH8 = wmax(lmax);
			max_t = en0;
			if (maxH8 > Hen8) {                                    // somebody beats H[en0]: replay the SSE tie order
				int pref = INT32_MIN;
				const int en1 = st0 + ((en0 - st0) >> 2 << 2);
				if (active && lmax == maxH8) {
#pragma unroll
					for (int k = 0; k < CPL; ++k)
						if (k >= klo && k < khi && Hk[k] == maxH8) {
							const int t = t0 + k, d = t - st0;
							const int p = t < en1 ? ((3 - (d & 3)) << 12) + (4095 - (d >> 2)) : -1 - (t - en1);
							if (p > pref) pref = p;
						}
				}
				pref = wmax(pref);
				max_t = pref >= 0 ? st0 + (4095 - (pref & 4095)) * 4 + (3 - (pref >> 12)) : en1 + (-1 - pref);
			} else maxH8 = Hen8;
		} else {                                                  // r == 0 (KSW:351)
			if (lane == 0) Hs[0] = (int)lo16u(V[0]) - bV - P.qe_as_passed * 8;
			wsync();
			Hen8 = Hst8 = maxH8 = Hs[0];
			max_t = 0;
		}
		{   // H[en0'-1] (or H[0]) as it stands now is what the next diagonal's last element starts from
			int st1, en1n;
			band(r + 1, qlen, tlen, w, st1, en1n);
			Hprev = (r + 1 < n_diag && st1 <= en1n) ? Hs[(en1n > 0 ? en1n - 1 : 0) & (W - 1)] : 0;
		}
		wsync();

		// -- ez bookkeeping (KSW:353-359), uniform across the warp
		if (en0 == tlen - 1 && Hen8 > mte8) { mte8 = Hen8; mte_q = r - en; }
		if (r - st0 == qlen - 1 && Hst8 > mqe8) { mqe8 = Hst8; mqe_t = st0; }
		if (maxH8 > ez_max8) { ez_max8 = maxH8; ez_max_t = max_t; ez_max_q = r - max_t; }
		else if (max_t >= ez_max_t && r - max_t >= ez_max_q) {
			const int tl = max_t - ez_max_t, ql = (r - max_t) - ez_max_q;
			const int l = tl > ql ? tl - ql : ql - tl;
			if (P.zdrop >= 0 && ez_max8 - maxH8 > (P.zdrop + l * P.e2) * 8) { zdropped = 1; break; }
		}
		if (r == n_diag - 1 && en0 == tlen - 1) score8 = Hen8;
		last_st = st;
	}

	// ---- traceback (KSW:382-391, K2H:119-151), the whole warp walks the path
	int n_cigar = 0, reach_end = 0, overflow = 0;
	const int ez_max = ez_max8 >> 3;
	const int mqe = mqe8 == NEG_INF ? NEG_INF : mqe8 >> 3;
	if (with_cigar) {
		int i = -1, j = -1;
		if (!zdropped && !(P.flag & F_EXTZ_ONLY)) { i = tlen - 1; j = qlen - 1; }
		else if (!zdropped && (P.flag & F_EXTZ_ONLY) && mqe + P.end_bonus > ez_max) { reach_end = 1; i = mqe_t; j = qlen - 1; }
		else if (ez_max_t >= 0 && ez_max_q >= 0) { i = ez_max_t; j = ez_max_q; }
		wsync();                                                  // traceback bytes written by other lanes
		int state = 0;
		uint32_t cur = 0;                                         // last CIGAR element, not yet stored
		auto push = [&](uint32_t op, int len) __attribute__((always_inline)) {
			if (n_cigar == 0 || op != (cur & 0xfu)) {
				if (n_cigar > 0) {
					if (n_cigar - 1 < cigar_cap) { if (lane == 0) cigar[n_cigar - 1] = cur; }
					else overflow = 1;
				}
				++n_cigar;
				cur = (uint32_t)len << 4 | op;
			} else cur += (uint32_t)len << 4;
		};
		while (i >= 0 && j >= 0) {
			const int di = (state == 0 || state == 1 || state == 3) ? 1 : 0;
			const int dj = (state == 0 || state == 2 || state == 4) ? 1 : 0;
			const int li = i - lane * di, lj = j - lane * dj;
			const bool valid = li >= 0 && lj >= 0;
			int forced = -1;
			uint32_t tmp = 0;
			bool clean = false;
			if (valid) {
				const int rr = li + lj;
				int lo0, hi0;
				band(rr, qlen, tlen, w, lo0, hi0);
				if (li < (lo0 & ~15)) forced = 2;
				if (li > (hi0 | 15)) forced = 1;
				if (forced < 0) {
#ifdef PANSVR_HOST_EMUL
					const uint32_t b = tb[(size_t)rr * W + (li & (W - 1))];
#else
					const uint32_t b = __ldcg(tb + (size_t)rr * W + (li & (W - 1)));
#endif
					tmp = (b & 0x78u) | (4u - (b & 7u));
				}
				clean = forced < 0 && (state == 0 ? (tmp & 7u) == 0 : ((tmp >> (state + 2)) & 1u) != 0);
			}
			const uint32_t stop = wballot(!clean);
			const int n = stop ? ffs32(stop) - 1 : 32;
			if (n > 0) {
				push(state == 0 ? 0u : (di ? 2u : 1u), n);
				i -= n * di; j -= n * dj;
			}
			if (n < 32) {                                         // the first cell that breaks the run, if it exists
				const int v = shfl((int)valid, n);
				const uint32_t t2 = shfl(tmp, n);
				const int f2 = shfl(forced, n);
				if (v) {
					int s2 = state;
					if (s2 == 0) s2 = t2 & 7;
					else if (!((t2 >> (s2 + 2)) & 1)) s2 = 0;
					if (s2 == 0) s2 = t2 & 7;
					if (f2 >= 0) s2 = f2;
					if (s2 == 0) { push(0, 1); --i; --j; }
					else if (s2 == 1 || s2 == 3) { push(2, 1); --i; }
					else { push(1, 1); --j; }
					state = s2;
				}
			}
		}
		if (i >= 0) push(2, i + 1);
		if (j >= 0) push(1, j + 1);
		if (n_cigar > 0) {
			if (n_cigar - 1 < cigar_cap) { if (lane == 0) cigar[n_cigar - 1] = cur; }
			else overflow = 1;
		}
		if (!(P.flag & F_REV_CIGAR) && !overflow) {
			wsync();
			for (int k = lane; k < n_cigar >> 1; k += 32) {
				const uint32_t a = cigar[k], b = cigar[n_cigar - 1 - k];
				cigar[k] = b; cigar[n_cigar - 1 - k] = a;
			}
		}
	}
	for (int k = n_cigar + lane; k < cigar_cap; k += 32) cigar[k] = 0;   // deterministic tail of the CIGAR row
	if (lane == 0) {
		res[0] = ez_max; res[1] = zdropped; res[2] = ez_max_q; res[3] = ez_max_t;
		res[4] = mqe; res[5] = mqe_t; res[6] = mte8 == NEG_INF ? NEG_INF : mte8 >> 3; res[7] = mte_q;
		res[8] = score8 == NEG_INF ? NEG_INF : score8 >> 3; res[9] = n_cigar; res[10] = reach_end;
		res[11] = overflow;
	}
	wsync();
}

} // namespace kswfast
