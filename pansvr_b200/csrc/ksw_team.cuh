// ksw_team.cuh -- banded two-piece-affine DP with on-device traceback; one lane per 16-cell
// block, 32/TEAM alignments per warp.
//
// Re-implements, bit for bit, what the reference computes in
//   /root/reference/src/kswlib/ksw2_extd2_sse.c:26-396  (ksw_extd2_sse)
//   /root/reference/src/kswlib/ksw2.h:106-151,238-261   (ksw_push_cigar, ksw_backtrack_D, zdrop)
// for flag subsets of {SCORE_ONLY, EXTZ_ONLY, REV_CIGAR} with match/mismatch/wildcard scoring
// (all that `panSVR fc_aln` and `fc_sv` ask for: read_realignment.cpp:889, SignalAssembly.hpp:463).
//
// What is reproduced is the reference's *machine*, not the textbook recurrence:
//   * anti-diagonal r updates the 16-cell blocks [st,en] that hull the band [st0,en0]; cells
//     outside the band are still computed, from stale rows, and feed later diagonals (KSW:139-267)
//   * the substitution row is refreshed only on [st0, st0+16*ceil((en0-st0+1)/16))  (KSW:158-173)
//   * first maximum wins in  H > E > F > E2 > F2                                  (KSW:225-236)
//   * exact int32 H side row with the 4-lane argmax tie order of the SSE scan       (KSW:316-351)
//   * int8 wrap-around: the out-of-band cells run an unstable recurrence whose values wrap all
//     the time, and with a clipping band they feed in-band cells.  WRAP keeps values as
//     value*8+bias with bias = 0x400 (mod 0x800) so that `& 0x07ff07ff` IS the int8 wrap (one LOP3
//     per pair of cells) before every comparison the reference makes; unclipped tasks never let an
//     out-of-band value reach an in-band cell and run without the masks (ksw_host.hpp).
//
// B200 mapping.  A lane is one __m128i of the reference: it owns the 16-cell block b of the
// seven rows (u v x y x2 y2 s) in registers, two cells per register as unsigned 16-bit halves, and
// block b lives in lane b % TEAM of its team; when the band's first block moves on, the lane is
// re-initialised for block b+TEAM.  TEAM = widest band in blocks (2..32), so 32/TEAM alignments
// share a warp and every per-diagonal scalar step is paid once for all of them.  Per pair of cells:
//   3 PRMT (neighbour shift) + 4 IADD3 (a b a2 b2, priority tag folded in) [+4 LOP3 wrap]
//   + 2 VIMNMX3.U16x2 (5-way argmax, tag in the low 3 bits = first maximum wins) + LOP3 + VIMNMX
//   + 2 IADD3 (u v) [+1 LOP3] + 4 IADD3 [+4 LOP3] + 4 VIMNMX (gap rows are stored as x+q+e, so the
//   ReLU is the whole update; VIADDMNMX without WRAP) + 4 VIMNMX + 4 IMAD/IADD3 (traceback byte).
// The left neighbour's last cell arrives with two SHFL per diagonal.  The exact H row lives in
// shared memory (cells that are not in [st0,en0) hold -2^30 so the update and its maximum need no
// masks); the argmax position is only materialised when a z-drop could fire or at the end (the H
// row of the diagonal that set the running maximum is snapshotted instead).  The traceback byte
// matrix (row = diagonal, column = t mod 16*TEAM) streams to HBM with one 16-byte store per lane
// per diagonal and is walked by the whole warp speculatively, 32 cells of a run per round trip.
#pragma once
#include "ksw_common.cuh"

namespace kswteam {
using namespace lanert;
using kswfast::Params;
using kswfast::band;
using kswfast::k32;
using kswfast::fadd;
using kswfast::fsub;
using kswfast::enc_t;
using kswfast::enc_q;
using kswfast::QS_PAD;
using kswfast::NEG_INF;
using kswfast::RES_WORDS;

enum { NEG_BIG = -0x40000000 };   // H of a cell that is not in [st0,en0)

template <int TEAM> LANE_FN int team_max(int v)
{
	if (TEAM == 32) return wmax(v);
#pragma unroll
	for (int o = TEAM / 2; o > 0; o >>= 1) { const int x = shfl_xor(v, o); v = x > v ? x : v; }
	return v;
}

// Refresh-mask table: row a (0..15) holds, for the 8 registers of a block, the halves whose cell index is >= a.
// 512 bytes of shared memory per CTA, filled once (ksw_batch.cu / the emulator harness).
LANE_HD void fill_mask_table(uint32_t *tab, int first, int step)
{
	for (int e = first; e < 16 * 8; e += step) {
		const int a = e >> 3, i = e & 7;
		tab[e] = (2 * i >= a ? 0x0000ffffu : 0u) | (2 * i + 1 >= a ? 0xffff0000u : 0u);
	}
}

// per-team shared memory, in bytes (Hs | Hsnap | QS | Ssp)
LANE_HD int team_smem_bytes(int team, int max_qlen) { return 16 * team * 4 * 2 + ((max_qlen + 2 + 15) & ~15) + 32; }

// The reference's argmax over one diagonal (KSW:319-350) for cells of [st0,en0) that equal M, from an H row
// in shared memory indexed by column (t mod W).  Team-wide; every lane of the team gets t.
template <int TEAM>
LANE_FN int resolve_argmax(const int32_t *H, int tl, int st0, int en0, int M)
{
	constexpr int W = 16 * TEAM;
	const int en1 = st0 + ((en0 - st0) >> 2 << 2);
	int pref = INT32_MIN;
	for (int k = 0; k < 16; ++k) {
		const int col = 16 * tl + k;
		const int t = st0 + ((col - st0) & (W - 1));
		if (t < en0 && H[col] == M) {
			const int d = t - st0;
			const int p = t < en1 ? ((3 - (d & 3)) << 12) + (4095 - (d >> 2)) : -1 - (t - en1);
			if (p > pref) pref = p;
		}
	}
	pref = team_max<TEAM>(pref);
	return pref >= 0 ? st0 + (4095 - (pref & 4095)) * 4 + (3 - (pref >> 12)) : en1 + (-1 - pref);
}

// 32/TEAM alignments, one per team of TEAM lanes; all arguments are team-uniform.
//   Hs, Hsnap : 16*TEAM int32 of shared memory each (this team's)     QS : >= qlen+2 bytes (this team's)
//   Ssp : 32 bytes (this team's)     scr : 8 words (this lane's)     mask_tab : fill_mask_table(), 128 words
//   tb : this team's traceback scratch, >= (anti-diagonals+1) * 16*TEAM bytes (unused with SCORE_ONLY)
template <int TEAM, bool WRAP, bool WC>
LANE_DEV void align_team(const Params &P, bool have_task, int qlen, const uint8_t *__restrict__ query, int tlen,
                         const uint8_t *__restrict__ target, int32_t *__restrict__ res, uint32_t *__restrict__ cigar,
                         int cigar_cap, uint8_t *__restrict__ tb, int32_t *Hs, int32_t *Hsnap, uint8_t *QS, uint8_t *Ssp,
                         uint32_t *scr, const uint32_t *mask_tab)
{
	constexpr int NR = 8;                // registers per row per lane (16 cells)
	constexpr int W = 16 * TEAM;         // cells resident in the team
	const int lane = lane_id(), tl = lane & (TEAM - 1);
	const int left = (lane & ~(TEAM - 1)) | ((tl + TEAM - 1) & (TEAM - 1));
	const int w = P.w < 0 ? (tlen > qlen ? tlen : qlen) : P.w;
	constexpr bool with_cigar = WC;   // WC == !(flag & SCORE_ONLY), chosen by the launcher
	const int q8 = P.q * 8, qe8 = (P.q + P.e) * 8, q28 = P.q2 * 8, qe28 = (P.q2 + P.e2) * 8;

	// packed constants (k32: added with a 32-bit add; dup16: operand of a 16x2 min/max)
	typedef kswfast::Bias<WRAP> BB;
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

	// ---- stage the query (wildcard-encoded, zero on both sides), clear H and the spill buffer
	bool q_wild = false;
	if (have_task) {
		for (int j = tl; j < qlen + 2; j += TEAM) {
			const uint32_t c = (j >= 1 && j <= qlen) ? enc_q(query[j - 1], P.wild) : 0u;
			QS[j] = (uint8_t)c;
			q_wild |= c == 0x20u;
		}
		for (int k = tl; k < W; k += TEAM) Hs[k] = NEG_BIG;
		for (int k = tl; k < 32; k += TEAM) Ssp[k] = 0;
	}
	q_wild = team_max<TEAM>(q_wild ? 1 : 0) != 0;
	wsync();

	// ---- per-lane rows of block blk
	uint32_t U[NR], V[NR], MX[NR], MY[NR], MX2[NR], MY2[NR], S[NR], TB[NR], QB[NR];
	int blk = tl;
	int ssp0 = -1, ssp1 = -1;            // block whose refreshed-ahead scores sit in Ssp[0..15] / Ssp[16..31]
	bool t_wild = false;
	auto load_group = [&](int r_for_q) __attribute__((always_inline)) {
		const int t0 = 16 * blk;
		const bool from_ssp = (blk & 1) ? ssp1 == blk : ssp0 == blk;
		t_wild = false;
#pragma unroll
		for (int i = 0; i < NR; ++i) {
			U[i] = U_DEF; V[i] = V_DEF;
			MX[i] = MY[i] = MX2[i] = MY2[i] = BMd;
			S[i] = S_ZERO;
			uint32_t tlo = 0, thi = 0;
			const int ta = t0 + 2 * i, tbb = ta + 1;
			if (ta < tlen) tlo = enc_t(target[ta], P.wild);
			if (tbb < tlen) thi = enc_t(target[tbb], P.wild);
			t_wild |= tlo == 0x10u || thi == 0x10u;
			TB[i] = pk(tlo, thi);
			int ja = r_for_q - ta, jb = r_for_q - tbb;           // query index of the cell on diagonal r_for_q
			ja = ja < -1 ? -1 : (ja > qlen ? qlen : ja);
			jb = jb < -1 ? -1 : (jb > qlen ? qlen : jb);
			QB[i] = pk(QS[ja + QS_PAD], QS[jb + QS_PAD]);
		}
		if (from_ssp) {
			const int8_t *sp = (const int8_t*)Ssp + 16 * (blk & 1);
#pragma unroll
			for (int i = 0; i < NR; ++i) S[i] = pk((uint32_t)(sp[2 * i] * 8 + 4 + bK), (uint32_t)(sp[2 * i + 1] * 8 + 4 + bK));
		}
	};
	if (have_task) load_group(-1);

	// ---- running ez (H-like quantities are scaled by 8)
	int ez_max8 = 0, ez_max_t = -1, ez_max_q = -1, mqe8 = NEG_INF, mqe_t = -1, mte8 = NEG_INF, mte_q = -1, score8 = NEG_INF;
	int zdropped = 0, last_bs = -1, st0_prev = 0, Hprev = 0, r_max = -1;
	bool max_known = true, done = !have_task;
	const int n_diag = qlen + tlen - 1;
	int32_t *hp = Hs + 16 * tl;

	int st0n = 0, en0n = -1;                 // band of the next diagonal (computed once, used twice)
	if (!done && n_diag > 0) band(0, qlen, tlen, w, st0n, en0n);
	for (int r = 0; ; ++r) {
		const int st0 = st0n, en0 = en0n;
		bool live = !done && r < n_diag;
		if (live && st0 > en0) { zdropped = 1; done = true; live = false; }
		if (!wballot(live)) break;
		const int bs = st0 >> 4, be = en0 >> 4, a = st0 & 15;

		// -- left-neighbour exchange on the rows as they stand after diagonal r-1
		const uint32_t snd1 = prmt(MX[NR - 1], V[NR - 1], 0x7632);   // lo = MX of my last cell, hi = V of it
		uint32_t rcv1 = shfl(snd1, left);
		uint32_t rcv2 = shfl(MX2[NR - 1], left);                      // hi half = MX2 of the neighbour's last cell
		bool active = false;
		int lmax = INT32_MIN, Hen8 = 0, Hst8 = 0, Mx = INT32_MIN;
		uint32_t pk4[4] = {0u, 0u, 0u, 0u};                       // this diagonal's 16 traceback bytes of the lane

		if (live) {
			const int nbk = ((en0 - st0) >> 4) + 1;                   // 16-byte chunks of the score refresh
			// -- window slide: the lane whose block fell left of the band takes the block TEAM further right
			if (blk < bs) {
				blk += TEAM;
				load_group(r);
#pragma unroll
				for (int k = 0; k < 16; k += 4) st4i(hp + k, NEG_BIG, NEG_BIG, NEG_BIG, NEG_BIG);
			} else {                                                  // query window moves one cell per diagonal
				int j = r - 16 * blk;
				j = j < -1 ? -1 : (j > qlen ? qlen : j);
				const uint32_t nq = QS[j + QS_PAD];
#pragma unroll
				for (int i = NR - 1; i > 0; --i) QB[i] = prmt(QB[i - 1], QB[i], 0x5432);
				QB[0] = (QB[0] << 16) | nq;
			}
			// -- scores refreshed ahead of the resident window go to the spill buffer (KSW:158-173)
			{
				const int sb = be + 1, p = sb & 1;
				const bool fresh = (p ? ssp1 : ssp0) != sb;
				const bool ahead = sb >= bs + TEAM;                   // block sb has no lane yet
				const int nsp = ahead ? st0 + 16 * nbk - 16 * sb : 0;
				if (fresh || nsp > 0) {
					for (int c = tl; c < (fresh ? 16 : nsp); c += TEAM) {
						int s = 0;
						if (c < nsp) {
							const int t = 16 * sb + c;
							const uint32_t tt = t < tlen ? enc_t(target[t], P.wild) : 0u;
							int j = r - t;
							j = j < -1 ? -1 : (j > qlen ? qlen : j);
							const uint32_t x = tt ^ QS[j + QS_PAD];
							s = x == 0 ? P.sc_mch : ((x & 0x30u) ? P.sc_N : P.sc_mis);
						}
						Ssp[16 * p + c] = (uint8_t)s;
					}
					if (p) ssp1 = sb; else ssp0 = sb;
				}
			}
			// -- boundary of the first block of the band (KSW:142-152)
			auto first_val = [&]() __attribute__((always_inline)) {   // value of the first row / column on diagonal r (KSW:151,155)
				return r == 0 ? -(P.q + P.e) : r < P.long_thres ? -P.e : r == P.long_thres ? P.long_diff : -P.e2;
			};
			if (blk == bs && !(bs > 0 && bs != last_bs)) {
				const int bv = bs > 0 ? -(P.q + P.e) : first_val();
				rcv1 = pk(bM, (uint32_t)(bv * 8 + bV));
				rcv2 = BMd;
			}
			// -- first-row cell t = r (KSW:153-156): one half of three rows, through the lane's scratch
			if ((be | 0) * 16 + 15 >= r && blk == (r >> 4)) {
				const int k = r & 15, kw = k >> 1, uval = first_val();
				const uint32_t keep = (k & 1) ? 0x0000ffffu : 0xffff0000u, sh = (k & 1) ? 16 : 0;
				st8(scr, MY); scr[kw] = (scr[kw] & keep) | ((uint32_t)bM << sh); ld8(scr, MY);
				st8(scr, MY2); scr[kw] = (scr[kw] & keep) | ((uint32_t)bM << sh); ld8(scr, MY2);
				st8(scr, U); scr[kw] = (scr[kw] & keep) | ((uint32_t)((uval * 8 + bU) & 0xffff) << sh); ld8(scr, U);
			}
			// -- substitution scores on [st0, st0+16*nbk): block bs cells >= a, then full blocks, then block bs+nbk cells < a
			{
				const bool in1 = blk >= bs && blk < bs + nbk, in2 = blk > bs && blk <= bs + nbk;
				if (in1 || in2) {
					// refreshed halves: MA (first block), ~MA (block after the last full one), all (in between)
					const uint32_t XM = (in2 && !in1) ? ~0u : 0u, YM = (in1 && in2) ? ~0u : 0u;
					uint32_t MA[NR];
					ld8(mask_tab + 8 * a, MA);
					if (q_wild || t_wild) {
#pragma unroll
						for (int i = 0; i < NR; ++i) {
							const uint32_t x = TB[i] ^ QB[i];
							const uint32_t sn = SBASE + minu(x, ONE2) * (uint32_t)D1 + minu(x & 0x00300030u, ONE2) * (uint32_t)E2;
							const uint32_t rm = xor_or(MA[i], XM, YM);
							S[i] = (S[i] & ~rm) | (sn & rm);
						}
					} else {
#pragma unroll
						for (int i = 0; i < NR; ++i) {
							const uint32_t sn = SBASE + minu(TB[i] ^ QB[i], ONE2) * (uint32_t)D1;
							const uint32_t rm = xor_or(MA[i], XM, YM);
							S[i] = (S[i] & ~rm) | (sn & rm);
						}
					}
				}
			}

			// -- the 16 cells of this lane, right to left so that [i-1] is still last diagonal's
			active = blk >= bs && blk <= be;
			if (active) {
				uint32_t tb_hi = 0;
#pragma unroll
				for (int i = NR - 1; i >= 0; --i) {
					const uint32_t mxt1 = i > 0 ? prmt(MX[i > 0 ? i - 1 : 0], MX[i], 0x5432) : prmt(rcv1, MX[0], 0x5410);
					const uint32_t vt1 = i > 0 ? prmt(V[i > 0 ? i - 1 : 0], V[i], 0x5432) : prmt(rcv1, V[0], 0x5432);
					const uint32_t mx2t1 = i > 0 ? prmt(MX2[i > 0 ? i - 1 : 0], MX2[i], 0x5432) : prmt(rcv2, MX2[0], 0x5432);
					const uint32_t ut = U[i];
					// a b a2 b2 with their priority tags (adds on the FMA pipe)
					uint32_t A = fadd(P, mxt1, fadd(P, vt1, CA)), Bv = fadd(P, MY[i], fadd(P, ut, CB));
					uint32_t A2 = fadd(P, mx2t1, fadd(P, vt1, CA2)), B2 = fadd(P, MY2[i], fadd(P, ut, CB2));
					if (WRAP) { A &= WM; Bv &= WM; A2 &= WM; B2 &= WM; }
					uint32_t zk = max3u(S[i], A, Bv);
					zk = max3u(zk, A2, B2);
					const uint32_t Z = minu(zk & 0xfff8fff8u, MCH);
					U[i] = fsub(P, fadd(P, Z, CU), vt1);
					V[i] = fsub(P, fadd(P, Z, CV), ut);
					if (WRAP) {
						V[i] &= WM;
						MX[i] = maxu(fsub(P, fadd(P, A, CNA), Z) & WM, BMd);
						MY[i] = maxu(fsub(P, fadd(P, Bv, CNB), Z) & WM, BMd);
						MX2[i] = maxu(fsub(P, fadd(P, A2, CNA2), Z) & WM, BMd);
						MY2[i] = maxu(fsub(P, fadd(P, B2, CNB2), Z) & WM, BMd);
					} else {
						MX[i] = addmaxu(A, fsub(P, CNA, Z), BMd);
						MY[i] = addmaxu(Bv, fsub(P, CNB, Z), BMd);
						MX2[i] = addmaxu(A2, fsub(P, CNA2, Z), BMd);
						MY2[i] = addmaxu(B2, fsub(P, CNB2, Z), BMd);
					}
					if (with_cigar) {                                 // traceback byte: tag + 8,16,32,64 for the four continuations
						uint32_t tw = fadd(P, zk & 0x00070007u, CTB);
						tw = minu(MY2[i], BM8d) * 8u + tw;
						tw = minu(MX2[i], BM8d) * 4u + tw;
						tw = minu(MY[i], BM8d) * 2u + tw;
						tw = fadd(P, minu(MX[i], BM8d), tw);
						if (i & 1) tb_hi = tw; else pk4[i >> 1] = prmt(tw, tb_hi, 0x6420);
					}
				}
			}

			// -- exact H row (KSW:316-351), scaled by 8, in shared memory at column t mod W.  Cells outside
			//    [st0,en0) hold NEG_BIG, so "H[t] += v[t]" and its maximum run over the whole block unmasked.
			const int ken = en0 - 16 * blk;
			const bool own_en = (unsigned)ken < 16u;
			if (r > 0) {
				if (st0 > st0_prev) {                                 // cell st0-1 left the band
					const int kd = st0 - 1 - 16 * blk;
					if ((unsigned)kd < 16u) hp[kd] = NEG_BIG;
				}
				if (own_en) hp[ken] = NEG_BIG;                        // H[en0] is not part of the += loop
				if (active) {
#pragma unroll
					for (int i = 0; i < NR; i += 2) {
						int h[4];
						ld4i(hp + 2 * i, h);
						const uint32_t nbv = (uint32_t)(-bV);
						h[0] = (int)fadd(P, lo16u(V[i]), fadd(P, (uint32_t)h[0], nbv)); h[1] = (int)fadd(P, hi16u(V[i]), fadd(P, (uint32_t)h[1], nbv));
						h[2] = (int)fadd(P, lo16u(V[i + 1]), fadd(P, (uint32_t)h[2], nbv)); h[3] = (int)fadd(P, hi16u(V[i + 1]), fadd(P, (uint32_t)h[3], nbv));
						st4i(hp + 2 * i, h[0], h[1], h[2], h[3]);
						lmax = max3s(lmax, h[0], h[1]);
						lmax = max3s(lmax, h[2], h[3]);
					}
				}
				if (own_en) {                                         // the special last element (KSW:322)
					int d;
					if (en0 > 0) {
						st8(scr, U);
						d = (int)((ken & 1) ? hi16u(scr[ken >> 1]) : lo16u(scr[ken >> 1]));
						d = WRAP ? (d & 0x7ff) - 0x400 : d - bU;        // stored u is not wrapped yet
					} else {                                          // single-column corner: H[0] += v[0]
						st8(scr, V);
						d = (int)((ken & 1) ? hi16u(scr[ken >> 1]) : lo16u(scr[ken >> 1])) - bV;
					}
					hp[ken] = Hprev + d;
				}
			} else if (blk == 0) hp[0] = (int)lo16u(V[0]) - bV - P.qe_as_passed * 8;   // r == 0 (KSW:351)
		}
		// -- one traceback byte per cell, row = diagonal, column = t mod W.  Two neighbouring lanes share a 32-byte sector: when only
		//    one of them is in the band the other one stores too (zeros, for cells no backtrack ever visits), so that DRAM sees
		//    whole sectors and never has to read one back to merge half of it.
		if (with_cigar) {
			const bool mine = live && active;
			const int partner = shfl_xor((int)mine, 1);                  // every lane takes part in the exchange
			if (mine || partner != 0) kswfast::store_cells<16>(tb + (size_t)r * W + 16 * tl, pk4);
		}
		wsync();
		Mx = team_max<TEAM>(lmax);
		bool slow = false;
		if (live) {
			Hen8 = Hs[en0 & (W - 1)];
			if (r - st0 == qlen - 1) Hst8 = Hs[st0 & (W - 1)];
			int maxH8 = Hen8;
			const bool beaten = Mx > Hen8;                            // a cell of [st0,en0) beats H[en0] (which has priority)
			if (beaten) maxH8 = Mx;
			// -- ez bookkeeping (KSW:353-359)
			if (en0 == tlen - 1 && Hen8 > mte8) { mte8 = Hen8; mte_q = r - (en0 | 15); }
			if (r - st0 == qlen - 1 && Hst8 > mqe8) { mqe8 = Hst8; mqe_t = st0; }
			if (maxH8 > ez_max8) {
				ez_max8 = maxH8; r_max = r;
				max_known = !beaten;
				if (!beaten) { ez_max_t = en0; ez_max_q = r - en0; }
				else {                                                // keep this diagonal's H row: the position is resolved lazily
#pragma unroll
					for (int k = 0; k < 16; k += 4) { int h[4]; ld4i(hp + k, h); st4i(Hsnap + 16 * tl + k, h[0], h[1], h[2], h[3]); }
				}
			} else if (P.zdrop >= 0 && ez_max8 - maxH8 > P.zdrop * 8) slow = true;   // only then can the z-drop fire (l*e2 >= 0)
		}
		// -- rare: exact z-drop test needs both argmax positions (K2H:245-261)
		if (wballot(slow)) {
			wsync();
			int st_m = 0, en_m = -1;
			const bool need_old = slow && !max_known;
			if (need_old) band(r_max, qlen, tlen, w, st_m, en_m);
			const int t_old = resolve_argmax<TEAM>(Hsnap, tl, st_m, en_m, need_old ? ez_max8 : INT32_MIN + 1);
			if (need_old) { ez_max_t = t_old; ez_max_q = r_max - t_old; max_known = true; }
			const bool beaten = slow && Mx > Hen8;
			const int t_new = resolve_argmax<TEAM>(Hs, tl, st0, en0, beaten ? Mx : INT32_MIN + 1);
			if (slow) {
				const int max_t = beaten ? t_new : en0, maxH8 = beaten ? Mx : Hen8;
				if (max_t >= ez_max_t && r - max_t >= ez_max_q) {
					const int tlv = max_t - ez_max_t, qlv = (r - max_t) - ez_max_q;
					const int l = tlv > qlv ? tlv - qlv : qlv - tlv;
					if (ez_max8 - maxH8 > (P.zdrop + l * P.e2) * 8) { zdropped = 1; done = true; }
				}
			}
		}
		if (live && !done && r == n_diag - 1 && en0 == tlen - 1) score8 = Hen8;   // not when the z-drop just fired (KSW:357-359)
		if (live) {
			// H[en0'-1] as it stands now is what the next diagonal's last element starts from; if that cell already
			// left the band before this diagonal it has not changed since the value we hold
			band(r + 1, qlen, tlen, w, st0n, en0n);
			if (r + 1 < n_diag && st0n <= en0n) {
				const int c = en0n > 0 ? en0n - 1 : 0;
				if (!(c < st0)) Hprev = Hs[c & (W - 1)];
			}
			last_bs = bs; st0_prev = st0;
		}
		wsync();
	}
	// ---- position of the running maximum, if it was left unresolved
	if (wballot(!max_known)) {
		wsync();
		int st_m = 0, en_m = -1;
		if (!max_known) band(r_max, qlen, tlen, w, st_m, en_m);
		const int t_old = resolve_argmax<TEAM>(Hsnap, tl, st_m, en_m, !max_known ? ez_max8 : INT32_MIN + 1);
		if (!max_known) { ez_max_t = t_old; ez_max_q = r_max - t_old; max_known = true; }
	}

	// ---- traceback (KSW:382-391, K2H:119-151): every team walks the path of its own alignment, TEAM cells of a run per round
	// trip to the traceback bytes -- the alignments of a warp side by side (a warp of sixteen 40 x 40 extensions makes about as
	// many dependent round trips for all of them as it used to make for one after the other)
	const int ez_max = ez_max8 >> 3;
	const int mqe = mqe8 == NEG_INF ? NEG_INF : mqe8 >> 3;
	int n_cigar = 0, reach_end = 0, overflow = 0;
	wsync();
	if (with_cigar) {
		const int base = lane & ~(TEAM - 1);
		const uint32_t tmask = TEAM == 32 ? 0xffffffffu : ((1u << TEAM) - 1u);
		int i = -1, j = -1;
		if (have_task) {
			if (!zdropped && !(P.flag & kswfast::F_EXTZ_ONLY)) { i = tlen - 1; j = qlen - 1; }
			else if (!zdropped && (P.flag & kswfast::F_EXTZ_ONLY) && mqe + P.end_bonus > ez_max) { reach_end = 1; i = mqe_t; j = qlen - 1; }
			else if (ez_max_t >= 0 && ez_max_q >= 0) { i = ez_max_t; j = ez_max_q; }
		}
		int state = 0;
		uint32_t cur = 0;                                             // last CIGAR element, not yet stored
		auto push = [&](uint32_t op, int len) __attribute__((always_inline)) {
			if (n_cigar == 0 || op != (cur & 0xfu)) {
				if (n_cigar > 0) {
					if (n_cigar - 1 < cigar_cap) { if (tl == 0) cigar[n_cigar - 1] = cur; }
					else overflow = 1;
				}
				++n_cigar;
				cur = (uint32_t)len << 4 | op;
			} else cur += (uint32_t)len << 4;
		};
		while (wballot(i >= 0 && j >= 0)) {
			const bool walking = i >= 0 && j >= 0;                    // (team-uniform: every lane of a team holds the same i, j, state)
			const int di = (state == 0 || state == 1 || state == 3) ? 1 : 0;
			const int dj = (state == 0 || state == 2 || state == 4) ? 1 : 0;
			const int li = i - tl * di, lj = j - tl * dj;
			const bool valid = walking && li >= 0 && lj >= 0;
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
					// a plain load: the bytes were stored by lanes of this same warp before the __syncwarp above
					const uint32_t bb = tb[(size_t)rr * W + (li & (W - 1))];
					tmp = (bb & 0x78u) | (4u - (bb & 7u));
				}
				clean = forced < 0 && (state == 0 ? (tmp & 7u) == 0 : ((tmp >> (state + 2)) & 1u) != 0);
			}
			const uint32_t stop = (wballot(!clean) >> base) & tmask;
			const int n = stop ? ffs32(stop) - 1 : TEAM;
			const int from = base + (n < TEAM ? n : 0);               // the first cell that breaks the run, if it exists
			const int v = shfl((int)valid, from);
			const uint32_t t2 = shfl(tmp, from);
			const int f2 = shfl(forced, from);
			if (walking) {
				if (n > 0) {
					push(state == 0 ? 0u : (di ? 2u : 1u), n);
					i -= n * di; j -= n * dj;
				}
				if (n < TEAM && v) {
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
		if (have_task) {
			if (i >= 0) push(2, i + 1);
			if (j >= 0) push(1, j + 1);
			if (n_cigar > 0) {
				if (n_cigar - 1 < cigar_cap) { if (tl == 0) cigar[n_cigar - 1] = cur; }
				else overflow = 1;
			}
		}
		if (!(P.flag & kswfast::F_REV_CIGAR)) {
			wsync();
			if (have_task && !overflow)
				for (int k = tl; k < n_cigar >> 1; k += TEAM) {
					const uint32_t x = cigar[k], y = cigar[n_cigar - 1 - k];
					cigar[k] = y; cigar[n_cigar - 1 - k] = x;
				}
		}
	}
	if (have_task) for (int k = n_cigar + tl; k < cigar_cap; k += TEAM) cigar[k] = 0;   // deterministic tail of the CIGAR row
	const int my_n_cigar = n_cigar, my_reach_end = reach_end, my_overflow = overflow;
	if (have_task && tl == 0) {
		res[0] = ez_max; res[1] = zdropped; res[2] = ez_max_q; res[3] = ez_max_t;
		res[4] = mqe; res[5] = mqe_t; res[6] = mte8 == NEG_INF ? NEG_INF : mte8 >> 3; res[7] = mte_q;
		res[8] = score8 == NEG_INF ? NEG_INF : score8 >> 3; res[9] = my_n_cigar; res[10] = my_reach_end; res[11] = my_overflow;
	}
	wsync();
}

} // namespace kswteam
