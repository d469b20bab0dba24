// ksw_generic.cuh -- catch-all device kernel: one thread per alignment, rows in global scratch.
//
// Covers what the team kernel (ksw_team.cuh) declines: KSW_EZ_RIGHT, KSW_EZ_GENERIC_SC,
// KSW_EZ_APPROX_MAX/DROP, bands wider than 32 blocks of 16 cells, gap costs above 127.  It executes the
// reference's 16-lane int8 machine literally (src/kswlib/ksw2_extd2_sse.c:26-396,
// src/kswlib/ksw2.h:106-151,238-261): seven int8 rows indexed by target position, band rounded
// to 16-cell blocks, byte arithmetic with wrap-around, int32 H row with the SSE scan's tie
// order, byte-per-cell traceback.  Throughput is not a goal here (no caller on the panSVR path
// uses these modes); exactness is.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/pansvr_b200.h"
#include "ksw_common.cuh"

namespace kswgeneric {

struct GArgs {
	kswfast::Params P;
	int m; int8_t mat[256];
	int n; const int *order;
	const uint8_t *qseq; const int64_t *qoff; const int32_t *qlen;
	const uint8_t *tseq; const int64_t *toff; const int32_t *tlen;
	int32_t *res; uint32_t *cigar; int cigar_cap;
	uint8_t *scratch; size_t per_slot; int n_slots;
};

__device__ __forceinline__ int8_t w8(int v) { return (int8_t)(v & 0xff); }

struct Ez { int max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end, overflow; };

__device__ inline bool g_zdrop(Ez &ez, int H, int r, int t, int zdrop, int e)
{
	if (H > ez.max) { ez.max = H; ez.max_t = t; ez.max_q = r - t; }
	else if (t >= ez.max_t && r - t >= ez.max_q) {
		int tl = t - ez.max_t, ql = (r - t) - ez.max_q, l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && ez.max - H > zdrop + l * e) { ez.zdropped = 1; return true; }
	}
	return false;
}

// CIGAR writer with ksw_push_cigar's merge rule (ksw2.h:106-116); the open element stays in a
// register, elements beyond `cap` are counted but not stored (status bit 0).
struct CigarOut {
	uint32_t *cigar; int cap; int n; uint32_t cur; int overflow;
	__device__ void flush() { if (n > 0) { if (n - 1 < cap) cigar[n - 1] = cur; else overflow = 1; } }
	__device__ void push(uint32_t op, int len)
	{
		if (n == 0 || op != (cur & 0xfu)) { flush(); ++n; cur = (uint32_t)len << 4 | op; }
		else cur += (uint32_t)len << 4;
	}
};

__device__ inline void g_backtrack(Ez &ez, uint32_t *cigar, int cap, bool rev, const uint8_t *dir, size_t stride, int qlen,
                                   int tlen, int w, int i, int j)
{
	int state = 0;
	CigarOut co = {cigar, cap, 0, 0, 0};
	while (i >= 0 && j >= 0) {
		int r = i + j, lo0, hi0, forced = -1;
		kswfast::band(r, qlen, tlen, w, lo0, hi0);
		const int lo = lo0 & ~15, hi = hi0 | 15;
		if (i < lo) forced = 2;
		if (i > hi) forced = 1;
		const uint32_t b = forced < 0 ? dir[(size_t)r * stride + (size_t)(i - lo)] : 0;
		if (state == 0) state = b & 7;
		else if (!((b >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = b & 7;
		if (forced >= 0) state = forced;
		if (state == 0) { co.push(0, 1); --i; --j; }
		else if (state == 1 || state == 3) { co.push(2, 1); --i; }
		else { co.push(1, 1); --j; }
	}
	if (i >= 0) co.push(2, i + 1);
	if (j >= 0) co.push(1, j + 1);
	co.flush();
	ez.n_cigar = co.n; ez.overflow = co.overflow;
	if (!rev && !co.overflow)
		for (int k = 0; k < co.n >> 1; ++k) {
			uint32_t t = cigar[k]; cigar[k] = cigar[co.n - 1 - k]; cigar[co.n - 1 - k] = t;
		}
}

__device__ inline void g_align(const GArgs &a, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int32_t *res,
                               uint32_t *cigar, uint8_t *mem)
{
	const kswfast::Params &P = a.P;
	const int flag = P.flag, q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, m = a.m;
	const bool with_cigar = !(flag & kswfast::F_SCORE_ONLY), approx = flag & kswfast::F_APPROX_MAX, right = flag & kswfast::F_RIGHT;
	const int w = P.w < 0 ? (tlen > qlen ? tlen : qlen) : P.w;
	const int T16 = (tlen + 15) / 16 * 16, Q16 = (qlen + 15) / 16 * 16;
	int nb = qlen < tlen ? qlen : tlen;
	nb = ((nb < w + 1 ? nb : w + 1) + 15) / 16 + 1;
	const size_t stride = (size_t)nb * 16;
	int8_t *U = (int8_t*)mem, *V = U + T16, *X = V + T16, *Y = X + T16, *X2 = Y + T16, *Y2 = X2 + T16, *S = Y2 + T16;
	uint8_t *SF = (uint8_t*)(S + T16), *QR = SF + T16;
	int32_t *H = (int32_t*)(QR + Q16 + 16);
	uint8_t *dir = (uint8_t*)(H + T16);
	Ez ez = {0, 0, -1, -1, kswfast::NEG_INF, -1, kswfast::NEG_INF, -1, kswfast::NEG_INF, 0, 0, 0};
	for (int t = 0; t < T16; ++t) { U[t] = V[t] = X[t] = Y[t] = w8(-q - e); X2[t] = Y2[t] = w8(-q2 - e2); S[t] = 0; SF[t] = 0; H[t] = kswfast::NEG_INF; }
	for (int t = 0; t < Q16 + 16; ++t) QR[t] = 0;
	for (int t = 0; t < qlen; ++t) QR[t] = query[qlen - 1 - t];
	for (int t = 0; t < tlen; ++t) SF[t] = target[t];
	int prev_lo = -1, prev_hi = -1, H0 = 0, last_H0_t = 0;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int lo0, hi0;
		kswfast::band(r, qlen, tlen, w, lo0, hi0);
		if (lo0 > hi0) { ez.zdropped = 1; break; }
		const int lo = lo0 & ~15, hi = hi0 | 15;
		const uint8_t *qrr = QR + (qlen - 1 - r);
		const int8_t first = r == 0 ? w8(-q - e) : r < P.long_thres ? w8(-e) : r == P.long_thres ? w8(P.long_diff) : w8(-e2);
		int8_t cx, cx2, cv;
		if (lo > 0) {
			if (lo - 1 >= prev_lo && lo - 1 <= prev_hi) { cx = X[lo - 1]; cx2 = X2[lo - 1]; cv = V[lo - 1]; }
			else { cx = w8(-q - e); cx2 = w8(-q2 - e2); cv = w8(-q - e); }
		} else { cx = w8(-q - e); cx2 = w8(-q2 - e2); cv = first; }
		if (hi >= r) { Y[r] = w8(-q - e); Y2[r] = w8(-q2 - e2); U[r] = first; }
		if (!(flag & kswfast::F_GENERIC_SC)) {
			for (int t = lo0; t <= hi0; t += 16)
				for (int k = 0; k < 16; ++k) {   // may read SF past T16 (into QR) and write S past T16 (into SF), as the reference does
					const uint8_t x = SF[t + k], y = qrr[t + k];
					int8_t sc = x == y ? (int8_t)P.sc_mch : (int8_t)P.sc_mis;
					if (x == (uint8_t)P.wild || y == (uint8_t)P.wild) sc = (int8_t)P.sc_N;
					S[t + k] = sc;
				}
		} else for (int t = lo0; t <= hi0; ++t) S[t] = a.mat[SF[t] * m + qrr[t]];
		for (int t = lo; t <= hi; ++t) {
			int8_t z = S[t];
			const int8_t xt1 = cx, x2t1 = cx2, vt1 = cv, ut = U[t];
			cx = X[t]; cx2 = X2[t]; cv = V[t];
			int8_t aa = w8(xt1 + vt1), bb = w8(Y[t] + ut), a2 = w8(x2t1 + vt1), b2 = w8(Y2[t] + ut), tmp;
			uint8_t d = 0;
			if (!right) {
				if (aa > z) { d = 1; z = aa; } if (bb > z) { d = 2; z = bb; }
				if (a2 > z) { d = 3; z = a2; } if (b2 > z) { d = 4; z = b2; }
			} else {
				if (!(z > aa)) { d = 1; z = aa; } if (!(z > bb)) { d = 2; z = bb; }
				if (!(z > a2)) { d = 3; z = a2; } if (!(z > b2)) { d = 4; z = b2; }
			}
			if (z > (int8_t)P.sc_mch) z = (int8_t)P.sc_mch;
			U[t] = w8(z - vt1); V[t] = w8(z - ut);
			tmp = w8(z - q); aa = w8(aa - tmp); bb = w8(bb - tmp);
			tmp = w8(z - q2); a2 = w8(a2 - tmp); b2 = w8(b2 - tmp);
			const bool ca = right ? aa >= 0 : aa > 0, cb = right ? bb >= 0 : bb > 0, ca2 = right ? a2 >= 0 : a2 > 0, cb2 = right ? b2 >= 0 : b2 > 0;
			X[t] = w8((ca ? aa : 0) - (q + e)); Y[t] = w8((cb ? bb : 0) - (q + e));
			X2[t] = w8((ca2 ? a2 : 0) - (q2 + e2)); Y2[t] = w8((cb2 ? b2 : 0) - (q2 + e2));
			d |= (ca ? 0x08 : 0) | (cb ? 0x10 : 0) | (ca2 ? 0x20 : 0) | (cb2 ? 0x40 : 0);
			if (with_cigar) dir[(size_t)r * stride + (size_t)(t - lo)] = d;
		}
		if (!approx) {
			int best, best_t;
			if (r > 0) {
				int lb[4], lt[4];
				const int vend = lo0 + (hi0 - lo0) / 4 * 4;
				best = H[hi0] = hi0 > 0 ? H[hi0 - 1] + U[hi0] : H[hi0] + V[hi0];
				best_t = hi0;
				for (int k = 0; k < 4; ++k) { lb[k] = best; lt[k] = best_t; }
				for (int t = lo0; t < vend; t += 4)
					for (int k = 0; k < 4; ++k) { H[t + k] += V[t + k]; if (H[t + k] > lb[k]) { lb[k] = H[t + k]; lt[k] = t; } }
				for (int k = 0; k < 4; ++k) if (best < lb[k]) { best = lb[k]; best_t = lt[k] + k; }
				for (int t = vend; t < hi0; ++t) { H[t] += V[t]; if (H[t] > best) { best = H[t]; best_t = t; } }
			} else { H[0] = V[0] - P.qe_as_passed; best = H[0]; best_t = 0; }
			if (hi0 == tlen - 1 && H[hi0] > ez.mte) { ez.mte = H[hi0]; ez.mte_q = r - hi; }
			if (r - lo0 == qlen - 1 && H[lo0] > ez.mqe) { ez.mqe = H[lo0]; ez.mqe_t = lo0; }
			if (g_zdrop(ez, best, r, best_t, P.zdrop, e2)) break;
			if (r == qlen + tlen - 2 && hi0 == tlen - 1) ez.score = H[tlen - 1];
		} else {
			if (r > 0) {
				if (last_H0_t >= lo0 && last_H0_t <= hi0 && last_H0_t + 1 >= lo0 && last_H0_t + 1 <= hi0) {
					const int d0 = V[last_H0_t], d1 = U[last_H0_t + 1];
					if (d0 > d1) H0 += d0; else { H0 += d1; ++last_H0_t; }
				} else if (last_H0_t >= lo0 && last_H0_t <= hi0) H0 += V[last_H0_t];
				else { ++last_H0_t; H0 += U[last_H0_t]; }
			} else { H0 = V[0] - P.qe_as_passed; last_H0_t = 0; }
			if ((flag & kswfast::F_APPROX_DROP) && g_zdrop(ez, H0, r, last_H0_t, P.zdrop, e2)) break;
			if (r == qlen + tlen - 2 && hi0 == tlen - 1) ez.score = H0;
		}
		prev_lo = lo; prev_hi = hi;
	}
	if (with_cigar) {
		const bool rev = flag & kswfast::F_REV_CIGAR;
		if (!ez.zdropped && !(flag & kswfast::F_EXTZ_ONLY)) g_backtrack(ez, cigar, a.cigar_cap, rev, dir, stride, qlen, tlen, w, tlen - 1, qlen - 1);
		else if (!ez.zdropped && (flag & kswfast::F_EXTZ_ONLY) && ez.mqe + P.end_bonus > ez.max) {
			ez.reach_end = 1;
			g_backtrack(ez, cigar, a.cigar_cap, rev, dir, stride, qlen, tlen, w, ez.mqe_t, qlen - 1);
		} else if (ez.max_t >= 0 && ez.max_q >= 0) g_backtrack(ez, cigar, a.cigar_cap, rev, dir, stride, qlen, tlen, w, ez.max_t, ez.max_q);
	}
	for (int k = ez.n_cigar; k < a.cigar_cap; ++k) cigar[k] = 0;
	res[0] = ez.max; res[1] = ez.zdropped; res[2] = ez.max_q; res[3] = ez.max_t; res[4] = ez.mqe; res[5] = ez.mqe_t;
	res[6] = ez.mte; res[7] = ez.mte_q; res[8] = ez.score; res[9] = ez.n_cigar; res[10] = ez.reach_end; res[11] = ez.overflow;
}

__global__ void ksw_generic_kernel(const __grid_constant__ GArgs a)
{
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= a.n_slots) return;
	uint8_t *mem = a.scratch + (size_t)slot * a.per_slot;
	for (int i = slot; i < a.n; i += a.n_slots) {
		const int t = a.order[i];
		g_align(a, a.qlen[t], a.qseq + a.qoff[t], a.tlen[t], a.tseq + a.toff[t], a.res + (size_t)t * kswfast::RES_WORDS,
		        a.cigar + (size_t)t * a.cigar_cap, mem);
	}
}

static inline int launch(cudaStream_t stream, int sm_count, const pansvr_ksw_params_t *pr, const kswfast::Params &P, int cnt,
                         const int *d_order, const uint8_t *d_qseq, const int64_t *d_qoff, const int32_t *d_qlen,
                         const uint8_t *d_tseq, const int64_t *d_toff, const int32_t *d_tlen, int32_t *d_res, uint32_t *d_cigar,
                         int cigar_cap, int max_qlen, int max_tlen, void **scratch, size_t *scratch_cap, int64_t *launches,
                         std::string &err)
{
	if (pr->m > 16) { err = "generic ksw kernel: alphabets above 16 symbols are not supported"; return PANSVR_E_UNSUPPORTED; }
	GArgs a;
	a.P = P; a.m = pr->m;
	memset(a.mat, 0, sizeof(a.mat));
	memcpy(a.mat, pr->mat, (size_t)pr->m * pr->m);
	a.n = cnt; a.order = d_order;
	a.qseq = d_qseq; a.qoff = d_qoff; a.qlen = d_qlen; a.tseq = d_tseq; a.toff = d_toff; a.tlen = d_tlen;
	a.res = d_res; a.cigar = d_cigar; a.cigar_cap = cigar_cap;
	const int w = P.w < 0 ? std::max(max_qlen, max_tlen) : P.w;
	const size_t T16 = ((size_t)max_tlen + 15) / 16 * 16, Q16 = ((size_t)max_qlen + 15) / 16 * 16;
	const size_t nb = ((size_t)std::min(std::min(max_qlen, max_tlen), w + 1) + 15) / 16 + 1;
	size_t per = T16 * 9 + Q16 + 16 + T16 * 4 + 64;
	if (!(P.flag & kswfast::F_SCORE_ONLY)) per += ((size_t)max_qlen + max_tlen) * nb * 16 + 16;
	per = (per + 255) & ~(size_t)255;
	size_t slots = std::min<size_t>((size_t)cnt, (size_t)sm_count * 512);
	const size_t cap_bytes = (size_t)8 << 30;
	if (per * slots > cap_bytes) slots = std::max<size_t>(1, cap_bytes / per);
	if (per * slots > *scratch_cap) {
		if (*scratch) cudaFree(*scratch);
		*scratch = nullptr; *scratch_cap = 0;
		cudaError_t e = cudaMalloc(scratch, per * slots);
		if (e != cudaSuccess) { err = std::string("generic ksw scratch: ") + cudaGetErrorString(e); return PANSVR_E_CUDA; }
		*scratch_cap = per * slots;
	}
	a.scratch = (uint8_t*)*scratch; a.per_slot = per; a.n_slots = (int)slots;
	ksw_generic_kernel<<<(unsigned)((slots + 63) / 64), 64, 0, stream>>>(a);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { err = std::string("generic ksw launch: ") + cudaGetErrorString(e); return PANSVR_E_CUDA; }
	++*launches;
	return 0;
}

} // namespace kswgeneric
