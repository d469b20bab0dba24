// ksw_common.cuh -- definitions shared by the ksw device kernels and their host-side planner:
// flag values (src/kswlib/ksw2.h:9-15), the normalised parameter block, the band of an
// anti-diagonal (src/kswlib/ksw2_extd2_sse.c:131-134) and the packed 16x2 number format.
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
	uint32_t one, mone;   // 1 and -1 the compiler cannot see through: a*one+b keeps an add on the FMA pipe (IMAD), see fadd()
};

// Integer adds that must not compete with LOP3/VIMNMX/PRMT for the ALU pipe: IMAD runs on the FMA pipe.
LANE_FN uint32_t fadd(const Params &P, uint32_t a, uint32_t b) { return a * P.one + b; }        // a + b
LANE_FN uint32_t fsub(const Params &P, uint32_t a, uint32_t z) { return z * P.mone + a; }       // a - z
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

} // namespace kswfast
