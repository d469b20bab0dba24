// lane_rt.cuh -- the handful of warp-level and packed-16x2 primitives the ksw kernels are
// written against.
//
// On the device (nvcc, sm_100a) every primitive is the single SASS instruction named in its
// comment.  When the same kernel source is compiled by g++ with -DPANSVR_HOST_EMUL (only done by
// tests/emul, to step the warp program lane by lane on a machine without a GPU) the primitives
// fall back to scalar code and the warp collectives to a 32-fibre lock-step scheduler
// (tests/emul/warp_emul.hpp).  The product library is never built that way.
#pragma once
#include <stdint.h>

#ifdef PANSVR_HOST_EMUL
#include "warp_emul.hpp"   // tests/emul: WarpEmul (fibres), provides the collectives below
#define LANE_FN inline
#define LANE_HD inline
#define LANE_DEV
#else
#define LANE_FN __device__ __forceinline__
#define LANE_HD __host__ __device__ __forceinline__
#define LANE_DEV __device__
#endif

namespace lanert {

// ---------------------------------------------------------------- packed unsigned 16x2
// Every DP quantity is kept as  true_value*8 + BIAS  in an unsigned 16-bit half, two cells per
// register.  Because both halves are non-negative and never overflow 16 bits, ordinary 32-bit
// IADD3/IMAD on the packed word are exact per half (the word is just hi*65536+lo); only
// min/max need the 16x2 forms.

LANE_FN uint32_t pk(uint32_t lo, uint32_t hi) { return (lo & 0xffffu) | (hi << 16); }
LANE_FN uint32_t dup16(int v) { uint32_t x = (uint32_t)v & 0xffffu; return x | (x << 16); }

#ifndef PANSVR_HOST_EMUL
LANE_FN uint32_t max3u(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }   // VIMNMX3.U16x2
LANE_FN uint32_t addmaxu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_u16x2(a, b, c); } // VIADDMNMX.U16x2
LANE_FN uint32_t minu(uint32_t a, uint32_t b) { return __vminu2(a, b); }                          // VIMNMX.U16x2
LANE_FN uint32_t maxu(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }                          // VIMNMX.U16x2
LANE_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }    // PRMT
LANE_FN int max3s(int a, int b, int c) { return __vimax3_s32(a, b, c); }                          // VIMNMX3.S32
LANE_FN uint32_t xor_or(uint32_t a, uint32_t b, uint32_t c)                                       // (a ^ b) | c as one LOP3
{
	uint32_t r;
	asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	return r;
}
#else
LANE_FN uint32_t xor_or(uint32_t a, uint32_t b, uint32_t c) { return (a ^ b) | c; }
LANE_FN uint32_t max3u(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t lo = a & 0xffff, hi = a >> 16;
	if ((b & 0xffff) > lo) lo = b & 0xffff;
	if ((c & 0xffff) > lo) lo = c & 0xffff;
	if ((b >> 16) > hi) hi = b >> 16;
	if ((c >> 16) > hi) hi = c >> 16;
	return lo | hi << 16;
}
LANE_FN uint32_t addmaxu(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t lo = (a + b) & 0xffff, hi = ((a >> 16) + (b >> 16)) & 0xffff;
	if ((c & 0xffff) > lo) lo = c & 0xffff;
	if ((c >> 16) > hi) hi = c >> 16;
	return lo | hi << 16;
}
LANE_FN uint32_t minu(uint32_t a, uint32_t b)
{
	uint32_t lo = (a & 0xffff) < (b & 0xffff) ? (a & 0xffff) : (b & 0xffff);
	uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
	return lo | hi << 16;
}
LANE_FN uint32_t maxu(uint32_t a, uint32_t b)
{
	uint32_t lo = (a & 0xffff) > (b & 0xffff) ? (a & 0xffff) : (b & 0xffff);
	uint32_t hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
	return lo | hi << 16;
}
LANE_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
	uint64_t src = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		uint32_t n = (sel >> (4 * i)) & 0xf;
		uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
		if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
LANE_FN int max3s(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }
#endif

// (hi half of a, lo half of b) -> the register that sits one cell to the left of b
LANE_FN uint32_t shl_cell(uint32_t left, uint32_t cur) { return prmt(left, cur, 0x5432); }
LANE_FN int lo16s(uint32_t v) { return (int)(int16_t)(v & 0xffff); }
LANE_FN int hi16s(uint32_t v) { return (int)v >> 16; }
LANE_FN uint32_t lo16u(uint32_t v) { return v & 0xffffu; }
LANE_FN uint32_t hi16u(uint32_t v) { return v >> 16; }

// ---------------------------------------------------------------- warp collectives
#ifndef PANSVR_HOST_EMUL
LANE_FN int lane_id() { return (int)(threadIdx.x & 31u); }
LANE_FN uint32_t shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }           // SHFL.IDX
LANE_FN int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
LANE_FN int shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }                  // SHFL.BFLY
LANE_FN int wmax(int v) { return __reduce_max_sync(0xffffffffu, v); }                             // CREDUX.MAX.S32
LANE_FN int wmin(int v) { return __reduce_min_sync(0xffffffffu, v); }
LANE_FN uint32_t wballot(bool p) { return __ballot_sync(0xffffffffu, p); }
LANE_FN void wsync() { __syncwarp(); }
LANE_FN int ffs32(uint32_t v) { return __ffs((int)v); }
LANE_FN int popc32(uint32_t v) { return __popc(v); }
#else
LANE_FN int lane_id() { return WarpEmul::lane(); }
LANE_FN uint32_t shfl(uint32_t v, int src) { return WarpEmul::shfl(v, src); }
LANE_FN int shfl(int v, int src) { return (int)WarpEmul::shfl((uint32_t)v, src); }
LANE_FN int shfl_xor(int v, int m) { return (int)WarpEmul::shfl((uint32_t)v, WarpEmul::lane() ^ m); }
LANE_FN int wmax(int v) { return WarpEmul::wmax(v); }
LANE_FN int wmin(int v) { return -WarpEmul::wmax(-v); }
LANE_FN uint32_t wballot(bool p) { return WarpEmul::ballot(p); }
LANE_FN void wsync() { WarpEmul::sync(); }
LANE_FN int ffs32(uint32_t v) { return __builtin_ffs((int)v); }
LANE_FN int popc32(uint32_t v) { return __builtin_popcount(v); }
#endif

// 8 registers <-> 32 bytes of 16-byte aligned shared memory (two 128-bit accesses on the device)
LANE_FN void st8(uint32_t *p, const uint32_t *r)
{
#ifndef PANSVR_HOST_EMUL
	((uint4*)p)[0] = make_uint4(r[0], r[1], r[2], r[3]);
	((uint4*)p)[1] = make_uint4(r[4], r[5], r[6], r[7]);
#else
	for (int i = 0; i < 8; ++i) p[i] = r[i];
#endif
}
LANE_FN void ld8(const uint32_t *p, uint32_t *r)
{
#ifndef PANSVR_HOST_EMUL
	const uint4 a = ((const uint4*)p)[0], b = ((const uint4*)p)[1];
	r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
#else
	for (int i = 0; i < 8; ++i) r[i] = p[i];
#endif
}
// 4 ints <-> 16 bytes of 16-byte aligned shared memory
LANE_FN void ld4i(const int32_t *p, int *r)
{
#ifndef PANSVR_HOST_EMUL
	const int4 a = *(const int4*)p; r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
#else
	for (int i = 0; i < 4; ++i) r[i] = p[i];
#endif
}
LANE_FN void st4i(int32_t *p, int a, int b, int c, int d)
{
#ifndef PANSVR_HOST_EMUL
	*(int4*)p = make_int4(a, b, c, d);
#else
	p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

} // namespace lanert
