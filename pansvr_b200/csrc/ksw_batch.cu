// ksw_batch.cu -- device kernels, batch planner and C ABI of the ksw extension stage.
//
// Replaces the reference's per-call CPU path
//   KSW_ALN_handler::align_non_splice -> ksw_extd2_sse
//   (src/PanSVgenerateVCF/read_realignment.cpp:872-891, src/kswlib/ksw2_extd2_sse.c:26-396)
// with a batched launch: the host plans the batch (kernel variant per task, longest-first
// order, scratch), persistent CTAs pull alignments from a global counter, one warp per alignment
// (ksw_fast.cuh).  No CPU fallback: every entry point fails if CUDA does.
#include <cuda_runtime.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pansvr_b200.h"
#include "ksw_common.cuh"
#include "ksw_team.cuh"
#include "ksw_generic.cuh"
#include "ksw_host.hpp"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(PANSVR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

constexpr int WARPS_PER_CTA = 4;
constexpr int THREADS = WARPS_PER_CTA * 32;

struct KArgs {
	kswfast::Params P;
	int n;                       // tasks of this launch
	const int *order;            // task ids, longest first
	int *counter;                // work queue head
	const uint8_t *qseq; const int64_t *qoff; const int32_t *qlen;
	const uint8_t *tseq; const int64_t *toff; const int32_t *tlen;
	int32_t *res; uint32_t *cigar; int cigar_cap;
	uint8_t *tb; size_t tb_per_team;
	int smem_per_team;
};

// Persistent CTAs; each warp pulls the next 32/TEAM alignments from the queue until it is empty.
template <int TEAM, bool WRAP, bool WC>
__global__ void __launch_bounds__(THREADS, 3) ksw_team_kernel(const __grid_constant__ KArgs a)
{
	extern __shared__ __align__(16) uint8_t smem[];
	constexpr int NT = 32 / TEAM, W = 16 * TEAM;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, team = lane / TEAM;
	uint8_t *wbase = smem + (size_t)warp * ((size_t)NT * a.smem_per_team + 32 * 32);
	uint8_t *base = wbase + (size_t)team * a.smem_per_team;
	int32_t *Hs = (int32_t*)base, *Hsnap = Hs + W;
	uint8_t *QS = (uint8_t*)(Hsnap + W);
	uint8_t *Ssp = base + a.smem_per_team - 32;
	uint32_t *scr = (uint32_t*)(wbase + (size_t)NT * a.smem_per_team) + lane * 8;
	uint32_t *mask_tab = (uint32_t*)(smem + (size_t)WARPS_PER_CTA * ((size_t)NT * a.smem_per_team + 32 * 32));
	kswteam::fill_mask_table(mask_tab, threadIdx.x, THREADS);
	__syncthreads();
	uint8_t *tb = a.tb + ((size_t)(blockIdx.x * WARPS_PER_CTA + warp) * NT + team) * a.tb_per_team;
	for (;;) {
		int idx = 0;
		if (lane == 0) idx = atomicAdd(a.counter, NT);
		idx = __shfl_sync(0xffffffffu, idx, 0);
		if (idx >= a.n) break;
		const bool have = idx + team < a.n;
		const int t = a.order ? a.order[have ? idx + team : idx] : (have ? idx + team : idx);   // order == NULL: identity
		kswteam::align_team<TEAM, WRAP, WC>(a.P, have, have ? a.qlen[t] : 0, a.qseq + a.qoff[t], have ? a.tlen[t] : 0, a.tseq + a.toff[t],
		                                a.res + (size_t)t * kswfast::RES_WORDS, a.cigar + (size_t)t * a.cigar_cap, a.cigar_cap,
		                                tb, Hs, Hsnap, QS, Ssp, scr, mask_tab);
	}
}

// Tasks the reference answers without running the DP (KSW:68, KSW:93): ksw_reset_extz only.
__global__ void ksw_reset_kernel(int n, const int *order, int32_t *res, uint32_t *cigar, int cigar_cap)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int32_t *o = res + (size_t)order[i] * kswfast::RES_WORDS;
	for (int k = 0; k < cigar_cap; ++k) cigar[(size_t)order[i] * cigar_cap + k] = 0;
	o[0] = 0; o[1] = 0; o[2] = -1; o[3] = -1; o[4] = kswfast::NEG_INF; o[5] = -1; o[6] = kswfast::NEG_INF; o[7] = -1;
	o[8] = kswfast::NEG_INF; o[9] = 0; o[10] = 0; o[11] = 0;
}

// Integer-ALU yardstick for the roofline: 32 independent dependency chains
// per thread of the 32-bit integer instructions the DP is made of (IADD3 / LOP3 / VIMNMX), so the
// ALU+FMA integer issue rate is the only limit.  One "op" = one 32-bit lane operation.
__global__ void __launch_bounds__(256) int_alu_peak_kernel(uint32_t *out, int iters, uint32_t seed)
{
	uint32_t a[16];
#pragma unroll
	for (int k = 0; k < 16; ++k) a[k] = seed * (threadIdx.x + 1) + k * 0x9e3779b9u;
	const uint32_t c1 = seed | 1u, c2 = seed ^ 0x5bd1e995u;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			a[k] = a[k] + c1 + (uint32_t)it;                      // IADD3
			a[k] = (a[k] ^ c2) & ~c1;                             // LOP3
			a[k] = (uint32_t)max((int)a[k], (int)c2);             // VIMNMX.S32
			a[k] = a[k] - c2 + c1;                                // IADD3
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < 16; ++k) r ^= a[k];
	if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;   // keeps the chains alive
}

// The same per pipe (the mixed chain above is a compiler-chosen blend of both integer pipes).
// MODE 1: LOP3 + VIMNMX only (ALU pipe).  MODE 2: IMAD only (FMA pipe; the multiplier is opaque so that the adds stay IMAD).
// MODE 3: eight LOP3/VIMNMX chains and eight IMAD chains side by side (both pipes, the issue limit).
template <int MODE>
__global__ void __launch_bounds__(256) int_pipe_peak_kernel(uint32_t *out, int iters, uint32_t seed, uint32_t one)
{
	uint32_t a[16];
#pragma unroll
	for (int k = 0; k < 16; ++k) a[k] = seed * (threadIdx.x + 1) + k * 0x9e3779b9u;
	const uint32_t c1 = seed | 1u, c2 = seed ^ 0x5bd1e995u;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			const bool alu = MODE == 1 || (MODE == 3 && (k & 1));
			if (alu) {
				a[k] = (a[k] ^ c2) & ~c1;                             // LOP3
				a[k] = (uint32_t)max((int)a[k], (int)c2);             // VIMNMX.S32
				a[k] = (a[k] | c1) ^ (uint32_t)it;                    // LOP3
				a[k] = (uint32_t)min((int)a[k], (int)c1);             // VIMNMX.S32
			} else {
				a[k] = a[k] * one + c1;                               // IMAD
				a[k] = a[k] * one + c2;
				a[k] = a[k] * one + (uint32_t)it;
				a[k] = a[k] * one + c1;
			}
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < 16; ++k) r ^= a[k];
	if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// Grow-only device buffer from the device's stream-ordered pool: growing it is no device-wide synchronisation (cudaFree is, and
// with several contexts at work on one GPU every one of them would stall whenever one buffer grows).
struct DevBuf {
	void *p = nullptr; size_t cap = 0; cudaStream_t st = nullptr;
	cudaError_t reserve(size_t bytes)
	{
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFreeAsync(p, st);                               // (whatever used it was waited for before the call that grows it)
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 4 + 256;
		cudaError_t e = cudaMallocAsync(&p, want, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);        // the buffer is used from other streams of the context too
		if (e == cudaSuccess) cap = want; else p = nullptr;
		return e;
	}
	void release() { if (p) { cudaFreeAsync(p, st); cudaStreamSynchronize(st); } p = nullptr; cap = 0; }
};

enum { V_TRIVIAL = 0, V_GENERIC = 1, V_FAST0 = 2 };   // team variants: V_FAST0 + 2*log2(TEAM/2) + wrap
constexpr int N_VARIANTS = V_FAST0 + 10;

} // namespace

struct pansvr_ksw_ctx {
	int device = 0, sm_count = 0;
	int smem_optin = 227 * 1024;      // cudaDevAttrMaxSharedMemoryPerBlockOptin
	cudaStream_t stream = nullptr;
	// the team variants of one batch run side by side on their own streams (each launch fills the device with persistent CTAs, so
	// what overlaps are their tails: a batch of a few hundred thousand short tasks is otherwise four launches with four tails)
	cudaStream_t vstream[4] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t vev_in = nullptr, vev_done[4] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	DevBuf qseq, tseq, qoff, toff, qlen, tlen, res, cigar, order, counters, tb, gscratch, plan_variant, plan_rows, plan_stats;
	DevBuf tb_v[16];                 // traceback scratch per variant (they run concurrently)
	pansvr_ksw_stats_t stats;
};

namespace {

// The per-task plan of a batch, made on the device from the lengths that are there anyway: the kernel variant of every task,
// task ids grouped by variant and, inside a variant, by anti-diagonal count (most first: 256 levels), per-variant counts and the
// largest row count / query / target length.  Four small kernels and one 240-byte download; the order inside a level is whatever
// the atomics give (results do not depend on it: every task writes its own row of the result arrays).
struct BatchPlan {
	bool identity;               // every task has the same shape: one variant, tasks in input order, no order[]
	int begin[N_VARIANTS + 1];
	int max_rows[N_VARIANTS], max_qlen[N_VARIANTS], max_tlen[N_VARIANTS];
};
enum { PS_COUNT = 0, PS_ROWS = N_VARIANTS, PS_QLEN = 2 * N_VARIANTS, PS_TLEN = 3 * N_VARIANTS, PS_UNIFORM = 4 * N_VARIANTS, PS_WORDS = 4 * N_VARIANTS + 4 };
struct PlanArgs { int n; const int32_t *qlen, *tlen; int w; int trivial, fast_params, nowrap_ok, smem_optin; uint8_t *variant; int *rows; int *stats; int *bins; int *order; };

__device__ __forceinline__ bool d_team_fits(int team, int qlen, int smem_optin)
{
	const long per_cta = ((long)kswteam::team_smem_bytes(team, qlen) * (32 / team) + 32 * 32) * WARPS_PER_CTA + 512;
	return per_cta <= (long)smem_optin;
}

__global__ void __launch_bounds__(256) plan_classify_kernel(PlanArgs a)
{
	__shared__ int s_stat[4 * N_VARIANTS];
	__shared__ int s_uniform;
	for (int k = threadIdx.x; k < 4 * N_VARIANTS; k += blockDim.x) s_stat[k] = 0;
	if (threadIdx.x == 0) s_uniform = 1;
	__syncthreads();
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.n) {
		const int ql = a.qlen[i], tl = a.tlen[i];
		int v, rows = 0;
		if (a.trivial || ql <= 0 || tl <= 0) v = V_TRIVIAL;
		else {
			rows = kswhost::n_diagonals(ql, tl, a.w);
			const int team = kswhost::pick_team(ql, tl, a.w);
			if (!a.fast_params || team == 0 || ql > 8000 || !d_team_fits(team, ql, a.smem_optin)) v = V_GENERIC;
			else {
				const bool wrap = !a.nowrap_ok || kswhost::band_clips(ql, tl, a.w);
				int lg = 0;
				while ((2 << lg) < team) ++lg;
				v = V_FAST0 + 2 * lg + (wrap ? 1 : 0);
			}
		}
		a.variant[i] = (uint8_t)v; a.rows[i] = rows;
		atomicAdd(&s_stat[PS_COUNT + v], 1);
		atomicMax(&s_stat[PS_ROWS + v], rows); atomicMax(&s_stat[PS_QLEN + v], ql); atomicMax(&s_stat[PS_TLEN + v], tl);
		if (ql != a.qlen[0] || tl != a.tlen[0]) s_uniform = 0;
	}
	__syncthreads();
	for (int k = threadIdx.x; k < 4 * N_VARIANTS; k += blockDim.x) {
		const int x = s_stat[k];
		if (x) { if (k < N_VARIANTS) atomicAdd(&a.stats[k], x); else atomicMax(&a.stats[k], x); }
	}
	if (threadIdx.x == 0 && !s_uniform) a.stats[PS_UNIFORM] = 1;       // (1 = the tasks are not all of one shape)
}

__device__ __forceinline__ int plan_level(const PlanArgs &a, int i, int v)
{
	const int mr = a.stats[PS_ROWS + v] > 1 ? a.stats[PS_ROWS + v] : 1;
	return 255 - (int)((long long)a.rows[i] * 255 / mr);
}

__global__ void __launch_bounds__(256) plan_hist_kernel(PlanArgs a)          // tasks per (variant, level)
{
	__shared__ int s_bin[N_VARIANTS * 256];
	for (int k = threadIdx.x; k < N_VARIANTS * 256; k += blockDim.x) s_bin[k] = 0;
	__syncthreads();
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.n) { const int v = a.variant[i]; atomicAdd(&s_bin[v * 256 + plan_level(a, i, v)], 1); }
	__syncthreads();
	for (int k = threadIdx.x; k < N_VARIANTS * 256; k += blockDim.x) if (s_bin[k]) atomicAdd(&a.bins[k], s_bin[k]);
}

__global__ void __launch_bounds__(1024) plan_scan_kernel(PlanArgs a)          // bins -> where each (variant, level) starts in order[]
{
	__shared__ int s_part[1024];
	constexpr int NB = N_VARIANTS * 256, PER = (NB + 1023) / 1024;
	int local[PER], sum = 0;
	for (int k = 0; k < PER; ++k) { const int b = threadIdx.x * PER + k; local[k] = b < NB ? a.bins[b] : 0; sum += local[k]; }
	s_part[threadIdx.x] = sum;
	__syncthreads();
	for (int o = 1; o < 1024; o <<= 1) {
		const int x = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
		__syncthreads();
		s_part[threadIdx.x] += x;
		__syncthreads();
	}
	int run = s_part[threadIdx.x] - sum;
	for (int k = 0; k < PER; ++k) { const int b = threadIdx.x * PER + k; if (b < NB) { a.bins[b] = run; run += local[k]; } }
}

__global__ void __launch_bounds__(256) plan_scatter_kernel(PlanArgs a)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n) return;
	const int v = a.variant[i];
	a.order[atomicAdd(&a.bins[v * 256 + plan_level(a, i, v)], 1)] = i;
}

int plan_batch(pansvr_ksw_ctx *ctx, const kswhost::Plan &pl, int64_t n, const int32_t *d_qlen, const int32_t *d_tlen, BatchPlan &bp)
{
	CU(ctx->plan_variant.reserve((size_t)n));
	CU(ctx->plan_rows.reserve(sizeof(int) * (size_t)n));
	CU(ctx->plan_stats.reserve(sizeof(int) * (PS_WORDS + N_VARIANTS * 256)));
	CU(ctx->order.reserve(sizeof(int) * (size_t)n));
	PlanArgs a;
	a.n = (int)n; a.qlen = d_qlen; a.tlen = d_tlen; a.w = pl.P.w;
	a.trivial = pl.trivial; a.fast_params = pl.fast_params; a.nowrap_ok = pl.nowrap_ok; a.smem_optin = ctx->smem_optin;
	a.variant = (uint8_t*)ctx->plan_variant.p; a.rows = (int*)ctx->plan_rows.p; a.stats = (int*)ctx->plan_stats.p; a.bins = a.stats + PS_WORDS;
	a.order = (int*)ctx->order.p;
	const unsigned grid = (unsigned)((n + 255) / 256);
	CU(cudaMemsetAsync(a.stats, 0, sizeof(int) * (PS_WORDS + N_VARIANTS * 256), ctx->stream));
	plan_classify_kernel<<<grid, 256, 0, ctx->stream>>>(a);
	plan_hist_kernel<<<grid, 256, 0, ctx->stream>>>(a);
	plan_scan_kernel<<<1, 1024, 0, ctx->stream>>>(a);
	plan_scatter_kernel<<<grid, 256, 0, ctx->stream>>>(a);
	CU(cudaGetLastError());
	ctx->stats.kernel_launches += 4;
	int h[PS_WORDS];
	CU(cudaMemcpyAsync(h, a.stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	bp.begin[0] = 0;
	int used = 0, which = -1;
	for (int v = 0; v < N_VARIANTS; ++v) {
		bp.begin[v + 1] = bp.begin[v] + h[PS_COUNT + v];
		bp.max_rows[v] = h[PS_ROWS + v]; bp.max_qlen[v] = h[PS_QLEN + v]; bp.max_tlen[v] = h[PS_TLEN + v];
		if (h[PS_COUNT + v]) { ++used; which = v; }
	}
	// uniform batches (the fixed-read-length case): the tasks run in input order, no indirection
	bp.identity = h[PS_UNIFORM] == 0 && used == 1 && which >= V_FAST0;
	return 0;
}

template <int TEAM, bool WRAP, bool WC>
int launch_team_wc(pansvr_ksw_ctx *ctx, KArgs a, int max_rows, int max_qlen, cudaStream_t st, DevBuf &tbuf)
{
	constexpr int W = 16 * TEAM, NT = 32 / TEAM;
	a.smem_per_team = kswteam::team_smem_bytes(TEAM, max_qlen);
	const int smem = (a.smem_per_team * NT + 32 * 32) * WARPS_PER_CTA + 512;
	auto kern = ksw_team_kernel<TEAM, WRAP, WC>;
	// The limit is a property of the function, not of the launch: several contexts launch this kernel at the same time with
	// different sizes (sub-blocks in flight, each with its own longest query), and a launch fails with "invalid argument" if
	// another thread lowered the limit between this thread's setting it and its launch.  So everybody sets the same value, the
	// most a CTA may have; what a launch uses is its own `smem`.
	if (smem > ctx->smem_optin) return fail(PANSVR_E_UNSUPPORTED, "ksw team kernel does not fit an SM (query too long)");
	CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin));
	int per_sm = 0;
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
	if (per_sm < 1) return fail(PANSVR_E_UNSUPPORTED, "ksw team kernel does not fit an SM (query too long)");
	int grid = ctx->sm_count * per_sm;
	grid = (int)std::min<int64_t>(grid, ((int64_t)a.n + WARPS_PER_CTA * NT - 1) / (WARPS_PER_CTA * NT));
	a.tb_per_team = (a.P.flag & kswfast::F_SCORE_ONLY) ? 0 : (((size_t)max_rows + 1) * W + 255) & ~(size_t)255;
	const size_t tb_cap = (size_t)24 << 30;               // keep the traceback scratch under 24 GiB
	while (grid > 1 && a.tb_per_team * (size_t)grid * WARPS_PER_CTA * NT > tb_cap) grid = (grid + 1) / 2;
	CU(tbuf.reserve(a.tb_per_team * (size_t)grid * WARPS_PER_CTA * NT + 256));
	a.tb = (uint8_t*)tbuf.p;
	ctx->stats.tb_bytes_per_warp = (int64_t)a.tb_per_team * NT;
	ctx->stats.resident_warps = (int64_t)grid * WARPS_PER_CTA;
	kern<<<grid, THREADS, smem, st>>>(a);
	CU(cudaGetLastError());
	++ctx->stats.kernel_launches;
	return 0;
}

template <int TEAM, bool WRAP>
int launch_team(pansvr_ksw_ctx *ctx, const KArgs &a, int max_rows, int max_qlen, cudaStream_t st, DevBuf &tbuf)
{
	return (a.P.flag & kswfast::F_SCORE_ONLY) ? launch_team_wc<TEAM, WRAP, false>(ctx, a, max_rows, max_qlen, st, tbuf)
	                                          : launch_team_wc<TEAM, WRAP, true>(ctx, a, max_rows, max_qlen, st, tbuf);
}

// everything after the inputs are on the device: plan, launch each variant, leave results on the device
int run_device(pansvr_ksw_ctx *ctx, int64_t n, const uint8_t *d_qseq, const int64_t *d_qoff, const int32_t *d_qlen,
               const uint8_t *d_tseq, const int64_t *d_toff, const int32_t *d_tlen, const int32_t *h_qlen,
               const int32_t *h_tlen, const pansvr_ksw_params_t *pr, int32_t *d_res, uint32_t *d_cigar, int cigar_cap)
{
	if (pr->m > 1 && !pr->mat) return fail(PANSVR_E_ARG, "params->mat is NULL");
	kswhost::Plan pl = kswhost::make_plan(pr->m, pr->mat, pr->gapo, pr->gape, pr->gapo2, pr->gape2, pr->w, pr->zdrop,
	                                      pr->end_bonus, pr->flag);
	(void)h_qlen; (void)h_tlen;                                   // (the plan is made on the device)
	BatchPlan bp;
	{ const int prc = plan_batch(ctx, pl, n, d_qlen, d_tlen, bp); if (prc) return prc; }
	CU(ctx->counters.reserve(sizeof(int) * N_VARIANTS));
	CU(cudaMemsetAsync(ctx->counters.p, 0, sizeof(int) * N_VARIANTS, ctx->stream));
	CU(cudaEventRecord(ctx->ev[1], ctx->stream));
	CU(cudaEventRecord(ctx->vev_in, ctx->stream));               // inputs and plan are in place: the variant streams start from here
	int n_side = 0;
	KArgs a;
	a.P = pl.P;
	a.qseq = d_qseq; a.qoff = d_qoff; a.qlen = d_qlen; a.tseq = d_tseq; a.toff = d_toff; a.tlen = d_tlen;
	a.res = d_res; a.cigar = d_cigar; a.cigar_cap = cigar_cap; a.tb = nullptr; a.tb_per_team = 0; a.smem_per_team = 0;
	for (int v = 0; v < N_VARIANTS; ++v) {
		const int cnt = bp.begin[v + 1] - bp.begin[v];
		if (cnt == 0) continue;
		a.n = cnt;
		a.order = bp.identity ? nullptr : (const int*)ctx->order.p + bp.begin[v];
		a.counter = (int*)ctx->counters.p + v;
		int rc = 0;
		if (v == V_TRIVIAL) {
			ksw_reset_kernel<<<(cnt + 255) / 256, 256, 0, ctx->stream>>>(cnt, a.order, d_res, d_cigar, cigar_cap);
			CU(cudaGetLastError());
			++ctx->stats.kernel_launches;
			ctx->stats.tasks_trivial += cnt;
		} else if (v == V_GENERIC) {
			rc = kswgeneric::launch(ctx->stream, ctx->sm_count, pr, pl.P, cnt, a.order, d_qseq, d_qoff, d_qlen, d_tseq, d_toff,
			                        d_tlen, d_res, d_cigar, cigar_cap, bp.max_qlen[v], bp.max_tlen[v], &ctx->gscratch.p,
			                        &ctx->gscratch.cap, &ctx->stats.kernel_launches, g_err);
			ctx->stats.tasks_generic += cnt;
		} else {
			const int lg = (v - V_FAST0) >> 1, wrap = (v - V_FAST0) & 1;
			const int mr = bp.max_rows[v], mq = bp.max_qlen[v];
			const int side = n_side++ & 3;
			cudaStream_t st = ctx->vstream[side];
			if (n_side > 4) CU(cudaStreamWaitEvent(st, ctx->vev_done[side], 0));      // (more than four variants: queue behind the earlier one)
			else CU(cudaStreamWaitEvent(st, ctx->vev_in, 0));
			DevBuf &tbuf = ctx->tb_v[v];
			switch (lg * 2 + wrap) {
			case 0: rc = launch_team<2, false>(ctx, a, mr, mq, st, tbuf); break;
			case 1: rc = launch_team<2, true>(ctx, a, mr, mq, st, tbuf); break;
			case 2: rc = launch_team<4, false>(ctx, a, mr, mq, st, tbuf); break;
			case 3: rc = launch_team<4, true>(ctx, a, mr, mq, st, tbuf); break;
			case 4: rc = launch_team<8, false>(ctx, a, mr, mq, st, tbuf); break;
			case 5: rc = launch_team<8, true>(ctx, a, mr, mq, st, tbuf); break;
			case 6: rc = launch_team<16, false>(ctx, a, mr, mq, st, tbuf); break;
			case 7: rc = launch_team<16, true>(ctx, a, mr, mq, st, tbuf); break;
			case 8: rc = launch_team<32, false>(ctx, a, mr, mq, st, tbuf); break;
			default: rc = launch_team<32, true>(ctx, a, mr, mq, st, tbuf); break;
			}
			if (rc == 0) CU(cudaEventRecord(ctx->vev_done[side], st));
			(wrap ? ctx->stats.tasks_fast_wrap : ctx->stats.tasks_fast_nowrap) += cnt;
		}
		if (rc) return rc;
	}
	for (int k = 0; k < std::min(n_side, 4); ++k) CU(cudaStreamWaitEvent(ctx->stream, ctx->vev_done[k], 0));     // join
	CU(cudaEventRecord(ctx->ev[2], ctx->stream));
	return 0;
}

int finish_stats(pansvr_ksw_ctx *ctx)
{
	CU(cudaEventRecord(ctx->ev[3], ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	float k = 0, t = 0;
	CU(cudaEventElapsedTime(&k, ctx->ev[1], ctx->ev[2]));
	CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[3]));
	ctx->stats.kernel_ms = k; ctx->stats.total_ms = t;
	return 0;
}

} // namespace

extern "C" {

const char *pansvr_last_error(void) { return g_err.c_str(); }

extern "C" int pansvr_ksw_create_prio(int device, int high_priority, pansvr_ksw_ctx **out);
void pansvr_ksw_destroy(pansvr_ksw_ctx *c);
int pansvr_ksw_create(int device, pansvr_ksw_ctx **out) { return pansvr_ksw_create_prio(device, 0, out); }

// (not part of the ABI in include/: the aln stage makes the context of its host path with high-priority streams, so that the few
// tasks of the pairs handed back to the host do not queue behind the bulk kernels of the sub-blocks in flight)
int pansvr_ksw_create_prio(int device, int high_priority, pansvr_ksw_ctx **out)
{
	if (!out) return fail(PANSVR_E_ARG, "out is NULL");
	*out = nullptr;
	int ndev = 0;
	CU(cudaGetDeviceCount(&ndev));
	if (device < 0 || device >= ndev) return fail(PANSVR_E_CUDA, "no such CUDA device");
	CU(cudaSetDevice(device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) return fail(PANSVR_E_CUDA, std::string("device is not sm_100 class: ") + prop.name);
	pansvr_ksw_ctx *c = new pansvr_ksw_ctx();
	// (a failure from here on must not leak the half-made context)
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { pansvr_ksw_destroy(c); \
	return fail(PANSVR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
	c->device = device; c->sm_count = prop.multiProcessorCount;
	if (cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) c->smem_optin = 227 * 1024;
	memset(&c->stats, 0, sizeof(c->stats));
	int prio_lo = 0, prio_hi = 0;
	CUX(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	const int prio = high_priority ? prio_hi : prio_lo;
	CUX(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio));
	for (auto &vs : c->vstream) CUX(cudaStreamCreateWithPriority(&vs, cudaStreamNonBlocking, prio));
	for (DevBuf *b : {&c->qseq, &c->tseq, &c->qoff, &c->toff, &c->qlen, &c->tlen, &c->res, &c->cigar, &c->order, &c->counters, &c->tb, &c->plan_variant, &c->plan_rows, &c->plan_stats}) b->st = c->stream;
	for (DevBuf &b : c->tb_v) b.st = c->stream;
	{   // freed buffers stay in the pool (the default threshold hands them back to the driver at the next synchronisation)
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { uint64_t keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
	}
	CUX(cudaEventCreateWithFlags(&c->vev_in, cudaEventDisableTiming));
	for (auto &e : c->vev_done) CUX(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	for (auto &e : c->ev) CUX(cudaEventCreate(&e));
#undef CUX
	*out = c;
	return 0;
}

void pansvr_ksw_destroy(pansvr_ksw_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	for (auto &vs : c->vstream) if (vs) { cudaStreamSynchronize(vs); cudaStreamDestroy(vs); }
	if (c->vev_in) cudaEventDestroy(c->vev_in);
	for (auto &e : c->vev_done) if (e) cudaEventDestroy(e);
	for (DevBuf &b : c->tb_v) b.release();
	for (DevBuf *b : {&c->qseq, &c->tseq, &c->qoff, &c->toff, &c->qlen, &c->tlen, &c->res, &c->cigar, &c->order, &c->counters,
	                  &c->tb, &c->plan_variant, &c->plan_rows, &c->plan_stats}) b->release();
	if (c->gscratch.p) cudaFree(c->gscratch.p);                   // (the generic kernel's launcher sizes this one itself, with cudaMalloc)
	for (auto &e : c->ev) if (e) cudaEventDestroy(e);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

void *pansvr_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
	return p;
}
void pansvr_host_free(void *p) { if (p) cudaFreeHost(p); }

int64_t pansvr_ksw_band_cells(int32_t qlen, int32_t tlen, int32_t w)
{
	if (qlen <= 0 || tlen <= 0) return 0;
	if (w < 0) w = std::max(qlen, tlen);
	int64_t n = 0;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int lo, hi;
		kswfast::band(r, qlen, tlen, w, lo, hi);
		if (lo > hi) break;
		n += hi - lo + 1;
	}
	return n;
}

// A call that fails half way must not leave copies or kernels in flight that still touch the caller's buffers (or the context's,
// which the next call may grow): whatever was queued is waited for before the error is returned.
static int drained(pansvr_ksw_ctx *ctx, int rc)
{
	if (rc != 0 && ctx) {
		const std::string keep = g_err;
		cudaStreamSynchronize(ctx->stream);
		for (auto &vs : ctx->vstream) if (vs) cudaStreamSynchronize(vs);
		cudaGetLastError();
		g_err = keep;
	}
	return rc;
}

static int batch_device_impl(pansvr_ksw_ctx *ctx, int64_t n, const uint8_t *d_qseq, const int64_t *d_qoff,
                                  const int32_t *d_qlen, const uint8_t *d_tseq, const int64_t *d_toff, const int32_t *d_tlen,
                                  const int32_t *h_qlen, const int32_t *h_tlen, const pansvr_ksw_params_t *params,
                                  int32_t *d_results, uint32_t *d_cigar, int32_t cigar_cap)
{
	if (!ctx || !params || n < 0 || n > 0x7fffffff || cigar_cap < 0) return fail(PANSVR_E_ARG, "bad argument");
	CU(cudaSetDevice(ctx->device));
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	if (n > 0) {
		int rc = run_device(ctx, n, d_qseq, d_qoff, d_qlen, d_tseq, d_toff, d_tlen, h_qlen, h_tlen, params, d_results, d_cigar, cigar_cap);
		if (rc) return rc;
	} else { CU(cudaEventRecord(ctx->ev[1], ctx->stream)); CU(cudaEventRecord(ctx->ev[2], ctx->stream)); }
	return finish_stats(ctx);
}

int pansvr_ksw_extd2_batch_device(pansvr_ksw_ctx *ctx, int64_t n, const uint8_t *d_qseq, const int64_t *d_qoff,
                                  const int32_t *d_qlen, const uint8_t *d_tseq, const int64_t *d_toff, const int32_t *d_tlen,
                                  const int32_t *h_qlen, const int32_t *h_tlen, const pansvr_ksw_params_t *params,
                                  int32_t *d_results, uint32_t *d_cigar, int32_t cigar_cap)
{
	return drained(ctx, batch_device_impl(ctx, n, d_qseq, d_qoff, d_qlen, d_tseq, d_toff, d_tlen, h_qlen, h_tlen, params, d_results, d_cigar, cigar_cap));
}

static int batch_impl(pansvr_ksw_ctx *ctx, int64_t n, const uint8_t *qseq, int64_t qseq_bytes, const int64_t *qoff,
                           const int32_t *qlen, const uint8_t *tseq, int64_t tseq_bytes, const int64_t *toff,
                           const int32_t *tlen, const pansvr_ksw_params_t *params, int32_t *results, uint32_t *cigar,
                           int32_t cigar_cap)
{
	if (!ctx || !params || n < 0 || n > 0x7fffffff || cigar_cap < 0 || qseq_bytes < 0 || tseq_bytes < 0)
		return fail(PANSVR_E_ARG, "bad argument");
	CU(cudaSetDevice(ctx->device));
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	if (n == 0) { CU(cudaEventRecord(ctx->ev[1], ctx->stream)); CU(cudaEventRecord(ctx->ev[2], ctx->stream)); return finish_stats(ctx); }
	for (int64_t i = 0; i < n; ++i) {
		if (qlen[i] > 0 && (qoff[i] < 0 || qoff[i] + qlen[i] > qseq_bytes)) return fail(PANSVR_E_ARG, "query window outside qseq");
		if (tlen[i] > 0 && (toff[i] < 0 || toff[i] + tlen[i] > tseq_bytes)) return fail(PANSVR_E_ARG, "target window outside tseq");
	}
	const size_t nn = (size_t)n;
	CU(ctx->qseq.reserve((size_t)qseq_bytes + 16)); CU(ctx->tseq.reserve((size_t)tseq_bytes + 16));
	CU(ctx->qoff.reserve(nn * 8)); CU(ctx->toff.reserve(nn * 8)); CU(ctx->qlen.reserve(nn * 4)); CU(ctx->tlen.reserve(nn * 4));
	CU(ctx->res.reserve(nn * sizeof(int32_t) * PANSVR_RES_WORDS));
	CU(ctx->cigar.reserve(nn * sizeof(uint32_t) * (size_t)std::max(cigar_cap, 1)));
	cudaStream_t s = ctx->stream;
	CU(cudaMemcpyAsync(ctx->qseq.p, qseq, (size_t)qseq_bytes, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(ctx->tseq.p, tseq, (size_t)tseq_bytes, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(ctx->qoff.p, qoff, nn * 8, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(ctx->toff.p, toff, nn * 8, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(ctx->qlen.p, qlen, nn * 4, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(ctx->tlen.p, tlen, nn * 4, cudaMemcpyHostToDevice, s));
	ctx->stats.h2d_bytes = qseq_bytes + tseq_bytes + (int64_t)nn * 24;
	int rc = run_device(ctx, n, (const uint8_t*)ctx->qseq.p, (const int64_t*)ctx->qoff.p, (const int32_t*)ctx->qlen.p,
	                    (const uint8_t*)ctx->tseq.p, (const int64_t*)ctx->toff.p, (const int32_t*)ctx->tlen.p, qlen, tlen, params,
	                    (int32_t*)ctx->res.p, (uint32_t*)ctx->cigar.p, cigar_cap);
	if (rc) return rc;
	CU(cudaMemcpyAsync(results, ctx->res.p, nn * sizeof(int32_t) * PANSVR_RES_WORDS, cudaMemcpyDeviceToHost, s));
	if (cigar_cap > 0 && !(params->flag & PANSVR_KSW_SCORE_ONLY))
		CU(cudaMemcpyAsync(cigar, ctx->cigar.p, nn * sizeof(uint32_t) * (size_t)cigar_cap, cudaMemcpyDeviceToHost, s));
	ctx->stats.d2h_bytes = (int64_t)(nn * sizeof(int32_t) * PANSVR_RES_WORDS) +
	                       ((cigar_cap > 0 && !(params->flag & PANSVR_KSW_SCORE_ONLY)) ? (int64_t)(nn * 4 * (size_t)cigar_cap) : 0);
	return finish_stats(ctx);
}

int pansvr_ksw_extd2_batch(pansvr_ksw_ctx *ctx, int64_t n, const uint8_t *qseq, int64_t qseq_bytes, const int64_t *qoff,
                           const int32_t *qlen, const uint8_t *tseq, int64_t tseq_bytes, const int64_t *toff,
                           const int32_t *tlen, const pansvr_ksw_params_t *params, int32_t *results, uint32_t *cigar,
                           int32_t cigar_cap)
{
	return drained(ctx, batch_impl(ctx, n, qseq, qseq_bytes, qoff, qlen, tseq, tseq_bytes, toff, tlen, params, results, cigar, cigar_cap));
}

int pansvr_int_alu_peak(pansvr_ksw_ctx *ctx, double *gops)
{
	if (!ctx || !gops) return fail(PANSVR_E_ARG, "bad argument");
	CU(cudaSetDevice(ctx->device));
	CU(ctx->counters.reserve(4096));
	const int iters = 4096, grid = ctx->sm_count * 8, ops_per_iter = 16 * 4;
	double best = 0;
	for (int rep = 0; rep < 5; ++rep) {
		CU(cudaEventRecord(ctx->ev[0], ctx->stream));
		int_alu_peak_kernel<<<grid, 256, 0, ctx->stream>>>((uint32_t*)ctx->counters.p, iters, 0x2545f491u + rep);
		CU(cudaGetLastError());
		CU(cudaEventRecord(ctx->ev[3], ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
		float ms = 0;
		CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
		const double g = (double)grid * 256 * iters * ops_per_iter / (ms * 1e-3) * 1e-9;
		if (rep > 0 && g > best) best = g;
	}
	*gops = best;
	return 0;
}

int pansvr_int_pipe_peaks(pansvr_ksw_ctx *ctx, double out[4])
{
	if (!ctx || !out) return fail(PANSVR_E_ARG, "bad argument");
	int rc = pansvr_int_alu_peak(ctx, &out[0]);
	if (rc) return rc;
	const int iters = 4096, grid = ctx->sm_count * 8, ops_per_iter = 16 * 4;
	for (int mode = 1; mode <= 3; ++mode) {
		double best = 0;
		for (int rep = 0; rep < 4; ++rep) {
			CU(cudaEventRecord(ctx->ev[0], ctx->stream));
			const uint32_t seed = 0x2545f491u + rep, one = (uint32_t)(rep >= 0);
			if (mode == 1) int_pipe_peak_kernel<1><<<grid, 256, 0, ctx->stream>>>((uint32_t*)ctx->counters.p, iters, seed, one);
			else if (mode == 2) int_pipe_peak_kernel<2><<<grid, 256, 0, ctx->stream>>>((uint32_t*)ctx->counters.p, iters, seed, one);
			else int_pipe_peak_kernel<3><<<grid, 256, 0, ctx->stream>>>((uint32_t*)ctx->counters.p, iters, seed, one);
			CU(cudaGetLastError());
			CU(cudaEventRecord(ctx->ev[3], ctx->stream));
			CU(cudaStreamSynchronize(ctx->stream));
			float ms = 0;
			CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
			const double g = (double)grid * 256 * iters * ops_per_iter / (ms * 1e-3) * 1e-9;
			if (rep > 0 && g > best) best = g;
		}
		out[mode] = best;
	}
	return 0;
}

int pansvr_ksw_last_stats(const pansvr_ksw_ctx *ctx, pansvr_ksw_stats_t *out)
{
	if (!ctx || !out) return fail(PANSVR_E_ARG, "bad argument");
	*out = ctx->stats;
	return 0;
}

void pansvr_ksw_extd2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                      int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int end_bonus, int flag,
                      pansvr_ksw_extz_t *ez)
{
	(void)km;
	// one context per calling thread, like the reference's per-thread KSW_ALN_handler; it goes when its thread ends (the main
	// thread's lives as long as the process: CUDA is on its way out by the time that thread's locals are destroyed)
	static thread_local struct Holder {
		pansvr_ksw_ctx *c = nullptr;
		~Holder() { if (c && (long)syscall(SYS_gettid) != (long)getpid()) pansvr_ksw_destroy(c); }
	} holder;
	pansvr_ksw_ctx *&ctx = holder.c;
	if (!ctx) {
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || pansvr_ksw_create(dev, &ctx) != 0) {
			fprintf(stderr, "pansvr_b200: ksw_extd2 needs a B200 and there is no CPU fallback: %s\n", pansvr_last_error());
			abort();
		}
	}
	pansvr_ksw_params_t pr;
	pr.m = m; pr.mat = mat; pr.gapo = gapo; pr.gape = gape; pr.gapo2 = gapo2; pr.gape2 = gape2;
	pr.w = w; pr.zdrop = zdrop; pr.end_bonus = end_bonus; pr.flag = flag;
	const int64_t zero = 0;
	int32_t res[PANSVR_RES_WORDS];
	int cap = std::max(ez->m_cigar, 16);
	std::vector<uint32_t> cig;
	for (;;) {
		cig.assign((size_t)cap, 0);
		const int32_t ql = qlen, tl = tlen;
		int rc = pansvr_ksw_extd2_batch(ctx, 1, query, qlen > 0 ? qlen : 0, &zero, &ql, target, tlen > 0 ? tlen : 0, &zero, &tl,
		                                &pr, res, cig.data(), cap);
		if (rc != 0) { fprintf(stderr, "pansvr_b200: ksw_extd2 failed: %s\n", pansvr_last_error()); abort(); }
		if (!(res[PANSVR_RES_STATUS] & 1)) break;
		cap = res[PANSVR_RES_N_CIGAR] + 1;
	}
	ez->max = (uint32_t)res[0]; ez->zdropped = (uint32_t)res[1]; ez->max_q = res[2]; ez->max_t = res[3];
	ez->mqe = res[4]; ez->mqe_t = res[5]; ez->mte = res[6]; ez->mte_q = res[7]; ez->score = res[8];
	ez->n_cigar = res[9]; ez->reach_end = res[10];
	if (ez->n_cigar > 0) {                  // grow ez->cigar the way ksw_push_cigar does (ksw2.h:106-116)
		int mc = ez->m_cigar;
		while (mc < ez->n_cigar) mc = mc ? mc << 1 : 4;
		if (mc != ez->m_cigar) { ez->cigar = (uint32_t*)realloc(ez->cigar, (size_t)mc << 2); ez->m_cigar = mc; }
		memcpy(ez->cigar, cig.data(), sizeof(uint32_t) * (size_t)ez->n_cigar);
	}
}

void ksw_extd2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                   int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int end_bonus, int flag,
                   pansvr_ksw_extz_t *ez)
{
	pansvr_ksw_extd2(km, qlen, query, tlen, target, m, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag, ez);
}

} // extern "C"
