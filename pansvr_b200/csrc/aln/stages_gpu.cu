// stages_gpu.cu -- CUDA backend of the device stages (stages_run.hpp): one thread per read / strand / candidate, grow-only
// buffers in HBM, cub prefix sums, everything of one block on one stream.  The per-read work of these stages is a few hundred
// dependent integer operations over a few hundred bytes, so the kernels are bound by HBM/L2 latency with every SM full of
// reads in flight -- the layout goal is that a read's data is touched once per stage and never leaves the device in between.
#include <cuda_runtime.h>
#include <unistd.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/pansvr_b200.h"
#include "stages_run.hpp"

namespace pansvr {

// device pointers of the seed service (seed_gpu.cu)
const IndexView &seed_service_view(const SeedService *s);

namespace {

template <class F> __global__ void __launch_bounds__(128) for_each_kernel(F f, size_t n)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) f(i);
}

template <class F> __global__ void __launch_bounds__(128) for_each_in_kernel(F f, const uint32_t *perm, size_t n)
{
	const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) f((size_t)perm[k]);
}

// SAM text, write pass: every thread formats the short fields of its record and sets its (up to four) long copies aside; the
// warp then runs the 32 x 4 copies with all lanes on one copy at a time, so that a 150-byte field is read and written as
// contiguous segments instead of one byte per lane at 32 different places.
__global__ void __launch_bounds__(128) text_write_kernel(FnText f, size_t n)
{
	__shared__ CopyJob jobs[4][32][4];
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	JobSink s;
	s.p = nullptr; s.n = 0; s.nj = 0; s.np = 0;
	if (i < n) f.prepare(i, s);
	for (int j = 0; j < 4; ++j) {
		CopyJob J;
		if (j < s.nj) J = s.job[j]; else { J.src = nullptr; J.dst = nullptr; J.len = 0; J.mode = 0; }
		jobs[warp][lane][j] = J;
	}
	__syncwarp();
	for (int l = 0; l < 32; ++l)
		for (int j = 0; j < 4; ++j) {
			const CopyJob J = jobs[warp][l][j];
			for (uint32_t k = (uint32_t)lane; k < J.len; k += 32) J.dst[k] = copy_byte(J.mode, J.src, k, J.len);
		}
	__syncwarp();
	for (int k = 0; k < s.np; ++k) s.p[s.patches[k]] = ',';
}

// Stage A: the census filter of every thread lives in shared memory, word k of thread t at [k * 128 + t]: whatever words the threads
// of a warp touch, they are in 32 different banks (a per-thread local array would put 32 random words in 32 different cache lines).
__global__ void __launch_bounds__(128) encode_kernel(FnEncode f, size_t n)
{
	__shared__ uint32_t filter[ENC_FILTER_WORDS * 128];
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) f.run(i, filter + threadIdx.x, 128);
}

struct CudaBackend {
	cudaStream_t st = nullptr;
	void *p[SL_COUNT]; size_t cap[SL_COUNT];
	bool failed = false; std::string why;
	std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
	struct Lap { cudaEvent_t a, b; int stage; };
	std::vector<Lap> laps;
	pansvr_ksw_ctx *ksw_ctx = nullptr; pansvr_ksw_params_t kp; int8_t mat[25];
	DevCounters dev;
	CudaBackend() { for (int i = 0; i < SL_COUNT; ++i) { p[i] = nullptr; cap[i] = 0; } }
	void check(cudaError_t e, const char *what) { if (e != cudaSuccess && !failed) { failed = true; why = std::string(what) + ": " + cudaGetErrorString(e); } }
	template <class T> T *buf(int slot, size_t n)
	{
		const size_t bytes = n * sizeof(T);
		if (bytes <= cap[slot]) return (T*)p[slot];
		// from the device's stream-ordered pool: growing a buffer is an operation of this block's stream like any other, not a
		// device-wide synchronisation (cudaFree is: every other block in flight would stall whenever a buffer grows)
		if (p[slot]) check(cudaFreeAsync(p[slot], st), "cudaFreeAsync");
		p[slot] = nullptr; cap[slot] = 0;
		const size_t want = bytes + bytes / 4 + 4096;
		cudaError_t e = cudaMallocAsync(&p[slot], want, st);
		if (e != cudaSuccess) { check(e, "cudaMallocAsync"); return nullptr; }
		cap[slot] = want;
		return (T*)p[slot];
	}
	void h2d(void *d, const void *h, size_t bytes) { if (bytes) { check(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync H2D"); dev.h2d_bytes += (int64_t)bytes; } }
	void d2h(void *h, const void *d, size_t bytes) { if (bytes) { check(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H"); dev.d2h_bytes += (int64_t)bytes; } }
	void zero(void *d, size_t bytes) { if (bytes) check(cudaMemsetAsync(d, 0, bytes, st), "cudaMemsetAsync"); }
	void sync() { check(cudaStreamSynchronize(st), "cudaStreamSynchronize"); }
	cudaEvent_t event()
	{
		if (ev_used == ev_pool.size()) { cudaEvent_t e; check(cudaEventCreate(&e), "cudaEventCreate"); ev_pool.push_back(e); }
		return ev_pool[ev_used++];
	}
	template <class F> void for_each(size_t n, const F &f, int stage)
	{
		if (n == 0 || failed) return;
		Lap l; l.a = event(); l.b = event(); l.stage = stage;
		check(cudaEventRecord(l.a, st), "cudaEventRecord");
		const unsigned threads = 128;
		for_each_kernel<F><<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(f, n);
		check(cudaGetLastError(), "kernel launch");
		check(cudaEventRecord(l.b, st), "cudaEventRecord");
		laps.push_back(l);
		++dev.launches;
	}
	template <class F> void for_each_in(size_t n, const uint32_t *perm, const F &f, int stage)
	{
		if (n == 0 || failed) return;
		Lap l; l.a = event(); l.b = event(); l.stage = stage;
		check(cudaEventRecord(l.a, st), "cudaEventRecord");
		for_each_in_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(f, perm, n);
		check(cudaGetLastError(), "kernel launch");
		check(cudaEventRecord(l.b, st), "cudaEventRecord");
		laps.push_back(l);
		++dev.launches;
	}
	void order_desc(const uint32_t *key, const uint32_t *idx, uint32_t *perm, size_t n)
	{
		if (n == 0 || failed) return;
		uint32_t *key_out = buf<uint32_t>(SL_SORT_KEYS, n);
		size_t tmp_bytes = 0;
		check(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, key, key_out, idx, perm, (int)n, 0, 8, st), "cub sort size");
		void *tmp = buf<uint8_t>(SL_SORT_TMP, tmp_bytes + 256);
		if (!key_out || !tmp) return;
		check(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, key, key_out, idx, perm, (int)n, 0, 8, st), "cub sort");
		++dev.launches;
	}
	void encode(size_t n, const FnEncode &f)
	{
		if (n == 0 || failed) return;
		Lap l; l.a = event(); l.b = event(); l.stage = 0;
		check(cudaEventRecord(l.a, st), "cudaEventRecord");
		encode_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(f, n);
		check(cudaGetLastError(), "kernel launch");
		check(cudaEventRecord(l.b, st), "cudaEventRecord");
		laps.push_back(l);
		++dev.launches;
	}
	void text_write(size_t n, const FnText &f)
	{
		if (n == 0 || failed) return;
		Lap l; l.a = event(); l.b = event(); l.stage = 7;
		check(cudaEventRecord(l.a, st), "cudaEventRecord");
		text_write_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(f, n);
		check(cudaGetLastError(), "kernel launch");
		check(cudaEventRecord(l.b, st), "cudaEventRecord");
		laps.push_back(l);
		++dev.launches;
	}
	void scan(const uint32_t *in, uint32_t *out, size_t n)
	{
		size_t tmp_bytes = 0;
		check(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, st), "cub scan size");
		void *tmp = buf<uint8_t>(SL_SCAN_TMP, tmp_bytes + 256);
		if (!tmp) return;
		check(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, st), "cub scan");
	}
	bool ksw(size_t n, const uint8_t *q, const int64_t *qoff, const int32_t *qlen, const uint8_t *t, const int64_t *toff, const int32_t *tlen,
	         int32_t *res, uint32_t *cig, int cigar_cap, std::string &err)
	{
		sync();
		if (failed) { err = why; return false; }
		// (the ksw context has its own stream; everything it reads was produced before the sync above, and the call returns
		// after its kernels have finished; the per-task plan is made on the device from the lengths there)
		const double th0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
		const int rc = pansvr_ksw_extd2_batch_device(ksw_ctx, (int64_t)n, q, qoff, qlen, t, toff, tlen, nullptr, nullptr, &kp, res, cig, cigar_cap);
		{
			FILE *tf = nullptr;
			if (trace_ref(st, &tf) && tf) fprintf(tf, "E %p %.6f %.6f %zu\n", (void*)this, th0, std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(), n);
		}
		if (rc != 0) { err = std::string("ksw batch: ") + pansvr_last_error(); return false; }
		pansvr_ksw_stats_t ks;
		if (pansvr_ksw_last_stats(ksw_ctx, &ks) == 0) { dev.launches += ks.kernel_launches; dev.h2d_bytes += ks.h2d_bytes; dev.ksw_kernel_ms += ks.kernel_ms; }
		return true;
	}
	// PANSVR_TRACE=<file>: every timed kernel of every sub-block with its start and end on one clock (ms since the first use), to see
	// what the device did when (diagnostics; the events exist anyway)
	static cudaEvent_t trace_ref(cudaStream_t st, FILE **fp)
	{
		static std::mutex m; static cudaEvent_t ref = nullptr; static FILE *f = nullptr; static bool tried = false;
		std::lock_guard<std::mutex> lk(m);
		if (!tried) {
			tried = true;
			if (const char *path = getenv("PANSVR_TRACE")) {
				f = fopen((std::string(path) + "." + std::to_string((long)getpid())).c_str(), "a");
				if (f && cudaEventCreate(&ref) == cudaSuccess) { cudaEventRecord(ref, st); cudaEventSynchronize(ref); fprintf(f, "# ref %.6f\n", std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count()); }
			}
		}
		*fp = f;
		return ref;
	}
	void collect_laps()
	{
		FILE *tf = nullptr;
		cudaEvent_t ref = trace_ref(st, &tf);
		for (const Lap &l : laps) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, l.a, l.b) != cudaSuccess) continue;
			if (ref && tf) { float t0 = 0; if (cudaEventElapsedTime(&t0, ref, l.a) == cudaSuccess) fprintf(tf, "K %p %d %.3f %.3f\n", (void*)this, l.stage, t0, t0 + ms); }
			if (l.stage == 1) dev.seed_kernel_ms += ms; else dev.stage_kernel_ms += ms;
			if (l.stage >= 0 && l.stage < 8) dev.by_stage_ms[l.stage] += ms;
		}
		laps.clear(); ev_used = 0;
	}
};

} // namespace

struct StageService {
	int device = 0;
	CudaBackend be;
	void *scan_tmp = nullptr; size_t scan_cap = 0;
	IndexView view;
	uint64_t *d_pos = nullptr, *d_ref = nullptr;
	uint32_t *d_csi = nullptr, *d_cen = nullptr; DevSv *d_sv = nullptr;
	RefView rf;
	PairIndexView pix;
	TextTables tt;                                                // target names and the anchors' tag strings, in device memory
	std::vector<void*> tt_mem;
	pansvr_ksw_ctx *own_ksw = nullptr;                            // every instance has its own ksw context (stream + scratch): no lock between blocks
};

StageService *stage_service_create(const DebgaIndex &idx, SeedService *seeds, void *ksw_ctx, int device, std::string &err)
{
	if (cudaSetDevice(device) != cudaSuccess) { err = "stage service: no usable CUDA device (the device stages have no CPU fallback)"; return nullptr; }
	StageService *s = new StageService();
	s->device = device;
	s->view = seed_service_view(seeds);
	(void)ksw_ctx;
	if (pansvr_ksw_create(device, &s->own_ksw) != 0) { err = std::string("stage service: ") + pansvr_last_error(); delete s; return nullptr; }
	s->be.ksw_ctx = s->own_ksw;
	std::vector<DevSv> svs(idx.sv_info.size());
	for (size_t i = 0; i < svs.size(); ++i) { svs[i].chr_id = idx.sv_info[i].chr_id; svs[i].st_pos = (uint32_t)idx.sv_info[i].st_pos; svs[i].end_offset = idx.sv_info[i].end_offset; svs[i].pad = 0; }
	auto up32 = [&](const void *h, size_t bytes, void **d) -> bool {
		if (cudaMalloc(d, bytes + 16) != cudaSuccess) return false;
		return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
	};
	auto up = [&](const std::vector<uint64_t> &h, uint64_t *&d) -> bool {
		if (cudaMalloc((void**)&d, (h.size() + 2) * 8) != cudaSuccess) return false;
		if (cudaMemset(d, 0, (h.size() + 2) * 8) != cudaSuccess) return false;
		return cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
	};
	if (cudaStreamCreateWithFlags(&s->be.st, cudaStreamNonBlocking) != cudaSuccess || !up(idx.pos, s->d_pos) || !up(idx.ref_seq, s->d_ref) ||
	    !up32(idx.chr_search_index.data(), idx.chr_search_index.size() * 4, (void**)&s->d_csi) || !up32(idx.chr_end_n.data(), idx.chr_end_n.size() * 4, (void**)&s->d_cen) ||
	    !up32(svs.data(), svs.size() * sizeof(DevSv), (void**)&s->d_sv)) {
		err = "stage service: CUDA set-up failed";
		stage_service_destroy(s);
		return nullptr;
	}
	s->rf.ref_seq = s->d_ref;
	s->pix.chr_search_index = s->d_csi; s->pix.chr_end_n = s->d_cen; s->pix.sv = s->d_sv;
	auto table = [&](const std::vector<std::string> &v, StrTable &t) -> bool {
		std::vector<uint32_t> off(v.size() + 1, 0);
		std::string pool;
		for (size_t i = 0; i < v.size(); ++i) { pool += v[i]; off[i + 1] = (uint32_t)pool.size(); }
		void *dp = nullptr, *dof = nullptr;
		if (!up32(pool.data(), pool.size(), &dp)) return false;
		s->tt_mem.push_back(dp);
		if (!up32(off.data(), off.size() * 4, &dof)) return false;
		s->tt_mem.push_back(dof);
		t.pool = (const char*)dp; t.off = (const uint32_t*)dof; t.n = (uint32_t)v.size();
		return true;
	};
	std::vector<std::string> prints(idx.sv_info.size()), ids(idx.sv_info.size());
	for (size_t i = 0; i < idx.sv_info.size(); ++i) { prints[i] = idx.sv_info[i].vcf_print; ids[i] = idx.sv_info[i].vcf_id; }
	if (!table(idx.target_names, s->tt.target_names) || !table(prints, s->tt.sv_print) || !table(ids, s->tt.sv_id)) {
		err = "stage service: CUDA set-up failed";
		stage_service_destroy(s);
		return nullptr;
	}
	s->tt.not_ori = 0;
	return s;
}

void stage_service_destroy(StageService *s)
{
	if (!s) return;
	cudaSetDevice(s->device);
	if (s->be.st) cudaStreamSynchronize(s->be.st);
	for (int i = 0; i < SL_COUNT; ++i) if (s->be.p[i]) cudaFreeAsync(s->be.p[i], s->be.st);
	if (s->be.st) cudaStreamSynchronize(s->be.st);
	for (cudaEvent_t e : s->be.ev_pool) cudaEventDestroy(e);
	if (s->d_pos) cudaFree(s->d_pos);
	if (s->d_ref) cudaFree(s->d_ref);
	for (void *q : {(void*)s->d_csi, (void*)s->d_cen, (void*)s->d_sv}) if (q) cudaFree(q);
	for (void *q : s->tt_mem) if (q) cudaFree(q);
	if (s->own_ksw) pansvr_ksw_destroy(s->own_ksw);
	if (s->be.st) cudaStreamDestroy(s->be.st);
	delete s;
}

void stage_service_set_scoring(StageService *s, const AlnScores &o, int zdrop)
{
	CudaBackend &be = s->be;
	const int8_t m = (int8_t)o.match, x = (int8_t)-o.mismatch;
	for (int a = 0, k = 0; a < 5; ++a) for (int b = 0; b < 5; ++b, ++k) be.mat[k] = (a == 4 || b == 4) ? 0 : (a == b ? m : x);   // ksw_gen_mat_D, RR:829-844
	be.kp.m = 5; be.kp.mat = be.mat; be.kp.gapo = (int8_t)o.gap_open; be.kp.gape = (int8_t)o.gap_ex; be.kp.gapo2 = (int8_t)o.gap_open2; be.kp.gape2 = (int8_t)o.gap_ex2;
	be.kp.w = 200; be.kp.zdrop = (uint16_t)zdrop; be.kp.end_bonus = -1; be.kp.flag = 0;               // copy_option, RR:817-827
}

bool stage_service_run(StageService *s, const DevStageIn &in, DevStageOut &out, std::string &err)
{
	if (cudaSetDevice(s->device) != cudaSuccess) { err = "cudaSetDevice failed"; return false; }
	CudaBackend &be = s->be;
	be.failed = false; be.why.clear(); be.dev = DevCounters();
	const bool ok = run_device_stages(be, s->view, s->d_pos, s->rf, s->pix, in, out, err);
	be.sync();
	be.collect_laps();
	out.dev.add(be.dev);
	if (be.failed) { err = "device stages: " + be.why; return false; }
	return ok;
}

bool stage_service_finalize(StageService *s, const PairOpts &o, int not_ori, size_t n_pairs, const int32_t *drawn, size_t n_drawn, const uint32_t *host_len,
                            const uint32_t *tie_pair, const DevPairState *tie_done, size_t n_ties, DevStageOut &out,
                            HostVec<char> &text_out, std::string &err)
{
	if (cudaSetDevice(s->device) != cudaSuccess) { err = "cudaSetDevice failed"; return false; }
	CudaBackend &be = s->be;
	be.failed = false; be.why.clear(); be.dev = DevCounters();
	TextTables T = s->tt;
	T.not_ori = not_ori;
	const bool ok = run_device_finalize(be, s->pix, o, T, n_pairs, drawn, n_drawn, host_len, tie_pair, tie_done, n_ties, out, text_out, err);
	be.sync();
	be.collect_laps();
	out.dev.add(be.dev);
	if (be.failed) { err = "device stages: " + be.why; return false; }
	return ok;
}

} // namespace pansvr
