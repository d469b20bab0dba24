// seed_gpu.cu -- device-resident deBGA index and the batched seed lookup kernel (stage B).
//
// The index arrays are uploaded once and stay in HBM for the life of the service: the k-mer/offset/unipath tables in
// their on-disk layout and the bucket table in its compacted form (index.hpp: the 2 GiB of bucket
// starts reduce to the non-empty buckets plus a 4 MB directory, which also keeps the lookups in L2).  One thread per read strand runs seed_read_strand()
// (seed_core.cuh): its probes are dependent random accesses, so throughput comes from having
// every SM full of strands in flight, not from per-thread speed.  Two passes (count, exclusive
// scan, fill) give a compact MEM list without a worst-case buffer per strand.
// This service runs the small batches of the HOST path (pairs with 'N' or random_r sampling, pipeline.cpp) on a high-priority
// stream; the device path seeds its sub-blocks with FnSeed / FnSeedPlace (stages_run.hpp) against the same index arrays
// (seed_service_view), in one pass.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <string>

#include "pipeline.hpp"

namespace pansvr {

namespace {

#define SCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

__global__ void seed_fill_kernel(IndexView ix, const uint64_t *bits, const uint8_t *seed_list, const SeedJob *jobs, int n,
                                 const uint32_t *off, Mem *mems)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const SeedJob j = jobs[i];
	const int cap = (int)(off[i + 1] - off[i]);
	if (cap > 0) seed_read_strand(ix, bits + j.bits_off, j.read_len, j.is_str != 0, seed_list + j.list_off, mems + off[i], cap);
}

// number of k-mer lookups of the batch (the work unit of the seeding roofline): warp-aggregated
__global__ void seed_count_kernel_probes(IndexView ix, const uint64_t *bits, const uint8_t *seed_list, const SeedJob *jobs, int n, uint32_t *count,
                                         unsigned long long *probes)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t p = 0;
	if (i < n) {
		const SeedJob j = jobs[i];
		count[i] = (uint32_t)seed_read_strand(ix, bits + j.bits_off, j.read_len, j.is_str != 0, seed_list + j.list_off, (Mem*)nullptr, 0, &p);
	}
	for (int o = 16; o > 0; o >>= 1) p += __shfl_down_sync(0xffffffffu, p, o);
	if ((threadIdx.x & 31) == 0 && p) atomicAdd(probes, (unsigned long long)p);
}

struct Buf {
	void *p = nullptr; size_t cap = 0;
	bool reserve(size_t bytes, std::string &err)
	{
		if (bytes <= cap) return true;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		const size_t want = bytes + bytes / 4 + 256;
		SCU(cudaMalloc(&p, want));
		cap = want;
		return true;
	}
	~Buf() { if (p) cudaFree(p); }
};

template <class T> bool upload(const std::vector<T> &h, T *&d, std::string &err)
{
	d = nullptr;
	SCU(cudaMalloc((void**)&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
	SCU(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
	return true;
}

} // namespace

static int g_staging_device = 0;      // device of the (last created) seed service: helper threads allocate on it, not on device 0

void *staging_alloc(size_t bytes)
{
	void *p = nullptr;
	cudaSetDevice(g_staging_device);
	return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void staging_free(void *p) { if (p) cudaFreeHost(p); }
bool staging_is_pinned(const void *p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}

struct SeedService {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	unsigned long long *d_probes = nullptr;
	uint64_t *seqb = nullptr, *seqf = nullptr, *posp = nullptr, *bkt_start = nullptr, *off_g = nullptr;
	uint32_t *kmer_g = nullptr, *bkt_dir = nullptr, *bkt_key = nullptr;
	IndexView view;
	Buf bits, list, jobs, count, off, mems, tmp;
	size_t index_bytes = 0;
};

SeedService *seed_service_create(const DebgaIndex &idx, int device, std::string &err)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
		err = "seed service: no usable CUDA device (the seeding stage has no CPU fallback)";
		return nullptr;
	}
	cudaSetDevice(device);
	g_staging_device = device;
	SeedService *s = new SeedService();
	s->device = device;
	auto fail = [&]() -> SeedService* { seed_service_destroy(s); return nullptr; };
	// (this service only runs the host path's small batches: its stream goes ahead of the bulk kernels of the sub-blocks in flight)
	int prio_lo = 0, prio_hi = 0;
	cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
	if (cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { err = "cudaStreamCreate failed"; return fail(); }
	for (cudaEvent_t &e : s->ev) if (cudaEventCreate(&e) != cudaSuccess) { err = "cudaEventCreate failed"; return fail(); }
	if (cudaMalloc((void**)&s->d_probes, 8) != cudaSuccess) { err = "cudaMalloc failed"; return fail(); }
	if (!upload(idx.seqb, s->seqb, err) || !upload(idx.seqf, s->seqf, err) || !upload(idx.posp, s->posp, err) ||
	    !upload(idx.bkt_dir, s->bkt_dir, err) || !upload(idx.bkt_key, s->bkt_key, err) || !upload(idx.bkt_start, s->bkt_start, err) ||
	    !upload(idx.off_g, s->off_g, err) || !upload(idx.kmer_g, s->kmer_g, err)) return fail();
	s->index_bytes = (idx.seqb.size() + idx.seqf.size() + idx.posp.size() + idx.bkt_start.size() + idx.off_g.size()) * 8 +
	                 (idx.kmer_g.size() + idx.bkt_dir.size() + idx.bkt_key.size()) * 4;
	s->view.seqb = s->seqb; s->view.seqf = s->seqf; s->view.posp = s->posp; s->view.off_g = s->off_g;
	s->view.bkt_dir = s->bkt_dir; s->view.bkt_key = s->bkt_key; s->view.bkt_start = s->bkt_start;
	s->view.kmer_g = s->kmer_g; s->view.n_seqf = idx.seqf.size();
	return s;
}

const IndexView &seed_service_view(const SeedService *s) { return s->view; }

void seed_service_destroy(SeedService *s)
{
	if (!s) return;
	cudaSetDevice(s->device);
	for (void *p : {(void*)s->seqb, (void*)s->seqf, (void*)s->posp, (void*)s->bkt_dir, (void*)s->bkt_key, (void*)s->bkt_start, (void*)s->off_g, (void*)s->kmer_g}) if (p) cudaFree(p);
	for (cudaEvent_t e : s->ev) if (e) cudaEventDestroy(e);
	if (s->d_probes) cudaFree(s->d_probes);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
}

bool seed_service_run(SeedService *s, SeedBatch &b, std::string &err)
{
	const int n = (int)b.jobs.size();
	b.mem_off.assign((size_t)n + 1, 0);
	b.mems.clear();
	if (n == 0) return true;
	SCU(cudaSetDevice(s->device));
	cudaStream_t st = s->stream;
	if (!s->bits.reserve(b.bits.size() * 8 + 8, err) || !s->list.reserve(b.seed_list.size() + 8, err) ||
	    !s->jobs.reserve((size_t)n * sizeof(SeedJob), err) || !s->count.reserve(((size_t)n + 1) * 4, err) ||
	    !s->off.reserve(((size_t)n + 1) * 4, err)) return false;
	SCU(cudaMemcpyAsync(s->bits.p, b.bits.data(), b.bits.size() * 8, cudaMemcpyHostToDevice, st));
	if (!b.seed_list.empty()) SCU(cudaMemcpyAsync(s->list.p, b.seed_list.data(), b.seed_list.size(), cudaMemcpyHostToDevice, st));
	SCU(cudaMemcpyAsync(s->jobs.p, b.jobs.data(), (size_t)n * sizeof(SeedJob), cudaMemcpyHostToDevice, st));
	SCU(cudaMemsetAsync(s->count.p, 0, ((size_t)n + 1) * 4, st));
	SCU(cudaMemsetAsync(s->d_probes, 0, 8, st));
	const int threads = 128, blocks = (n + threads - 1) / threads;
	SCU(cudaEventRecord(s->ev[0], st));
	seed_count_kernel_probes<<<blocks, threads, 0, st>>>(s->view, (const uint64_t*)s->bits.p, (const uint8_t*)s->list.p, (const SeedJob*)s->jobs.p, n,
	                                                     (uint32_t*)s->count.p, s->d_probes);
	SCU(cudaGetLastError());
	SCU(cudaEventRecord(s->ev[1], st));
	unsigned long long probes = 0;
	SCU(cudaMemcpyAsync(&probes, s->d_probes, 8, cudaMemcpyDeviceToHost, st));
	b.dev.launches += 1;
	b.dev.h2d_bytes += (int64_t)(b.bits.size() * 8 + b.seed_list.size() + (size_t)n * sizeof(SeedJob));
	size_t tmp_bytes = 0;
	SCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint32_t*)s->count.p, (uint32_t*)s->off.p, n + 1, st));
	if (!s->tmp.reserve(tmp_bytes, err)) return false;
	SCU(cub::DeviceScan::ExclusiveSum(s->tmp.p, tmp_bytes, (const uint32_t*)s->count.p, (uint32_t*)s->off.p, n + 1, st));
	SCU(cudaMemcpyAsync(b.mem_off.data(), s->off.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
	SCU(cudaStreamSynchronize(st));
	{
		float ms = 0;
		SCU(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
		b.dev.seed_kernel_ms += ms; b.dev.seed_probes += (int64_t)probes; b.dev.d2h_bytes += (int64_t)((size_t)n + 1) * 4 + 8;
	}
	const size_t total = b.mem_off[n];
	b.mems.resize(total);
	if (total == 0) return true;
	if (!s->mems.reserve(total * sizeof(Mem), err)) return false;
	SCU(cudaEventRecord(s->ev[2], st));
	seed_fill_kernel<<<blocks, threads, 0, st>>>(s->view, (const uint64_t*)s->bits.p, (const uint8_t*)s->list.p, (const SeedJob*)s->jobs.p, n,
	                                             (const uint32_t*)s->off.p, (Mem*)s->mems.p);
	SCU(cudaGetLastError());
	SCU(cudaEventRecord(s->ev[3], st));
	SCU(cudaMemcpyAsync(b.mems.data(), s->mems.p, total * sizeof(Mem), cudaMemcpyDeviceToHost, st));
	SCU(cudaStreamSynchronize(st));
	{
		float ms = 0;
		SCU(cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]));
		b.dev.seed_kernel_ms += ms; b.dev.launches += 1; b.dev.d2h_bytes += (int64_t)(total * sizeof(Mem));
	}
	return true;
}

} // namespace pansvr
