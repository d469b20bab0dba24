// pipeline.hpp -- the `panSVR fc_aln` block pipeline, restructured for a batched GPU back end.
//
// Reference: src/PanSVgenerateVCF/read_realignment.{cpp,hpp} (fc_aln), src/cpp_lib/graph.{cpp,hpp}
// (chaining), src/PanSVgenerateVCF/deBGA_index.cpp (merge/expand of seeds).  The reference handles
// one pair at a time on a CPU thread; here one block of pairs goes through stages so that the two
// hot steps run as device batches:
//
//   A  host   2-bit encode, STR k-mer census, 64-bit packing                    (RR:646-654, 538-598, 295-300)
//   B  GPU    seed lookup + MEM extension for every read strand                 (seed_core.cuh)
//   C  host   merge MEMs per unipath, expand to reference positions, chain DP   (deBGA_index.cpp:151-305, graph.cpp:53-150)
//   D  host   plan the ksw windows of every chain end that can still be chosen  (RR:308-400, 910-986)
//   E  GPU    ksw_extd2 batch                                                   (ksw_team.cuh)
//   F  host   chain selection, result sort, pairing, SAM text                   (RR:212-293, 406-476, 479-536, 745-799; RRH:434-628)
//
// Everything runs on the helper threads except what has to see the reference's libc random streams in the reference's
// order.  rand(): stage D plans a superset because which chain ends get extended depends on rand() tie-breaks whose stream
// position depends on the results of earlier pairs (get_ksw_score is a pure function of (strand, end node), so every node
// with dist >= max(30, best-30) is planned); stage F finishes every pair against a probe first, enumerates the outcomes of
// the ties of a read (usually they all give the same candidates, and the read then only advances the stream), records the
// pairing ties as events, and a short in-order pass redraws those from the real stream; reads with 'N' (rand()%4 per base)
// are prepared once per possible substitution.  random_r(): reads that sample a repeat's positions are chained one after
// the other in stage C.  DESIGN.md section 4.5 has the details.
//
// The output is defined against `panSVR fc_aln -t 1` (the only deterministic mode).
#pragma once
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "index.hpp"
#include "refrand.hpp"
#include "seed_core.cuh"

namespace pansvr {

struct AlnOptions {                   // MAP_PARA, read_realignment.hpp:43-128 (defaults :33-41)
	int match = 2, mismatch = 12, gap_open = 16, gap_ex = 1, gap_open2 = 32, gap_ex2 = 0;
	int zdrop = 400, bw = 500;        // -w is parsed and ignored by the reference (RR:817-827): ksw runs with w=200
	bool not_ori = false;             // -Q
	int max_use_read = 0x7fffffff;
	int threads = 1;                  // host helper threads of the parallel stages (the output does not depend on it)
	// read statistics (STAT_ field of the first comment, RR:134-148)
	bool stat_set = false;
	int read_len = 150, isize_min = 100, isize_mid = 500, isize_max = 900;
};

struct FastqRec {                     // views into the caller's FASTQ text (no copies)
	const char *name = nullptr, *comment = nullptr, *seq = nullptr, *qual = nullptr;
	uint32_t name_l = 0, comment_l = 0, seq_l = 0, qual_l = 0;
};

// ---- device services (link-time: CUDA in the product library, host emulation in tests/emul) ---------------------
// Staging memory of the two device batches: page-locked in the product (cudaHostAlloc in seed_gpu.cu), so that the
// copies of a block are direct DMA; plain malloc in the host test build.
void *staging_alloc(size_t bytes);
void staging_free(void *p);
bool staging_is_pinned(const void *p);   // page-locked memory the device can copy from directly (a caller's buffer from pansvr_host_alloc)

// The part of std::vector the pipeline needs, for trivially copyable T, on staging memory; grow-only, contents are
// not initialised by resize().  The pipeline keeps its batch buffers across blocks, so they are pinned once.
template <class T> class HostVec {
public:
	HostVec() {}
	HostVec(const HostVec&) = delete;
	HostVec &operator=(const HostVec&) = delete;
	~HostVec() { if (p_) staging_free(p_); }
	T *data() { return p_; }
	const T *data() const { return p_; }
	size_t size() const { return n_; }
	bool empty() const { return n_ == 0; }
	void clear() { n_ = 0; }
	T &operator[](size_t i) { return p_[i]; }
	const T &operator[](size_t i) const { return p_[i]; }
	void reserve(size_t m)
	{
		if (m <= cap_) return;
		const size_t want = m + m / 2 + 64;
		T *q = (T*)staging_alloc(want * sizeof(T));
		if (!q) throw std::bad_alloc();
		if (n_) memcpy(q, p_, n_ * sizeof(T));
		if (p_) staging_free(p_);
		p_ = q; cap_ = want;
	}
	void resize(size_t m) { reserve(m); n_ = m; }
	void assign(size_t m, const T &v) { resize(m); for (size_t i = 0; i < m; ++i) p_[i] = v; }
	void push_back(const T &v) { reserve(n_ + 1); p_[n_++] = v; }
	void append(const T *b, const T *e) { const size_t k = (size_t)(e - b); reserve(n_ + k); if (k) memcpy(p_ + n_, b, k * sizeof(T)); n_ += k; }
private:
	T *p_ = nullptr;
	size_t n_ = 0, cap_ = 0;
};

struct DevCounters {                  // what the device services did for a block (bench.py: gpu_launches, e2e bytes, rooflines)
	int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
	int64_t seed_probes = 0;          // k-mer lookups of the seeding kernels (every probe is made once: the counting pass keeps what it finds)
	double seed_kernel_ms = 0, ksw_kernel_ms = 0, stage_kernel_ms = 0;   // CUDA-event time of our kernels on their streams
	double by_stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // the same by group: 0 records / original alignments / encode + census, 1 seeding,
	                                                    // 2 merge + chain, 3 ksw planning, 4 candidate resolution, 5 cell count, 6 pairing probe + finalize, 7 SAM text
	void add(const DevCounters &o)
	{
		launches += o.launches; h2d_bytes += o.h2d_bytes; d2h_bytes += o.d2h_bytes; seed_probes += o.seed_probes;
		seed_kernel_ms += o.seed_kernel_ms; ksw_kernel_ms += o.ksw_kernel_ms; stage_kernel_ms += o.stage_kernel_ms;
		for (int i = 0; i < 8; ++i) by_stage_ms[i] += o.by_stage_ms[i];
	}
};

struct SeedJob {                      // one read strand
	uint32_t bits_off, read_len;      // word offset into SeedBatch::bits
	uint32_t list_off;                // offset into SeedBatch::seed_list (STR reads only)
	uint32_t is_str;
};
struct SeedBatch {
	HostVec<uint64_t> bits;           // packed reads, 32 bases per word, one spare zero word after each read
	std::vector<uint8_t> seed_list;
	HostVec<SeedJob> jobs;
	HostVec<Mem> mems;                // out: MEMs of job i are mems[mem_off[i] .. mem_off[i+1])
	HostVec<uint32_t> mem_off;
	DevCounters dev;                  // out: added to by seed_service_run
	void clear() { bits.clear(); seed_list.clear(); jobs.clear(); mems.clear(); mem_off.clear(); }
};
struct KswBatchBuf {                  // the ksw tasks of one block, joined (stage D -> E -> F)
	HostVec<uint8_t> q, t;
	HostVec<int64_t> qoff, toff;
	HostVec<int32_t> qlen, tlen, res;
	HostVec<uint32_t> cig;
	int cap = 16;                     // CIGAR words per task; a block whose longest CIGAR does not fit is run again with room
};
struct StageService;                  // opaque; the device stages (stages_run.hpp)
struct BlockHooks;
struct SeedService;                   // opaque; owns the device-resident index
SeedService *seed_service_create(const DebgaIndex &idx, int device, std::string &err);
void seed_service_destroy(SeedService *s);
bool seed_service_run(SeedService *s, SeedBatch &b, std::string &err);

// ---- results of one pair --------------------------------------------------------------------------------------------
struct BlockOutput {                  // record text of one block: buffer t holds the lines ("...\n") of the t-th chunk of pairs,
	std::vector<std::string> sam;     // concatenating the buffers in order gives the block's output in input order
	std::vector<std::string> ori;     // `-p` output: pairs still poorly aligned
	HostVec<char> sam_text;           // device path: the block's main output as one text, straight from the device (then `sam` is empty)
	// ... or straight into the caller's buffer: place(total) is asked once per block for room for `total` bytes (nullptr = none: the
	// text goes to sam_text); placed_bytes = what went there.  A block that has no text for it calls place(0) all the same, so
	// that callers can hand the room out block by block, in order.
	std::function<char*(size_t)> place;
	size_t placed_bytes = 0; const char *placed_ptr = nullptr; bool placed = false, place_called = false;
};

// PANSVR_TRACE=<file>: a mark in the process's phase trace (<file>.<pid>.host: "H <seq> <what> <steady clock seconds>")
void trace_mark(uint64_t seq, const char *what);

struct CigarPath { uint8_t type; int16_t size; };

class AlnPipeline {
public:
	// `stages` (may be null): the device stages A..F1 (stages_run.hpp); without them every stage but seeding and ksw runs on the host
	AlnPipeline(const DebgaIndex &idx, const AlnOptions &opt, SeedService *seeds, void *ksw_ctx, StageService *stages = nullptr, int device = 0);
	~AlnPipeline();
	AlnPipeline(const AlnPipeline&) = delete;
	AlnPipeline &operator=(const AlnPipeline&) = delete;
	// static-chunk parallel loop over [0,n) on the pipeline's helper threads: fn(begin, end, chunk_index)
	void parallel(size_t n, const std::function<void(size_t, size_t, int)> &fn, size_t serial_below = 256);
	// Aligns n_reads/2 interleaved pairs (recs[2i], recs[2i+1]); `out` receives the SAM text in input order.
	// Blocks are numbered by the caller (`seq` = 0, 1, 2, ... since the last reset) and two blocks with consecutive numbers
	// may be in flight on two threads: everything that consumes a random stream is serialised in `seq` order inside, the
	// rest (stages A-E, the probe, the record text) of one block overlaps the in-order replay of the other.
	bool align_block(const FastqRec *recs, size_t n_reads, BlockOutput &out, std::string &err, uint64_t seq);
	bool align_block_host(const FastqRec *recs, size_t n_reads, BlockOutput &out, std::string &err, uint64_t seq, BlockHooks *hooks);
	// the same block given as text (strict 4-line FASTQ of n_pairs interleaved pairs): the records are found on the device
	bool align_block_text(const char *text, size_t bytes, size_t n_pairs, BlockOutput &out, std::string &err, uint64_t seq, bool *reparse);
	bool align_block_dev(const char *text, size_t bytes, const FastqRec *recs, size_t n_pairs, BlockOutput &out, std::string &err, uint64_t seq, bool *reparse);
	bool has_device_stages() const { return stages_ != nullptr; }
	uint64_t next_seq() { return seq_issued_++; }
	std::atomic<uint64_t> bad_cigar_records_{0};   // records left out because their CIGAR does not span the read (see output_bam)
	void ensure_read_stats(const FastqRec &first);   // STAT_ fields of the input's first comment; call before overlapping blocks
	void reset();                     // back to the state of a freshly started `fc_aln` (rand() streams, counters)
	// ---- one input sharded over several processes: the only state that flows from pair to pair are the libc
	// random streams the replay consumes in input order.  A process that handles pairs [b, e) of the input takes the streams as the
	// process before it left them and hands them on when its own in-order passes are done; everything else runs without waiting.
	struct StreamState { uint32_t magic; GlibcRandom rand, rand_r[2]; };
	StreamState export_streams();                       // waits for every block issued so far to finish its in-order pass
	void import_streams(const StreamState &s);           // before any block since the last reset
	void await_streams(const std::string &path);         // the first in-order section from now on waits for `path` and imports it
	bool publish_streams(const std::string &path);       // export_streams() into `path` (written beside it and renamed: readers never see a part)
	bool awaiting_streams() { std::lock_guard<std::mutex> lk(turn_m_); return !await_path_.empty(); }
	// the same per block: the in-order section of block `seq` first waits for the file `await` and imports it, and writes `publish`
	// the moment it is done (before the block's later stages) -- pieces of one input dealt to several processes in turn
	void chain_at(uint64_t seq, const char *await, const char *publish);
	// a block that ends without having had its in-order section (an error on its way) must still let the blocks behind it have theirs
	void pass_turn_if_pending(uint64_t seq);
	struct Stats { uint64_t reads = 0, probes_reads = 0, mems = 0, ksw_tasks = 0, ksw_cells = 0, deferred_pairs = 0, in_order_pairs = 0, in_order_draws = 0, host_pairs = 0, tie_pairs = 0;
	               double t_in_order = 0;
	               double t_stage[8] = {0, 0, 0, 0, 0, 0, 0, 0}; DevCounters dev; } stats;   // A..F, FASTQ parse, output assembly
	AlnOptions opt;                   // stat_set / read_len / isize_* are filled from the first comment
private:
	struct Impl;
	struct Workers;
	Workers *workers_ = nullptr;
	const DebgaIndex &idx_;
	SeedService *seeds_;
	void *ksw_;
	StageService *stages_;
	int device_ = 0;
	// one block's trip through the device stages: a stage service instance (device buffers, stream, ksw context) plus the host
	// side of its transfers; blocks in flight at the same time hold different ones, instances are created on demand and kept
	struct DevBuffers;
	std::vector<DevBuffers*> dev_free_, dev_all_;
	std::mutex dev_pool_m_;
	DevBuffers *acquire_dev(std::string &err, uint64_t seq = 0);
	// further GPUs of the same box: sub-block k goes to site k mod (1 + extra sites); each site has its own index replica
	struct DevSite { int device; SeedService *seeds; };
	// how many sub-blocks may be in their first trip on one device at a time: the trips of the sub-blocks in flight share the GPU,
	// so the fewer run side by side the sooner the oldest one is through -- and it is the oldest one every in-order pass (of this
	// process, and of the process that has the next piece of the input) waits for.  Second trips are not counted.
	std::mutex trip1_m_; std::condition_variable trip1_cv_; std::vector<int> trip1_busy_; std::vector<std::set<uint64_t>> trip1_wait_; int trip1_cap_ = 3;
	std::vector<DevSite> sites_;          // [0] = the context's first device
public:
	void add_device(int device, SeedService *seeds) { sites_.push_back(DevSite{device, seeds}); }
	size_t n_devices() const { return sites_.size(); }
private:
	void release_dev(DevBuffers *d);
	// batch buffers of the host path live across blocks (staging memory is pinned once); every block in flight holds one set
	struct HostSlot { SeedBatch seeds; KswBatchBuf ksw; };
	std::vector<HostSlot*> host_free_, host_all_;
	HostSlot *acquire_host();
	void release_host(HostSlot *h);
	SeedBatch seed_small_;
	std::mutex dev_m_;                    // the two device services take one batch at a time
	std::mutex stats_m_;
	std::mutex turn_m_;                   // blocks take their in-order sections by sequence number
	std::condition_variable turn_cv_;
	uint64_t replay_turn_ = 0, seq_issued_ = 0;
	GlibcRandom rand_;                // the process-global rand() of the reference
	GlibcRandom rand_r_[2];           // per-handler random_r states (RRH:339-340), seeded from rand_ at start-up
	int min_filter_score_ = 0;
	std::string await_path_;              // guarded by turn_m_
	std::vector<uint8_t> host_sv_;        // per-anchor (chr_id, st_pos, end_offset) as the device stages see them (DevSv), for the pairs finished in order on the host
	std::map<uint64_t, std::pair<std::string, std::string>> chain_;   // seq -> (await, publish); guarded by turn_m_
	bool write_streams(const std::string &path, const StreamState &s);
	friend struct Impl;
};

} // namespace pansvr
