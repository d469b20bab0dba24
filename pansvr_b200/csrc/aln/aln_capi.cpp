// aln_capi.cpp -- C ABI and command line of the aln stage (declared in include/pansvr_b200.h).
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/pansvr_b200.h"
#include "bam_out.hpp"
#include "text_in.hpp"
#include "pipeline.hpp"
#include "stages_run.hpp"

using namespace pansvr;

namespace { thread_local std::string g_aln_err; }

// Output buffers the device writes the SAM text into directly (page-locked staging memory, so the copy from the device is one
// DMA per sub-block and nothing is copied on the host afterwards).  They are handed to the caller as the result of
// pansvr_aln_block and come back through pansvr_free; a few are kept for reuse, since page-locking is not cheap.
namespace {
struct OutPool {
	struct Buf { char *p; size_t cap; bool busy; };
	std::mutex m;
	std::vector<Buf> bufs;
	char *acquire(size_t cap)
	{
		std::lock_guard<std::mutex> lk(m);
		for (Buf &b : bufs) if (!b.busy && b.cap >= cap) { b.busy = true; return b.p; }
		for (size_t i = 0; i < bufs.size(); ++i) if (!bufs[i].busy) { staging_free(bufs[i].p); bufs.erase(bufs.begin() + (long)i); break; }   // one too small: replace it
		char *p = (char*)staging_alloc(cap);
		if (!p) return nullptr;
		bufs.push_back(Buf{p, cap, true});
		return p;
	}
	bool release(void *p)                                     // false: not one of ours
	{
		std::lock_guard<std::mutex> lk(m);
		size_t idle = 0;
		for (Buf &b : bufs) idle += !b.busy;
		for (size_t i = 0; i < bufs.size(); ++i) if (bufs[i].p == p) {
			if (idle >= 3) { staging_free(bufs[i].p); bufs.erase(bufs.begin() + (long)i); } else bufs[i].busy = false;
			return true;
		}
		return false;
	}
};
OutPool g_out_pool;
}

extern "C" int pansvr_ksw_create_prio(int device, int high_priority, pansvr_ksw_ctx **out);   // ksw_batch.cu (not in the public header)

struct pansvr_aln_ctx {
	DebgaIndex idx;
	AlnOptions opt;
	SeedService *seeds = nullptr;
	std::vector<SeedService*> more_seeds;   // index replicas on the further GPUs of a device list
	StageService *stages = nullptr;
	pansvr_ksw_ctx *ksw = nullptr;
	AlnPipeline *pipe = nullptr;
	BamHeaderInfo bam_hdr;
	std::deque<BlockOutput> outs;     // chunk buffers of the record text of every sub-block, kept across calls
};

struct pansvr_bam_file {
	FILE *f = nullptr;
	BamWriter *w = nullptr;
	~pansvr_bam_file() { delete w; if (f) fclose(f); }
};

namespace {

// 4-line FASTQ records out of a memory buffer: "@name comment\nseq\n+...\nqual\n" (kseq_read, clib/utils.c:953-990).
// Records are views into the buffer.
void parse_fastq(const char *p, size_t n, std::vector<FastqRec> &out)
{
	size_t i = 0;
	auto line = [&](const char *&b, uint32_t &l) -> bool {
		if (i >= n) return false;
		const char *nl = (const char*)memchr(p + i, '\n', n - i);
		const size_t e = nl ? (size_t)(nl - p) : n;
		size_t len = e - i;
		// (a '\r' before the newline stays part of the line: the reference's reader keeps it, clib/utils.c:953-990)
		b = p + i; l = (uint32_t)len;
		i = e + 1;
		return true;
	};
	out.reserve(n / 300 + 16);
	for (;;) {
		FastqRec r;
		const char *h; uint32_t hl;
		do { if (!line(h, hl)) return; } while (hl == 0);
		if (h[0] != '@' && h[0] != '>') return;
		uint32_t sp = 1;
		while (sp < hl && h[sp] != ' ' && h[sp] != '\t') ++sp;
		r.name = h + 1; r.name_l = sp - 1;
		while (sp < hl && (h[sp] == ' ' || h[sp] == '\t')) ++sp;
		r.comment = h + sp; r.comment_l = hl - sp;
		const char *plus; uint32_t pl;
		if (!line(r.seq, r.seq_l) || !line(plus, pl) || !line(r.qual, r.qual_l)) return;
		out.push_back(r);
	}
}

// The same records, found on all helper threads: strict 4-line FASTQ (what fc_signal writes) has its record starts at the
// lines whose index is a multiple of four, so every thread counts the newlines of its slice of the buffer, a prefix sum
// gives the index of the first line starting in each slice, and each thread parses the records that start in its slice
// straight into their final slots.  Anything that does not look like strict 4-line FASTQ (blank lines, a header not
// starting with '@', a third line not starting with '+') returns false and the sequential parser above is used.
bool parse_fastq_parallel(const char *p, size_t n, AlnPipeline &pipe, int threads, std::vector<FastqRec> &out)
{
	if (threads <= 1 || n < (1u << 20)) return false;
	const size_t T = (size_t)threads, per = (n + T - 1) / T;
	std::vector<size_t> nl(T + 1, 0);
	pipe.parallel(T, [&](size_t b, size_t e, int) {
		for (size_t t = b; t < e; ++t) {
			const size_t s0 = std::min(n, per * t), s1 = std::min(n, per * (t + 1));
			size_t c = 0;
			for (const char *q = p + s0, *qe = p + s1; (q = (const char*)memchr(q, '\n', (size_t)(qe - q))) != nullptr; ++q) ++c;
			nl[t + 1] = c;
		}
	}, 2);
	for (size_t t = 0; t < T; ++t) nl[t + 1] += nl[t];
	const size_t lines = nl[T] + (n && p[n - 1] != '\n' ? 1 : 0);
	if (lines % 4 != 0) return false;
	out.assign(lines / 4, FastqRec());
	std::vector<uint8_t> bad(T, 0);
	pipe.parallel(T, [&](size_t b, size_t e, int) {
		for (size_t t = b; t < e; ++t) {
			const size_t s0 = std::min(n, per * t), s1 = std::min(n, per * (t + 1));
			if (s0 >= s1) continue;
			size_t q = s0, idx = nl[t];                           // idx = newlines before q = index of the line containing q
			if (q != 0 && p[q - 1] != '\n') {                      // q is inside a line that started earlier: go to the next line start
				const char *x = (const char*)memchr(p + q, '\n', s1 - q);
				if (!x) continue;                                  // no line starts in this slice
				q = (size_t)(x - p) + 1; ++idx;
			}
			while (idx % 4 != 0 && q < s1) {                       // the tail of a record that started in an earlier slice
				const char *x = (const char*)memchr(p + q, '\n', n - q);
				if (!x) { q = n; break; }
				q = (size_t)(x - p) + 1; ++idx;
			}
			while (q < s1 && q < n) {                              // records starting in [s0, s1)
				const char *ln[4]; uint32_t ll[4];
				for (int k = 0; k < 4; ++k) {
					const char *x = q < n ? (const char*)memchr(p + q, '\n', n - q) : nullptr;
					const size_t e2 = x ? (size_t)(x - p) : n;
					size_t len = e2 - q;
					// (a '\r' stays part of the line, like in parse_fastq)
					ln[k] = p + q; ll[k] = (uint32_t)len;
					q = e2 + 1;
				}
				if (ll[0] == 0 || (ln[0][0] != '@' && ln[0][0] != '>') || ll[2] == 0 || ln[2][0] != '+') { bad[t] = 1; break; }
				FastqRec r;
				uint32_t sp = 1;
				while (sp < ll[0] && ln[0][sp] != ' ' && ln[0][sp] != '\t') ++sp;
				r.name = ln[0] + 1; r.name_l = sp - 1;
				while (sp < ll[0] && (ln[0][sp] == ' ' || ln[0][sp] == '\t')) ++sp;
				r.comment = ln[0] + sp; r.comment_l = ll[0] - sp;
				r.seq = ln[1]; r.seq_l = ll[1]; r.qual = ln[3]; r.qual_l = ll[3];
				out[idx / 4] = r;
				idx += 4;
			}
		}
	}, 2);
	for (uint8_t x : bad) if (x) { out.clear(); return false; }
	return true;
}

// bounded hand-over between the three steps of the command line
template <class T> class Queue {
public:
	explicit Queue(size_t depth) : depth_(depth) {}
	void push(T &&v)
	{
		std::unique_lock<std::mutex> lk(m_);
		room_.wait(lk, [&]() { return q_.size() < depth_; });
		q_.push_back(std::move(v));
		item_.notify_one();
	}
	T pop()
	{
		std::unique_lock<std::mutex> lk(m_);
		item_.wait(lk, [&]() { return !q_.empty(); });
		T v = std::move(q_.front());
		q_.pop_front();
		room_.notify_one();
		return v;
	}
private:
	size_t depth_;
	std::deque<T> q_;
	std::mutex m_;
	std::condition_variable room_, item_;
};

AlnOptions from_c(const pansvr_aln_options_t *o)
{
	AlnOptions a;
	a.threads = 0;
	if (!o) return a;
	// a zero means "the reference's default" unless the field's bit in explicit_mask says the caller meant zero (`-E 0` on the
	// command line is honoured by the reference's option parser)
	auto given = [&](int bit, int32_t v) { return v != 0 || ((o->explicit_mask >> bit) & 1); };
	if (given(0, o->match)) a.match = o->match;
	if (given(1, o->mismatch)) a.mismatch = o->mismatch;
	if (given(2, o->gap_open)) a.gap_open = o->gap_open;
	if (given(3, o->gap_ex)) a.gap_ex = o->gap_ex;
	if (given(4, o->gap_open2)) a.gap_open2 = o->gap_open2;
	a.gap_ex2 = o->gap_ex2;
	if (given(6, o->zdrop)) a.zdrop = o->zdrop;
	if (given(7, o->band_width)) a.bw = o->band_width;
	a.not_ori = o->not_ori != 0;
	if (o->max_use_read > 0 || ((o->explicit_mask >> 9) & 1)) a.max_use_read = o->max_use_read;
	a.threads = o->threads;
	return a;
}

} // namespace

extern "C" {

int pansvr_aln_create(const char *index_dir, const char *header_sam, const pansvr_aln_options_t *opt, int device, pansvr_aln_ctx **out)
{
	if (!index_dir || !header_sam || !out) return PANSVR_E_ARG;
	*out = nullptr;
	pansvr_aln_ctx *c = new pansvr_aln_ctx();
	c->opt = from_c(opt);
	if (c->opt.threads <= 0) c->opt.threads = (int)std::min(48u, std::max(1u, std::thread::hardware_concurrency()));   // the reference's -t limit (RRH:121)
	std::string err;
	const auto tick = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const bool timing = getenv("PANSVR_TIMING") != nullptr;
	double t0 = tick();
	auto lap = [&](const char *what) { if (timing) { const double t = tick(); fprintf(stderr, "[timing] create/%s %.3f s\n", what, t - t0); t0 = t; } };
	// the index files are read (and the bucket table compacted) while the CUDA runtime starts up
	bool idx_ok = false;
	std::string idx_err;
	std::thread loader([&]() { idx_ok = c->idx.load(index_dir, header_sam, idx_err); });
	const int ksw_rc = pansvr_ksw_create_prio(device, 1, &c->ksw);     // the host path's context: its few tasks go ahead of the bulk kernels
	lap("ksw context (CUDA init)");
	loader.join();
	lap("index files (remainder)");
	if (!idx_ok) { g_aln_err = idx_err; if (ksw_rc == 0) pansvr_ksw_destroy(c->ksw); delete c; return PANSVR_E_ARG; }
	if (ksw_rc != 0) { g_aln_err = pansvr_last_error(); delete c; return PANSVR_E_CUDA; }
	c->seeds = seed_service_create(c->idx, device, err);
	if (!c->seeds) { g_aln_err = err; pansvr_ksw_destroy(c->ksw); delete c; return PANSVR_E_CUDA; }
	c->stages = stage_service_create(c->idx, c->seeds, c->ksw, device, err);
	if (!c->stages) { g_aln_err = err; seed_service_destroy(c->seeds); pansvr_ksw_destroy(c->ksw); delete c; return PANSVR_E_CUDA; }
	lap("index upload");
	c->pipe = new AlnPipeline(c->idx, c->opt, c->seeds, c->ksw, c->stages, device);
	c->bam_hdr.parse(c->idx.header_text);
	lap("pipeline");
	*out = c;
	return 0;
}

int pansvr_aln_create_multi(const char *index_dir, const char *header_sam, const pansvr_aln_options_t *opt, const int *devices, int n_devices, pansvr_aln_ctx **out)
{
	if (!devices || n_devices < 1) return PANSVR_E_ARG;
	const int rc = pansvr_aln_create(index_dir, header_sam, opt, devices[0], out);
	if (rc != 0) return rc;
	pansvr_aln_ctx *c = *out;
	for (int k = 1; k < n_devices; ++k) {                          // an index replica per further GPU; stage services are made on demand
		std::string err;
		SeedService *s = seed_service_create(c->idx, devices[k], err);
		if (!s) { g_aln_err = err; pansvr_aln_destroy(c); *out = nullptr; return PANSVR_E_CUDA; }
		c->more_seeds.push_back(s);
		c->pipe->add_device(devices[k], s);
	}
	return 0;
}

void pansvr_aln_destroy(pansvr_aln_ctx *c)
{
	if (!c) return;
	delete c->pipe;
	for (SeedService *s : c->more_seeds) seed_service_destroy(s);
	stage_service_destroy(c->stages);
	seed_service_destroy(c->seeds);
	pansvr_ksw_destroy(c->ksw);
	delete c;
}

const char *pansvr_aln_header_text(const pansvr_aln_ctx *c) { return c ? c->idx.header_text.c_str() : ""; }
const char *pansvr_aln_last_error(void) { return g_aln_err.c_str(); }

namespace {

// parse + align one block; `outp` receives the record text of every pair
// parse + align one block; `text` receives, in input order, the buffers that hold the record text.
// A large block is cut into sub-blocks and two of them are in flight at a time: the single-threaded in-order replay of one
// overlaps the parallel stages of the next (AlnPipeline::align_block keeps the random streams in sequence).
struct Part { const char *p; size_t n; };                    // a piece of record text; pieces end at line ends

// cuts [p, p + n) into pieces of about 4 MB at line ends, so that whoever consumes the text can do so on all helper threads
void add_parts(std::vector<Part> &parts, const char *p, size_t n)
{
	const size_t want = (size_t)4 << 20;
	size_t at = 0;
	while (at < n) {
		size_t e = std::min(n, at + want);
		if (e < n) { const char *nl = (const char*)memchr(p + e, '\n', n - e); e = nl ? (size_t)(nl - p) + 1 : n; }
		parts.push_back(Part{p + at, e - at});
		at = e;
	}
}

// One piece of an input handed to a block call: whole pairs as 4-line FASTQ text.  await / publish: files through which the
// reference's random streams are handed from the piece before it (realigned by another process) and to the piece after it.
struct Piece { const char *p = nullptr; size_t n = 0; const char *await_path = nullptr, *publish_path = nullptr; size_t sam_bytes = 0, ori_bytes = 0; };

size_t count_line_ends(const char *b, const char *e)
{
	size_t c = 0;
	for (const char *q = b; (q = (const char*)memchr(q, '\n', (size_t)(e - q))) != nullptr; ++q) ++c;
	return c;
}

// Position just behind line end number `want` counted from `at`, found on all helper threads by counting the line ends of slices of
// a window that probably holds it (`guess` bytes); lines = `want`, or how many line ends there were if the text ends first (then n
// is returned).  Strict 4-line FASTQ (what fc_signal writes) has 8 lines per pair, so this is where a run of pairs ends without
// looking at the records.
size_t line_end_after(const char *p, size_t n, size_t at, size_t want, size_t guess, AlnPipeline &pipe, int threads, size_t &lines)
{
	lines = 0;
	const size_t T = (size_t)std::max(1, threads);
	std::vector<size_t> nl(T + 1);
	while (at < n) {
		const size_t win = std::min(n - at, std::max<size_t>(guess, (size_t)1 << 16)), slice = (win + T - 1) / T;
		std::fill(nl.begin(), nl.end(), 0);
		pipe.parallel(T, [&](size_t b, size_t e, int) {
			for (size_t t = b; t < e; ++t) nl[t + 1] = count_line_ends(p + at + std::min(win, slice * t), p + at + std::min(win, slice * (t + 1)));
		}, 2);
		for (size_t t = 0; t < T; ++t) nl[t + 1] += nl[t];
		if (lines + nl[T] < want) {                                // not in this window: go on behind it
			lines += nl[T]; at += win;
			guess = std::max<size_t>(guess / 8, (size_t)1 << 20);
			continue;
		}
		const size_t need = want - lines;
		size_t t = 0;
		while (nl[t + 1] < need) ++t;
		size_t seen = nl[t];
		const char *q = p + at + std::min(win, slice * t);
		for (;;) { q = (const char*)memchr(q, '\n', (size_t)(p + n - q)); ++seen; ++q; if (seen == need) break; }
		lines = want;
		return (size_t)(q - p);
	}
	return n;
}

size_t pairs_per_sub_block(size_t n_pairs, int threads)
{
	size_t per = std::max<size_t>(n_pairs, 1);
	if (threads > 1 && n_pairs >= 65536) per = std::min<size_t>(262144, std::max<size_t>(32768, (n_pairs + 3) / 4));
	if (const char *e = getenv("PANSVR_SUB_PAIRS")) { const long v = atol(e); if (v > 0) per = (size_t)v; }   // tests: force the cut
	return per;
}

// `room`: a buffer the sub-blocks' SAM text goes into directly, one behind the other in input order (device path); *room_used =
// how much of it was filled.  A sub-block whose text does not fit keeps it in its own buffer (and so do all later ones).
struct TextRoom { char *p = nullptr; size_t cap = 0, used = 0; size_t next = 0; bool full = false; std::mutex m; std::condition_variable cv; };

// The pieces of a call, cut into sub-blocks that go through the pipeline a few at a time.  On the device path the text goes up as
// it is and the records are found there, so the host only has to know where to cut: sub-block after sub-block, just before each
// one is started (the first one is on its way after a look at a few MB, not after a pass over the whole text).
int run_block(pansvr_aln_ctx *c, Piece *pieces, size_t n_pieces, std::vector<Part> &sam, std::vector<Part> &ori,
              const std::function<void(const BlockOutput&, size_t)> *on_done = nullptr, TextRoom *room = nullptr)
{
	const auto tick = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const int threads = c->opt.threads;
	struct Sub {
		size_t piece = 0; const char *p = nullptr; size_t n = 0, pairs = 0; bool as_text = false;
		const FastqRec *recs = nullptr;                             // host-parsed piece: this sub-block's records
		uint64_t seq = 0; std::thread th; std::string err; uint8_t ok = 1;
	};
	std::deque<Sub> subs;
	std::deque<std::vector<FastqRec>> parsed;                       // records of the pieces the host parser took
	bool chained = false;
	for (size_t i = 0; i < n_pieces; ++i) { pieces[i].sam_bytes = pieces[i].ori_bytes = 0; chained |= pieces[i].await_path || pieces[i].publish_path; }
	// ---- the cutter: next_sub() describes the next sub-block, in input order
	size_t pi = 0, at = 0, per = 0, rec_at = 0, guess = 0;
	bool open = false, text_mode = false, piece_first = true;
	const std::vector<FastqRec> *cur_recs = nullptr;
	auto open_host = [&](const Piece &P) {
		text_mode = false;
		parsed.emplace_back();
		std::vector<FastqRec> &recs = parsed.back();
		if (!parse_fastq_parallel(P.p, P.n, *c->pipe, threads, recs)) parse_fastq(P.p, P.n, recs);
		cur_recs = &recs; rec_at = 0;
		per = pairs_per_sub_block(recs.size() / 2, threads);
		if (chained && recs.size() / 2 <= 262144 && !getenv("PANSVR_SUB_PAIRS")) per = std::max<size_t>(recs.size() / 2, 1);   // pieces of a dealt input are the caller's sub-blocks
		if (!recs.empty()) c->pipe->ensure_read_stats(recs[0]);
	};
	auto open_piece = [&]() {
		const Piece &P = pieces[pi];
		open = true; at = 0; piece_first = true;
		text_mode = c->pipe->has_device_stages() && P.n > 0 && P.p[P.n - 1] == '\n';
		if (text_mode) {
			const size_t head = std::min(P.n, (size_t)4 << 20);
			const size_t lines = count_line_ends(P.p, P.p + head);
			std::vector<FastqRec> first;
			const char *q = P.p; int l4 = 0;
			while (l4 < 4 && q < P.p + P.n) { q = (const char*)memchr(q, '\n', (size_t)(P.p + P.n - q)); if (!q) break; ++q; ++l4; }
			if (l4 == 4) parse_fastq(P.p, (size_t)(q - P.p), first);
			if (first.empty() || lines < 8) text_mode = false;
			else {
				c->pipe->ensure_read_stats(first[0]);
				const double bytes_per_pair = 8.0 * (double)head / (double)lines;
				const size_t est_pairs = (size_t)((double)P.n / bytes_per_pair);
				per = pairs_per_sub_block(est_pairs, threads);
				if ((per >= est_pairs || (chained && est_pairs <= 262144)) && !getenv("PANSVR_SUB_PAIRS")) per = (size_t)-1 / 16;   // one sub-block, however far the estimate is off
				guess = per < (size_t)1 << 40 ? (size_t)((double)per * bytes_per_pair * 1.02) + 4096 : P.n;
			}
		}
		if (!text_mode) open_host(P);
	};
	auto next_sub = [&](Sub &S) -> bool {
		for (;;) {
			if (!open) { if (pi >= n_pieces) return false; open_piece(); }
			const Piece &P = pieces[pi];
			S.piece = pi;
			if (text_mode) {
				if (at < P.n) {
					size_t lines = 0;
					const size_t e = line_end_after(P.p, P.n, at, 8 * per, guess, *c->pipe, threads, lines);
					if (lines % 8 != 0 && at == 0) { open_host(P); continue; }      // not whole pairs of 4-line records: the host parser decides
					S.p = P.p + at; S.n = e - at; S.pairs = lines / 8; S.as_text = lines % 8 == 0; S.recs = nullptr;
					if (!S.as_text) {                                               // (an odd tail behind sub-blocks already on their way)
						parsed.emplace_back();
						parse_fastq(S.p, S.n, parsed.back());
						S.recs = parsed.back().data(); S.pairs = parsed.back().size() / 2;
					}
					at = e; piece_first = false;
					return true;
				}
			} else {
				const size_t n_pairs = cur_recs->size() / 2;
				if (rec_at < n_pairs || piece_first) {
					const size_t pe = std::min(n_pairs, rec_at + per);
					S.p = nullptr; S.n = 0; S.as_text = false; S.recs = cur_recs->data() + 2 * rec_at; S.pairs = pe - rec_at;
					rec_at = pe; piece_first = false;
					return true;
				}
			}
			open = false; ++pi;
		}
	};
	// on_done: the caller takes every sub-block's output as soon as it and all earlier ones are finished, while later ones still run
	size_t handed = 0;
	auto hand_over = [&](size_t upto) { if (on_done) for (; handed < upto; ++handed) if (subs[handed].ok) (*on_done)(c->outs[handed], subs[handed].piece); };
	auto run_sub = [&](Sub *sp, BlockOutput *bop, size_t k) {          // (pointers: the deques grow while sub-blocks run)
		Sub &S = *sp;
		BlockOutput &bo = *bop;
		bo.place = nullptr;
		if (room) bo.place = [room, k](size_t total) -> char* {      // sub-block k's text goes behind that of 0 .. k-1
			std::unique_lock<std::mutex> lk(room->m);
			room->cv.wait(lk, [&]() { return room->next == k; });
			char *at2 = nullptr;
			if (!room->full && room->used + total + 1 <= room->cap) { at2 = room->p + room->used; room->used += total; }
			else if (total) room->full = true;
			room->next = k + 1;
			lk.unlock();
			room->cv.notify_all();
			return at2;
		};
		struct Always { BlockOutput &b; ~Always() { if (b.place && !b.place_called) b.place(0); b.place = nullptr; } } always{bo};
		if (S.as_text) {
			bool reparse = false;
			S.ok = c->pipe->align_block_text(S.p, S.n, S.pairs, bo, S.err, S.seq, &reparse) ? 1 : 0;
			if (!S.ok || !reparse) return;
			std::vector<FastqRec> mine;                             // not what it looked like (a header line without '@', ...): the host parser's records
			parse_fastq(S.p, S.n, mine);
			S.ok = c->pipe->align_block(mine.data(), mine.size(), bo, S.err, S.seq) ? 1 : 0;
			return;
		}
		S.ok = c->pipe->align_block(S.recs, 2 * S.pairs, bo, S.err, S.seq) ? 1 : 0;
	};
	// eight in flight per device (their device trips and host passes overlap; at most three of them in their first trip at a time,
	// pipeline.hpp: trip1_cap_); all of them while the context waits for another process's
	// stream state (pansvr_aln_await_state), so that only the in-order passes wait and every other stage of the shard is done by
	// the time the state arrives
	size_t flight = c->pipe->has_device_stages() ? 8 * c->pipe->n_devices() : 2;
	if (const char *e = getenv("PANSVR_FLIGHT")) { const long v = atol(e); if (v > 0) flight = (size_t)v; }
	if (c->pipe->awaiting_streams()) flight = (size_t)-1;
	double t_cut = 0;
	for (size_t k = 0;; ++k) {
		const double t0 = tick();
		Sub next;
		const bool more = next_sub(next);
		t_cut += tick() - t0;
		if (!more) break;
		if (k >= flight) { trace_mark(k - flight, "join_wait"); subs[k - flight].th.join(); trace_mark(k - flight, "joined"); hand_over(k - flight + 1); }
		subs.emplace_back(std::move(next));
		Sub &S = subs.back();
		S.seq = c->pipe->next_seq();
		const Piece &P = pieces[S.piece];
		const bool first_of_piece = k == 0 || subs[k - 1].piece != S.piece;
		// (the last sub-block of a piece is the one after which the cutter moves on: at == P.n in text mode, rec_at == all in host mode)
		const bool last_of_piece = text_mode ? at >= P.n : rec_at >= cur_recs->size() / 2;
		if ((first_of_piece && P.await_path) || (last_of_piece && P.publish_path))
			c->pipe->chain_at(S.seq, first_of_piece ? P.await_path : nullptr, last_of_piece ? P.publish_path : nullptr);
		if (c->outs.size() <= k) c->outs.emplace_back();
		trace_mark(S.seq, "launch");
		S.th = std::thread(run_sub, &S, &c->outs[k], k);
	}
	c->pipe->stats.t_stage[6] += t_cut;
	const size_t n_sub = subs.size();
	for (size_t k = 0; k < n_sub; ++k) if (subs[k].th.joinable()) { trace_mark(k, "join_wait"); subs[k].th.join(); trace_mark(k, "joined"); hand_over(k + 1); }
	trace_mark(n_sub, "block_done");
	for (size_t k = 0; k < n_sub; ++k) if (!subs[k].ok) { g_aln_err = subs[k].err; return PANSVR_E_CUDA; }
	hand_over(n_sub);
	sam.clear(); ori.clear();
	if (on_done) return 0;
	for (size_t k = 0; k < n_sub; ++k) {
		if (c->outs[k].sam_text.size()) add_parts(sam, c->outs[k].sam_text.data(), c->outs[k].sam_text.size());     // device path: one text per sub-block
		for (const std::string &x : c->outs[k].sam) if (!x.empty()) add_parts(sam, x.data(), x.size());
		for (const std::string &x : c->outs[k].ori) if (!x.empty()) add_parts(ori, x.data(), x.size());
	}
	return 0;
}

struct CallReport {                                           // PANSVR_TIMING=1: wall time of the call next to the sum of its stages
	pansvr_aln_ctx *c; double t_enter, t_sum0;
	static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	explicit CallReport(pansvr_aln_ctx *c_) : c(c_), t_enter(now()), t_sum0(0) { for (int i = 0; i < 8; ++i) t_sum0 += c->pipe->stats.t_stage[i]; }
	~CallReport()
	{
		if (!getenv("PANSVR_TIMING")) return;
		double s = 0;
		for (int i = 0; i < 8; ++i) s += c->pipe->stats.t_stage[i];
		fprintf(stderr, "[timing] block call %.3f s, stages %.3f s\n", now() - t_enter, s - t_sum0);
	}
};

} // namespace

namespace {
int block_of_pieces(pansvr_aln_ctx *c, Piece *pieces, size_t n_pieces, char **sam, size_t *sam_bytes, char **ori, size_t *ori_bytes)
{
	CallReport report(c);
	size_t n = 0;
	for (size_t i = 0; i < n_pieces; ++i) n += pieces[i].n;
	// The output text is assembled while the block is still being aligned: every finished sub-block is copied to its place behind the
	// earlier ones.  Room for 2.5 x the input (a record grows by its fixed fields and tags) is only address space until it is written;
	// it grows if that should not be enough.
	struct Grow {
		char *p = nullptr; size_t cap = 0, used = 0;
		bool room(size_t more) { if (used + more + 1 <= cap) return true; const size_t want = std::max(cap * 2, used + more + 1); char *q = (char*)realloc(p, want); if (!q) return false; p = q; cap = want; return true; }
	} gs, go;
	TextRoom room;
	const bool direct = c->pipe->has_device_stages();            // device path: the text comes straight into a page-locked buffer
	gs.cap = n / 2 * 5 + (1 << 16);
	if (direct) { room.cap = gs.cap; if (const char *e = getenv("PANSVR_ROOM_BYTES")) room.cap = (size_t)std::max(1L, atol(e)); room.p = g_out_pool.acquire(room.cap); }   // (the variable: tests force the spill)
	if (!room.p) gs.p = (char*)malloc(gs.cap);
	go.cap = 1 << 16; go.p = (char*)malloc(go.cap);
	bool oom = (!gs.p && !room.p) || !go.p;
	bool spilled = false;                                         // some text did not go into the room (did not fit)
	size_t placed_prefix = 0;                                     // text of the sub-blocks handed over so far that sits in the room, contiguous from its start
	double t_join = 0;
	const std::function<void(const BlockOutput&, size_t)> take = [&](const BlockOutput &o, size_t piece) {
		const double t0 = CallReport::now();
		std::vector<Part> ps, po;
		if (room.p && !o.placed && !spilled) {
			// this sub-block's text is not in the room (no room left): from here on the text is assembled by copying; what the earlier
			// sub-blocks put into the room moves over first
			spilled = true;
			gs.p = (char*)malloc(gs.cap);
			if (!gs.p || !gs.room(placed_prefix)) { oom = true; return; }
			memcpy(gs.p, room.p, placed_prefix);
			gs.used = placed_prefix;
		}
		if (o.placed && !spilled) placed_prefix += o.placed_bytes;
		if (o.placed && spilled) add_parts(ps, o.placed_ptr, o.placed_bytes);
		if (o.sam_text.size()) add_parts(ps, o.sam_text.data(), o.sam_text.size());
		for (const std::string &x : o.sam) if (!x.empty()) add_parts(ps, x.data(), x.size());
		for (const std::string &x : o.ori) if (!x.empty()) add_parts(po, x.data(), x.size());
		auto append = [&](Grow &g, const std::vector<Part> &parts) {
			size_t tot = 0;
			for (const Part &q : parts) tot += q.n;
			if (oom || !g.room(tot)) { oom = true; return; }
			std::vector<size_t> off(parts.size() + 1, g.used);
			for (size_t i = 0; i < parts.size(); ++i) off[i + 1] = off[i] + parts[i].n;
			c->pipe->parallel(parts.size(), [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) memcpy(g.p + off[i], parts[i].p, parts[i].n); }, 2);
			g.used += tot;
		};
		for (const Part &q : ps) pieces[piece].sam_bytes += q.n;
		if (o.placed && !spilled) pieces[piece].sam_bytes += o.placed_bytes;
		for (const Part &q : po) pieces[piece].ori_bytes += q.n;
		append(gs, ps); append(go, po);
		t_join += CallReport::now() - t0;
	};
	std::vector<Part> unused_s, unused_o;
	const int rc = run_block(c, pieces, n_pieces, unused_s, unused_o, &take, room.p ? &room : nullptr);
	if (rc != 0 || oom) { free(gs.p); free(go.p); if (room.p) g_out_pool.release(room.p); if (rc == 0) { g_aln_err = "out of memory"; return PANSVR_E_ARG; } return rc; }
	if (room.p && !spilled) { gs.p = room.p; gs.used = placed_prefix; }  // the usual case: everything is already in place
	else if (room.p) g_out_pool.release(room.p);
	gs.p[gs.used] = 0; go.p[go.used] = 0;
	*sam = gs.p; *ori = go.p;
	if (sam_bytes) *sam_bytes = gs.used;
	if (ori_bytes) *ori_bytes = go.used;
	c->pipe->stats.t_stage[7] += t_join;
	return 0;
}
} // namespace

int pansvr_aln_block(pansvr_aln_ctx *c, const char *fastq, size_t n, char **sam, size_t *sam_bytes, char **ori, size_t *ori_bytes)
{
	if (!c || !fastq || !sam || !ori) return PANSVR_E_ARG;
	Piece one;
	one.p = fastq; one.n = n;
	return block_of_pieces(c, &one, 1, sam, sam_bytes, ori, ori_bytes);
}

int pansvr_aln_pieces(pansvr_aln_ctx *c, pansvr_aln_piece_t *pieces, int n_pieces, char **sam, size_t *sam_bytes, char **ori, size_t *ori_bytes)
{
	if (!c || !pieces || n_pieces < 1 || !sam || !ori) return PANSVR_E_ARG;
	std::vector<Piece> ps((size_t)n_pieces);
	for (int i = 0; i < n_pieces; ++i) {
		if (!pieces[i].fastq) return PANSVR_E_ARG;
		ps[(size_t)i].p = pieces[i].fastq; ps[(size_t)i].n = pieces[i].fastq_bytes;
		ps[(size_t)i].await_path = pieces[i].await_path && *pieces[i].await_path ? pieces[i].await_path : nullptr;
		ps[(size_t)i].publish_path = pieces[i].publish_path && *pieces[i].publish_path ? pieces[i].publish_path : nullptr;
	}
	const int rc = block_of_pieces(c, ps.data(), ps.size(), sam, sam_bytes, ori, ori_bytes);
	for (int i = 0; i < n_pieces; ++i) { pieces[i].sam_bytes = ps[(size_t)i].sam_bytes; pieces[i].ori_bytes = ps[(size_t)i].ori_bytes; }
	return rc;
}

// Same block, records in BAM form: every record as bam_write1 hands it to BGZF ([block_size][core][name][cigar][seq][qual][aux]).
int pansvr_aln_block_bam(pansvr_aln_ctx *c, const char *fastq, size_t n, uint8_t **bam, size_t *bam_bytes, uint8_t **ori, size_t *ori_bytes)
{
	if (!c || !fastq || !bam || !ori) return PANSVR_E_ARG;
	CallReport report(c);
	std::vector<Part> text_sam, text_ori;
	Piece one;
	one.p = fastq; one.n = n;
	const int rc = run_block(c, &one, 1, text_sam, text_ori);
	if (rc != 0) return rc;
	const double t0 = CallReport::now();
	const size_t np = text_sam.size(), npo = text_ori.size();
	std::vector<std::vector<uint8_t>> part_s(np), part_o(npo);
	std::vector<std::string> errs(np + npo);
	// A record htslib's sam_parse1 would reject (the reference then writes a half-parsed bam1_t with undefined content, e.g. the
	// -p record of a read whose original CIGAR does not match its sequence) is left out and counted, not fatal.
	auto encode_lines = [&](const Part &text, std::vector<uint8_t> &dst, std::string &) {
		dst.reserve(text.n / 2);
		for (size_t p = 0; p < text.n;) {
			const char *nl = (const char*)memchr(text.p + p, '\n', text.n - p);
			const size_t e = nl ? (size_t)(nl - text.p) : text.n;
			if (e > p) {
				const size_t at = dst.size();
				std::string why;
				if (!bam_encode_record(text.p + p, e - p, c->bam_hdr, dst, why)) { dst.resize(at); ++c->pipe->bad_cigar_records_; }
			}
			p = e + 1;
		}
	};
	c->pipe->parallel(np + npo, [&](size_t b, size_t e, int) {
		for (size_t i = b; i < e; ++i) { if (i < np) encode_lines(text_sam[i], part_s[i], errs[i]); else encode_lines(text_ori[i - np], part_o[i - np], errs[i]); }
	}, 2);
	for (const std::string &e : errs) if (!e.empty()) { g_aln_err = e; return PANSVR_E_ARG; }
	auto join = [&](std::vector<std::vector<uint8_t>> &parts, uint8_t **out, size_t *bytes) -> bool {
		size_t tot = 0;
		for (auto &p : parts) tot += p.size();
		uint8_t *buf = (uint8_t*)malloc(tot + 1);
		if (!buf) return false;
		size_t off = 0;
		for (auto &p : parts) { if (!p.empty()) memcpy(buf + off, p.data(), p.size()); off += p.size(); }
		*out = buf;
		if (bytes) *bytes = tot;
		return true;
	};
	if (!join(part_s, bam, bam_bytes) || !join(part_o, ori, ori_bytes)) { g_aln_err = "out of memory"; return PANSVR_E_ARG; }
	c->pipe->stats.t_stage[7] += CallReport::now() - t0;
	return 0;
}

// ---- BAM files (hts_open(.., "wb") + sam_hdr_write + sam_write1 + hts_close of the reference, RR:85-94,165-175)
int pansvr_bam_open(pansvr_aln_ctx *c, const char *path, pansvr_bam_file **out)
{
	if (!c || !path || !out) return PANSVR_E_ARG;
	*out = nullptr;
	pansvr_bam_file *b = new pansvr_bam_file();
	b->f = fopen(path, "wb");
	if (!b->f) { g_aln_err = std::string("cannot open ") + path; delete b; return PANSVR_E_ARG; }
	AlnPipeline *pipe = c->pipe;
	b->w = new BamWriter(b->f, [pipe](size_t n, const std::function<void(size_t, size_t, int)> &fn) { pipe->parallel(n, fn, 2); });
	if (!b->w->write_header(c->bam_hdr)) { g_aln_err = "BAM header write failed"; delete b; return PANSVR_E_ARG; }
	*out = b;
	return 0;
}

int pansvr_bam_write(pansvr_bam_file *b, const uint8_t *records, size_t bytes)
{
	if (!b || (!records && bytes)) return PANSVR_E_ARG;
	std::vector<uint32_t> sizes;
	for (size_t off = 0; off < bytes;) {
		if (off + 4 > bytes) { g_aln_err = "truncated BAM record stream"; return PANSVR_E_ARG; }
		uint32_t bl; memcpy(&bl, records + off, 4);
		if (off + 4 + bl > bytes) { g_aln_err = "truncated BAM record stream"; return PANSVR_E_ARG; }
		sizes.push_back(4 + bl);
		off += 4 + (size_t)bl;
	}
	if (!b->w->write_records(records, sizes)) { g_aln_err = "BAM write failed"; return PANSVR_E_ARG; }
	return 0;
}

int pansvr_bam_close(pansvr_bam_file *b)
{
	if (!b) return PANSVR_E_ARG;
	const bool ok = b->w->close();
	delete b;
	if (!ok) { g_aln_err = "BAM close failed"; return PANSVR_E_ARG; }
	return 0;
}

int pansvr_aln_last_stats(const pansvr_aln_ctx *c, pansvr_aln_stats_t *out)
{
	if (!c || !out) return PANSVR_E_ARG;
	const AlnPipeline::Stats &s = c->pipe->stats;
	out->reads = (int64_t)s.reads; out->mems = (int64_t)s.mems; out->ksw_tasks = (int64_t)s.ksw_tasks; out->ksw_cells = (int64_t)s.ksw_cells;
	out->deferred_pairs = (int64_t)s.deferred_pairs;
	out->in_order_seconds = s.t_in_order; out->in_order_pairs = (int64_t)s.in_order_pairs; out->in_order_draws = (int64_t)s.in_order_draws; out->host_pairs = (int64_t)s.host_pairs; out->tie_pairs = (int64_t)s.tie_pairs;
	for (int i = 0; i < 8; ++i) out->stage_seconds[i] = s.t_stage[i];
	out->kernel_launches = s.dev.launches; out->h2d_bytes = s.dev.h2d_bytes; out->d2h_bytes = s.dev.d2h_bytes;
	out->seed_probes = s.dev.seed_probes;
	out->seed_kernel_ms = s.dev.seed_kernel_ms; out->ksw_kernel_ms = s.dev.ksw_kernel_ms; out->stage_kernel_ms = s.dev.stage_kernel_ms;
	for (int i = 0; i < 8; ++i) out->stage_kernel_ms_by[i] = s.dev.by_stage_ms[i];
	return 0;
}

void pansvr_free(void *p) { if (p && !g_out_pool.release(p)) free(p); }

int pansvr_aln_prime_read_stats(pansvr_aln_ctx *c, const char *fastq_head, size_t n)
{
	if (!c || !fastq_head) return PANSVR_E_ARG;
	std::vector<FastqRec> recs;
	const char *nl = fastq_head;
	int lines = 0;
	for (size_t i = 0; i < n && lines < 4; ++i) if (fastq_head[i] == '\n') { ++lines; nl = fastq_head + i + 1; }
	parse_fastq(fastq_head, lines == 4 ? (size_t)(nl - fastq_head) : n, recs);
	if (recs.empty()) { g_aln_err = "no FASTQ record in the given text"; return PANSVR_E_ARG; }
	c->pipe->ensure_read_stats(recs[0]);
	return 0;
}

int pansvr_aln_await_state(pansvr_aln_ctx *c, const char *path)
{
	if (!c || !path || !*path) return PANSVR_E_ARG;
	c->pipe->await_streams(path);
	return 0;
}

int pansvr_aln_publish_state(pansvr_aln_ctx *c, const char *path)
{
	if (!c || !path || !*path) return PANSVR_E_ARG;
	if (!c->pipe->publish_streams(path)) { g_aln_err = std::string("cannot write ") + path; return PANSVR_E_ARG; }
	return 0;
}

int pansvr_aln_reset(pansvr_aln_ctx *c)
{
	if (!c) return PANSVR_E_ARG;
	c->pipe->reset();
	return 0;
}

// ---- `panSVR fc_aln` command line (MAP_PARA::get_option, read_realignment.hpp:82-128)
int pansvr_fc_aln_main(int argc, char **argv)
{
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);               // the sub-blocks in flight use some fifty streams: hardware queues of their own (read when CUDA starts)
	pansvr_aln_options_t o;
	memset(&o, 0, sizeof o);
	std::string out_path = "./output.bam", ori_path = "./output_ori.bam";
	bool sam = false;
	int threads = 4;
	std::vector<int> devices;
	static const struct option lo[] = {
		{"thread", 1, 0, 't'}, {"gap-open1", 1, 0, 'O'}, {"gap-open2", 1, 0, 'P'}, {"gap-extension1", 1, 0, 'E'}, {"gap-extension2", 1, 0, 'F'},
		{"match-score", 1, 0, 'M'}, {"mis-score", 1, 0, 'm'}, {"zdrop", 1, 0, 'z'}, {"band-width", 1, 0, 'w'}, {"output", 1, 0, 'o'},
		{"output_signal_ori", 1, 0, 'p'}, {"not-ori", 0, 0, 'Q'}, {"SAM", 0, 0, 'S'}, {"max_use_read", 1, 0, 'R'}, {"device", 1, 0, 'd'}, {0, 0, 0, 0}};
	optind = 1;
	int ch;
	bool f_given = false;
	while ((ch = getopt_long(argc, argv, "t:O:P:E:F:M:m:z:w:o:p:QSR:d:", lo, 0)) != -1) {
		switch (ch) {
		case 't': threads = atoi(optarg); o.threads = threads; break;
		case 'O': o.gap_open = atoi(optarg); o.explicit_mask |= 1 << 2; break;
		case 'P': o.gap_open2 = atoi(optarg); o.explicit_mask |= 1 << 4; break;
		case 'E': o.gap_ex = atoi(optarg); o.explicit_mask |= 1 << 3; break;
		case 'F': o.gap_ex2 = atoi(optarg); f_given = true; break;
		case 'M': o.match = atoi(optarg); o.explicit_mask |= 1 << 0; break;
		case 'm': o.mismatch = atoi(optarg); o.explicit_mask |= 1 << 1; break;
		case 'z': o.zdrop = atoi(optarg); o.explicit_mask |= 1 << 6; break;
		case 'w': o.band_width = atoi(optarg); o.explicit_mask |= 1 << 7; break;     // parsed and ignored, like the reference (RR:817-827)
		case 'o': out_path = optarg; break;
		case 'p': ori_path = optarg; break;
		case 'Q': o.not_ori = 1; break;
		case 'S': sam = true; break;
		case 'R': o.max_use_read = atoi(optarg); o.explicit_mask |= 1 << 9; break;
		case 'd': for (const char *q = optarg; *q;) { devices.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; } break;
		default: return 1;
		}
	}
	(void)f_given;                                            // -t: host helper threads; the output is that of `-t 1`
	if (o.threads <= 0) o.threads = threads;
	if (argc - optind < 3) {
		fprintf(stderr, "Usage: fc_aln [Options] <IndexDir> <ReadFiles.fq|-> <ori_header.sam>\n");
		return 1;
	}
	// plain text, gzip (one inflate stream, as in the reference) or BGZF (blocks inflated on all helper threads): text_in.hpp
	TextInput in;
	{
		std::string why;
		if (!in.open(argv[optind + 1], std::max(1, threads), why)) { fprintf(stderr, "pansvr_b200 fc_aln: %s\n", why.c_str()); return 1; }
	}
	// The input is opened first: the reader fills its queue while the context is created (CUDA start-up, index files).
	// Three overlapping steps like the reference's kt_pipeline (RR:110-119): a reader thread cuts the input into blocks at
	// pair boundaries, this thread aligns them in order, a writer thread puts the records out.  The block size is ours
	// (512 k pairs): the output does not depend on it.
	struct Job { std::string fastq; bool last = false; };
	struct Result { void *main = nullptr, *ori = nullptr; size_t main_n = 0, ori_n = 0; bool last = false; };
	const long max_pairs = (o.explicit_mask >> 9) & 1 ? std::max(0, o.max_use_read) : (o.max_use_read > 0 ? o.max_use_read : 0x7fffffff);   // (-R 0 reads nothing, like the reference)
	Queue<Job> jobs(2);
	Queue<Result> results(2);
	std::atomic<bool> failed(false);
	std::thread reader([&]() {
		std::string cur;
		std::vector<char> chunk(4 << 20);
		long lines = 0, total_pairs = 0, pairs_in_block = 0;
		size_t boundary = 0;                                    // end of the last complete pair in `cur`
		bool stop = max_pairs <= 0;
		auto hand_over = [&](size_t upto, bool last) {
			Job j; j.fastq.assign(cur, 0, upto); j.last = last;
			cur.erase(0, upto); boundary = 0; pairs_in_block = 0;
			jobs.push(std::move(j));
		};
		while (!stop && !failed) {
			const long got = in.read(chunk.data(), chunk.size());
			if (got < 0) { fprintf(stderr, "pansvr_b200 fc_aln: %s\n", in.why().c_str()); failed = true; break; }
			if (got == 0) break;
			const size_t base = cur.size();
			cur.append(chunk.data(), (size_t)got);
			for (const char *q = cur.data() + base, *e = cur.data() + cur.size(); (q = (const char*)memchr(q, '\n', (size_t)(e - q))) != nullptr; ++q) {
				if (++lines % 8 != 0) continue;
				boundary = (size_t)(q - cur.data()) + 1;
				++pairs_in_block;
				if (++total_pairs >= max_pairs) { stop = true; break; }
			}
			if (stop) { cur.resize(boundary); break; }
			if (pairs_in_block >= 524288 || boundary >= ((size_t)256 << 20)) hand_over(boundary, false);
		}
		hand_over(cur.size(), true);                            // whatever is left (an unterminated last line included)
	});
	pansvr_aln_ctx *ctx = nullptr;
	if (devices.empty()) devices.push_back(0);
	int rc = pansvr_aln_create_multi(argv[optind], argv[optind + 2], &o, devices.data(), (int)devices.size(), &ctx);
	auto stop_reader = [&]() { failed = true; for (;;) { Job j = jobs.pop(); if (j.last) break; } reader.join(); in.close(); };
	if (rc != 0) { fprintf(stderr, "pansvr_b200 fc_aln: %s\n", pansvr_aln_last_error()); stop_reader(); return 1; }
	FILE *fo = nullptr, *fp = nullptr;
	pansvr_bam_file *bo = nullptr, *bp = nullptr;
	if (sam) {
		fo = fopen(out_path.c_str(), "w"); fp = fopen(ori_path.c_str(), "w");
		if (!fo || !fp) { fprintf(stderr, "pansvr_b200 fc_aln: cannot open the output files\n"); stop_reader(); return 1; }
		fputs(pansvr_aln_header_text(ctx), fo);
		fputs(pansvr_aln_header_text(ctx), fp);
	} else if (pansvr_bam_open(ctx, out_path.c_str(), &bo) != 0 || pansvr_bam_open(ctx, ori_path.c_str(), &bp) != 0) {
		fprintf(stderr, "pansvr_b200 fc_aln: %s\n", pansvr_aln_last_error());
		stop_reader();
		return 1;
	}
	std::thread writer([&]() {
		for (;;) {
			Result r = results.pop();
			if (!failed) {
				if (sam) { if (fwrite(r.main, 1, r.main_n, fo) != r.main_n || fwrite(r.ori, 1, r.ori_n, fp) != r.ori_n) failed = true; }
				else if (pansvr_bam_write(bo, (const uint8_t*)r.main, r.main_n) != 0 || pansvr_bam_write(bp, (const uint8_t*)r.ori, r.ori_n) != 0) failed = true;
			}
			pansvr_free(r.main); pansvr_free(r.ori);
			if (r.last) break;
		}
	});
	bool ok = true;
	for (;;) {
		trace_mark(0, "cli_wait_job");
		Job j = jobs.pop();
		trace_mark(j.fastq.size(), "cli_job");
		Result r; r.last = j.last;
		if (ok && !failed && !j.fastq.empty()) {
			int rc;
			if (sam) rc = pansvr_aln_block(ctx, j.fastq.data(), j.fastq.size(), (char**)&r.main, &r.main_n, (char**)&r.ori, &r.ori_n);
			else rc = pansvr_aln_block_bam(ctx, j.fastq.data(), j.fastq.size(), (uint8_t**)&r.main, &r.main_n, (uint8_t**)&r.ori, &r.ori_n);
			if (rc != 0) { fprintf(stderr, "pansvr_b200 fc_aln: %s\n", pansvr_aln_last_error()); ok = false; failed = true; }
		}
		results.push(std::move(r));
		if (j.last) break;
	}
	trace_mark(0, "cli_join_reader");
	reader.join();
	trace_mark(0, "cli_join_writer");
	writer.join();
	trace_mark(0, "cli_joined");
	if (failed && ok) { fprintf(stderr, "pansvr_b200 fc_aln: writing the output failed: %s\n", pansvr_aln_last_error()); ok = false; }
	in.close();
	if (sam) { fclose(fo); fclose(fp); }
	else if (pansvr_bam_close(bo) != 0 || pansvr_bam_close(bp) != 0) { fprintf(stderr, "pansvr_b200 fc_aln: %s\n", pansvr_aln_last_error()); ok = false; }
	if (const uint64_t nbad = ctx->pipe->bad_cigar_records_.load())
		fprintf(stderr, "pansvr_b200 fc_aln: %llu records left out: htslib would reject them (a CIGAR that does not span the read after a z-dropped "
		        "extension, or an original CIGAR that does not match the sequence); the reference writes such records with undefined content\n",
		        (unsigned long long)nbad);
	pansvr_aln_stats_t st;
	pansvr_aln_last_stats(ctx, &st);
	fprintf(stderr, "pansvr_b200 fc_aln: %ld pairs finished in order from device results, %ld by the host path\n", (long)st.tie_pairs, (long)st.host_pairs);
	fprintf(stderr, "pansvr_b200 fc_aln: %ld reads, %ld MEMs, %ld ksw tasks; stage seconds A %.3f B %.3f C %.3f D %.3f E %.3f F %.3f parse %.3f emit %.3f\n",
	        (long)st.reads, (long)st.mems, (long)st.ksw_tasks, st.stage_seconds[0], st.stage_seconds[1], st.stage_seconds[2], st.stage_seconds[3],
	        st.stage_seconds[4], st.stage_seconds[5], st.stage_seconds[6], st.stage_seconds[7]);
	trace_mark(0, "cli_destroy");
	pansvr_aln_destroy(ctx);
	trace_mark(0, "cli_exit");
	return ok ? 0 : 1;
}

} // extern "C"
