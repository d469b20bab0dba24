// pipeline.cpp -- see pipeline.hpp for the stage layout and the reference line numbers.
#include "pipeline.hpp"

#include <stdio.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "../../../include/pansvr_b200.h"
#include "stages_run.hpp"

namespace pansvr {

struct AlnPipeline::DevBuffers {                                  // one block's trip through the device stages
	StageService *svc = nullptr; bool owned = false; size_t site = 0;
	HostVec<uint8_t> text;                                        // the block's FASTQ text (pinned staging)
	HostVec<DevRead> reads;
	HostVec<DevRec> recs;
	HostVec<int32_t> drawn;                                       // what the in-order pass took from the rand() stream for the block's device pairs
	HostVec<uint32_t> tie_pair; HostVec<DevPairState> tie_done;   // the pairs the in-order pass finished itself, and what came out
	HostVec<uint32_t> host_len;
	DevStageOut out;
};

namespace {

enum { FORWARD = 1, REVERSE = 0 };                         // clib/utils.h:72-73
enum { ALN_LEFT = 0, ALN_RIGHT = 1, ALN_E2E = 2 };        // KSW_ALN_* read_realignment.hpp:185-187
enum { MAX_OUTPUT_NUMBER = 6, MIN_CHAIN_SCORE = 20, MAX_CHAIN_SCORE_DIFF = 30, MIN_CHAIN_SCORE2 = 30, MIN_ALN_SCORE = 40 };
enum { POS_N_MAX = 500, POS_N_MAX_LEVEL2 = 8000, RANDOM_NUM = 500, WAITING_LEN = 3, EINDEL = 1 };
enum { MIN_STR_REPEAT_COUNT = 4, MIN_STR_DETECT_LEN = 15 };
const uint32_t U32MAX = 0xffffffffu;
const int I32MAX = 0x7fffffff;

} // namespace

// Persistent helper threads.  Chunk t of every parallel region runs on worker t, and regions over reads are cut at
// pair boundaries, so the per-read containers a worker allocates are later grown and freed by the same thread (its own
// malloc arena: no cross-thread frees, no lock contention when a block of a million reads is torn down).
struct AlnPipeline::Workers {
	std::vector<std::thread> th;
	std::mutex m, region;                                       // region: one parallel region at a time (two blocks may be in flight)
	std::condition_variable go, done;
	uint64_t generation = 0;
	int chunks = 0, pending = 0;
	bool stop = false;
	const std::function<void(int)> *job = nullptr;
	explicit Workers(int n)
	{
		for (int i = 0; i < n; ++i) th.emplace_back([this, i]() {
			uint64_t seen = 0;
			for (;;) {
				const std::function<void(int)> *f;
				{
					std::unique_lock<std::mutex> lk(m);
					go.wait(lk, [&]() { return stop || generation != seen; });
					if (stop) return;
					seen = generation;
					if (i >= chunks) continue;
					f = job;
				}
				(*f)(i);
				{
					std::lock_guard<std::mutex> lk(m);
					if (--pending == 0) done.notify_one();
				}
			}
		});
	}
	~Workers()
	{
		{ std::lock_guard<std::mutex> lk(m); stop = true; }
		go.notify_all();
		for (std::thread &t : th) t.join();
	}
	void run(int n_chunks, const std::function<void(int)> &f)
	{
		std::lock_guard<std::mutex> one_region(region);
		std::unique_lock<std::mutex> lk(m);
		job = &f; chunks = n_chunks; pending = n_chunks; ++generation;
		go.notify_all();
		done.wait(lk, [&]() { return pending == 0; });
	}
};

// static-chunk parallel loop over [0,n): fn(begin, end, chunk_index)
void AlnPipeline::parallel(size_t n, const std::function<void(size_t, size_t, int)> &fn, size_t serial_below)
{
	const int T = workers_ ? (int)workers_->th.size() : 1;
	if (T <= 1 || n < serial_below) { fn((size_t)0, n, 0); return; }
	const size_t per = (n + T - 1) / T;
	const int chunks = (int)((n + per - 1) / per);
	workers_->run(chunks, [&](int t) { fn(std::min(n, per * t), std::min(n, per * (t + 1)), t); });
}

namespace {

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// PANSVR_TRACE=<file>: host-side phase marks of every sub-block (steady clock, s), one file per process
void trace_host(uint64_t seq, const char *what, double at = -1)
{
	static const char *path = getenv("PANSVR_TRACE");
	if (!path) return;
	static std::mutex m; static FILE *f = nullptr;
	std::lock_guard<std::mutex> lk(m);
	if (!f) f = fopen((std::string(path) + "." + std::to_string((long)getpid()) + ".host").c_str(), "a");
	if (f) { fprintf(f, "H %llu %s %.6f\n", (unsigned long long)seq, what, at >= 0 ? at : now()); fflush(f); }
}

// ---------------------------------------------------------------------------------------------- small pieces
struct Dna5Table {                                              // charToDna5n, read_realignment.cpp:180-202, as a table (no branches)
	uint8_t t[256];
	Dna5Table() { memset(t, 0, sizeof t); t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; t['n'] = 4; }
};
const Dna5Table g_dna5;
inline uint8_t dna5(unsigned char c) { return g_dna5.t[c]; }
struct RevTable {                                               // getReverseChar, clib/bam_file.c:320-328, as a table
	char t[256];
	RevTable() { memset(t, 'N', sizeof t); t['A'] = t['a'] = 'T'; t['C'] = t['c'] = 'G'; t['G'] = t['g'] = 'C'; t['T'] = t['t'] = 'A'; }
};
const RevTable g_rev;
struct Nt16Norm {                                               // a base after htslib's sam_parse1 -> sam_format1 round trip:
	char t[256];                                                // seq_nt16_str[seq_nt16_table[c]] (upper case, IUPAC kept, anything else 'N')
	Nt16Norm()
	{
		memset(t, 'N', sizeof t);
		const char *code = "=ACMGRSVTWYHKDBN";
		for (int i = 0; i < 16; ++i) { t[(unsigned char)code[i]] = code[i]; if (code[i] >= 'A') t[(unsigned char)(code[i] + 32)] = code[i]; }
		t['0'] = 'A'; t['1'] = 'C'; t['2'] = 'G'; t['3'] = 'T';
	}
};
const Nt16Norm g_nt16_norm;
inline char rev_char(char c) { return g_rev.t[(unsigned char)c]; }
void rev_str(char *s, int len)                                  // getReverseStr_char, clib/bam_file.c:330-340
{
	const int half = len >> 1;
	for (int i = 0; i < half; ++i) { const char t = s[i]; s[i] = rev_char(s[len - 1 - i]); s[len - 1 - i] = rev_char(t); }
	if (len & 1) s[half] = rev_char(s[half]);
}
template <class T> void rev_qual(T *q, int len)                 // getReverseStr_qual(_char), clib/bam_file.c:342-360
{                                                               // sic: i runs to len/2 inclusive, so an even length re-swaps its middle pair
	const int half = len >> 1;
	for (int i = 0; i < half + 1; ++i) { const int ri = len - 1 - i; const T t = q[i]; q[i] = q[ri]; q[ri] = t; }
}

CigarPath cig_char(char t, int size)                            // CIGAR_PATH(char, uint16_t), read_realignment.hpp:133-149
{
	static const char *ops = "MIDNSHP=XB";
	const char *p = strchr(ops, t);
	CigarPath c;
	c.type = p ? (uint8_t)(p - ops) : 0;
	c.size = (int16_t)(uint16_t)size;
	return c;
}
CigarPath cig_bin(uint32_t w) { CigarPath c; c.type = (uint8_t)(w & 0xf); c.size = (int16_t)(w >> 4); return c; }
bool cig_try_merge(CigarPath &a, const CigarPath &b)            // CIGAR_PATH::try_merge, read_realignment.hpp:159-178
{
	if (b.size < 0) {
		if (a.type == 0) { a.size = (int16_t)(a.size + b.size); return true; }
		if (a.type == 2) { a.size = (int16_t)(a.size - b.size); return true; }
		return true;                                            // the reference asserts here
	}
	if (a.type == b.type || b.size == 0) { a.size = (int16_t)(a.size + b.size); return true; }
	return false;
}

struct VertexU { uint64_t uid; uint32_t read_pos, uni_pos_off, length1, length2, pos_n, cov; };
struct UniSeed { uint32_t read_begin, read_end, seed_id, ref_begin, ref_end, cov; };   // UNI_SEED, graph.hpp:42-49
struct PathNode { float dist; int32_t pre_node; uint8_t used; };
struct Edge { uint32_t to, from; int weight; float penalty; };

int cmp_mem(const void *a, const void *b)                       // vertex_MEM::cmp, deBGA_index.hpp:33-50
{
	const Mem *x = (const Mem*)a, *y = (const Mem*)b;
	if (x->uid > y->uid) return 1;
	if (x->uid < y->uid) return -1;
	if (x->read_pos > y->read_pos) return 1;
	if (x->read_pos < y->read_pos) return -1;
	return 0;
}
int cmp_seed(const void *a, const void *b)                      // UNI_SEED::cmp, graph.cpp:14-33
{
	const UniSeed *x = (const UniSeed*)a, *y = (const UniSeed*)b;
	if (x->ref_end > y->ref_end) return 1;
	if (x->ref_end < y->ref_end) return -1;
	if (x->ref_begin > y->ref_begin) return 1;
	if (x->ref_begin < y->ref_begin) return -1;
	return 0;
}

// one chained strand of a read: sorted seeds + longest-path table (Graph_handler)
struct Graph {
	std::vector<UniSeed> v;
	std::vector<PathNode> path;
	bool is_str = false;
	// state of sort_output
	float max_distance = 0; uint32_t max_index = 0;
	std::vector<int> same_top;
};

// a candidate alignment (MAX_IDX_OUTPUT, read_realignment.hpp:232-318)
struct Result {
	uint32_t align_score = 0, chain_score = 0, max_index = 0, read_bg = 0;
	const SvInfo *sv = nullptr;
	uint8_t mapq = 0;
	bool has_mate = false;
	uint32_t mate_chr = 0, mate_ref_bg = 0;
	const SvInfo *mate_sv = nullptr;
	bool is_ori = false;
	uint32_t chr = 0, ref_bg = 0;
	int direction = FORWARD;
	std::vector<CigarPath> cigar;
	bool cigar_ok = true;          // the CIGAR spans the read (reverseGIGAR's check, RRH:296-299)
	int rst_idx = 0;
};

// what get_ksw_score yields for one (strand, end node): pieces in emission order
struct Piece {
	int kind;                      // 0 = literal CIGAR entry, 1 = ksw task
	CigarPath lit;
	int task, type;                // ksw task id in the block list, KSW_ALN_* type
};
struct NodeAln {
	std::vector<Piece> pieces;
	int fixed_score = 0;           // everything but the ksw scores
	int read_begin_alignment = 0;
	bool planned = false;
	// filled the first time a candidate ending at this node is extended (finish_read); a pair that is finished twice
	// -- against the probe, then in the replay -- reuses it
	bool resolved = false;
	uint32_t align_score = 0;
	std::vector<CigarPath> cigar;
	bool cigar_ok = true;
};

} // namespace

// ================================================================================================ per-read state
struct ReadState {
	const FastqRec *rec = nullptr;
	std::string comment;                       // mutable copy (the reference edits its kseq_t in place)
	int read_l = 0;
	bool seq_twice_reversed = false;           // the reference reverse-complemented its kseq_t in place and back (output_BAM on a
	                                           // reverse-strand record): anything but ACGT is 'N' from then on, e.g. in the -p record
	bool has_n = false, skip = false;          // skip: early-out of single_end_handler::align (RR:413-414)
	// original alignment (parse_ori_mapping_rst)
	Result ori;
	bool ori_unmapped = false;
	// encoded read
	std::vector<uint8_t> bin[2];
	bool is_str = false;
	std::vector<uint8_t> seed_list[2];
	std::vector<uint64_t> bits[2];             // 32 bases per word, one spare zero word
	// A read with 1..3 'N' is prepared once per possible substitution ("variant" read states appended after the real ones);
	// the replay draws the real rand()%4 values and adopts the matching variant.
	int n_draws = 0;                           // 'N' bases = rand() draws of binary_read_2_bit
	int64_t var_base = -1;                     // real read: index of its first variant (4^n_draws of them), -1 = none
	int64_t var_of = -1;                       // variant: index of the real read
	uint32_t var_code = 0;                     // variant: substitution of the j-th N = (var_code >> 2j) & 3
	bool in_order_only = false;                // real read: must be prepared during the replay (too many N, or random_r needed)
	bool batched = false, needs_rand = false;  // needs_rand: a unipath with > 500 positions draws from random_r (expand_seed)
	bool dev = false;                          // stages A..F1 of this read state ran on the device (stages_run.hpp)
	int64_t dev_index = -1;                    // its row in the block's device read table
	std::vector<VertexU> vu[2];
	int job[2] = {-1, -1};
	Graph g[2];
	std::vector<std::pair<uint64_t, NodeAln>> node_aln;   // key = strand << 32 | node, ascending
	// results
	int result_num = 0;
	std::vector<Result> result;                // at most 2 * MAX_OUTPUT_NUMBER, reserved once so that pointers stay valid
	Result *primary = nullptr, *secondary = nullptr;
};

struct KswView { const int32_t *res; const uint32_t *cig; int cap; };   // results of a task list, as finish_read reads them
struct KswTaskList {
	std::vector<uint8_t> q, t;
	std::vector<int64_t> qoff, toff;
	std::vector<int32_t> qlen, tlen;
	std::vector<int32_t> res;
	std::vector<uint32_t> cig;
	int cap = 64;
	int add(const uint8_t *qs, int ql, const uint8_t *ts, int tl)
	{
		qoff.push_back((int64_t)q.size()); toff.push_back((int64_t)t.size());
		qlen.push_back(ql); tlen.push_back(tl);
		q.insert(q.end(), qs, qs + ql); t.insert(t.end(), ts, ts + tl);
		return (int)qlen.size() - 1;
	}
	void clear() { q.clear(); t.clear(); qoff.clear(); toff.clear(); qlen.clear(); tlen.clear(); res.clear(); cig.clear(); }
	KswView view() const { KswView v; v.res = res.data(); v.cig = cig.data(); v.cap = cap; return v; }
};

// The process-wide rand() stream as the replay sees it: the real generator, or a probe that only records that it
// was asked (a pair that never asks is independent of the stream position and can be finished on any thread).
struct RandTap {
	enum { MAX_SCRIPT = 12 };
	GlibcRandom *real = nullptr;
	uint32_t calls = 0;
	// probe mode: the outcome of call k is script[k] (0 beyond the script) and its modulus is recorded, so that every
	// combination of outcomes of a read can be enumerated (explore_read)
	uint8_t script[MAX_SCRIPT], moduli[MAX_SCRIPT];
	uint32_t script_len = 0;
	bool too_deep = false;
	int32_t draw(int32_t m)                                  // rand() % m
	{
		const uint32_t k = calls++;
		if (real) return real->next() % m;
		if (k >= MAX_SCRIPT || m > 255) { too_deep = true; return 0; }
		moduli[k] = (uint8_t)m;
		return k < script_len ? (int32_t)script[k] : 0;
	}
	void restart(uint32_t len) { calls = 0; script_len = len; too_deep = false; }
	std::vector<CigarPath> cigar_scratch;                    // (per-thread scratch of finish_read rides along)
};
struct PairEvent { int8_t i, j; uint8_t tie; };             // store_pair reached with fin >= max_score: candidates i/j (-1 = none)

struct AlnPipeline::Impl {
	AlnPipeline &P;
	const DebgaIndex &idx;
	bool count_bad = true;                                       // output_bam counts the records it leaves out (off for scratch runs)
	explicit Impl(AlnPipeline &p) : P(p), idx(p.idx_) {}

	// ---------------------------------------------------------------- stage A
	void parse_ori(ReadState &r)                                // parse_ori_mapping_rst, read_realignment.hpp:392-429
	{
		Result &o = r.ori;
		o = Result();
		o.is_ori = true;
		std::string &c = r.comment;
		const char *tok[10]; size_t tok_l[10]; size_t nul_at[10];
		int nt = 0, nn = 0;
		size_t i = 0;
		while (nt < 10 && i < c.size()) {                        // strtok_r(.., "_"): skip separators, cut at the next one
			while (i < c.size() && c[i] == '_') ++i;
			if (i >= c.size()) break;
			size_t j = i;
			while (j < c.size() && c[j] != '_') ++j;
			tok[nt] = c.data() + i; tok_l[nt] = j - i; ++nt;
			if (j < c.size()) nul_at[nn++] = j;
			i = j + 1;
		}
		auto num = [&](int k) -> int {                            // atoi of token k
			if (k >= nt) return 0;
			const char *p = tok[k], *e = p + tok_l[k];
			while (p < e && (*p == ' ' || *p == '\t')) ++p;
			bool neg = false;
			if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
			long v = 0;
			while (p < e && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
			return (int)(neg ? -v : v);
		};
		o.chr = (uint32_t)num(0);
		o.ref_bg = (uint32_t)num(1);
		o.read_bg = (uint32_t)num(2);
		o.align_score = (uint32_t)num(3);
		o.mapq = (uint8_t)num(4);
		o.direction = (nt > 9 && tok_l[9] > 0 && tok[9][0] == 'F') ? FORWARD : REVERSE;
		r.ori_unmapped = nt > 9 && tok_l[9] > 1 && tok[9][1] == 'Y';
		o.cigar.clear();
		o.cigar.reserve(2);
		if (o.read_bg > 0) o.cigar.push_back(cig_char('S', (int)o.read_bg));
		o.cigar.push_back(cig_char('M', r.read_l - (int)o.read_bg));
		o.sv = nullptr; o.has_mate = false;
		if (o.ref_bg >= (uint32_t)I32MAX) o.ref_bg = 1;
		for (int k = 0; k < nn; ++k) if (nul_at[k] + 1 < c.size()) c[nul_at[k]] = ',';  // separators the tokenizer consumed come back as ','
	}

	void encode(ReadState &r)                                    // binary_read_2_bit, read_realignment.cpp:646-654
	{
		const int L = r.read_l;
		r.bin[0].resize(L); r.bin[1].resize(L);
		uint8_t *f = r.bin[0].data(), *rv = r.bin[1].data();
		const char *seq = r.rec->seq;
		if (r.n_draws == 0 && r.var_of < 0) {                     // no 'N': nothing to draw
			for (int i = 0; i < L; ++i) { const uint8_t c = dna5((unsigned char)seq[i]); f[i] = c; rv[L - i - 1] = c ^ 3; }
			return;
		}
		int nth = 0;
		for (int i = 0; i < L; ++i) {
			char ch = seq[i];
			if (ch == 'N') ch = "ACGT"[r.var_of >= 0 ? (r.var_code >> (2 * nth++)) & 3 : (uint32_t)(P.rand_.next() % 4)];
			const uint8_t c = dna5((unsigned char)ch);
			f[i] = c;
			rv[L - i - 1] = c ^ 3;
		}
	}

	// `words` words are written at out[off..]: the packed bases, then zeros (the spare word the k-mer and window reads rely on)
	static void pack64(const std::vector<uint8_t> &b, uint64_t *out, size_t off, size_t words)   // binary_read_64_bit, RR:295-300
	{
		const size_t n = b.size();
		for (size_t k = n >> 5; k < words; ++k) out[off + k] = 0;
		size_t i = 0;
		for (; i + 32 <= n; i += 32) {
			uint64_t w = 0, any = 0;
			for (int k = 0; k < 32; k += 8) {                      // eight codes at a time: x * 0x40100401 gathers four 2-bit codes into one byte
				uint64_t x;
				memcpy(&x, b.data() + i + k, 8);
				any |= x;
				const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
				w = (w << 16) | (uint64_t)(((lo * 0x40100401u) >> 24) << 8) | (uint64_t)((hi * 0x40100401u) >> 24);
			}
			if (any & 0xfcfcfcfcfcfcfcfcull) {                     // a code 4 (lower-case 'n') spills into its neighbour in the reference: literal path
				w = 0;
				for (int k = 0; k < 32; ++k) w = (w << 2) | b[i + k];
			}
			out[off + (i >> 5)] = w;
		}
		if (i < n) {
			uint64_t w = 0;
			for (size_t k = i; k < n; ++k) w = (w << 2) | b[k];
			out[off + (i >> 5)] = w << ((32 - (n - i)) << 1);
		}
	}

	// scratch of str_census, one per worker: slots are valid when their stamp equals the current read's
	struct CensusScratch {
		std::vector<uint64_t> keys;
		std::vector<uint32_t> stamp;
		std::vector<uint16_t> cnt, slot_of;
		uint32_t now = 0;
	};

	// STR census of the forward strand (read_realignment.cpp:553-598): k-mer multiplicities, mask, forced seeds
	void str_census(ReadState &r, const uint64_t *bits, CensusScratch &cs)
	{
		const uint32_t L = (uint32_t)r.read_l, kn = L - LEN_KMER + 1;
		// multiplicity of every 20-mer (the reference's std::map<kmer,count>): open addressing, 4x slots
		uint32_t cap = 64;
		while (cap < 4 * kn) cap <<= 1;
		if (cs.keys.size() < cap) { cs.keys.assign(cap, 0); cs.stamp.assign(cap, 0); cs.cnt.assign(cap, 0); cs.now = 0; }
		if (cs.slot_of.size() < kn) cs.slot_of.resize(kn);
		if (++cs.now == 0) { std::fill(cs.stamp.begin(), cs.stamp.end(), 0u); cs.now = 1; }
		uint64_t *keys = cs.keys.data(); uint32_t *stamp = cs.stamp.data(); uint16_t *cnt = cs.cnt.data(), *slot_of = cs.slot_of.data();
		const uint32_t tick = cs.now;
		uint32_t distinct = 0;
		// the 20-mer at i is get_kmer(i, bits); when every base code is 0..3 (no lower-case 'n', whose code 4 spills into
		// the neighbouring base of the packed word) it is also the rolling value over the byte codes, which is cheaper
		const uint8_t *bin = r.bin[0].data();
		bool plain = true;
		for (uint32_t i = 0; i < L; ++i) plain &= bin[i] < 4;
		const uint64_t kmask = (1ull << (2 * LEN_KMER)) - 1;
		uint64_t roll = 0;
		if (plain) for (uint32_t i = 0; i + 1 < LEN_KMER; ++i) roll = (roll << 2) | bin[i];
		for (uint32_t i = 0; i < kn; ++i) {
			uint64_t k;
			if (plain) { roll = ((roll << 2) | bin[i + LEN_KMER - 1]) & kmask; k = roll; }
			else k = get_kmer(i, bits);
			uint32_t h = (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 40) & (cap - 1);
			while (stamp[h] == tick && keys[h] != k) h = (h + 1) & (cap - 1);
			if (stamp[h] != tick) { stamp[h] = tick; keys[h] = k; cnt[h] = 0; ++distinct; }
			++cnt[h];
			slot_of[i] = (uint16_t)h;
		}
		r.is_str = (size_t)distinct < (size_t)kn - MIN_STR_DETECT_LEN;
		r.seed_list[0].clear(); r.seed_list[1].clear();
		if (!r.is_str) return;
		std::vector<uint8_t> &sl = r.seed_list[0];
		sl.assign(L, 0);                                         // repeat_seed_info is per handler and larger; only [0,kn) is read
		for (uint32_t i = 0; i < kn; ++i) sl[i] = cnt[slot_of[i]] >= MIN_STR_REPEAT_COUNT ? 0 : 1;
		int bg = 0, ed = 0;
		for (uint32_t i = 0; i < SEED_STEP; ++i) {
			bg += sl[i] == 0; ed += sl[L - LEN_KMER - i] == 0;
			sl[i] += 2; sl[L - LEN_KMER - i] += 4;
		}
		if (bg < SEED_STEP && ed < SEED_STEP) {
			int n = 0;
			for (uint32_t i = 0; n < SEED_STEP && i < kn; ++i) { if (sl[i] > 0) continue; sl[i] += 8; ++n; }
		}
		r.seed_list[1].assign(sl.begin(), sl.begin() + kn);
		rev_qual(r.seed_list[1].data(), (int)kn);                // getReverseStr_qual(seed_list, kn), RR:601
		r.seed_list[0].resize(kn);
	}

	// ---------------------------------------------------------------- stage C
	// m[0..n) is sorted in place.  The reference reads one element past the end in its loop conditions (a slot that can never
	// match); the `j < n` tests below come first instead.
	void merge_mems(Mem *m, uint32_t n, std::vector<VertexU> &out)   // merge_seed_in_unipath, deBGA_index.cpp:151-217
	{
		out.clear();
		if (n == 0) return;
		out.reserve(n);
		if (n == 1) {
			VertexU u; u.uid = m[0].uid; u.read_pos = m[0].read_pos; u.uni_pos_off = m[0].uni_pos_off; u.pos_n = m[0].pos_n;
			u.length1 = u.length2 = u.cov = m[0].length;
			out.push_back(u);
			return;
		}
		qsort(m, n, sizeof(Mem), cmp_mem);
		uint64_t uid_t = m[0].uid;
		uint32_t j = 0;
		while (j < n) {
			const uint32_t s1 = j;
			uint32_t cov = m[s1].length;
			++j;
			while (j < n && uid_t == m[j].uid && m[j].uni_pos_off > m[j - 1].uni_pos_off) {
				const int diff = (int)(m[j].read_pos - m[j - 1].read_pos - m[j - 1].length);
				if (diff > WAITING_LEN) break;
				const int c_eindel = (int)((m[j].uni_pos_off - m[j - 1].uni_pos_off) - (m[j].read_pos - m[j - 1].read_pos));
				if (std::abs(c_eindel) < EINDEL) { cov += diff > 0 ? m[j].length : (uint32_t)(diff + (int)m[j].length); ++j; }
				else break;
			}
			const uint32_t e1 = j - 1;
			VertexU u;
			u.uid = m[s1].uid; u.read_pos = m[s1].read_pos; u.uni_pos_off = m[s1].uni_pos_off; u.pos_n = m[s1].pos_n; u.cov = cov;
			if (s1 == e1) u.length1 = u.length2 = m[s1].length;
			else {
				u.length1 = m[e1].read_pos + m[e1].length - m[s1].read_pos;
				u.length2 = m[e1].uni_pos_off + m[e1].length - m[s1].uni_pos_off;
			}
			out.push_back(u);
			if (j < n) uid_t = m[j].uid;
		}
	}

	void expand(const std::vector<VertexU> &vu, std::vector<UniSeed> &out, GlibcRandom &rr)   // expand_seed, deBGA_index.cpp:219-258
	{
		size_t total = 0;
		for (const VertexU &u : vu) total += u.pos_n > POS_N_MAX ? (size_t)RANDOM_NUM : u.pos_n;
		out.reserve(total);
		for (uint32_t i = 0; i < vu.size(); ++i) {
			const VertexU &u = vu[i];
			auto push = [&](uint32_t mpos) {
				UniSeed s;
				s.seed_id = i; s.read_begin = u.read_pos; s.read_end = u.read_pos + u.length1 - 1;
				s.ref_begin = (uint32_t)(idx.pos[mpos + idx.posp[u.uid]] + u.uni_pos_off - 1);
				s.ref_end = s.ref_begin + u.length2 - 1; s.cov = u.cov;
				out.push_back(s);
			};
			if (u.pos_n > POS_N_MAX) {
				if (u.pos_n > POS_N_MAX_LEVEL2) return;          // sic: abandons every remaining vertex
				for (int k = 0; k < RANDOM_NUM; ++k) push((uint32_t)(rr.next() % (int32_t)u.pos_n));
			} else for (uint32_t mpos = 0; mpos < u.pos_n; ++mpos) push(mpos);
		}
	}

	void chain(Graph &g, std::vector<Edge> &edges)                // Graph_handler::process + dynamic_programming_path, graph.cpp:53-150
	{
		std::vector<UniSeed> &v = g.v;
		const uint32_t n = (uint32_t)v.size();
		g.path.clear();
		if (n == 0) return;
		qsort(v.data(), n, sizeof(UniSeed), cmp_seed);
		const int max_ref_dis = g.is_str ? 400 : 50, max_read_dis = g.is_str ? 400 : 50;
		const uint32_t max_step = g.is_str ? 80 : 40, max_gap = g.is_str ? 20 : 50;
		const uint32_t step = std::min(n, max_step);
		g.path.resize(n);
		for (uint32_t i = 0; i < n; ++i) { g.path[i].dist = (float)v[i].cov; g.path[i].pre_node = -1; g.path[i].used = 0; }
		edges.clear();
		for (uint32_t a = 0; a + 1 < n; ++a) {
			const uint32_t read_end = v[a].read_end, ref_end = v[a].ref_end, seed_id = v[a].seed_id;
			const uint32_t stop = std::min(n, a + step);
			for (uint32_t b = a + 1; b < stop; ++b) {
				if (v[b].seed_id == seed_id) continue;
				if (v[b].ref_end == ref_end) continue;
				const int32_t dis_ref = (int32_t)(v[b].ref_begin - ref_end);
				if (dis_ref > max_ref_dis) break;
				const int32_t dis_read = (int32_t)(v[b].read_begin - read_end);
				if (dis_read > max_read_dis) continue;
				const uint32_t abs_gap = dis_read > dis_ref ? (uint32_t)(dis_read - dis_ref) : (uint32_t)(dis_ref - dis_read);
				if (abs_gap > max_gap) continue;
				const float penalty = abs_gap == 0 ? 0.f : (float)((abs_gap >> 3) + 3);
				uint32_t weight;
				if (dis_read == dis_ref) weight = v[b].cov - (uint32_t)std::max(1 - dis_read, 0);
				else if (dis_read > 0 && dis_ref > 0) weight = v[b].cov;
				else if (dis_read >= -5 && dis_read <= 0 && dis_ref >= -5) weight = v[b].cov + (uint32_t)std::min(dis_read, dis_ref);
				else continue;
				Edge e; e.to = b; e.from = a; e.weight = (int)weight; e.penalty = penalty;
				edges.push_back(e);
			}
		}
		if (edges.empty()) return;
		std::stable_sort(edges.begin(), edges.end(), [](const Edge &x, const Edge &y) { return x.to < y.to; });   // per node, in insertion order
		size_t k = 0;
		while (k < edges.size()) {
			const uint32_t to = edges[k].to;
			float cur = 0; int32_t pre = -1;
			for (; k < edges.size() && edges[k].to == to; ++k) {
				const float temp = g.path[edges[k].from].dist + (float)edges[k].weight - edges[k].penalty;
				if (cur <= temp) { cur = temp; pre = (int32_t)edges[k].from; }
			}
			g.path[to].dist = cur; g.path[to].pre_node = pre;
		}
	}

	// ---------------------------------------------------------------- stage D: get_ksw_score as a plan (RR:308-400, 893-986)
	struct PlanScratch { std::vector<uint8_t> tseq, qrev; };
	struct Planner {
		Impl &I; ReadState &r; int strand; KswTaskList &tasks; NodeAln &out;
		std::vector<uint8_t> &tseq, &qrev;
		int total_q_len = 0; bool last_simple = false;
		Planner(Impl &i, ReadState &rs, int s, KswTaskList &t, NodeAln &o, PlanScratch &sc) : I(i), r(rs), strand(s), tasks(t), out(o), tseq(sc.tseq), qrev(sc.qrev) {}
		void lit(char t, int size) { Piece p; p.kind = 0; p.lit = cig_char(t, size); p.task = -1; p.type = 0; out.pieces.push_back(p); }
		int mismatch(int rs, int re, int fs, int fe)             // get_misMatch, RR:893-908
		{
			int qlen = re - rs, tlen = fe - fs;
			if (fe < fs) { tlen = 0; qlen += fs - fe; }
			tseq.resize(std::max(tlen, qlen) + 1);
			I.idx.refseq(tseq.data(), (uint32_t)tlen, (uint32_t)fs);
			const uint8_t *q = r.bin[strand].data() + rs;
			int nm = 0;
			for (int i = 0; i < qlen; ++i) nm += (i < tlen ? q[i] != tseq[i] : 1);
			return nm > 3 ? 3 : nm;
		}
		void alignment(int rs, int re, int fs, int fe, int type)  // KSW_ALN_handler::alignment, RR:910-986
		{
			int qlen = re - rs, tlen = fe - fs;
			if (fe < fs) { tlen = 0; qlen += fs - fe; }
			tseq.assign((size_t)std::max(tlen, 0) + 1, 0);
			I.idx.refseq(tseq.data(), (uint32_t)tlen, (uint32_t)fs);
			const uint8_t *q = r.bin[strand].data() + rs;
			if (type == ALN_LEFT) {
				std::reverse(tseq.begin(), tseq.begin() + tlen);
				qrev.assign(q, q + qlen);
				std::reverse(qrev.begin(), qrev.end());
				q = qrev.data();
			}
			total_q_len += qlen;
			bool simple = false; uint32_t nm = 0;
			if (qlen == 0 || tlen == 0) { simple = true; nm = (uint32_t)(qlen + tlen); }
			else if (qlen == tlen || type != ALN_E2E) {
				for (int i = 0; i < qlen && nm < 6; ++i) nm += (i < tlen ? q[i] != tseq[i] : 1);
				if (nm == 1 || (nm < 6 && (int)(nm << 3) < qlen)) simple = true;
			}
			last_simple = simple;
			const AlnOptions &o = I.P.opt;
			if (simple) {
				if (qlen == 0 || tlen == 0) {
					if (nm != 0) out.fixed_score -= std::min(o.gap_open + ((int)nm - 1) * o.gap_ex, o.gap_open2 + ((int)nm - 1) * o.gap_ex2);
				} else out.fixed_score += qlen * o.match - (int)nm * (o.match + o.mismatch);
				if (qlen == 0) lit('D', tlen); else if (tlen == 0) lit('I', qlen); else lit('M', qlen);
				if (fe < fs) lit('D', fe - fs);
			} else if ((int64_t)tlen * qlen > 1000000) {           // align_non_splice guard, RR:874-887: fixed 2-element CIGAR, reset ez
				const int sc = type == ALN_E2E ? 0 : PANSVR_KSW_NEG_INF;
				out.fixed_score += sc;
				const uint32_t c0 = (uint32_t)qlen << 4 | 1, c1 = (uint32_t)tlen << 4 | 3;
				Piece a, b; a.kind = b.kind = 0; a.task = b.task = -1; a.type = b.type = 0;
				a.lit = cig_bin(type == ALN_LEFT ? c0 : c1); b.lit = cig_bin(type == ALN_LEFT ? c1 : c0);
				out.pieces.push_back(a); out.pieces.push_back(b);
			} else {
				Piece p; p.kind = 1; p.type = type; p.lit = CigarPath{0, 0};
				p.task = tasks.add(q, qlen, tseq.data(), tlen);
				out.pieces.push_back(p);
			}
		}
		void run(int first_node)
		{
			const Graph &g = r.g[strand];
			const AlnOptions &o = I.P.opt;
			const int read_l = r.read_l;
			int aln_read_begin = read_l, aln_read_end = read_l, aln_ref_begin = I32MAX, aln_ref_end = I32MAX;
			int last_aln_begin = read_l, last_ref_begin = I32MAX, unitig_mis = 0;
			for (int node = first_node; node != -1;) {
				const UniSeed &s = g.v[node];
				const int m_rb = (int)s.read_begin, m_re = (int)s.read_end, m_fb = (int)s.ref_begin, m_fe = (int)s.ref_end;
				aln_read_begin = std::min(aln_read_begin, m_re);
				aln_ref_begin = std::min(aln_ref_begin, m_fe);
				if (aln_read_begin <= aln_read_end) {
					if (aln_read_end < last_aln_begin) {
						const int mem_len = last_aln_begin - aln_read_end;
						unitig_mis += mismatch(aln_read_end, aln_read_end + mem_len, last_ref_begin, last_ref_begin + mem_len);
						lit('M', mem_len);
					}
					last_aln_begin = aln_read_begin;
					if (aln_ref_end == I32MAX) {
						aln_ref_end = aln_ref_begin + (aln_read_end - aln_read_begin) + 30;
						alignment(aln_read_begin, aln_read_end, aln_ref_begin, aln_ref_end, ALN_RIGHT);
					} else alignment(aln_read_begin, aln_read_end, aln_ref_begin, aln_ref_end, ALN_E2E);
				} else {
					const int d_read = aln_read_end - aln_read_begin, d_ref = aln_ref_end - aln_ref_begin;
					if (d_read != d_ref) {
						const int del = std::abs(d_ref - d_read);
						out.fixed_score -= std::min(o.gap_open + (del - 1) * o.gap_ex, o.gap_open2 + (del - 1) * o.gap_ex2);
					}
				}
				aln_read_end = m_rb; last_ref_begin = m_fb; aln_ref_end = m_fb;
				(void)m_re;
				const int next = g.path[node].pre_node;
				if (next == -1) break;
				node = next;
			}
			if (aln_read_end < last_aln_begin) {
				const int mem_len = last_aln_begin - aln_read_end;
				unitig_mis += mismatch(aln_read_end, aln_read_end + mem_len, last_ref_begin, last_ref_begin + mem_len);
				lit('M', mem_len);
			}
			aln_read_begin = 0; aln_ref_begin = 0;
			int rba = 0;
			if (aln_read_begin < aln_read_end) {
				aln_ref_begin = std::max(0, aln_ref_end - (aln_read_end - aln_read_begin) - 30);
				alignment(aln_read_begin, aln_read_end, aln_ref_begin, aln_ref_end, ALN_LEFT);
				if (aln_ref_end > aln_ref_begin) rba = last_simple ? aln_ref_end - aln_ref_begin - 30 : aln_ref_end - aln_ref_begin;
			}
			out.fixed_score += (read_l - total_q_len) * o.match;
			out.fixed_score -= unitig_mis * (o.match + o.mismatch);
			out.read_begin_alignment = rba;
			out.planned = true;
		}
	};

	void plan_read(ReadState &r, KswTaskList &tasks, PlanScratch &scratch)
	{
		r.node_aln.clear();
		uint32_t best = 0;                                         // chain scores are compared as uint32 (RR:423-429, 442)
		for (int s = 0; s < 2; ++s) for (const PathNode &p : r.g[s].path) best = std::max(best, (uint32_t)p.dist);
		if (best < (uint32_t)MIN_CHAIN_SCORE) return;
		for (int s = 0; s < 2; ++s) {
			const Graph &g = r.g[s];
			for (uint32_t n = 0; n < g.path.size(); ++n) {
				const uint32_t c = (uint32_t)g.path[n].dist;
				if (c < (uint32_t)MIN_CHAIN_SCORE2 || c + MAX_CHAIN_SCORE_DIFF < best) continue;
				r.node_aln.emplace_back((uint64_t)s << 32 | n, NodeAln());
				NodeAln &na = r.node_aln.back().second;
				na.pieces.reserve(8);
				Planner pl(*this, r, s, tasks, na, scratch);       // g[1] is the reverse strand, bin[1] its sequence
				pl.run((int)n);
			}
		}
	}

	// ---------------------------------------------------------------- stage F
	int sort_output(ReadState &r, int s, Result &rst, int direction, RandTap &rnd)   // read_realignment.cpp:212-293
	{
		Graph &g = r.g[s];
		const int n = (int)g.path.size();
		if (n == 0) return 0;
		for (;;) {
			g.max_index = U32MAX; g.max_distance = 0;
			g.same_top.clear(); g.same_top.push_back((int)g.max_index);
			for (int i = n - 1; i >= 0; --i) {
				if (g.path[i].used) continue;
				const float d = g.path[i].dist;
				if (g.max_distance < d) { g.max_distance = d; g.max_index = (uint32_t)i; g.same_top.clear(); g.same_top.push_back(i); }
				else if (g.max_distance == d) g.same_top.push_back(i);
			}
			if (g.max_index == U32MAX) return 0;
			int used = 0, fresh = 0;
			const uint32_t same = (uint32_t)g.same_top.size();
			if (same > 1) g.max_index = (uint32_t)g.same_top[rnd.draw((int32_t)same)];
			int node = (int)g.max_index;
			const int first = node;
			for (; node != -1;) {
				if (g.path[node].used) ++used; else ++fresh;
				g.path[node].used = 1;
				const int next = g.path[node].pre_node;
				if (next == -1) break;
				node = next;
			}
			const int last = node;
			if (first - last > ((fresh + used + 5) << 1))
				for (int k = last; k < first; ++k) g.path[k].used = 1;
			if (used >= fresh) continue;                           // tail recursion in the reference
			const int ref_begin = (int)g.v[node].ref_begin;
			const int chr = idx.chromosome_id((uint32_t)ref_begin);
			rst.direction = direction;
			rst.max_index = g.max_index;
			rst.chain_score = (uint32_t)g.max_distance;
			rst.read_bg = g.v[node].read_begin;
			rst.chr = (uint32_t)chr;
			rst.ref_bg = (uint32_t)(ref_begin - (int)idx.chr_end_before(chr));
			return 1;
		}
	}

	static int cmp_chain(const void *a, const void *b)           // cmp_chain_score, read_realignment.hpp:303-308 (returns 0/1, sic)
	{
		const Result *x = *(Result* const*)a, *y = *(Result* const*)b;
		if (x->chain_score != y->chain_score) return x->chain_score < y->chain_score;
		return x->max_index > y->max_index;
	}
	static int cmp_align(const void *a, const void *b)           // cmp_align_score, read_realignment.hpp:310-315
	{
		const Result *x = *(Result* const*)a, *y = *(Result* const*)b;
		if (x->align_score != y->align_score) return x->align_score < y->align_score;
		return x->max_index > y->max_index;
	}
	// qsort of an array of large structs: glibc sorts pointers and permutes afterwards, same comparison order
	void sort_results(Result *res, int n, int (*cmp)(const void*, const void*))
	{
		if (n < 2) return;
		Result *p[2 * MAX_OUTPUT_NUMBER];
		for (int i = 0; i < n; ++i) p[i] = res + i;
		qsort(p, n, sizeof(Result*), cmp);
		bool moved = false;
		for (int i = 0; i < n; ++i) moved |= p[i] != res + i;
		if (!moved) return;
		Result tmp[2 * MAX_OUTPUT_NUMBER];
		for (int i = 0; i < n; ++i) tmp[i] = std::move(*p[i]);
		for (int i = 0; i < n; ++i) res[i] = std::move(tmp[i]);
	}

	bool reverse_cigar(std::vector<CigarPath> &out, const std::vector<CigarPath> &tmp, int read_len)   // reverseGIGAR, read_realignment.hpp:277-301
	{
		out.clear();
		if (tmp.empty()) return false;
		out.reserve(tmp.size());
		out.push_back(tmp.back());
		for (int i = (int)tmp.size() - 2; i >= 0; --i)
			if (!cig_try_merge(out.back(), tmp[i])) out.push_back(tmp[i]);
		if (!out.empty() && out[0].size == 0) out.erase(out.begin());
		int total = 0;
		for (const CigarPath &ci : out) if (ci.type == 0 || ci.type == 1 || ci.type == 3 || ci.type == 4) total += ci.size;
		return total == read_len;
	}

	// the rest of single_end_handler::align once chains and ksw results exist (RR:416-475)
	void finish_read(ReadState &r, const KswView &tasks, RandTap &rnd)
	{
		r.result_num = 0; r.primary = r.secondary = nullptr;
		r.result.clear();
		if (r.skip) return;
		for (int s = 0; s < 2; ++s) for (PathNode &p : r.g[s].path) p.used = 0;   // a probed pair is finished a second time
		r.result.reserve(2 * MAX_OUTPUT_NUMBER);
		uint32_t max_chain = 0;
		for (int s = 0; s < 2; ++s) {
			const int direction = s == 0 ? FORWARD : REVERSE;
			for (int i = 0; i < MAX_OUTPUT_NUMBER; ++i) {
				Result slot;
				if (!sort_output(r, s, slot, direction, rnd)) break;
				const uint32_t c = slot.chain_score;
				max_chain = std::max(c, max_chain);
				if (c + MAX_CHAIN_SCORE_DIFF < max_chain || c < MIN_CHAIN_SCORE2) break;
				r.result.push_back(slot);
				++r.result_num;
			}
		}
		sort_results(r.result.data(), r.result_num, cmp_chain);
		if (r.result_num == 0 || max_chain < MIN_CHAIN_SCORE) return;
		for (int k = 0; k < r.result_num; ++k) {
			Result &c = r.result[k];
			if (c.chain_score + MAX_CHAIN_SCORE_DIFF < max_chain) { r.result_num = k; break; }
			const int s = c.direction == REVERSE ? 1 : 0;
			const uint64_t key = (uint64_t)s << 32 | c.max_index;
			auto it = std::lower_bound(r.node_aln.begin(), r.node_aln.end(), key, [](const std::pair<uint64_t, NodeAln> &a, uint64_t k) { return a.first < k; });
			if (it == r.node_aln.end() || it->first != key || !it->second.planned) {     // cannot happen: the plan is a superset
				fprintf(stderr, "pansvr_b200: internal error, chain end %u of strand %d was not planned\n", c.max_index, s);
				abort();
			}
			NodeAln &na = it->second;
			if (!na.resolved) {
			int score = na.fixed_score;
			std::vector<CigarPath> &tmp = rnd.cigar_scratch;
			tmp.clear();
			for (const Piece &p : na.pieces) {
				if (p.kind == 0) { tmp.push_back(p.lit); continue; }
				const int32_t *res = tasks.res + (size_t)p.task * PANSVR_RES_WORDS;
				const uint32_t *cg = tasks.cig + (size_t)p.task * tasks.cap;
				const int nc = res[PANSVR_RES_N_CIGAR];
				if (p.type == ALN_E2E) { score += res[PANSVR_RES_SCORE]; for (int i = nc - 1; i >= 0; --i) tmp.push_back(cig_bin(cg[i])); }
				else if (p.type == ALN_LEFT) { score += res[PANSVR_RES_MQE]; for (int i = 0; i < nc; ++i) tmp.push_back(cig_bin(cg[i])); }
				else { score += res[PANSVR_RES_MQE]; for (int i = nc - 1; i >= 0; --i) tmp.push_back(cig_bin(cg[i])); }
			}
			na.align_score = (uint32_t)std::max(score, 0);
			na.cigar_ok = reverse_cigar(na.cigar, tmp, r.read_l);
			na.resolved = true;
			}
			c.ref_bg -= (uint32_t)na.read_begin_alignment;
			c.align_score = na.align_score;
			c.cigar = na.cigar;
			c.cigar_ok = na.cigar_ok;
		}
		sort_results(r.result.data(), r.result_num, cmp_align);
		if (r.result[0].align_score < (uint32_t)MIN_ALN_SCORE) { r.result_num = 0; return; }
		for (int i = 0; i < r.result_num; ++i) {
			Result &c = r.result[i];
			const uint32_t sv_id = c.chr;
			c.sv = &idx.sv_info[sv_id];
			c.chr = c.sv->chr_id;
			c.ref_bg += (uint32_t)c.sv->st_pos;
			if (c.ref_bg >= (uint32_t)I32MAX) c.ref_bg = 5;
			c.is_ori = false; c.rst_idx = i; c.mapq = 0; c.has_mate = false;
		}
		if (r.result_num > 0) {
			const int32_t d = (int32_t)(r.result[0].align_score - (r.result_num > 1 ? r.result[1].align_score : 0));
			r.result[0].mapq = (uint8_t)(d > 40 ? 40 : d);
		}
	}

	// ---- PE_score (read_realignment.hpp:434-628)
	struct PE {
		int max_same = 1, max_score = 0, cur_isize = 0;
		bool proper = false, gain = false;
		Result *m1 = nullptr, *m2 = nullptr;
	};
	int get_isize(int p1, int p2, int d1, int d2) const
	{
		if (d1 == d2) return 0;
		const int max_isize = P.opt.isize_max + 200, min_isize = std::max(0, P.opt.isize_min - 200);
		const int isize = P.opt.read_len + (d1 == FORWARD ? p2 - p1 : p1 - p2);
		return (isize < max_isize && isize > min_isize) ? isize : 0;
	}
	int proper_mated(const Result *a, const Result *b) const
	{
		if (!a || !b || a->chr != b->chr) return 0;
		const int a1 = (int)a->ref_bg, a2 = a1 + (a->is_ori ? 0 : a->sv->end_offset);
		const int b1 = (int)b->ref_bg, b2 = b1 + (b->is_ori ? 0 : b->sv->end_offset);
		int v;
		if ((v = get_isize(a1, b1, a->direction, b->direction)) > 0) return v;
		if ((v = get_isize(a1, b2, a->direction, b->direction)) > 0) return v;
		if ((v = get_isize(a2, b1, a->direction, b->direction)) > 0) return v;
		if ((v = get_isize(a2, b2, a->direction, b->direction)) > 0) return v;
		return 0;
	}
	void store_pair(PE &pe, Result *a, Result *b, RandTap &rnd, std::vector<PairEvent> *ev = nullptr, int ci = -1, int cj = -1)
	{
		const int isize = proper_mated(a, b);
		const int basic = (a ? (int)a->align_score : 0) + (b ? (int)b->align_score : 0);
		const bool one_new = (a && !a->is_ori) || (b && !b->is_ori);
		const int fin = basic + (isize > 0 ? 0 : -60) + (one_new ? 0 : 1);
		if (fin >= pe.max_score) {
			bool store = true;
			if (ev) { PairEvent e; e.i = (int8_t)ci; e.j = (int8_t)cj; e.tie = fin == pe.max_score; ev->push_back(e); }
			if (fin > pe.max_score) pe.max_same = 1;
			else if (fin == pe.max_score) { ++pe.max_same; if (rnd.draw(pe.max_same) != 0) store = false; }
			if (store) { pe.m1 = a; pe.m2 = b; pe.max_score = fin; pe.cur_isize = isize; pe.proper = isize > 0; }
		}
	}
	static Result *pick(ReadState &r, int i) { return i < 0 ? nullptr : (i < r.result_num ? &r.result[i] : &r.ori); }
	// `ev`: also record every candidate combination that reaches `fin >= max_score`, in order.  Which of them wins the ties
	// is the only thing the random stream decides (the running maximum does not depend on it), so the replay can redraw
	// a pair's ties from the events alone (replay_pairing) without touching the candidates.
	void pair_up(ReadState *se, PE &pe, RandTap &rnd, std::vector<PairEvent> *ev = nullptr)
	{
		pe = PE();
		int n0 = se[0].result_num, n1 = se[1].result_num;
		if (!se[0].ori_unmapped) ++n0;
		if (!se[1].ori_unmapped) ++n1;
		for (int i = 0; i < n0; ++i) store_pair(pe, pick(se[0], i), nullptr, rnd, ev, i, -1);
		for (int j = 0; j < n1; ++j) store_pair(pe, nullptr, pick(se[1], j), rnd, ev, -1, j);
		for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) store_pair(pe, pick(se[0], i), pick(se[1], j), rnd, ev, i, j);
		pe.gain = pe.max_score > 0 && ((pe.m1 && !pe.m1->is_ori) || (pe.m2 && !pe.m2->is_ori));
	}
	// the pairing decided by the winner (i, j) of the events; max_score is what pair_up already found
	void apply_pairing(ReadState *se, PE &pe, int i, int j)
	{
		pe.m1 = pick(se[0], i); pe.m2 = pick(se[1], j);
		pe.cur_isize = proper_mated(pe.m1, pe.m2);
		pe.proper = pe.cur_isize > 0;
		pe.gain = pe.max_score > 0 && ((pe.m1 && !pe.m1->is_ori) || (pe.m2 && !pe.m2->is_ori));
	}
	// what finish_read left in r, as far as anything downstream can see it
	static void result_signature(const ReadState &r, std::vector<uint32_t> &sig)
	{
		sig.clear();
		sig.push_back((uint32_t)r.result_num);
		for (int k = 0; k < r.result_num; ++k) {
			const Result &c = r.result[k];
			sig.push_back(c.align_score); sig.push_back(c.chain_score); sig.push_back(c.max_index); sig.push_back(c.read_bg);
			sig.push_back(c.chr); sig.push_back(c.ref_bg); sig.push_back((uint32_t)c.direction); sig.push_back(c.mapq);
			sig.push_back((uint32_t)(c.sv ? c.sv->id : 0xffffffffu)); sig.push_back((uint32_t)c.cigar.size());
			for (const CigarPath &ci : c.cigar) sig.push_back((uint32_t)ci.type << 16 | (uint16_t)ci.size);
		}
	}
	// A read whose candidate list had ties: run finish_read for every combination of tie outcomes.  If all of them leave the
	// same candidates and make the same number of draws, the read does not depend on the stream at all -- the replay only
	// has to advance the stream by that number.  Returns the number of draws, or -1 (too many combinations, or they differ).
	int explore_read(ReadState &r, const KswView &tasks, RandTap &probe, std::vector<uint32_t> &sig0, std::vector<uint32_t> &sig)
	{
		enum { MAX_LEAVES = 24 };
		probe.restart(0);
		finish_read(r, tasks, probe);
		if (probe.too_deep) return -1;
		const uint32_t c0 = probe.calls;
		if (c0 == 0) return 0;
		result_signature(r, sig0);
		uint8_t choice[RandTap::MAX_SCRIPT] = {0}, mod[RandTap::MAX_SCRIPT];
		memcpy(mod, probe.moduli, c0);
		uint32_t c = c0;
		for (int leaves = 1; ; ++leaves) {
			int p = (int)c - 1;                                   // next combination: odometer over the moduli of the last run
			while (p >= 0 && choice[p] + 1 >= mod[p]) --p;
			if (p < 0) break;
			if (leaves >= MAX_LEAVES) return -1;
			++choice[p];
			for (uint32_t k = (uint32_t)p + 1; k < RandTap::MAX_SCRIPT; ++k) choice[k] = 0;
			probe.restart((uint32_t)p + 1);
			memcpy(probe.script, choice, (size_t)p + 1);
			finish_read(r, tasks, probe);
			if (probe.too_deep || probe.calls != c0) return -1;
			c = probe.calls;
			memcpy(mod, probe.moduli, c);
			result_signature(r, sig);
			if (sig != sig0) return -1;
		}
		return (int)c0;
	}
	void set_primary(ReadState *se, PE &pe)                      // set_primary_secondary_mate, RRH:501-534
	{
		for (int i = 0; i < 2; ++i) {
			Result *c = i == 0 ? pe.m1 : pe.m2;
			if (!c) continue;
			ReadState &h = se[i];
			h.primary = c; h.secondary = nullptr;
			if (c->is_ori && h.result_num > 0) h.secondary = &h.result[0];
			else if (h.result_num > 1) h.secondary = c->rst_idx == 0 ? &h.result[1] : &h.result[0];
			Result *m = i == 0 ? pe.m2 : pe.m1;
			if (m && m->chr != U32MAX) {
				c->has_mate = true; c->mate_chr = m->chr; c->mate_ref_bg = m->ref_bg; c->mate_sv = m->sv;
				if (c->is_ori) c->sv = c->mate_sv;
			} else { c->has_mate = false; c->mate_chr = 0; c->mate_sv = nullptr; }
		}
	}

	// SAM text of one record after htslib's parse->format round trip (RR:479-536; sam.c sam_parse1 / sam_format1)
	static void append_int(std::string &s, long v)
	{
		char b[24]; int n = 24;
		unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
		do { b[--n] = (char)('0' + u % 10); u /= 10; } while (u);
		if (v < 0) b[--n] = '-';
		s.append(b + n, 24 - n);
	}
	// SEQ and QUAL columns: the reference reverses its kseq_t in place around the write and restores it afterwards
	void append_seq_qual(std::string &out, const ReadState &r, bool reversed)
	{
		const size_t at = out.size();
		out.append(r.rec->seq, r.rec->seq_l); out += '\t'; out.append(r.rec->qual, r.rec->qual_l);
		if (r.seq_twice_reversed) for (size_t i = at, e = at + r.rec->seq_l; i < e; ++i) out[i] = rev_char(rev_char(out[i]));
		if (reversed) { rev_str(&out[at], r.read_l); rev_qual(&out[at + r.rec->seq_l + 1], r.read_l); }
		for (size_t i = at, e = at + r.rec->seq_l; i < e; ++i) out[i] = g_nt16_norm.t[(unsigned char)out[i]];   // 4-bit round trip of SEQ
	}
	std::string target_name(uint32_t id) const { return id < idx.target_names.size() ? idx.target_names[id] : std::string("*"); }

	void output_bam(ReadState &r, std::string &out, bool first, int abs_isize)
	{
		Result *p = r.primary;                                   // appends one line (with its newline) to `out`, or nothing
		if (!p || p->chr == U32MAX) return;
		if (P.opt.not_ori && p->is_ori) return;
		// A z-dropped extension (only with -z well below the default) can leave a CIGAR shorter than the read.  The reference
		// logs "ERROR cigar", htslib rejects the record and the reference then writes the half-parsed bam1_t with stale buffer
		// bytes; there is nothing defined to reproduce, so the record is left out and counted.
		if (!p->is_ori && !p->cigar_ok) { if (count_bad) ++P.bad_cigar_records_; if (p->direction == REVERSE) r.seq_twice_reversed = true; return; }
		const int dir = p->direction;
		const uint8_t flag = (uint8_t)((first ? 0x40 : 0) + (dir == REVERSE ? 0x10 : 0) + (p->has_mate ? 0 : 0x08));
		out.append(r.rec->name, r.rec->name_l); out += '\t'; append_int(out, flag); out += '\t';
		out += target_name(p->chr); out += '\t'; append_int(out, (int)p->ref_bg); out += '\t'; append_int(out, p->mapq); out += '\t';
		if (p->cigar.empty()) out += '*';
		for (const CigarPath &c : p->cigar) { append_int(out, c.size); out += "MIDNSHP=XB"[c.type]; }
		out += '\t';
		const int isize = dir == FORWARD ? abs_isize : -abs_isize;
		if (p->has_mate) {
			out += (p->mate_chr == p->chr) ? std::string("=") : target_name(p->mate_chr);
			out += '\t'; append_int(out, (int)p->mate_ref_bg); out += '\t'; append_int(out, isize); out += '\t';
		} else out += "*\t0\t0\t";
		append_seq_qual(out, r, dir == REVERSE); out += '\t';
		if (dir == REVERSE) r.seq_twice_reversed = true;
		out += "AS:i:"; append_int(out, (int)p->align_score);
		out += "\tOS:i:"; append_int(out, (int)r.ori.align_score);
		out += "\tOA:Z:"; append_int(out, (int)r.ori.chr); out += ','; append_int(out, (int)r.ori.ref_bg); out += ',';
		append_int(out, (int)r.ori.read_bg); out += ','; append_int(out, r.ori.mapq); out += ','; out += r.ori_unmapped ? 'U' : 'M'; out += ';';
		if (!p->is_ori) { out += "\tCS:i:"; append_int(out, (int)p->chain_score); }
		if (p->sv) { out += "\tSV:Z:"; out += p->sv->vcf_print; }
		if (p->mate_sv) { out += "\tMV:Z:"; out += p->mate_sv->vcf_print; }
		if (r.secondary) {
			const Result *s = r.secondary;
			out += "\tXA:Z:"; append_int(out, (int)s->chr); out += ','; append_int(out, (int)s->ref_bg); out += ','; append_int(out, (int)s->read_bg);
			out += ','; append_int(out, (int)s->align_score); out += ','; out += s->direction == FORWARD ? 'F' : 'R'; out += ',';
			out += s->sv ? s->sv->vcf_id : std::string("*"); out += ';';
		}
		out += "\tRC:Z:"; out += r.comment;
		out += '\n';
	}

	// output_ori_bam (RR:656-717): the original alignment rebuilt from the comment; returns whether the original CIGAR shows a clip >= 25 / unmapped
	void output_ori(ReadState &r, std::string &out, int max_score, bool &clip_or_unmapped)
	{
		clip_or_unmapped = true;                                 // appends one line (with its newline) to `out`, or nothing
		std::string &c = r.comment;
		const size_t fpos = c.find("FLAG_");
		if (fpos == std::string::npos) return;
		unsigned flag = 0, qual = 0;
		sscanf(c.c_str() + fpos + 5, "%u_%u_", &flag, &qual);
		const size_t cpos = c.find("CIGAR_", fpos + 5);
		if (cpos == std::string::npos) return;
		const size_t cig_b = cpos + 6;
		size_t cig_e = cig_b;
		while (cig_e < c.size() && c[cig_e] != '_') ++cig_e;
		const std::string cigar = c.substr(cig_b, cig_e - cig_b);
		const size_t mate_b = cig_e + 1 + 5;
		int mchr = 0, mpos = 0, isize = 0;
		if (mate_b < c.size()) sscanf(c.c_str() + mate_b, "%d_%d_%d_", &mchr, &mpos, &isize);
		mpos += 1;
		std::string tags;
		const size_t tpos = c.find("TAG_", mate_b < c.size() ? mate_b : c.size());
		if (tpos != std::string::npos) tags = c.substr(tpos + 4);
		const int tag_len = (int)tags.size();
		for (int i = 0; i < tag_len - 5; ++i) if (tags[i] == '_' && tags[i + 3] == ':' && tags[i + 5] == ':') tags[i] = '\t';
		if (tag_len > 0) tags.resize(tag_len - 1);
		// the reference cuts the comment in place at the end of the CIGAR and edits the tags
		c[cig_e < c.size() ? cig_e : c.size() - 1] = '\0';
		out.append(r.rec->name, r.rec->name_l); out += '\t'; append_int(out, flag); out += '\t'; out += target_name(r.ori.chr); out += '\t';
		append_int(out, (long)(r.ori.ref_bg + 1)); out += '\t'; append_int(out, qual); out += '\t';
		out += cigar.empty() ? std::string("*") : cigar; out += '\t';
		out += ((uint32_t)mchr == r.ori.chr) ? std::string("=") : target_name((uint32_t)mchr);
		out += '\t'; append_int(out, mpos); out += '\t'; append_int(out, isize); out += '\t';
		const bool fwd = (flag & 0x10) == 0;
		append_seq_qual(out, r, !fwd);
		if (tag_len) { out += '\t'; out += tags; }
		out += "\tMS:i:"; append_int(out, max_score);
		out += '\n';
		// bam_has_clip_or_unmapped_ori (RR:721-733) on the CIGAR just written
		if (cigar.empty()) { clip_or_unmapped = true; return; }
		std::vector<std::pair<int, char>> ops;
		for (size_t i = 0; i < cigar.size();) {
			int len = 0;
			while (i < cigar.size() && cigar[i] >= '0' && cigar[i] <= '9') len = len * 10 + (cigar[i++] - '0');
			if (i < cigar.size()) ops.push_back(std::make_pair(len, cigar[i++]));
		}
		int clip = 0;
		if (!ops.empty()) {
			if (ops.front().second == 'S' || ops.front().second == 'H') clip += ops.front().first;
			if (ops.back().second == 'S' || ops.back().second == 'H') clip += ops.back().first;
		}
		clip_or_unmapped = ops.empty() || clip >= 25;
	}

	// the `-p` records of a pair (RR:776-797): both originals, unless the pair turned out proper after all
	void output_ori_pair(ReadState *se, const PE &pe, std::string &, std::string &ori, int min_filter_score)
	{
		if (!(pe.max_score <= min_filter_score && (int)se[0].ori.chr != -1 && (int)se[1].ori.chr != -1)) return;
		bool clip[2] = {true, true};
		const size_t ori_mark = ori.size();
		for (int k = 0; k < 2; ++k) output_ori(se[k], ori, pe.max_score, clip[k]);
		bool proper = pe.proper;
		for (int k = 0; proper && k < 2; ++k) {
			const Result *c = k == 0 ? pe.m1 : pe.m2;
			if (!c) { proper = false; break; }
			if (c->is_ori && clip[k]) proper = false;
			if (proper && !c->is_ori) {
				int ins = 0;
				for (const CigarPath &ci : c->cigar) if (ci.type == 1) ins += ci.size;
				if (c->cigar.empty() || ins >= 25) proper = false;
			}
		}
		if (proper) ori.resize(ori_mark);                              // a proper pair after all: nothing goes to the -p file
	}
};

// ================================================================================================ public
AlnPipeline::AlnPipeline(const DebgaIndex &idx, const AlnOptions &o, SeedService *seeds, void *ksw_ctx, StageService *stages, int device)
	: opt(o), idx_(idx), seeds_(seeds), ksw_(ksw_ctx), stages_(stages), device_(device), rand_(1)
{
	if (opt.threads > 1) workers_ = new Workers(opt.threads);
	if (getenv("PANSVR_HOST_STAGES")) stages_ = nullptr;          // differential runs: stages A, C, D on the host (the round-1 path)
	if (stages_) {
		AlnScores sc{opt.match, opt.mismatch, opt.gap_open, opt.gap_ex, opt.gap_open2, opt.gap_ex2};
		stage_service_set_scoring(stages_, sc, opt.zdrop);
		DevBuffers *b = new DevBuffers();
		b->svc = stages_;
		dev_all_.push_back(b); dev_free_.push_back(b);
	}
	sites_.push_back(DevSite{device_, seeds_});
	host_sv_.resize(idx_.sv_info.size() * sizeof(DevSv) + sizeof(DevSv));
	for (size_t i = 0; i < idx_.sv_info.size(); ++i) {
		DevSv v; v.chr_id = idx_.sv_info[i].chr_id; v.st_pos = (uint32_t)idx_.sv_info[i].st_pos; v.end_offset = idx_.sv_info[i].end_offset; v.pad = 0;
		memcpy(host_sv_.data() + i * sizeof(DevSv), &v, sizeof v);
	}
	if (const char *e = getenv("PANSVR_TRIP1")) { const int v = atoi(e); if (v > 0) trip1_cap_ = v; }
	reset();
}

AlnPipeline::~AlnPipeline()
{
	delete workers_;
	for (DevBuffers *b : dev_all_) { if (b->owned) stage_service_destroy(b->svc); delete b; }
	for (HostSlot *h : host_all_) delete h;
}

AlnPipeline::DevBuffers *AlnPipeline::acquire_dev(std::string &err, uint64_t seq)
{
	const size_t site = (size_t)(seq % sites_.size());               // sub-blocks are dealt to the GPUs round robin
	{
		std::lock_guard<std::mutex> lk(dev_pool_m_);
		for (size_t k = 0; k < dev_free_.size(); ++k) if (dev_free_[k]->site == site) { DevBuffers *b = dev_free_[k]; dev_free_.erase(dev_free_.begin() + (long)k); return b; }
	}
	StageService *svc = stage_service_create(idx_, sites_[site].seeds, ksw_, sites_[site].device, err);   // another block in flight: its own device state
	if (!svc) return nullptr;
	AlnScores sc{opt.match, opt.mismatch, opt.gap_open, opt.gap_ex, opt.gap_open2, opt.gap_ex2};
	stage_service_set_scoring(svc, sc, opt.zdrop);
	DevBuffers *b = new DevBuffers();
	b->svc = svc; b->owned = true; b->site = site;
	std::lock_guard<std::mutex> lk(dev_pool_m_);
	dev_all_.push_back(b);
	return b;
}

void AlnPipeline::release_dev(DevBuffers *b) { std::lock_guard<std::mutex> lk(dev_pool_m_); dev_free_.push_back(b); }

AlnPipeline::HostSlot *AlnPipeline::acquire_host()
{
	std::lock_guard<std::mutex> lk(dev_pool_m_);
	if (!host_free_.empty()) { HostSlot *h = host_free_.back(); host_free_.pop_back(); return h; }
	HostSlot *h = new HostSlot();
	host_all_.push_back(h);
	return h;
}

void AlnPipeline::release_host(HostSlot *h) { std::lock_guard<std::mutex> lk(dev_pool_m_); host_free_.push_back(h); }

// read statistics from the first comment of the input (load_reads, RR:134-148); must have run before two blocks are in flight
void AlnPipeline::ensure_read_stats(const FastqRec &first)
{
	if (opt.stat_set) return;
	const std::string c0(first.comment, first.comment_l);
	const char *st = strstr(c0.c_str(), "STAT_");
	if (!st || sscanf(st + 5, "%d_%d_%d_%d_", &opt.read_len, &opt.isize_min, &opt.isize_mid, &opt.isize_max) == -1) {
		opt.read_len = 150; opt.isize_min = 100; opt.isize_mid = 500; opt.isize_max = 900;
	}
	min_filter_score_ = std::max(opt.read_len * opt.match * 2 - 80, 50);
	opt.stat_set = true;
}

AlnPipeline::StreamState AlnPipeline::export_streams()
{
	std::unique_lock<std::mutex> lk(turn_m_);
	turn_cv_.wait(lk, [&]() { return replay_turn_ == seq_issued_; });
	StreamState s;
	memset((void*)&s, 0, sizeof s);
	s.magic = 0x70535652u;
	s.rand = rand_; s.rand_r[0] = rand_r_[0]; s.rand_r[1] = rand_r_[1];
	return s;
}

void AlnPipeline::import_streams(const StreamState &s) { rand_ = s.rand; rand_r_[0] = s.rand_r[0]; rand_r_[1] = s.rand_r[1]; }

void AlnPipeline::await_streams(const std::string &path) { std::lock_guard<std::mutex> lk(turn_m_); await_path_ = path; }

void AlnPipeline::chain_at(uint64_t seq, const char *await, const char *publish)
{
	std::lock_guard<std::mutex> lk(turn_m_);
	chain_[seq] = std::make_pair(std::string(await ? await : ""), std::string(publish ? publish : ""));
}

bool AlnPipeline::publish_streams(const std::string &path) { return write_streams(path, export_streams()); }

void AlnPipeline::pass_turn_if_pending(uint64_t seq)
{
	std::unique_lock<std::mutex> lk(turn_m_);
	turn_cv_.wait(lk, [&]() { return replay_turn_ >= seq; });
	if (replay_turn_ != seq) return;                                     // it had its turn
	replay_turn_ = seq + 1;
	chain_.erase(seq);
	lk.unlock();
	turn_cv_.notify_all();
}

bool AlnPipeline::write_streams(const std::string &path, const StreamState &s)
{
	const std::string tmp = path + ".part";
	FILE *f = fopen(tmp.c_str(), "wb");
	if (!f) return false;
	const bool ok = fwrite(&s, sizeof s, 1, f) == 1;
	if (fclose(f) != 0 || !ok) { remove(tmp.c_str()); return false; }
	return rename(tmp.c_str(), path.c_str()) == 0;
}

void AlnPipeline::reset()
{
	replay_turn_ = 0; seq_issued_ = 0;
	await_path_.clear(); chain_.clear();
	bad_cigar_records_ = 0;
	rand_.reseed(1);
	stats = Stats();
	opt.stat_set = false;
	// Classify_buff_pool of thread 0: two single_end_handlers, each seeds its random_r state with rand() (RR:62-67, RRH:339-340)
	for (int i = 0; i < 2; ++i) rand_r_[i].reseed((unsigned)rand_.next());
}

// The host path: every stage but seeding and ksw on the helper threads (round 1's pipeline).  It finishes what the device path
// leaves to it -- pairs with 'N' or a lower-case 'n', pairs whose unipaths need random_r sampling, pairs whose rand() ties change
// the outcome -- and, without a stage service, whole blocks.  With hooks, `recs` are the pairs left to it out of a larger block:
// the in-order pass and the record text are merged with the device path's pairs in input order.
struct BlockHooks {
	const uint32_t *global_pair = nullptr;                            // pair k of `recs` is pair global_pair[k] of the whole block
	std::function<void(uint64_t)> fast_until;                         // in-order work of the device path's pairs below this block index
	std::function<bool(const std::function<void(size_t, std::string&, std::string&)>&, std::string&)> emit;   // record text of the whole block
};

bool AlnPipeline::align_block_host(const FastqRec *recs, size_t n_reads_in, BlockOutput &out, std::string &err, uint64_t seq, BlockHooks *H)
{
	Impl I(*this);
	const size_t n_reads = n_reads_in & ~(size_t)1, n_pairs = n_reads / 2;
	// in-order sections: wait until every earlier block has finished its replay; leave by passing the turn on
	double t_turn = -1;                                                   // when this block's in-order section began
	auto wait_turn = [&]() {
		std::unique_lock<std::mutex> lk(turn_m_);
		turn_cv_.wait(lk, [&]() { return replay_turn_ == seq; });
		struct Mark { double &t; ~Mark() { if (t < 0) t = now(); } } mark{t_turn};
		// this process continues another one's input (await_streams, chain_at): take the random streams where that one left them
		std::string path;
		if (!await_path_.empty()) { path = await_path_; await_path_.clear(); }
		auto ch = chain_.find(seq);
		if (ch != chain_.end() && !ch->second.first.empty()) { path = ch->second.first; ch->second.first.clear(); }
		if (path.empty()) return;
		lk.unlock();
		for (uint64_t spins = 0;; ++spins) {
			if (spins == 6000000) { fprintf(stderr, "pansvr_b200: still no stream state at %s after 10 minutes (did the process before this one fail?)\n", path.c_str()); abort(); }
			if (FILE *f = fopen(path.c_str(), "rb")) {
				StreamState s;
				const bool ok = fread(&s, sizeof s, 1, f) == 1 && s.magic == 0x70535652u;
				fclose(f);
				if (ok) { import_streams(s); remove(path.c_str()); return; }
			}
			std::this_thread::sleep_for(std::chrono::microseconds(100));
		}
	};
	struct TurnGuard {                                                    // whatever happens, the next block must not wait for ever
		AlnPipeline &P; uint64_t seq; bool passed = false;
		void pass()
		{
			if (passed) return;
			passed = true;
			std::string pub;
			{ std::lock_guard<std::mutex> lk(P.turn_m_); auto ch = P.chain_.find(seq); if (ch != P.chain_.end()) { pub = ch->second.second; P.chain_.erase(ch); } }
			if (!pub.empty()) {                                               // the process that has the next piece of the input goes on from here
				StreamState s;
				memset((void*)&s, 0, sizeof s);
				s.magic = 0x70535652u; s.rand = P.rand_; s.rand_r[0] = P.rand_r_[0]; s.rand_r[1] = P.rand_r_[1];
				if (!P.write_streams(pub, s)) { fprintf(stderr, "pansvr_b200: cannot write the stream state to %s\n", pub.c_str()); abort(); }
			}
			{ std::lock_guard<std::mutex> lk(P.turn_m_); if (P.replay_turn_ == seq) P.replay_turn_ = seq + 1; }
			P.turn_cv_.notify_all();
		}
		~TurnGuard() { if (!passed) { std::unique_lock<std::mutex> lk(P.turn_m_); P.turn_cv_.wait(lk, [&]() { return P.replay_turn_ == seq; }); lk.unlock(); pass(); } }
	} turn{*this, seq};
	auto add_time = [&](int stage, double dt) { std::lock_guard<std::mutex> lk(stats_m_); stats.t_stage[stage] += dt; };
	out.sam.resize((size_t)std::max(1, opt.threads));                     // the caller may keep `out` across blocks: capacity is reused
	out.ori.resize((size_t)std::max(1, opt.threads));
	for (std::string &x : out.sam) x.clear();
	for (std::string &x : out.ori) x.clear();
	if (n_pairs == 0 && !H) {
		bool chained;
		{ std::lock_guard<std::mutex> lk(turn_m_); chained = chain_.count(seq) != 0; }
		if (chained) wait_turn();                                          // an empty piece still takes the streams over and hands them on
		return true;
	}
	double t0 = now();

	if (n_pairs) ensure_read_stats(recs[0]);

	// ---- stage A (parallel over reads; reads of a pair with an 'N' are left for the replay, see below)
	const int T = std::max(1, opt.threads);
	// per-read state is constructed and destroyed by the worker threads (it is ~0.5 KB of containers per read)
	// loops over reads are cut at pair boundaries: the same worker owns a pair's two reads in every stage
	auto par_reads = [&](const std::function<void(size_t, size_t, int)> &fn) { parallel(n_pairs, [&](size_t b, size_t e, int t) { fn(2 * b, 2 * e, t); }); };
	// 'N' census: a read with 1..3 N gets 4^n variant states behind the real reads (indices n_reads ..)
	const bool timing = getenv("PANSVR_TIMING") != nullptr;
	double ta = now();
	auto lap = [&](const char *what) { if (timing) { const double t = now(); fprintf(stderr, "[timing]   A/%s %.3f s\n", what, t - ta); ta = t; } };
	enum { MAX_VARIANT_DRAWS = 3 };
	std::vector<uint8_t> n_count(n_reads, 0), lower_n(n_reads, 0);
	par_reads([&](size_t b, size_t e, int) {
		for (size_t i = b; i < e; ++i) {
			uint32_t c = 0;
			const char *q = recs[i].seq, *qe = q + recs[i].seq_l;
			while ((q = (const char*)memchr(q, 'N', (size_t)(qe - q))) != nullptr) { ++c; ++q; }
			n_count[i] = (uint8_t)std::min<uint32_t>(c, 255);
			lower_n[i] = memchr(recs[i].seq, 'n', recs[i].seq_l) != nullptr;     // code 4 spills into the packed neighbour: such a read stays on the host
		}
	});
	std::vector<std::pair<size_t, size_t>> var_src;                       // (real read, first variant index) of reads with variants
	size_t n_var = 0;
	for (size_t i = 0; i < n_reads; ++i)
		if (n_count[i] >= 1 && n_count[i] <= MAX_VARIANT_DRAWS && recs[i].seq_l >= LEN_KMER) { var_src.push_back(std::make_pair(i, n_reads + n_var)); n_var += (size_t)1 << (2 * n_count[i]); }
	const size_t n_all = n_reads + n_var;
	lap("N census");
	// every loop over read states: the real reads in pair chunks, then the variants
	auto par_all = [&](const std::function<void(size_t, size_t, int)> &fn) {
		par_reads(fn);
		if (n_var) parallel(n_var, [&](size_t b, size_t e, int t) { fn(n_reads + b, n_reads + e, t); });
	};
	struct ReadArray {
		ReadState *p; size_t n; decltype(par_all) &par;
		ReadArray(size_t n_, decltype(par_all) &par_) : p((ReadState*)malloc(sizeof(ReadState) * std::max<size_t>(n_, 1))), n(n_), par(par_)
		{ par([&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) new (p + i) ReadState(); }); }
		~ReadArray() { par([&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) p[i].~ReadState(); }); free(p); }
		ReadState &operator[](size_t i) { return p[i]; }
	} rs(n_all, par_all);
	lap("state array");
	par_reads([&](size_t b, size_t e, int) {
		for (size_t i = b; i < e; ++i) {
			ReadState &r = rs[i];
			r.rec = &recs[i];
			r.comment.assign(recs[i].comment, recs[i].comment_l);
			r.read_l = (int)recs[i].seq_l;
			I.parse_ori(r);
			if (r.ori.chr > 24) r.ori_unmapped = true;                    // RR:413
			r.skip = !r.ori_unmapped && r.ori.align_score == (uint32_t)(r.read_l * opt.match);   // RR:414
			r.n_draws = n_count[i];
			r.has_n = n_count[i] != 0;
			r.in_order_only = n_count[i] > MAX_VARIANT_DRAWS;
		}
	});
	for (const auto &vs : var_src) {
		ReadState &base = rs[vs.first];
		base.var_base = (int64_t)vs.second;
		const size_t nv = (size_t)1 << (2 * base.n_draws);
		for (size_t c = 0; c < nv; ++c) {
			ReadState &v = rs[vs.second + c];
			v.rec = base.rec; v.read_l = base.read_l; v.skip = base.skip;
			v.var_of = (int64_t)vs.first; v.var_code = (uint32_t)c;
		}
	}
	lap("comment parse");
	Impl::CensusScratch census_main;
	auto prepare_read = [&](ReadState &r) {                                   // encode + pack + STR census
		I.encode(r);
		const size_t words = (size_t)(r.read_l >> 5) + 2;
		for (int s = 0; s < 2; ++s) { r.bits[s].assign(words, 0); Impl::pack64(r.bin[s], r.bits[s].data(), 0, words); }
		I.str_census(r, r.bits[0].data(), census_main);
	};
	auto register_jobs = [&](ReadState &r, SeedBatch &sb) {
		for (int s = 0; s < 2; ++s) {
			SeedJob j;
			j.bits_off = (uint32_t)sb.bits.size(); j.read_len = (uint32_t)r.read_l; j.is_str = r.is_str ? 1 : 0; j.list_off = 0;
			sb.bits.append(r.bits[s].data(), r.bits[s].data() + r.bits[s].size());
			if (r.is_str) { j.list_off = (uint32_t)sb.seed_list.size(); sb.seed_list.insert(sb.seed_list.end(), r.seed_list[s].begin(), r.seed_list[s].end()); }
			r.job[s] = (int)sb.jobs.size();
			sb.jobs.push_back(j);
		}
	};
	auto merge_read = [&](ReadState &r, SeedBatch &b) {                      // stage C, part 1: no random numbers
		r.needs_rand = false;
		for (int s = 0; s < 2; ++s) {
			const int j = r.job[s];
			I.merge_mems(b.mems.data() + b.mem_off[j], b.mem_off[j + 1] - b.mem_off[j], r.vu[s]);
			for (const VertexU &u : r.vu[s]) if (u.pos_n > POS_N_MAX) r.needs_rand = true;
		}
	};
	// `i` = index of the (real) read: reads 2p draw from the first handler's random_r stream, reads 2p+1 from the second's
	auto chain_read = [&](ReadState &r, size_t i, std::vector<Edge> &edges) {   // stage C, part 2: expand + chain
		for (int s = 0; s < 2; ++s) {
			Graph &g = r.g[s];
			g.v.clear(); g.is_str = r.is_str;
			I.expand(r.vu[s], g.v, rand_r_[i & 1]);
			I.chain(g, edges);
		}
	};
	// a real read with N is not encoded here (its bases depend on the draws); the N-free mate of such a read is
	for (size_t i = 0; i < n_all; ++i) { ReadState &r = rs[i]; r.batched = !(r.skip || r.has_n || r.read_l < LEN_KMER); }
	HostSlot *hslot = acquire_host();
	struct HostRelease { AlnPipeline &P; HostSlot *h; ~HostRelease() { P.release_host(h); } } host_release{*this, hslot};
	SeedBatch &sb = hslot->seeds;
	sb.clear();
	{
		std::vector<uint32_t> word_off(n_all + 1, 0), job_of(n_all + 1, 0);
		for (size_t i = 0; i < n_all; ++i) {                                 // layout of the packed-read pool (two strands per read)
			ReadState &r = rs[i];
			word_off[i + 1] = word_off[i] + (r.batched ? 2u * (uint32_t)((r.read_l >> 5) + 2) : 0u);
			job_of[i + 1] = job_of[i] + (r.batched ? 2u : 0u);
		}
		sb.bits.resize(word_off[n_all]);
		sb.jobs.resize(job_of[n_all]);
		lap("layout");
		par_all([&](size_t b, size_t e, int) {
			Impl::CensusScratch census;
			for (size_t i = b; i < e; ++i) {
				ReadState &r = rs[i];
				if (!r.batched) continue;
				I.encode(r);
				const uint32_t words = (uint32_t)((r.read_l >> 5) + 2);
				for (int s = 0; s < 2; ++s) {
					const uint32_t off = word_off[i] + (uint32_t)s * words;
					Impl::pack64(r.bin[s], sb.bits.data(), off, words);
					SeedJob &j = sb.jobs[job_of[i] + s];
					j.bits_off = off; j.read_len = (uint32_t)r.read_l; j.is_str = 0; j.list_off = 0;
					r.job[s] = (int)(job_of[i] + s);
				}
				I.str_census(r, sb.bits.data() + word_off[i], census);
			}
		});
		lap("encode + pack + STR census");
		for (size_t i = 0; i < n_all; ++i) {                                 // STR reads are rare: their seed lists are appended in order
			ReadState &r = rs[i];
			if (!r.batched || !r.is_str) continue;
			for (int s = 0; s < 2; ++s) {
				SeedJob &j = sb.jobs[r.job[s]];
				j.is_str = 1; j.list_off = (uint32_t)sb.seed_list.size();
				sb.seed_list.insert(sb.seed_list.end(), r.seed_list[s].begin(), r.seed_list[s].end());
			}
		}
	}
	add_time(0, now() - t0); t0 = now();
	// ---- stage B
	{
		std::lock_guard<std::mutex> dev(dev_m_);
		if (!sb.jobs.empty() && !seed_service_run(seeds_, sb, err)) return false;
	}
	{ std::lock_guard<std::mutex> lk(stats_m_); stats.mems += sb.mems.size(); stats.dev.add(sb.dev); sb.dev = DevCounters(); }
	add_time(1, now() - t0); t0 = now();
	// ---- stage C: reads whose expansion draws from the per-handler random_r stream go in input order, the rest in parallel
	par_all([&](size_t b, size_t e, int) {
		std::vector<Edge> edges;
		for (size_t i = b; i < e; ++i) {
			ReadState &r = rs[i];
			if (!r.batched) continue;
			merge_read(r, sb);
			if (!r.needs_rand) chain_read(r, i, edges);
		}
	});
	std::vector<Edge> edges_main;
	Impl::PlanScratch plan_main;
	// Reads whose seeds include a unipath with more than 500 positions sample them with their handler's random_r stream
	// (expand_seed, IDX:219-258): those streams are consumed in input order, one stream per mate.  So these reads are chained
	// here one after the other; a read with 'N' runs every variant from the same stream position (the variants must leave the
	// stream at one position, otherwise the read waits for the replay); and once a read of a handler has to wait for the
	// replay -- where it draws in its turn -- every later read of that handler that needs the stream waits too.
	{
		bool any = false;
		for (size_t i = 0; i < n_all && !any; ++i) any = rs[i].batched && rs[i].needs_rand;
		for (size_t i = 0; i < n_reads && !any; ++i) any = rs[i].in_order_only;
		if (any) {
			wait_turn();                                                   // after every earlier block's replay
			bool tainted[2] = {false, false};
			for (size_t i = 0; i < n_reads; ++i) {
				ReadState &r = rs[i];
				const int h = (int)(i & 1);
				if (r.skip || r.read_l < LEN_KMER) continue;
				if (r.in_order_only) { tainted[h] = true; continue; }         // (too many N: what it will draw is not known yet)
				if (!r.has_n) {
					if (!r.needs_rand) continue;
					if (tainted[h]) { r.in_order_only = true; r.batched = false; continue; }
					chain_read(r, i, edges_main);
					continue;
				}
				if (r.var_base < 0) continue;
				const size_t nv = (size_t)1 << (2 * r.n_draws);
				bool needs = false;
				for (size_t c = 0; c < nv; ++c) needs |= rs[(size_t)r.var_base + c].needs_rand;
				if (!needs) continue;
				bool same = !tainted[h];
				const GlibcRandom start = rand_r_[h];
				GlibcRandom after = start;
				for (size_t c = 0; c < nv && same; ++c) {
					ReadState &v = rs[(size_t)r.var_base + c];
					rand_r_[h] = start;
					chain_read(v, i, edges_main);                            // (a variant that does not need the stream leaves it where it was)
					if (c == 0) after = rand_r_[h]; else same = rand_r_[h] == after;
				}
				if (same) rand_r_[h] = after;
				else {                                                       // the variants disagree (or the handler already waits)
					rand_r_[h] = start;
					r.in_order_only = true; tainted[h] = true;
					for (size_t c = 0; c < nv; ++c) rs[(size_t)r.var_base + c].batched = false;
				}
			}
		}
	}
	add_time(2, now() - t0); t0 = now();
	// ---- stage D: per-thread task lists, concatenated afterwards
	KswBatchBuf &tasks = hslot->ksw;
	{
		std::vector<KswTaskList> part((size_t)T + 1);                      // one list per chunk of reads, the last one for the variants
		std::vector<size_t> lo((size_t)T, 0), hi((size_t)T, 0);
		par_reads([&](size_t b, size_t e, int t) {
			lo[t] = b; hi[t] = e;
			Impl::PlanScratch scratch;
			for (size_t i = b; i < e; ++i) if (rs[i].batched) I.plan_read(rs[i], part[t], scratch);
		});
		for (size_t i = n_reads; i < n_all; ++i)
			if (rs[i].batched && !rs[rs[i].var_of].in_order_only) I.plan_read(rs[i], part[T], plan_main);
		// joined on the helper threads: every list is copied to its place and the task ids of its reads are rebased
		std::vector<size_t> kb((size_t)T + 2, 0), qb((size_t)T + 2, 0), tb((size_t)T + 2, 0);
		for (int t = 0; t <= T; ++t) { kb[t + 1] = kb[t] + part[t].qlen.size(); qb[t + 1] = qb[t] + part[t].q.size(); tb[t + 1] = tb[t] + part[t].t.size(); }
		const size_t nk = kb[T + 1];
		tasks.q.resize(qb[T + 1]); tasks.t.resize(tb[T + 1]);
		tasks.qoff.resize(nk); tasks.toff.resize(nk); tasks.qlen.resize(nk); tasks.tlen.resize(nk);
		parallel((size_t)T + 1, [&](size_t b, size_t e, int) {
			for (size_t t = b; t < e; ++t) {
				const KswTaskList &p = part[t];
				if (!p.q.empty()) memcpy(tasks.q.data() + qb[t], p.q.data(), p.q.size());
				if (!p.t.empty()) memcpy(tasks.t.data() + tb[t], p.t.data(), p.t.size());
				for (size_t k = 0; k < p.qlen.size(); ++k) {
					tasks.qoff[kb[t] + k] = p.qoff[k] + (int64_t)qb[t]; tasks.toff[kb[t] + k] = p.toff[k] + (int64_t)tb[t];
					tasks.qlen[kb[t] + k] = p.qlen[k]; tasks.tlen[kb[t] + k] = p.tlen[k];
				}
				const int base = (int)kb[t];
				if (!base) continue;
				const size_t ib = t < (size_t)T ? lo[t] : n_reads, ie = t < (size_t)T ? hi[t] : n_all;
				for (size_t i = ib; i < ie; ++i)
					for (auto &kv : rs[i].node_aln) for (Piece &pc : kv.second.pieces) if (pc.kind == 1) pc.task += base;
			}
		}, 2);
	}
	add_time(3, now() - t0); t0 = now();
	// ---- stage E
	const int8_t m = (int8_t)opt.match, x = (int8_t)-opt.mismatch;
	int8_t mat[25];
	for (int a = 0, k = 0; a < 5; ++a) for (int b = 0; b < 5; ++b, ++k) mat[k] = (a == 4 || b == 4) ? 0 : (a == b ? m : x);   // ksw_gen_mat_D, RR:829-844
	pansvr_ksw_params_t kp;
	kp.m = 5; kp.mat = mat; kp.gapo = (int8_t)opt.gap_open; kp.gape = (int8_t)opt.gap_ex; kp.gapo2 = (int8_t)opt.gap_open2; kp.gape2 = (int8_t)opt.gap_ex2;
	kp.w = 200; kp.zdrop = (uint16_t)opt.zdrop; kp.end_bonus = -1; kp.flag = 0;      // copy_option, RR:817-827
	// one ksw batch over a task list; the CIGAR rows are `cap` words, a list whose longest CIGAR does not fit is run again
	auto ksw_batch = [&](size_t n, const uint8_t *q, size_t q_bytes, const int64_t *qoff, const int32_t *qlen, const uint8_t *t, size_t t_bytes,
	                     const int64_t *toff, const int32_t *tlen, int &cap, const std::function<void(int32_t*&, uint32_t*&)> &out_buffers) -> bool {
		if (n == 0) return true;
		for (;;) {
			int32_t *res; uint32_t *cig;
			out_buffers(res, cig);
			std::unique_lock<std::mutex> dev(dev_m_);
			const int rc = pansvr_ksw_extd2_batch((pansvr_ksw_ctx*)ksw_, (int64_t)n, q, (int64_t)q_bytes, qoff, qlen, t, (int64_t)t_bytes, toff, tlen, &kp, res, cig, cap);
			if (rc != 0) { err = std::string("ksw batch: ") + pansvr_last_error(); return false; }
			pansvr_ksw_stats_t ks;
			if (pansvr_ksw_last_stats((pansvr_ksw_ctx*)ksw_, &ks) == 0) {
				std::lock_guard<std::mutex> lk(stats_m_);
				stats.dev.launches += ks.kernel_launches; stats.dev.h2d_bytes += ks.h2d_bytes; stats.dev.d2h_bytes += ks.d2h_bytes;
				stats.dev.ksw_kernel_ms += ks.kernel_ms;
			}
			dev.unlock();
			std::atomic<int> need(0);
			std::atomic<uint64_t> cells(0);
			parallel(n, [&](size_t b, size_t e, int) {
				int nd = 0; uint64_t c = 0;
				for (size_t i = b; i < e; ++i) {
					if (res[i * PANSVR_RES_WORDS + PANSVR_RES_STATUS] & 1) nd = std::max(nd, res[i * PANSVR_RES_WORDS + PANSVR_RES_N_CIGAR]);
					c += (uint64_t)pansvr_ksw_band_cells(qlen[i], tlen[i], kp.w);
				}
				cells += c;
				int cur = need.load();
				while (nd > cur && !need.compare_exchange_weak(cur, nd)) {}
			});
			if (!need.load()) { std::lock_guard<std::mutex> lk(stats_m_); stats.ksw_tasks += n; stats.ksw_cells += cells.load(); return true; }
			cap = need.load() + 8;
		}
	};
	auto run_ksw = [&](KswTaskList &tl) -> bool {
		const size_t n = tl.qlen.size();
		return ksw_batch(n, tl.q.data(), tl.q.size(), tl.qoff.data(), tl.qlen.data(), tl.t.data(), tl.t.size(), tl.toff.data(), tl.tlen.data(), tl.cap,
		                 [&](int32_t *&res, uint32_t *&cig) { tl.res.resize(n * PANSVR_RES_WORDS); tl.cig.resize(n * (size_t)tl.cap); res = tl.res.data(); cig = tl.cig.data(); });
	};
	{
		const size_t n = tasks.qlen.size();
		tasks.cap = 16;
		if (!ksw_batch(n, tasks.q.data(), tasks.q.size(), tasks.qoff.data(), tasks.qlen.data(), tasks.t.data(), tasks.t.size(), tasks.toff.data(), tasks.tlen.data(),
		               tasks.cap, [&](int32_t *&res, uint32_t *&cig) { tasks.res.resize(n * PANSVR_RES_WORDS); tasks.cig.resize(n * (size_t)tasks.cap); res = tasks.res.data(); cig = tasks.cig.data(); }))
			return false;
	}
	const KswView tasks_view{tasks.res.data(), tasks.cig.data(), tasks.cap};
	add_time(4, now() - t0); t0 = now();

	// ---- stage F: chain selection, result sort and pairing.  Only exact ties consume rand() (RR:247, RRH:553), so every
	// pair is first finished on a worker thread against a probe; pairs that asked for a random number, and the deferred
	// pairs, are then replayed in input order against the real stream -- the stream sees exactly the reference's calls.
	std::vector<Impl::PE> pes(n_pairs);
	// what the in-order pass has to do for a pair:
	//   0  nothing: no random number is drawn anywhere
	//   1  advance the stream by the pair's draws: the candidate lists have ties, but every outcome of them gives the same lists
	//      (explore_read) and the pairing has no ties
	//   2  the same, then redraw the pairing ties from the recorded events; the winner is applied on the helper threads afterwards
	//   4 + bits  finish read 0 (bit 0) / read 1 (bit 1) against the real stream (their outcomes differ, or are too many to
	//      enumerate), then pair up
	//   16 a pair with 'N' (see below)
	std::vector<uint8_t> redo(n_pairs, 0), draws0(n_pairs, 0), draws1(n_pairs, 0), ev_cnt(n_pairs, 0), ev_chunk(n_pairs, 0);
	std::vector<uint32_t> ev_off(n_pairs, 0);
	std::vector<int8_t> win_i(n_pairs, -1), win_j(n_pairs, -1);
	std::vector<std::vector<PairEvent>> events((size_t)std::max(1, opt.threads));
	parallel(n_pairs, [&](size_t pb, size_t pe_, int t) {
		RandTap probe;
		std::vector<uint32_t> sig0, sig;
		std::vector<PairEvent> ev, &evs = events[(size_t)t];
		for (size_t pi = pb; pi < pe_; ++pi) {
			ReadState *se = &rs[2 * pi];
			if (se[0].has_n || se[1].has_n || se[0].in_order_only || se[1].in_order_only) { redo[pi] = 16; continue; }
			const int c0 = I.explore_read(se[0], tasks_view, probe, sig0, sig);
			const int c1 = I.explore_read(se[1], tasks_view, probe, sig0, sig);
			if (c0 < 0 || c1 < 0 || c0 > 250 || c1 > 250) {
				redo[pi] = (uint8_t)(4 | (c0 < 0 || c0 > 250 ? 1 : 0) | (c1 < 0 || c1 > 250 ? 2 : 0));
				draws0[pi] = (uint8_t)(c0 < 0 || c0 > 250 ? 0 : c0); draws1[pi] = (uint8_t)(c1 < 0 || c1 > 250 ? 0 : c1);
				continue;
			}
			draws0[pi] = (uint8_t)c0; draws1[pi] = (uint8_t)c1;
			probe.restart(0);
			ev.clear();
			I.pair_up(se, pes[pi], probe, &ev);
			if (probe.calls == 0) {                                          // the pairing is decided
				redo[pi] = c0 + c1 ? 1 : 0;
				if (pes[pi].gain) I.set_primary(se, pes[pi]);
			} else if (ev.size() <= 255) {
				redo[pi] = 2;
				ev_chunk[pi] = (uint8_t)t; ev_off[pi] = (uint32_t)evs.size(); ev_cnt[pi] = (uint8_t)ev.size();
				evs.insert(evs.end(), ev.begin(), ev.end());
			} else redo[pi] = 4;                                             // (cannot happen with <= 13 x 13 candidates; stay safe)
		}
	});
	const double t_probe = now() - t0;
	size_t n_redo = 0, n_full = 0, n_in_order = 0, n_deferred = 0, n_mems_late = 0;
	RandTap real; real.real = &rand_;
	// the replay walks cold per-read data on one thread: pull the state of the pairs a few steps ahead into the cache
	std::vector<uint32_t> redo_list;
	for (size_t pi = 0; pi < n_pairs; ++pi) if (redo[pi]) redo_list.push_back((uint32_t)pi);
	wait_turn();                                                          // ---- in input order from here: the rand() stream
	auto prefetch_state = [&](size_t k) { if (k < redo_list.size()) { const char *p = (const char*)&rs[2 * (size_t)redo_list[k]]; for (size_t o = 0; o < 2 * sizeof(ReadState); o += 64) __builtin_prefetch(p + o); } };
	auto prefetch_arrays = [&](size_t k) {
		if (k >= redo_list.size()) return;
		for (int m = 0; m < 2; ++m) {
			const ReadState &r = rs[2 * (size_t)redo_list[k] + m];
			for (int s2 = 0; s2 < 2; ++s2) {
				const char *p = (const char*)r.g[s2].path.data(); const size_t n = r.g[s2].path.size() * sizeof(PathNode);
				for (size_t o = 0; o < n && o < 512; o += 64) __builtin_prefetch(p + o);
				__builtin_prefetch(r.g[s2].v.data());
			}
			__builtin_prefetch(r.node_aln.data());
			__builtin_prefetch(r.result.data());
		}
	};
	for (size_t ri = 0; ri <= redo_list.size(); ++ri) {
		if (H) H->fast_until(ri < redo_list.size() ? (uint64_t)H->global_pair[redo_list[ri]] : ~(uint64_t)0);   // the device path's pairs before this one
		if (ri == redo_list.size()) break;
		const size_t pi = redo_list[ri];
		if (redo[pi] >= 4) { prefetch_state(ri + 8); prefetch_arrays(ri + 3); ++n_full; }
		++n_redo;
		ReadState *se = &rs[2 * pi];
		if (redo[pi] == 16) {                                             // deferred pair: its rand() / random_r draws happen now
			++n_deferred;
			KswTaskList local;
			bool own[2] = {false, false};
			for (int k = 0; k < 2; ++k) {
				ReadState &r = se[k];
				if (r.skip || r.read_l < LEN_KMER || (!r.has_n && !r.in_order_only)) { /* nothing drawn: skipped, too short, or prepared in the batch */ }
				else if (r.has_n && r.var_base >= 0 && !r.in_order_only) {     // draw the substitutions, adopt the variant prepared for them
					uint32_t code = 0;
					for (int j = 0; j < r.n_draws; ++j) code |= (uint32_t)(rand_.next() % 4) << (2 * j);
					ReadState &v = rs[(size_t)r.var_base + code];
					for (int s = 0; s < 2; ++s) { r.bin[s].swap(v.bin[s]); r.vu[s].swap(v.vu[s]); std::swap(r.g[s], v.g[s]); }
					r.is_str = v.is_str; r.node_aln.swap(v.node_aln);
				} else {
					++n_in_order;
					own[k] = true;
					SeedBatch &one = seed_small_;
					one.clear();
					prepare_read(r);
					register_jobs(r, one);
					{
						std::lock_guard<std::mutex> dev(dev_m_);
						if (!seed_service_run(seeds_, one, err)) return false;
					}
					n_mems_late += one.mems.size();
					{ std::lock_guard<std::mutex> lk(stats_m_); stats.dev.add(one.dev); one.dev = DevCounters(); }
					merge_read(r, one);
					chain_read(r, 2 * pi + k, edges_main);
					I.plan_read(r, local, plan_main);
					if (!run_ksw(local)) return false;
				}
				I.finish_read(r, own[k] ? local.view() : tasks_view, real);
			}
		} else if (redo[pi] & 4) {
			if (redo[pi] & 1) I.finish_read(se[0], tasks_view, real); else for (int k = 0; k < draws0[pi]; ++k) rand_.next();
			if (redo[pi] & 2) I.finish_read(se[1], tasks_view, real); else for (int k = 0; k < draws1[pi]; ++k) rand_.next();
		} else {                                                          // 1 or 2: the candidate lists stand whatever is drawn
			for (int k = 0, n = draws0[pi] + draws1[pi]; k < n; ++k) rand_.next();
			if (redo[pi] == 2) {                                           // store_pair's tie rule over the recorded events (RRH:553)
				const PairEvent *e = events[ev_chunk[pi]].data() + ev_off[pi];
				int max_same = 1, wi = -1, wj = -1;
				for (int k = 0; k < ev_cnt[pi]; ++k) {
					if (!e[k].tie) { max_same = 1; wi = e[k].i; wj = e[k].j; }
					else { ++max_same; if (rand_.next() % max_same == 0) { wi = e[k].i; wj = e[k].j; } }
				}
				win_i[pi] = (int8_t)wi; win_j[pi] = (int8_t)wj;
			}
			continue;
		}
		I.pair_up(se, pes[pi], real);
		if (pes[pi].gain) I.set_primary(se, pes[pi]);
	}
	turn.pass();                                                          // the next block may replay now
	{ std::lock_guard<std::mutex> lk(stats_m_); stats.t_in_order += now() - t_turn; stats.in_order_pairs += n_redo; }
	trace_host(seq, "io_begin", t_turn);
	trace_host(seq, "io_end");
	parallel(n_pairs, [&](size_t pb, size_t pe_, int) {                    // winners of the redrawn pairings
		for (size_t pi = pb; pi < pe_; ++pi) {
			if (redo[pi] != 2) continue;
			I.apply_pairing(&rs[2 * pi], pes[pi], win_i[pi], win_j[pi]);
			if (pes[pi].gain) I.set_primary(&rs[2 * pi], pes[pi]);
		}
	});
	{ std::lock_guard<std::mutex> lk(stats_m_); stats.reads += 2 * n_pairs; stats.deferred_pairs += n_deferred; stats.mems += n_mems_late; }
	if (getenv("PANSVR_TIMING"))
		fprintf(stderr, "[timing] finish: probe %.3f s, in-order pass over %zu/%zu pairs (%zu of them re-finished) %.3f s (%zu variant states for reads with N, %zu reads prepared in order)\n",
		        t_probe, n_redo, n_pairs, n_full, now() - t0 - t_probe, n_var, n_in_order);
	const double t_text = now();
	// ---- SAM text of every pair (no random numbers involved any more: parallel)
	auto text_pair = [&](size_t pi, std::string &sam, std::string &ori) {
		ReadState *se = &rs[2 * pi];
		const Impl::PE &pe = pes[pi];
		if (pe.gain)
			for (int k = 0; k < 2; ++k) I.output_bam(se[k], sam, k == 0, pe.cur_isize);
		I.output_ori_pair(se, pe, sam, ori, min_filter_score_);
	};
	if (H) { if (!H->emit(text_pair, err)) return false; }
	else parallel(n_pairs, [&](size_t pb, size_t pe_, int t) {            // chunk t writes its pairs, in order, into buffer t
		// the string headers of neighbouring chunks share cache lines and every append updates the length: work on locals
		std::string sam, ori;
		sam.swap(out.sam[(size_t)t]); ori.swap(out.ori[(size_t)t]);
		sam.reserve((pe_ - pb) * 2 * (2 * (size_t)opt.read_len + 400));
		for (size_t pi = pb; pi < pe_; ++pi) text_pair(pi, sam, ori);
		sam.swap(out.sam[(size_t)t]); ori.swap(out.ori[(size_t)t]);
	});
	add_time(5, now() - t0);
	if (timing) fprintf(stderr, "[timing]   F/record text %.3f s\n", now() - t_text);
	if (getenv("PANSVR_TIMING")) { double a = 0; for (int i = 0; i < 6; ++i) a += stats.t_stage[i]; fprintf(stderr, "[timing] align_block body done, stages A-F %.3f s\n", a); }
	return true;
}


void trace_mark(uint64_t seq, const char *what) { trace_host(seq, what); }

// ================================================================================================ the device path
bool AlnPipeline::align_block(const FastqRec *recs, size_t n_reads_in, BlockOutput &out, std::string &err, uint64_t seq)
{
	const size_t n_pairs = n_reads_in / 2, nd = 2 * n_pairs;
	out.sam_text.clear(); out.placed = false; out.placed_bytes = 0; out.place_called = false;
	if (!stages_ || n_pairs == 0) return align_block_host(recs, n_reads_in, out, err, seq, nullptr);
	ensure_read_stats(recs[0]);
	// records are views into one buffer, in input order: the text they span goes up as it is
	const char *base = recs[0].name;
	const char *end = recs[nd - 1].qual + recs[nd - 1].qual_l;
	for (size_t t = 0; t < nd; t += 1 + (nd > 2 ? nd - 2 : 0)) {          // (first and last suffice for a parsed buffer; be safe about odd callers)
		base = std::min(base, recs[t].name); end = std::max(end, recs[t].qual + recs[t].qual_l);
	}
	return align_block_dev(base, (size_t)(end - base), recs, n_pairs, out, err, seq, nullptr);
}

// A block given as text: strict 4-line FASTQ of n_pairs interleaved pairs (the caller counted the lines).  The record table is
// made on the device; *reparse = true (and nothing done) if the text turns out not to be what it was taken for.
bool AlnPipeline::align_block_text(const char *text, size_t bytes, size_t n_pairs, BlockOutput &out, std::string &err, uint64_t seq, bool *reparse)
{
	*reparse = false;
	out.sam_text.clear(); out.placed = false; out.placed_bytes = 0; out.place_called = false;
	if (!stages_ || n_pairs == 0) { *reparse = true; return true; }
	return align_block_dev(text, bytes, nullptr, n_pairs, out, err, seq, reparse);
}

bool AlnPipeline::align_block_dev(const char *base, size_t text_bytes, const FastqRec *recs, size_t n_pairs, BlockOutput &out, std::string &err, uint64_t seq, bool *reparse)
{
	trace_host(seq, "enter");
	// whatever happens on the way (but not when the block is handed back for the host parser: it comes again under the same number)
	struct TurnOnExit { AlnPipeline &P; uint64_t seq; bool again; ~TurnOnExit() { if (!again) P.pass_turn_if_pending(seq); } } turn_on_exit{*this, seq, false};
	const size_t nd = 2 * n_pairs;
	Impl I(*this);
	const DebgaIndex &idx = idx_;
	auto add_time = [&](int stage, double dt) { std::lock_guard<std::mutex> lk(stats_m_); stats.t_stage[stage] += dt; };
	double t0 = now();
	DevBuffers *db = acquire_dev(err, seq);
	if (!db) return false;
	struct Release { AlnPipeline &P; DevBuffers *b; ~Release() { P.release_dev(b); } } release{*this, db};
	if (text_bytes >= 0xfffffff0ull) { err = "block too large for the device stages (cut it into smaller blocks)"; return false; }
	// the text goes up from page-locked memory: the caller's own buffer if it is (pansvr_host_alloc), else a staging copy of it
	const bool pinned_input = staging_is_pinned(base);
	if (!pinned_input) {
		db->text.resize(text_bytes + 1);
		const size_t piece = (size_t)4 << 20, n_piece = (text_bytes + piece - 1) / piece;
		parallel(n_piece, [&](size_t b, size_t e, int) { for (size_t k = b; k < e; ++k) memcpy(db->text.data() + k * piece, base + k * piece, std::min(piece, text_bytes - k * piece)); }, 2);
	}
	size_t words = 0, list_bytes = 0;
	if (recs) {
		db->reads.resize(nd); db->recs.resize(nd);
		for (size_t t = 0; t < nd; ++t) {
			const FastqRec &r = recs[t];
			DevRec &dr = db->recs[t];
			dr.name_off = (uint32_t)(r.name - base); dr.comment_off = (uint32_t)(r.comment - base); dr.seq_off = (uint32_t)(r.seq - base); dr.qual_off = (uint32_t)(r.qual - base);
			dr.name_l = r.name_l; dr.comment_l = r.comment_l; dr.seq_l = r.seq_l; dr.qual_l = r.qual_l;
			DevRead &d = db->reads[t];
			d.seq_off = dr.seq_off; d.len = r.seq_l; d.var_code = 0; d.bits_off = (uint32_t)words; d.list_off = (uint32_t)list_bytes;
			if (r.seq_l >= LEN_KMER) { words += 2 * (size_t)((r.seq_l >> 5) + 2); list_bytes += 2 * (size_t)(r.seq_l - LEN_KMER + 1); }
		}
		if (words >= 0xffffffffull || list_bytes >= 0xffffffffull) { err = "block too large for the device stages (cut it into smaller blocks)"; return false; }
	}
	add_time(0, now() - t0); t0 = now();
	// ---- first trip: the original alignments, stages A..F1 and the probe of stage F
	DevStageIn in;
	in.text = pinned_input ? (const uint8_t*)base : db->text.data(); in.text_bytes = text_bytes; in.reads = db->reads.data(); in.n_reads = nd; in.bits_words = words; in.list_bytes = list_bytes;
	in.scores = AlnScores{opt.match, opt.mismatch, opt.gap_open, opt.gap_ex, opt.gap_open2, opt.gap_ex2};
	in.recs = recs ? db->recs.data() : nullptr; in.parse_text = recs == nullptr;
	in.pair_opts = PairOpts{opt.isize_max, opt.isize_min, opt.read_len, min_filter_score_};
	const char *dump = getenv("PANSVR_DUMP_STAGES");
	in.want_tables = dump != nullptr;
	{
		const size_t site = db->site;
		{
			std::unique_lock<std::mutex> lk(trip1_m_);
			if (trip1_busy_.size() <= site) { trip1_busy_.resize(site + 1, 0); trip1_wait_.resize(site + 1); }
			trip1_wait_[site].insert(seq);                              // admitted oldest first
			trip1_cv_.wait(lk, [&]() { return trip1_busy_[site] < trip1_cap_ && *trip1_wait_[site].begin() == seq; });
			trip1_wait_[site].erase(seq);
			++trip1_busy_[site];
			lk.unlock();
			trip1_cv_.notify_all();
		}
		struct Leave { AlnPipeline &P; size_t site; ~Leave() { { std::lock_guard<std::mutex> lk(P.trip1_m_); --P.trip1_busy_[site]; } P.trip1_cv_.notify_all(); } } leave{*this, site};
		trace_host(seq, "trip1_begin");
		if (const char *e = getenv("PANSVR_TEST_FAIL_SEQ")) if ((uint64_t)atol(e) == seq) { err = "injected failure of this sub-block's first trip (test)"; return false; }
		if (!stage_service_run(db->svc, in, db->out, err)) return false;
		trace_host(seq, "trip1_end");
	}
	DevStageOut &o = db->out;
	if (!o.parse_ok) { if (reparse) { *reparse = true; turn_on_exit.again = true; return true; } err = "internal error: record table rejected"; return false; }
	// a record of the block as the host path sees it
	const DevRec *drecs = recs ? db->recs.data() : o.recs.data();
	auto rec_at = [&](size_t t) -> FastqRec {
		if (recs) return recs[t];
		const DevRec &d = drecs[t];
		FastqRec r;
		r.name = base + d.name_off; r.comment = base + d.comment_off; r.seq = base + d.seq_off; r.qual = base + d.qual_off;
		r.name_l = d.name_l; r.comment_l = d.comment_l; r.seq_l = d.seq_l; r.qual_l = d.qual_l;
		return r;
	};
	add_time(1, now() - t0); t0 = now();
	if (dump) {                                                           // tests: what the device stages returned, for the differential
		const std::string path = std::string(dump) + "." + std::to_string(seq);   // between the CUDA backend and the host-stepped one
		if (FILE *f = fopen(path.c_str(), "wb")) {
			const uint64_t hdr[4] = {nd, o.seed_off[2 * nd], o.cands.size(), o.mem_off[2 * nd]};
			fwrite(hdr, 8, 4, f);
			fwrite(o.flags.data(), 1, nd, f); fwrite(o.mem_off.data(), 4, 2 * nd + 1, f); fwrite(o.seed_off.data(), 4, 2 * nd + 1, f);
			fwrite(o.seeds.data(), sizeof(DevSeed), o.seeds.size(), f); fwrite(o.dist.data(), 4, o.dist.size(), f); fwrite(o.pre.data(), 4, o.pre.size(), f);
			fwrite(o.cand_off.data(), 4, nd + 1, f);
			for (size_t c = 0; c < o.cands.size(); ++c) {
				DevCand cd = o.cands[c];
				fwrite(o.cigs.data() + cd.cig_off, sizeof(DevCigar), cd.n_cig, f);
				cd.piece_off = cd.cig_off = cd.cig_cap = 0;
				fwrite(&cd, sizeof cd, 1, f);
			}
			for (size_t k = 0; k < o.pair_probe.size(); ++k) {             // (only what the probe defines: events beyond ev_cnt are not written by it)
				const DevProbe &pr = o.pair_probe[k];
				fwrite(&pr, 4, 1, f); fwrite(pr.ev_i, 1, pr.ev_cnt, f); fwrite(pr.ev_j, 1, pr.ev_cnt, f); fwrite(&pr.tie_mask, 4, 1, f);
			}
			fclose(f);
		}
	}
	// ---- the pairs the host path finishes: 'N' / 'n' in a read, a unipath that needs random_r
	// (PANSVR_TIES_ON_HOST_PATH=1, tests: also the pairs whose ties decide their outcome, as before round 2's last change)
	static const bool ties_on_host_path = getenv("PANSVR_TIES_ON_HOST_PATH") != nullptr;
	std::vector<uint32_t> host_list;
	size_t io_pairs = 0;
	for (size_t p = 0; p < n_pairs; ++p) {
		const uint8_t rd = o.redo[p];
		if (rd == PR_REDO_HOST || (ties_on_host_path && rd == PR_REDO_TIES)) host_list.push_back((uint32_t)p);
		else io_pairs += rd != 0;
	}
	std::vector<FastqRec> hrecs(2 * host_list.size());
	for (size_t s = 0; s < host_list.size(); ++s) { hrecs[2 * s] = rec_at(2 * (size_t)host_list[s]); hrecs[2 * s + 1] = rec_at(2 * (size_t)host_list[s] + 1); }
	// in-order pass of the device path's pairs: each advances the stream by a count the probe knows (its reads' draws, one per tied
	// pairing event), so the pass only takes the numbers -- draw_off[p] of them before pair p -- and the second trip redraws the
	// pairing winners from them on the device (RRH:553).  In between, in their turn: the pairs whose ties decide their outcome,
	// finished here against the stream itself from the seeds, chain tables and candidates the device computed (no stage is run
	// again for them); and the host path's pairs.
	const size_t n_drawn = n_pairs ? o.draw_off[n_pairs] : 0;
	db->drawn.resize(n_drawn + 1);
	const size_t n_ties = ties_on_host_path ? 0 : o.ties.size();
	db->tie_pair.resize(n_ties); db->tie_done.resize(n_ties);
	size_t taken = 0, tie_cur = 0;
	std::vector<uint8_t> tie_used;
	PairIndexView host_pix_;
	host_pix_.chr_search_index = idx.chr_search_index.data(); host_pix_.chr_end_n = idx.chr_end_n.data(); host_pix_.sv = (const DevSv*)host_sv_.data();
	auto finish_tie = [&](size_t k) {
		const DevTie &t = o.ties[k];
		ReadView R[2];
		tie_used.assign((size_t)(t.seed_off[4] - t.seed_off[0]) + 1, 0);
		for (int m = 0; m < 2; ++m) {
			for (int s2 = 0; s2 < 2; ++s2) {
				const uint32_t rel = t.seed_off[2 * m + s2] - t.seed_off[0], at = t.seed_at + rel;
				R[m].v[s2] = o.tie_seeds.data() + at; R[m].dist[s2] = o.tie_dist.data() + at; R[m].pre[s2] = o.tie_pre.data() + at;
				R[m].used[s2] = tie_used.data() + rel; R[m].n[s2] = t.seed_off[2 * m + s2 + 1] - t.seed_off[2 * m + s2];
			}
			R[m].cands = o.tie_cands.data();
			R[m].cand_b = t.cand_at + (t.cand_off[m] - t.cand_off[0]); R[m].cand_e = t.cand_at + (t.cand_off[m + 1] - t.cand_off[0]);
			R[m].ori = t.ori[m];
		}
		DevTap tap;
		tap.restart(0);
		tap.real = &rand_;
		DevPairState &st = db->tie_done[k];
		memset((void*)&st, 0, sizeof st);
		dev_finish_pair_in_order(host_pix_, in.pair_opts, R, st, tap);
		const int32_t delta = (int32_t)t.cand_off[0] - (int32_t)t.cand_at;     // candidate numbers back to the device's
		for (int m = 0; m < 2; ++m) for (int x = 0; x < st.n[m]; ++x) if (st.res[m][x].cand >= 0) st.res[m][x].cand += delta;
		db->tie_pair[k] = t.pair;
	};
	BlockHooks H;
	H.global_pair = host_list.data();
	H.fast_until = [&](uint64_t upto) {
		int32_t *r = db->drawn.data();
		for (;;) {
			const uint64_t next_tie = tie_cur < n_ties ? (uint64_t)o.ties[tie_cur].pair : ~(uint64_t)0;
			const size_t to = o.draw_off[(size_t)std::min<uint64_t>(std::min(upto, next_tie), n_pairs)];
			for (; taken < to; ++taken) r[taken] = rand_.next();
			if (next_tie >= upto) break;
			finish_tie(tie_cur++);
		}
	};
	H.emit = [&](const std::function<void(size_t, std::string&, std::string&)> &text_host, std::string &e2) -> bool {
		double t1 = now();
		// the host path's pairs first: the device leaves room for their records
		const size_t nh = host_list.size();
		std::vector<std::string> h_sam(nh), h_ori(nh);
		parallel(nh, [&](size_t b, size_t e, int) { for (size_t s = b; s < e; ++s) text_host(s, h_sam[s], h_ori[s]); }, 64);
		db->host_len.resize(n_pairs + 1);
		for (size_t p = 0; p < n_pairs; ++p) db->host_len[p] = 0;
		for (size_t s = 0; s < nh; ++s) db->host_len[host_list[s]] = (uint32_t)h_sam[s].size();
		// ---- second trip: the winners go up; primary / secondary / mate of every read and the block's SAM text come back
		bool got = false;
		o.text_dest = [&](size_t total) -> char* { char *p = out.place ? out.place(total) : nullptr; out.place_called = true; got = p != nullptr; return p; };
		trace_host(seq, "trip2_begin");
		const bool fin_ok = stage_service_finalize(db->svc, in.pair_opts, opt.not_ori ? 1 : 0, n_pairs, db->drawn.data(), n_drawn, db->host_len.data(), db->tie_pair.data(), db->tie_done.data(), n_ties, o, out.sam_text, e2);
		o.text_dest = nullptr;
		trace_host(seq, "trip2_end");
		if (!fin_ok) return false;
		out.placed = got;
		if (out.placed) { out.placed_bytes = o.text_total; out.placed_ptr = o.text_ptr; }
		for (size_t s = 0; s < nh; ++s) if (!h_sam[s].empty()) memcpy(o.text_ptr + o.txt_off[2 * (size_t)host_list[s]], h_sam[s].data(), h_sam[s].size());
		bad_cigar_records_ += o.bad_records;
		add_time(2, now() - t1); t1 = now();
		// ---- the `-p` records (output_ori_bam, RR:656-717, 776-797) are rare: written here from what came back
		std::vector<int32_t> slot_of(n_pairs, -1);
		for (size_t s = 0; s < nh; ++s) slot_of[host_list[s]] = (int32_t)s;
		Impl Iq(*this);
		Iq.count_bad = false;
		// the device selected the pairs that may qualify (score at most min_filter_score, both originals on a chromosome)
		std::vector<uint32_t> sel_order(o.sel_pair.size());
		for (size_t k = 0; k < sel_order.size(); ++k) sel_order[k] = (uint32_t)k;
		std::sort(sel_order.begin(), sel_order.end(), [&](uint32_t x, uint32_t y) { return o.sel_pair[x] < o.sel_pair[y]; });
		parallel(n_pairs, [&](size_t pb, size_t pe_, int t) {             // chunk t writes its pairs, in order, into buffer t
			std::string ori, scratch;
			ori.swap(out.ori[(size_t)t]);
			// what this chunk has to look at, in pair order: the host path's pairs and the selected ones
			size_t hi = std::lower_bound(host_list.begin(), host_list.end(), (uint32_t)pb) - host_list.begin();
			size_t si = std::lower_bound(sel_order.begin(), sel_order.end(), (uint32_t)pb, [&](uint32_t x, uint32_t v) { return o.sel_pair[x] < v; }) - sel_order.begin();
			size_t pi_all = pb;                                               // (sel_all: every pair is looked at)
			for (;;) {
				size_t pi; const DevFinal *f2 = nullptr; const DevPairFinal *pfp = nullptr; bool host = false;
				if (o.sel_all) {
					if (pi_all >= pe_) break;
					pi = pi_all++;
					if (slot_of[pi] >= 0) host = true; else { f2 = &o.fin[2 * pi]; pfp = &o.pfin[pi]; }
				} else {
					const size_t hp = hi < host_list.size() ? host_list[hi] : (size_t)-1, sp_ = si < sel_order.size() ? o.sel_pair[sel_order[si]] : (size_t)-1;
					pi = std::min(hp, sp_);
					if (pi >= pe_) break;
					if (hp == pi) { host = true; ++hi; if (sp_ == pi) ++si; }
					else { f2 = &o.sel_fin[2 * (size_t)sel_order[si]]; pfp = &o.sel_pfin[sel_order[si]]; ++si; }
				}
				if (host) { ori += h_ori[(size_t)slot_of[pi]]; continue; }
				const DevPairFinal &pf = *pfp;
				if (!pf.valid || !(pf.max_score <= min_filter_score_)) continue;
				// rebuild what output_ori_bam reads of the two handlers
				ReadState se[2];
				Result prim[2], sec[2];
				Impl::PE pe;
				const FastqRec two[2] = {rec_at(2 * pi), rec_at(2 * pi + 1)};
				pe.max_score = pf.max_score; pe.cur_isize = pf.cur_isize; pe.gain = pf.gain != 0; pe.proper = pf.proper != 0;
				for (int m = 0; m < 2; ++m) {
					ReadState &r = se[m];
					r.rec = &two[m];
					r.comment.assign(r.rec->comment, r.rec->comment_l);
					r.read_l = (int)r.rec->seq_l;
					Iq.parse_ori(r);
					if (r.ori.chr > 24) r.ori_unmapped = true;
					const DevFinal &f = f2[m];
					if (!(f.flags & FIN_PRIMARY)) continue;
					Result *c;
					if (f.flags & FIN_P_ORI) c = &r.ori;
					else {
						c = &prim[m];
						c->is_ori = false; c->chr = f.p_chr; c->ref_bg = f.p_ref_bg; c->align_score = f.p_align; c->chain_score = f.p_chain;
						c->mapq = (uint8_t)f.p_mapq; c->direction = (f.flags & FIN_P_FWD) ? FORWARD : REVERSE; c->cigar_ok = (f.flags & FIN_P_CIGAR_OK) != 0;
						c->cigar.clear();                                       // (stand-in with the two properties the `-p` decision reads: empty or not, inserted bases)
						if (f.p_ncig) { c->cigar.push_back(CigarPath{0, 1}); if (f.p_ins) c->cigar.push_back(CigarPath{1, (int16_t)std::min<uint32_t>(f.p_ins, 32767)}); }
					}
					c->sv = f.p_sv >= 0 ? &idx.sv_info[(size_t)f.p_sv] : nullptr;
					c->has_mate = (f.flags & FIN_HAS_MATE) != 0; c->mate_chr = f.mate_chr; c->mate_ref_bg = f.mate_ref_bg;
					c->mate_sv = f.p_mate_sv >= 0 ? &idx.sv_info[(size_t)f.p_mate_sv] : nullptr;
					(m == 0 ? pe.m1 : pe.m2) = c;
					// the main record (written by the device) came first: a reverse-strand record leaves the read reverse-complemented
					// in place and back (RR:502-510), after which anything but ACGT reads 'N'
					if (pe.gain && c->chr != U32MAX && !(opt.not_ori && c->is_ori) && c->direction == REVERSE) r.seq_twice_reversed = true;
				}
				Iq.output_ori_pair(se, pe, scratch, ori, min_filter_score_);
			}
			ori.swap(out.ori[(size_t)t]);
		});
		add_time(5, now() - t1);
		return true;
	};
	const bool ok = align_block_host(hrecs.data(), hrecs.size(), out, err, seq, &H);
	trace_host(seq, "done");
	if (ok) {
		std::lock_guard<std::mutex> lk(stats_m_);
		stats.reads += 2 * (n_pairs - host_list.size());
		stats.in_order_pairs += io_pairs; stats.in_order_draws += n_drawn; stats.host_pairs += host_list.size(); stats.tie_pairs += n_ties;
		stats.mems += o.mem_off[2 * nd]; stats.ksw_tasks += o.n_tasks; stats.ksw_cells += o.n_cells;
		stats.dev.add(o.dev); stats.dev.seed_probes += (int64_t)o.probes;
		o.dev = DevCounters();
	}
	return ok;
}

} // namespace pansvr
