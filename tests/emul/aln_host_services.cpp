// aln_host_services.cpp -- TEST INFRASTRUCTURE ONLY.
// Host stand-ins for the two device services of the aln pipeline, so that the host-side logic of
// pansvr_b200/csrc/aln/pipeline.cpp (stages A, C, D, F) can be checked against the reference's SAM on a machine
// without a GPU: seeding steps seed_core.cuh on the host, ksw calls the oracle (oracle/libksw_oracle.so).
// The product library links seed_gpu.cu and ksw_batch.cu instead; nothing here is ever part of it.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>

#include "../../include/pansvr_b200.h"
#include "../../pansvr_b200/csrc/aln/pipeline.hpp"
#include "../../pansvr_b200/csrc/aln/stages_run.hpp"

namespace pansvr {

void *staging_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void staging_free(void *p) { free(p); }
bool staging_is_pinned(const void *) { return false; }

struct SeedService { IndexView view; };

SeedService *seed_service_create(const DebgaIndex &idx, int, std::string &)
{
	SeedService *s = new SeedService();
	s->view.seqb = idx.seqb.data(); s->view.seqf = idx.seqf.data(); s->view.posp = idx.posp.data();
	s->view.bkt_dir = idx.bkt_dir.data(); s->view.bkt_key = idx.bkt_key.data(); s->view.bkt_start = idx.bkt_start.data();
	s->view.off_g = idx.off_g.data(); s->view.kmer_g = idx.kmer_g.data(); s->view.n_seqf = idx.seqf.size();
	return s;
}
void seed_service_destroy(SeedService *s) { delete s; }
bool seed_service_run(SeedService *s, SeedBatch &b, std::string &)
{
	const size_t n = b.jobs.size();
	b.mem_off.assign(n + 1, 0);
	b.mems.clear();
	std::vector<Mem> tmp(1024);
	for (size_t i = 0; i < n; ++i) {
		const SeedJob &j = b.jobs[i];
		int c = seed_read_strand(s->view, b.bits.data() + j.bits_off, j.read_len, j.is_str != 0, b.seed_list.data() + j.list_off, tmp.data(), (int)tmp.size());
		if (c > (int)tmp.size()) { tmp.resize(c); c = seed_read_strand(s->view, b.bits.data() + j.bits_off, j.read_len, j.is_str != 0, b.seed_list.data() + j.list_off, tmp.data(), c); }
		b.mems.append(tmp.data(), tmp.data() + c);
		b.mem_off[i + 1] = (uint32_t)b.mems.size();
	}
	return true;
}

// ---- host backend of the device stages (stages_run.hpp): the same functors, stepped by a loop over plain memory
struct HostBackend {
	void *p[SL_COUNT]; size_t cap[SL_COUNT];
	pansvr_ksw_params_t kp; int8_t mat[25];
	HostBackend() { for (int i = 0; i < SL_COUNT; ++i) { p[i] = nullptr; cap[i] = 0; } }
	~HostBackend() { for (int i = 0; i < SL_COUNT; ++i) free(p[i]); }
	template <class T> T *buf(int slot, size_t n)
	{
		const size_t bytes = n * sizeof(T) + 64;
		if (bytes > cap[slot]) { free(p[slot]); p[slot] = malloc(bytes); cap[slot] = bytes; }
		return (T*)p[slot];
	}
	void h2d(void *d, const void *h, size_t bytes) { if (bytes) memcpy(d, h, bytes); }
	void d2h(void *h, const void *d, size_t bytes) { if (bytes) memcpy(h, d, bytes); }
	void zero(void *d, size_t bytes) { if (bytes) memset(d, 0, bytes); }
	void sync() {}
	template <class F> void for_each(size_t n, const F &f, int) { for (size_t i = 0; i < n; ++i) f(i); }
	template <class F> void for_each_in(size_t n, const uint32_t *perm, const F &f, int) { for (size_t k = 0; k < n; ++k) f((size_t)perm[k]); }
	void order_desc(const uint32_t *key, const uint32_t *idx, uint32_t *perm, size_t n)
	{
		for (size_t i = 0; i < n; ++i) perm[i] = idx[i];
		std::stable_sort(perm, perm + n, [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });   // (idx is the identity here)
	}
	void encode(size_t n, const FnEncode &f)                      // strided scratch like the CUDA kernel's (word k of "thread" t at [k * 4 + t])
	{
		uint32_t filter[ENC_FILTER_WORDS * 4];
		for (size_t i = 0; i < n; ++i) f.run(i, filter + (i & 3), 4);
	}
	void text_write(size_t n, const FnText &f)                    // the text kernel's three steps, one record at a time
	{
		for (size_t i = 0; i < n; ++i) {
			JobSink s;
			f.prepare(i, s);
			for (int j = 0; j < s.nj; ++j) for (uint32_t k = 0; k < s.job[j].len; ++k) s.job[j].dst[k] = copy_byte(s.job[j].mode, s.job[j].src, k, s.job[j].len);
			for (int k = 0; k < s.np; ++k) s.p[s.patches[k]] = ',';
		}
	}
	void scan(const uint32_t *in, uint32_t *out, size_t n) { uint32_t run = 0; for (size_t i = 0; i < n; ++i) { const uint32_t v = in[i]; out[i] = run; run += v; } }
	bool ksw(size_t n, const uint8_t *q, const int64_t *qoff, const int32_t *qlen, const uint8_t *t, const int64_t *toff, const int32_t *tlen,
	         int32_t *res, uint32_t *cig, int cigar_cap, std::string &err)
	{
		int64_t qb = 0, tb = 0;
		for (size_t i = 0; i < n; ++i) { qb = std::max<int64_t>(qb, qoff[i] + qlen[i]); tb = std::max<int64_t>(tb, toff[i] + tlen[i]); }
		if (pansvr_ksw_extd2_batch((pansvr_ksw_ctx*)1, (int64_t)n, q, qb, qoff, qlen, t, tb, toff, tlen, &kp, res, cig, cigar_cap) != 0) { err = pansvr_last_error(); return false; }
		return true;
	}
};

struct HostStrTable {
	std::string pool; std::vector<uint32_t> off;
	void build(const std::vector<std::string> &v) { off.assign(v.size() + 1, 0); pool.clear(); for (size_t i = 0; i < v.size(); ++i) { pool += v[i]; off[i + 1] = (uint32_t)pool.size(); } }
	StrTable view() const { StrTable t; t.pool = pool.data(); t.off = off.data(); t.n = (uint32_t)off.size() - 1; return t; }
};
struct StageService { HostBackend be; IndexView view; const uint64_t *pos; RefView rf; std::vector<DevSv> svs; PairIndexView pix; HostStrTable names, prints, ids; };

StageService *stage_service_create(const DebgaIndex &idx, SeedService *seeds, void *, int, std::string &)
{
	StageService *s = new StageService();
	s->view = seeds->view; s->pos = idx.pos.data(); s->rf.ref_seq = idx.ref_seq.data();
	s->svs.resize(idx.sv_info.size());
	for (size_t i = 0; i < s->svs.size(); ++i) { s->svs[i].chr_id = idx.sv_info[i].chr_id; s->svs[i].st_pos = (uint32_t)idx.sv_info[i].st_pos; s->svs[i].end_offset = idx.sv_info[i].end_offset; s->svs[i].pad = 0; }
	s->pix.chr_search_index = idx.chr_search_index.data(); s->pix.chr_end_n = idx.chr_end_n.data(); s->pix.sv = s->svs.data();
	std::vector<std::string> prints(idx.sv_info.size()), ids(idx.sv_info.size());
	for (size_t i = 0; i < idx.sv_info.size(); ++i) { prints[i] = idx.sv_info[i].vcf_print; ids[i] = idx.sv_info[i].vcf_id; }
	s->names.build(idx.target_names); s->prints.build(prints); s->ids.build(ids);
	return s;
}
void stage_service_destroy(StageService *s) { delete s; }
void stage_service_set_scoring(StageService *s, const AlnScores &o, int zdrop)
{
	HostBackend &be = s->be;
	const int8_t m = (int8_t)o.match, x = (int8_t)-o.mismatch;
	for (int a = 0, k = 0; a < 5; ++a) for (int b = 0; b < 5; ++b, ++k) be.mat[k] = (a == 4 || b == 4) ? 0 : (a == b ? m : x);
	be.kp.m = 5; be.kp.mat = be.mat; be.kp.gapo = (int8_t)o.gap_open; be.kp.gape = (int8_t)o.gap_ex; be.kp.gapo2 = (int8_t)o.gap_open2; be.kp.gape2 = (int8_t)o.gap_ex2;
	be.kp.w = 200; be.kp.zdrop = (uint16_t)zdrop; be.kp.end_bonus = -1; be.kp.flag = 0;
}
bool stage_service_run(StageService *s, const DevStageIn &in, DevStageOut &out, std::string &err)
{
	return run_device_stages(s->be, s->view, s->pos, s->rf, s->pix, in, out, err);
}
bool stage_service_finalize(StageService *s, const PairOpts &o, int not_ori, size_t n_pairs, const int32_t *drawn, size_t n_drawn, const uint32_t *host_len,
                            const uint32_t *tie_pair, const DevPairState *tie_done, size_t n_ties, DevStageOut &out,
                            HostVec<char> &text_out, std::string &err)
{
	TextTables T;
	T.target_names = s->names.view(); T.sv_print = s->prints.view(); T.sv_id = s->ids.view(); T.not_ori = not_ori;
	return run_device_finalize(s->be, s->pix, o, T, n_pairs, drawn, n_drawn, host_len, tie_pair, tie_done, n_ties, out, text_out, err);
}

} // namespace pansvr

// ---- ksw through the oracle
typedef double (*batch_fn)(const char*, const char*, int, const uint8_t*, const int64_t*, const int32_t*, const uint8_t*, const int64_t*,
                           const int32_t*, int, const int8_t*, int, int, int, int, int, int, int, int, int, int32_t*, uint32_t*, int);
typedef int64_t (*cells_fn)(int, int, int);
static batch_fn g_batch = nullptr;
static cells_fn g_cells = nullptr;
static std::string g_err;
static bool load_oracle()
{
	if (g_batch) return true;
	const char *p = getenv("PANSVR_ORACLE_SO");
	void *h = dlopen(p ? p : "oracle/libksw_oracle.so", RTLD_NOW | RTLD_LOCAL);
	if (!h) { g_err = dlerror(); return false; }
	g_batch = (batch_fn)dlsym(h, "ksw_batch_run");
	g_cells = (cells_fn)dlsym(h, "ksw_extd2_oracle_cells");
	return g_batch && g_cells;
}
extern "C" const char *pansvr_last_error(void);
extern "C" {
const char *pansvr_last_error(void) { return g_err.c_str(); }
int pansvr_ksw_create(int, pansvr_ksw_ctx **out) { *out = (pansvr_ksw_ctx*)1; return load_oracle() ? 0 : PANSVR_E_CUDA; }
int pansvr_ksw_create_prio(int d, int, pansvr_ksw_ctx **out) { return pansvr_ksw_create(d, out); }
void pansvr_ksw_destroy(pansvr_ksw_ctx*) {}
int pansvr_ksw_last_stats(const pansvr_ksw_ctx*, pansvr_ksw_stats_t *out) { memset(out, 0, sizeof *out); return 0; }
int64_t pansvr_ksw_band_cells(int32_t q, int32_t t, int32_t w) { return load_oracle() ? g_cells(q, t, w) : 0; }
int pansvr_ksw_extd2_batch(pansvr_ksw_ctx*, int64_t n, const uint8_t *qseq, int64_t, const int64_t *qoff, const int32_t *qlen, const uint8_t *tseq,
                           int64_t, const int64_t *toff, const int32_t *tlen, const pansvr_ksw_params_t *p, int32_t *res, uint32_t *cig, int32_t cap)
{
	if (!load_oracle()) return PANSVR_E_CUDA;
	const double s = g_batch("", "", (int)n, qseq, qoff, qlen, tseq, toff, tlen, p->m, p->mat, p->gapo, p->gape, p->gapo2, p->gape2, p->w, p->zdrop,
	                         p->end_bonus, p->flag, 4, res, cig, cap);
	return s < 0 ? PANSVR_E_CUDA : 0;
}
}
