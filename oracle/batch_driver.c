/*
 * batch_driver.c -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Runs a list of ksw tasks through any function that has the reference's ksw_extd2_sse
 * signature (/root/reference/src/kswlib/ksw2.h:63-64) on a pool of pthreads, one
 * ksw_extz_t per thread, exactly how the reference's workers use it
 * (/root/reference/src/PanSVgenerateVCF/read_realignment.cpp:872-891: one KSW_ALN_handler
 * with one `ez` per thread, reused across calls).  The function is either the scalar
 * restatement in this directory or the reference's own object in oracle/_ref/libksw_ref.so
 * (resolved with dlopen, so nothing of the reference is linked in here).
 *
 * Used by tests/ (as the checker) and by bench.py's cpu_baseline / --impl reference legs.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <dlfcn.h>
#include <pthread.h>
#include <time.h>

typedef struct {
	uint32_t max:31, zdropped:1;
	int max_q, max_t, mqe, mqe_t, mte, mte_q, score, m_cigar, n_cigar, reach_end;
	uint32_t *cigar;
} drv_extz_t;

typedef void (*ksw_fn_t)(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                         const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop,
                         int end_bonus, int flag, drv_extz_t *ez);

extern void ksw_extd2_oracle(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                             const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop,
                             int end_bonus, int flag, void *ez);

#define RES_WORDS 12 /* max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar reach_end status */

typedef struct {
	ksw_fn_t fn;
	int n;
	const uint8_t *qseq; const int64_t *qoff; const int32_t *qlen;
	const uint8_t *tseq; const int64_t *toff; const int32_t *tlen;
	int8_t m; const int8_t *mat; int8_t q, e, q2, e2;
	int w, zdrop, end_bonus, flag;
	int32_t *res; uint32_t *cigar; int cigar_cap;
	volatile int64_t *next;
} job_t;

static void *worker(void *arg)
{
	job_t *j = (job_t*)arg;
	drv_extz_t ez;
	memset(&ez, 0, sizeof(ez));
	for (;;) {
		int64_t b = __sync_fetch_and_add(j->next, 64), i;
		if (b >= j->n) break;
		for (i = b; i < b + 64 && i < j->n; ++i) {
			int32_t *o = j->res + (size_t)i * RES_WORDS;
			j->fn(0, j->qlen[i], j->qseq + j->qoff[i], j->tlen[i], j->tseq + j->toff[i],
			      j->m, j->mat, j->q, j->e, j->q2, j->e2, j->w, j->zdrop, j->end_bonus, j->flag, &ez);
			o[0] = (int32_t)ez.max; o[1] = (int32_t)ez.zdropped; o[2] = ez.max_q; o[3] = ez.max_t;
			o[4] = ez.mqe; o[5] = ez.mqe_t; o[6] = ez.mte; o[7] = ez.mte_q; o[8] = ez.score;
			o[9] = ez.n_cigar; o[10] = ez.reach_end; o[11] = ez.n_cigar > j->cigar_cap;
			if (j->cigar && j->cigar_cap > 0) {
				int k, nc = ez.n_cigar < j->cigar_cap ? ez.n_cigar : j->cigar_cap;
				uint32_t *c = j->cigar + (size_t)i * j->cigar_cap;
				for (k = 0; k < nc; ++k) c[k] = ez.cigar[k];
				for (; k < j->cigar_cap; ++k) c[k] = 0;
			}
		}
	}
	free(ez.cigar);
	return 0;
}

/* lib_path == NULL or "" -> the scalar restatement; otherwise dlopen(lib_path) and use `symbol`.
 * Returns wall seconds of the compute region (threads started to threads joined), <0 on error. */
double ksw_batch_run(const char *lib_path, const char *symbol, int n,
                     const uint8_t *qseq, const int64_t *qoff, const int32_t *qlen,
                     const uint8_t *tseq, const int64_t *toff, const int32_t *tlen,
                     int m, const int8_t *mat, int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag,
                     int n_threads, int32_t *res, uint32_t *cigar, int cigar_cap)
{
	job_t j;
	pthread_t *th;
	volatile int64_t next = 0;
	struct timespec t0, t1;
	void *h = 0;
	int i;
	memset(&j, 0, sizeof(j));
	if (lib_path && lib_path[0]) {
		h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
		if (!h) { fprintf(stderr, "ksw_batch_run: %s\n", dlerror()); return -1.0; }
		j.fn = (ksw_fn_t)dlsym(h, symbol);
		if (!j.fn) { fprintf(stderr, "ksw_batch_run: symbol %s not found\n", symbol); return -2.0; }
	} else j.fn = (ksw_fn_t)ksw_extd2_oracle;
	j.n = n; j.qseq = qseq; j.qoff = qoff; j.qlen = qlen; j.tseq = tseq; j.toff = toff; j.tlen = tlen;
	j.m = (int8_t)m; j.mat = mat; j.q = (int8_t)q; j.e = (int8_t)e; j.q2 = (int8_t)q2; j.e2 = (int8_t)e2;
	j.w = w; j.zdrop = zdrop; j.end_bonus = end_bonus; j.flag = flag;
	j.res = res; j.cigar = cigar; j.cigar_cap = cigar_cap; j.next = &next;
	if (n_threads < 1) n_threads = 1;
	th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (i = 1; i < n_threads; ++i) pthread_create(&th[i], 0, worker, &j);
	worker(&j);
	for (i = 1; i < n_threads; ++i) pthread_join(th[i], 0);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(th);
	/* the library is left loaded on purpose: dlclose+dlopen per call would dominate small batches */
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
