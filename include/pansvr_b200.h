/*
 * pansvr_b200.h -- C ABI of the B200-native `panSVR aln` realignment hot path.
 *
 * Plain C types only; every entry point names the reference interface it replaces.
 * Reference paths are relative to the hitbc/panSVR tree.
 *
 * Error model: the reference has no error codes on this path (ksw returns void and signals
 * trouble through ez->zdropped / KSW_NEG_INF; src/kswlib/ksw2_extd2_sse.c:68,93).  Entry points
 * that can fail for reasons the reference does not have (no GPU, out of device memory) return
 * 0 on success and a negative PANSVR_E_* code otherwise; pansvr_last_error() holds the text.
 * There is no CPU fallback: without a usable B200 every compute entry point fails loudly.
 */
#ifndef PANSVR_B200_H_
#define PANSVR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PANSVR_E_CUDA      (-1)   /* a CUDA call failed (no device, OOM, launch error) */
#define PANSVR_E_ARG       (-2)   /* invalid argument */
#define PANSVR_E_UNSUPPORTED (-3) /* parameter set outside what the device kernels implement */

/* flags of src/kswlib/ksw2.h:9-15 (same values) */
#define PANSVR_KSW_SCORE_ONLY  0x01
#define PANSVR_KSW_RIGHT       0x02
#define PANSVR_KSW_GENERIC_SC  0x04
#define PANSVR_KSW_APPROX_MAX  0x08
#define PANSVR_KSW_APPROX_DROP 0x10
#define PANSVR_KSW_EXTZ_ONLY   0x40
#define PANSVR_KSW_REV_CIGAR   0x80
#define PANSVR_KSW_NEG_INF     (-0x40000000)

/* Same layout as the reference's ksw_extz_t (src/kswlib/ksw2.h:26-35). */
typedef struct {
	uint32_t max:31, zdropped:1;
	int max_q, max_t;
	int mqe, mqe_t;
	int mte, mte_q;
	int score;
	int m_cigar, n_cigar;
	int reach_end;
	uint32_t *cigar;
} pansvr_ksw_extz_t;

/* One result row of a batch: the fields of ksw_extz_t as 12 int32 words. */
enum {
	PANSVR_RES_MAX = 0, PANSVR_RES_ZDROPPED, PANSVR_RES_MAX_Q, PANSVR_RES_MAX_T, PANSVR_RES_MQE, PANSVR_RES_MQE_T,
	PANSVR_RES_MTE, PANSVR_RES_MTE_Q, PANSVR_RES_SCORE, PANSVR_RES_N_CIGAR, PANSVR_RES_REACH_END,
	PANSVR_RES_STATUS,           /* bit 0: CIGAR longer than cigar_cap (n_cigar is still exact) */
	PANSVR_RES_WORDS
};

/* Scoring / band parameters of one batch: the trailing arguments of ksw_extd2_sse
 * (src/kswlib/ksw2.h:63-64); fc_aln passes one set for the whole run
 * (src/PanSVgenerateVCF/read_realignment.cpp:817-827,889). */
typedef struct {
	int32_t m;              /* alphabet size; last symbol is the wildcard */
	const int8_t *mat;      /* m*m scores (host pointer) */
	int8_t gapo, gape, gapo2, gape2;
	int32_t w, zdrop, end_bonus, flag;
} pansvr_ksw_params_t;

/* Counters of the last batch call on a context (for bench.py's gpu_launches / roofline). */
typedef struct {
	int64_t kernel_launches;    /* our kernels launched */
	int64_t tasks_fast_wrap, tasks_fast_nowrap, tasks_generic, tasks_trivial;
	double kernel_ms;           /* CUDA-event time of the kernels only, on the context's stream */
	double total_ms;            /* CUDA-event time of the whole call: H2D + kernels + D2H */
	int64_t h2d_bytes, d2h_bytes;
	int64_t tb_bytes_per_warp, resident_warps;
} pansvr_ksw_stats_t;

typedef struct pansvr_ksw_ctx pansvr_ksw_ctx;

/* One context = one GPU + one stream + reusable device scratch.  Not thread-safe; use one per
 * host thread, like the reference's per-thread KSW_ALN_handler (read_realignment.hpp:183-275). */
int  pansvr_ksw_create(int device, pansvr_ksw_ctx **out);
void pansvr_ksw_destroy(pansvr_ksw_ctx *ctx);
const char *pansvr_last_error(void);

/* Pinned host memory for the caller's batch buffers (the read/candidate-window batcher). */
void *pansvr_host_alloc(size_t bytes);
void  pansvr_host_free(void *p);

/*
 * Batched ksw_extd2_sse: task i aligns query qseq[qoff[i] .. +qlen[i]) against target
 * tseq[toff[i] .. +tlen[i]) (one base per byte, values < m) and fills results[i*12 .. +12) and
 * cigar[i*cigar_cap .. +n_cigar) exactly as the reference would fill ksw_extz_t for that call.
 * All pointers are HOST pointers; copies to and from the device happen inside.
 * Replaces: the per-call loop around KSW_ALN_handler::align_non_splice
 * (read_realignment.cpp:872-891) -> ksw_extd2_sse (ksw2_extd2_sse.c:26).
 */
int pansvr_ksw_extd2_batch(pansvr_ksw_ctx *ctx, int64_t n,
                           const uint8_t *qseq, int64_t qseq_bytes, const int64_t *qoff, const int32_t *qlen,
                           const uint8_t *tseq, int64_t tseq_bytes, const int64_t *toff, const int32_t *tlen,
                           const pansvr_ksw_params_t *params,
                           int32_t *results, uint32_t *cigar, int32_t cigar_cap);

/* Same, with every array already resident in device memory (DEVICE pointers) and the results
 * left there.  The per-task plan (kernel variant, order) is made on the device from the lengths
 * there; h_qlen / h_tlen (host copies of the lengths) are not read any more and may be NULL.
 * Used when the sequences were produced on the GPU (seed/chain stage) and for kernel-only timing. */
int pansvr_ksw_extd2_batch_device(pansvr_ksw_ctx *ctx, int64_t n,
                                  const uint8_t *d_qseq, const int64_t *d_qoff, const int32_t *d_qlen,
                                  const uint8_t *d_tseq, const int64_t *d_toff, const int32_t *d_tlen,
                                  const int32_t *h_qlen, const int32_t *h_tlen,
                                  const pansvr_ksw_params_t *params,
                                  int32_t *d_results, uint32_t *d_cigar, int32_t cigar_cap);

int pansvr_ksw_last_stats(const pansvr_ksw_ctx *ctx, pansvr_ksw_stats_t *out);

/* Measured 32-bit integer-ALU throughput of the device in Gop/s (IADD3/LOP3/VIMNMX chains, no
 * memory): the denominator of the DP kernel's roofline (cells/s x 55 ops per cell). */
int pansvr_int_alu_peak(pansvr_ksw_ctx *ctx, double *gops);
/* The same measurement per pipe, so that the yardstick can be read against the hardware: out[0] = the mixed chain above,
 * out[1] = ALU pipe only (LOP3 / VIMNMX chains), out[2] = FMA pipe only (IMAD chains), out[3] = both pipes fed at once
 * (independent LOP3 and IMAD chains: the issue limit).  Gop/s, one op = one 32-bit lane operation. */
int pansvr_int_pipe_peaks(pansvr_ksw_ctx *ctx, double out[4]);

/* In-band DP cells of one task, the work unit GCUPS is quoted in (ksw2_extd2_sse.c:131-138). */
int64_t pansvr_ksw_band_cells(int32_t qlen, int32_t tlen, int32_t w);

/*
 * Drop-in for the reference kernel, identical signature and ownership rules
 * (src/kswlib/ksw2.h:63-64): `km` is ignored, `ez` is caller-owned and reused, ez->cigar grows
 * with realloc and is never freed here.  A batch of one on a process-wide context for the
 * current device; meant for parity checks and for callers that have not been batched yet.
 * Aborts with a message if no GPU is usable (the reference signature cannot carry an error).
 */
void pansvr_ksw_extd2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                      const int8_t *mat, int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop,
                      int end_bonus, int flag, pansvr_ksw_extz_t *ez);
/* The same function under the reference's own symbol name, so that the reference links
 * against libpansvr_b200.so instead of compiling src/kswlib/ksw2_extd2_sse.c. */
void ksw_extd2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                   const int8_t *mat, int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop,
                   int end_bonus, int flag, pansvr_ksw_extz_t *ez);

/* ------------------------------------------------------------------------------------------------
 * The aln stage itself (`panSVR fc_aln`): seed -> chain -> ksw -> pairing -> SAM, one block of read pairs
 * per call.  Replaces deCOY_CLASSIFY_MAIN::init_run / classify_pipeline step 1 / align_read_pair
 * (src/PanSVgenerateVCF/read_realignment.cpp:26-176, 745-799).  Output is defined against
 * `fc_aln -t 1` (the reference is only deterministic single-threaded).
 */
typedef struct {                /* MAP_PARA, read_realignment.hpp:43-128; 0 in every field = the reference's defaults */
	int32_t match, mismatch, gap_open, gap_ex, gap_open2, gap_ex2, zdrop, band_width;
	int32_t not_ori;            /* -Q */
	int32_t max_use_read;       /* -R */
	int32_t threads;            /* -t: host helper threads of the parallel stages; 0 = all cores (max 48, the reference's limit).  The output is that of `-t 1`. */
	int32_t explicit_mask;      /* bit k set: field k of this struct (match = 0, mismatch = 1, ...) is meant as given even when it is 0 */
} pansvr_aln_options_t;

typedef struct {
	int64_t reads, mems, ksw_tasks, ksw_cells, deferred_pairs;
	double stage_seconds[8];    /* A encode/census, B seeding (GPU), C merge/expand/chain, D ksw planning, E ksw (GPU), F replay + SAM text,
	                               FASTQ parse, output assembly */
	/* device side, summed over all GPUs of the context since the last reset */
	int64_t kernel_launches;    /* our kernels launched */
	int64_t h2d_bytes, d2h_bytes;
	int64_t seed_probes;        /* k-mer lookups made by the seeding kernels */
	double seed_kernel_ms, ksw_kernel_ms, stage_kernel_ms;   /* CUDA-event time of the seeding / ksw / other stage kernels */
	double stage_kernel_ms_by[8];   /* the stage kernels by group: 0 records + original alignments + encode/census, 1 seeding, 2 merge + chain,
	                                   3 ksw planning, 4 candidate resolution, 5 cell count, 6 pairing probe + finalize, 7 SAM text */
	/* the in-order passes (the only part of a block that waits for the block before it: the reference's rand() stream) */
	double in_order_seconds;    /* time inside the in-order sections, waiting for the turn not included */
	int64_t in_order_pairs;     /* pairs that drew at least one random number */
	int64_t in_order_draws;     /* rand() calls replayed */
	int64_t host_pairs;         /* pairs the device path handed to the host path ('N', random_r sampling) */
	int64_t tie_pairs;          /* pairs whose rand() ties decide their outcome: finished by the in-order pass from the device's seeds, chains and candidates */
} pansvr_aln_stats_t;

typedef struct pansvr_aln_ctx pansvr_aln_ctx;

/* Loads the deBGA index directory (the 8 files `deBGA index -k 22` writes + unipath.chr) and the header SAM of the
 * original BAM, uploads the index to `device`.  Replaces deBGA_INDEX::load_index_file (deBGA_index.cpp:33-80). */
int  pansvr_aln_create(const char *index_dir, const char *header_sam, const pansvr_aln_options_t *opt, int device, pansvr_aln_ctx **out);
/* The same on several GPUs of one box: every device gets a replica of the index, the sub-blocks of a block are dealt to the devices
 * round robin (each on its own host thread and stream), the results are merged by pair index as on one device -- the reference's
 * kt_for fan-out over a block and its ordered write (read_realignment.cpp:114,160,165-176).  `-d 0,1,2,3` on the command line. */
int  pansvr_aln_create_multi(const char *index_dir, const char *header_sam, const pansvr_aln_options_t *opt, const int *devices, int n_devices, pansvr_aln_ctx **out);
void pansvr_aln_destroy(pansvr_aln_ctx *ctx);
/* Header text the reference writes in front of its output (the original header, verbatim). */
const char *pansvr_aln_header_text(const pansvr_aln_ctx *ctx);
const char *pansvr_aln_last_error(void);
/* One block: `fastq` holds interleaved pairs in the wire format of `fc_signal` (4-line FASTQ, alignment of the original
 * BAM in the comment).  *sam / *ori receive malloc'ed, NUL-terminated SAM body text (records only) of the main output
 * and of the `-p` output; free with pansvr_free.  The rand() replay state carries over from block to block. */
int  pansvr_aln_block(pansvr_aln_ctx *ctx, const char *fastq, size_t fastq_bytes, char **sam, size_t *sam_bytes, char **ori, size_t *ori_bytes);
/* Same block, BAM records instead of text: *bam / *ori receive the concatenated records exactly as htslib's bam_write1
 * hands them to BGZF after sam_parse1 (little-endian [block_size][refID,pos,bin_mq_nl,flag_nc,l_seq,next_refID,next_pos,tlen]
 * [read_name][cigar][seq][qual][aux]); replaces sam_parse1 in output_BAM / output_ori_bam (read_realignment.cpp:479-536,656-717). */
int  pansvr_aln_block_bam(pansvr_aln_ctx *ctx, const char *fastq, size_t fastq_bytes, uint8_t **bam, size_t *bam_bytes, uint8_t **ori, size_t *ori_bytes);
/* BAM file writer: hts_open(path, "wb") + sam_hdr_write, sam_write1 per record, hts_close (read_realignment.cpp:85-94,
 * 165-175).  BGZF blocks are cut where htslib 1.9 cuts them and deflated at zlib's default level on the context's helper
 * threads, so the file equals the reference's byte for byte (same zlib).  `records` = output of pansvr_aln_block_bam. */
typedef struct pansvr_bam_file pansvr_bam_file;
int  pansvr_bam_open(pansvr_aln_ctx *ctx, const char *path, pansvr_bam_file **out);
int  pansvr_bam_write(pansvr_bam_file *f, const uint8_t *records, size_t bytes);
int  pansvr_bam_close(pansvr_bam_file *f);
int  pansvr_aln_last_stats(const pansvr_aln_ctx *ctx, pansvr_aln_stats_t *out);
/* Puts the replay back to the state of a freshly started `fc_aln` (rand() streams, counters); the index stays resident. */
int  pansvr_aln_reset(pansvr_aln_ctx *ctx);
void pansvr_free(void *p);
/* ---- one input sharded over several processes / GPUs (reads shard trivially; results are merged by pair index, like the
 * reference's kt_for fan-out and ordered write, read_realignment.cpp:114,160,165-176).  A process that realigns pairs [b, e) of an
 * input must see (1) the STAT_ fields of the input's FIRST record, which the reference parses once (read_realignment.cpp:134-148):
 * pansvr_aln_prime_read_stats with the head of the input; (2) the libc random streams (rand(), random_r()) as the process
 * handling the pairs before b left them: pansvr_aln_await_state(path) makes the context's first in-order pass wait for the file
 * `path` and continue from the state in it -- all other stages of its blocks run meanwhile -- and pansvr_aln_publish_state(path)
 * writes the state after the context's last in-order pass (atomically: readers never see a partial file).  No collective, no
 * device traffic between processes. */
int  pansvr_aln_prime_read_stats(pansvr_aln_ctx *ctx, const char *fastq_head, size_t bytes);
/* Block-cyclic form of the same, for one input dealt to N processes piece by piece (piece b goes to process b mod N): one call
 * realigns all pieces this process owns, a few sub-blocks in flight across piece boundaries.  A piece's first in-order pass waits
 * for `await_path` (NULL: none -- the first piece of the input) and its last one writes `publish_path` (NULL: none) the moment it
 * is done, before the piece's remaining stages: the in-order passes of consecutive pieces follow each other through the files while
 * everything else of every piece overlaps.  *sam / *ori: the texts of the pieces one behind the other; sam_bytes / ori_bytes of
 * each piece = its share.  (pansvr_aln_block = one piece without files.) */
typedef struct {
	const char *fastq; size_t fastq_bytes;
	const char *await_path, *publish_path;
	size_t sam_bytes, ori_bytes;                /* out */
} pansvr_aln_piece_t;
int  pansvr_aln_pieces(pansvr_aln_ctx *ctx, pansvr_aln_piece_t *pieces, int n_pieces, char **sam, size_t *sam_bytes, char **ori, size_t *ori_bytes);
int  pansvr_aln_await_state(pansvr_aln_ctx *ctx, const char *path);
int  pansvr_aln_publish_state(pansvr_aln_ctx *ctx, const char *path);
/* Same command line as `panSVR fc_aln` (classify_main, src/main.cpp:18-25): [options] <IndexDir> <reads.fq|-> <header.sam>.
 * Writes BAM, or SAM text with -S, like the reference; -d <gpu>[,<gpu>...] selects the device(s); -t is the number of host helper threads
 * (the output is that of the reference's `-t 1`, the only deterministic mode). */
int  pansvr_fc_aln_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif /* PANSVR_B200_H_ */
