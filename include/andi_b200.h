/* include/andi_b200.h -- C ABI of libandi_b200.so
 *
 * B200-native (sm_100a) replacement for the all-pairs anchor-distance hot path of
 * EvolBioInf/andi v1.15: enhanced-suffix-array construction (src/esa.c), the anchor walk
 * (src/process.c:29-214) and substitution counting (src/model.c:246-337), driven the way
 * src/dist_hack.h:34-96 drives them. Plain C types only; every entry point names the reference
 * interface it stands in for. There is no CPU fallback: every call that computes needs a CUDA
 * device and fails with ANDI_ERR_CUDA otherwise.
 *
 * Layers
 *   1. native handles   andi_ctx / andi_esa / andi_pool_* / andi_dist_*   (this file, part A)
 *   2. the reference's own C surface (esa_init, esa_free, get_match, get_match_cached,
 *      dist_anchor, calculate rows) with the reference's struct layouts     (part B)
 *
 * Thread-safety: a context is single-threaded (one host thread per GPU, as SURVEY 7.3-8
 * recommends); the part-B wrappers serialise on an internal mutex so that the reference's
 * OpenMP callers (src/dist_hack.h:8,16) stay correct.
 */
#ifndef ANDI_B200_H
#define ANDI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ part A: native ABI */

typedef struct andi_ctx andi_ctx; /* one per GPU */
typedef struct andi_esa andi_esa; /* device-resident index of one subject */

/* src/model.h:52-57 `struct model`: counts[(subject_base << 2) + query_base], A0 C1 G2 T3
 * (enum at src/model.h:14-32), then the query length. 68 bytes, same layout. */
typedef struct andi_model {
	uint32_t counts[16];
	uint32_t seq_len;
} andi_model;

/* src/esa.h:25-34 `lcp_inter_t`: inclusive SA bounds, empty = i == j == -1. Same layout. */
typedef struct andi_lcp_inter {
	int32_t l, i, j, m;
} andi_lcp_inter;

/* src/global.h:50 */
enum { ANDI_M_RAW = 0, ANDI_M_JC = 1, ANDI_M_KIMURA = 2, ANDI_M_LOGDET = 3, ANDI_M_ANI = 4 };

enum {
	ANDI_OK = 0,
	ANDI_ERR_ARG = 1,	 /* NULL / out-of-range argument (esa_init returns 1 for these, src/esa.c:255) */
	ANDI_ERR_CUDA = 2,	 /* CUDA runtime failure, including "no device" */
	ANDI_ERR_NOMEM = 3,	 /* host or device allocation failed */
	ANDI_ERR_TOO_LONG = 4 /* sequence longer than (INT_MAX-1)/2, the limit of src/andi.c:296-300 */
};

/* andi_esa_build flags */
enum {
	ANDI_ESA_SEARCH = 0, /* SA + LCP + k-mer directory: all the anchor walk needs */
	ANDI_ESA_FULL = 1	 /* additionally CLD, FVC and the 4^10 prefix cache of src/esa.c:73-245,312-363 */
};

/* Number of usable CUDA devices (0 when there is none or the runtime fails). */
int andi_device_count(void);

/* Create a context on CUDA device `device`. `stream` is a cudaStream_t (e.g. torch's current
 * stream) or NULL for a private stream. */
int andi_ctx_create(int device, void *stream, andi_ctx **out);
void andi_ctx_destroy(andi_ctx *ctx);
/* Message of the last failure on this context (never NULL). ctx may be NULL for create errors. */
const char *andi_last_error(const andi_ctx *ctx);

/* Sequence pool = the `seq_t sequences[n]` array handed to calculate_distances
 * (src/process.h:11). Sequences must already be normalized as src/sequence.c:260-282 does
 * (only A C G T and the contig separator '!'). They are 2-bit packed on the device;
 * GC content (src/sequence.c:196-207) is counted by the same kernel.
 * _host copies from host memory (pageable or pinned); _device takes chars already in HBM
 * (one buffer, sequence k at d_chars + offsets[k]). A new pool replaces the old one. */
int andi_pool_set_host(andi_ctx *ctx, const char *const *seqs, const size_t *lens, size_t n);
int andi_pool_set_device(andi_ctx *ctx, const char *d_chars, const size_t *offsets,
						 const size_t *lens, size_t n);
size_t andi_pool_size(const andi_ctx *ctx);
/* gc as calc_gc; threshold = min_anchor_length(p_value, gc, 2*len+1) (src/sequence.c:210-219,
 * 296-373), evaluated in host double precision exactly like the reference. */
int andi_pool_info(const andi_ctx *ctx, size_t k, size_t *len, double *gc, int *has_separator);
size_t andi_threshold(double p_value, double gc, size_t rs_len);

/* esa_init (src/esa.h:63, src/esa.c:254-277) for pool sequence `subject`: builds
 * RS = revcomp '#' forward on the device (src/sequence.c:143-189), its suffix array
 * (stands where src/esa.c:303 calls divsufsort), LCP (src/esa.c:373-426) and, with
 * ANDI_ESA_FULL, CLD / FVC / cache. */
int andi_esa_build(andi_ctx *ctx, size_t subject, unsigned flags, andi_esa **out);
/* Same from an RS string the caller already built (seq_subject.RS, src/sequence.h:32-46). */
int andi_esa_build_rs(andi_ctx *ctx, const char *rs, size_t rs_len, unsigned flags, andi_esa **out);
/* esa_free (src/esa.h:64, src/esa.c:280-287) */
void andi_esa_free(andi_esa *esa);
int32_t andi_esa_len(const andi_esa *esa);
/* Copy the arrays of src/esa.h:42-59 to host memory. Any pointer may be NULL.
 * SA: len int32; LCP: len+1; CLD: len+1 (CLD[len] := 0, the reference leaves it unwritten);
 * FVC: len chars; cache: 4^10 entries. CLD/FVC/cache need ANDI_ESA_FULL. */
int andi_esa_download(const andi_esa *esa, int32_t *SA, int32_t *LCP, int32_t *CLD, char *FVC,
					  andi_lcp_inter *cache);
/* get_match (src/esa.h:62, src/esa.c:614-624) for a batch of host queries, evaluated on the
 * device with the search the anchor walk uses. out[k] = {l, i, j, -1}: longest prefix of query
 * k occurring in RS and the exact SA range holding it. */
int andi_esa_get_match(const andi_esa *esa, const char *const *queries, const size_t *lens,
					   size_t nq, andi_lcp_inter *out);

/* dist_anchor (src/process.c:141-142) for many queries of the pool against one index:
 * out[k] = dist_anchor(esa, pool[query_ids[k]], len, threshold). `model` selects how anchor
 * interiors are classified (the global MODEL read at src/model.c:247). */
int andi_dist_row(andi_ctx *ctx, const andi_esa *esa, const size_t *query_ids, size_t nq,
				  size_t threshold, int model, andi_model *out);
/* dist_anchor for one host query string (normalized chars). */
int andi_dist_anchor(andi_ctx *ctx, const andi_esa *esa, const char *query, size_t qlen,
					 size_t threshold, int model, andi_model *out);
/* distMatrix / distMatrixLM (src/dist_hack.h:34-96) restricted to subjects [s_begin, s_end):
 * out[(i - s_begin) * n + j] = M(i, j), including the diagonal cells {seq_len 9, AtoA 9}
 * (src/dist_hack.h:61-64). p_value is ANCHOR_P_VALUE. low_memory != 0 keeps one index resident
 * at a time (F_LOW_MEMORY); results are identical either way (test/test_extra.sh:19-22). */
int andi_dist_rows(andi_ctx *ctx, size_t s_begin, size_t s_end, double p_value, int model,
				   int low_memory, andi_model *out);
/* Same, rows left in device memory (d_out is a device pointer): the form the multi-GPU driver
 * uses before the NCCL gather of row blocks, and the one bench.py times as `value`. */
int andi_dist_rows_device(andi_ctx *ctx, size_t s_begin, size_t s_end, double p_value, int model,
						  int low_memory, andi_model *d_out);

/* ---- several GPUs (src/dist_hack.h:8,46-47: the reference spreads the subjects of distMatrix
 * over THREADS OpenMP threads; here they are spread over devices).
 *
 * The packed pool of one context can be handed to another one without a second upload:
 * andi_pool_export describes the planes of `ctx` (device pointers on ctx's device plus the
 * host-side per-sequence facts); andi_pool_import copies them into `ctx` -- over NVLink /
 * cudaMemcpyPeerAsync when `src_device` differs from ctx's device (pass the exporting context's
 * device; -1 = "same device as ctx"). The view's pointers stay valid until the exporting
 * context changes its pool. bench.py's ranks (one process per GPU) move the same planes with an
 * NCCL broadcast and import them from their own device. */
typedef struct andi_pool_view {
	const void *d_code; /* 2-bit code plane, `words` u64 words */
	const void *d_spec; /* separator plane, same geometry; all zero when any_separator == 0 */
	size_t words;
	size_t n;
	const size_t *lens;		  /* n */
	const double *gc;		  /* n */
	const int *has_separator; /* n */
	int any_separator;
} andi_pool_view;
int andi_pool_export(const andi_ctx *ctx, andi_pool_view *out);
int andi_pool_import(andi_ctx *ctx, const andi_pool_view *view, int src_device);

/* distMatrix / distMatrixLM (src/dist_hack.h:34-96) for the whole matrix on `n_devices` GPUs:
 * one host thread per device, the pool uploaded and packed ONCE (on devices[0]) and copied
 * packed to the others, subjects handed out in small batches from a shared queue (their cost
 * varies with length and divergence), rows written straight into out[n * n]. `progress`
 * (may be NULL) is called from the worker threads, serialised, after every batch with the
 * number of (subject, query) pairs finished so far -- what src/dist_hack.h:37-43,74-95 prints.
 * Returns ANDI_OK or the first failing device's error code; its message is copied to errbuf. */
typedef void (*andi_progress_fn)(size_t pairs_done, size_t pairs_total, void *user);
int andi_dist_matrix_multi(const int *devices, int n_devices, const char *const *seqs, const size_t *lens,
						   size_t n, double p_value, int model, int low_memory, andi_model *out,
						   andi_progress_fn progress, void *user, char *errbuf, size_t errbuf_len);

/* Device-side timing of the last andi_dist_rows / andi_esa_build call on this context, taken
 * with CUDA events on the context's stream (milliseconds), plus launch counts. */
typedef struct andi_stats {
	double esa_ms;		   /* all index-construction kernels */
	double walk_ms;		   /* anchor-walk kernels only */
	double total_ms;	   /* first launch to last completion */
	uint64_t esa_launches; /* OUR kernels launched for index construction (CUB calls not counted) */
	uint64_t cub_calls;	   /* cub::DeviceRadixSort / DeviceScan / DeviceSelect calls */
	uint64_t walk_launches;
	uint64_t pairs;		/* ordered (subject, query) pairs walked */
	uint64_t subjects;	/* indexes built */
	uint64_t sa_rounds; /* prefix-doubling refinement rounds, summed over subjects */
	uint64_t h2d_bytes, d2h_bytes;
	uint64_t p2p_bytes; /* packed pool planes received from another device (andi_pool_import) */
	double rows_ms;		/* device wall time of the andi_dist_rows calls (esa_ms and walk_ms are per-kernel
						 * sums and overlap when two subjects are in flight) */
} andi_stats;
int andi_get_stats(const andi_ctx *ctx, andi_stats *out);
void andi_reset_stats(andi_ctx *ctx);

/* ------------------------------------------- part B: the reference's own C surface
 * Same names, argument meaning and error behaviour as the reference, so its driver
 * (src/dist_hack.h) and unit tests (test/test_esa.c) link against this library unchanged.
 * Struct layouts are those of src/esa.h:25-59, src/sequence.h:32-46 and src/model.h:52-57;
 * they are declared in include/andi_compat.h to keep this header free of name clashes. */

#ifdef __cplusplus
}
#endif
#endif
