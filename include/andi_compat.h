/* include/andi_compat.h -- part B of the C ABI: the reference's own C surface.
 *
 * libandi_b200.so exports the functions below with the reference's names, argument meaning
 * and struct layouts, so that the reference's driver (src/dist_hack.h:47-90) and unit tests
 * (test/test_esa.c) can link against it in place of src/esa.c + src/process.c. The structs are
 * redeclared here (rather than including the reference's headers) so this file is
 * self-contained; layouts are those of
 *     lcp_inter_t  src/esa.h:25-34        esa_s        src/esa.h:42-59
 *     seq_subject  src/sequence.h:32-46   model        src/model.h:52-57
 * A program that already includes the reference's headers must NOT include this one.
 *
 * Behaviour
 *   esa_init        builds the index on the GPU (device 0 unless ANDI_B200_DEVICE is set) and
 *                   fills SA, LCP, CLD, FVC and cache with host copies that are bit-identical
 *                   to the reference's (CLD[len] excepted, which the reference never writes).
 *                   Returns 0 on success, 1 on NULL arguments (src/esa.c:255), a non-zero
 *                   code on device failure. C->S borrows S->RS exactly like the reference.
 *   esa_free        frees the host arrays and the device index, zeroes the struct
 *                   (src/esa.c:280-287); calling it twice is harmless.
 *   get_match       longest-prefix search evaluated on the device index; {l,i,j} as in
 *   get_match_cached src/esa.c:614-656 (m is -1). Returns {-1,-1,-1,-1} for an ESA that was
 *                   not built by this library (src/esa.c:616-618).
 *   dist_anchor     src/process.c:141-142. Reads the caller's global MODEL when the executable
 *                   defines one (weak reference), else ANDI_M_JC / andi_compat_set_model().
 * All five serialise on one mutex: the reference calls them from OpenMP threads
 * (src/dist_hack.h:8,16); one GPU context does the work.
 */
#ifndef ANDI_COMPAT_H
#define ANDI_COMPAT_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t saidx_t;

typedef struct {
	saidx_t l, i, j, m;
} lcp_inter_t;

typedef struct esa_s {
	const char *S;
	saidx_t *SA;
	saidx_t *LCP;
	saidx_t len;
	lcp_inter_t *cache;
	char *FVC;
	saidx_t *CLD;
} esa_s;

typedef struct seq_subject {
	char *RS;
	size_t RSlen;
	double gc;
	size_t threshold;
} seq_subject;

typedef struct model {
	unsigned int counts[16];
	unsigned int seq_len;
} model;

/* src/sequence.h:19-26 */
typedef struct seq_s {
	char *S;
	size_t len;
	char *name;
} seq_t;

int esa_init(esa_s *C, const seq_subject *S);
void esa_free(esa_s *C);
lcp_inter_t get_match(const esa_s *C, const char *query, size_t qlen);
lcp_inter_t get_match_cached(const esa_s *C, const char *query, size_t qlen);
model dist_anchor(const esa_s *C, const char *query, size_t query_length, size_t threshold);

/* The driver level (SURVEY 8b "the real offload point"): distMatrix / distMatrixLM
 * (src/dist_hack.h:34-96) fill M[n * n] for all subjects through ONE batched GPU run -- pool packed
 * once, subjects spread over the devices of ANDI_B200_DEVICES (e.g. "0-7"; default: the device of
 * ANDI_B200_DEVICE, else 0) -- and print the reference's progress line when FLAGS has
 * F_PRINT_PROGRESS. They read the host program's globals ANCHOR_P_VALUE, MODEL and FLAGS
 * (src/global.h:20-48; link the program with -rdynamic so the library sees them). The two differ
 * only in what they promise about memory; one index is resident per GPU in both.
 * calculate_distances (src/process.h:11, src/process.c:230-321) is the reference's own sequence:
 * allocate M, fill it, print_distances / print_coverages / bootstrap matrices through the HOST
 * PROGRAM's own io.c and model.c (print_distances, print_coverages, model_average,
 * model_bootstrap are taken from the executable). With it src/esa.c and src/process.c can be
 * dropped from the reference's build altogether. */
void distMatrix(model *M, const seq_t *sequences, size_t n);
void distMatrixLM(model *M, const seq_t *sequences, size_t n);
void calculate_distances(seq_t *sequences, size_t n);

/* Used when the hosting executable does not define the reference's global `int MODEL`. */
void andi_compat_set_model(int model_id);

#ifdef __cplusplus
}
#endif
#endif
