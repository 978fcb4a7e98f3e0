/* oracle/andi_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's all-pairs anchor-distance hot path, written
 * from the behaviour of /root/reference (EvolBioInf/andi v1.15); every function cites the
 * reference file:line it follows. It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check the CUDA path; nothing under andi_b200/ may include,
 * link or call it.
 *
 * Pinning: tests/test_oracle_vs_ref.py compares every function below with the unmodified
 * reference compiled into oracle/_ref/libandi_ref.so (oracle/Makefile, target `ref`), and
 * tests/golden/ holds vectors generated from that library (tests/golden/make_golden.py).
 */
#ifndef ANDI_ORACLE_H
#define ANDI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/model.h:14-32,52-57 : 4x4 substitution counts, index = (subject base << 2) + query base,
 * A=0 C=1 G=2 T=3, plus the query length. 68 bytes. */
typedef struct {
	uint32_t counts[16];
	uint32_t seq_len;
} orc_model;

/* src/esa.h:25-34 : lcp-interval, inclusive bounds, empty = i == j == -1. */
typedef struct {
	int32_t l, i, j, m;
} orc_interval;

/* src/esa.h:42-59 : the enhanced suffix array (same members, own ordering is irrelevant). */
typedef struct {
	const char *S; /* borrowed: RS, RS[len] == '\0' */
	int32_t len;
	int32_t *SA;		 /* len */
	int32_t *LCP;		 /* len + 1 */
	int32_t *CLD;		 /* len + 1 */
	char *FVC;			 /* len */
	orc_interval *cache; /* 4^10 */
} orc_esa;

enum { ORC_RAW = 0, ORC_JC = 1, ORC_KIMURA = 2, ORC_LOGDET = 3, ORC_ANI = 4 }; /* src/global.h:50 */

/* src/sequence.c:260-282 (normalize): keep ACGT!, upper-case acgt, drop the rest, in place.
 * Returns the new length; *non_acgt is set to 1 when something was dropped. */
size_t orc_normalize(char *s, int *non_acgt);

/* src/sequence.c:143-189 (revcomp + catcomp): RS = revcomp(s) '#' s '\0'; out has 2n+2 bytes. */
void orc_make_rs(const char *s, size_t n, char *out);

/* src/sequence.c:196-207 (calc_gc) */
double orc_gc(const char *s, size_t n);

/* src/sequence.c:296-304, 314-373 */
double orc_shustring_cum_prob(size_t x, double p, size_t l);
size_t orc_min_anchor_length(double p, double g, size_t l);

/* src/esa.c:294-304 via divsufsort: suffix array under unsigned byte order. 0 on success. */
int orc_suffix_array(const unsigned char *T, int32_t *SA, int32_t n);

/* src/esa.c:254-277 (esa_init and its five stages). E->S must stay alive. 0 on success. */
int orc_esa_build(orc_esa *E, const char *RS, int32_t len);
void orc_esa_free(orc_esa *E);

/* Spec of src/esa.c:614-624 / 636-656: longest prefix of query[0..qlen)
 * occurring in RS and the exact SA range holding it. m is not part of the spec (set to -1). */
orc_interval orc_get_match(const orc_esa *E, const char *query, size_t qlen);

/* The reference's own search procedure restated (child-table walk + prefix cache),
 * src/esa.c:441-511, 531-601, 636-656; returns the same four fields the reference does. */
orc_interval orc_get_match_cld(const orc_esa *E, const char *query, size_t qlen);
orc_interval orc_get_match_cached(const orc_esa *E, const char *query, size_t qlen);

/* test/test_esa.c:172-203 restated: exhaustive agreement sweep over all 4^depth queries */
size_t orc_sweep_check(const orc_esa *E, int depth);

/* src/process.c:141-214 */
orc_model orc_dist_anchor(const orc_esa *E, const char *query, size_t qlen, size_t threshold,
						  int model_id);
/* same walk with the ESA lookup replaced by the spec search orc_get_match */
orc_model orc_dist_anchor_spec(const orc_esa *E, const char *query, size_t qlen,
							   size_t threshold, int model_id);

/* src/model.c:246-279, 309-337 */
void orc_model_count_equal(orc_model *M, const char *q, size_t len, int model_id);
void orc_model_count(orc_model *M, const char *s, const char *q, size_t len);

/* src/model.c:39-46, 68-73, 81-209 */
orc_model orc_model_average(const orc_model *a, const orc_model *b);
double orc_model_coverage(const orc_model *M);
double orc_estimate(const orc_model *M, int model_id);

/* src/dist_hack.h:34-96 : rows [s_begin, s_end) of the matrix; out has (s_end-s_begin)*n cells.
 * p_value is ANCHOR_P_VALUE (src/andi.c:48). */
int orc_rows(const char *const *seqs, const size_t *lens, size_t n, size_t s_begin, size_t s_end,
			 int model_id, double p_value, orc_model *out);

#ifdef __cplusplus
}
#endif
#endif
