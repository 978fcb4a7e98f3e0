/* andi_b200/host/andi_host.h -- host-side pieces of the andi command line that stay on the
 * CPU: FASTA ingest, normalisation, join mode, estimators, PHYLIP output, bootstrap.
 * Plain C; calls the GPU only through include/andi_b200.h. Reference lines are cited at each
 * function in the .c files. */
#ifndef ANDI_HOST_H
#define ANDI_HOST_H

#include "../../include/andi_b200.h"
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

/* src/global.h:56-67 */
enum {
	HF_TRUNCATE_NAMES = 1,
	HF_VERBOSE = 2,
	HF_EXTRA_VERBOSE = 4,
	HF_NON_ACGT = 8,
	HF_JOIN = 16,
	HF_LOW_MEMORY = 32,
	HF_SHORT = 64,
	HF_PRINT_PROGRESS = 128,
	HF_SOFT_ERROR = 256
};

typedef struct {
	char *S;	 /* normalized: A C G T and '!' between joined contigs */
	size_t len;
	char *name;
} host_seq;

typedef struct {
	host_seq *data;
	size_t size, capacity;
} host_seqs;

typedef struct {
	int flags;
	int model;				  /* ANDI_M_* */
	double p_value;			  /* ANCHOR_P_VALUE, src/andi.c:48 */
	unsigned long bootstrap;  /* number of extra matrices (b - 1), src/andi.c:198 */
	unsigned long seed;		  /* --seed; without it the CLI seeds with time(NULL) like the reference */
	int device;
} host_config;

/* fasta.c */
void seqs_init(host_seqs *v);
void seqs_push(host_seqs *v, host_seq s);
void seqs_free(host_seqs *v);
/* Reads every record of a FASTA file ("-" = stdin) into out; returns 0 on success, sets
 * HF_NON_ACGT in *flags when characters were stripped. Errors are reported like the
 * reference does (warn + soft error flag) and the file is skipped. */
int fasta_read(const char *file_name, host_seqs *out, int *flags);
/* src/io.c:159-189 + src/sequence.c:78-125: all records of one file glued with '!' */
int fasta_read_join(const char *file_name, host_seqs *out, int *flags);
/* The same two without printing: a failure leaves the warning text in msg. The command line reads
 * its files with all cores (one core parses 0.3 GB/s, a 3085-genome pool is 6.5 GB)
 * and then reports in file order. */
int fasta_read_quiet(const char *file_name, host_seqs *out, int *flags, char *msg, size_t msg_len);
int fasta_read_join_quiet(const char *file_name, host_seqs *out, int *flags, char *msg, size_t msg_len);

/* model_host.c : src/model.c:39-209 */
andi_model model_average(const andi_model *a, const andi_model *b);
double model_coverage(const andi_model *m);
double model_estimate(const andi_model *m, int model_id);
/* src/model.c:222-232 with an own MT19937 + multinomial from exact conditional binomials
 * (PARITY UNPINNED: no GSL here, so the distribution is the reference's, the stream is not) */
typedef struct host_rng host_rng;
host_rng *host_rng_new(unsigned long seed);
void host_rng_free(host_rng *r);
uint32_t host_rng_binomial(host_rng *r, double p, uint32_t n); /* exact: inversion / BTPE */
andi_model host_model_bootstrap(host_rng *r, andi_model datum);

/* output.c : src/io.c:246-338 */
void host_print_distances(FILE *out, const andi_model *M, const host_seqs *seqs, const host_config *cfg, int warnings,
					 int *flags);
void host_print_coverages(FILE *out, const andi_model *M, size_t n);

#endif
