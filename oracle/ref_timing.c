/* oracle/ref_timing.c -- TEST INFRASTRUCTURE ONLY.
 * Thin timing harness around the reference's own entry points (linked into
 * oracle/_ref/libandi_ref.so only). It calls the unmodified reference functions
 * seq_subject_init / esa_init / dist_anchor / esa_free exactly as the reference's driver does
 * (src/dist_hack.h:47-90) so bench.py's cpu_baseline / --impl reference leg can time the hot
 * path on in-memory sequences without going through FASTA files.
 */
#define _GNU_SOURCE
#include "esa.h"
#include "global.h"
#include "model.h"
#include "sequence.h"
#include <omp.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

model dist_anchor(const esa_s *C, const char *query, size_t query_length, size_t threshold);

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Rows [s_begin, s_end) of the all-pairs matrix over n in-memory sequences, the way
 * distMatrix (FAST mode: omp over subjects) does it. out: (s_end-s_begin) * n models.
 * t_out[0] = wall seconds total, t_out[1] = summed esa_init seconds (all threads),
 * t_out[2] = summed dist_anchor seconds (all threads). Returns 0 on success. */
int ref_rows(const char *const *seqs, const size_t *lens, size_t n, size_t s_begin, size_t s_end,
			 int threads, int model_id, double p_value, model *out, double *t_out) {
	MODEL = model_id;
	ANCHOR_P_VALUE = p_value;
	double t_esa = 0.0, t_walk = 0.0;
	int fail = 0;
	double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : t_esa, t_walk)
	for (size_t i = s_begin; i < s_end; i++) {
		seq_t base = {.S = (char *)seqs[i], .len = lens[i], .name = (char *)"x"};
		seq_subject subject;
		esa_s E;
		double a = now_s();
		if (seq_subject_init(&subject, &base) || esa_init(&E, &subject)) {
#pragma omp atomic write
			fail = 1;
			continue;
		}
		double b = now_s();
		t_esa += b - a;
		for (size_t j = 0; j < n; j++) {
			model *cell = &out[(i - s_begin) * n + j];
			if (j == i) {
				*cell = (model){.seq_len = 9, .counts = {9}};
				continue;
			}
			*cell = dist_anchor(&E, seqs[j], lens[j], subject.threshold);
		}
		t_walk += now_s() - b;
		esa_free(&E);
		seq_subject_free(&subject);
	}
	if (t_out) {
		t_out[0] = now_s() - t0;
		t_out[1] = t_esa;
		t_out[2] = t_walk;
	}
	return fail;
}

/* Low-memory mode shape (distMatrixLM, src/dist_hack.h:16,59-60): subjects serial, omp over
 * queries sharing one ESA. */
int ref_rows_lm(const char *const *seqs, const size_t *lens, size_t n, size_t s_begin,
				size_t s_end, int threads, int model_id, double p_value, model *out,
				double *t_out) {
	MODEL = model_id;
	ANCHOR_P_VALUE = p_value;
	double t_esa = 0.0, t_walk = 0.0;
	double t0 = now_s();
	for (size_t i = s_begin; i < s_end; i++) {
		seq_t base = {.S = (char *)seqs[i], .len = lens[i], .name = (char *)"x"};
		seq_subject subject;
		esa_s E;
		double a = now_s();
		if (seq_subject_init(&subject, &base) || esa_init(&E, &subject)) return 1;
		double b = now_s();
		t_esa += b - a;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
		for (size_t j = 0; j < n; j++) {
			model *cell = &out[(i - s_begin) * n + j];
			if (j == i) {
				*cell = (model){.seq_len = 9, .counts = {9}};
				continue;
			}
			*cell = dist_anchor(&E, seqs[j], lens[j], subject.threshold);
		}
		t_walk += now_s() - b;
		esa_free(&E);
		seq_subject_free(&subject);
	}
	if (t_out) {
		t_out[0] = now_s() - t0;
		t_out[1] = t_esa;
		t_out[2] = t_walk;
	}
	return 0;
}

/* Seconds spent in each esa_init stage cannot be separated without touching the reference
 * sources, so only the whole of esa_init is timed here (shim SA included). */
int ref_sizeof_model(void) { return (int)sizeof(model); }
int ref_sizeof_esa(void) { return (int)sizeof(esa_s); }
