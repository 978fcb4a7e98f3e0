/* oracle/shim/divsufsort_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Re-entrant stand-in for libdivsufsort's divsufsort() (called from the reference at
 * src/esa.c:303, concurrently from OpenMP threads in src/dist_hack.h:8,52).
 *
 * Method (own design, not libdivsufsort's): the distinct bytes of T are ranked densely
 * (b bits per symbol, 0 = "past the end"), all suffixes are counting-sorted on their first
 * floor(24/b) symbols (<= 2^24 buckets; 8 symbols for DNA + separators), then each bucket
 * is finished with qsort_r whose comparator compares 8 bytes at a time in big-endian
 * order. Fast on genome-like text, quadratic on pathological repeats -- fine for a shim.
 * Timings of this function are always reported as "shim SA", never as libdivsufsort. */
#define _GNU_SOURCE
#include "divsufsort.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
	const unsigned char *T;
	saidx_t n;
} sufctx;

static inline uint64_t be64(const unsigned char *p) {
	uint64_t v;
	memcpy(&v, p, 8);
	return __builtin_bswap64(v);
}

static int suffix_cmp(const void *pa, const void *pb, void *vctx) {
	const sufctx *c = (const sufctx *)vctx;
	saidx_t a = *(const saidx_t *)pa, b = *(const saidx_t *)pb;
	if (a == b) return 0;
	const unsigned char *T = c->T;
	saidx_t n = c->n;
	saidx_t la = n - a, lb = n - b;
	saidx_t m = la < lb ? la : lb;
	saidx_t k = 0;
	for (; k + 8 <= m; k += 8) {
		uint64_t x = be64(T + a + k), y = be64(T + b + k);
		if (x != y) return x < y ? -1 : 1;
	}
	for (; k < m; k++) {
		if (T[a + k] != T[b + k]) return T[a + k] < T[b + k] ? -1 : 1;
	}
	/* one is a proper prefix of the other: the shorter suffix is smaller */
	return la < lb ? -1 : 1;
}

typedef struct {
	unsigned char rank[256];
	int bits, k;
} keycfg;

static inline uint32_t prefix_key(const keycfg *kc, const unsigned char *T, saidx_t n, saidx_t i) {
	uint32_t k = 0;
	for (int d = 0; d < kc->k; d++) {
		k <<= kc->bits;
		if (i + d < n) k |= kc->rank[T[i + d]];
	}
	return k;
}

int divsufsort(const unsigned char *T, saidx_t *SA, saidx_t n) {
	if (!T || !SA || n < 0) return -1;
	if (n == 0) return 0;
	keycfg kc;
	memset(&kc, 0, sizeof kc);
	{
		unsigned char seen[256] = {0};
		for (saidx_t i = 0; i < n; i++) seen[T[i]] = 1;
		int sigma = 0;
		for (int c = 0; c < 256; c++)
			if (seen[c]) kc.rank[c] = (unsigned char)(++sigma);
		kc.bits = 1;
		while ((1 << kc.bits) <= sigma) kc.bits++;
		kc.k = 24 / kc.bits;
		/* keep the bucket table no larger than ~4n entries */
		while (kc.k > 1 && ((size_t)1 << (kc.bits * kc.k)) > 4 * (size_t)n + 256) kc.k--;
	}
	const size_t NB = (size_t)1 << (kc.bits * kc.k);
	uint32_t *cnt = calloc(NB + 1, sizeof(uint32_t));
	if (!cnt) return -2;
	for (saidx_t i = 0; i < n; i++) cnt[prefix_key(&kc, T, n, i) + 1]++;
	for (size_t b = 0; b < NB; b++) cnt[b + 1] += cnt[b];
	uint32_t *cur = malloc(NB * sizeof(uint32_t));
	if (!cur) {
		free(cnt);
		return -2;
	}
	memcpy(cur, cnt, NB * sizeof(uint32_t));
	for (saidx_t i = 0; i < n; i++) SA[cur[prefix_key(&kc, T, n, i)]++] = i;
	free(cur);
	sufctx ctx = {T, n};
	for (size_t b = 0; b < NB; b++) {
		uint32_t lo = cnt[b], hi = cnt[b + 1];
		if (hi - lo > 1) qsort_r(SA + lo, hi - lo, sizeof(saidx_t), suffix_cmp, &ctx);
	}
	free(cnt);
	return 0;
}
