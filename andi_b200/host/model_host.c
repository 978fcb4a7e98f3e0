/* andi_b200/host/model_host.c -- estimators and bootstrap on the 68-byte cells the GPU path
 * returns. Host FP64 like the reference (src/model.c:39-232): 16 integers
 * in, one double out -- nothing for a GPU to do. */
#include "andi_host.h"
#include <math.h>
#include <stdlib.h>

andi_model model_average(const andi_model *a, const andi_model *b) {
	/* src/model.c:39-46: the "average" is the cell-wise sum */
	andi_model r = *a;
	for (int k = 0; k < 16; k++) r.counts[k] += b->counts[k];
	r.seq_len += b->seq_len;
	return r;
}

static size_t cell_sum(const andi_model *m, const int *cells, int n) {
	size_t s = 0;
	for (int k = 0; k < n; k++) s += m->counts[cells[k]];
	return s;
}

static size_t total(const andi_model *m) {
	size_t s = 0;
	for (int k = 0; k < 16; k++) s += m->counts[k];
	return s;
}

double model_coverage(const andi_model *m) { return (double)total(m) / (double)m->seq_len; } /* model.c:68-73 */

static double raw(const andi_model *m) {
	/* src/model.c:81-93 */
	static const int off_diag[12] = {1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14};
	size_t nucl = total(m), snps = cell_sum(m, off_diag, 12);
	if (nucl <= 3) return NAN;
	return (double)snps / (double)nucl;
}

static double kimura(const andi_model *m) {
	/* src/model.c:115-130 */
	static const int ts[4] = {2, 8, 7, 13};
	static const int tv[8] = {1, 4, 3, 12, 9, 6, 11, 14};
	size_t nucl = total(m);
	double P = (double)cell_sum(m, ts, 4) / (double)nucl;
	double Q = (double)cell_sum(m, tv, 8) / (double)nucl;
	double tmp = 1.0 - 2.0 * P - Q;
	double d = -0.25 * log((1.0 - 2.0 * Q) * tmp * tmp);
	return d <= 0.0 ? 0.0 : d;
}

static double logdet(const andi_model *m) {
	/* src/model.c:161-198 */
	double nucl = (double)total(m), P[16];
	for (int k = 0; k < 16; k++) P[k] = m->counts[k] / nucl;
	double margins = 0.0;
	for (int r = 0; r < 4; r++) {
		const int row[4] = {4 * r, 4 * r + 1, 4 * r + 2, 4 * r + 3};
		margins += log(cell_sum(m, row, 4) / nucl);
	}
	for (int c = 0; c < 4; c++) {
		const int col[4] = {c, 4 + c, 8 + c, 12 + c};
		margins += log(cell_sum(m, col, 4) / nucl);
	}
	/* determinant of the 4x4 frequency matrix: first row times 3x3 cofactors, each cofactor
	 * written with the 2x2 minors of the last two rows */
#define MINOR(a, b) (P[8 + (a)] * P[12 + (b)] - P[12 + (a)] * P[8 + (b)])
	double det = P[0] * P[5] * MINOR(2, 3) - P[0] * P[6] * MINOR(1, 3) + P[0] * P[7] * MINOR(1, 2) -
				 P[1] * P[4] * MINOR(2, 3) + P[1] * P[6] * MINOR(0, 3) - P[1] * P[7] * MINOR(0, 2) +
				 P[2] * P[4] * MINOR(1, 3) - P[2] * P[5] * MINOR(0, 3) + P[2] * P[7] * MINOR(0, 1) -
				 P[3] * P[4] * MINOR(1, 2) + P[3] * P[5] * MINOR(0, 2) - P[3] * P[6] * MINOR(0, 1);
#undef MINOR
	double d = -0.25 * (log(det) - 0.5 * margins);
	return d <= 0.0 ? 0.0 : d;
}

double model_estimate(const andi_model *m, int model_id) {
	switch (model_id) {
		case ANDI_M_RAW: return raw(m);
		case ANDI_M_KIMURA: return kimura(m);
		case ANDI_M_LOGDET: return logdet(m);
		case ANDI_M_ANI: return (1.0 - raw(m)) * 100; /* src/model.c:206-209 */
		case ANDI_M_JC:
		default: {
			/* src/model.c:101-107 */
			double d = -0.75 * log(1.0 - (4.0 / 3.0) * raw(m));
			return d <= 0.0 ? 0.0 : d;
		}
	}
}

/* ---- bootstrap: src/model.c:222-232 draws the 16 counts from a multinomial with GSL.
 * GSL is not available here, so this is MT19937 (Matsumoto & Nishimura) + conditional
 * binomials; distributionally equivalent, NOT stream-identical to GSL (parity unpinned). */
struct host_rng {
	uint32_t mt[624];
	int at;
};

host_rng *host_rng_new(unsigned long seed) {
	host_rng *r = malloc(sizeof *r);
	if (!r) return NULL;
	if (seed == 0) seed = 4357;
	r->mt[0] = (uint32_t)seed;
	for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
	r->at = 624;
	return r;
}

void host_rng_free(host_rng *r) { free(r); }

static uint32_t rng_u32(host_rng *r) {
	if (r->at >= 624) {
		for (int k = 0; k < 624; k++) {
			uint32_t y = (r->mt[k] & 0x80000000u) | (r->mt[(k + 1) % 624] & 0x7fffffffu);
			r->mt[k] = r->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
		}
		r->at = 0;
	}
	uint32_t y = r->mt[r->at++];
	y ^= y >> 11;
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= y >> 18;
	return y;
}

static double rng_unit(host_rng *r) { return (rng_u32(r) + 0.5) / 4294967296.0; }

static double rng_normal(host_rng *r) {
	double u = rng_unit(r), v = rng_unit(r);
	return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
}

/* Binomial(n, p): exact geometric-skip sampling for small n*p, otherwise a continuity-
 * corrected normal draw clamped to [0, n] (n*p*(1-p) is in the thousands for genome counts). */
static uint32_t rng_binomial(host_rng *r, double p, uint32_t n) {
	if (p <= 0.0 || n == 0) return 0;
	if (p >= 1.0) return n;
	int flip = p > 0.5;
	double q = flip ? 1.0 - p : p;
	uint32_t k;
	if ((double)n * q < 64.0) {
		double lq = log1p(-q), pos = 0.0;
		k = 0;
		for (;;) {
			pos += floor(log(rng_unit(r)) / lq) + 1.0;
			if (pos > (double)n) break;
			k++;
		}
	} else {
		double x = floor((double)n * q + sqrt((double)n * q * (1.0 - q)) * rng_normal(r) + 0.5);
		if (x < 0) x = 0;
		if (x > (double)n) x = (double)n;
		k = (uint32_t)x;
	}
	return flip ? n - k : k;
}

andi_model model_bootstrap(host_rng *r, andi_model datum) {
	size_t nucl = total(&datum);
	double p[16], norm = 0.0, used_p = 0.0;
	for (int k = 0; k < 16; k++) p[k] = datum.counts[k] / (double)nucl, norm += p[k];
	uint32_t left = (uint32_t)nucl;
	for (int k = 0; k < 16; k++) {
		uint32_t draw = 0;
		if (p[k] > 0.0 && left) {
			double cond = p[k] / (norm - used_p);
			draw = rng_binomial(r, cond > 1.0 ? 1.0 : cond, left);
		}
		datum.counts[k] = draw;
		used_p += p[k];
		left -= draw;
	}
	return datum;
}
