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
 * GSL is not available here, so this is MT19937 (Matsumoto & Nishimura) + the conditional
 * binomials gsl_ran_multinomial uses, each drawn EXACTLY (below); the same distribution as the
 * reference's, NOT the same random stream as GSL's (parity unpinned). */
struct host_rng {
	uint32_t mt[624];
	int at;
};

host_rng *host_rng_new(unsigned long seed) {
	host_rng *r = malloc(sizeof *r);
	if (!r) return NULL;
	r->mt[0] = (uint32_t)seed; /* every seed is taken as given, 0 included (the CLI maps "no --seed" to the time) */
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

/* Binomial(n, p), EXACT (the reference draws through gsl_ran_binomial, src/model.c:229 via
 * gsl_ran_multinomial): inversion by sequential search for small means, and for the rest the
 * published BTPE algorithm (Kachitvichyanukul & Schmeiser, "Binomial random variate generation",
 * CACM 31(2), 1988, steps 0-5.3): triangle / parallelogram / two exponential tails as the
 * majorising function, squeeze, then the exact acceptance test. Same distribution as GSL, not
 * the same random stream (parity with GSL unpinned). */
static uint32_t binomial_inversion(host_rng *r, double p, uint32_t n) {
	const double q = 1.0 - p, s = p / q, a = (n + 1.0) * s;
	const double bound = fmin((double)n, n * p + 10.0 * sqrt(n * p * q + 1.0));
	for (;;) {
		double f = pow(q, (double)n), u = rng_unit(r);
		uint32_t x = 0;
		for (;;) {
			if (u < f) return x;
			if ((double)x > bound) break; /* numerical leftovers in the far tail: draw again */
			u -= f;
			x++;
			f *= a / x - s;
		}
	}
}

static double stirling_tail(double x) {
	const double x2 = x * x;
	return (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / x2) / x2) / x2) / x2) / x / 166320.0;
}

static uint32_t binomial_btpe(host_rng *r, double p, uint32_t nn) {
	/* step 0: set-up (p <= 0.5, n*p >= 30) */
	const double n = (double)nn, q = 1.0 - p, npq = n * p * q, fm = n * p + p;
	const double m = floor(fm);
	const double p1 = floor(2.195 * sqrt(npq) - 4.6 * q) + 0.5;
	const double xm = m + 0.5, xl = xm - p1, xr = xm + p1;
	const double c = 0.134 + 20.5 / (15.3 + m);
	double a = (fm - xl) / (fm - xl * p);
	const double laml = a * (1.0 + a / 2.0);
	a = (xr - fm) / (xr * q);
	const double lamr = a * (1.0 + a / 2.0);
	const double p2 = p1 * (1.0 + 2.0 * c), p3 = p2 + c / laml, p4 = p3 + c / lamr;
	for (;;) {
		/* step 1: the triangle is accepted at once */
		double u = rng_unit(r) * p4, v = rng_unit(r), y;
		if (u <= p1) return (uint32_t)floor(xm - p1 * v + u);
		if (u <= p2) { /* step 2: parallelograms */
			const double x = xl + (u - p1) / c;
			v = v * c + 1.0 - fabs(m - x + 0.5) / p1;
			if (v > 1.0) continue;
			y = floor(x);
		} else if (u <= p3) { /* step 3: left exponential tail */
			y = floor(xl + log(v) / laml);
			if (y < 0.0) continue;
			v = v * (u - p2) * laml;
		} else { /* step 4: right exponential tail */
			y = floor(xr - log(v) / lamr);
			if (y > n) continue;
			v = v * (u - p3) * lamr;
		}
		/* step 5: acceptance. 5.1: near the mode (or npq small) evaluate f(y)/f(m) by recursion */
		const double k = fabs(y - m);
		if (k <= 20.0 || k >= npq / 2.0 - 1.0) {
			const double s = p / q, aa = s * (n + 1.0);
			double f = 1.0;
			if (m < y) {
				for (double i = m + 1.0; i <= y; i += 1.0) f *= aa / i - s;
			} else if (m > y) {
				for (double i = y + 1.0; i <= m; i += 1.0) f /= aa / i - s;
			}
			if (v > f) continue;
			return (uint32_t)y;
		}
		/* 5.2: squeezing on log f */
		const double rho = (k / npq) * ((k * (k / 3.0 + 0.625) + 0.1666666666666) / npq + 0.5);
		const double tt = -k * k / (2.0 * npq), lv = log(v);
		if (lv < tt - rho) return (uint32_t)y;
		if (lv > tt + rho) continue;
		/* 5.3: the final test with Stirling's formula */
		const double x1 = y + 1.0, f1 = m + 1.0, z = n + 1.0 - m, w = n - y + 1.0;
		const double bound = xm * log(f1 / x1) + (n - m + 0.5) * log(z / w) + (y - m) * log(w * p / (x1 * q)) + stirling_tail(f1) +
							 stirling_tail(z) + stirling_tail(x1) + stirling_tail(w);
		if (lv > bound) continue;
		return (uint32_t)y;
	}
}

uint32_t host_rng_binomial(host_rng *r, double p, uint32_t n) {
	if (!(p > 0.0) || n == 0) return 0;
	if (p >= 1.0) return n;
	const int flip = p > 0.5;
	const double q = flip ? 1.0 - p : p;
	const uint32_t k = (double)n * q < 30.0 ? binomial_inversion(r, q, n) : binomial_btpe(r, q, n);
	return flip ? n - k : k;
}

andi_model host_model_bootstrap(host_rng *r, andi_model datum) {
	size_t nucl = total(&datum);
	double p[16], norm = 0.0, used_p = 0.0;
	for (int k = 0; k < 16; k++) p[k] = datum.counts[k] / (double)nucl, norm += p[k];
	uint32_t left = (uint32_t)nucl;
	for (int k = 0; k < 16; k++) {
		uint32_t draw = 0;
		if (p[k] > 0.0 && left) {
			double cond = p[k] / (norm - used_p);
			draw = host_rng_binomial(r, cond > 1.0 ? 1.0 : cond, left);
		}
		datum.counts[k] = draw;
		used_p += p[k];
		left -= draw;
	}
	return datum;
}
