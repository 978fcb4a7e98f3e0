/* oracle/shim/gsl_shim.c -- TEST INFRASTRUCTURE ONLY. PARITY UNPINNED.
 * MT19937 (published reference algorithm) + a multinomial sampler built from conditional
 * binomials (the decomposition GSL documents for gsl_ran_multinomial). The binomial here is a
 * plain inversion/normal-free exact sampler (sum of Bernoulli blocks via geometric skips);
 * it is distributionally correct but does NOT reproduce GSL's BTPE stream. */
#include <gsl/gsl_randist.h>
#include <math.h>
#include <stdlib.h>

static const gsl_rng_type mt_type = {"mt19937-shim"};
const gsl_rng_type *gsl_rng_default = &mt_type;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
	gsl_rng *r = calloc(1, sizeof *r);
	if (!r) return NULL;
	r->type = T;
	gsl_rng_set(r, 0);
	return r;
}

void gsl_rng_set(gsl_rng *r, unsigned long seed) {
	if (seed == 0) seed = 4357; /* GSL's documented default seed for mt19937 */
	r->mt[0] = seed & 0xffffffffUL;
	for (int i = 1; i < 624; i++)
		r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
	r->mti = 624;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

unsigned long gsl_rng_get(gsl_rng *r) {
	unsigned long *mt = r->mt;
	if (r->mti >= 624) {
		for (int k = 0; k < 624; k++) {
			unsigned long y = (mt[k] & 0x80000000UL) | (mt[(k + 1) % 624] & 0x7fffffffUL);
			mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
		}
		r->mti = 0;
	}
	unsigned long y = mt[r->mti++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680UL;
	y ^= (y << 15) & 0xefc60000UL;
	y ^= (y >> 18);
	return y & 0xffffffffUL;
}

double gsl_rng_uniform(gsl_rng *r) { return gsl_rng_get(r) / 4294967296.0; }

unsigned int gsl_ran_binomial(gsl_rng *r, double p, unsigned int n) {
	if (p <= 0.0 || n == 0) return 0;
	if (p >= 1.0) return n;
	int flip = p > 0.5;
	double q = flip ? 1.0 - p : p;
	/* geometric-skip sampler: count successes among n trials */
	double lq = log1p(-q);
	unsigned int k = 0;
	double pos = 0.0;
	for (;;) {
		double u = gsl_rng_uniform(r);
		if (u <= 0.0) u = 1.0 / 4294967296.0;
		pos += floor(log(u) / lq) + 1.0;
		if (pos > (double)n) break;
		k++;
	}
	return flip ? n - k : k;
}

void gsl_ran_multinomial(gsl_rng *r, size_t K, unsigned int N, const double p[], unsigned int n[]) {
	double norm = 0.0, sum_p = 0.0;
	unsigned int sum_n = 0;
	for (size_t k = 0; k < K; k++) norm += p[k];
	for (size_t k = 0; k < K; k++) {
		if (p[k] > 0.0)
			n[k] = gsl_ran_binomial(r, p[k] / (norm - sum_p), N - sum_n);
		else
			n[k] = 0;
		sum_p += p[k];
		sum_n += n[k];
	}
}
