/* oracle/shim/gsl/gsl_rng.h -- TEST INFRASTRUCTURE ONLY.
 * GSL is not installed in this image (unpinned in the reference's configure.ac:22-27).
 * The reference touches gsl_rng only at src/andi.c:272-279,330 and src/model.c:229.
 * This shim restates the published MT19937 generator (Matsumoto & Nishimura 1998), which is
 * what gsl_rng_default resolves to. PARITY UNPINNED: no GSL binary is available to check
 * the stream against. */
#ifndef ANDI_ORACLE_SHIM_GSL_RNG_H
#define ANDI_ORACLE_SHIM_GSL_RNG_H
#include <stddef.h>
typedef struct {
	const char *name;
} gsl_rng_type;
typedef struct {
	const gsl_rng_type *type;
	unsigned long mt[624];
	int mti;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
void gsl_rng_free(gsl_rng *r);
unsigned long gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
#endif
