/* oracle/shim/gsl/gsl_randist.h -- TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see gsl_rng.h).
 * Only gsl_ran_multinomial is used by the reference (src/model.c:229). */
#ifndef ANDI_ORACLE_SHIM_GSL_RANDIST_H
#define ANDI_ORACLE_SHIM_GSL_RANDIST_H
#include <gsl/gsl_rng.h>
void gsl_ran_multinomial(gsl_rng *r, size_t K, unsigned int N, const double p[], unsigned int n[]);
unsigned int gsl_ran_binomial(gsl_rng *r, double p, unsigned int n);
#endif
