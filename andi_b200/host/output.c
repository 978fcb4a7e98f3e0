/* andi_b200/host/output.c -- PHYLIP distance matrix and coverage output, byte-compatible with
 * the reference (src/io.c:246-338). */
#include "andi_host.h"
#include <err.h>
#include <errno.h>
#include <math.h>
#include <stdlib.h>

void host_print_distances(FILE *out, const andi_model *M, const host_seqs *seqs, const host_config *cfg, int warnings,
					 int *flags) {
	const size_t n = seqs->size;
	double *D = malloc(n * n * sizeof *D);
	if (!D) err(errno, "Out of memory");
	int scientific = 0;
	for (size_t i = 0; i < n; i++) {
		for (size_t j = 0; j < n; j++) {
			const andi_model *ij = &M[i * n + j], *ji = &M[j * n + i];
			andi_model datum = *ij;
			if (!(cfg->flags & HF_EXTRA_VERBOSE)) datum = model_average(ij, ji); /* io.c:272-276 */
			double d = D[i * n + j] = (i == j) ? 0.0 : model_estimate(&datum, cfg->model);
			if (d > 0 && d < 0.001) scientific = 1; /* io.c:280-282 */
			if (isnan(d) && warnings) {
				*flags |= HF_SOFT_ERROR;
				warnx("For the two sequences '%s' and '%s' the distance computation failed and is reported as "
					  "nan. Please refer to the documentation for further details.",
					  seqs->data[i].name, seqs->data[j].name);
			}
			if (!isnan(d) && i < j && warnings) {
				double c1 = model_coverage(ij), c2 = model_coverage(ji);
				if (c1 < 0.2 || c2 < 0.2) { /* io.c:292-303 */
					*flags |= HF_SOFT_ERROR;
					warnx("For the two sequences '%s' and '%s' very little homology was found (%f and %f, "
						  "respectively).",
						  seqs->data[i].name, seqs->data[j].name, c1, c2);
				}
			}
		}
	}
	fprintf(out, "%zu\n", n);
	for (size_t i = 0; i < n; i++) {
		fprintf(out, (cfg->flags & HF_TRUNCATE_NAMES) ? "%-10.10s" : "%-10s", seqs->data[i].name);
		for (size_t j = 0; j < n; j++) fprintf(out, scientific ? " %1.4e" : " %1.4f", D[i * n + j]);
		fprintf(out, "\n");
	}
	free(D);
}

void host_print_coverages(FILE *out, const andi_model *M, size_t n) {
	/* src/io.c:329-338 */
	fprintf(out, "\nCoverage:\n");
	for (size_t i = 0; i < n; i++) {
		for (size_t j = 0; j < n; j++) fprintf(out, "%1.4e ", model_coverage(&M[i * n + j]));
		fprintf(out, "\n");
	}
}
