/* andi_b200/host/andi_main.c -- the andi command line on top of libandi_b200.so.
 *
 * Keeps the reference's options with the reference's meanings (src/andi.c:63-333):
 *   -j/--join (a flag, NOT a thread count), -t/--threads, -m/--model RAW|JC|KIMURA|LOGDET|ANI,
 *   -b/--bootstrap, -p, -l/--low-memory, -v, --file-of-filenames, --truncate-names,
 *   --progress, -h, --version.
 * calculate_distances (src/process.c:230-270) becomes: pack the pool on the GPU, one call to
 * andi_dist_matrix_multi (subjects spread over the GPUs named by --devices, as the reference spreads
 * them over -t threads, src/dist_hack.h:8), then the reference's host-side printing. Extra, not in
 * the reference: --seed N (bootstrap RNG; the reference seeds with time(NULL)), --device N,
 * --devices LIST. */
#define _GNU_SOURCE
#include "andi_host.h"
#include <err.h>
#include <errno.h>
#include <getopt.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>
#include <unistd.h>

static void usage(int status) {
	static const char text[] =
		"Usage: andi [OPTIONS...] FILES...\n"
		"\tFILES... can be any sequence of FASTA files.\n"
		"\tUse '-' as file name to read from stdin.\n"
		"Options:\n"
		"  -b, --bootstrap=INT  Print additional bootstrap matrices\n"
		"      --file-of-filenames=FILE  Read additional filenames from FILE; one per line\n"
		"  -j, --join           Treat all sequences from one file as a single genome\n"
		"  -l, --low-memory     Use less memory at the cost of speed\n"
		"  -m, --model=MODEL    Pick an evolutionary model of 'Raw', 'JC', 'Kimura', 'LogDet', 'ANI'; default: JC\n"
		"  -p FLOAT             Significance of an anchor; default: 0.025\n"
		"      --progress=WHEN  Print a progress bar 'always', 'never', or 'auto'; default: auto\n"
		"  -t, --threads=INT    Accepted for compatibility; the work runs on the GPU\n"
		"      --truncate-names Truncate names to ten characters\n"
		"  -v, --verbose        Prints additional information\n"
		"  -h, --help           Display this help and exit\n"
		"      --version        Output version information and acknowledgments\n"
		"      --seed=INT       Seed of the bootstrap generator (default: time)\n"
		"      --device=INT     CUDA device to use (default 0)\n"
		"      --devices=LIST   CUDA devices to spread the subjects over, e.g. 0-7 or 0,2,3 or all\n";
	fputs(text, status == EXIT_SUCCESS ? stdout : stderr);
	exit(status);
}

typedef struct {
	char **data;
	size_t size, capacity;
} names;

/* "0-7", "0,2,3", "1", "all" -> device numbers; returns how many (0 = malformed) */
static int parse_devices(const char *arg, int *out, int cap) {
	int n = 0;
	if (!strcasecmp(arg, "all")) {
		int count = andi_device_count();
		for (int d = 0; d < count && n < cap; d++) out[n++] = d;
		return n;
	}
	const char *p = arg;
	while (*p) {
		char *end;
		long a = strtol(p, &end, 10), b;
		if (end == p || a < 0) return 0;
		b = a;
		if (*end == '-') {
			p = end + 1;
			b = strtol(p, &end, 10);
			if (end == p || b < a) return 0;
		}
		for (long d = a; d <= b && n < cap; d++) out[n++] = (int)d;
		if (*end == ',') end++;
		else if (*end) return 0;
		p = end;
	}
	return n;
}

/* src/dist_hack.h:37-43,74-95: the progress line on stderr */
typedef struct {
	size_t n;
} progress_state;

static void print_progress(size_t done, size_t total, void *user) {
	const progress_state *st = user;
	fprintf(stderr, "\rComparing %zu sequences: %5.1f%% (%zu/%zu)", st->n, total ? 100.0 * (double)done / (double)total : 100.0, done,
			total);
}

static void names_push(names *v, char *s) {
	if (v->size == v->capacity) {
		v->capacity = v->capacity ? 2 * v->capacity : 8;
		v->data = realloc(v->data, v->capacity * sizeof *v->data);
		if (!v->data) err(errno, "Out of memory");
	}
	v->data[v->size++] = s;
}

/* src/io.c:103-144: one file name per line, empty lines ignored, "-" = stdin */
static void read_file_of_filenames(const char *fof, names *v, int *flags) {
	FILE *f = strcmp(fof, "-") ? fopen(fof, "r") : stdin;
	if (!f) {
		*flags |= HF_SOFT_ERROR;
		warn("%s", fof);
		return;
	}
	char *line = NULL;
	size_t cap = 0;
	ssize_t got;
	while ((got = getline(&line, &cap, f)) != -1) {
		char *nl = strchr(line, '\n');
		if (nl) *nl = '\0';
		if (*line) names_push(v, strdup(line));
	}
	free(line);
	if (f != stdin) fclose(f);
}

int main(int argc, char *argv[]) {
	static const struct option long_options[] = {{"version", no_argument, NULL, 0},
												 {"truncate-names", no_argument, NULL, 0},
												 {"file-of-filenames", required_argument, NULL, 0},
												 {"progress", optional_argument, NULL, 0},
												 {"seed", required_argument, NULL, 0},
												 {"device", required_argument, NULL, 0},
												 {"devices", required_argument, NULL, 0},
												 {"help", no_argument, NULL, 'h'},
												 {"verbose", no_argument, NULL, 'v'},
												 {"join", no_argument, NULL, 'j'},
												 {"low-memory", no_argument, NULL, 'l'},
												 {"threads", required_argument, NULL, 't'},
												 {"bootstrap", required_argument, NULL, 'b'},
												 {"model", required_argument, NULL, 'm'},
												 {0, 0, 0, 0}};
	host_config cfg = {.flags = 0, .model = ANDI_M_JC, .p_value = 0.025, .bootstrap = 0, .seed = 0, .device = 0};
	names files = {0};
	enum { P_AUTO, P_NEVER, P_ALWAYS } progress = P_AUTO; /* src/andi.c:83 */
	int devices[64], n_devices = 0, seed_given = 0;

	for (;;) {
		int idx = 0;
		int c = getopt_long(argc, argv, "jvht:p:m:b:l", long_options, &idx);
		if (c == -1) break;
		switch (c) {
			case 0: {
				const char *name = long_options[idx].name;
				if (!strcmp(name, "version")) {
					printf("andi (andi_b200, GPU hot path) compatible with andi 1.15\n");
					return EXIT_SUCCESS;
				} else if (!strcmp(name, "truncate-names")) {
					cfg.flags |= HF_TRUNCATE_NAMES;
				} else if (!strcmp(name, "file-of-filenames")) {
					read_file_of_filenames(optarg, &files, &cfg.flags);
				} else if (!strcmp(name, "progress")) {
					/* src/andi.c:111-122 */
					if (!optarg || !strcasecmp(optarg, "always")) progress = P_ALWAYS;
					else if (!strcasecmp(optarg, "auto")) progress = P_AUTO;
					else if (!strcasecmp(optarg, "never")) progress = P_NEVER;
					else
						warnx("invalid argument to --progress '%s'. Expected one of 'auto', 'always', or 'never'.", optarg);
				} else if (!strcmp(name, "seed")) {
					cfg.seed = strtoul(optarg, NULL, 10);
					seed_given = 1;
				} else if (!strcmp(name, "device")) {
					cfg.device = atoi(optarg);
				} else if (!strcmp(name, "devices")) {
					n_devices = parse_devices(optarg, devices, 64);
					if (!n_devices) errx(1, "Expected a list of CUDA devices for --devices (e.g. 0-7 or 0,2), but '%s' was given.", optarg);
				}
				break;
			}
			case 'h': usage(EXIT_SUCCESS); break;
			case 'v': cfg.flags |= (cfg.flags & HF_VERBOSE) ? HF_EXTRA_VERBOSE : HF_VERBOSE; break;
			case 'p': {
				errno = 0;
				char *end;
				double prop = strtod(optarg, &end);
				if (errno || end == optarg || *end != '\0') {
					cfg.flags |= HF_SOFT_ERROR;
					warnx("Expected a floating point number for -p argument, but '%s' was given. Skipping argument.", optarg);
				} else if (prop <= 0.0 || prop >= 1.0) {
					cfg.flags |= HF_SOFT_ERROR;
					warnx("A probability should be a value between 0 and 1, exclusive; Ignoring -p %f argument.", prop);
				} else {
					cfg.p_value = prop;
				}
				break;
			}
			case 'l': cfg.flags |= HF_LOW_MEMORY; break;
			case 'j': cfg.flags |= HF_JOIN; break;
			case 't': {
				errno = 0;
				char *end;
				(void)strtoul(optarg, &end, 10);
				if (errno || end == optarg || *end != '\0')
					warnx("Expected a number for -t argument, but '%s' was given. Ignoring -t argument.", optarg);
				break;
			}
			case 'b': {
				errno = 0;
				char *end;
				unsigned long b = strtoul(optarg, &end, 10);
				if (errno || end == optarg || *end != '\0' || b == 0) {
					cfg.flags |= HF_SOFT_ERROR;
					warnx("Expected a positive number for -b argument, but '%s' was given. Ignoring -b argument.", optarg);
				} else {
					cfg.bootstrap = b - 1;
				}
				break;
			}
			case 'm': {
				if (!strcasecmp(optarg, "RAW")) cfg.model = ANDI_M_RAW;
				else if (!strcasecmp(optarg, "JC")) cfg.model = ANDI_M_JC;
				else if (!strcasecmp(optarg, "KIMURA")) cfg.model = ANDI_M_KIMURA;
				else if (!strcasecmp(optarg, "LOGDET")) cfg.model = ANDI_M_LOGDET;
				else if (!strcasecmp(optarg, "ANI")) cfg.model = ANDI_M_ANI;
				else {
					cfg.flags |= HF_SOFT_ERROR;
					warnx("Ignoring argument for --model. Expected Raw, JC, Kimura, LogDet or ANI");
				}
				break;
			}
			default: usage(EXIT_FAILURE);
		}
	}
	for (int i = optind; i < argc; i++) names_push(&files, strdup(argv[i]));

	if ((cfg.flags & HF_JOIN) && files.size == 0) errx(1, "In join mode at least one filename needs to be supplied.");
	size_t minfiles = (cfg.flags & HF_JOIN) ? 2 : 1;
	if (files.size < minfiles) {
		if (!isatty(STDIN_FILENO))
			names_push(&files, strdup("-"));
		else
			usage(EXIT_FAILURE);
	}

	/* every file is read by one thread (src/io.c:159-233 reads them one after the other); sequences,
	 * flags and warnings are then collected in file order, so nothing depends on the thread count */
	host_seqs seqs;
	seqs_init(&seqs);
	{
		host_seqs *part = calloc(files.size ? files.size : 1, sizeof *part);
		int *part_flags = calloc(files.size ? files.size : 1, sizeof *part_flags);
		char(*msg)[512] = calloc(files.size ? files.size : 1, sizeof *msg);
		if (!part || !part_flags || !msg) err(errno, "Out of memory");
		const int join = (cfg.flags & HF_JOIN) != 0;
#pragma omp parallel for schedule(dynamic, 1)
		for (size_t i = 0; i < files.size; i++) {
			seqs_init(&part[i]);
			if (join)
				fasta_read_join_quiet(files.data[i], &part[i], &part_flags[i], msg[i], sizeof msg[i]);
			else
				fasta_read_quiet(files.data[i], &part[i], &part_flags[i], msg[i], sizeof msg[i]);
		}
		for (size_t i = 0; i < files.size; i++) {
			if (msg[i][0]) warnx("%s", msg[i]);
			cfg.flags |= part_flags[i];
			for (size_t k = 0; k < part[i].size; k++) seqs_push(&seqs, part[i].data[k]);
			free(part[i].data); /* the sequences themselves moved into seqs */
			free(files.data[i]);
		}
		free(part), free(part_flags), free(msg);
	}
	free(files.data);

	const size_t n = seqs.size;
	if (n < 2) errx(1, "I am truly sorry, but with less than two sequences (%zu given) there is nothing to compare.", n);
	if (cfg.flags & HF_NON_ACGT)
		warnx("The input sequences contained characters other than acgtACGT. These were automatically stripped to "
			  "ensure correct results.");
	for (size_t i = 0; i < n; i++) {
		const host_seq *s = &seqs.data[i];
		if ((cfg.flags & HF_TRUNCATE_NAMES) && strlen(s->name) > 10)
			warnx("The sequence name '%s' is longer than ten characters. It will be truncated in the output to '%.10s'.",
				  s->name, s->name);
		const size_t limit = (INT_MAX - 1) / 2;
		if (s->len > limit) errx(1, "The sequence %s is too long. The technical limit is %zu.", s->name, limit);
		if (s->len == 0) errx(1, "The sequence %s is empty.", s->name);
		if (s->len < 1000) cfg.flags |= HF_SHORT;
	}
	if (cfg.flags & HF_SHORT) {
		cfg.flags |= HF_SOFT_ERROR;
		warnx("One of the given input sequences is shorter than a thousand nucleotides. This may result in "
			  "inaccurate distances. Try an alignment instead.");
	}

	/* src/andi.c:318-324 */
	if (progress == P_AUTO) progress = isatty(STDERR_FILENO) ? P_ALWAYS : P_NEVER;
	if (progress == P_ALWAYS) cfg.flags |= HF_PRINT_PROGRESS;

	/* ---- calculate_distances (src/process.c:230-270) on the GPU(s) */
	if (!n_devices) devices[0] = cfg.device, n_devices = 1;
	const char **ptr = malloc(n * sizeof *ptr);
	size_t *len = malloc(n * sizeof *len);
	andi_model *M = malloc(n * n * sizeof *M);
	if (!ptr || !len || !M)
		err(errno, "Could not allocate enough memory for the comparison matrix. Try using --join or --low-memory.");
	for (size_t i = 0; i < n; i++) ptr[i] = seqs.data[i].S, len[i] = seqs.data[i].len;
	char msg[512];
	progress_state pst = {n};
	const int show = (cfg.flags & HF_PRINT_PROGRESS) != 0;
	if (show) print_progress(0, n * n - n, &pst);
	if (andi_dist_matrix_multi(devices, n_devices, ptr, len, n, cfg.p_value, cfg.model, (cfg.flags & HF_LOW_MEMORY) != 0, M,
							   show ? print_progress : NULL, &pst, msg, sizeof msg))
		errx(1, "Failed to create index: %s", msg);
	if (show) fprintf(stderr, ", done.\n");

	host_print_distances(stdout, M, &seqs, &cfg, 1, &cfg.flags);
	if (cfg.flags & HF_VERBOSE) host_print_coverages(stdout, M, n);

	if (cfg.bootstrap) {
		/* src/process.c:289-321 */
		host_rng *rng = host_rng_new(seed_given ? cfg.seed : (unsigned long)time(NULL));
		andi_model *B = malloc(n * n * sizeof *B);
		if (!rng || !B) err(errno, "Out of memory");
		while (cfg.bootstrap--) {
			for (size_t i = 0; i < n; i++) {
				for (size_t j = i; j < n; j++) {
					if (i == j) {
						memset(&B[i * n + j], 0, sizeof *B);
						B[i * n + j].seq_len = 1, B[i * n + j].counts[0] = 1;
						continue;
					}
					andi_model datum = model_average(&M[i * n + j], &M[j * n + i]);
					datum = host_model_bootstrap(rng, datum);
					B[i * n + j] = B[j * n + i] = datum;
				}
			}
			host_print_distances(stdout, B, &seqs, &cfg, 0, &cfg.flags);
		}
		free(B);
		host_rng_free(rng);
	}

	free(M), free(ptr), free(len);
	seqs_free(&seqs);
	return (cfg.flags & HF_SOFT_ERROR) ? EXIT_FAILURE : EXIT_SUCCESS;
}
