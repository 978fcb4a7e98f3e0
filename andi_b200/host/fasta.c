/* andi_b200/host/fasta.c -- FASTA ingest for the command line.
 * Stands where the reference uses its vendored pfasta parser (libs/pfasta.c) through
 * read_fasta / read_fasta_join (src/io.c:159-233), seq_init + normalize
 * (src/sequence.c:234-282) and dsa_join (src/sequence.c:78-125). Own implementation: the whole
 * file is slurped and scanned once. Accepted grammar follows pfasta: a record is '>' name
 * [comment] newline, then whitespace-separated words that start with a letter, '-' or '*'. */
#define _GNU_SOURCE
#include "andi_host.h"
#include <ctype.h>
#include <err.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void seqs_init(host_seqs *v) {
	v->data = NULL;
	v->size = v->capacity = 0;
}

void seqs_push(host_seqs *v, host_seq s) {
	if (v->size == v->capacity) {
		size_t cap = v->capacity ? v->capacity * 2 : 8;
		host_seq *p = realloc(v->data, cap * sizeof *p);
		if (!p) err(errno, "Out of memory");
		v->data = p;
		v->capacity = cap;
	}
	v->data[v->size++] = s;
}

void seqs_free(host_seqs *v) {
	for (size_t i = 0; i < v->size; i++) {
		free(v->data[i].S);
		free(v->data[i].name);
	}
	free(v->data);
	seqs_init(v);
}

static char *slurp(const char *file_name, size_t *len) {
	FILE *f = strcmp(file_name, "-") ? fopen(file_name, "rb") : stdin;
	if (!f) return NULL;
	size_t cap = 1 << 20, n = 0;
	char *buf = malloc(cap);
	if (!buf) err(errno, "Out of memory");
	for (;;) {
		if (n == cap) {
			cap *= 2;
			char *p = realloc(buf, cap);
			if (!p) err(errno, "Out of memory");
			buf = p;
		}
		size_t got = fread(buf + n, 1, cap - n, f);
		n += got;
		if (got == 0) break;
	}
	int bad = ferror(f);
	if (f != stdin) fclose(f);
	if (bad) {
		free(buf);
		errno = EIO;
		return NULL;
	}
	*len = n;
	return buf;
}

/* src/sequence.c:260-282: keep ACGT and '!', upper-case acgt, drop everything else */
static size_t normalize_into(char *dst, const char *src, size_t n, int *non_acgt) {
	size_t w = 0;
	for (size_t i = 0; i < n; i++) {
		char c = src[i];
		switch (c) {
			case 'A': case 'C': case 'G': case 'T': case '!': dst[w++] = c; break;
			case 'a': case 'c': case 'g': case 't': dst[w++] = (char)(c - 32); break;
			default:
				if (!isspace((unsigned char)c)) *non_acgt = 1;
				break;
		}
	}
	dst[w] = '\0';
	return w;
}

/* The reader proper. It prints nothing: a failure leaves its message in msg (what the reference
 * would have warned), so that files can be read by several threads and reported in order. */
int fasta_read_quiet(const char *file_name, host_seqs *out, int *flags, char *msg, size_t msg_len) {
	size_t len = 0;
	if (msg_len) msg[0] = '\0';
	char *buf = slurp(file_name, &len);
	if (!buf) {
		*flags |= HF_SOFT_ERROR;
		snprintf(msg, msg_len, "%s: %s", file_name, strerror(errno));
		return 1;
	}
	const char *fail = NULL;
	size_t line = 1, p = 0, fail_line = 0;
	char failbuf[128];
	if (len == 0) {
		fail = "File is empty.";
	} else if (buf[0] != '>') {
		fail = "File must start with '>'.";
	}
	while (!fail && p < len) {
		if (buf[p] != '>') {
			snprintf(failbuf, sizeof failbuf, "Expected '>' but found '%c' on line %zu.", buf[p], line);
			fail = failbuf;
			break;
		}
		p++;
		size_t name_begin = p;
		while (p < len && !isspace((unsigned char)buf[p])) p++;
		if (p == name_begin) {
			snprintf(failbuf, sizeof failbuf, "Empty name on line %zu.", line);
			fail = failbuf;
			break;
		}
		size_t name_end = p;
		while (p < len && buf[p] != '\n') p++; /* comment */
		if (p >= len) {
			snprintf(failbuf, sizeof failbuf, "Unexpected EOF in %s on line %zu.", name_end == len ? "name" : "comment", line);
			fail = failbuf;
			break;
		}
		/* sequence: words that start with a letter, '-' or '*' */
		size_t seq_begin = p;
		for (;;) {
			while (p < len && isspace((unsigned char)buf[p])) {
				if (buf[p] == '\n') line++;
				p++;
			}
			if (p >= len) break;
			char c = buf[p];
			if (!(isalpha((unsigned char)c) || c == '-' || c == '*')) break;
			while (p < len && !isspace((unsigned char)buf[p])) p++;
		}
		size_t seq_end = p;
		host_seq s;
		s.S = malloc(seq_end - seq_begin + 1);
		s.name = strndup(buf + name_begin, name_end - name_begin);
		if (!s.S || !s.name) err(errno, "Out of memory");
		int dropped = 0;
		s.len = normalize_into(s.S, buf + seq_begin, seq_end - seq_begin, &dropped);
		/* did the record hold any sequence characters at all? (pfasta: "Empty sequence") */
		int any = 0;
		for (size_t i = seq_begin; i < seq_end && !any; i++) any = !isspace((unsigned char)buf[i]);
		if (!any) {
			free(s.S), free(s.name);
			snprintf(failbuf, sizeof failbuf, "Empty sequence on line %zu.", line);
			fail = failbuf;
			break;
		}
		if (dropped) *flags |= HF_NON_ACGT;
		seqs_push(out, s);
		(void)fail_line;
	}
	free(buf);
	if (fail) {
		*flags |= HF_SOFT_ERROR;
		snprintf(msg, msg_len, "%s: %s", file_name, fail);
		return 1;
	}
	return 0;
}

int fasta_read(const char *file_name, host_seqs *out, int *flags) {
	char msg[512];
	int rc = fasta_read_quiet(file_name, out, flags, msg, sizeof msg);
	if (msg[0]) warnx("%s", msg);
	return rc;
}

int fasta_read_join(const char *file_name, host_seqs *out, int *flags) {
	char msg[512];
	int rc = fasta_read_join_quiet(file_name, out, flags, msg, sizeof msg);
	if (msg[0]) warnx("%s", msg);
	return rc;
}

int fasta_read_join_quiet(const char *file_name, host_seqs *out, int *flags, char *msg, size_t msg_len) {
	host_seqs single;
	seqs_init(&single);
	fasta_read_quiet(file_name, &single, flags, msg, msg_len);
	if (single.size == 0) {
		seqs_free(&single);
		return 1;
	}
	size_t total = 0;
	for (size_t i = 0; i < single.size; i++) total += single.data[i].len + 1;
	host_seq joined;
	joined.S = malloc(total);
	if (!joined.S) err(errno, "Out of memory");
	char *w = joined.S;
	for (size_t i = 0; i < single.size; i++) {
		if (i) *w++ = '!';
		memcpy(w, single.data[i].S, single.data[i].len);
		w += single.data[i].len;
	}
	*w = '\0';
	joined.len = total - 1;
	/* name = file name without directory and without everything after the first dot */
	const char *base = strrchr(file_name, '/');
	base = base ? base + 1 : file_name;
	const char *dot = strchrnul(base, '.');
	joined.name = strndup(base, (size_t)(dot - base));
	if (!joined.name) err(errno, "Out of memory");
	seqs_push(out, joined);
	seqs_free(&single);
	return 0;
}
