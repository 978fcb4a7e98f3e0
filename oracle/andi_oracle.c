/* oracle/andi_oracle.c -- TEST INFRASTRUCTURE ONLY (see andi_oracle.h).
 *
 * CPU restatement of the reference's hot path. Written from the reference's behaviour, not
 * from its text: the suffix sorter is a plain prefix-doubling sorter, the child table is
 * computed from its nearest-smaller-value characterisation, the prefix cache is evaluated
 * per 10-mer instead of by recursion, and the search has two forms (the spec by binary
 * search, and the reference's child-table procedure). Each function names the reference
 * lines it must agree with; tests/test_oracle_vs_ref.py checks that agreement against
 * oracle/_ref/libandi_ref.so.
 */
#define _GNU_SOURCE
#include "andi_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* sequence preparation (src/sequence.c)                                                 */

size_t orc_normalize(char *s, int *non_acgt) {
	/* src/sequence.c:260-282 */
	size_t w = 0;
	int dropped = 0;
	for (size_t r = 0; s[r]; r++) {
		char c = s[r];
		if (c == 'a' || c == 'c' || c == 'g' || c == 't') c = (char)(c - 'a' + 'A');
		if (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == '!')
			s[w++] = c;
		else
			dropped = 1;
	}
	s[w] = '\0';
	if (non_acgt && dropped) *non_acgt = 1;
	return w;
}

static char complement_char(char c) {
	/* src/sequence.c:154-163 : anything below 'A' (only '!' survives normalize) becomes ';' */
	switch (c) {
		case 'A': return 'T';
		case 'C': return 'G';
		case 'G': return 'C';
		case 'T': return 'A';
		default: return c < 'A' ? ';' : c;
	}
}

void orc_make_rs(const char *s, size_t n, char *out) {
	/* src/sequence.c:143-189 ; pinned by test/test_seq.c:34,53,69 */
	for (size_t k = 0; k < n; k++) out[k] = complement_char(s[n - 1 - k]);
	out[n] = '#';
	memcpy(out + n + 1, s, n);
	out[2 * n + 1] = '\0';
}

double orc_gc(const char *s, size_t n) {
	/* src/sequence.c:196-207 */
	size_t gc = 0;
	for (size_t k = 0; k < n; k++) gc += (s[k] == 'G' || s[k] == 'C');
	return (double)gc / (double)n;
}

static size_t choose(size_t n, size_t k) {
	/* src/sequence.c:314-335 (integer arithmetic, same evaluation order) */
	if (n == 0 || k > n) return 0;
	if (k == 0 || k == n) return 1;
	if (k > n - k) k = n - k;
	size_t r = 1;
	for (size_t i = 1; i <= k; i++) {
		r *= n - k + i;
		r /= i;
	}
	return r;
}

double orc_shustring_cum_prob(size_t x, double p, size_t l) {
	/* src/sequence.c:352-373 ; the floating point expression is kept term for term so the
	 * threshold comes out identical. */
	double xx = (double)x, ll = (double)l, s = 0.0;
	for (size_t k = 0; k <= x; k++) {
		double kk = (double)k;
		double t = pow(p, kk) * pow(0.5 - p, xx - kk);
		s += pow(2, xx) * (t * pow(1 - t, ll)) * (double)choose(x, k);
		if (s >= 1.0) {
			s = 1.0;
			break;
		}
	}
	return s;
}

size_t orc_min_anchor_length(double p, double g, size_t l) {
	/* src/sequence.c:296-304 ; pinned by test/test_process.c:16-29 */
	size_t x = 1;
	while (orc_shustring_cum_prob(x, g / 2, l) < 1 - p) x++;
	return x;
}

/* ------------------------------------------------------------------------------------ */
/* suffix array (stands where the reference calls divsufsort, src/esa.c:303)              */

typedef struct {
	const int32_t *rank;
	int32_t n, h;
} dbl_ctx;

static int cmp_second_key(const void *pa, const void *pb, void *vc) {
	const dbl_ctx *c = (const dbl_ctx *)vc;
	int32_t a = *(const int32_t *)pa, b = *(const int32_t *)pb;
	int32_t ra = a + c->h < c->n ? c->rank[a + c->h] : -1;
	int32_t rb = b + c->h < c->n ? c->rank[b + c->h] : -1;
	return (ra > rb) - (ra < rb);
}

typedef struct {
	uint64_t key;
	int32_t pos;
} keyed;

static int cmp_keyed(const void *pa, const void *pb) {
	const keyed *a = (const keyed *)pa, *b = (const keyed *)pb;
	return (a->key > b->key) - (a->key < b->key);
}

int orc_suffix_array(const unsigned char *T, int32_t *SA, int32_t n) {
	/* Prefix doubling: order by the first 8 bytes (missing bytes count as 0, which is below
	 * every text byte, so a proper prefix sorts first), then refine groups with h = 8,16,.. */
	if (!T || !SA || n < 0) return -1;
	if (n == 0) return 0;
	keyed *ks = malloc((size_t)n * sizeof *ks);
	int32_t *rank = malloc((size_t)n * sizeof *rank);
	int32_t *nrank = malloc((size_t)n * sizeof *nrank);
	if (!ks || !rank || !nrank) {
		free(ks), free(rank), free(nrank);
		return -2;
	}
	for (int32_t i = 0; i < n; i++) {
		uint64_t k = 0;
		for (int d = 0; d < 8; d++) k = (k << 8) | (i + d < n ? T[i + d] : 0u);
		ks[i].key = k;
		ks[i].pos = i;
	}
	qsort(ks, (size_t)n, sizeof *ks, cmp_keyed);
	int unresolved = 0;
	for (int32_t j = 0; j < n; j++) {
		SA[j] = ks[j].pos;
		int32_t r = (j > 0 && ks[j].key == ks[j - 1].key) ? rank[SA[j - 1]] : j;
		rank[SA[j]] = r;
		unresolved |= (r != j);
	}
	free(ks);
	for (int32_t h = 8; unresolved; h *= 2) {
		dbl_ctx ctx = {rank, n, h};
		unresolved = 0;
		int32_t j = 0;
		while (j < n) {
			int32_t e = j + 1;
			while (e < n && rank[SA[e]] == rank[SA[j]]) e++;
			if (e - j > 1) {
				qsort_r(SA + j, (size_t)(e - j), sizeof *SA, cmp_second_key, &ctx);
				int32_t head = j;
				for (int32_t k = j; k < e; k++) {
					if (k > j && cmp_second_key(&SA[k - 1], &SA[k], &ctx) != 0) head = k;
					nrank[SA[k]] = head;
					unresolved |= (head != k);
				}
			} else {
				nrank[SA[j]] = j;
			}
			j = e;
		}
		int32_t *t = rank;
		rank = nrank;
		nrank = t;
		if (h > n) break;
	}
	free(rank), free(nrank);
	return 0;
}

/* ------------------------------------------------------------------------------------ */
/* ESA stages (src/esa.c)                                                                */

static int build_lcp(orc_esa *E) {
	/* src/esa.c:373-426 : phi array, then a left-to-right scan in text order that re-uses
	 * l-1 matched characters; LCP[0] = LCP[len] = -1; comparison stops at RS[len] == '\0'. */
	int32_t n = E->len;
	const char *S = E->S;
	const int32_t *SA = E->SA;
	int32_t *LCP = malloc(((size_t)n + 1) * sizeof *LCP);
	int32_t *phi = malloc((size_t)n * sizeof *phi);
	if (!LCP || !phi) {
		free(LCP), free(phi);
		return -2;
	}
	phi[SA[0]] = -1;
	for (int32_t r = 1; r < n; r++) phi[SA[r]] = SA[r - 1];
	int64_t keep = 0;
	for (int32_t pos = 0; pos < n; pos++) {
		int32_t other = phi[pos];
		if (other < 0) {
			phi[pos] = -1; /* now holds PLCP */
			continue;
		}
		while (S[other + keep] == S[pos + keep]) keep++;
		phi[pos] = (int32_t)keep;
		keep = keep > 0 ? keep - 1 : 0;
	}
	LCP[0] = -1;
	LCP[n] = -1;
	for (int32_t r = 1; r < n; r++) LCP[r] = phi[SA[r]];
	free(phi);
	E->LCP = LCP;
	return 0;
}

static int build_cld(orc_esa *E) {
	/* src/esa.c:312-363 restated through its nearest-smaller-value characterisation:
	 * for x in [1, len-1], y = nearest index left of x with LCP[y] <= LCP[x],
	 * k = nearest index right of x with LCP[k] < LCP[x]. Then
	 *   LCP[y] == LCP[x]            -> CLD[y]   = x   (next l-index)
	 *   else LCP[k] <  LCP[y]       -> CLD[y]   = x   (first l-index of the child below y)
	 *   else                         -> CLD[k-1] = x   (first l-index, stored left of k)
	 * CLD[0] = len; CLD[len] is never written by the reference (left 0 here). */
	int32_t n = E->len;
	const int32_t *LCP = E->LCP;
	int32_t *CLD = calloc((size_t)n + 1, sizeof *CLD);
	int32_t *left = malloc(((size_t)n + 1) * sizeof *left);
	int32_t *right = malloc(((size_t)n + 1) * sizeof *right);
	int32_t *stk = malloc(((size_t)n + 2) * sizeof *stk);
	if (!CLD || !left || !right || !stk) {
		free(CLD), free(left), free(right), free(stk);
		return -2;
	}
	int32_t top = 0;
	stk[top++] = 0;
	for (int32_t x = 1; x < n; x++) {
		while (LCP[stk[top - 1]] > LCP[x]) top--;
		left[x] = stk[top - 1];
		stk[top++] = x;
	}
	top = 0;
	stk[top++] = n;
	for (int32_t x = n - 1; x >= 1; x--) {
		while (LCP[stk[top - 1]] >= LCP[x]) top--;
		right[x] = stk[top - 1];
		stk[top++] = x;
	}
	CLD[0] = n;
	for (int32_t x = 1; x < n; x++) {
		int32_t y = left[x], k = right[x];
		if (LCP[y] == LCP[x] || LCP[k] < LCP[y])
			CLD[y] = x;
		else
			CLD[k - 1] = x;
	}
	free(left), free(right), free(stk);
	E->CLD = CLD;
	return 0;
}

static int build_fvc(orc_esa *E) {
	/* src/esa.c:229-245 : FVC[i] = S[SA[i] + LCP[i]], including i = 0 where LCP is -1 */
	int32_t n = E->len;
	char *F = malloc((size_t)n);
	if (!F) return -2;
	for (int32_t r = 0; r < n; r++) F[r] = E->S[E->SA[r] + E->LCP[r]];
	E->FVC = F;
	return 0;
}

static int base_code(char c) {
	/* src/esa.c:49-58 */
	return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}

static int is_empty(orc_interval v) { return v.i == -1 && v.j == -1; }

static orc_interval root_interval(const orc_esa *E) {
	/* src/esa.c:82-83, 620-621 */
	int32_t m = E->CLD[E->len - 1];
	orc_interval r = {.l = E->LCP[m], .i = 0, .j = E->len - 1, .m = m};
	return r;
}

static orc_interval child_interval(const orc_esa *E, orc_interval p, char a) {
	/* src/esa.c:441-511. Children of [p.i, p.j] start at p.i and at each l-index
	 * (p.m, CLD[p.m], ...); the first character of a child is read from the text for the
	 * first child and from FVC for the others. The fields returned for each kind of child
	 * are the ones the reference produces (they end up in the prefix cache). */
	const int32_t *SA = E->SA, *LCP = E->LCP, *CLD = E->CLD;
	orc_interval none = p;
	none.i = none.j = -1;
	if (p.i == p.j) return E->S[SA[p.i] + p.l] == a ? p : none;

	int32_t start = p.i, split = p.m, depth = p.l;
	char c = E->S[SA[start] + depth];
	for (;;) {
		if (c == a) {
			orc_interval r;
			r.i = start;
			if (start == split - 1) {
				r.j = start, r.m = -1, r.l = LCP[start];
			} else {
				r.j = split - 1, r.m = CLD[split - 1], r.l = LCP[r.m];
			}
			return r;
		}
		if (c > a) return none;
		start = split;
		if (start != p.j) {
			split = CLD[split];
			if (LCP[split] == depth) {
				c = E->FVC[start];
				continue;
			}
		}
		/* [start, p.j] is the last child; split is its first l-index (or start itself) */
		if (E->FVC[start] != a) return none;
		orc_interval r = {.l = LCP[split], .i = start, .j = p.j, .m = split};
		return r;
	}
}

static orc_interval descend(const orc_esa *E, const char *q, size_t qlen, int32_t k,
							orc_interval at) {
	/* src/esa.c:531-601 */
	const char *S = E->S;
	if (is_empty(at)) return at;
	if (at.i == at.j) {
		int32_t p = E->SA[at.i];
		size_t d = (size_t)at.l;
		while (d < qlen && S[p + d] && S[p + d] == q[d]) d++;
		at.l = (int32_t)d;
		return at;
	}
	orc_interval best = at;
	do {
		at = child_interval(E, at, q[k]);
		if (is_empty(at)) {
			best.l = k;
			return best;
		}
		best.i = at.i;
		best.j = at.j;
		int32_t upto = (int32_t)qlen;
		if (at.i < at.j && at.l < upto) upto = at.l;
		k++;
		int32_t p = E->SA[at.i];
		for (; k < upto; k++) {
			if (S[p + k] != q[k]) {
				best.l = k;
				return best;
			}
		}
	} while (k < (int32_t)qlen);
	best.l = (int32_t)qlen;
	return best;
}

orc_interval orc_get_match_cld(const orc_esa *E, const char *query, size_t qlen) {
	/* src/esa.c:614-624 */
	return descend(E, query, qlen, 0, root_interval(E));
}

#define ORC_CACHE_LEN 10

orc_interval orc_get_match_cached(const orc_esa *E, const char *query, size_t qlen) {
	/* src/esa.c:636-656 */
	if (qlen <= ORC_CACHE_LEN) return orc_get_match_cld(E, query, qlen);
	int64_t slot = 0;
	for (int d = 0; d < ORC_CACHE_LEN; d++) {
		int c = base_code(query[d]);
		if (c < 0) return orc_get_match_cld(E, query, qlen);
		slot = (slot << 2) | c;
	}
	orc_interval at = E->cache[slot];
	if (is_empty(at)) return orc_get_match_cld(E, query, qlen);
	return descend(E, query, qlen, at.l, at);
}

static orc_interval cache_entry(const orc_esa *E, const char *w) {
	/* src/esa.c:73-215 evaluated for ONE 10-mer w: follow the depth-first filling order of
	 * the reference along the characters of w and report what ends up in w's slot.
	 * `at` is the interval of w[0..pos); the value written for everything below a prefix
	 * that cannot be extended (or whose interval is too deep) is the parent interval. */
	orc_interval at = root_interval(E);
	size_t pos = 0;
	while (pos < ORC_CACHE_LEN) {
		if (is_empty(at)) return at; /* esa.c:104-107 */
		orc_interval sub = child_interval(E, at, w[pos]);
		if (is_empty(sub)) return at; /* esa.c:123-127 */
		if (sub.i == sub.j) {		  /* esa.c:130-135 */
			sub.l = (int32_t)pos + 1;
			return sub;
		}
		if (sub.l <= (int32_t)(pos + 1)) { /* esa.c:137-142 */
			at = sub;
			pos++;
			continue;
		}
		if (sub.l >= ORC_CACHE_LEN) return at; /* esa.c:146-150 */
		/* esa.c:152-186 : the child interval is deeper than pos+1; only the one spelled out by
		 * the text continues, every other extension keeps the parent. */
		size_t k = pos + 1;
		int special = 0;
		for (; k < (size_t)sub.l; k++) {
			char c = E->S[E->SA[sub.i] + k];
			if (base_code(c) < 0) {
				special = 1;
				break;
			}
			if (w[k] != c) return at;
		}
		if (special) return sub; /* esa.c:182-183 : filled from depth k on, with the deep l */
		at = sub;
		pos = k;
	}
	return at;
}

static int build_cache(orc_esa *E) {
	size_t slots = (size_t)1 << (2 * ORC_CACHE_LEN);
	orc_interval *cache = malloc(slots * sizeof *cache);
	if (!cache) return -2;
	char w[ORC_CACHE_LEN + 1];
	w[ORC_CACHE_LEN] = '\0';
	for (size_t s = 0; s < slots; s++) {
		for (int d = 0; d < ORC_CACHE_LEN; d++) w[d] = "ACGT"[(s >> (2 * (ORC_CACHE_LEN - 1 - d))) & 3];
		cache[s] = cache_entry(E, w);
	}
	E->cache = cache;
	return 0;
}

int orc_esa_build(orc_esa *E, const char *RS, int32_t len) {
	/* src/esa.c:254-277 */
	if (!E || !RS) return 1;
	memset(E, 0, sizeof *E);
	E->S = RS;
	E->len = len;
	E->SA = malloc((size_t)len * sizeof *E->SA);
	if (!E->SA) return -2;
	int rc = orc_suffix_array((const unsigned char *)RS, E->SA, len);
	if (!rc) rc = build_lcp(E);
	if (!rc) rc = build_cld(E);
	if (!rc) rc = build_fvc(E);
	if (!rc) rc = build_cache(E);
	if (rc) orc_esa_free(E);
	return rc;
}

void orc_esa_free(orc_esa *E) {
	/* src/esa.c:280-287 */
	free(E->SA), free(E->LCP), free(E->CLD), free(E->FVC), free(E->cache);
	memset(E, 0, sizeof *E);
}

/* ------------------------------------------------------------------------------------ */
/* search, spec form                                                  */

static int32_t common_len(const char *a, const char *b, size_t cap) {
	size_t k = 0;
	while (k < cap && a[k] == b[k]) k++; /* b is RS: its terminating '\0' never equals a[k] */
	return (int32_t)k;
}

orc_interval orc_get_match(const orc_esa *E, const char *query, size_t qlen) {
	/* Longest prefix of the query found anywhere in RS and the exact SA range of suffixes that
	 * start with it: binary search for the query's insertion point, take the better of the
	 * two neighbours, then widen over LCP values >= l. */
	const int32_t n = E->len;
	const unsigned char *S = (const unsigned char *)E->S;
	int32_t lo = 0, hi = n; /* first suffix >= query lies in [lo, hi] */
	while (lo < hi) {
		int32_t mid = lo + (hi - lo) / 2;
		int32_t p = E->SA[mid];
		int32_t c = common_len(query, E->S + p, qlen);
		int suffix_less;
		if ((size_t)c == qlen)
			suffix_less = 0; /* query is a prefix of the suffix */
		else
			suffix_less = S[p + c] < (unsigned char)query[c]; /* S[n] == 0 sorts first */
		if (suffix_less)
			lo = mid + 1;
		else
			hi = mid;
	}
	int32_t lm = lo > 0 ? common_len(query, E->S + E->SA[lo - 1], qlen) : -1;
	int32_t rm = lo < n ? common_len(query, E->S + E->SA[lo], qlen) : -1;
	orc_interval r = {.m = -1};
	if (lm <= 0 && rm <= 0) {
		r.l = 0, r.i = 0, r.j = n - 1;
		return r;
	}
	int32_t at = rm >= lm ? lo : lo - 1;
	r.l = rm >= lm ? rm : lm;
	r.i = r.j = at;
	while (r.i > 0 && E->LCP[r.i] >= r.l) r.i--;
	while (r.j + 1 < n && E->LCP[r.j + 1] >= r.l) r.j++;
	return r;
}

/* ------------------------------------------------------------------------------------ */
/* counting (src/model.c)                                                                */

static int bits_of(char c) {
	/* src/model.c:295-299 : A0 C1 G2 T3 from bits 1..2 of the ASCII code */
	unsigned v = (unsigned char)c & 6u;
	v ^= v >> 1;
	return (int)(v >> 1);
}

void orc_model_count_equal(orc_model *M, const char *q, size_t len, int model_id) {
	/* src/model.c:246-279 */
	if (model_id == ORC_RAW || model_id == ORC_JC || model_id == ORC_KIMURA) {
		uint32_t quarter = (uint32_t)(len / 4);
		M->counts[0] += quarter;
		M->counts[5] += quarter;
		M->counts[10] += quarter;
		M->counts[15] += quarter + (uint32_t)(len & 3);
		return;
	}
	for (size_t k = 0; k < len; k++) {
		if (q[k] < 'A') continue;
		int b = bits_of(q[k]);
		M->counts[b * 5]++;
	}
}

void orc_model_count(orc_model *M, const char *s, const char *q, size_t len) {
	/* src/model.c:309-337 */
	for (size_t k = 0; k < len; k++) {
		if (s[k] < 'A' || q[k] < 'A') continue;
		M->counts[(bits_of(s[k]) << 2) + bits_of(q[k])]++;
	}
}

orc_model orc_model_average(const orc_model *a, const orc_model *b) {
	/* src/model.c:39-46 */
	orc_model r = *a;
	for (int k = 0; k < 16; k++) r.counts[k] += b->counts[k];
	r.seq_len += b->seq_len;
	return r;
}

static size_t total_of(const orc_model *M) {
	size_t t = 0;
	for (int k = 0; k < 16; k++) t += M->counts[k];
	return t;
}

double orc_model_coverage(const orc_model *M) {
	/* src/model.c:68-73 */
	return (double)total_of(M) / (double)M->seq_len;
}

static double est_raw(const orc_model *M) {
	/* src/model.c:81-93 */
	size_t nucl = total_of(M), same = M->counts[0] + (size_t)M->counts[5] + M->counts[10] + M->counts[15];
	if (nucl <= 3) return NAN;
	return (double)(nucl - same) / (double)nucl;
}

double orc_estimate(const orc_model *M, int model_id) {
	const uint32_t *c = M->counts;
	switch (model_id) {
		case ORC_RAW: return est_raw(M);
		case ORC_ANI: return (1.0 - est_raw(M)) * 100; /* src/model.c:206-209 */
		case ORC_KIMURA: {							  /* src/model.c:115-130 */
			size_t nucl = total_of(M);
			size_t ts = (size_t)c[2] + c[8] + c[7] + c[13];
			size_t tv = (size_t)c[1] + c[4] + c[3] + c[12] + c[9] + c[6] + c[11] + c[14];
			double P = (double)ts / (double)nucl, Q = (double)tv / (double)nucl;
			double tmp = 1.0 - 2.0 * P - Q;
			double d = -0.25 * log((1.0 - 2.0 * Q) * tmp * tmp);
			return d <= 0.0 ? 0.0 : d;
		}
		case ORC_LOGDET: { /* src/model.c:161-198 */
			double nucl = (double)total_of(M), P[16];
			for (int k = 0; k < 16; k++) P[k] = c[k] / nucl;
			double ld = 0.0;
			for (int r = 0; r < 4; r++) ld += log(((size_t)c[4 * r] + c[4 * r + 1] + c[4 * r + 2] + c[4 * r + 3]) / nucl);
			for (int q = 0; q < 4; q++) ld += log(((size_t)c[q] + c[4 + q] + c[8 + q] + c[12 + q]) / nucl);
			/* 4x4 determinant by expansion along the first row, 2x2 minors of rows 2,3 */
#define M2(a, b) (P[8 + (a)] * P[12 + (b)] - P[12 + (a)] * P[8 + (b)])
			double det = P[0] * P[5] * M2(2, 3) - P[0] * P[6] * M2(1, 3) + P[0] * P[7] * M2(1, 2) -
						 P[1] * P[4] * M2(2, 3) + P[1] * P[6] * M2(0, 3) - P[1] * P[7] * M2(0, 2) +
						 P[2] * P[4] * M2(1, 3) - P[2] * P[5] * M2(0, 3) + P[2] * P[7] * M2(0, 1) -
						 P[3] * P[4] * M2(1, 2) + P[3] * P[5] * M2(0, 2) - P[3] * P[6] * M2(0, 1);
#undef M2
			double d = -0.25 * (log(det) - 0.5 * ld);
			return d <= 0.0 ? 0.0 : d;
		}
		case ORC_JC:
		default: { /* src/model.c:101-107 */
			double d = est_raw(M);
			d = -0.75 * log(1.0 - (4.0 / 3.0) * d);
			return d <= 0.0 ? 0.0 : d;
		}
	}
}

/* ------------------------------------------------------------------------------------ */
/* the anchor walk (src/process.c:29-214)                                                */

typedef struct {
	size_t s, q, len;
} hit;

static orc_model walk(const orc_esa *E, const char *query, size_t qlen, size_t t, int model_id,
					  int spec_search) {
	orc_model out;
	memset(&out, 0, sizeof out);
	out.seq_len = (uint32_t)qlen;
	hit prev = {0, 0, 0}, cur = {0, 0, 0};
	int prev_paired = 0;
	const size_t half = (size_t)E->len / 2; /* process.c:148 : reverse strand below, forward above */

	while (cur.q < qlen) {
		int found = 0;
		/* process.c:82-100 : try the diagonal of the previous anchor first, no uniqueness test */
		size_t step = cur.q - prev.q;
		size_t hole = step - prev.len;
		size_t guess = prev.s + step;
		if (guess < (size_t)E->len && hole <= t) {
			cur.s = guess;
			cur.len = (size_t)common_len(query + cur.q, E->S + guess, qlen - cur.q);
			found = cur.len >= t;
		}
		if (!found) {
			/* process.c:113-123 */
			orc_interval m = spec_search ? orc_get_match(E, query + cur.q, qlen - cur.q)
										 : orc_get_match_cached(E, query + cur.q, qlen - cur.q);
			cur.s = (size_t)E->SA[m.i];
			cur.len = m.l > 0 ? (size_t)m.l : 0;
			found = (m.i == m.j) && cur.len >= t;
		}
		if (found) {
			/* process.c:160-193 */
			size_t end_s = prev.s + prev.len, end_q = prev.q + prev.len;
			int pairs = cur.s > end_s && (cur.q - end_q) == (cur.s - end_s) &&
						((cur.s < half) == (prev.s < half));
			if (pairs) {
				orc_model_count_equal(&out, query + prev.q, prev.len, model_id);
				orc_model_count(&out, E->S + end_s, query + end_q, cur.q - end_q);
				prev_paired = 1;
			} else {
				if (prev_paired || prev.len >= 2 * t)
					orc_model_count_equal(&out, query + prev.q, prev.len, model_id);
				prev_paired = 0;
			}
			prev = cur;
		}
		cur.q += cur.len + 1; /* process.c:196 */
	}
	/* process.c:199-211 */
	if (prev.len >= qlen) {
		orc_model_count_equal(&out, query, qlen, model_id);
		return out;
	}
	if (prev_paired || prev.len >= 2 * t) orc_model_count_equal(&out, query + prev.q, prev.len, model_id);
	return out;
}

orc_model orc_dist_anchor(const orc_esa *E, const char *query, size_t qlen, size_t threshold,
						  int model_id) {
	return walk(E, query, qlen, threshold, model_id, 0);
}

/* Same walk, but the ESA lookup is the spec (orc_get_match) instead of the reference's cached
 * child-table procedure; the two differ only in the prefix-cache corner documented in
 * DESIGN.md ("known divergence") and SURVEY 7.3-4. */
orc_model orc_dist_anchor_spec(const orc_esa *E, const char *query, size_t qlen, size_t threshold,
							   int model_id) {
	return walk(E, query, qlen, threshold, model_id, 1);
}

int orc_rows(const char *const *seqs, const size_t *lens, size_t n, size_t s_begin, size_t s_end,
			 int model_id, double p_value, orc_model *out) {
	/* src/dist_hack.h:46-90 */
	for (size_t i = s_begin; i < s_end; i++) {
		size_t len = lens[i];
		char *RS = malloc(2 * len + 2);
		if (!RS) return -2;
		orc_make_rs(seqs[i], len, RS);
		size_t t = orc_min_anchor_length(p_value, orc_gc(seqs[i], len), 2 * len + 1);
		orc_esa E;
		int rc = orc_esa_build(&E, RS, (int32_t)(2 * len + 1));
		if (rc) {
			free(RS);
			return rc;
		}
		for (size_t j = 0; j < n; j++) {
			orc_model *cell = &out[(i - s_begin) * n + j];
			if (j == i) {
				memset(cell, 0, sizeof *cell);
				cell->seq_len = 9;
				cell->counts[0] = 9;
			} else {
				*cell = orc_dist_anchor(&E, seqs[j], lens[j], t, model_id);
			}
		}
		orc_esa_free(&E);
		free(RS);
	}
	return 0;
}

/* ------------------------------------------------------------------------------------ */
/* test/test_esa.c:38-44,172-203 ("/esa/full cache"): for ALL ACGT strings of length depth,
 * the cached search, the uncached child-table search and the spec must agree on (l, i, j),
 * the match must spell the query prefix and must be maximal. Returns the number of
 * strings that violate any of these (0 = pass). */
size_t orc_sweep_check(const orc_esa *E, int depth) {
	size_t bad = 0, total = (size_t)1 << (2 * depth);
	char q[40];
	if (depth >= (int)sizeof q) return (size_t)-1;
	q[depth] = '\0';
	for (size_t s = 0; s < total; s++) {
		for (int d = 0; d < depth; d++) q[d] = "ACGT"[(s >> (2 * (depth - 1 - d))) & 3];
		orc_interval a = orc_get_match_cached(E, q, (size_t)depth);
		orc_interval b = orc_get_match_cld(E, q, (size_t)depth);
		orc_interval c = orc_get_match(E, q, (size_t)depth);
		int ok = a.l == b.l && a.i == b.i && a.j == b.j && a.l == c.l && a.i == c.i && a.j == c.j;
		ok = ok && strncmp(q, E->S + E->SA[a.i], (size_t)a.l) == 0;
		ok = ok && (q[a.l] != E->S[a.l + E->SA[a.i]] || q[a.l] == '\0');
		bad += !ok;
	}
	return bad;
}
