// andi_b200/csrc/walk_kernels.cuh -- the anchor walk.
//
// One thread owns one ordered (subject, query) pair and runs the reference's state machine
// (src/process.c:141-214) exactly: lucky anchor on the previous diagonal first
// (process.c:82-100), otherwise a longest-match lookup in the subject index
// (process.c:113-123), then the right-anchor pairing and the one-anchor-late accounting
// (process.c:160-211). All text is 2-bit packed; comparisons are XOR + find-first-set on
// 32-character windows, gap columns are classified with popcounts.
#pragma once
#include "esa_kernels.cuh"

#define ANDI_SCAN_MAX 8

struct SubjectIndex {
	TextView rs;		   // RS planes, len = N = 2n+1, mid = n
	const u32 *SA;		   // N
	const int32_t *LCP;	   // N + 1
	const u32 *dir;		   // 4^K + 1
	PresenceLevels present;
	int K;				   // directory depth, 0 = none (lookups use the generic search only)
	u32 threshold;		   // minimum anchor length for this subject
	u32 self;			   // pool index of the subject (its own query is skipped)
	u32 has_sep;		   // RS contains '!' / ';'
};

struct QueryView {
	TextView t;
	u32 has_sep;
};

struct MatchResult {
	u32 len;	 // longest prefix of the query found in RS
	u32 at;		 // SA index of one suffix carrying it (valid when found_pos)
	bool unique; // exactly one suffix carries it
	bool found_pos;
};

// ---- generic search: binary search of the query among the suffixes SA[lo, hi) in the
// reference's byte order, then the better neighbour; uniqueness from the LCP array.
template <bool SPEC>
__device__ MatchResult search_range(const SubjectIndex &S, const TextView &q, u32 qpos, u32 rem,
									u32 lo, u32 hi) {
	const u32 lo0 = lo, hi0 = hi;
	while (lo < hi) {
		u32 mid = lo + ((hi - lo) >> 1);
		u32 p = S.SA[mid];
		u32 lim = min(rem, rs_run<SPEC>(S.rs, p));
		u32 c = match_len<SPEC>(q, qpos, S.rs, p, lim);
		bool suffix_less;
		if (c == rem)
			suffix_less = false;  // the query is a prefix of this suffix
		else
			suffix_less = sym3<SPEC>(S.rs, p + c) < sym3<true>(q, qpos + c);
		if (suffix_less)
			lo = mid + 1;
		else
			hi = mid;
	}
	MatchResult r;
	r.found_pos = true;
	int lm = -1, rm = -1;
	if (lo > lo0) {
		u32 p = S.SA[lo - 1];
		lm = (int)match_len<SPEC>(q, qpos, S.rs, p, min(rem, rs_run<SPEC>(S.rs, p)));
	}
	if (lo < hi0) {
		u32 p = S.SA[lo];
		rm = (int)match_len<SPEC>(q, qpos, S.rs, p, min(rem, rs_run<SPEC>(S.rs, p)));
	}
	if (lm <= 0 && rm <= 0) {
		r.len = 0, r.at = 0, r.unique = false;
		return r;
	}
	if (rm >= lm) {
		r.len = (u32)rm, r.at = lo;
		r.unique = (rm > lm) && (S.LCP[lo + 1] < rm);
	} else {
		r.len = (u32)lm, r.at = lo - 1;
		r.unique = S.LCP[lo - 1] < lm;
	}
	return r;
}

// ---- the lookup the walk uses. Fast path: k-mer directory -> short scan of the bucket;
// when the k-mer is absent only the LENGTH of the match matters (it is < K <= threshold, so
// no anchor can result) and the presence bitmaps give it without touching the suffix array.
template <bool SPEC>
__device__ MatchResult longest_match(const SubjectIndex &S, const TextView &q, u32 qpos, u32 rem) {
	const int K = S.K;
	bool direct = K > 0 && rem >= (u32)K;
	u64 cw = 0;
	if (direct) {
		cw = window32(q.code, qpos);
		if (SPEC) {
			u64 sw = window32(q.spec, qpos);
			if (sw & ((1ULL << (2 * K)) - 1ULL)) direct = false;
		}
	}
	if (!direct) return search_range<SPEC>(S, q, qpos, rem, 0, S.rs.len);

	u32 key = kmer_key(cw, K);
	u32 lo = __ldg(S.dir + key), hi = __ldg(S.dir + key + 1);
	if (hi > lo) {
		MatchResult r;
		if (hi - lo <= ANDI_SCAN_MAX) {
			u32 best = 0, cnt = 0, at = lo;
			for (u32 c = lo; c < hi; c++) {
				u32 p = __ldg(S.SA + c);
				u32 m = match_len<SPEC>(q, qpos, S.rs, p, min(rem, rs_run<SPEC>(S.rs, p)));
				if (m > best)
					best = m, cnt = 1, at = c;
				else if (m == best)
					cnt++;
			}
			r.len = best, r.at = at, r.unique = cnt == 1, r.found_pos = true;
		} else {
			r = search_range<SPEC>(S, q, qpos, rem, lo, hi);
		}
		if (r.len >= (u32)K) return r;
	}
	MatchResult r;
	r.unique = false, r.found_pos = false, r.at = 0, r.len = 0;
	for (int m = K - 1; m >= 1; m--) {
		u32 x = key >> (2 * (K - m));
		if ((__ldg(S.present.bits + S.present.offset[m] + (x >> 5)) >> (x & 31u)) & 1u) {
			r.len = (u32)m;
			break;
		}
	}
	return r;
}

// ---- counting (src/model.c)
struct Counts {
	u32 c[16];
};

// src/model.c:309-337: classify `len` aligned columns; columns with a separator on either
// side are skipped. 32 columns per step: four "subject is base a" masks, four "query is
// base b" masks, sixteen popcounts.
template <bool SPEC>
__device__ __forceinline__ void count_columns(Counts &M, const TextView &s, u32 ps, const TextView &q,
											  u32 pq, u32 len) {
	for (u32 k = 0; k < len; k += 32) {
		u32 span = min(32u, len - k);
		u64 valid = span == 32 ? ANDI_EVEN_BITS : (ANDI_EVEN_BITS & ((1ULL << (2 * span)) - 1ULL));
		u64 sw = window32(s.code, ps + k), qw = window32(q.code, pq + k);
		if (SPEC) valid &= ~(window32(s.spec, ps + k) | window32(q.spec, pq + k));
		if (span == 1) {  // by far the most common gap: a single substitution
			if (valid) M.c[((u32)sw & 3u) * 4u + ((u32)qw & 3u)]++;
			continue;
		}
		u64 s_lo = sw & ANDI_EVEN_BITS, s_hi = (sw >> 1) & ANDI_EVEN_BITS;
		u64 q_lo = qw & ANDI_EVEN_BITS, q_hi = (qw >> 1) & ANDI_EVEN_BITS;
		u64 sm[4] = {~s_hi & ~s_lo, ~s_hi & s_lo, s_hi & ~s_lo, s_hi & s_lo};
		u64 qm[4] = {~q_hi & ~q_lo, ~q_hi & q_lo, q_hi & ~q_lo, q_hi & q_lo};
#pragma unroll
		for (int a = 0; a < 4; a++)
#pragma unroll
			for (int b = 0; b < 4; b++) M.c[a * 4 + b] += __popcll(sm[a] & qm[b] & valid);
	}
}

// src/model.c:246-279. QUARTER (RAW/JC/KIMURA): len/4 to each diagonal cell, remainder to
// TtoT, no text is read. Otherwise (LOGDET/ANI) the composition of the query slice, separators
// skipped.
template <bool QUARTER, bool SPEC>
__device__ __forceinline__ void count_anchor(Counts &M, const TextView &q, u32 pq, u32 len) {
	if (QUARTER) {
		u32 f = len >> 2;
		M.c[0] += f, M.c[5] += f, M.c[10] += f, M.c[15] += f + (len & 3u);
		return;
	}
	for (u32 k = 0; k < len; k += 32) {
		u32 span = min(32u, len - k);
		u64 valid = span == 32 ? ANDI_EVEN_BITS : (ANDI_EVEN_BITS & ((1ULL << (2 * span)) - 1ULL));
		u64 qw = window32(q.code, pq + k);
		if (SPEC) valid &= ~window32(q.spec, pq + k);
		u64 lo = qw & ANDI_EVEN_BITS, hi = (qw >> 1) & ANDI_EVEN_BITS;
		M.c[0] += __popcll(~hi & ~lo & valid);
		M.c[5] += __popcll(~hi & lo & valid);
		M.c[10] += __popcll(hi & ~lo & valid);
		M.c[15] += __popcll(hi & lo & valid);
	}
}

// ---- src/process.c:141-214 for one pair
template <bool QUARTER, bool SPEC>
__device__ void walk_pair(const SubjectIndex &S, const TextView &q, u32 threshold, u32 *out17) {
	Counts M;
#pragma unroll
	for (int k = 0; k < 16; k++) M.c[k] = 0;
	const u32 qlen = q.len, N = S.rs.len, border = N / 2, t = threshold;
	u32 pos_q = 0;							 // this_match.pos_Q
	u32 cur_s = 0, cur_len = 0;				 // this_match.pos_S, .length
	u32 last_s = 0, last_q = 0, last_len = 0; // last_match
	bool last_paired = false;				 // last_was_right_anchor

	while (pos_q < qlen) {
		const u32 rem = qlen - pos_q;
		bool found = false;
		// process.c:86-99: same diagonal as the previous anchor, no uniqueness test
		u32 advance = pos_q - last_q;
		u32 gap = advance - last_len;
		u32 guess = last_s + advance;
		if (guess < N && gap <= t) {
			cur_s = guess;
			cur_len = match_len<SPEC>(q, pos_q, S.rs, guess, min(rem, rs_run<SPEC>(S.rs, guess)));
			found = cur_len >= t;
		}
		if (!found) {
			// process.c:117-122
			MatchResult m = longest_match<SPEC>(S, q, pos_q, rem);
			cur_len = m.len;
			found = m.unique && m.len >= t;
			if (found) cur_s = __ldg(S.SA + m.at);
		}
		if (found) {
			u32 end_s = last_s + last_len, end_q = last_q + last_len;
			bool pairs = cur_s > end_s && (pos_q - end_q) == (cur_s - end_s) &&
						 ((cur_s < border) == (last_s < border));
			if (pairs) {
				count_anchor<QUARTER, SPEC>(M, q, last_q, last_len);
				count_columns<SPEC>(M, S.rs, end_s, q, end_q, pos_q - end_q);
				last_paired = true;
			} else {
				if (last_paired || last_len >= 2 * t) count_anchor<QUARTER, SPEC>(M, q, last_q, last_len);
				last_paired = false;
			}
			last_s = cur_s, last_q = pos_q, last_len = cur_len;
		}
		pos_q += cur_len + 1;
	}
	// process.c:199-211
	if (last_len >= qlen) {
		count_anchor<QUARTER, SPEC>(M, q, 0, qlen);
	} else if (last_paired || last_len >= 2 * t) {
		count_anchor<QUARTER, SPEC>(M, q, last_q, last_len);
	}
#pragma unroll
	for (int k = 0; k < 16; k++) out17[k] = M.c[k];
	out17[16] = qlen;
}

// Pair p of a batch: subject slot p / nq, query p % nq. A persistent grid pulls pair ids from
// a global counter so finished threads pick up new work immediately.
template <bool QUARTER, bool SPEC>
__global__ void __launch_bounds__(256)
k_walk(const SubjectIndex *__restrict__ subjects, u32 nslots, const QueryView *__restrict__ queries,
	   const u32 *__restrict__ query_ids, u32 nq, u32 threshold_override, u32 *__restrict__ out,
	   unsigned long long *__restrict__ next_pair) {
	const unsigned long long total = (unsigned long long)nslots * nq;
	for (;;) {
		unsigned long long p = atomicAdd(next_pair, 1ULL);
		if (p >= total) return;
		u32 slot = (u32)(p / nq), k = (u32)(p % nq);
		u32 qid = query_ids ? query_ids[k] : k;
		const SubjectIndex &S = subjects[slot];
		u32 *cell = out + (size_t)p * 17;
		if (qid == S.self) {
			// src/dist_hack.h:61-64
			cell[0] = 9;
			for (int c = 1; c < 16; c++) cell[c] = 0;
			cell[16] = 9;
			continue;
		}
		const TextView q = queries[qid].t;
		walk_pair<QUARTER, SPEC>(S, q, threshold_override ? threshold_override : S.threshold, cell);
	}
}

// get_match for a batch of packed queries (tests / andi_esa_get_match): the lookup of the walk
// plus, for the interval bounds the spec asks for, a widening over LCP >= l.
template <bool SPEC>
__global__ void k_get_match(SubjectIndex S, const QueryView *__restrict__ queries, u32 nq,
							Inter *__restrict__ out) {
	u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nq) return;
	const TextView &q = queries[k].t;
	MatchResult m = longest_match<SPEC>(S, q, 0, q.len);
	MatchResult g = search_range<SPEC>(S, q, 0, q.len, 0, S.rs.len);
	Inter r;
	r.m = -1;
	if (m.len != g.len || (m.found_pos && m.unique != g.unique)) {
		r.l = -2, r.i = (int32_t)m.len, r.j = (int32_t)g.len;  // internal disagreement
		out[k] = r;
		return;
	}
	if (g.len == 0) {
		r.l = 0, r.i = 0, r.j = (int32_t)S.rs.len - 1;
		out[k] = r;
		return;
	}
	int32_t i = (int32_t)g.at, j = (int32_t)g.at, l = (int32_t)g.len;
	while (i > 0 && S.LCP[i] >= l) i--;
	while (j + 1 < (int32_t)S.rs.len && S.LCP[j + 1] >= l) j++;
	r.l = l, r.i = i, r.j = j;
	out[k] = r;
}
