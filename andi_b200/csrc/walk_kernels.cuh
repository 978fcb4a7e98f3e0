// andi_b200/csrc/walk_kernels.cuh -- the anchor walk.
//
// The reference walks a query sequentially (src/process.c:141-214); the state that carries
// from one loop iteration to the next is exactly
//     (pos_Q, last_match.{pos_S,pos_Q,length}, last_was_right_anchor)             [WalkState]
// plus the running counts. Two walks that reach the same WalkState behave identically from
// there on. That gives the parallel form used here (exact, not an approximation):
//
//   k_walk_chunks  one thread per (query, chunk). It walks its chunk from an "amnesic" state
//                  (no previous anchor) -> counts U_c and exit state E_c. Then it plays the
//                  TRUE successor: it continues from E_c into chunk c+1 while also replaying
//                  the amnesic walk of chunk c+1, one step at a time, until both chains are in
//                  the same WalkState. The counts of the true chain minus those of the amnesic
//                  chain up to that point are the boundary correction D_c.
//   k_walk_reduce  one warp per pair: total = sum_c (U_c + D_c) while every boundary
//                  synchronised; where one did not (e.g. a long anchor-free stretch) it walks on
//                  sequentially from the last known true state until the chains meet again.
//                  (The kernels of this file stop a boundary's replay at the end of chunk c+1.
//                  k_walk_v3, walk_v3.cuh, lets it run until the chains meet, wherever that is:
//                  its records need no sequential path -- k_walk_reduce_finish(extended = 1).)
//
// All threads of a launch work on ONE subject, so its index (SA, directory, packed text) stays
// resident in the 126 MB L2 while the queries stream through. The main loop has a single back
// edge with a __syncwarp so the 32 walks of a warp stay converged step by step.
#pragma once
#include "esa_kernels.cuh"

#define ANDI_SCAN_MAX 8
#define ANDI_WALK_THREADS 256
#define ANDI_UNIT_WORDS 38	// U[16] D[16] E[5] flag

struct SubjectIndex {
	TextView rs;		   // RS planes, len = N = 2n+1, mid = n
	const u32 *SA;		   // N
	const int32_t *LCP;	   // N + 1
	const u64 *dir;		   // 4^K: first SA index | count << 32 of every k-mer
	const unsigned char *plen;  // 4^K: longest prefix of each k-mer present in RS (< K)
	const u64 *fdir;	   // 4^K: the phase-pipeline kernels' view of dir / plen / SA (k_prefix_len)
	int K;				   // directory depth, 0 = none (lookups use the generic search only)
	u32 threshold;		   // minimum anchor length for this subject
	u32 self;			   // pool index of the subject (its own query is skipped)
	u32 has_sep;		   // RS contains '!' / ';'
	// query side, LOGDET / ANI fast path only: base of the pool's code plane and of its
	// prefix-composition table (same word geometry), so comp = qcomp_base + (code - qcode_base)
	const u64 *qcode_base;
	const uint4 *qcomp_base;
	// every query's spec plane lies at the same word distance from its code plane (pool planes and
	// temporary query planes are each one allocation pair): spec = code + qspec_delta
	long long qspec_delta;
	// SPEC fast path: sep3[j] != 0 iff spec words j, j+1 or j+2 hold a separator (k_sep3); the
	// query's table has the geometry of its code plane: q_sep3 = qsep3_base + (code - qcode_base)
	const unsigned char *s_sep3;
	const unsigned char *qsep3_base;
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
// reference's byte order, then the better neighbour (looked for inside [nlo, nhi));
// uniqueness from the LCP array.
template <bool SPEC>
__device__ MatchResult search_range(const SubjectIndex &S, const TextView &q, u32 qpos, u32 rem,
									u32 lo, u32 hi, u32 nlo, u32 nhi) {
	const u32 lo0 = nlo, hi0 = nhi;
	// Manber-Myers acceleration: the query shares m_lo characters with the suffix just below the
	// current range and m_hi with the one just above it (0 while a bound is still the initial one);
	// every suffix in between shares at least min(m_lo, m_hi) with it, so the compare starts
	// there. In a bucket of copies of a long repeat (all suffixes agree for kilobases) this turns
	// log(bucket) full-length compares into one. (r2: a 2.1 Mbp genome with 30 copies of a
	// 1.5 kbp element walked 2.6x slower than a repeat-free one, all of it here.)
	u32 m_lo = 0, m_hi = 0;
	while (lo < hi) {
		u32 mid = lo + ((hi - lo) >> 1);
		u32 p = S.SA[mid];
		u32 lim = min(rem, rs_run<SPEC>(S.rs, p));
		u32 k0 = min(min(m_lo, m_hi), lim);
		u32 c = k0 + match_len<SPEC>(q, qpos + k0, S.rs, p + k0, lim - k0);
		bool suffix_less;
		if (c == rem)
			suffix_less = false;  // the query is a prefix of this suffix
		else
			suffix_less = sym3<SPEC>(S.rs, p + c) < sym3<SPEC>(q, qpos + c);  // q.mid is never hit
		if (suffix_less)
			lo = mid + 1, m_lo = c;
		else
			hi = mid, m_hi = c;
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
// no anchor can result) and the plen table gives it without touching the suffix array.
template <bool SPEC>
__device__ __forceinline__ MatchResult longest_match(const SubjectIndex &S, const TextView &q, u32 qpos,
													 u32 rem) {
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
	if (!direct) {
		// A separator among the first K query characters, or a query tail shorter than K. The
		// padded bucket key (sa_bucket.cuh) is monotone in suffix order, so the query's place
		// among the suffixes lies inside the bucket of ITS padded key: [end of bucket key-1,
		// end of bucket key]. The binary search starts from that range instead of [0, N).
		u32 lo = 0, hi = S.rs.len;
		if (K > 0 && rem > 0) {
			u32 r = min(rem, (u32)K);
			u64 w = window32(q.code, qpos);
			if (SPEC) {
				u64 sw = window32(q.spec, qpos);
				if (sw) r = min(r, (u32)(__ffsll((long long)sw) - 1) >> 1);
			}
			w &= (1ULL << (2u * r)) - 1ULL;
			u32 key = kmer_key(w, K);
			u64 de = __ldg(S.dir + key);
			bool known = (u32)de != 0xffffffffu;
			u32 end = (u32)de + (u32)(de >> 32), start = 0;
			if (key) {
				u64 dp = __ldg(S.dir + key - 1);
				known = known && (u32)dp != 0xffffffffu;
				start = (u32)dp + (u32)(dp >> 32);
			}
			if (known && start <= end && end <= S.rs.len) lo = start, hi = end;
		}
		return search_range<SPEC>(S, q, qpos, rem, lo, hi, 0, S.rs.len);
	}

	u32 key = kmer_key(cw, K);
	u64 de = __ldg(S.dir + key);
	u32 lo = (u32)de, hi = lo + (u32)(de >> 32);
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
			r = search_range<SPEC>(S, q, qpos, rem, lo, hi, lo, hi);
		}
		if (r.len >= (u32)K) return r;
	}
	MatchResult r;
	r.unique = false, r.found_pos = false, r.at = 0, r.len = __ldg(S.plen + key);
	return r;
}

// ---- count accumulators. add(cell, v) adds to one of the 16 cells of src/model.h:14-32.
// Shared-memory flavour: one column of a [16][threads] array per thread (conflict free,
// dynamically indexable without local memory); `sign` is +1 or -1 (as u32) so the amnesic
// chain of a boundary replay can be subtracted in place.
struct SharedAcc {
	u32 *col;  // &cells[0][threadIdx.x]
	u32 sign;
	__device__ __forceinline__ void add(u32 cell, u32 v) { col[cell * ANDI_WALK_THREADS] += v * sign; }
};
struct LocalAcc {
	u32 c[16];
	__device__ __forceinline__ void add(u32 cell, u32 v) { c[cell] += v; }
};

// src/model.c:309-337: classify `len` aligned columns; columns with a separator on either
// side are skipped. 32 columns per step: four "subject is base a" masks, four "query is
// base b" masks, sixteen popcounts.
template <bool SPEC, class Acc>
__device__ __forceinline__ void count_columns(Acc &M, const TextView &s, u32 ps, const TextView &q, u32 pq,
											  u32 len) {
	for (u32 k = 0; k < len; k += 32) {
		u32 span = min(32u, len - k);
		u64 valid = span == 32 ? ANDI_EVEN_BITS : (ANDI_EVEN_BITS & ((1ULL << (2 * span)) - 1ULL));
		u64 sw = window32(s.code, ps + k), qw = window32(q.code, pq + k);
		if (SPEC) valid &= ~(window32(s.spec, ps + k) | window32(q.spec, pq + k));
		if (span == 1) {  // by far the most common gap: a single substitution
			if (valid) M.add(((u32)sw & 3u) * 4u + ((u32)qw & 3u), 1);
			continue;
		}
		u64 s_lo = sw & ANDI_EVEN_BITS, s_hi = (sw >> 1) & ANDI_EVEN_BITS;
		u64 q_lo = qw & ANDI_EVEN_BITS, q_hi = (qw >> 1) & ANDI_EVEN_BITS;
		u64 sm[4] = {~s_hi & ~s_lo, ~s_hi & s_lo, s_hi & ~s_lo, s_hi & s_lo};
		u64 qm[4] = {~q_hi & ~q_lo, ~q_hi & q_lo, q_hi & ~q_lo, q_hi & q_lo};
#pragma unroll
		for (int a = 0; a < 4; a++)
#pragma unroll
			for (int b = 0; b < 4; b++) {
				u32 n = __popcll(sm[a] & qm[b] & valid);
				if (n) M.add(a * 4 + b, n);
			}
	}
}

// src/model.c:246-279. QUARTER (RAW/JC/KIMURA): len/4 to each diagonal cell, remainder to
// TtoT, no text is read. Otherwise (LOGDET/ANI) the composition of the query slice, separators
// skipped.
template <bool QUARTER, bool SPEC, class Acc>
__device__ __forceinline__ void count_anchor(Acc &M, const TextView &q, u32 pq, u32 len) {
	if (QUARTER) {
		u32 f = len >> 2;
		M.add(0, f), M.add(5, f), M.add(10, f), M.add(15, f + (len & 3u));
		return;
	}
	u32 a = 0, c = 0, g = 0, t = 0;
	for (u32 k = 0; k < len; k += 32) {
		u32 span = min(32u, len - k);
		u64 valid = span == 32 ? ANDI_EVEN_BITS : (ANDI_EVEN_BITS & ((1ULL << (2 * span)) - 1ULL));
		u64 qw = window32(q.code, pq + k);
		if (SPEC) valid &= ~window32(q.spec, pq + k);
		u64 lo = qw & ANDI_EVEN_BITS, hi = (qw >> 1) & ANDI_EVEN_BITS;
		a += __popcll(~hi & ~lo & valid);
		c += __popcll(~hi & lo & valid);
		g += __popcll(hi & ~lo & valid);
		t += __popcll(hi & lo & valid);
	}
	M.add(0, a), M.add(5, c), M.add(10, g), M.add(15, t);
}

// ---- separator hints for the SPEC fast path: out[j] = spec words j, j+1, j+2 hold a separator.
// A 64-base window starting in word j touches exactly these words, so one byte load decides
// whether the spec planes have to be read at all. (On concatenated planes a hint may look into
// the next sequence: a false positive only costs the spec loads.)
__global__ void k_sep3(const u64 *__restrict__ spec, size_t nwords, unsigned char *__restrict__ out) {
	size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= nwords) return;
	u64 x = spec[j];
	if (j + 1 < nwords) x |= spec[j + 1];
	if (j + 2 < nwords) x |= spec[j + 2];
	out[j] = x != 0;
}

// ---- prefix composition of the pool (LOGDET / ANI fast path): entry j of a sequence holds the
// number of A, C, G, T among its first 32*j bases, so the composition of any slice
// (src/model.c:259-278 counts an anchor by its characters) is two entries plus two partial
// words. One block per sequence, tiles of blockDim words with a running carry. Separators
// are not counted.
__global__ void __launch_bounds__(256) k_comp_prefix(const QueryView *__restrict__ queries,
													 const u64 *__restrict__ pool_code, uint4 *__restrict__ pool_comp) {
	const u64 *code = queries[blockIdx.x].t.code, *spec = queries[blockIdx.x].t.spec;
	const u32 len = queries[blockIdx.x].t.len, entries = (len >> 5) + 1;
	uint4 *out = pool_comp + (code - pool_code);
	__shared__ uint4 warp_sum[8];
	const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	uint4 carry = make_uint4(0, 0, 0, 0);
	for (u32 base = 0; base < entries; base += blockDim.x) {
		u32 j = base + threadIdx.x;
		uint4 v = make_uint4(0, 0, 0, 0);
		if (j < entries) {
			u32 nb = min(32u, len - 32u * j);
			u64 valid = nb == 32u ? ANDI_EVEN_BITS : (ANDI_EVEN_BITS & ((1ULL << (2u * nb)) - 1ULL));
			valid &= ~spec[j];	// separators are not counted
			u64 w = code[j], lo = w & valid, hi = (w >> 1) & valid;
			v.y = (u32)__popcll(lo & ~hi), v.z = (u32)__popcll(hi & ~lo), v.w = (u32)__popcll(hi & lo);
			v.x = (u32)__popcll(valid) - v.y - v.z - v.w;
		}
		uint4 inc = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			u32 x = __shfl_up_sync(0xffffffffu, inc.x, d), y = __shfl_up_sync(0xffffffffu, inc.y, d);
			u32 z = __shfl_up_sync(0xffffffffu, inc.z, d), w = __shfl_up_sync(0xffffffffu, inc.w, d);
			if (lane >= (u32)d) inc.x += x, inc.y += y, inc.z += z, inc.w += w;
		}
		if (lane == 31u) warp_sum[wid] = inc;
		__syncthreads();
		uint4 pre = carry, tot = carry;
		for (u32 k = 0; k < 8; k++) {
			uint4 ws = warp_sum[k];
			if (k < wid) pre.x += ws.x, pre.y += ws.y, pre.z += ws.z, pre.w += ws.w;
			tot.x += ws.x, tot.y += ws.y, tot.z += ws.z, tot.w += ws.w;
		}
		if (j < entries)
			out[j] = make_uint4(pre.x + inc.x - v.x, pre.y + inc.y - v.y, pre.z + inc.z - v.z, pre.w + inc.w - v.w);
		carry = tot;
		__syncthreads();
	}
}

// ---- the state that carries across iterations of the loop at src/process.c:153-197
struct WalkState {
	u32 pos_q;						 // this_match.pos_Q
	u32 last_s, last_q, last_len;	 // last_match
	u32 paired;						 // last_was_right_anchor
};

__device__ __forceinline__ bool same_state(const WalkState &a, const WalkState &b) {
	return a.pos_q == b.pos_q && a.last_s == b.last_s && a.last_q == b.last_q && a.last_len == b.last_len &&
		   a.paired == b.paired;
}

__device__ __forceinline__ WalkState amnesic_state(u32 pos) {
	WalkState w;
	w.pos_q = pos, w.last_s = 0, w.last_q = 0, w.last_len = 0, w.paired = 0;
	return w;
}

// One iteration of the loop at src/process.c:153-197.
template <bool QUARTER, bool SPEC, class Acc>
__device__ __forceinline__ void walk_step(const SubjectIndex &S, const TextView &q, u32 t, WalkState &w,
										  Acc &M) {
	const u32 qlen = q.len, N = S.rs.len, border = N / 2;
	const u32 pos_q = w.pos_q, rem = qlen - pos_q;
	u32 cur_s = 0, cur_len = 0;
	bool found = false;
	// process.c:86-99: same diagonal as the previous anchor, no uniqueness test
	u32 advance = pos_q - w.last_q;
	u32 gap = advance - w.last_len;
	u32 guess = w.last_s + advance;
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
		// process.c:160-193
		u32 end_s = w.last_s + w.last_len, end_q = w.last_q + w.last_len;
		bool pairs = cur_s > end_s && (pos_q - end_q) == (cur_s - end_s) && ((cur_s < border) == (w.last_s < border));
		if (pairs) {
			count_anchor<QUARTER, SPEC>(M, q, w.last_q, w.last_len);
			count_columns<SPEC>(M, S.rs, end_s, q, end_q, pos_q - end_q);
			w.paired = 1;
		} else {
			if (w.paired || w.last_len >= 2 * t) count_anchor<QUARTER, SPEC>(M, q, w.last_q, w.last_len);
			w.paired = 0;
		}
		w.last_s = cur_s, w.last_q = pos_q, w.last_len = cur_len;
	}
	w.pos_q = pos_q + cur_len + 1;	// process.c:196
}

// src/process.c:199-211: what is still owed once the loop has ended in state w.
template <bool QUARTER, bool SPEC, class Acc>
__device__ __forceinline__ void walk_tail(const TextView &q, u32 t, const WalkState &w, Acc &M) {
	if (w.last_len >= q.len) {
		count_anchor<QUARTER, SPEC>(M, q, 0, q.len);
	} else if (w.paired || w.last_len >= 2 * t) {
		count_anchor<QUARTER, SPEC>(M, q, w.last_q, w.last_len);
	}
}

// ---- k_walk_chunks: see the header of this file. Record layout per unit (ANDI_UNIT_WORDS):
// [0,16) U_c   counts of the amnesic walk of chunk c
// [16,32) D_c  (true chain - amnesic chain) of chunk c+1 up to their meeting point
// [32,37) E_c  exit state of the amnesic walk of chunk c
// [37]    1 if the boundary into chunk c+1 synchronised (or c is the last chunk), else 0
template <bool QUARTER, bool SPEC>
__global__ void __launch_bounds__(ANDI_WALK_THREADS)
k_walk_chunks(const SubjectIndex S, const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids,
			  u32 nq, u32 chunk, u32 cpq, u32 threshold, u32 *__restrict__ records,
				   unsigned long long *__restrict__ next_unit) {
	__shared__ u32 cells[2][16][ANDI_WALK_THREADS];
	const u32 tid = threadIdx.x;
	const unsigned long long total = (unsigned long long)nq * cpq;
	// units are handed out dynamically: their cost varies with the divergence of the pair
	unsigned long long unit = atomicAdd(next_unit, 1ULL);
	const u32 t = threshold;

	bool active = false;
	int phase = 1;
	TextView q;
	q.code = nullptr, q.spec = nullptr, q.len = 0, q.mid = 0xffffffffu;
	WalkState A = amnesic_state(0), B = amnesic_state(0), E = amnesic_state(0);
	u32 a_true = 1;			   // phase 2: A is the true chain (B the amnesic one) or the other way round
	u32 c_end = 0, c2_end = 0;  // end of the own chunk / of the next chunk

	for (;;) {
		if (!active) {
			while (unit < total) {
				u32 k = (u32)(unit / cpq), c = (u32)(unit % cpq);
				u32 qid = query_ids ? query_ids[k] : k;
				u32 qlen = queries[qid].t.len;
				unsigned long long cs = (unsigned long long)c * chunk;
				if (qid != S.self && cs < qlen) {
					q = queries[qid].t;
					A = amnesic_state((u32)cs);
					c_end = (u32)min((unsigned long long)qlen, cs + chunk);
					c2_end = (u32)min((unsigned long long)qlen, cs + 2ULL * chunk);
					phase = 1;
					active = true;
#pragma unroll
					for (int x = 0; x < 16; x++) cells[0][x][tid] = 0, cells[1][x][tid] = 0;
					break;
				}
				unit = atomicAdd(next_unit, 1ULL);
			}
		}
		if (!__any_sync(0xffffffffu, active)) break;
		if (active) {
			bool finished = false;
			u32 flag = 1;
			if (phase == 1) {
				SharedAcc acc = {&cells[0][0][tid], 1u};
				walk_step<QUARTER, SPEC>(S, q, t, A, acc);
				if (A.pos_q >= c_end) {
					E = A;
					if (c_end >= q.len) {
						finished = true;  // last chunk: nothing follows
					} else {
						phase = 2;
						B = amnesic_state(c_end);
						a_true = 1;
					}
				}
			} else {
				// A and B are the two chains inside chunk c+1; a_true says which one A is.
				const WalkState &T = a_true ? A : B;
				const WalkState &P = a_true ? B : A;
				if (same_state(A, B)) {
					finished = true;
				} else if (T.pos_q >= c2_end || P.pos_q >= c2_end) {
					finished = true, flag = 0;	// no meeting point inside chunk c+1
				} else {
					// advance the chain that is behind (the true chain on a tie)
					bool step_true = T.pos_q <= P.pos_q;
					if (step_true != (a_true != 0)) {
						WalkState tmp = A;
						A = B, B = tmp;
						a_true ^= 1u;
					}
					SharedAcc acc = {&cells[1][0][tid], a_true ? 1u : 0xffffffffu};
					walk_step<QUARTER, SPEC>(S, q, t, A, acc);
				}
			}
			if (finished) {
				u32 *rec = records + unit * ANDI_UNIT_WORDS;
#pragma unroll
				for (int x = 0; x < 16; x++) rec[x] = cells[0][x][tid];
#pragma unroll
				for (int x = 0; x < 16; x++) rec[16 + x] = flag ? cells[1][x][tid] : 0u;
				rec[32] = E.pos_q, rec[33] = E.last_s, rec[34] = E.last_q, rec[35] = E.last_len, rec[36] = E.paired;
				rec[37] = flag;
				active = false;
				unit = atomicAdd(next_unit, 1ULL);
			}
		}
		__syncwarp();
	}
}

// ---- the reduction of the unit records of one pair, see the header of this file.
//
// reduce_pair_sequential: one WARP per pair. Lanes sum the records of 32 consecutive chunks at a
// time; the first boundary that did not synchronise (if any) is located with a ballot and lane 0
// carries the true chain on sequentially from there. Exact for every input; used directly by the
// round-1 path and as the fallback of the parallel form below.
//
// (r2) k_walk_reduce_sum + k_walk_reduce_finish: with all boundaries synchronised (the rule) the
// result is a plain sum over the records plus the tail, and one warp per pair is far too little
// parallelism for few, long queries (config 2: 28 warps on the whole GPU, 244 us per subject, a
// quarter of the walk; config 5: 2 warps, 5.3 ms). Now a grid of (pair, slice of chunks) CTAs sums
// the records -- lane x of a warp accumulates word x of every record: coalesced -- and adds them
// to the pair's cell with atomics, noting whether any boundary failed; the finish kernel adds the
// tail (src/process.c:199-211) or, for a pair with a failed boundary, redoes it sequentially.
template <bool QUARTER, bool SPEC>
__device__ void reduce_pair_sequential(const SubjectIndex &S, const TextView &q, u32 t, u32 chunk, const u32 *base, u32 *cell) {
	const u32 lane = threadIdx.x & 31u;
	const u32 qlen = q.len;
	const u32 nch = (u32)(((unsigned long long)qlen + chunk - 1) / chunk);
	LocalAcc total;	 // per lane partial sums
#pragma unroll
	for (int x = 0; x < 16; x++) total.c[x] = 0;
	WalkState fin = amnesic_state(0);  // meaningful on lane 0 only
	bool have_fin = false;			   // lane 0: fin was produced by the sequential path at the very end
	u32 c0 = 0;
	while (c0 < nch) {
		u32 c = c0 + lane;
		bool valid = c < nch;
		const u32 *rec = base + (unsigned long long)c * ANDI_UNIT_WORDS;
		u32 flag = valid ? rec[37] : 1u;
		unsigned bad = __ballot_sync(0xffffffffu, valid && !flag);
		u32 first_bad = bad ? (u32)(__ffs((int)bad) - 1) : 32u;
		if (valid && lane <= first_bad) {
#pragma unroll
			for (int x = 0; x < 16; x++) total.c[x] += rec[x] + rec[16 + x];
		}
		if (!bad) {
			c0 += 32;
			continue;
		}
		// The boundary after chunk cb did not synchronise: the true chain continues from E_cb.
		u32 cb = c0 + first_bad;
		u32 resume = 0xffffffffu;
		if (lane == 0) {
			const u32 *rb = base + (unsigned long long)cb * ANDI_UNIT_WORDS;
			WalkState T;
			T.pos_q = rb[32], T.last_s = rb[33], T.last_q = rb[34], T.last_len = rb[35], T.paired = rb[36];
			while (T.pos_q < qlen) {
				u32 cc = T.pos_q / chunk;  // chunk the true chain is in
				u32 cc_end = (u32)min((unsigned long long)qlen, (unsigned long long)(cc + 1) * chunk);
				WalkState P = amnesic_state(cc * chunk);
				LocalAcc minus;
#pragma unroll
				for (int x = 0; x < 16; x++) minus.c[x] = 0;
				bool met = false;
				for (;;) {
					if (same_state(T, P)) {
						met = true;
						break;
					}
					if (T.pos_q >= cc_end || P.pos_q >= cc_end) break;
					if (T.pos_q <= P.pos_q)
						walk_step<QUARTER, SPEC>(S, q, t, T, total);
					else
						walk_step<QUARTER, SPEC>(S, q, t, P, minus);
				}
				if (met) {
					// from here the amnesic walk of chunk cc IS the true walk: take its record,
					// minus what the amnesic chain had counted before the meeting point
#pragma unroll
					for (int x = 0; x < 16; x++) total.c[x] -= minus.c[x];
					resume = cc;
					break;
				}
				while (T.pos_q < cc_end) walk_step<QUARTER, SPEC>(S, q, t, T, total);
			}
			if (resume == 0xffffffffu) fin = T, have_fin = true;
		}
		resume = __shfl_sync(0xffffffffu, resume, 0);
		if (resume == 0xffffffffu) break;  // the true chain reached the end of the query
		c0 = resume;
	}
	if (lane == 0 && !have_fin) {
		const u32 *rl = base + (unsigned long long)(nch - 1) * ANDI_UNIT_WORDS;
		fin.pos_q = rl[32], fin.last_s = rl[33], fin.last_q = rl[34], fin.last_len = rl[35], fin.paired = rl[36];
	}
	if (lane == 0) walk_tail<QUARTER, SPEC>(q, t, fin, total);
#pragma unroll
	for (int x = 0; x < 16; x++) {
		u32 v = total.c[x];
		for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
		if (lane == 0) cell[x] = v;
	}
	if (lane == 0) cell[16] = qlen;
}

template <bool QUARTER, bool SPEC>
__global__ void __launch_bounds__(128)
k_walk_reduce(const SubjectIndex S, const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids,
			  u32 nq, u32 chunk, u32 cpq, u32 threshold, const u32 *__restrict__ records, u32 *__restrict__ out) {
	const u32 lane = threadIdx.x & 31u;
	const u32 k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (k >= nq) return;  // whole warp
	u32 qid = query_ids ? query_ids[k] : k;
	u32 *cell = out + (size_t)k * 17;
	if (qid == S.self) {
		// src/dist_hack.h:61-64
		if (lane < 17) cell[lane] = (lane == 0 || lane == 16) ? 9u : 0u;
		return;
	}
	reduce_pair_sequential<QUARTER, SPEC>(S, queries[qid].t, threshold, chunk, records + (unsigned long long)k * cpq * ANDI_UNIT_WORDS, cell);
}

#define ANDI_REDUCE_SLICE 2048u	 // chunks per CTA of k_walk_reduce_sum

// out (nq cells, zeroed by the caller) += sum over the records of the slice; bad[k] (zeroed by the
// caller) = 0xffffffff - c for the first chunk c of pair k whose boundary did not synchronise. Grid: (nq, slices); 256 threads = 8 warps, one record per warp and turn.
__global__ void __launch_bounds__(256)
k_walk_reduce_sum(const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids, u32 self, u32 chunk, u32 cpq,
				  const u32 *__restrict__ records, u32 *__restrict__ out, u32 *__restrict__ bad, unsigned long long *__restrict__ stat) {
	const u32 k = blockIdx.x, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;	 // grid: (pair, slice)
	const u32 qid = query_ids ? query_ids[k] : k;
	if (qid == self) return;
	const u32 qlen = queries[qid].t.len;
	const u32 nch = (u32)(((unsigned long long)qlen + chunk - 1) / chunk);
	const u32 c0 = blockIdx.y * ANDI_REDUCE_SLICE, c1 = min(nch, c0 + ANDI_REDUCE_SLICE);
	if (c0 >= nch) return;
	const u32 *base = records + (unsigned long long)k * cpq * ANDI_UNIT_WORDS;
	u32 sum = 0, first_bad = 0xffffffffu, anchor_sum = 0;
	for (u32 c = c0 + wid; c < c1; c += 8u) {
		const u32 *rec = base + (unsigned long long)c * ANDI_UNIT_WORDS;
		sum += rec[lane];												 // lanes 0..15: U_c, lanes 16..31: D_c
		if (lane == 5u && rec[37] == 0u) first_bad = min(first_bad, c);	 // the flag of the boundary behind chunk c
		if (lane == 6u) anchor_sum += min(rec[35], 1u << 20);				 // length of the anchor the chunk's walk ended with
	}
	// stat[0] += those lengths, stat[1] += chunks: their mean tells the host what kind of pool this is
	// (a few dozen bases: ordinary divergence; kilobases: near-identical genomes -> k_walk_v3<.., BURST>)
	if (stat && lane == 6u && c0 + wid < c1) {
		atomicAdd(stat, (unsigned long long)anchor_sum);
		atomicAdd(stat + 1, (unsigned long long)((c1 - c0 - wid + 7u) / 8u));
	}
	sum += __shfl_down_sync(0xffffffffu, sum, 16);	// lane x < 16: U[x] + D[x]
	__shared__ u32 part[8][16];
	__shared__ u32 s_bad;
	if (threadIdx.x == 0) s_bad = 0;
	__syncthreads();
	if (lane < 16u) part[wid][lane] = sum;
	if (first_bad != 0xffffffffu) atomicMax(&s_bad, 0xffffffffu - first_bad);
	__syncthreads();
	if (threadIdx.x < 16u) {
		u32 v = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) v += part[w][threadIdx.x];
		if (v) atomicAdd(out + (size_t)k * 17 + threadIdx.x, v);
	}
	if (threadIdx.x == 0 && s_bad) atomicMax(bad + k, s_bad);
}

// One warp per pair: the diagonal cell, the tail of a pair whose boundaries all synchronised, or the
// whole pair again, sequentially, where one did not.
template <bool QUARTER, bool SPEC>
__global__ void __launch_bounds__(128)
k_walk_reduce_finish(const SubjectIndex S, const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids, u32 nq,
					 u32 chunk, u32 cpq, u32 threshold, const u32 *__restrict__ records, const u32 *__restrict__ bad,
					 u32 *__restrict__ out, u32 extended) {
	const u32 lane = threadIdx.x & 31u;
	const u32 k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (k >= nq) return;  // whole warp
	const u32 qid = query_ids ? query_ids[k] : k;
	u32 *cell = out + (size_t)k * 17;
	if (qid == S.self) {
		// src/dist_hack.h:61-64
		if (lane < 17) cell[lane] = (lane == 0 || lane == 16) ? 9u : 0u;
		return;
	}
	const TextView q = queries[qid].t;
	const u32 *base = records + (unsigned long long)k * cpq * ANDI_UNIT_WORDS;
	// extended (records of k_walk_v3): every boundary's correction is valid wherever its chains met, the
	// sum is complete; chains that reached the end of the query apart left the true final state in the
	// record of the first such chunk. Else (k_walk_chunks, k_walk_chunks_fast: the replay stops at the end
	// of the next chunk): the pair is redone sequentially from its first unsynchronised boundary.
	if (bad[k] && !extended) {
		reduce_pair_sequential<QUARTER, SPEC>(S, q, threshold, chunk, base, cell);
		return;
	}
	if (lane == 0) {
		const u32 nch = (u32)(((unsigned long long)q.len + chunk - 1) / chunk);
		const u32 *rl = base + (unsigned long long)(bad[k] ? 0xffffffffu - bad[k] : nch - 1) * ANDI_UNIT_WORDS;
		WalkState fin;
		fin.pos_q = rl[32], fin.last_s = rl[33], fin.last_q = rl[34], fin.last_len = rl[35], fin.paired = rl[36];
		LocalAcc tail;
#pragma unroll
		for (int x = 0; x < 16; x++) tail.c[x] = 0;
		walk_tail<QUARTER, SPEC>(q, threshold, fin, tail);
#pragma unroll
		for (int x = 0; x < 16; x++) cell[x] += tail.c[x];
		cell[16] = q.len;
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
	MatchResult g = search_range<SPEC>(S, q, 0, q.len, 0, S.rs.len, 0, S.rs.len);
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
