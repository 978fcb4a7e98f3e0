// andi_b200/csrc/sa_bucket.cuh -- suffix array by k-mer bucketing, the fast
// front end of the index build. Stands where src/esa.c:303 calls divsufsort.
//
//   k_bucket_hist     count suffixes per k-mer bucket (atomics on an L2-resident table)
//   (exclusive scan)  bucket starts
//   k_bucket_scatter  drop every suffix into its bucket
//   k_bucket_sort     one thread per bucket: order its (usually 1-3) suffixes by direct
//                     comparison of the packed text, at most ANDI_SORT_CAP characters deep,
//                     and emit what the rest of the build needs: group ids / ranks for the
//                     doubling rounds, the k-mer directory entry of the walk, presence bits
//
// Suffixes that are still tied afterwards (repeats longer than the cap, buckets larger than
// ANDI_SORT_MAX) are finished by the prefix-doubling rounds of andi_b200.cu starting at h = K;
// genome-like text without long repeats needs none.
//
// Bucket key of a suffix = its first K characters as 2-bit codes, with everything from the
// first separator / the text end onwards replaced by 'A' (0). The reference's byte order puts
// every separator below 'A', so this key is monotone in suffix order; the suffixes that were
// padded sort to the front of their bucket by the full comparison.
#pragma once
#include "text.cuh"

#define ANDI_SORT_MAX 16   // buckets up to this size are sorted by one thread
#define ANDI_SORT_CAP 128  // characters compared before two suffixes are declared tied
// Buckets the capped per-thread sort cannot finish (suffixes that agree on ANDI_SORT_CAP characters, or
// more than ANDI_SORT_MAX of them) go to a list and are sorted by one warp each with compares to the
// end (k_sort_deep): that is where the repeats of real genomes (IS elements, rRNA operons: a few
// dozen copies, kilobases long) end up. Whatever does not fit -- too many such buckets, a bucket
// beyond ANDI_DEEP_BUCKET, a match beyond ANDI_DEEP_CAP (tandem repeats, poly-A) -- raises flags[0]
// and takes the prefix-doubling rounds as before.
#define ANDI_DEEP_BUCKET 64u
#define ANDI_DEEP_CAP 16384u
struct TieSink {
	u32 *flags;	 // [0] suffixes left tied (-> doubling rounds), [1] LCP values left capped (-> phi), [4] / [5] list lengths
	u32 *deep;	 // (first slot, end slot) of the listed buckets; behind them the slots of the listed LCP values
	u32 deep_cap, lcp_cap;
};
// true = listed; false = no room (or no list: join mode), the caller reports the ties the old way
__device__ __forceinline__ bool deep_append(const TieSink &t, u32 b, u32 e) {
	if (!t.deep || e - b > ANDI_DEEP_BUCKET) return false;
	const u32 at = atomicAdd(t.flags + 4, 1u);
	if (at >= t.deep_cap) return false;
	t.deep[2 * at] = b, t.deep[2 * at + 1] = e;
	return true;
}

// Number of leading nucleotides of the suffix at p (capped at 32) and its padded key.
__device__ __forceinline__ u32 padded_key(const TextView &rs, u32 p, int K, u32 &run) {
	u64 sw = window32(rs.spec, p);
	u32 r = sw ? (u32)(__ffsll((long long)sw) - 1) >> 1 : 32u;
	r = min(r, rs.len - p);
	run = r;
	u64 cw = window32(rs.code, p);
	if (r < 32u) cw &= (1ULL << (2u * r)) - 1ULL;
	return kmer_key(cw, K);
}

__global__ void k_bucket_hist(TextView rs, int K, u32 *__restrict__ hist) {
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= rs.len) return;
	u32 run;
	atomicAdd(hist + padded_key(rs, i, K, run), 1u);
}

__global__ void k_bucket_scatter(TextView rs, int K, u32 *__restrict__ cursor, u32 *__restrict__ SA) {
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= rs.len) return;
	u32 run;
	u32 key = padded_key(rs, i, K, run);
	SA[atomicAdd(cursor + key, 1u)] = i;  // cursor starts at the bucket start
}

// Compare suffixes a != b of RS in the reference's byte order, looking at most `cap`
// characters deep. Returns <0, >0, or 0 when they agree on the first `cap` characters.
// SPEC=false: RS holds no separator but '#' (handled by position, no spec-plane loads).
template <bool SPEC>
__device__ __forceinline__ int compare_suffixes(const TextView &rs, u32 a, u32 b, u32 cap) {
	u32 lim = min(cap, SPEC ? rs.len - max(a, b) : pair_limit_fast(rs, a, b));
	u32 m = match_len<SPEC>(rs, a, rs, b, lim);
	if (m == cap) return 0;
	u32 sa = sym3<SPEC>(rs, a + m), sb = sym3<SPEC>(rs, b + m);
	return sa < sb ? -1 : 1;  // different suffixes always differ here (end of text ranks lowest)
}

// Does the suffix at p have a separator or the text end inside its first K characters?
template <bool SPEC>
__device__ __forceinline__ bool is_padded(const TextView &rs, u32 p, int K) {
	if (SPEC) {
		u32 run;
		padded_key(rs, p, K, run);
		return run < (u32)K;
	}
	return p + (u32)K > rs.len || (p <= rs.mid && rs.mid < p + (u32)K);
}

// ---- texts with separators ('!' / ';' of join mode): the padded suffixes.
// Every suffix with a separator (or the text end) inside its first K characters has a padded
// key, and they pile up: all suffixes that START with a separator share key 0, a genome of
// 2000 contigs puts ~5000 suffixes there. They are therefore kept out of the per-bucket
// insertion sort: the bucketing pass appends them to a list with an order-preserving 63-bit
// key (the byte-order ranks of their first 21 characters, 3 bits each), the list is sorted by
// a library radix sort, the rare equal keys are settled by direct comparison, and every padded
// suffix is written to the FRONT of its bucket (the padded key is monotone, so the sorted list
// restricted to a bucket is the bucket's order). The valid suffixes fill the bucket from the
// back; fvalid[key] = index of the first valid one.
struct PaddedList {
	u64 *key;
	u32 *idx;
	u32 *count;	 // appended so far (may exceed cap: the host then grows the list and rebuilds)
	u32 cap;
};

__device__ __forceinline__ u64 order_key21(const TextView &rs, u32 p) {
	u64 cw = window32(rs.code, p), sw = window32(rs.spec, p);
	u32 left = rs.len - p;
	u64 key = 0;
#pragma unroll
	for (u32 c = 0; c < 21; c++) {
		u32 code = (u32)(cw >> (2 * c)) & 3u, sp = (u32)(sw >> (2 * c)) & 1u;
		u32 r = c < left ? (sp ? code + 1u : code + 4u) : 0u;
		key = (key << 3) | r;
	}
	return key;
}

// The padded bucket key of a suffix from its order key (K <= 14 < 21): codes up to the first
// separator / end, zeros from there.
__device__ __forceinline__ u32 bucket_of_order_key(u64 key21, int K) {
	u32 out = 0;
	bool open = true;
	for (int c = 0; c < K; c++) {
		u32 r = (u32)(key21 >> (3 * (20 - c))) & 7u;
		open = open && r >= 4u;
		out = (out << 2) | (open ? r - 4u : 0u);
	}
	return out;
}

__device__ __forceinline__ void padded_append(const PaddedList &pl, const TextView &rs, u32 p) {
	u32 slot = atomicAdd(pl.count, 1u);
	if (slot < pl.cap) pl.key[slot] = order_key21(rs, p), pl.idx[slot] = p;
}

// Scatter of the counting-sort path: valid suffixes fill their bucket from the back
// (cursor_end starts at the bucket end and ends at fvalid), padded ones go to the list.
__global__ void k_bucket_scatter_spec(TextView rs, int K, u32 *__restrict__ cursor_end, u32 *__restrict__ SA,
									  PaddedList pl) {
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= rs.len) return;
	u32 run;
	u32 key = padded_key(rs, i, K, run);
	if (run < (u32)K)
		padded_append(pl, rs, i);
	else
		SA[atomicSub(cursor_end + key, 1u) - 1u] = i;
}

// Radix-sort path: sort key = bucket key * 2 + valid, so the padded suffixes come first.
__global__ void k_bucket_keys_spec(TextView rs, int K, u32 *__restrict__ keys, u32 *__restrict__ idx, PaddedList pl) {
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= rs.len) return;
	u32 run;
	u32 key = padded_key(rs, i, K, run);
	bool valid = run >= (u32)K;
	if (!valid) padded_append(pl, rs, i);
	keys[i] = (key << 1) | (valid ? 1u : 0u);
	idx[i] = i;
}

__global__ void k_bucket_bounds_spec(const u32 *__restrict__ keys, u32 N, u32 *__restrict__ bstart,
									 u32 *__restrict__ bend, u32 *__restrict__ fvalid) {
	u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= N) return;
	u32 kv = keys[j], k = kv >> 1;
	bool first_of_run = j == 0 || keys[j - 1] != kv, last_of_run = j + 1 == N || keys[j + 1] != kv;
	if (j == 0 || (keys[j - 1] >> 1) != k) bstart[k] = j;
	if (j + 1 == N || (keys[j + 1] >> 1) != k) bend[k] = j + 1;
	if ((kv & 1u) && first_of_run) fvalid[k] = j;
	if (!(kv & 1u) && last_of_run) fvalid[k] = j + 1;
}

// Equal order keys (two padded suffixes agreeing on 21 characters: duplicated contig starts):
// the thread at the head of such a run orders it by direct comparison.
template <bool SPEC>
__global__ void k_padded_ties(TextView rs, const u64 *__restrict__ key, u32 *__restrict__ idx,
							  const u32 *__restrict__ count, u32 cap) {
	u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	u32 P = min(*count, cap);
	if (j >= P) return;
	u64 k = key[j];
	if (j > 0 && key[j - 1] == k) return;
	u32 e = j + 1;
	while (e < P && key[e] == k) e++;
	for (u32 x = j + 1; x < e; x++) {
		u32 cur = idx[x];
		u32 y = x;
		while (y > j && compare_suffixes<SPEC>(rs, idx[y - 1], cur, 0xffffffffu) > 0) {
			idx[y] = idx[y - 1];
			y--;
		}
		idx[y] = cur;
	}
}

// Sorted padded suffix j goes to slot (j - first list index of its bucket) of its bucket.
__global__ void k_padded_place(int K, const u64 *__restrict__ key, const u32 *__restrict__ idx,
							   const u32 *__restrict__ count, u32 cap, const u32 *__restrict__ bstart,
							   u32 *__restrict__ SA) {
	u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	u32 P = min(*count, cap);
	if (j >= P) return;
	u32 bk = bucket_of_order_key(key[j], K);
	u32 lo = 0, hi = j;
	while (lo < hi) {
		u32 m = lo + ((hi - lo) >> 1);
		if (bucket_of_order_key(key[m], K) < bk)
			lo = m + 1;
		else
			hi = m;
	}
	SA[bstart[bk] + (j - lo)] = idx[j];
}

// dir64[key] = first SA index of the suffixes that really start with this k-mer (no
// separator inside) | their number << 32; first + number = end of the bucket. *n_ambiguous counts the suffixes that are still tied
// with a neighbour; their groups are materialised by k_bucket_groups only when there are any.
template <bool SPEC>
__device__ __forceinline__ u32 bucket_sort_range(const TextView &rs, int K, u32 key, u32 b, u32 e, bool presorted_front,
												 u32 *__restrict__ SA, u64 *__restrict__ dir64, const TieSink &sink);

// One bucket, found through the tables of the bucketing pass: [bstart[key], bend[key]).
template <bool SPEC>
__device__ __forceinline__ u32 bucket_sort_key(const TextView &rs, int K, u32 key, const u32 *__restrict__ bstart,
											   const u32 *__restrict__ bend, const u32 *__restrict__ fvalid,
											   u32 *__restrict__ SA, u64 *__restrict__ dir64,
											   const TieSink &sink, u32 empty_known) {
	const u32 e = bend[key];
	if (e == bstart[key]) {
		// first + count is the END of the bucket for every key, so the start of bucket k is the
		// end of k-1 (the generic search narrows its range with that). The radix bucketing path
		// does not know where an empty bucket lies: 0xffffffff = unknown.
		dir64[key] = empty_known ? (u64)e : 0xffffffffULL;
		return 0;
	}
	// fvalid given: the padded suffixes already stand sorted in [bstart, fvalid) (k_padded_place)
	return bucket_sort_range<SPEC>(rs, K, key, fvalid ? min(fvalid[key], e) : bstart[key], e, fvalid != nullptr, SA, dir64, sink);
}

// Order the suffixes SA[b, e) of one bucket (b < e unless presorted_front) and emit its directory entry.
template <bool SPEC>
__device__ __forceinline__ u32 bucket_sort_range(const TextView &rs, int K, u32 key, u32 b, u32 e, bool presorted_front,
												 u32 *__restrict__ SA, u64 *__restrict__ dir64, const TieSink &sink) {
	u32 *n_ambiguous = sink.flags;
	const bool fvalid = presorted_front;
	const u32 s = e - b;
	if (fvalid && s > ANDI_SORT_MAX) {	// valid suffixes only: one group of depth K for the doubling rounds
		atomicAdd(n_ambiguous, s);
		dir64[key] = (u64)b | ((u64)s << 32);
		return s;
	}
	u32 valid = 0;
	if (s > ANDI_SORT_MAX) {
		// Too large for one thread: the suffixes that really start with this k-mer are left to
		// the doubling rounds as one group of depth K. The padded ones (a separator or the text
		// end inside their first K characters) do NOT share K characters with anything, so they
		// are moved to the front and ordered here by direct comparison (they are few).
		u32 front = b;
		for (u32 j = b; j < e; j++) {
			u32 p = SA[j];
			if (is_padded<SPEC>(rs, p, K)) {
				SA[j] = SA[front];
				SA[front] = p;
				front++;
			}
		}
		valid = e - front;
		for (u32 x = b + 1; x < front; x++) {
			u32 cur = SA[x];
			u32 y = x;
			while (y > b && compare_suffixes<SPEC>(rs, SA[y - 1], cur, 0xffffffffu) > 0) {
				SA[y] = SA[y - 1];
				y--;
			}
			SA[y] = cur;
		}
		// a few dozen suffixes: k_sort_deep orders the whole bucket (should it give up, the bucket is
		// still in the form the doubling rounds expect)
		if (valid > 1 && !(!SPEC && deep_append(sink, b, e))) atomicAdd(n_ambiguous, valid);
		dir64[key] = (u64)front | ((u64)valid << 32);
		return valid;
	}
	if (s <= 2) {
		// nine buckets in ten: one suffix (nothing to order) or two (one comparison decides order
		// and tie) -- kept out of the local-memory array of the general case
		if (s == 0) {  // only padded suffixes (already placed)
			dir64[key] = (u64)e;
			return 0;
		}
		u32 p0 = SA[b], p1 = s == 2 ? SA[b + 1] : 0u;
		valid = fvalid ? s : (u32)!is_padded<SPEC>(rs, p0, K);
		if (s == 2) {
			if (!fvalid) valid += (u32)!is_padded<SPEC>(rs, p1, K);
			int c = compare_suffixes<SPEC>(rs, p0, p1, ANDI_SORT_CAP);
			if (c > 0) SA[b] = p1, SA[b + 1] = p0;
			if (c == 0 && !(!SPEC && deep_append(sink, b, e))) atomicAdd(n_ambiguous, 2u);
		}
		dir64[key] = (u64)(e - valid) | ((u64)valid << 32);
		return valid;
	}
	u32 v[ANDI_SORT_MAX];
#pragma unroll
	for (int x = 0; x < ANDI_SORT_MAX; x++) v[x] = x < (int)s ? SA[b + x] : 0u;
	// insertion sort; ties (equal up to the cap) keep their relative order
	for (u32 x = 1; x < s; x++) {
		u32 cur = v[x];
		u32 y = x;
		while (y > 0 && compare_suffixes<SPEC>(rs, v[y - 1], cur, ANDI_SORT_CAP) > 0) {
			v[y] = v[y - 1];
			y--;
		}
		v[y] = cur;
	}
	u32 tied = 0;
	for (u32 x = 0; x < s; x++) {
		u32 p = v[x];
		valid += fvalid ? 1u : (u32)!is_padded<SPEC>(rs, p, K);
		if (s > 1) SA[b + x] = p;
		if (x > 0 && compare_suffixes<SPEC>(rs, v[x - 1], p, ANDI_SORT_CAP) == 0) tied++;
	}
	if (tied && !(!SPEC && deep_append(sink, b, e))) atomicAdd(n_ambiguous, tied + 1);
	dir64[key] = (u64)(e - valid) | ((u64)valid << 32);
	return valid;
}

// One thread per k-mer. Besides the directory entry the kernel leaves level K-1 of the
// presence bitmaps (esa_kernels.cuh): bit y is set iff one of the four k-mers 4y..4y+3 occurs,
// folded from a warp ballot (the words must be zero before the launch).
template <bool SPEC>
__global__ void k_bucket_sort(TextView rs, int K, const u32 *__restrict__ bstart, const u32 *__restrict__ bend,
							  const u32 *__restrict__ fvalid, u32 *__restrict__ SA, u64 *__restrict__ dir64,
							  const TieSink sink, u32 empty_known, u32 *__restrict__ present_top) {
	u32 key = blockIdx.x * blockDim.x + threadIdx.x;
	u32 count = 0;
	if (key < (1u << (2 * K))) count = bucket_sort_key<SPEC>(rs, K, key, bstart, bend, fvalid, SA, dir64, sink, empty_known);
	u32 m = __ballot_sync(0xffffffffu, count != 0);
	if ((threadIdx.x & 31u) == 0 && m) {
		m |= m >> 1, m |= m >> 2;  // bit 4g = any of the four k-mers of group g
		u32 folded = 0;
#pragma unroll
		for (int g = 0; g < 8; g++) folded |= ((m >> (4 * g)) & 1u) << g;
		u32 y0 = key >> 2;	// first (K-1)-mer of this warp (key is a multiple of 32 here)
		atomicOr(present_top + (y0 >> 5), folded << (y0 & 31u));
	}
}



// ---- k_bucket_sort_slots (texts without separators): the same per-bucket work with one thread per
// suffix-array SLOT instead of one per k-mer. At depth K = log4(N) + 1 three buckets in four are
// empty, and a thread per k-mer spends most of the kernel reading table entries of empty buckets
// (18 % of the whole build in round 1). Here thread j computes the k-mer of the suffix the bucketing
// pass left in slot j and reads that ONE bucket's bounds (the scatter cursor table = bucket ends):
// if the bucket starts at j, j is its head and sorts it. The entries of EMPTY buckets are written
// by the scan kernel below. Presence bits (level K-1) by atomicOr.
__device__ __forceinline__ u32 padded_key_nosep(const TextView &rs, u32 p, int K) {
	const u32 run = p <= rs.mid ? rs.mid - p : rs.len - p;	// nucleotides before '#' / the end
	u64 cw = window32(rs.code, p);
	if (run < 32u) cw &= (1ULL << (2u * run)) - 1ULL;
	return kmer_key(cw, K);
}

__global__ void k_bucket_sort_slots(TextView rs, int K, const u32 *__restrict__ bend, u32 *__restrict__ SA, u64 *__restrict__ dir64,
									const TieSink sink, u32 *__restrict__ present_top) {
	const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= rs.len) return;
	// the bucket of the suffix in this slot: [bend[key - 1], bend[key]) (bend[-1] is a zero in front of the table)
	const u32 key = padded_key_nosep(rs, SA[j], K);
	if (bend[(int)key - 1] != j) return;  // not the head of its bucket
	const u32 count = bucket_sort_range<false>(rs, K, key, j, bend[key], false, SA, dir64, sink);
	if (count) {
		const u32 y = key >> 2;	 // its (K-1)-mer
		atomicOr(present_top + (y >> 5), 1u << (y & 31u));
	}
}

// ---- k_scan_buckets: exclusive prefix sums over the k-mer histogram in one launch (stands where
// round 1 called cub::DeviceScan). The grid is small enough to be co-resident (two CTAs per SM),
// every CTA owns one contiguous stretch of keys: (1) it sums its stretch and publishes the sum,
// (2) its threads wait for the sums of all CTAs before it and add them up -- everybody is resident,
// so nobody waits for a CTA that cannot run --, (3) it scans its stretch tile by tile.
// out[k] = number of suffixes in buckets < k (may alias hist: the histogram becomes the scatter
// cursor), out[n] = N. With dir64 given, the directory is initialised on the way: the entries of
// EMPTY buckets are final (first = end of the bucket, count 0) -- k_bucket_sort_slots never sees
// those buckets --, the others are rewritten by the bucket sort.
#define ANDI_SCAN_TILE 4096u  // 256 threads x 16 keys
__device__ __forceinline__ u32 block_sum_256(u32 v, u32 *warp_sum) {
	for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	__syncthreads();  // warp_sum may still be read from the previous use
	if ((threadIdx.x & 31u) == 0) warp_sum[threadIdx.x >> 5] = v;
	__syncthreads();
	u32 t = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) t += warp_sum[k];
	return t;
}

__global__ void __launch_bounds__(256) k_scan_buckets(const u32 *hist, u32 n, u32 per_cta, u32 *out, u64 *__restrict__ dir64,
													  unsigned long long *cta_sum) {
	__shared__ u32 warp_sum[8];
	const u32 b0 = min(n, blockIdx.x * per_cta), b1 = min(n, b0 + per_cta);	 // per_cta is a multiple of the tile
	const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	// (1) the sum of the stretch
	u32 mine = 0;
	for (u32 k = b0 + threadIdx.x * 4u; k < b1; k += 1024u) {
		if (k + 3u < b1) {
			const uint4 x = *reinterpret_cast<const uint4 *>(hist + k);
			mine += x.x + x.y + x.z + x.w;
		} else {
			for (u32 j = k; j < b1; j++) mine += hist[j];
		}
	}
	const u32 total = block_sum_256(mine, warp_sum);
	volatile unsigned long long *sums = cta_sum;
	if (threadIdx.x == 0) {
		__threadfence();
		sums[blockIdx.x] = (1ULL << 32) | total;
	}
	// (2) everything in front of this stretch
	u32 before = 0;
	for (u32 c = threadIdx.x; c < blockIdx.x; c += 256u) {
		unsigned long long x;
		do x = sums[c];
		while (!(x >> 32));
		before += (u32)x;
	}
	u32 run0 = block_sum_256(before, warp_sum);
	// (3) scan the stretch, 1024 keys at a time: thread t owns keys 4t .. 4t+3 of the tile, so that
	// loads and stores are 16 contiguous bytes per lane (a thread that owns 16 consecutive keys makes
	// every load and store touch 32 different sectors: measured 245 us instead of 35 for 4^12 keys)
	for (u32 t0 = b0; t0 < b1; t0 += 1024u) {
		const u32 base = t0 + threadIdx.x * 4u;
		uint4 x = make_uint4(0, 0, 0, 0);
		if (base + 3u < b1)
			x = *reinterpret_cast<const uint4 *>(hist + base);
		else {
			if (base + 0u < b1) x.x = hist[base + 0u];
			if (base + 1u < b1) x.y = hist[base + 1u];
			if (base + 2u < b1) x.z = hist[base + 2u];
		}
		const u32 part = x.x + x.y + x.z + x.w;
		u32 inc = part;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			u32 y = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= (u32)d) inc += y;
		}
		__syncthreads();
		if (lane == 31u) warp_sum[wid] = inc;
		__syncthreads();
		u32 in_front = 0, tile_total = 0;
#pragma unroll
		for (u32 k = 0; k < 8; k++) {
			const u32 ws = warp_sum[k];
			in_front += k < wid ? ws : 0u;
			tile_total += ws;
		}
		const u32 r0 = run0 + in_front + inc - part, r1 = r0 + x.x, r2 = r1 + x.y, r3 = r2 + x.z;
		if (base + 3u < b1) {
			*reinterpret_cast<uint4 *>(out + base) = make_uint4(r0, r1, r2, r3);
		} else {
			if (base + 0u < b1) out[base + 0u] = r0;
			if (base + 1u < b1) out[base + 1u] = r1;
			if (base + 2u < b1) out[base + 2u] = r2;
		}
		if (dir64) {
			// every entry, as full 32-byte stores (predicated 8-byte stores to the empty ones alone made
			// partial sectors: 190 MB read + 160 MB written for a 134 MB table). An empty bucket's
			// entry is final (first = its end, count 0); a non-empty one gets first | count here and
			// is rewritten by the bucket sort with the count of suffixes that really carry the k-mer.
			if (base + 3u < b1) {
				ulonglong2 *d = reinterpret_cast<ulonglong2 *>(dir64 + base);
				d[0] = make_ulonglong2((u64)r0 | ((u64)x.x << 32), (u64)r1 | ((u64)x.y << 32));
				d[1] = make_ulonglong2((u64)r2 | ((u64)x.z << 32), (u64)r3 | ((u64)x.w << 32));
			} else {
				if (base + 0u < b1) dir64[base + 0u] = (u64)r0 | ((u64)x.x << 32);
				if (base + 1u < b1) dir64[base + 1u] = (u64)r1 | ((u64)x.y << 32);
				if (base + 2u < b1) dir64[base + 2u] = (u64)r2 | ((u64)x.z << 32);
			}
		}
		run0 += tile_total;
	}
	if (blockIdx.x == gridDim.x - 1u && threadIdx.x == 0) out[n] = run0;	// = N (the last CTA ends at n)
}

// ---- two-level counting sort for DEEP directories (K >= 13: texts of hundreds of Mbp, config 5).
// A histogram of 4^K counters (1 GB at K = 14) is far beyond L2, and random atomics on it run at
// DRAM-sector speed; round 1 went through cub::DeviceRadixSort there. Own replacement:
//   level 1  partition all suffixes by their first K1 = ceil(K/2) bases (<= 16384 parts): every CTA
//            histograms its stretch of the text in shared memory, reserves its share of every part
//            with ONE global atomic per non-empty part, and scatters through shared-memory cursors
//            (k_part_hist, k_scan_buckets, k_part_scatter) -> positions grouped by part in `tmp`
//   level 2  one CTA per part: counting sort of its suffixes on the remaining K2 = K - K1 bases with
//            a 4^K2-counter histogram in shared memory; it knows every bucket of its part, so it
//            writes the bucket ends, and the directory entries of the empty buckets, as coalesced
//            streams, and then sorts its buckets in place (k_part_sort) -> SA ordered up to ties
// Texts without separators only. Traffic: the text twice (L2), 3 x 4N bytes of positions.
__global__ void __launch_bounds__(1024) k_part_hist(TextView rs, int K, int K2, u32 per_cta, u32 *__restrict__ hist1) {
	extern __shared__ u32 sh[];
	const u32 parts = 1u << (2 * (K - K2));
	for (u32 x = threadIdx.x; x < parts; x += blockDim.x) sh[x] = 0;
	__syncthreads();
	const u32 c0 = blockIdx.x * per_cta, c1 = min(rs.len, c0 + per_cta);
	for (u32 i = c0 + threadIdx.x; i < c1; i += blockDim.x) atomicAdd(&sh[padded_key_nosep(rs, i, K) >> (2 * K2)], 1u);
	__syncthreads();
	for (u32 x = threadIdx.x; x < parts; x += blockDim.x)
		if (sh[x]) atomicAdd(hist1 + x, sh[x]);
}

__global__ void __launch_bounds__(1024) k_part_scatter(TextView rs, int K, int K2, u32 per_cta, u32 *__restrict__ cursor1,
														u32 *__restrict__ tmp) {
	extern __shared__ u32 sh[];
	const u32 parts = 1u << (2 * (K - K2));
	for (u32 x = threadIdx.x; x < parts; x += blockDim.x) sh[x] = 0;
	__syncthreads();
	const u32 c0 = blockIdx.x * per_cta, c1 = min(rs.len, c0 + per_cta);
	for (u32 i = c0 + threadIdx.x; i < c1; i += blockDim.x) atomicAdd(&sh[padded_key_nosep(rs, i, K) >> (2 * K2)], 1u);
	__syncthreads();
	for (u32 x = threadIdx.x; x < parts; x += blockDim.x)
		if (sh[x]) sh[x] = atomicAdd(cursor1 + x, sh[x]);  // this CTA's slots of part x start here
	__syncthreads();
	for (u32 i = c0 + threadIdx.x; i < c1; i += blockDim.x) tmp[atomicAdd(&sh[padded_key_nosep(rs, i, K) >> (2 * K2)], 1u)] = i;
}

// start1[x] = first slot of part x (x = blockIdx.x), start1[x + 1] its end. bend[key] = end of bucket key.
__global__ void __launch_bounds__(1024) k_part_sort(TextView rs, int K, int K2, const u32 *__restrict__ start1,
													 const u32 *__restrict__ tmp, u32 *__restrict__ SA, u32 *__restrict__ bend,
													 u64 *__restrict__ dir64, const TieSink sink,
													 u32 *__restrict__ present_top) {
	extern __shared__ u32 sh[];	 // 4^K2 counters, then 32 warp sums
	const u32 bins = 1u << (2 * K2), mask = bins - 1u, part = blockIdx.x;
	const u32 s = start1[part], e = start1[part + 1];
	u32 *warp_sum = sh + bins;
	for (u32 x = threadIdx.x; x < bins; x += blockDim.x) sh[x] = 0;
	__syncthreads();
	for (u32 j = s + threadIdx.x; j < e; j += blockDim.x) atomicAdd(&sh[padded_key_nosep(rs, tmp[j], K) & mask], 1u);
	__syncthreads();
	// exclusive scan of the counters: every thread owns bins / 1024 consecutive ones (4 or 16)
	const u32 per = bins / blockDim.x, b0 = threadIdx.x * per, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	u32 mine = 0;
	for (u32 x = 0; x < per; x++) mine += sh[b0 + x];
	u32 inc = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		u32 y = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (u32)d) inc += y;
	}
	if (lane == 31u) warp_sum[wid] = inc;
	__syncthreads();
	u32 before = 0;
	for (u32 k = 0; k < wid; k++) before += warp_sum[k];
	u32 run = s + before + inc - mine;
	const size_t key0 = ((size_t)part << (2 * K2)) + b0;
	for (u32 x = 0; x < per; x++) {
		const u32 cnt = sh[b0 + x];
		if (cnt == 0) dir64[key0 + x] = (u64)run;  // empty bucket: first = its end, count 0
		sh[b0 + x] = run;						   // cursor
		run += cnt;
		bend[key0 + x] = run;
	}
	__syncthreads();
	for (u32 j = s + threadIdx.x; j < e; j += blockDim.x) {
		const u32 i = tmp[j];
		SA[atomicAdd(&sh[padded_key_nosep(rs, i, K) & mask], 1u)] = i;
	}
	__syncthreads();
	// the buckets of this part are complete and this CTA knows where each one lies (the cursors now
	// hold the bucket ENDS): every thread sorts the buckets of its own bins right here -- no
	// separate pass over all slots that finds the bucket heads again through the text
	u32 start = b0 ? sh[b0 - 1] : s;
	for (u32 x = 0; x < per; x++) {
		const u32 end = sh[b0 + x];
		if (end > start) {
			const u32 key = (u32)(key0 + x);
			if (bucket_sort_range<false>(rs, K, key, start, end, false, SA, dir64, sink))
				atomicOr(present_top + ((key >> 2) >> 5), 1u << ((key >> 2) & 31u));
		}
		start = end;
	}
}

// The walk's view of the directory, one 8-byte entry per k-mer so that the common lookups cost
// ONE table access (the tables compete with the streaming queries for L2):
//   tag 0 (absent k-mer)    low bits = plen: the longest prefix of the k-mer present in RS, which
//                           is all the walk needs (no anchor can result below K)
//   tag 1 (one suffix)      bits 0..30 = its text position, bits 31..60 = the 15 bases that follow
//                           the k-mer there: k_walk_v3 reads the match length up to K + 15 off the
//                           entry (no candidate window at all), k_walk_chunks_fast starts its compare
//                           without the dependent suffix-array load
//   tag 2 (two suffixes)    bits 0..30 and 31..61 = their text positions in suffix order (N < 2^31):
//                           most buckets with company hold exactly two, no suffix-array load either
//   tag 3 (three or more)   low 32 bits = first SA index, bits 32..61 = their number
// Written by k_prefix_len (esa_kernels.cuh) together with the prefix lengths.
#define ANDI_FDIR_TAG(e) ((u32)((e) >> 62))

// One warp per listed bucket [b, e) (at most ANDI_DEEP_BUCKET suffixes): every suffix is ranked by
// counting the bucket members below it, compares running to the end of the match (cap: ANDI_DEEP_CAP
// characters; a compare that gets there raises flags[0] and everybody stops: the doubling rounds
// take over). Texts without separators.
__global__ void __launch_bounds__(256) k_sort_deep(TextView rs, u32 *__restrict__ SA, const TieSink sink) {
	const u32 lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const u32 n = min(*(volatile u32 *)(sink.flags + 4), sink.deep_cap);
	for (u32 x = warp; x < n; x += nwarps) {
		if (*(volatile u32 *)sink.flags) return;  // somebody met a match beyond the cap
		const u32 b = sink.deep[2 * x], m = sink.deep[2 * x + 1] - b;
		u32 p[ANDI_DEEP_BUCKET / 32], r[ANDI_DEEP_BUCKET / 32];
		bool capped = false;
#pragma unroll
		for (u32 k = 0; k < ANDI_DEEP_BUCKET / 32; k++) {
			const u32 i = lane + 32u * k;
			p[k] = i < m ? SA[b + i] : 0u, r[k] = 0;
		}
		for (u32 j = 0; j < m; j++) {
			const u32 pj = SA[b + j];  // the same word for all lanes
#pragma unroll
			for (u32 k = 0; k < ANDI_DEEP_BUCKET / 32; k++) {
				const u32 i = lane + 32u * k;
				if (i < m && i != j) {
					const int c = compare_suffixes<false>(rs, pj, p[k], ANDI_DEEP_CAP);
					capped |= c == 0;
					r[k] += c < 0 ? 1u : 0u;
				}
			}
		}
		if (__any_sync(0xffffffffu, capped)) {
			if (lane == 0) atomicAdd(sink.flags, m);
			return;	 // (the bucket stays as it was: a permutation of its suffixes)
		}
		__syncwarp();  // all reads of SA[b, e) are done
#pragma unroll
		for (u32 k = 0; k < ANDI_DEEP_BUCKET / 32; k++)
			if (lane + 32u * k < m) SA[b + r[k]] = p[k];
	}
}

// Only when k_bucket_sort reported ties: group heads, ranks and "ambiguous" flags of every
// suffix, the input of the doubling rounds (index_host.cuh). Buckets are laid out as
// k_bucket_sort left them.
template <bool SPEC>
__global__ void k_bucket_groups(TextView rs, int K, const u32 *__restrict__ bstart, const u32 *__restrict__ bend,
								const u32 *__restrict__ fvalid, const u32 *__restrict__ SA, u32 *__restrict__ grp,
								u32 *__restrict__ rank, unsigned char *__restrict__ amb) {
	u32 key = blockIdx.x * blockDim.x + threadIdx.x;
	if (key >= (1u << (2 * K))) return;
	u32 b = bstart[key];
	const u32 e = bend[key];
	if (e == b) return;
	if (fvalid) {  // sorted padded suffixes in front: singleton groups
		u32 f = min(fvalid[key], e);
		for (u32 j = b; j < f; j++) grp[j] = j, rank[SA[j]] = j, amb[j] = 0;
		b = f;
		if (e == b) return;
	}
	const u32 s = e - b;
	if (s > ANDI_SORT_MAX) {
		u32 front = b;
		for (; front < e && !fvalid; front++) {
			if (!is_padded<SPEC>(rs, SA[front], K)) break;
		}
		for (u32 j = b; j < front; j++) grp[j] = j, rank[SA[j]] = j, amb[j] = 0;
		for (u32 j = front; j < e; j++) grp[j] = front, rank[SA[j]] = front, amb[j] = (e - front) > 1;
		return;
	}
	u32 ties = 0;  // bit x: SA[b+x] agrees with SA[b+x-1] on the first ANDI_SORT_CAP characters
	for (u32 x = 1; x < s; x++)
		if (compare_suffixes<SPEC>(rs, SA[b + x - 1], SA[b + x], ANDI_SORT_CAP) == 0) ties |= 1u << x;
	u32 head = b;
	for (u32 x = 0; x < s; x++) {
		bool same = (ties >> x) & 1u;
		if (!same) head = b + x;
		grp[b + x] = head;
		rank[SA[b + x]] = head;
		amb[b + x] = same || ((ties >> (x + 1)) & 1u);
	}
}

// LCP[j] = lcp(SA[j-1], SA[j]) by direct comparison, at most `cap` characters; pairs that reach
// the cap raise *overflow and the caller recomputes everything through the phi array
// (src/esa.c:373-426, k_phi / k_plcp) -- only repeat-rich texts get there.
template <bool SPEC>
__global__ void k_lcp_direct(TextView rs, const u32 *__restrict__ SA, u32 cap, int32_t *__restrict__ LCP, const TieSink sink) {
	u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j > rs.len) return;
	if (j == 0 || j == rs.len) {
		LCP[j] = -1;
		return;
	}
	u32 a = SA[j - 1], b = SA[j];
	u32 lim = min(cap, SPEC ? rs.len - max(a, b) : pair_limit_fast(rs, a, b));
	u32 m = match_len<SPEC>(rs, a, rs, b, lim);
	if (m == cap) {
		// the value goes on from here in k_lcp_deep if the list has room (texts without separators)
		const u32 at = (!SPEC && sink.deep) ? atomicAdd(sink.flags + 5, 1u) : 0xffffffffu;
		if (at < sink.lcp_cap)
			sink.deep[2 * sink.deep_cap + at] = j;
		else
			atomicExch(sink.flags + 1, 1u);
	}
	LCP[j] = (int32_t)m;
}

// The listed LCP values (they reached the cap of k_lcp_direct: neighbours inside a repeat), one warp
// each: 32 windows of 64 characters per round, on to the first mismatch or the limit.
__global__ void __launch_bounds__(256) k_lcp_deep(TextView rs, const u32 *__restrict__ SA, u32 cap, int32_t *__restrict__ LCP,
												   const TieSink sink) {
	const u32 lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	if (*(volatile u32 *)(sink.flags + 1)) return;	// the list overflowed: everything is redone through phi
	const u32 n = min(*(volatile u32 *)(sink.flags + 5), sink.lcp_cap);
	for (u32 x = warp; x < n; x += nwarps) {
		const u32 j = sink.deep[2 * sink.deep_cap + x];
		const u32 a = SA[j - 1], b = SA[j], lim = pair_limit_fast(rs, a, b);
		u32 m = cap;
		while (m < lim) {
			const u32 at = m + 64u * lane;
			u32 d = 64u;
			if (at < lim) {
				u64 a0, a1, b0, b1;
				window64(rs.code, a + at, a0, a1);
				window64(rs.code, b + at, b0, b1);
				const u64 x0 = a0 ^ b0, x1 = a1 ^ b1;
				d = x0 ? (u32)(__ffsll((long long)x0) - 1) >> 1 : (x1 ? 32u + ((u32)(__ffsll((long long)x1) - 1) >> 1) : 64u);
			}
			const unsigned hit = __ballot_sync(0xffffffffu, d < 64u);
			if (hit) {
				const int first = __ffs((int)hit) - 1;
				m += 64u * (u32)first + __shfl_sync(0xffffffffu, d, first);
				break;
			}
			m += 2048u;
		}
		if (lane == 0) LCP[j] = (int32_t)min(m, lim);
	}
}
