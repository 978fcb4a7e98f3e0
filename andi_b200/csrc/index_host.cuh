// andi_b200/csrc/index_host.cuh -- host side of index construction (esa_init, src/esa.c:254-277);
// included by andi_b200.cu.
//
// Optimistic pipeline, ONE host synchronisation per subject in the common case:
//   bucket suffix sort (sa_bucket.cuh) -> direct LCP -> k-mer directory / presence / prefix
//   lengths -> read back two flags {tied suffixes, LCP overflow}.
// Buckets the capped per-thread sort cannot finish (repeats longer than ANDI_SORT_CAP, 17 to 64
// suffixes) and LCP values that reach the direct cap are LISTED and finished by one warp each
// (k_sort_deep, k_lcp_deep: the repeats of real genomes). Only when suffixes are still tied after
// that (full list, larger buckets, matches beyond ANDI_DEEP_CAP) do the prefix-doubling rounds run
// (group, rank[i+h]) with h = K, 2K, ...; only when the LCP list overflows (join mode: when any
// value reached the cap) is the LCP recomputed through the phi array (src/esa.c:373-426).
// The radix sort / scan / select used by the doubling rounds are CUB device primitives.
#pragma once
#include "sa_bucket.cuh"

#include <thrust/iterator/counting_iterator.h>

#define ANDI_LCP_DIRECT_CAP 1024u
#define ANDI_DEEP_LIST 65536u		// buckets k_sort_deep can take (sa_bucket.cuh, TieSink)
#define ANDI_DEEP_LCP_LIST 524288u	// LCP values k_lcp_deep can take

struct MaxOp {
	__host__ __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; }
};

// Refine the suffixes flagged in amb[0..N) (grp / rank hold their current groups, all of depth
// >= h0) until every group is a singleton.
static int doubling_rounds(andi_ctx *ctx, andi_esa *E, u32 *grp, u32 *rank, unsigned char *amb, u32 h0) {
	const u32 N = E->N;
	cudaStream_t st = ctx->stream;
	u32 *pos_a = nullptr, *pos_b = nullptr, *d_count = nullptr;
	CK(dalloc(ctx, &pos_a, N));
	CK(dalloc(ctx, &d_count, 1));
	size_t sel_bytes = 0;
	thrust::counting_iterator<u32> iota(0);
	cub::DeviceSelect::Flagged(nullptr, sel_bytes, iota, amb, pos_a, d_count, (int)N, st);
	void *tmp = nullptr;
	CK(cudaMallocAsync(&tmp, sel_bytes, st));
	CK(cub::DeviceSelect::Flagged(tmp, sel_bytes, iota, amb, pos_a, d_count, (int)N, st));
	ctx->st.cub_calls++;
	u32 m = 0;
	CK(cudaMemcpyAsync(&m, d_count, sizeof(u32), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	cudaFreeAsync(tmp, st);
	if (m == 0) {
		dfree(ctx, pos_a), dfree(ctx, d_count);
		return ANDI_OK;
	}
	u64 *keys_a = nullptr, *keys_b = nullptr;
	u32 *vals_a = nullptr, *vals_b = nullptr, *v = nullptr, *g = nullptr;
	unsigned char *amb2 = nullptr;
	CK(dalloc(ctx, &keys_a, m));
	CK(dalloc(ctx, &keys_b, m));
	CK(dalloc(ctx, &vals_a, m));
	CK(dalloc(ctx, &vals_b, m));
	CK(dalloc(ctx, &v, m));
	CK(dalloc(ctx, &g, m));
	CK(dalloc(ctx, &amb2, m));
	CK(dalloc(ctx, &pos_b, m));
	size_t b1 = 0, b2 = 0, b3 = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, b1, keys_a, keys_b, vals_a, vals_b, (int)m, 0, 64, st);
	cub::DeviceScan::InclusiveScan(nullptr, b2, v, g, MaxOp(), (int)m, st);
	cub::DeviceSelect::Flagged(nullptr, b3, pos_a, amb2, pos_b, d_count, (int)m, st);
	size_t tmp_bytes = std::max(b1, std::max(b2, b3));
	CK(cudaMallocAsync(&tmp, tmp_bytes, st));
	int bits = 33;
	while (bits < 64 && (1ULL << (bits - 32)) <= (u64)N) bits++;  // the group index needs log2(N) bits
	for (u32 h = h0; m > 0; h *= 2) {
		size_t tb = tmp_bytes;
		k_round_keys<<<nblocks(m, 256), 256, 0, st>>>(pos_a, m, E->SA, grp, rank, h, N, keys_a, vals_a);
		CK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys_a, keys_b, vals_a, vals_b, (int)m, 0, bits, st));
		k_head_values<<<nblocks(m, 256), 256, 0, st>>>(keys_b, m, pos_a, v);
		tb = tmp_bytes;
		CK(cub::DeviceScan::InclusiveScan(tmp, tb, v, g, MaxOp(), (int)m, st));
		k_apply_groups<<<nblocks(m, 256), 256, 0, st>>>(g, m, pos_a, vals_b, E->SA, grp, rank, amb2);
		tb = tmp_bytes;
		CK(cub::DeviceSelect::Flagged(tmp, tb, pos_a, amb2, pos_b, d_count, (int)m, st));
		ctx->st.esa_launches += 3;
		ctx->st.cub_calls += 3;
		ctx->st.sa_rounds++;
		CK(cudaMemcpyAsync(&m, d_count, sizeof(u32), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		std::swap(pos_a, pos_b);
		if (h > N) break;
	}
	dfree(ctx, keys_a), dfree(ctx, keys_b), dfree(ctx, vals_a), dfree(ctx, vals_b);
	dfree(ctx, v), dfree(ctx, g), dfree(ctx, amb2), dfree(ctx, pos_a), dfree(ctx, pos_b), dfree(ctx, d_count);
	cudaFreeAsync(tmp, st);
	return ANDI_OK;
}

// Fallback sorter for indexes without a directory (K == 0: tiny thresholds): LSD radix sort
// on the first 16 characters (48-bit keys), then doubling from h = 16.
static int build_sa_radix(andi_ctx *ctx, andi_esa *E) {
	const u32 N = E->N;
	cudaStream_t st = ctx->stream;
	u64 *keys_a = nullptr, *keys_b = nullptr;
	u32 *vals_a = nullptr, *v = nullptr, *g = nullptr, *grp = nullptr, *rank = nullptr;
	unsigned char *amb = nullptr;
	CK(dalloc(ctx, &keys_a, N));
	CK(dalloc(ctx, &keys_b, N));
	CK(dalloc(ctx, &vals_a, N));
	CK(dalloc(ctx, &v, N));
	CK(dalloc(ctx, &g, N));
	CK(dalloc(ctx, &grp, N));
	CK(dalloc(ctx, &rank, N));
	CK(dalloc(ctx, &amb, N));
	size_t b1 = 0, b2 = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, b1, keys_a, keys_b, vals_a, E->SA, (int)N, 0, 48, st);
	cub::DeviceScan::InclusiveScan(nullptr, b2, v, g, MaxOp(), (int)N, st);
	size_t tmp_bytes = std::max(b1, b2);
	void *tmp = nullptr;
	CK(cudaMallocAsync(&tmp, tmp_bytes, st));
	TextView rs = rs_view(E);
	k_suffix_keys<<<nblocks(N, 256), 256, 0, st>>>(rs, keys_a, vals_a);
	size_t tb = tmp_bytes;
	CK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys_a, keys_b, vals_a, E->SA, (int)N, 0, 48, st));
	k_head_values<<<nblocks(N, 256), 256, 0, st>>>(keys_b, N, nullptr, v);
	tb = tmp_bytes;
	CK(cub::DeviceScan::InclusiveScan(tmp, tb, v, g, MaxOp(), (int)N, st));
	k_apply_groups<<<nblocks(N, 256), 256, 0, st>>>(g, N, nullptr, E->SA, E->SA, grp, rank, amb);
	ctx->st.esa_launches += 3;
	ctx->st.cub_calls += 2;
	dfree(ctx, keys_a), dfree(ctx, keys_b), dfree(ctx, vals_a), dfree(ctx, v), dfree(ctx, g);
	cudaFreeAsync(tmp, st);
	int rc = doubling_rounds(ctx, E, grp, rank, amb, 16);
	dfree(ctx, grp), dfree(ctx, rank), dfree(ctx, amb);
	return rc;
}

// src/esa.c:373-426 in full: phi array, blocked Kasai, permute.
static int build_lcp_phi(andi_ctx *ctx, andi_esa *E) {
	const u32 N = E->N;
	cudaStream_t st = ctx->stream;
	int32_t *phi = nullptr;
	CK(dalloc(ctx, &phi, N));
	TextView rs = rs_view(E);
	k_phi<<<nblocks(N, 256), 256, 0, st>>>(E->SA, N, phi);
	u32 slices = (N + 31) / 32;
	if (E->has_sep)
		k_plcp<true><<<nblocks(slices, 128), 128, 0, st>>>(rs, phi);
	else
		k_plcp<false><<<nblocks(slices, 128), 128, 0, st>>>(rs, phi);
	k_lcp_from_plcp<<<nblocks((size_t)N + 1, 256), 256, 0, st>>>(E->SA, phi, N, E->LCP);
	ctx->st.esa_launches += 3;
	dfree(ctx, phi);
	return ANDI_OK;
}

// Word offsets of the presence levels 1..K-1 inside E->present.bits (sized by esa_ensure).
static void presence_layout(andi_esa *E) {
	size_t words = 0;
	for (int m = 1; m < E->K; m++) {
		E->present.offset[m] = (u32)words;
		words += (((size_t)1 << (2 * m)) + 31) / 32;
	}
}

// Level K-1 of the presence bitmaps was left by k_bucket_sort; derive the lower levels, patch in
// the suffixes that end early, then the prefix-length table and the walk's directory view.
// Needs the final suffix array (the view holds text positions).
static int build_prefix_lengths(andi_ctx *ctx, andi_esa *E) {
	cudaStream_t st = ctx->stream;
	const int K = E->K;
	TextView rs = rs_view(E);
	if (K >= 3) {  // level K-2 by the whole grid, everything below it by one CTA
		u32 nb = 1u << (2 * (K - 2));
		k_presence_down<<<nblocks((nb + 31) / 32, 256), 256, 0, st>>>(E->present.bits + E->present.offset[K - 1], nb,
																	   E->present.bits + E->present.offset[K - 2]);
		if (K >= 4) k_presence_down_all<<<1, 1024, 0, st>>>(E->present, K - 2);
	}
	if (E->has_sep) {
		k_presence_patch<<<nblocks(E->N, 256), 256, 0, st>>>(rs, K, E->present, 0, E->N);
	} else {
		// only the K positions in front of '#' and of the text end can hold a shorter run
		u32 k = (u32)K;
		k_presence_patch<<<1, 32, 0, st>>>(rs, K, E->present, E->n >= k ? E->n - k : 0, k + 1);
		k_presence_patch<<<1, 32, 0, st>>>(rs, K, E->present, E->N >= k ? E->N - k : 0, k + 1);
	}
	k_prefix_len<<<nblocks((size_t)1 << (2 * (K - 1)), 256), 256, 0, st>>>(E->present, K, E->dir, E->SA, E->code, E->plen, E->fdir);
	ctx->st.esa_launches += 3 + (K >= 4 ? 2 : (K >= 3 ? 1 : 0));
	return ANDI_OK;
}

static int build_full(andi_ctx *ctx, andi_esa *E) {
	// CLD, FVC, prefix cache: only for the reference-visible esa_s (ANDI_ESA_FULL)
	cudaStream_t st = ctx->stream;
	const u32 N = E->N;
	dfree(ctx, E->CLD), dfree(ctx, E->FVC), dfree(ctx, E->cache);
	CK(dalloc(ctx, &E->CLD, (size_t)N + 1));
	CK(dalloc(ctx, &E->FVC, (size_t)N));
	CK(dalloc(ctx, &E->cache, (size_t)1 << 20));
	TextView rs = rs_view(E);
	k_fvc<<<nblocks(N, 256), 256, 0, st>>>(rs, E->SA, E->LCP, E->FVC);
	MinPyramid P{};
	P.level[0] = E->LCP;
	P.size[0] = N + 1;
	P.levels = 1;
	std::vector<int32_t *> owned;
	while (P.size[P.levels - 1] > 32) {
		u32 ns = (P.size[P.levels - 1] + 31) / 32;
		int32_t *buf = nullptr;
		CK(dalloc(ctx, &buf, ns));
		owned.push_back(buf);
		k_min_reduce32<<<nblocks(ns, 256), 256, 0, st>>>(P.level[P.levels - 1], P.size[P.levels - 1], buf, ns);
		P.level[P.levels] = buf;
		P.size[P.levels] = ns;
		P.levels++;
		ctx->st.esa_launches++;
	}
	k_cld<<<nblocks(N, 256), 256, 0, st>>>(P, N, E->CLD);
	for (auto b : owned) cudaFreeAsync(b, st);
	EsaView V;
	V.rs = rs, V.SA = E->SA, V.LCP = E->LCP, V.CLD = E->CLD, V.FVC = E->FVC;
	k_prefix_cache<<<nblocks(1u << 20, 128), 128, 0, st>>>(V, E->cache);
	ctx->st.esa_launches += 3;
	E->full = true;
	return ANDI_OK;
}

static void esa_release(andi_esa *E) {
	andi_ctx *ctx = E->ctx;
	dfree(ctx, E->code), dfree(ctx, E->spec), dfree(ctx, E->sep3), dfree(ctx, E->SA), dfree(ctx, E->LCP);
	dfree(ctx, E->dir), dfree(ctx, E->present.bits), dfree(ctx, E->plen), dfree(ctx, E->fdir), dfree(ctx, E->CLD), dfree(ctx, E->FVC);
	dfree(ctx, E->cache);
	E->cap_words = E->cap_n = E->cap_kmers = E->cap_present = 0;
	E->full = false;
}

// Make sure the arrays of E can hold an index of E->N characters with directory depth E->K.
static int esa_ensure(andi_ctx *ctx, andi_esa *E) {
	const size_t nw = plane_words(E->N);
	if (nw > E->cap_words) {
		dfree(ctx, E->code), dfree(ctx, E->spec), dfree(ctx, E->sep3);
		CK(dalloc(ctx, &E->code, nw));
		CK(dalloc(ctx, &E->spec, nw));
		CK(dalloc(ctx, &E->sep3, nw));
		E->cap_words = nw;
	}
	if ((size_t)E->N > E->cap_n) {
		dfree(ctx, E->SA), dfree(ctx, E->LCP);
		CK(dalloc(ctx, &E->SA, E->N));
		CK(dalloc(ctx, &E->LCP, (size_t)E->N + 1));
		E->cap_n = E->N;
	}
	if (E->K >= 2) {
		const size_t kmers = (size_t)1 << (2 * E->K);
		if (kmers > E->cap_kmers) {
			dfree(ctx, E->dir), dfree(ctx, E->plen), dfree(ctx, E->fdir);
			CK(dalloc(ctx, &E->dir, kmers));
			CK(dalloc(ctx, &E->plen, kmers));
			CK(dalloc(ctx, &E->fdir, kmers));
			E->cap_kmers = kmers;
		}
		size_t words = 0;
		for (int m = 1; m < E->K; m++) words += (((size_t)1 << (2 * m)) + 31) / 32;
		if (words > E->cap_present) {
			dfree(ctx, E->present.bits);
			CK(dalloc(ctx, &E->present.bits, words));
			E->cap_present = words;
		}
	}
	return ANDI_OK;
}

static int scratch_ensure(andi_ctx *ctx, size_t kmers, size_t N) {
	auto &b = ctx->bs;
	if (!b.flags) CK(dalloc(ctx, &b.flags, 8));
	if (!b.deep) CK(dalloc(ctx, &b.deep, 2 * (size_t)ANDI_DEEP_LIST + ANDI_DEEP_LCP_LIST));
	if (kmers > b.kmers_cap) {
		dfree(ctx, b.hist_alloc), dfree(ctx, b.bstart), dfree(ctx, b.scan_state);
		CK(dalloc(ctx, &b.hist_alloc, kmers + 1 + 4));
		CK(cudaMemsetAsync(b.hist_alloc, 0, 4 * sizeof(u32), ctx->stream));
		b.hist = b.hist_alloc + 4;
		CK(dalloc(ctx, &b.bstart, kmers + 1));
		CK(dalloc(ctx, &b.scan_state, 1024));  // one word per CTA of k_scan_buckets
		b.kmers_cap = kmers;
	}
	if (N > b.n_cap) {
		dfree(ctx, b.grp), dfree(ctx, b.rank), dfree(ctx, b.amb);
		CK(dalloc(ctx, &b.grp, N));
		CK(dalloc(ctx, &b.rank, N));
		CK(dalloc(ctx, &b.amb, N));
		b.n_cap = N;
	}
	return ANDI_OK;
}

// Exclusive prefix sums of hist[0, n) into out[0, n] (out may be hist), see k_scan_buckets.
static int scan_buckets(andi_ctx *ctx, const u32 *hist, size_t n, u32 *out, u64 *dir64) {
	auto &b = ctx->bs;
	// co-resident grid: two CTAs of 256 threads per SM (the kernel spins on its predecessors' sums)
	unsigned ctas = (unsigned)ctx->sm_count * 2u;
	const size_t tiles = (n + ANDI_SCAN_TILE - 1) / ANDI_SCAN_TILE;
	if (tiles < ctas) ctas = (unsigned)std::max<size_t>(tiles, 1);
	const u32 per_cta = (u32)(((tiles + ctas - 1) / ctas) * ANDI_SCAN_TILE);
	CK(cudaMemsetAsync(b.scan_state, 0, (size_t)ctas * sizeof(unsigned long long), ctx->stream));
	k_scan_buckets<<<ctas, 256, 0, ctx->stream>>>(hist, (u32)n, per_cta, out, dir64, b.scan_state);
	ctx->st.esa_launches++;
	return ANDI_OK;
}

// The padded-suffix list (texts with separators), `cap` entries.
static int padded_ensure(andi_ctx *ctx, size_t cap) {
	auto &b = ctx->bs;
	if (cap <= b.pl_cap) return ANDI_OK;
	for (int x = 0; x < 2; x++) {
		dfree(ctx, b.pl_key[x]), dfree(ctx, b.pl_idx[x]);
		CK(dalloc(ctx, &b.pl_key[x], cap));
		CK(dalloc(ctx, &b.pl_idx[x], cap));
	}
	if (b.pl_tmp) cudaFreeAsync(b.pl_tmp, ctx->stream);
	b.pl_tmp = nullptr, b.pl_tmp_bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, b.pl_tmp_bytes, b.pl_key[0], b.pl_key[1], b.pl_idx[0], b.pl_idx[1], (int)cap, 0, 63,
									ctx->stream);
	CK(cudaMallocAsync(&b.pl_tmp, b.pl_tmp_bytes, ctx->stream));
	b.pl_cap = cap;
	return ANDI_OK;
}

// Sort the padded list and write it to the bucket fronts.
static int padded_finish(andi_ctx *ctx, andi_esa *E, const TextView &rs) {
	auto &b = ctx->bs;
	cudaStream_t st = ctx->stream;
	const u32 cap = (u32)b.pl_cap;
	CK(cub::DeviceRadixSort::SortPairs(b.pl_tmp, b.pl_tmp_bytes, b.pl_key[0], b.pl_key[1], b.pl_idx[0], b.pl_idx[1], (int)cap, 0,
									   63, st));
	k_padded_ties<true><<<nblocks(cap, 128), 128, 0, st>>>(rs, b.pl_key[1], b.pl_idx[1], b.flags + 2, cap);
	k_padded_place<<<nblocks(cap, 128), 128, 0, st>>>(E->K, b.pl_key[1], b.pl_idx[1], b.flags + 2, cap, b.bstart, E->SA);
	ctx->st.esa_launches += 2;
	ctx->st.cub_calls += 1;
	return ANDI_OK;
}

// Directories up to this many k-mers (K <= 12: a 64 MB histogram) are bucketed by counting sort
// with atomics on L2-resident tables; deeper ones (texts of hundreds of Mbp) go through a radix sort.
#define ANDI_BUCKET_ATOMIC_KMERS ((size_t)1 << 24)

static int build_index_bucket(andi_ctx *ctx, andi_esa *E) {
	const u32 N = E->N;
	const int K = E->K;
	cudaStream_t st = ctx->stream;
	const size_t kmers = (size_t)1 << (2 * K);
	TextView rs = rs_view(E);
	int rc = scratch_ensure(ctx, kmers, N);
	if (rc) return rc;
	auto &b = ctx->bs;
	const bool sep = E->has_sep;
	const u32 *fvalid = nullptr, *bend = nullptr;
	if (sep && b.pl_cap == 0) {
		rc = padded_ensure(ctx, (size_t)1 << 16);
		if (rc) return rc;
	}
rebuild:
	CK(cudaMemsetAsync(b.flags, 0, 8 * sizeof(u32), st));
	const TieSink sink{b.flags, sep ? nullptr : b.deep, ANDI_DEEP_LIST, ANDI_DEEP_LCP_LIST};
	PaddedList pl{b.pl_key[0], b.pl_idx[0], b.flags + 2, (u32)b.pl_cap};
	if (sep) {	// unused list slots sort to the end
		CK(cudaMemsetAsync(b.pl_key[0], 0xff, b.pl_cap * sizeof(u64), st));
		CK(cudaMemsetAsync(b.pl_idx[0], 0xff, b.pl_cap * sizeof(u32), st));
	}
	const bool atomic_path = kmers <= ANDI_BUCKET_ATOMIC_KMERS;
	bool slots = atomic_path && !sep;  // bucket ends are in hist (hist - 1 = bucket starts), not in bstart / bend tables
	bool sorted = false;			   // the bucketing pass has sorted the buckets already (k_part_sort)
	if (atomic_path) {
		// counting sort with L2-resident tables: histogram, scan, scatter
		CK(cudaMemsetAsync(b.hist, 0, (kmers + 1) * sizeof(u32), st));
		k_bucket_hist<<<nblocks(N, 256), 256, 0, st>>>(rs, K, b.hist);
		if (!sep) {
			// the histogram becomes the scatter cursor in place (and the scan writes the directory entries of
			// the empty buckets); after the scatter hist[key] = end of bucket key, hist[key - 1] its start
			rc = scan_buckets(ctx, b.hist, kmers, b.hist, E->dir);
			if (rc) return rc;
			k_bucket_scatter<<<nblocks(N, 256), 256, 0, st>>>(rs, K, b.hist, E->SA);
			bend = b.hist;
		} else {
			// valid suffixes fill their bucket from the back: the cursor starts at the bucket end and
			// stops at the first valid one; the padded ones go through the list to the front
			rc = scan_buckets(ctx, b.hist, kmers, b.bstart, nullptr);
			if (rc) return rc;
			CK(cudaMemcpyAsync(b.hist, b.bstart + 1, kmers * sizeof(u32), cudaMemcpyDeviceToDevice, st));
			k_bucket_scatter_spec<<<nblocks(N, 256), 256, 0, st>>>(rs, K, b.hist, E->SA, pl);
			bend = b.bstart + 1, fvalid = b.hist;
		}
		ctx->st.esa_launches += 2;
	} else if (!sep) {
		// deep directory, no separators: own two-level counting sort (sa_bucket.cuh)
		const int K2 = K / 2, K1 = K - K2;
		const u32 parts = 1u << (2 * K1), bins = 1u << (2 * K2);
		// one CTA per SM: every CTA keeps one open write sector per part, and 148 x 16384 x 32 B has to stay in L2
		// (two per SM measured the same: 32.2 vs 32.5 ms per 120 Mbp subject)
		const unsigned ctas = (unsigned)ctx->sm_count;
		const u32 per_cta = (N + ctas - 1) / ctas;
		if (!ctx->part_attr_set) {	// per device: a function attribute belongs to the device's context
			CK(cudaFuncSetAttribute(k_part_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
			CK(cudaFuncSetAttribute(k_part_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
			CK(cudaFuncSetAttribute(k_part_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 128));
			ctx->part_attr_set = true;
		}
		u32 *hist1 = b.bstart, *start1 = b.bstart + parts + 4;	 // level-1 tables live in the (unused) bucket-start scratch
		CK(cudaMemsetAsync(hist1, 0, (parts + 1) * sizeof(u32), st));
		k_part_hist<<<ctas, 1024, parts * sizeof(u32), st>>>(rs, K, K2, per_cta, hist1);
		rc = scan_buckets(ctx, hist1, parts, start1, nullptr);
		if (rc) return rc;
		CK(cudaMemcpyAsync(hist1, start1, parts * sizeof(u32), cudaMemcpyDeviceToDevice, st));	// cursors
		k_part_scatter<<<ctas, 1024, parts * sizeof(u32), st>>>(rs, K, K2, per_cta, hist1, b.grp);
		presence_layout(E);
		CK(cudaMemsetAsync(E->present.bits + E->present.offset[K - 1], 0, ((((size_t)1 << (2 * (K - 1))) + 31) / 32) * sizeof(u32), st));
		k_part_sort<<<parts, 1024, bins * sizeof(u32) + 128, st>>>(rs, K, K2, start1, b.grp, E->SA, b.hist, E->dir, sink,
																	 E->present.bits + E->present.offset[K - 1]);
		bend = b.hist;	// bucket ends; hist - 1 = bucket starts (hist_alloc holds zeros in front)
		slots = true, sorted = true;
		ctx->st.esa_launches += 3;
	} else {
		// deep directory WITH separators (join mode on a genome of hundreds of Mbp): (key, position)
		// pairs through the library radix sort, bounds from the sorted keys
		u32 *keys_a = nullptr, *keys_b = nullptr, *idx = nullptr;
		CK(dalloc(ctx, &keys_a, N));
		CK(dalloc(ctx, &keys_b, N));
		CK(dalloc(ctx, &idx, N));
		const int bits = 2 * K + 1;
		size_t sort_bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_a, keys_b, idx, E->SA, (int)N, 0, bits, st);
		void *tmp = nullptr;
		CK(cudaMallocAsync(&tmp, sort_bytes, st));
		k_bucket_keys_spec<<<nblocks(N, 256), 256, 0, st>>>(rs, K, keys_a, idx, pl);
		CK(cub::DeviceRadixSort::SortPairs(tmp, sort_bytes, keys_a, keys_b, idx, E->SA, (int)N, 0, bits, st));
		CK(cudaMemsetAsync(b.bstart, 0, kmers * sizeof(u32), st));
		CK(cudaMemsetAsync(b.hist, 0, kmers * sizeof(u32), st));
		if (kmers > b.fvalid_cap) {
			dfree(ctx, b.fvalid);
			CK(dalloc(ctx, &b.fvalid, kmers));
			b.fvalid_cap = kmers;
		}
		CK(cudaMemsetAsync(b.fvalid, 0, kmers * sizeof(u32), st));
		k_bucket_bounds_spec<<<nblocks(N, 256), 256, 0, st>>>(keys_b, N, b.bstart, b.hist, b.fvalid);
		fvalid = b.fvalid;
		bend = b.hist;
		cudaFreeAsync(tmp, st);
		dfree(ctx, keys_a), dfree(ctx, keys_b), dfree(ctx, idx);
		ctx->st.esa_launches += 2;
		ctx->st.cub_calls += 1;
	}
	presence_layout(E);
	u32 *present_top = E->present.bits + E->present.offset[K - 1];
	if (!sorted) CK(cudaMemsetAsync(present_top, 0, ((((size_t)1 << (2 * (K - 1))) + 31) / 32) * sizeof(u32), st));
	if (sep) {
		rc = padded_finish(ctx, E, rs);
		if (rc) return rc;
		k_bucket_sort<true><<<nblocks(kmers, 128), 128, 0, st>>>(rs, K, b.bstart, bend, fvalid, E->SA, E->dir, sink, atomic_path, present_top);
		k_lcp_direct<true><<<nblocks((size_t)N + 1, 256), 256, 0, st>>>(rs, E->SA, ANDI_LCP_DIRECT_CAP, E->LCP, sink);
	} else {
		// (texts without separators: the counting-sort path sorts its buckets with one thread per
		// slot, the two-level path has sorted them inside k_part_sort already)
		if (!sorted) k_bucket_sort_slots<<<nblocks(N, 128), 128, 0, st>>>(rs, K, b.hist, E->SA, E->dir, sink, present_top);
		// the buckets and, below, the LCP values that met a repeat: listed, finished by one warp each
		// (two small launches that find empty lists on repeat-free texts)
		k_sort_deep<<<2 * ctx->sm_count, 256, 0, st>>>(rs, E->SA, sink);
		k_lcp_direct<false><<<nblocks((size_t)N + 1, 256), 256, 0, st>>>(rs, E->SA, ANDI_LCP_DIRECT_CAP, E->LCP, sink);
		k_lcp_deep<<<2 * ctx->sm_count, 256, 0, st>>>(rs, E->SA, ANDI_LCP_DIRECT_CAP, E->LCP, sink);
		ctx->st.esa_launches += 2;
	}
	ctx->st.esa_launches += 2;
	u32 h_flags[4] = {0, 0, 0, 0};
	if (!rc) {
		CK(cudaMemcpyAsync(h_flags, b.flags, sizeof h_flags, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	if (!rc && sep && h_flags[2] > b.pl_cap) {
		// more padded suffixes than the list held: grow it and build again (first subject only, as a rule)
		size_t want = (size_t)h_flags[2] * 2;
		if (want > (size_t)INT_MAX) {
			ctx->err = "too many separators";
			return ANDI_ERR_TOO_LONG;
		}
		rc = padded_ensure(ctx, want);
		if (rc) return rc;
		goto rebuild;
	}
	if (!rc && h_flags[0]) {
		// tied suffixes: materialise their groups, refine them, then take the LCP again
		if (E->has_sep)
			k_bucket_groups<true><<<nblocks(kmers, 128), 128, 0, st>>>(rs, K, b.bstart, bend, fvalid, E->SA, b.grp, b.rank, b.amb);
		else
			// (slots: the scatter left the bucket ends in hist, so hist - 1 is the array of bucket starts)
			k_bucket_groups<false><<<nblocks(kmers, 128), 128, 0, st>>>(rs, K, slots ? b.hist - 1 : b.bstart, bend, nullptr, E->SA, b.grp, b.rank, b.amb);
		ctx->st.esa_launches++;
		rc = doubling_rounds(ctx, E, b.grp, b.rank, b.amb, (u32)K);
		if (!rc) {
			CK(cudaMemsetAsync(b.flags + 1, 0, sizeof(u32), st));
			CK(cudaMemsetAsync(b.flags + 5, 0, sizeof(u32), st));
			if (E->has_sep) {
				k_lcp_direct<true><<<nblocks((size_t)N + 1, 256), 256, 0, st>>>(rs, E->SA, ANDI_LCP_DIRECT_CAP, E->LCP, sink);
			} else {
				k_lcp_direct<false><<<nblocks((size_t)N + 1, 256), 256, 0, st>>>(rs, E->SA, ANDI_LCP_DIRECT_CAP, E->LCP, sink);
				k_lcp_deep<<<2 * ctx->sm_count, 256, 0, st>>>(rs, E->SA, ANDI_LCP_DIRECT_CAP, E->LCP, sink);
			}
			ctx->st.esa_launches++;
			CK(cudaMemcpyAsync(h_flags, b.flags, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
		}
	}
	if (!rc && h_flags[1]) rc = build_lcp_phi(ctx, E);
	if (!rc) rc = build_prefix_lengths(ctx, E);	 // the suffix array is final
	return rc;
}

// RS planes are in place; build everything else.
static int build_index(andi_ctx *ctx, andi_esa *E, unsigned flags) {
	NvtxRange range("andi: index build (esa_init)");
	cudaEvent_t e0 = get_event(ctx), e1 = get_event(ctx);
	mark(ctx, e0);
	if (!ctx->first_ev) {
		ctx->first_ev = get_event(ctx);
		mark(ctx, ctx->first_ev);
	}
	int rc = esa_ensure(ctx, E);
	if (rc) return rc;
	{  // separator hints for a join-mode walk (the queries may hold separators even if RS has only '#')
		const size_t nw = plane_words(E->N);
		k_sep3<<<nblocks(nw, 256), 256, 0, ctx->stream>>>(E->spec, nw, E->sep3);
		ctx->st.esa_launches++;
	}
	if (E->K >= 2) {
		rc = build_index_bucket(ctx, E);
	} else {
		rc = build_sa_radix(ctx, E);
		if (!rc) rc = build_lcp_phi(ctx, E);
	}
	if (!rc && (flags & ANDI_ESA_FULL)) rc = build_full(ctx, E);
	mark(ctx, e1);
	ctx->esa_ev.emplace_back(e0, e1);
	if (!ctx->last_ev) ctx->last_ev = get_event(ctx);
	mark(ctx, ctx->last_ev);
	ctx->st.subjects++;
	if (!rc) {
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) {
			ctx->err = std::string("index kernels: ") + cudaGetErrorString(e);
			rc = ANDI_ERR_CUDA;
		}
	}
	return rc;
}

