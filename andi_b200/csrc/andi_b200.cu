// andi_b200/csrc/andi_b200.cu -- C ABI of libandi_b200.so (include/andi_b200.h), host side.
//
// Orchestrates the sm_100a kernels of esa_kernels.cuh / walk_kernels.cuh. One context per
// GPU, one stream, stream-ordered allocations (cudaMallocAsync) so index construction for
// thousands of subjects never calls cudaMalloc in the steady state. Radix sort / scan / select
// are CUB device primitives (library code, like cuBLAS for a GEMM); every other kernel is ours.
#include "../../include/andi_b200.h"
#include "walk_kernels.cuh"
#include "pack_tma.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>	// doubling fallback (repeat-rich texts) only
#include <cub/device/device_select.cuh>

#include <nvtx3/nvToolsExt.h>  // header-only; ranges show up in Nsight Systems (tracing)

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// ------------------------------------------------------------------ plumbing

// NVTX range for the lifetime of a scope: pool upload / index build / walk of one subject
struct NvtxRange {
	explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
};

struct andi_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	std::string err;
	int sm_count = 148;

	// pool (src/process.h:11 `seq_t *sequences, size_t n`)
	size_t n = 0;
	std::vector<size_t> len;
	std::vector<double> gc;
	std::vector<int> has_sep;
	std::vector<size_t> word_off;  // u64-word offset of each sequence's planes
	u64 *pool_code = nullptr, *pool_spec = nullptr;
	size_t pool_words = 0;
	// The planes and the character staging buffer are KEPT across pools (capacities below): setting a
	// pool of the same size again allocates nothing. (Re-allocating 10 GB per call made every second
	// andi_pool_set_host take 0.7 - 2.5 s instead of 0.12 s while the memory pool re-mapped.)
	size_t plane_cap = 0;
	unsigned char *stage_chars = nullptr;
	size_t stage_cap = 0;
	uint4 *pool_comp = nullptr;	 // prefix composition per word, built on first LOGDET / ANI use
	unsigned char *pool_sep3 = nullptr;	 // separator hints per word, built on first join-mode walk
	QueryView *d_queries = nullptr;
	bool any_sep = false;

	// index-build scratch, grown on demand and reused for every subject (no allocation in the
	// steady state of andi_dist_rows)
	struct {
		// hist has four zero entries in front (hist_alloc): hist - 1 is then the array of bucket starts
		// once the scatter has turned hist into the array of bucket ends
		u32 *hist_alloc = nullptr, *hist = nullptr, *bstart = nullptr, *grp = nullptr, *rank = nullptr, *flags = nullptr;
		u32 *deep = nullptr;  // sa_bucket.cuh, TieSink: listed buckets (pairs), then listed LCP slots
		unsigned char *amb = nullptr;
		unsigned long long *scan_state = nullptr;  // k_scan_buckets: one word per tile + the ticket counter
		size_t kmers_cap = 0, n_cap = 0;
		// padded-suffix list of texts with separators (sa_bucket.cuh), double-buffered for the sort
		u64 *pl_key[2] = {nullptr, nullptr};
		u32 *pl_idx[2] = {nullptr, nullptr};
		void *pl_tmp = nullptr;
		size_t pl_tmp_bytes = 0, pl_cap = 0;
		u32 *fvalid = nullptr;	// radix path only
		size_t fvalid_cap = 0;
	} bs;

	unsigned long long *walk_counter = nullptr;  // unit dispenser of the walk kernels
	bool part_attr_set = false;					 // k_part_* may use 64 KB of dynamic shared memory on this device
	u32 *walk_bad = nullptr;					 // per pair: a chunk boundary did not synchronise (k_walk_reduce_sum)
	size_t walk_bad_cap = 0;
	// sum of the anchor lengths the chunk walks of the last launch ended with, and their number
	// (k_walk_reduce_sum); copied to the host after every walk, looked at before the next one
	unsigned long long *walk_stat = nullptr, *h_walk_stat = nullptr;
	bool walk_burst = false;  // launch k_walk_v3<.., BURST = true>: the pool is one of near-identical genomes

	// pinned staging planes of andi_pool_set_host (the pool is packed on the host, host_pack.c)
	u64 *h_code = nullptr, *h_spec = nullptr;
	size_t h_words = 0;

	// Second lane of andi_dist_rows: a helper context on its own stream that BORROWS this pool, so
	// that the index build and the walk of subject i+1 fill the tail of subject i's walk
	// (walk_host.cuh). Created on first use, destroyed with its owner.
	andi_ctx *helper = nullptr;
	bool borrowed_pool = false;
	cudaEvent_t fork_ev = nullptr, join_ev = nullptr;

	andi_stats st{};
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> esa_ev, walk_ev;
	std::vector<cudaEvent_t> free_ev;
	cudaEvent_t first_ev = nullptr, last_ev = nullptr;
};

struct andi_esa {
	andi_ctx *ctx = nullptr;
	u32 n = 0, N = 0;
	u64 *code = nullptr, *spec = nullptr;
	unsigned char *sep3 = nullptr;	// separator hints of the RS planes (k_sep3), has_sep only
	u32 *SA = nullptr;
	int32_t *LCP = nullptr;
	u64 *dir = nullptr;
	PresenceLevels present{};
	unsigned char *plen = nullptr;
	u64 *fdir = nullptr;  // 4^K: the walk's own view of the directory (written by k_prefix_len)
	int K = 0;
	bool has_sep = false;
	bool full = false;
	int32_t *CLD = nullptr;
	char *FVC = nullptr;
	Inter *cache = nullptr;
	u32 self = 0xffffffffu;
	u32 threshold = 0;
	// capacities of the arrays above (an andi_esa can be rebuilt in place for another subject)
	size_t cap_words = 0, cap_n = 0, cap_kmers = 0, cap_present = 0;
};

static thread_local std::string g_create_err;

#define CK(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess) {                                                                   \
			ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
			return ANDI_ERR_CUDA;                                                                  \
		}                                                                                          \
	} while (0)

static inline unsigned nblocks(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }
static inline size_t plane_words(size_t chars) { return chars / 32 + 4; }  // >= 3 guard words (window64)

template <class T>
static cudaError_t dalloc(andi_ctx *ctx, T **p, size_t count) {
	return cudaMallocAsync((void **)p, std::max<size_t>(count, 1) * sizeof(T), ctx->stream);
}
template <class T>
static void dfree(andi_ctx *ctx, T *&p) {
	if (p) cudaFreeAsync((void *)p, ctx->stream);
	p = nullptr;
}

static cudaEvent_t get_event(andi_ctx *ctx) {
	if (!ctx->free_ev.empty()) {
		cudaEvent_t e = ctx->free_ev.back();
		ctx->free_ev.pop_back();
		return e;
	}
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}

static void mark(andi_ctx *ctx, cudaEvent_t e) {
	cudaEventRecord(e, ctx->stream);
}

// Fold finished event pairs into the stats (call after a stream synchronize).
static void harvest_events(andi_ctx *ctx) {
	float ms;
	for (auto &p : ctx->esa_ev) {
		if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) ctx->st.esa_ms += ms;
		ctx->free_ev.push_back(p.first), ctx->free_ev.push_back(p.second);
	}
	for (auto &p : ctx->walk_ev) {
		if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) ctx->st.walk_ms += ms;
		ctx->free_ev.push_back(p.first), ctx->free_ev.push_back(p.second);
	}
	ctx->esa_ev.clear(), ctx->walk_ev.clear();
	if (ctx->first_ev && ctx->last_ev) {
		if (cudaEventElapsedTime(&ms, ctx->first_ev, ctx->last_ev) == cudaSuccess) ctx->st.total_ms += ms;
		ctx->free_ev.push_back(ctx->first_ev), ctx->free_ev.push_back(ctx->last_ev);
		ctx->first_ev = ctx->last_ev = nullptr;
	}
}

extern "C" int andi_device_count(void) {
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return count;
}

extern "C" int andi_ctx_create(int device, void *stream, andi_ctx **out) {
	if (!out) return ANDI_ERR_ARG;
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || device < 0 || device >= count) {
		g_create_err = e != cudaSuccess ? std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)
										: "no such CUDA device";
		return ANDI_ERR_CUDA;
	}
	andi_ctx *ctx = new andi_ctx();
	ctx->device = device;
	if ((e = cudaSetDevice(device)) != cudaSuccess) {
		g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
		delete ctx;
		return ANDI_ERR_CUDA;
	}
	if (stream) {
		ctx->stream = (cudaStream_t)stream;
	} else {
		if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
			g_create_err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
			delete ctx;
			return ANDI_ERR_CUDA;
		}
		ctx->own_stream = true;
	}
	cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
	// keep freed blocks in the stream-ordered pool instead of returning them to the driver
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
		uint64_t keep = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
	}
	*out = ctx;
	return ANDI_OK;
}

static void pool_release(andi_ctx *ctx) {
	if (ctx->helper) ctx->helper->n = 0;  // its borrowed pointers die with this pool
	if (ctx->borrowed_pool) {
		ctx->pool_code = ctx->pool_spec = nullptr, ctx->pool_comp = nullptr, ctx->pool_sep3 = nullptr, ctx->d_queries = nullptr;
		ctx->plane_cap = 0;
		ctx->borrowed_pool = false;
	}
	// (pool_code / pool_spec stay allocated for the next pool: planes_ensure)
	dfree(ctx, ctx->pool_comp);
	dfree(ctx, ctx->pool_sep3);
	dfree(ctx, ctx->d_queries);
	ctx->n = 0;
	ctx->len.clear(), ctx->gc.clear(), ctx->has_sep.clear(), ctx->word_off.clear();
	ctx->any_sep = false;
}

extern "C" void andi_ctx_destroy(andi_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->helper) {
		andi_ctx_destroy(ctx->helper);
		ctx->helper = nullptr;
	}
	if (ctx->h_code) cudaFreeHost(ctx->h_code);
	if (ctx->h_spec) cudaFreeHost(ctx->h_spec);
	if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
	if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
	cudaStreamSynchronize(ctx->stream);
	harvest_events(ctx);
	pool_release(ctx);
	dfree(ctx, ctx->pool_code), dfree(ctx, ctx->pool_spec), dfree(ctx, ctx->stage_chars);
	dfree(ctx, ctx->bs.hist_alloc), dfree(ctx, ctx->bs.bstart), dfree(ctx, ctx->bs.grp), dfree(ctx, ctx->bs.rank);
	dfree(ctx, ctx->bs.scan_state);
	dfree(ctx, ctx->walk_stat);
	if (ctx->h_walk_stat) cudaFreeHost(ctx->h_walk_stat);
	dfree(ctx, ctx->bs.deep), dfree(ctx, ctx->bs.flags), dfree(ctx, ctx->bs.amb), dfree(ctx, ctx->walk_counter), dfree(ctx, ctx->walk_bad);
	dfree(ctx, ctx->bs.pl_key[0]), dfree(ctx, ctx->bs.pl_key[1]), dfree(ctx, ctx->bs.pl_idx[0]), dfree(ctx, ctx->bs.pl_idx[1]);
	dfree(ctx, ctx->bs.fvalid);
	if (ctx->bs.pl_tmp) cudaFreeAsync(ctx->bs.pl_tmp, ctx->stream);
	cudaStreamSynchronize(ctx->stream);
	for (auto e : ctx->free_ev) cudaEventDestroy(e);
	if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" const char *andi_last_error(const andi_ctx *ctx) {
	return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int andi_get_stats(const andi_ctx *ctx, andi_stats *out) {
	if (!ctx || !out) return ANDI_ERR_ARG;
	*out = ctx->st;
	return ANDI_OK;
}

extern "C" void andi_reset_stats(andi_ctx *ctx) {
	if (ctx) ctx->st = andi_stats{};
}

// ------------------------------------------------------------------ threshold (host, FP64)
// src/sequence.c:296-373. Kept on the host in double precision with the reference's
// evaluation order so that the integer threshold is identical.

static size_t n_choose_k(size_t n, size_t k) {
	if (n == 0 || k > n) return 0;
	if (k == 0 || k == n) return 1;
	k = std::min(k, n - k);
	size_t r = 1;
	for (size_t i = 1; i <= k; i++) r = r * (n - k + i) / i;
	return r;
}

static double shustring_cdf(size_t x, double p, size_t l) {
	const double xx = (double)x, ll = (double)l;
	double acc = 0.0;
	for (size_t k = 0; k <= x; k++) {
		double kk = (double)k;
		double t = pow(p, kk) * pow(0.5 - p, xx - kk);
		acc += pow(2, xx) * (t * pow(1 - t, ll)) * (double)n_choose_k(x, k);
		if (acc >= 1.0) return 1.0;
	}
	return acc;
}

extern "C" size_t andi_threshold(double p_value, double gc, size_t rs_len) {
	size_t x = 1;
	while (shustring_cdf(x, gc / 2, rs_len) < 1 - p_value) x++;
	return x;
}

// ------------------------------------------------------------------ pool

// Planes for a pool of `words` words: the ones this context already has, if they are large enough.
static int planes_ensure(andi_ctx *ctx, size_t words) {
	if (words > ctx->plane_cap) {
		dfree(ctx, ctx->pool_code), dfree(ctx, ctx->pool_spec);
		ctx->plane_cap = 0;
		CK(dalloc(ctx, &ctx->pool_code, words));
		CK(dalloc(ctx, &ctx->pool_spec, words));
		ctx->plane_cap = words;
	}
	ctx->pool_words = words;
	return ANDI_OK;
}

static int pool_finish(andi_ctx *ctx, const unsigned char *d_chars, const std::vector<size_t> &offs) {
	// d_chars: all sequences in HBM; pack them, count GC and separators.
	const size_t n = ctx->n;
	size_t words = 0;
	ctx->word_off.resize(n);
	for (size_t k = 0; k < n; k++) {
		ctx->word_off[k] = words;
		words += (plane_words(ctx->len[k]) + 1) & ~(size_t)1;  // keep 16-byte alignment
	}
	{
		int rc = planes_ensure(ctx, words);
		if (rc) return rc;
	}
	unsigned long long *d_cnt = nullptr;
	CK(dalloc(ctx, &d_cnt, 2 * n));
	CK(cudaMemsetAsync(d_cnt, 0, 2 * n * sizeof(unsigned long long), ctx->stream));
	// one TMA-staged launch for all full 8 KiB tiles of 16-byte aligned sequences, one launch for
	// the tails (pack_tma.cuh)
	{
		std::vector<PackSeq> ps(n);
		std::vector<u32> full(n);
		u32 ntiles = 0;
		for (size_t k = 0; k < n; k++) {
			bool aligned = ((uintptr_t)(d_chars + offs[k]) & 15u) == 0;
			full[k] = aligned ? (u32)(ctx->len[k] / ANDI_PACK_TILE) : 0u;
			ps[k].char_off = offs[k], ps[k].word_off = ctx->word_off[k], ps[k].len = (u32)ctx->len[k], ps[k].tile0 = ntiles;
			ntiles += full[k];
		}
		PackSeq *d_ps = nullptr;
		u32 *d_full = nullptr;
		CK(dalloc(ctx, &d_ps, n));
		CK(dalloc(ctx, &d_full, n));
		CK(cudaMemcpyAsync(d_ps, ps.data(), n * sizeof(PackSeq), cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(d_full, full.data(), n * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
		if (ntiles) {
			unsigned grid = std::min<unsigned>(ntiles, (unsigned)ctx->sm_count * 4u);
			k_pack_tma<<<grid, 256, 0, ctx->stream>>>(d_chars, d_ps, (u32)n, ntiles, ctx->pool_code, ctx->pool_spec, d_cnt);
			ctx->st.esa_launches++;
		}
		k_pack_tails<<<(unsigned)n, 256, 0, ctx->stream>>>(d_chars, d_ps, d_full, ctx->pool_code, ctx->pool_spec, d_cnt);
		ctx->st.esa_launches++;
		CK(cudaStreamSynchronize(ctx->stream));	 // ps / full go out of scope
		dfree(ctx, d_ps), dfree(ctx, d_full);
	}
	std::vector<unsigned long long> cnt(2 * n);
	CK(cudaMemcpyAsync(cnt.data(), d_cnt, 2 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	dfree(ctx, d_cnt);
	ctx->gc.resize(n), ctx->has_sep.resize(n);
	std::vector<QueryView> qv(n);
	for (size_t k = 0; k < n; k++) {
		ctx->gc[k] = (double)cnt[2 * k] / (double)ctx->len[k];
		ctx->has_sep[k] = cnt[2 * k + 1] != 0;
		ctx->any_sep |= ctx->has_sep[k] != 0;
		qv[k].t.code = ctx->pool_code + ctx->word_off[k];
		qv[k].t.spec = ctx->pool_spec + ctx->word_off[k];
		qv[k].t.len = (u32)ctx->len[k];
		qv[k].t.mid = 0xffffffffu;
		qv[k].has_sep = ctx->has_sep[k];
	}
	CK(dalloc(ctx, &ctx->d_queries, n));
	CK(cudaMemcpyAsync(ctx->d_queries, qv.data(), n * sizeof(QueryView), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return ANDI_OK;
}

static int pool_check(andi_ctx *ctx, const size_t *lens, size_t n) {
	if (!ctx || !lens || n == 0) return ANDI_ERR_ARG;
	for (size_t k = 0; k < n; k++) {
		if (lens[k] == 0) {
			ctx->err = "empty sequence";
			return ANDI_ERR_ARG;
		}
		if (lens[k] > (size_t)(INT_MAX - 1) / 2) {
			ctx->err = "sequence longer than (INT_MAX-1)/2";
			return ANDI_ERR_TOO_LONG;
		}
	}
	return ANDI_OK;
}

extern "C" {
void andi_host_pack_pool(const char *const *seqs, const size_t *lens, const size_t *word_off, const size_t *nwords, size_t k0, size_t k1,
						 uint64_t *code, uint64_t *spec, uint64_t *gc, uint64_t *sep);
}

// The pool from host memory, packed to 2 bits per base ON THE HOST (host_pack.c: AVX2, all cores)
// into pinned staging planes and uploaded packed, chunk by chunk, so that packing chunk c + 1
// overlaps the upload of chunk c: a quarter of the bytes on PCIe. Opt-in (ANDI_B200_HOST_PACK=1, see
// andi_pool_set_host); the default uploads the characters and packs on the device (k_pack_tma).
static int pool_set_host_packed(andi_ctx *ctx, const char *const *seqs, const size_t *lens, size_t n) {
	std::vector<size_t> nwords(n);
	size_t words = 0;
	ctx->word_off.resize(n);
	for (size_t k = 0; k < n; k++) {
		ctx->word_off[k] = words;
		nwords[k] = plane_words(lens[k]);
		words += (nwords[k] + 1) & ~(size_t)1;	// keep 16-byte alignment
	}
	if (words > ctx->h_words) {
		if (ctx->h_code) cudaFreeHost(ctx->h_code);
		if (ctx->h_spec) cudaFreeHost(ctx->h_spec);
		ctx->h_code = ctx->h_spec = nullptr, ctx->h_words = 0;
		CK(cudaHostAlloc((void **)&ctx->h_code, words * sizeof(u64), cudaHostAllocDefault));
		CK(cudaHostAlloc((void **)&ctx->h_spec, words * sizeof(u64), cudaHostAllocDefault));
		ctx->h_words = words;
	}
	{
		int rc = planes_ensure(ctx, words);
		if (rc) return rc;
	}
	std::vector<uint64_t> gc(n), sep(n);
	const size_t chunk_chars = (size_t)256 << 20;
	for (size_t k0 = 0; k0 < n;) {
		size_t k1 = k0, chars = 0;
		while (k1 < n && (k1 == k0 || chars + lens[k1] <= chunk_chars)) chars += lens[k1++];
		andi_host_pack_pool(seqs, lens, ctx->word_off.data(), nwords.data(), k0, k1, reinterpret_cast<uint64_t *>(ctx->h_code), reinterpret_cast<uint64_t *>(ctx->h_spec), gc.data(), sep.data());
		bool any = false;
		for (size_t k = k0; k < k1; k++) {
			any |= sep[k] != 0;
			if (nwords[k] & 1) ctx->h_code[ctx->word_off[k] + nwords[k]] = 0, ctx->h_spec[ctx->word_off[k] + nwords[k]] = 0;	// alignment pad
		}
		const size_t w0 = ctx->word_off[k0], w1 = k1 < n ? ctx->word_off[k1] : words;
		CK(cudaMemcpyAsync(ctx->pool_code + w0, ctx->h_code + w0, (w1 - w0) * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
		ctx->st.h2d_bytes += (w1 - w0) * sizeof(u64);
		if (any) {
			CK(cudaMemcpyAsync(ctx->pool_spec + w0, ctx->h_spec + w0, (w1 - w0) * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
			ctx->st.h2d_bytes += (w1 - w0) * sizeof(u64);
		} else {
			CK(cudaMemsetAsync(ctx->pool_spec + w0, 0, (w1 - w0) * sizeof(u64), ctx->stream));
		}
		k0 = k1;
	}
	ctx->gc.resize(n), ctx->has_sep.resize(n);
	std::vector<QueryView> qv(n);
	for (size_t k = 0; k < n; k++) {
		ctx->gc[k] = (double)gc[k] / (double)lens[k];  // src/sequence.c:196-207
		ctx->has_sep[k] = sep[k] != 0;
		ctx->any_sep |= ctx->has_sep[k] != 0;
		qv[k].t.code = ctx->pool_code + ctx->word_off[k], qv[k].t.spec = ctx->pool_spec + ctx->word_off[k];
		qv[k].t.len = (u32)lens[k], qv[k].t.mid = 0xffffffffu, qv[k].has_sep = ctx->has_sep[k];
	}
	CK(dalloc(ctx, &ctx->d_queries, n));
	CK(cudaMemcpyAsync(ctx->d_queries, qv.data(), n * sizeof(QueryView), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return ANDI_OK;
}

extern "C" int andi_pool_set_host(andi_ctx *ctx, const char *const *seqs, const size_t *lens, size_t n) {
	if (!ctx || !seqs) return ANDI_ERR_ARG;
	NvtxRange range("andi: pool upload + pack");
	int rc = pool_check(ctx, lens, n);
	if (rc) return rc;
	CK(cudaSetDevice(ctx->device));
	pool_release(ctx);
	ctx->n = n;
	ctx->len.assign(lens, lens + n);
	// Measured on the B200 boxes of this pool (PCIe Gen5: 55 GB/s from pinned memory, 16 host cores):
	// uploading the characters and packing on the device costs 140 ms for 6.5 GB, packing on the
	// host first 210 ms -- so the host packer is opt-in (ANDI_B200_HOST_PACK=1: hosts with many
	// cores behind a slow link).
	const char *hp = getenv("ANDI_B200_HOST_PACK");
	if (hp && atoi(hp) == 1) {
		rc = pool_set_host_packed(ctx, seqs, lens, n);
		if (rc) pool_release(ctx);
		return rc;
	}
	std::vector<size_t> offs(n);
	size_t total = 0;
	for (size_t k = 0; k < n; k++) {
		offs[k] = total;
		total += (lens[k] + 15) & ~(size_t)15;
	}
	if (total > ctx->stage_cap) {
		dfree(ctx, ctx->stage_chars);
		ctx->stage_cap = 0;
		CK(dalloc(ctx, &ctx->stage_chars, total));
		ctx->stage_cap = total;
	}
	unsigned char *d_chars = ctx->stage_chars;
	// sequences that lie in host memory the way they will lie on the device (one buffer, 16-byte
	// stride rounding: what a caller with a pool buffer has) go up in ONE copy per run
	for (size_t k = 0; k < n;) {
		size_t e = k + 1;
		while (e < n && seqs[e] == seqs[k] + (offs[e] - offs[k]) && offs[e] - offs[k] < ((size_t)512 << 20)) e++;  // 121 vs 136 ms per 6.5 GB pool
		const size_t bytes = offs[e - 1] - offs[k] + lens[e - 1];
		CK(cudaMemcpyAsync(d_chars + offs[k], seqs[k], bytes, cudaMemcpyHostToDevice, ctx->stream));
		k = e;
	}
	ctx->st.h2d_bytes += total;
	rc = pool_finish(ctx, d_chars, offs);
	if (rc) pool_release(ctx);
	return rc;
}

extern "C" int andi_pool_set_device(andi_ctx *ctx, const char *d_chars, const size_t *offsets,
									const size_t *lens, size_t n) {
	if (!ctx || !d_chars || !offsets) return ANDI_ERR_ARG;
	int rc = pool_check(ctx, lens, n);
	if (rc) return rc;
	CK(cudaSetDevice(ctx->device));
	pool_release(ctx);
	ctx->n = n;
	ctx->len.assign(lens, lens + n);
	std::vector<size_t> offs(offsets, offsets + n);
	rc = pool_finish(ctx, (const unsigned char *)d_chars, offs);
	if (rc) pool_release(ctx);	// no half-set pool
	return rc;
}

extern "C" size_t andi_pool_size(const andi_ctx *ctx) { return ctx ? ctx->n : 0; }

extern "C" int andi_pool_info(const andi_ctx *ctx, size_t k, size_t *len, double *gc, int *has_separator) {
	if (!ctx || k >= ctx->n) return ANDI_ERR_ARG;
	if (len) *len = ctx->len[k];
	if (gc) *gc = ctx->gc[k];
	if (has_separator) *has_separator = ctx->has_sep[k];
	return ANDI_OK;
}

// ------------------------------------------------------------------ index construction

static TextView rs_view(const andi_esa *E) {
	TextView t;
	t.code = E->code, t.spec = E->spec, t.len = E->N, t.mid = E->n;
	return t;
}

static int choose_depth(u32 N, u32 threshold, unsigned long long query_bases = 0) {
	// Directory depth ~ log4(N): about one suffix per bucket; never deeper than the anchor
	// threshold (below it only the match LENGTH matters, see longest_match) nor than 14.
	// One level deeper (a quarter of a suffix per bucket: most lookups end at the directory
	// entry or at its only suffix, and whole warps skip the candidate phase) makes the walk 12 %
	// faster but every table pass of the build four times longer; that pays once a subject is
	// walked by about a gigabase of queries (measured on 3085 x 2.1 Mbp: K 11 -> 12 moves the walk
	// 7.43 -> 6.55 ms and the build 0.46 -> 0.74 ms per subject; K 13: 6.30 and 2.2 ms).
	int k = (int)floor(log((double)N) / log(4.0) + 0.5);
	if (query_bases >= 1000000000ULL) k += 1;
	if (const char *bias = getenv("ANDI_B200_DEPTH_BIAS")) k += atoi(bias);	 // experiments only
	k = std::max(k, 4);
	k = std::min(k, 14);
	if (threshold < (u32)k) k = (int)threshold;
	return k >= 2 ? k : 0;
}

#include "index_host.cuh"

extern "C" int andi_esa_build(andi_ctx *ctx, size_t subject, unsigned flags, andi_esa **out) {
	if (!ctx || !out || subject >= ctx->n) return ANDI_ERR_ARG;
	*out = nullptr;
	CK(cudaSetDevice(ctx->device));
	andi_esa *E = new andi_esa();
	E->ctx = ctx;
	E->n = (u32)ctx->len[subject];
	E->N = 2 * E->n + 1;
	E->has_sep = ctx->has_sep[subject] != 0;
	E->self = (u32)subject;
	size_t nw = plane_words(E->N);
	E->threshold = (u32)andi_threshold(0.025, ctx->gc[subject], E->N);
	E->K = choose_depth(E->N, E->threshold);
	if (esa_ensure(ctx, E)) {
		esa_release(E);
		delete E;
		return ANDI_ERR_NOMEM;
	}
	k_build_rs<<<nblocks(nw, 256), 256, 0, ctx->stream>>>(ctx->pool_code + ctx->word_off[subject],
														   ctx->pool_spec + ctx->word_off[subject], E->n, E->code,
														   E->spec, (u32)nw);
	ctx->st.esa_launches++;
	*out = E;
	// The directory depth is tied to the default threshold (p = 0.025); a walk that is given a
	// smaller threshold falls back to the generic search (see launch sites).
	int rc = build_index(ctx, E, flags);
	if (rc) {
		esa_release(E);
		delete E;
		*out = nullptr;
	}
	return rc;
}

extern "C" int andi_esa_build_rs(andi_ctx *ctx, const char *rs, size_t rs_len, unsigned flags, andi_esa **out) {
	if (!ctx || !rs || !out) return ANDI_ERR_ARG;
	*out = nullptr;
	if (rs_len == 0 || rs_len > (size_t)INT_MAX) {
		ctx->err = "bad RS length";
		return ANDI_ERR_TOO_LONG;
	}
	CK(cudaSetDevice(ctx->device));
	andi_esa *E = new andi_esa();
	E->ctx = ctx;
	E->N = (u32)rs_len;
	E->n = (u32)(rs_len / 2);
	size_t nw = plane_words(E->N);
	unsigned char *d_chars = nullptr;
	unsigned long long *d_cnt = nullptr;
	unsigned long long cnt[2] = {0, 0};
	cudaError_t e = cudaSuccess;
	if ((e = dalloc(ctx, &E->code, nw)) != cudaSuccess || (e = dalloc(ctx, &E->spec, nw)) != cudaSuccess ||
		(e = dalloc(ctx, &E->sep3, nw)) != cudaSuccess || (e = dalloc(ctx, &d_chars, rs_len)) != cudaSuccess || (e = dalloc(ctx, &d_cnt, 2)) != cudaSuccess) {
		ctx->err = std::string("device allocation failed: ") + cudaGetErrorString(e);
		esa_release(E);
		delete E;
		return ANDI_ERR_NOMEM;
	}
	E->cap_words = nw;
	cudaMemcpyAsync(d_chars, rs, rs_len, cudaMemcpyHostToDevice, ctx->stream);
	cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), ctx->stream);
	ctx->st.h2d_bytes += rs_len;
	k_pack_rs_bytes<<<nblocks(nw, 256), 256, 0, ctx->stream>>>(d_chars, E->N, E->code, E->spec, (u32)nw, d_cnt);
	ctx->st.esa_launches++;
	cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream);
	e = cudaStreamSynchronize(ctx->stream);
	dfree(ctx, d_chars), dfree(ctx, d_cnt);
	if (e != cudaSuccess) {
		ctx->err = std::string("pack RS: ") + cudaGetErrorString(e);
		esa_release(E);
		delete E;
		return ANDI_ERR_CUDA;
	}
	// A general RS string may hold '#' anywhere, several times or not at all; the SPEC path makes
	// no assumption about it, so use it unless the string has the canonical shape: no other
	// separator and exactly ONE '#', in the middle (the non-SPEC kernels know '#' by position only).
	E->has_sep = cnt[1] != 0 || cnt[0] != 1 || (rs_len % 2 == 0) || rs[rs_len / 2] != '#';
	// gc of the forward half for the default threshold
	size_t gcn = 0;
	for (size_t k = E->n + 1; k < rs_len; k++) gcn += (rs[k] == 'G' || rs[k] == 'C');
	double gc = E->n ? (double)gcn / (double)E->n : 0.5;
	E->threshold = (u32)andi_threshold(0.025, gc, E->N);
	E->K = choose_depth(E->N, E->threshold);
	int rc = build_index(ctx, E, flags);
	if (rc) {
		esa_release(E);
		delete E;
		return rc;
	}
	*out = E;
	return ANDI_OK;
}

extern "C" void andi_esa_free(andi_esa *E) {
	if (!E) return;
	cudaSetDevice(E->ctx->device);
	esa_release(E);
	delete E;
}

extern "C" int32_t andi_esa_len(const andi_esa *E) { return E ? (int32_t)E->N : 0; }

extern "C" int andi_esa_download(const andi_esa *E, int32_t *SA, int32_t *LCP, int32_t *CLD, char *FVC,
								 andi_lcp_inter *cache) {
	if (!E) return ANDI_ERR_ARG;
	andi_ctx *ctx = E->ctx;
	CK(cudaSetDevice(ctx->device));
	if ((CLD || FVC || cache) && !E->full) {
		ctx->err = "CLD/FVC/cache need ANDI_ESA_FULL";
		return ANDI_ERR_ARG;
	}
	cudaStream_t st = ctx->stream;
	const size_t N = E->N;
	if (SA) CK(cudaMemcpyAsync(SA, E->SA, N * 4, cudaMemcpyDeviceToHost, st));
	if (LCP) CK(cudaMemcpyAsync(LCP, E->LCP, (N + 1) * 4, cudaMemcpyDeviceToHost, st));
	if (CLD) CK(cudaMemcpyAsync(CLD, E->CLD, (N + 1) * 4, cudaMemcpyDeviceToHost, st));
	if (FVC) CK(cudaMemcpyAsync(FVC, E->FVC, N, cudaMemcpyDeviceToHost, st));
	if (cache) CK(cudaMemcpyAsync(cache, E->cache, sizeof(Inter) << 20, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	ctx->st.d2h_bytes += (SA ? N * 4 : 0) + (LCP ? (N + 1) * 4 : 0) + (CLD ? (N + 1) * 4 : 0) + (FVC ? N : 0) +
						 (cache ? (sizeof(Inter) << 20) : 0);
	return ANDI_OK;
}

static SubjectIndex subject_index(const andi_esa *E) {
	SubjectIndex S;
	S.rs = rs_view(E);
	S.SA = E->SA, S.LCP = E->LCP, S.dir = E->dir, S.plen = E->plen, S.fdir = E->fdir;
	S.K = E->K, S.threshold = E->threshold, S.self = E->self, S.has_sep = E->has_sep;
	S.qcode_base = nullptr, S.qcomp_base = nullptr, S.qspec_delta = 0;
	S.s_sep3 = E->sep3, S.qsep3_base = nullptr;
	return S;
}

// Upload + pack a list of host strings as temporary queries.
struct TempQueries {
	u64 *code = nullptr, *spec = nullptr;
	unsigned char *sep3 = nullptr;
	QueryView *d_views = nullptr;
	bool any_sep = false;
};

// Many short queries (get_match batches): one thread packs one whole query.
__global__ void k_pack_many(const unsigned char *__restrict__ chars, const size_t *__restrict__ coff,
							const u32 *__restrict__ lens, const size_t *__restrict__ woff, u32 nq,
							u64 *__restrict__ code, u64 *__restrict__ spec, unsigned long long *counters) {
	u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nq) return;
	const unsigned char *src = chars + coff[k];
	u32 n = lens[k], nw = n / 32 + 4, sep = 0;  // = plane_words(n)
	for (u32 w = 0; w < nw; w++) {
		u64 cw = 0, sw = 0;
		for (u32 d = 0; d < 32 && w * 32 + d < n; d++) {
			u32 c = src[w * 32 + d];
			bool nuc = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
			u32 v = c & 6u;
			v ^= v >> 1;
			v >>= 1;
			if (!nuc) v = 0, sep++;
			cw |= (u64)v << (2 * d);
			sw |= (u64)(!nuc) << (2 * d);
		}
		code[woff[k] + w] = cw;
		spec[woff[k] + w] = sw;
	}
	if (sep) atomicAdd(&counters[1], (unsigned long long)sep);
}

static int temp_queries(andi_ctx *ctx, const char *const *qs, const size_t *lens, size_t nq, TempQueries &T) {
	std::vector<size_t> coff(nq), woff(nq);
	size_t chars = 0, words = 0;
	for (size_t k = 0; k < nq; k++) {
		coff[k] = chars, woff[k] = words;
		chars += (lens[k] + 15) & ~(size_t)15;
		words += (plane_words(lens[k]) + 1) & ~(size_t)1;
	}
	unsigned char *d_chars = nullptr;
	unsigned long long *d_cnt = nullptr;
	CK(dalloc(ctx, &d_chars, chars));
	CK(dalloc(ctx, &d_cnt, 2));
	CK(dalloc(ctx, &T.code, words));
	CK(dalloc(ctx, &T.spec, words));
	CK(dalloc(ctx, &T.sep3, words));
	CK(dalloc(ctx, &T.d_views, nq));
	CK(cudaMemsetAsync(d_cnt, 0, 16, ctx->stream));
	std::vector<QueryView> qv(nq);
	for (size_t k = 0; k < nq; k++) {
		qv[k].t.code = T.code + woff[k], qv[k].t.spec = T.spec + woff[k];
		qv[k].t.len = (u32)lens[k], qv[k].t.mid = 0xffffffffu, qv[k].has_sep = 0;
	}
	if (nq <= 8) {
		for (size_t k = 0; k < nq; k++) {
			if (lens[k]) CK(cudaMemcpyAsync(d_chars + coff[k], qs[k], lens[k], cudaMemcpyHostToDevice, ctx->stream));
			u32 nw = (u32)plane_words(lens[k]);
			k_pack<<<nblocks(nw, 256), 256, 0, ctx->stream>>>(d_chars + coff[k], (u32)lens[k], T.code + woff[k],
															   T.spec + woff[k], nw, d_cnt);
		}
	} else {
		// gather on the host, one copy, one kernel
		std::vector<unsigned char> host(chars);
		std::vector<u32> l32(nq);
		for (size_t k = 0; k < nq; k++) {
			memcpy(host.data() + coff[k], qs[k], lens[k]);
			l32[k] = (u32)lens[k];
		}
		size_t *d_coff = nullptr, *d_woff = nullptr;
		u32 *d_len = nullptr;
		CK(dalloc(ctx, &d_coff, nq));
		CK(dalloc(ctx, &d_woff, nq));
		CK(dalloc(ctx, &d_len, nq));
		CK(cudaMemcpyAsync(d_chars, host.data(), chars, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(d_coff, coff.data(), nq * sizeof(size_t), cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(d_woff, woff.data(), nq * sizeof(size_t), cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(d_len, l32.data(), nq * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
		k_pack_many<<<nblocks(nq, 128), 128, 0, ctx->stream>>>(d_chars, d_coff, d_len, d_woff, (u32)nq, T.code, T.spec,
																d_cnt);
		CK(cudaStreamSynchronize(ctx->stream));	 // host staging buffers go out of scope
		dfree(ctx, d_coff), dfree(ctx, d_woff), dfree(ctx, d_len);
	}
	k_sep3<<<nblocks(words, 256), 256, 0, ctx->stream>>>(T.spec, words, T.sep3);
	ctx->st.h2d_bytes += chars;
	unsigned long long cnt[2];
	CK(cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(T.d_views, qv.data(), nq * sizeof(QueryView), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	T.any_sep = cnt[1] != 0;
	dfree(ctx, d_chars), dfree(ctx, d_cnt);
	return ANDI_OK;
}

static void temp_release(andi_ctx *ctx, TempQueries &T) {
	dfree(ctx, T.code), dfree(ctx, T.spec), dfree(ctx, T.sep3), dfree(ctx, T.d_views);
}

extern "C" int andi_esa_get_match(const andi_esa *E, const char *const *queries, const size_t *lens, size_t nq,
								  andi_lcp_inter *out) {
	if (!E || !queries || !lens || !out) return ANDI_ERR_ARG;
	if (nq == 0) return ANDI_OK;
	andi_ctx *ctx = E->ctx;
	CK(cudaSetDevice(ctx->device));
	TempQueries T;
	int rc = temp_queries(ctx, queries, lens, nq, T);
	if (rc) return rc;
	Inter *d_out = nullptr;
	CK(dalloc(ctx, &d_out, nq));
	SubjectIndex S = subject_index(E);
	if (E->has_sep || T.any_sep)
		k_get_match<true><<<nblocks(nq, 128), 128, 0, ctx->stream>>>(S, T.d_views, (u32)nq, d_out);
	else
		k_get_match<false><<<nblocks(nq, 128), 128, 0, ctx->stream>>>(S, T.d_views, (u32)nq, d_out);
	CK(cudaMemcpyAsync(out, d_out, nq * sizeof(Inter), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	dfree(ctx, d_out);
	temp_release(ctx, T);
	// l == -2 is k_get_match's mark for "the directory lookup and the full-range search disagree":
	// an internal error, never an answer
	for (size_t k = 0; k < nq; k++)
		if (out[k].l == -2) {
			ctx->err = "get_match: directory lookup and generic search disagree (internal error)";
			return ANDI_ERR_CUDA;
		}
	return ANDI_OK;
}

// ------------------------------------------------------------------ the walk (host side)
#include "walk_host.cuh"

// ------------------------------------------------------------------ several GPUs
#include "multi_host.cuh"

// ------------------------------------------------------------------ part B: reference symbols
#include "compat.cuh"
