// andi_b200/csrc/experimental/walk_binned.cuh -- EXPERIMENTAL, NOT PART OF THE LIBRARY.
//
// Round-2 groundwork for the anchor walk: the phase-binned form of k_walk_chunks_fast<1,0>
// (RAW / JC / KIMURA counting, no separators). `make experimental` compiles it as a syntax /
// ptxas check; libandi_b200.so does not contain it (a development build with
// -DANDI_EXPERIMENTAL_BINNED launches it through ANDI_B200_WALK=binned). Status at the end of
// round 1: the phase logic passes tests/test_binned_emulation.py on the CPU and one GPU run gave
// rows equal to the oracle on five stress groups, but this untuned form is 3.5x slower than
// k_walk_chunks_fast -- see DESIGN.md section 4 ("Round-2 plan") for the numbers and the reasons.
//
// The measured problem of k_walk_chunks_fast (profiles/r1j_*): a warp executes every phase of a
// trip for the 10-18 of its 32 lanes that need it (13.6 lanes per instruction on average), and
// every per-instruction saving moved the result by <= 1 %. Here the lane <-> unit binding is
// given up. A CTA owns BIN_SLOTS units whose whole walk state lives in shared memory, and one
// queue per phase. The CTA advances in super-steps:
//
//     __syncthreads            the queues of this super-step are complete
//     every warp, repeatedly:  take 32 units off a queue (one shared atomic), run THAT phase for
//                              them with all lanes, append each unit to the queue of its next
//                              phase for the NEXT super-step
//     __syncthreads            swap the two queue sets
//
// so every phase body runs with full warps except for at most one ragged batch per queue and
// super-step. Two CTAs per SM overlap one CTA's memory round with the other's arithmetic. Units,
// records and results are those of walk_kernels.cuh: k_walk_reduce consumes the records as is.
//
// Phases (one memory round each):
//   FETCH   take a unit from the global dispenser, initialise the slot, BEGIN
//   CMP     one 64-base window of a compare (lucky diagonal or directory candidate)
//   DIR     directory view fdir: absent -> DECIDE, one suffix -> CMP, several -> CAND
//   CAND    SA[candidate] -> CMP
//   SLOW    generic search (rare) -> DECIDE
//   DECIDE  src/process.c:160-196: pairing, accounting, advance; then COLS or BEGIN
//   COLS    up to 16 gap columns (src/model.c:309-337); then COLS or BEGIN
// BEGIN (chunk / boundary-replay bookkeeping and the set-up of the lucky compare,
// src/process.c:86-99) is pure arithmetic and runs at the end of FETCH / DECIDE / COLS.
#pragma once
#include "../walk_kernels.cuh"
#include "../sa_bucket.cuh"

// ---- the primitives of walk_binned_phases.h on the device
#define BIN_FN __device__ __forceinline__
#define BIN_OUTLINE_FN __device__ __noinline__
__device__ __forceinline__ u64 bin_ld64(const u64 *p) { return __ldg(p); }
__device__ __forceinline__ u32 bin_ld32(const u32 *p) { return __ldg(p); }
__device__ __forceinline__ u32 bin_ffs64(u64 x) { return (u32)(__ffsll((long long)x) - 1); }
__device__ __forceinline__ u32 bin_ffs32(u32 x) { return (u32)(__ffs((int)x) - 1); }
__device__ __forceinline__ u32 bin_popc32(u32 x) { return (u32)__popc(x); }
__device__ __forceinline__ u32 bin_atomic_inc(u32 *p) { return atomicAdd(p, 1u); }
__device__ __forceinline__ unsigned long long bin_next_unit(unsigned long long *p) { return atomicAdd(p, 1ULL); }
__device__ __forceinline__ void bin_slow_lookup(const SubjectIndex &S, const u64 *q_code, u32 qlen, u32 pos, u32 &len,
												bool &unique, u32 &at) {
	TextView qv;
	qv.code = q_code, qv.spec = nullptr, qv.len = qlen, qv.mid = 0xffffffffu;
	MatchResult m = longest_match<false>(S, qv, pos, qlen - pos);
	len = m.len, unique = m.unique;
	at = (m.unique && m.found_pos) ? __ldg(S.SA + m.at) : 0u;
}
#include "walk_binned_phases.h"

__global__ void __launch_bounds__(ANDI_BIN_SLOTS, 2)
k_walk_binned(const SubjectIndex S, const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids, u32 nq,
			  u32 chunk, u32 cpq, u32 threshold, u32 *__restrict__ records, unsigned long long *__restrict__ next_unit) {
	extern __shared__ __align__(16) unsigned char bin_smem[];
	BinShared &sh = *reinterpret_cast<BinShared *>(bin_smem);
	const u32 tid = threadIdx.x, lane = tid & 31u;
	const unsigned long long total = (unsigned long long)nq * cpq;	// the launcher guarantees total < 2^32
	BinConst c;
	c.t = threshold, c.N = S.rs.len, c.mid = S.rs.mid, c.border = c.N / 2, c.chunk = chunk, c.cpq = cpq, c.K = S.K;

	// every slot starts in FETCH
	if (tid < BQ_N) sh.count[0][tid] = 0, sh.count[1][tid] = 0, sh.head[tid] = 0;
	__syncthreads();
	sh.queue[0][BQ_FETCH][tid] = (unsigned short)tid;
	if (tid == 0) sh.count[0][BQ_FETCH] = ANDI_BIN_SLOTS;
	u32 cur = 0, waited[BQ_N];
#pragma unroll
	for (u32 q = 0; q < BQ_N; q++) waited[q] = 0;

	for (;;) {
		__syncthreads();  // the queues of this super-step are complete
		u32 pending = 0;
#pragma unroll
		for (u32 q = 0; q < BQ_N; q++) pending += sh.count[cur][q];
		if (pending == 0) break;  // every slot ran out of units
		const u32 nxt = cur ^ 1u;

		for (u32 q = 0; q < BQ_N; q++) {
			const u32 have = sh.count[cur][q];
			if (bin_defer(have, waited[q], pending)) {	// a thin queue waits for a full batch (block-uniform decision)
				if (tid < have) bin_push(sh, nxt, q, sh.queue[cur][q][tid]);
				waited[q]++;
				continue;
			}
			waited[q] = 0;
			for (;;) {
				u32 base = 0;
				if (lane == 0) base = atomicAdd(&sh.head[q], 32u);
				base = __shfl_sync(0xffffffffu, base, 0);
				if (base >= have) break;
				if (base + lane >= have) continue;	// ragged last batch of this queue
				const u32 s = sh.queue[cur][q][base + lane];
				bin_phase(q, s, sh, nxt, S, queries, query_ids, c, total, records, next_unit);
			}
		}

		__syncthreads();  // every unit of this super-step has been moved to the next queue set
		if (tid < BQ_N) sh.count[cur][tid] = 0, sh.head[tid] = 0;
		cur = nxt;
	}
}

// Host side (what launch_walk of walk_host.cuh would do for this kernel).
static inline cudaError_t launch_walk_binned(const SubjectIndex &S, const QueryView *d_queries, const u32 *d_query_ids,
											 u32 nq, u32 chunk, u32 cpq, u32 threshold, u32 *d_records,
											 unsigned long long *d_counter, int sm_count, cudaStream_t stream) {
	if ((unsigned long long)nq * cpq >= 0xffffffffULL || !S.qcode_base || S.K <= 0) return cudaErrorInvalidValue;
	cudaError_t e = cudaFuncSetAttribute(k_walk_binned, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BinShared));
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_walk_binned, ANDI_BIN_SLOTS, sizeof(BinShared));
	if (e != cudaSuccess) return e;
	if (per_sm < 1) per_sm = 1;
	e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
	if (e != cudaSuccess) return e;
	k_walk_binned<<<(unsigned)(per_sm * sm_count), ANDI_BIN_SLOTS, sizeof(BinShared), stream>>>(
		S, d_queries, d_query_ids, nq, chunk, cpq, threshold, d_records, d_counter);
	return cudaGetLastError();
}
