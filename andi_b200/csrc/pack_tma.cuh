// andi_b200/csrc/pack_tma.cuh -- pool packing with TMA-staged input.
//
// k_pack of esa_kernels.cuh reads its 32 input bytes per thread straight from global memory
// (32-byte strides inside a warp: ~350 GB/s). Here the bytes of a whole pool stream through
// shared memory instead: every CTA owns 8 KiB tiles, one elected thread issues a 1-D bulk
// tensor copy (cp.async.bulk global -> shared, completion counted on an mbarrier), two tiles in
// flight, and the 256 threads pack 32 bytes each with byte-SIMD arithmetic (4 characters per
// 32-bit operation). One launch for the whole pool; the < 8 KiB tails go through k_pack_tails.
//
// Requirements of cp.async.bulk: 16-byte aligned source and a size that is a multiple of 16 --
// full tiles of 16-aligned sequences satisfy both; anything else takes the plain kernel.
#pragma once
#include "esa_kernels.cuh"

#define ANDI_PACK_TILE 8192u

struct PackSeq {
	unsigned long long char_off;  // offset of the sequence in the char buffer
	unsigned long long word_off;  // offset of its planes in the pool (u64 words)
	u32 len;					  // characters
	u32 tile0;					  // index of its first full tile in the global tile numbering
};

__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, u32 count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, u32 bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@!p bra WAIT_%=;\n\t}" ::"r"(smem_addr(bar)),
		"r"(parity)
		: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, unsigned long long *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
				 "l"(src), "r"(bytes), "r"(smem_addr(bar))
				 : "memory");
}

// Four input bytes -> 8 bits of 2-bit codes, 8 bits of spec pairs; returns G+C and separator counts.
__device__ __forceinline__ void pack4(u32 w, u32 &code8, u32 &spec8, u32 &gc, u32 &sep) {
	// A 0x41, C 0x43, G 0x47, T 0x54: bits 1..2 give 0,1,3,2 -> xor with bit 2 gives 0,1,2,3
	u32 c4 = ((w >> 1) & 0x03030303u) ^ ((w >> 2) & 0x01010101u);
	u32 is_t = __vcmpeq4(w, 0x54545454u);
	u32 is_acg = __vcmpeq4(w | 0x06060606u, 0x47474747u) & ~__vcmpeq4(w, 0x45454545u);	// excludes 'E'
	u32 nuc = is_t | is_acg;  // 0xff per nucleotide byte
	c4 &= nuc;
	u32 s4 = ~nuc & 0x01010101u;
	// gather byte k into bit pair k: (b0 + b1<<8 + b2<<16 + b3<<24) * (2^24+2^18+2^12+2^6) >> 24
	code8 = (c4 * 0x01041040u) >> 24;
	spec8 = (s4 * 0x01041040u) >> 24;
	gc += __popc((c4 ^ (c4 >> 1)) & 0x01010101u);
	sep += __popc(s4);
}

__global__ void __launch_bounds__(256)
k_pack_tma(const unsigned char *__restrict__ chars, const PackSeq *__restrict__ seqs, u32 nseq, u32 ntiles,
		   u64 *__restrict__ code, u64 *__restrict__ spec, unsigned long long *__restrict__ counters) {
	__shared__ __align__(128) unsigned char buf[2][ANDI_PACK_TILE];
	__shared__ __align__(8) unsigned long long bar[2];
	__shared__ u32 s_seq[2];
	const u32 tid = threadIdx.x;
	if (tid == 0) {
		mbar_init(&bar[0], 1);
		mbar_init(&bar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	// tile -> sequence: last sequence whose first tile is <= tile (sequences without full tiles
	// share their tile0 with the next one and are never selected because of the "last" rule
	// combined with the tile count check below)
	auto issue = [&](u32 stage, u32 tile) {
		u32 lo = 0, hi = nseq;
		while (hi - lo > 1) {
			u32 m = (lo + hi) >> 1;
			if (seqs[m].tile0 <= tile)
				lo = m;
			else
				hi = m;
		}
		s_seq[stage] = lo;
		const unsigned char *src = chars + seqs[lo].char_off + (unsigned long long)(tile - seqs[lo].tile0) * ANDI_PACK_TILE;
		mbar_expect_tx(&bar[stage], ANDI_PACK_TILE);
		tma_load_1d(buf[stage], src, ANDI_PACK_TILE, &bar[stage]);
	};

	u32 tile = blockIdx.x, stage = 0, parity0 = 0, parity1 = 0;
	if (tid == 0 && tile < ntiles) issue(0, tile);
	for (; tile < ntiles; tile += gridDim.x) {
		u32 next = tile + gridDim.x;
		if (tid == 0 && next < ntiles) issue(stage ^ 1u, next);
		mbar_wait(&bar[stage], stage ? parity1 : parity0);
		if (stage) parity1 ^= 1u; else parity0 ^= 1u;
		const u32 k = s_seq[stage];
		const uint4 *in = reinterpret_cast<const uint4 *>(buf[stage] + tid * 32u);
		uint4 v0 = in[0], v1 = in[1];
		u32 w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
		u64 cw = 0, sw = 0;
		u32 gc = 0, sep = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			u32 c8, s8;
			pack4(w[j], c8, s8, gc, sep);
			cw |= (u64)c8 << (8 * j);
			sw |= (u64)s8 << (8 * j);
		}
		unsigned long long word = seqs[k].word_off + (unsigned long long)(tile - seqs[k].tile0) * (ANDI_PACK_TILE / 32u) + tid;
		code[word] = cw;
		spec[word] = sw;
		// warp-reduce the counters, one atomic pair per warp
		for (int o = 16; o; o >>= 1) {
			gc += __shfl_down_sync(0xffffffffu, gc, o);
			sep += __shfl_down_sync(0xffffffffu, sep, o);
		}
		if ((tid & 31u) == 0) {
			if (gc) atomicAdd(&counters[2 * k], (unsigned long long)gc);
			if (sep) atomicAdd(&counters[2 * k + 1], (unsigned long long)sep);
		}
		__syncthreads();  // buf[stage] is free again (it is refilled in the next iteration)
		stage ^= 1u;
	}
}

// Everything behind the last full tile of every sequence, guard words included: one CTA per
// sequence, plain loads (at most 8 KiB + guard per sequence).
__global__ void __launch_bounds__(256)
k_pack_tails(const unsigned char *__restrict__ chars, const PackSeq *__restrict__ seqs, const u32 *__restrict__ full_tiles,
			 u64 *__restrict__ code, u64 *__restrict__ spec, unsigned long long *__restrict__ counters) {
	const u32 k = blockIdx.x;
	const PackSeq s = seqs[k];
	const u32 first_word = full_tiles[k] * (ANDI_PACK_TILE / 32u);
	const u32 nwords = s.len / 32u + 4u;  // = plane_words(len)
	u32 gc = 0, sep = 0;
	for (u32 w = first_word + threadIdx.x; w < nwords; w += blockDim.x) {
		u64 cw = 0, sw = 0;
		u32 base = w * 32u;
		if (base < s.len) {
			u32 cnt = min(32u, s.len - base);
			const unsigned char *src = chars + s.char_off + base;
			for (u32 d = 0; d < cnt; d++) {
				u32 c = src[d];
				bool nuc = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
				u32 v = c & 6u;
				v ^= v >> 1;
				v >>= 1;
				if (!nuc) v = 0;
				cw |= (u64)v << (2 * d);
				sw |= (u64)(!nuc) << (2 * d);
				gc += nuc & ((v == 1) | (v == 2));
				sep += !nuc;
			}
		}
		code[s.word_off + w] = cw;
		spec[s.word_off + w] = sw;
	}
	if (gc) atomicAdd(&counters[2 * k], (unsigned long long)gc);
	if (sep) atomicAdd(&counters[2 * k + 1], (unsigned long long)sep);
}
