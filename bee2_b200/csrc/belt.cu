// belt.cu — belt-CTR / belt-ECB / belt-hash batch kernels for sm_100a + C-ABI launchers.
//
// Replaces beltCTRStepE's block loop (belt_ctr.c:85-97), beltECBStepE/D's block loop
// (belt_ecb.c:68-75, :92-99) and beltHash (belt_hash.c) on batches.
//
// Work decomposition: one 16-octet block per thread per step, BELT_ILP independent
// blocks in flight per thread, persistent grid of one 1024-thread CTA per SM (the
// 128 KiB bank-replicated T-tables are built once per CTA), grid-stride over blocks
// so that each warp stores 512 contiguous octets (4 full 128-B lines) per STG.128.
#include "belt_dev.cuh"
#include "gf128.cuh"

#define BELT_THREADS 1024
#define BELT_ILP 2

// ---------------------------------------------------------------- CTR
// counter block for keystream block n (n = first_block + j): s + n + 1 mod 2^128
// (belt_ctr.c:27-35: 128-bit little-endian increment applied before each encryption)
__device__ __forceinline__ uint4 ctr_block(const uint4 s, u64 n)
{
	const u64 inc = n + 1;              // callers keep first_block + nblocks < 2^64
	const u64 lo = ((u64)s.y << 32 | s.x) + inc;
	const u64 hi = ((u64)s.w << 32 | s.z) + (lo < inc ? 1 : 0);
	return make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
}

template <bool XOR_SRC>
__global__ void __launch_bounds__(BELT_THREADS, 1)
belt_ctr_kernel(uint4* dst, const uint4* src, u64 nblocks, u32 tail,
	const BeltKey key, const uint4 s, u64 first)
{
	extern __shared__ __align__(1024) u8 sm[];
	BeltBigT::fill(sm);
	__syncthreads();
	const BeltBigT S(sm);
	const u64 step = (u64)gridDim.x * BELT_THREADS;
	// blocks that can be stored as whole uint4 (a ragged last block is handled apart)
	const u64 nfull = tail ? nblocks - 1 : nblocks;
	for (u64 j0 = (u64)blockIdx.x * BELT_THREADS + threadIdx.x; j0 < nblocks; j0 += step * BELT_ILP)
	{
		uint4 v[BELT_ILP];
#pragma unroll
		for (int u = 0; u < BELT_ILP; ++u)
			v[u] = ctr_block(s, first + j0 + u * step);
#pragma unroll
		for (int u = 0; u < BELT_ILP; ++u)
			belt_encr(S, v[u].x, v[u].y, v[u].z, v[u].w, key.k);
#pragma unroll
		for (int u = 0; u < BELT_ILP; ++u)
		{
			const u64 j = j0 + u * step;
			if (j < nfull)
			{
				if (XOR_SRC)
				{
					const uint4 d = ldg_stream(src + j);
					v[u].x ^= d.x, v[u].y ^= d.y, v[u].z ^= d.z, v[u].w ^= d.w;
				}
				stg_stream(dst + j, v[u]);
			}
			else if (j < nblocks)
			{
				// ragged tail: `tail` octets of the last keystream block (belt_ctr.c:99-110)
				const u32 w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
				u8* d8 = reinterpret_cast<u8*>(dst + j);
				const u8* s8 = reinterpret_cast<const u8*>(src + j);
				for (u32 i = 0; i < tail; ++i)
				{
					u8 g = (u8)(w[i >> 2] >> (8 * (i & 3)));
					if (XOR_SRC)
						g ^= s8[i];
					d8[i] = g;
				}
			}
		}
	}
}

// ---------------------------------------------------------------- CHE keystream
// belt-CHE encrypts with E_K over the LFSR counter s_j = s_(j-1) x ^ 1 (belt_che.c:75-110). Each
// warp owns a span of 32*iters consecutive blocks; lane L takes blocks span + 32 i + L, so a
// warp stores 512 contiguous octets per step and a lane advances its counter by x^32 (a word
// shift) between steps. The start counter of every lane comes from the closed form in gf128.cuh.
template <bool XOR_SRC>
__global__ void __launch_bounds__(BELT_THREADS, 1)
belt_che_kernel(uint4* dst, const uint4* src, u64 nblocks, u32 tail, const BeltKey key, const uint4 s0,
	u64 first, u64 iters)
{
	extern __shared__ __align__(1024) u8 sm[];
	BeltBigT::fill(sm);
	__syncthreads();
	const BeltBigT S(sm);
	const u64 warp = ((u64)blockIdx.x * BELT_THREADS + threadIdx.x) >> 5;
	const u32 lane = threadIdx.x & 31u;
	const u64 span = warp * 32 * iters;
	if (span >= nblocks)
		return;
	const u64 nfull = tail ? nblocks - 1 : nblocks;
	gf128 s;
	{
		const gf128 s0g = {{s0.x, s0.y, s0.z, s0.w}};
		s = che_counter(s0g, first + span + lane + 1);   // gamma block j uses s_(j+1), j 0-based
	}
#pragma unroll 1
	for (u64 i = 0; i < iters; ++i)
	{
		const u64 j = span + 32 * i + lane;
		if (span + 32 * i >= nblocks)
			break;
		uint4 v = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
		belt_encr(S, v.x, v.y, v.z, v.w, key.k);
		if (j < nfull)
		{
			if (XOR_SRC)
			{
				const uint4 d = ldg_stream(src + j);
				v.x ^= d.x, v.y ^= d.y, v.z ^= d.z, v.w ^= d.w;
			}
			stg_stream(dst + j, v);
		}
		else if (j < nblocks)
		{
			const u32 w[4] = {v.x, v.y, v.z, v.w};
			u8* d8 = reinterpret_cast<u8*>(dst + j);
			const u8* s8 = reinterpret_cast<const u8*>(src + j);
			for (u32 k = 0; k < tail; ++k)
			{
				u8 g = (u8)(w[k >> 2] >> (8 * (k & 3)));
				if (XOR_SRC)
					g ^= s8[k];
				d8[k] = g;
			}
		}
		s = che_step32(s);
	}
}

// ---------------------------------------------------------------- ECB
// MULTIKEY: block j is encrypted under its own 32-octet key keys[2j..2j+1] (config 5).
template <bool DEC, bool MULTIKEY>
__global__ void __launch_bounds__(BELT_THREADS, 1)
belt_ecb_kernel(uint4* dst, const uint4* src, const uint4* __restrict__ keys, u64 nblocks,
	const BeltKey key, const bool keys32 = false)
{
	extern __shared__ __align__(1024) u8 sm[];
	BeltBigT::fill(sm);
	__syncthreads();
	const BeltBigT S(sm);
	const u64 step = (u64)gridDim.x * BELT_THREADS;
	for (u64 j0 = (u64)blockIdx.x * BELT_THREADS + threadIdx.x; j0 < nblocks; j0 += step * BELT_ILP)
	{
		uint4 v[BELT_ILP];
		u32 k[BELT_ILP][8];
#pragma unroll
		for (int u = 0; u < BELT_ILP; ++u)
		{
			const u64 j = j0 + u * step;
			if (j < nblocks)
			{
				v[u] = ldg_stream(src + j);
				if (MULTIKEY)
				{
					// one 256-bit request per thread when the key array is 32-byte aligned: each 32-byte
					// sector is fetched once (two 128-bit loads at stride 32 fetch every sector twice)
					uint4 k0, k1;
					if (keys32)
						ldg_stream256(keys + 2 * j, k0, k1);
					else
						k0 = ldg_stream(keys + 2 * j), k1 = ldg_stream(keys + 2 * j + 1);
					k[u][0] = k0.x, k[u][1] = k0.y, k[u][2] = k0.z, k[u][3] = k0.w;
					k[u][4] = k1.x, k[u][5] = k1.y, k[u][6] = k1.z, k[u][7] = k1.w;
				}
			}
			else
				v[u] = make_uint4(0, 0, 0, 0);
			if (!MULTIKEY || j >= nblocks)
			{
#pragma unroll
				for (int i = 0; i < 8; ++i)
					k[u][i] = key.k[i];
			}
		}
#pragma unroll
		for (int u = 0; u < BELT_ILP; ++u)
		{
			if (DEC)
				belt_decr(S, v[u].x, v[u].y, v[u].z, v[u].w, k[u]);
			else
				belt_encr(S, v[u].x, v[u].y, v[u].z, v[u].w, k[u]);
		}
#pragma unroll
		for (int u = 0; u < BELT_ILP; ++u)
		{
			const u64 j = j0 + u * step;
			if (j < nblocks)
				stg_stream(dst + j, v[u]);
		}
	}
}

// S-box policy of the batch belt-hash kernel (-DBELT_HASH_SBOX=BeltSmallT for the 1 KiB single table)
#ifndef BELT_HASH_SBOX
#define BELT_HASH_SBOX BeltT4
#endif
// ---------------------------------------------------------------- belt-hash batch
// One message per thread; messages of equal length msg_len at msgs + i*stride.
template <class SB> __device__ __forceinline__ void belt_hash_msg(const SB& S, u8* __restrict__ o, const u8* __restrict__ m,
	u64 msg_len, bool al4)
{
	u32 h[8], ls[8] = {0, 0, 0, 0, 0, 0, 0, 0}, X[8];
	belt_hash_init(h);
	u64 off = 0;
#pragma unroll 1
	for (; off < msg_len; off += 32)
	{
		if (al4 && off + 32 <= msg_len)
		{
			if ((reinterpret_cast<uintptr_t>(m) & 15) == 0)
			{
				// two 128-bit loads per block: every lane reads its own message, so a load instruction costs one LSU
				// wavefront per lane whatever its width — eight 32-bit loads were 256 wavefronts per block and
				// warp on top of the 672 of the S-box reads
				const uint4 lo = ldg_stream(reinterpret_cast<const uint4*>(m + off)), hi = ldg_stream(reinterpret_cast<const uint4*>(m + off) + 1);
				X[0] = lo.x, X[1] = lo.y, X[2] = lo.z, X[3] = lo.w, X[4] = hi.x, X[5] = hi.y, X[6] = hi.z, X[7] = hi.w;
			}
			else
			{
#pragma unroll
				for (int j = 0; j < 8; ++j)
					X[j] = reinterpret_cast<const u32*>(m + off)[j];
			}
		}
		else
		{
#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				u32 w = 0;
#pragma unroll
				for (int b = 0; b < 4; ++b)
				{
					const u64 p = off + 4 * j + b;
					if (p < msg_len)
						w |= (u32)m[p] << (8 * b);
				}
				X[j] = w;
			}
		}
		belt_compress(S, ls + 4, h, X);
	}
	// length in bits as a 128-bit LE integer (belt_lcl.c:25-48)
	ls[0] = (u32)(msg_len << 3), ls[1] = (u32)(msg_len >> 29), ls[2] = (u32)(msg_len >> 61), ls[3] = 0;
	belt_compress(S, (u32*)0, h, ls);
#pragma unroll
	for (int j = 0; j < 8; ++j)
		reinterpret_cast<u32*>(o)[j] = h[j];
}
// small batches: 256-thread CTAs with the four 1 KiB tables (random bank conflicts, cheap to set up)
__global__ void __launch_bounds__(256)
belt_hash_kernel(u8* __restrict__ hashes, const u8* __restrict__ msgs, u64 msg_len, u64 stride,
	u64 count)
{
	__shared__ u32 tab[BELT_HASH_SBOX::WORDS];
	BELT_HASH_SBOX::fill(tab);
	__syncthreads();
	const BELT_HASH_SBOX S(tab);
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	belt_hash_msg(S, hashes + 32 * i, msgs + i * stride, msg_len, (((uintptr_t)msgs | stride) & 3) == 0);
}
// large batches: the shape of the CTR kernel — one persistent 1024-thread CTA per SM with the bank-replicated
// 128 KiB tables (every S-box read conflict-free), grid-stride over the messages
__global__ void __launch_bounds__(BELT_THREADS, 1)
belt_hash_big_kernel(u8* __restrict__ hashes, const u8* __restrict__ msgs, u64 msg_len, u64 stride, u64 count)
{
	extern __shared__ __align__(1024) u8 sm[];
	BeltBigT::fill(sm);
	__syncthreads();
	const BeltBigT S(sm);
	const bool al4 = (((uintptr_t)msgs | stride) & 3) == 0;
	for (u64 i = (u64)blockIdx.x * BELT_THREADS + threadIdx.x; i < count; i += (u64)gridDim.x * BELT_THREADS)
		belt_hash_msg(S, hashes + 32 * i, msgs + i * stride, msg_len, al4);
}

// ---------------------------------------------------------------- launchers (C ABI)
static u32 belt_grid(u64 nblocks)
{
	const u64 sms = (u64)b2g_sm_count();
	const u64 want = (nblocks + (u64)BELT_THREADS * BELT_ILP - 1) / ((u64)BELT_THREADS * BELT_ILP);
	return (u32)(want < sms ? (want ? want : 1) : sms);
}

// opt every big-table kernel in to 128 KiB of dynamic shared memory (once, at bring-up)
static u32 belt_optin_all(void)
{
	const void* ks[] = {(const void*)belt_ctr_kernel<true>, (const void*)belt_ctr_kernel<false>,
		(const void*)belt_che_kernel<true>, (const void*)belt_che_kernel<false>,
		(const void*)belt_ecb_kernel<false, false>, (const void*)belt_ecb_kernel<true, false>,
		(const void*)belt_ecb_kernel<false, true>, (const void*)belt_hash_big_kernel};
	for (size_t i = 0; i < sizeof ks / sizeof ks[0]; ++i)
		if (cudaFuncSetAttribute(ks[i], cudaFuncAttributeMaxDynamicSharedMemorySize, BELT_BIGT_BYTES) != cudaSuccess)
			return b2g_check_launch("cudaFuncSetAttribute(belt)");
	return B2G_OK;
}

extern "C" u32 b2g_belt_upload_tables(const u8 H[256])
{
	const u32 e = belt_upload_H(H);
	return e ? e : belt_optin_all();
}

extern "C" u32 b2g_beltCTR_dev(void* d_dest, const void* d_src, size_t count, const u32 key[8],
	const u32 ctr0[4], u64 first_block, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (count == 0) return B2G_OK;
	if (((uintptr_t)d_dest & 15) || ((uintptr_t)d_src & 15)) return B2G_BAD_INPUT;
	const u64 nblocks = ((u64)count + 15) / 16;
	const u32 tail = (u32)(count & 15);
	BeltKey k;
	for (int i = 0; i < 8; ++i) k.k[i] = key[i];
	const uint4 s = make_uint4(ctr0[0], ctr0[1], ctr0[2], ctr0[3]);
	cudaStream_t st = (cudaStream_t)stream;
	if (d_src)
	{
		belt_ctr_kernel<true><<<belt_grid(nblocks), BELT_THREADS, BELT_BIGT_BYTES, st>>>(
			(uint4*)d_dest, (const uint4*)d_src, nblocks, tail, k, s, first_block);
	}
	else
	{
		belt_ctr_kernel<false><<<belt_grid(nblocks), BELT_THREADS, BELT_BIGT_BYTES, st>>>(
			(uint4*)d_dest, (const uint4*)0, nblocks, tail, k, s, first_block);
	}
	b2g_note_launch();
	return b2g_check_launch("belt_ctr_kernel");
}

// belt-CHE data pass: dest = src ^ gamma, gamma block j = E_K(s_(first_block + j + 1)), s_0 = E_K(iv)
extern "C" u32 b2g_beltCHE_dev(void* d_dest, const void* d_src, size_t count, const u32 key[8],
	const u32 s0[4], u64 first_block, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (count == 0) return B2G_OK;
	if (((uintptr_t)d_dest & 15) || ((uintptr_t)d_src & 15)) return B2G_BAD_INPUT;
	const u64 nblocks = ((u64)count + 15) / 16;
	const u32 tail = (u32)(count & 15);
	BeltKey k;
	for (int i = 0; i < 8; ++i) k.k[i] = key[i];
	const uint4 s = make_uint4(s0[0], s0[1], s0[2], s0[3]);
	// one span of 32*iters blocks per warp; all warps of the persistent grid busy, spans >= 64 steps
	const u64 warps_max = (u64)b2g_sm_count() * (BELT_THREADS / 32);
	u64 iters = (nblocks + 32 * warps_max - 1) / (32 * warps_max);
	if (iters < 64) iters = 64;
	const u64 nwarps = (nblocks + 32 * iters - 1) / (32 * iters);
	const u32 grid = (u32)((nwarps + BELT_THREADS / 32 - 1) / (BELT_THREADS / 32));
	cudaStream_t st = (cudaStream_t)stream;
	if (d_src)
		belt_che_kernel<true><<<grid, BELT_THREADS, BELT_BIGT_BYTES, st>>>((uint4*)d_dest, (const uint4*)d_src,
			nblocks, tail, k, s, first_block, iters);
	else
		belt_che_kernel<false><<<grid, BELT_THREADS, BELT_BIGT_BYTES, st>>>((uint4*)d_dest, (const uint4*)0,
			nblocks, tail, k, s, first_block, iters);
	b2g_note_launch();
	return b2g_check_launch("belt_che_kernel");
}

// A handful of blocks under one key (beltBlockEncr and friends, the E_K(iv) of beltCTRStart, ciphertext
// stealing): one small CTA per 128 blocks with the 4 KiB four-table S-box instead of a 1024-thread CTA that
// first fills the 128 KiB bank-replicated tables. Latency path, not a throughput path.
#define BELT_SMALL_BLOCKS 4096
template <bool DEC> __global__ void __launch_bounds__(128) belt_ecb_small_kernel(uint4* dst, const uint4* src,
	u32 nblocks, const BeltKey key)
{
	__shared__ u32 tab[BeltT4::WORDS];
	BeltT4::fill(tab);
	__syncthreads();
	const BeltT4 S(tab);
	const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= nblocks)
		return;
	uint4 v = src[j];
	if (DEC)
		belt_decr(S, v.x, v.y, v.z, v.w, key.k);
	else
		belt_encr(S, v.x, v.y, v.z, v.w, key.k);
	dst[j] = v;
}

extern "C" u32 b2g_beltECB_dev(void* d_dest, const void* d_src, size_t nblocks, const u32 key[8],
	int decrypt, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (nblocks == 0) return B2G_OK;
	if (((uintptr_t)d_dest & 15) || ((uintptr_t)d_src & 15)) return B2G_BAD_INPUT;
	BeltKey k;
	for (int i = 0; i < 8; ++i) k.k[i] = key[i];
	cudaStream_t st = (cudaStream_t)stream;
	if (nblocks <= BELT_SMALL_BLOCKS)
	{
		const u32 grid = (u32)((nblocks + 127) / 128);
		if (decrypt)
			belt_ecb_small_kernel<true><<<grid, 128, 0, st>>>((uint4*)d_dest, (const uint4*)d_src, (u32)nblocks, k);
		else
			belt_ecb_small_kernel<false><<<grid, 128, 0, st>>>((uint4*)d_dest, (const uint4*)d_src, (u32)nblocks, k);
		b2g_note_launch();
		return b2g_check_launch("belt_ecb_small_kernel");
	}
	if (decrypt)
	{
		belt_ecb_kernel<true, false><<<belt_grid(nblocks), BELT_THREADS, BELT_BIGT_BYTES, st>>>(
			(uint4*)d_dest, (const uint4*)d_src, (const uint4*)0, nblocks, k);
	}
	else
	{
		belt_ecb_kernel<false, false><<<belt_grid(nblocks), BELT_THREADS, BELT_BIGT_BYTES, st>>>(
			(uint4*)d_dest, (const uint4*)d_src, (const uint4*)0, nblocks, k);
	}
	b2g_note_launch();
	return b2g_check_launch("belt_ecb_kernel");
}

// block j of d_src under key j -> block j of d_dst (d_dst may be d_src, or another device's HBM mapped
// with b2g_ipc_open: the final gather of a sharded batch then happens inside this kernel)
extern "C" u32 b2g_beltECBEncrBatch2_dev(void* d_dst, const void* d_src, const void* d_keys32, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (count == 0) return B2G_OK;
	if (((uintptr_t)d_dst & 15) || ((uintptr_t)d_src & 15) || ((uintptr_t)d_keys32 & 15)) return B2G_BAD_INPUT;
	BeltKey k = {};
	belt_ecb_kernel<false, true><<<belt_grid(count), BELT_THREADS, BELT_BIGT_BYTES, (cudaStream_t)stream>>>(
		(uint4*)d_dst, (const uint4*)d_src, (const uint4*)d_keys32, count, k, ((uintptr_t)d_keys32 & 31) == 0);
	b2g_note_launch();
	return b2g_check_launch("belt_ecb_kernel<multikey>");
}
extern "C" u32 b2g_beltECBEncrBatch_dev(void* d_blocks, const void* d_keys32, size_t count, void* stream)
{
	return b2g_beltECBEncrBatch2_dev(d_blocks, d_blocks, d_keys32, count, stream);
}

// ---------------------------------------------------------------- belt-hash, streaming form
// One sponge-like chain (beltHashStepH / StepG, belt_hash.c:52-131): state = s (4 words) || h (8 words).
// Absorbs nblocks 32-octet blocks; if `final`, then compresses the length block len || s without
// touching s (belt_hash.c:118-120). Sequential by construction: one thread works.
__global__ void belt_hash_step_kernel(u32* __restrict__ state, const u8* __restrict__ data, u64 nblocks,
	u32 final, uint4 len)
{
	__shared__ u32 tab[256];
	BeltSmallT::fill(tab);
	__syncthreads();
	if (threadIdx.x != 0)
		return;
	const BeltSmallT S(tab);
	u32 s[4], h[8], X[8];
#pragma unroll
	for (int j = 0; j < 4; ++j) s[j] = state[j];
#pragma unroll
	for (int j = 0; j < 8; ++j) h[j] = state[4 + j];
#pragma unroll 1
	for (u64 i = 0; i < nblocks; ++i)
	{
		const u8* p = data + 32 * i;
#pragma unroll
		for (int j = 0; j < 8; ++j)
			X[j] = (u32)p[4 * j] | (u32)p[4 * j + 1] << 8 | (u32)p[4 * j + 2] << 16 | (u32)p[4 * j + 3] << 24;
		belt_compress(S, s, h, X);
	}
	if (final)
	{
		X[0] = len.x, X[1] = len.y, X[2] = len.z, X[3] = len.w;
#pragma unroll
		for (int j = 0; j < 4; ++j) X[4 + j] = s[j];
		belt_compress(S, (u32*)0, h, X);
	}
#pragma unroll
	for (int j = 0; j < 4; ++j) state[j] = s[j];
#pragma unroll
	for (int j = 0; j < 8; ++j) state[4 + j] = h[j];
}

extern "C" u32 b2g_beltHashStep_dev(void* d_state, const void* d_data, size_t nblocks, int final,
	const u32 len[4], void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if ((uintptr_t)d_state & 3) return B2G_BAD_INPUT;
	const uint4 l = make_uint4(len[0], len[1], len[2], len[3]);
	belt_hash_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((u32*)d_state, (const u8*)d_data, nblocks,
		final ? 1u : 0u, l);
	b2g_note_launch();
	return b2g_check_launch("belt_hash_step_kernel");
}

extern "C" u32 b2g_beltHashBatch_dev(void* d_hashes, const void* d_msgs, size_t msg_len, size_t stride,
	size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_hashes & 3) return B2G_BAD_INPUT;
	// from half a wave of the persistent shape on (and messages long enough to pay for filling the 128 KiB tables)
	if (count >= (size_t)b2g_sm_count() * BELT_THREADS / 2 && msg_len >= 64)
	{
		const u64 ctas = (count + BELT_THREADS - 1) / BELT_THREADS;
		const u32 grid = (u32)(ctas < (u64)b2g_sm_count() ? ctas : (u64)b2g_sm_count());
		belt_hash_big_kernel<<<grid, BELT_THREADS, BELT_BIGT_BYTES, (cudaStream_t)stream>>>((u8*)d_hashes, (const u8*)d_msgs,
			msg_len, stride, count);
		b2g_note_launch();
		return b2g_check_launch("belt_hash_big_kernel");
	}
	const u32 grid = (u32)((count + 255) / 256);
	belt_hash_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((u8*)d_hashes, (const u8*)d_msgs, msg_len, stride, count);
	b2g_note_launch();
	return b2g_check_launch("belt_hash_kernel");
}
