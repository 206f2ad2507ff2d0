// belt_dev.cuh — device-side belt block function (STB 34.101.31) for sm_100a.
//
// Replaces the reference's E/D macros and T-tables (belt_block.c:121-195, :210-295).
// Two S-box policies:
//   BeltBigT  — 4 pre-rotated u32 T-tables (H5,H13,H21,H29 in the reference's naming),
//               each replicated once per shared-memory bank so that lane L only ever
//               touches bank L: every LDS is conflict-free regardless of the data.
//               Byte offset of table t, entry x, lane L:
//                   (t>>1)*65536 + x*256 + (t&1)*128 + L*4          (128 KiB total)
//               so one PRMT builds the address: byte1 = x, byte0 = (t&1)*128 + L*4.
//   BeltSmallT — one 1 KiB table of H[x] as u32; rotation done in the ALU. For kernels
//               where belt is a minor cost (belt-hash inside bign) and smem is scarce.
#pragma once
#include "common.cuh"

#define BELT_BIGT_BYTES (128 * 1024)

// One copy per translation unit (no relocatable device code): each .cu that includes this
// header uploads the S-box into its own copy through belt_upload_H().
__constant__ u8 c_beltH[256];
static inline u32 belt_upload_H(const u8 H[256])
{
	if (cudaMemcpyToSymbol(c_beltH, H, 256) != cudaSuccess)
		return b2g_check_launch("cudaMemcpyToSymbol(beltH)");
	return B2G_OK;
}

struct BeltKey { u32 k[8]; };

// ---------------------------------------------------------------- S-box policies
struct BeltBigT
{
	const u8* base;   // shared memory, BELT_BIGT_BYTES
	u32 la, lb;       // lane*4, lane*4 + 128

	__device__ __forceinline__ static void fill(u8* sm)
	{
		u32* w = reinterpret_cast<u32*>(sm);
		for (u32 i = threadIdx.x; i < 4u * 256u * 32u; i += blockDim.x)
		{
			const u32 L = i & 31u, x = (i >> 5) & 255u, t = i >> 13;
			const u32 v = rotl32((u32)c_beltH[x], 5 + 8 * (int)t);
			w[(((t >> 1) << 16) + (x << 8) + ((t & 1u) << 7) + (L << 2)) >> 2] = v;
		}
	}
	__device__ __forceinline__ BeltBigT(const u8* sm) : base(sm)
	{
		la = (threadIdx.x & 31u) << 2;
		lb = la + 128u;
	}
	// G_r with r = 5 + 8*T0: byte k of x goes through table (T0 + k) mod 4
	template <int T0> __device__ __forceinline__ u32 g(u32 x) const
	{
		u32 v[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			const int t = (T0 + k) & 3;
			const u32 addr = __byte_perm(x, (t & 1) ? lb : la, 0x5504 + (k << 4));
			v[k] = *reinterpret_cast<const u32*>(base + (t >> 1) * 65536 + addr);
		}
		return v[0] ^ v[1] ^ v[2] ^ v[3];
	}
};

// The small-table policies keep the table's 32-bit SHARED-WINDOW address, not a generic pointer, and read
// it with ld.shared: the policy object travels by value into out-of-line routines (belt_hash_words inside
// the bign kernels), where a generic pointer made every lookup a generic LD with 64-bit address arithmetic
// (IADD3 + IMAD.X per lookup; r02 SASS histogram: 672 LD + 504 IMAD.X per three block encryptions).
__device__ __forceinline__ u32 belt_lds(u32 saddr)
{
	u32 v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
	return v;
}
__device__ __forceinline__ u32 belt_saddr(const void* sm) { return (u32)__cvta_generic_to_shared(sm); }

struct BeltSmallT
{
	static constexpr int WORDS = 256;
	u32 tab;   // shared-window byte address of 256 words: tab[x] = H[x]

	__device__ __forceinline__ static void fill(u32* sm)
	{
		for (u32 i = threadIdx.x; i < 256u; i += blockDim.x)
			sm[i] = c_beltH[i];
	}
	__device__ __forceinline__ BeltSmallT(const u32* sm) : tab(belt_saddr(sm)) {}
	template <int T0> __device__ __forceinline__ u32 g(u32 x) const
	{
		const u32 v = belt_lds(tab + ((x & 255u) << 2)) | belt_lds(tab + (((x >> 8) & 255u) << 2)) << 8 |
			belt_lds(tab + (((x >> 16) & 255u) << 2)) << 16 | belt_lds(tab + ((x >> 24) << 2)) << 24;
		return rotl32(v, 5 + 8 * T0);
	}
};

// Four pre-rotated 1 KiB tables (the reference's H5 / H13 / H21 / H29, belt_block.c:121-195) without the
// per-bank replication of BeltBigT: 4 KiB of shared memory, 4 LDS + 3 XOR per G-box, bank conflicts as
// the data fall (the LSU pipe is idle in the kernels that use it: belt inside bign).
struct BeltT4
{
	// The table is a static __shared__ array of the policy itself: its shared-window address is a link-time
	// constant, so every lookup is one LDS with an immediate base (R + imm) — no base-register add per
	// lookup (r02 SASS: 304 IMAD.IADD per block encryption with a run-time base). Kernels still declare
	// `__shared__ u32 tab[SB::WORDS]` and pass it around for the other policies; here it is one word.
	static constexpr int WORDS = 1;
	__device__ __forceinline__ static u32* table()
	{
		__shared__ u32 t[1024];   // t[k * 256 + x] = rotl32(H[x], 5 + 8 k)
		return t;
	}
	__device__ __forceinline__ static void fill(u32*)
	{
		u32* t = table();
		for (u32 i = threadIdx.x; i < 1024u; i += blockDim.x)
			t[i] = rotl32((u32)c_beltH[i & 255u], 5 + 8 * (int)(i >> 8));
	}
	__device__ __forceinline__ BeltT4(const u32*) {}
	// G_r with r = 5 + 8*T0: byte k of x goes through table (T0 + k) mod 4
	template <int T0> __device__ __forceinline__ u32 g(u32 x) const
	{
		const u32* t = table();
		return t[((T0 + 0) & 3) * 256 + (x & 255u)] ^ t[((T0 + 1) & 3) * 256 + ((x >> 8) & 255u)] ^
			t[((T0 + 2) & 3) * 256 + ((x >> 16) & 255u)] ^ t[((T0 + 3) & 3) * 256 + (x >> 24)];
	}
};

// ---------------------------------------------------------------- block function
// One round (steps 2.1-2.9; belt_block.c:231-240). KI(j) = index of the j-th round key.
#define BELT_ROUND(S, a, b, c, d, k, i, KI)                      \
	{                                                            \
		b ^= S.template g<0>(a + k[KI(i, 0)]);                   \
		c ^= S.template g<2>(d + k[KI(i, 1)]);                   \
		a -= S.template g<1>(b + k[KI(i, 2)]);                   \
		const u32 e_ = S.template g<2>(b + c + k[KI(i, 3)]) ^ (u32)(i); \
		b += e_;                                                 \
		c -= e_;                                                 \
		d += S.template g<1>(c + k[KI(i, 4)]);                   \
		b ^= S.template g<2>(a + k[KI(i, 5)]);                   \
		c ^= S.template g<0>(d + k[KI(i, 6)]);                   \
	}
#define BELT_KE(i, j) ((7 * (i) - 7 + (j)) & 7)
#define BELT_KD(i, j) ((7 * (i) - 1 - (j)) & 7)

// E_K: (a,b,c,d) in, result returned in place (word order of the 128-bit block)
template <class SB> __device__ __forceinline__ void belt_encr(const SB& S, u32& a, u32& b, u32& c, u32& d, const u32 (&k)[8])
{
	// operand rotation per round is pure renaming (belt_block.c:258-269)
	BELT_ROUND(S, a, b, c, d, k, 1, BELT_KE)
	BELT_ROUND(S, b, d, a, c, k, 2, BELT_KE)
	BELT_ROUND(S, d, c, b, a, k, 3, BELT_KE)
	BELT_ROUND(S, c, a, d, b, k, 4, BELT_KE)
	BELT_ROUND(S, a, b, c, d, k, 5, BELT_KE)
	BELT_ROUND(S, b, d, a, c, k, 6, BELT_KE)
	BELT_ROUND(S, d, c, b, a, k, 7, BELT_KE)
	BELT_ROUND(S, c, a, d, b, k, 8, BELT_KE)
	const u32 ta = a, tb = b, tc = c, td = d;
	a = tb, b = td, c = ta, d = tc;
}

// D_K (belt_block.c:284-295)
template <class SB> __device__ __forceinline__ void belt_decr(const SB& S, u32& a, u32& b, u32& c, u32& d, const u32 (&k)[8])
{
	BELT_ROUND(S, a, b, c, d, k, 8, BELT_KD)
	BELT_ROUND(S, c, a, d, b, k, 7, BELT_KD)
	BELT_ROUND(S, d, c, b, a, k, 6, BELT_KD)
	BELT_ROUND(S, b, d, a, c, k, 5, BELT_KD)
	BELT_ROUND(S, a, b, c, d, k, 4, BELT_KD)
	BELT_ROUND(S, c, a, d, b, k, 3, BELT_KD)
	BELT_ROUND(S, d, c, b, a, k, 2, BELT_KD)
	BELT_ROUND(S, b, d, a, c, k, 1, BELT_KD)
	const u32 ta = a, tb = b, tc = c, td = d;
	a = tc, b = ta, c = td, d = tb;
}

// ---------------------------------------------------------------- belt-hash (belt_compr.c:27-87, belt_hash.c)
// sigma compression; s may be null (final block).
template <class SB> __device__ __forceinline__ void belt_compress(const SB& S, u32* s, u32 (&h)[8], const u32 (&X)[8])
{
	u32 t0 = h[0] ^ h[4], t1 = h[1] ^ h[5], t2 = h[2] ^ h[6], t3 = h[3] ^ h[7];
	u32 S0 = t0, S1 = t1, S2 = t2, S3 = t3;
	belt_encr(S, S0, S1, S2, S3, X);
	S0 ^= t0, S1 ^= t1, S2 ^= t2, S3 ^= t3;
	if (s)
		s[0] ^= S0, s[1] ^= S1, s[2] ^= S2, s[3] ^= S3;
	const u32 k1[8] = {S0, S1, S2, S3, h[4], h[5], h[6], h[7]};
	const u32 k2[8] = {~S0, ~S1, ~S2, ~S3, h[0], h[1], h[2], h[3]};
	u32 y0 = X[0], y1 = X[1], y2 = X[2], y3 = X[3];
	belt_encr(S, y0, y1, y2, y3, k1);
	u32 z0 = X[4], z1 = X[5], z2 = X[6], z3 = X[7];
	belt_encr(S, z0, z1, z2, z3, k2);
	h[0] = y0 ^ X[0], h[1] = y1 ^ X[1], h[2] = y2 ^ X[2], h[3] = y3 ^ X[3];
	h[4] = z0 ^ X[4], h[5] = z1 ^ X[5], h[6] = z2 ^ X[6], h[7] = z3 ^ X[7];
}

__device__ __forceinline__ void belt_hash_init(u32 (&h)[8])
{
	const u32* H32 = reinterpret_cast<const u32*>(c_beltH);
#pragma unroll
	for (int i = 0; i < 8; ++i)
		h[i] = H32[i];
}

// Hash of a message held as zero-padded 32-bit words (nwords32 = 8 * ceil(len/32)).
// `msg` may live in local or global memory. out = 8 words (32 octets).
struct belt_digest { u32 w[8]; };
// (result by value: callers keep no address-taken output buffer, cf. sc256 in bign.cu)
template <class SB> __device__ __noinline__ belt_digest belt_hash_words(const SB S, const u32* msg, u32 len_bytes)
{
	belt_digest out;
	u32 h[8], ls[8] = {0, 0, 0, 0, 0, 0, 0, 0}, X[8];
	belt_hash_init(h);
	const u32 nblk = (len_bytes + 31u) >> 5;
#pragma unroll 1
	for (u32 i = 0; i < nblk; ++i)
	{
#pragma unroll
		for (int j = 0; j < 8; ++j)
			X[j] = msg[8 * i + j];
		belt_compress(S, ls + 4, h, X);
	}
	ls[0] = len_bytes << 3, ls[1] = len_bytes >> 29;
	belt_compress(S, (u32*)0, h, ls);
#pragma unroll
	for (int j = 0; j < 8; ++j)
		out.w[j] = h[j];
	return out;
}
