// gf128.cuh — GF(2^128) = GF(2)[x]/(x^128 + x^7 + x^2 + x + 1) for the belt-DWP tag.
//
// Replaces beltPolyMul = ppMul + ppRedBelt (belt_lcl.c:119-132, pp_red.c:129-141). Bit i of the
// 128-bit little-endian integer (word 0 least significant) is the coefficient of x^i.
#pragma once
#include "common.cuh"

struct gf128 { u32 w[4]; };

__device__ __forceinline__ gf128 gf_xor(const gf128 a, const gf128 b)
{
	gf128 r;
#pragma unroll
	for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] ^ b.w[i];
	return r;
}

// a * x mod f
__device__ __forceinline__ gf128 gf_mulx(const gf128 a)
{
	gf128 r;
	const u32 top = a.w[3] >> 31;
	r.w[3] = __funnelshift_l(a.w[2], a.w[3], 1);
	r.w[2] = __funnelshift_l(a.w[1], a.w[2], 1);
	r.w[1] = __funnelshift_l(a.w[0], a.w[1], 1);
	r.w[0] = (a.w[0] << 1) ^ (top ? 0x87u : 0u);
	return r;
}

// generic product, bit-serial (used once per thread, for the chunk weight)
static __device__ __noinline__ gf128 gf_mul(const gf128 a, const gf128 b)
{
	gf128 r = {{0, 0, 0, 0}}, v = b;
#pragma unroll 1
	for (int i = 0; i < 128; ++i)
	{
		const u32 m = 0u - ((a.w[i >> 5] >> (i & 31)) & 1u);
#pragma unroll
		for (int k = 0; k < 4; ++k) r.w[k] ^= v.w[k] & m;
		v = gf_mulx(v);
	}
	return r;
}

// 32 bits -> 64 bits with a zero between neighbours (squaring in GF(2)[x] spreads the bits)
__device__ __forceinline__ u64 gf_spread(u32 x)
{
	u64 v = x;
	v = (v | v << 16) & 0x0000FFFF0000FFFFull;
	v = (v | v << 8) & 0x00FF00FF00FF00FFull;
	v = (v | v << 4) & 0x0F0F0F0F0F0F0F0Full;
	v = (v | v << 2) & 0x3333333333333333ull;
	v = (v | v << 1) & 0x5555555555555555ull;
	return v;
}

// a^2 mod f
__device__ __forceinline__ gf128 gf_sqr(const gf128 a)
{
	u64 p[4];
#pragma unroll
	for (int i = 0; i < 4; ++i) p[i] = gf_spread(a.w[i]);
	// p[3]:p[2] is the high half h; x^128 = x^7 + x^2 + x + 1, so fold h*(x^7 + x^2 + x + 1)
	const u64 h0 = p[2], h1 = p[3];
	// t = h * (x^7 + x^2 + x + 1) as a 192-bit value t2:t1:t0 (t2 < 2^7)
	const u64 t0 = h0 ^ h0 << 1 ^ h0 << 2 ^ h0 << 7;
	const u64 t1 = h1 ^ h1 << 1 ^ h1 << 2 ^ h1 << 7 ^ h0 >> 63 ^ h0 >> 62 ^ h0 >> 57;
	const u64 t2 = h1 >> 63 ^ h1 >> 62 ^ h1 >> 57;
	// second fold of the small overflow t2 * x^128
	const u64 lo0 = p[0] ^ t0 ^ t2 ^ t2 << 1 ^ t2 << 2 ^ t2 << 7;
	const u64 lo1 = p[1] ^ t1;
	gf128 r;
	r.w[0] = (u32)lo0, r.w[1] = (u32)(lo0 >> 32), r.w[2] = (u32)lo1, r.w[3] = (u32)(lo1 >> 32);
	return r;
}

// ---- multiplication by a FIXED r through 4-bit window tables in shared memory ----
// TAB[p][v] = (v * x^(4p)) * r, p = 0..31, v = 0..15, each entry replicated 8 times in one
// 128-byte row so that lane L reads copy L & 7: the 8 lanes of every LDS.128 quarter-warp phase
// touch 8 different 4-bank groups -> conflict-free for any data (64 KiB per CTA).
#define GF_TAB_BYTES (32 * 16 * 128)

// cooperative build by the whole CTA (blockDim.x threads); r must be CTA-uniform
__device__ __forceinline__ void gf_tab_build(u8* tab, const gf128 r)
{
	for (u32 e = threadIdx.x; e < 32u * 16u; e += blockDim.x)
	{
		const u32 p = e >> 4, v = e & 15u;
		gf128 b = r;                       // r * x^(4p)
		for (u32 i = 0; i < 4 * p; ++i) b = gf_mulx(b);
		gf128 acc = {{0, 0, 0, 0}};
#pragma unroll
		for (int bit = 0; bit < 4; ++bit)
		{
			if (v >> bit & 1u) acc = gf_xor(acc, b);
			b = gf_mulx(b);
		}
		uint4* row = reinterpret_cast<uint4*>(tab + (size_t)e * 128);
		const uint4 val = make_uint4(acc.w[0], acc.w[1], acc.w[2], acc.w[3]);
#pragma unroll
		for (int c = 0; c < 8; ++c) row[c] = val;
	}
}

// a * r via the tables
__device__ __forceinline__ gf128 gf_mul_tab(const u8* tab, const gf128 a)
{
	const u8* base = tab + ((threadIdx.x & 7u) << 4);
	u32 x0 = 0, x1 = 0, x2 = 0, x3 = 0;
#pragma unroll
	for (int p = 0; p < 32; ++p)
	{
		const u32 v = (a.w[p >> 3] >> (4 * (p & 7))) & 15u;
		const uint4 t = *reinterpret_cast<const uint4*>(base + ((size_t)(p * 16) << 7) + (v << 7));
		x0 ^= t.x, x1 ^= t.y, x2 ^= t.z, x3 ^= t.w;
	}
	gf128 r = {{x0, x1, x2, x3}};
	return r;
}

// r^e by square-and-multiply (squarings are bit spreads, multiplications by r use the tables)
__device__ __forceinline__ gf128 gf_pow_r(const u8* tab, u64 e)
{
	gf128 acc = {{1, 0, 0, 0}};
	if (e == 0) return acc;
	int top = 63 - __clzll((long long)e);
#pragma unroll 1
	for (int i = top; i >= 0; --i)
	{
		acc = gf_sqr(acc);
		if (e >> i & 1) acc = gf_mul_tab(tab, acc);
	}
	return acc;
}

// x^e mod f: squarings (bit spreads) and multiplications by x only
__device__ __forceinline__ gf128 gf_pow_x(u64 e)
{
	gf128 acc = {{1, 0, 0, 0}};
	if (e == 0) return acc;
	const int top = 63 - __clzll((long long)e);
#pragma unroll 1
	for (int i = top; i >= 0; --i)
	{
		acc = gf_sqr(acc);
		if (e >> i & 1) acc = gf_mulx(acc);
	}
	return acc;
}

// a * x^32 mod f: a word shift; the word shifted out (< 2^32) folds back as top * (x^7 + x^2 + x + 1)
__device__ __forceinline__ gf128 gf_mulx32(const gf128 a)
{
	const u32 t = a.w[3];
	gf128 r;
	r.w[0] = t ^ t << 1 ^ t << 2 ^ t << 7;
	r.w[1] = a.w[0] ^ t >> 31 ^ t >> 30 ^ t >> 25;
	r.w[2] = a.w[1];
	r.w[3] = a.w[2];
	return r;
}

// belt-CHE counter: s_j = s_(j-1) * x ^ 1 (belt_che.c:89, belt_lcl.c:99-108). Closed form used to
// start any thread anywhere: s_j = s_0 x^j ^ (x^j ^ 1) (x + 1)^(-1), and (x + 1)^(-1) =
// (f(x) + 1) / (x + 1) = x^127 + ... + x^7 + x because f(1) = 1.
__device__ __forceinline__ gf128 che_counter(const gf128 s0, u64 j)
{
	const gf128 inv_xp1 = {{0xFFFFFF82u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}};
	gf128 xj = gf_pow_x(j);
	gf128 a = gf_mul(s0, xj);
	xj.w[0] ^= 1u;
	return gf_xor(a, gf_mul(xj, inv_xp1));
}
// s -> s_(+32): 32 steps at once, s x^32 ^ (x^31 + ... + 1)
__device__ __forceinline__ gf128 che_step32(const gf128 s)
{
	gf128 r = gf_mulx32(s);
	r.w[0] ^= 0xFFFFFFFFu;
	return r;
}
