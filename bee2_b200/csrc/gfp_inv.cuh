// gfp_inv.cuh — field inversion by Bernstein–Yang division steps ("safegcd", 2019), one thread per element.
//
// Replaces the power a^(p-2) that gfpInv computes (gfp.c:33-44, qr.c power ladder): for p = 2^256 - 189 that
// is 255 squarings + 13 multiplications = 33 000 dependent-chain instructions, and it sits on the critical path
// of every CTA of the bign kernels (bign.cu block_inv: one warp inverts the 32 nodes at which the CTA's product
// tree ends while the other warps wait at a barrier). The division-step iteration needs ~12 000 instructions
// (l = 128) with short dependent chains.
//
//   divstep(delta, f, g) = (1 - delta, g, (g - f) / 2)          if delta > 0 and g odd
//                          (1 + delta, f, (g + (g mod 2) f) / 2) otherwise
// from (1, p, a); f stays odd, after enough steps g = 0 and f = +-gcd = +-1. Steps are taken 30 at a time on
// the low 32 bits of f and g only, which gives a 2x2 integer matrix t with t (f, g) = 2^30 (f', g'); the
// matrix is then applied to the full-size f, g (exactly) and to the cofactors d, e (modulo p, with the
// division by 2^30 made exact by adding a multiple of p), where d a = f and e a = g (mod p) throughout.
// Numbers are held in signed 30-bit limbs (top limb carries the sign) so that every row of a matrix product
// fits a 64-bit accumulator — the layout of libsecp256k1's modinv32, restated for the three bign fields.
//
// Step count: for f, g < 2^d the iteration ends within floor((49 d + 57) / 17) steps (Bernstein–Yang,
// Theorem 11.2, d >= 46): 741 / 1110 / 1479 for d = 256 / 384 / 512 -> 25 / 37 / 50 batches of 30.
// For 256-bit moduli the variant that starts from delta = 1/2 is known to end within 590 steps (the
// convex-hull bound libsecp256k1's modinv32 relies on: 20 batches of 30); it differs in one line — the
// update of eta = -(delta + 1/2) — and is what N = 8 uses. The wider fields keep delta = 1 and the theorem.
// CT = true always runs all batches (the instruction stream does not depend on a: secret-dependent inputs —
// the Z coordinates of k G in the signing kernels); CT = false stops at the first batch that ends with
// g = 0 (public inputs: verification). Either way the loop goes on while g != 0, so the RESULT never
// depends on a step bound being right — only the claim of a fixed instruction count does. (Batches the
// early-exit form needs on 20 000 random elements per field: 17-18 of 20 / 26-28 of 37 / 34-37 of 50;
// tests/test_host_gfp.py::test_inversion_step_bound_holds_on_samples.)
#pragma once
#include "common.cuh"

typedef int32_t i32;
typedef int64_t i64;

#ifndef GFP_INV_HD
#define GFP_INV_HD __host__ __device__ __forceinline__
#endif

template <int N> struct inv30
{
	static constexpr int L = (32 * N + 2 + 29) / 30;               // 9 / 13 / 18 limbs
	static constexpr bool HALF_DELTA = N == 8;                      // delta starts at 1/2 (256-bit bound: 590 steps)
	static constexpr int BATCHES = HALF_DELTA ? 20 : ((49 * 32 * N + 57) / 17 + 29) / 30;
	static constexpr i32 M30 = (i32)0x3FFFFFFF;
};

struct inv_mat { i32 u, v, q, r; };

// 30 division steps on the low bits; eta = -delta (HALF: eta = -(delta + 1/2)). Branch-free.
template <bool HALF> GFP_INV_HD i32 inv_divsteps30(i32 eta, u32 f0, u32 g0, inv_mat& t)
{
	u32 u = 1, v = 0, q = 0, r = 1;
	u32 f = f0, g = g0;
#pragma unroll 6
	for (int i = 0; i < 30; ++i)
	{
		const u32 c1 = (u32)(eta >> 31);    // delta > 0
		const u32 c2 = 0u - (g & 1u);       // g odd
		const u32 c3 = c1 & c2;             // swap case
		const u32 one = c3 & 1u;
		// g, q, r += (delta > 0 ? -(f, u, v) : (f, u, v)) if g is odd; -x = (x ^ ~0) + 1: one LOP3 + one IADD3 each
		g += ((f ^ c1) & c2) + one, q += ((u ^ c1) & c2) + one, r += ((v ^ c1) & c2) + one;
		// swap: delta <- 1 - delta, else delta <- 1 + delta
		eta = HALF ? (i32)(((u32)eta ^ c3) - 1u) : (i32)(((u32)eta ^ c3) - (c3 + 1u));
		f += g & c3, u += q & c3, v += r & c3;
		g >>= 1, u <<= 1, v <<= 1;
	}
	t.u = (i32)u, t.v = (i32)v, t.q = (i32)q, t.r = (i32)r;
	return eta;
}

// (f, g) <- t (f, g) / 2^30, exact
template <int L> GFP_INV_HD void inv_update_fg(i32* f, i32* g, const inv_mat& t)
{
	constexpr i32 M30 = (i32)0x3FFFFFFF;
	i64 cf = (i64)t.u * f[0] + (i64)t.v * g[0];
	i64 cg = (i64)t.q * f[0] + (i64)t.r * g[0];
	cf >>= 30, cg >>= 30;
#pragma unroll
	for (int i = 1; i < L; ++i)
	{
		cf += (i64)t.u * f[i] + (i64)t.v * g[i];
		cg += (i64)t.q * f[i] + (i64)t.r * g[i];
		f[i - 1] = (i32)cf & M30, cf >>= 30;
		g[i - 1] = (i32)cg & M30, cg >>= 30;
	}
	f[L - 1] = (i32)cf, g[L - 1] = (i32)cg;
}

// (d, e) <- t (d, e) / 2^30 mod p; d, e stay in (-2p, p). m = p in 30-bit limbs, minv = p^-1 mod 2^30
template <int L> GFP_INV_HD void inv_update_de(i32* d, i32* e, const inv_mat& t, const i32* m, u32 minv)
{
	constexpr i32 M30 = (i32)0x3FFFFFFF;
	const i32 sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
	// a first multiple of p that brings negative d, e up ...
	i32 md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
	i64 cd = (i64)t.u * d[0] + (i64)t.v * e[0];
	i64 ce = (i64)t.q * d[0] + (i64)t.r * e[0];
	// ... corrected so that the low 30 bits of t (d, e) + p (md, me) vanish
	md -= (i32)((minv * (u32)cd + (u32)md) & (u32)M30);
	me -= (i32)((minv * (u32)ce + (u32)me) & (u32)M30);
	cd += (i64)m[0] * md, ce += (i64)m[0] * me;
	cd >>= 30, ce >>= 30;
#pragma unroll
	for (int i = 1; i < L; ++i)
	{
		cd += (i64)t.u * d[i] + (i64)t.v * e[i];
		ce += (i64)t.q * d[i] + (i64)t.r * e[i];
		cd += (i64)m[i] * md, ce += (i64)m[i] * me;
		d[i - 1] = (i32)cd & M30, cd >>= 30;
		e[i - 1] = (i32)ce & M30, ce >>= 30;
	}
	d[L - 1] = (i32)cd, e[L - 1] = (i32)ce;
}

// r = 1/a mod p for p = 2^(32N) - c, a any N-limb value (0 -> 0); the result is canonical (< p)
template <int N, bool CT> GFP_INV_HD void inv_safegcd(u32* r, const u32* a, u32 c, int* batches_used = 0)
{
	constexpr int L = inv30<N>::L;
	constexpr i32 M30 = inv30<N>::M30;
	i32 m[L], f[L], g[L], d[L], e[L];
	// p and a in 30-bit limbs
	{
		u32 pw[N];
		pw[0] = 0u - c;
#pragma unroll
		for (int i = 1; i < N; ++i) pw[i] = 0xFFFFFFFFu;
#pragma unroll
		for (int i = 0; i < L; ++i)
		{
			const int bit = 30 * i, w = bit >> 5, s = bit & 31;
			u32 xp = w < N ? pw[w] >> s : 0u, xa = w < N ? a[w] >> s : 0u;
			if (s > 2 && w + 1 < N)
				xp |= pw[w + 1] << (32 - s), xa |= a[w + 1] << (32 - s);
			m[i] = (i32)(xp & (u32)M30), g[i] = (i32)(xa & (u32)M30);
			f[i] = m[i], d[i] = 0, e[i] = 0;
		}
		e[0] = 1;
	}
	// p^-1 mod 2^30 by Newton's iteration from p^-1 = p (mod 8)
	u32 minv = (u32)m[0];
#pragma unroll
	for (int i = 0; i < 4; ++i) minv *= 2u - (u32)m[0] * minv;
	minv &= (u32)M30;
	i32 eta = -1;
#pragma unroll 1
	for (int b = 0;; ++b)
	{
		inv_mat t;
		eta = inv_divsteps30<inv30<N>::HALF_DELTA>(eta, (u32)f[0] | (u32)f[1] << 30, (u32)g[0] | (u32)g[1] << 30, t);
		inv_update_de<L>(d, e, t, m, minv);
		inv_update_fg<L>(f, g, t);
		if (!CT || b >= inv30<N>::BATCHES - 1)
		{
			// CT: only looked at after the last scheduled batch (g = 0 by then: never another round)
			i32 z = 0;
#pragma unroll
			for (int i = 0; i < L; ++i) z |= g[i];
			if (z == 0)
			{
				if (batches_used)   // tests: how many batches the iteration needed (early-exit form)
					*batches_used = b + 1;
				break;
			}
		}
	}
	// now f = +-1 and d = +-1/a in (-2p, p): d <- d + p if negative; d <- -d if f < 0; d <- d + p if negative
	{
		const i32 sf = f[L - 1] >> 31;
		i32 s = d[L - 1] >> 31, cy = 0;
#pragma unroll
		for (int i = 0; i < L; ++i)
		{
			i32 x = d[i] + (m[i] & s);
			x = (x ^ sf) - sf;
			x += cy;
			cy = x >> 30, d[i] = i < L - 1 ? (x & M30) : x;
		}
		s = d[L - 1] >> 31, cy = 0;
#pragma unroll
		for (int i = 0; i < L; ++i)
		{
			i32 x = d[i] + (m[i] & s) + cy;
			cy = x >> 30, d[i] = i < L - 1 ? (x & M30) : x;
		}
	}
	// back to 32-bit limbs
#pragma unroll
	for (int w = 0; w < N; ++w)
	{
		const int bit = 32 * w, i = bit / 30, s = bit % 30;
		u32 x = (u32)d[i] >> s;
		if (i + 1 < L)
			x |= (u32)d[i + 1] << (30 - s);
		if (s > 28 && i + 2 < L)
			x |= (u32)d[i + 2] << (60 - s);
		r[w] = x;
	}
}
