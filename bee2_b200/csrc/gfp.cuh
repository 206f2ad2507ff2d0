// gfp.cuh — GF(p) for the three bign fields, p = 2^(32 N) - c:
//   N =  8: p = 2^256 - 189 (bign-curve256v1, bign_params.c:36-41)
//   N = 12: p = 2^384 - 317 (bign-curve384v1, bign_params.c:78-85)
//   N = 16: p = 2^512 - 569 (bign-curve512v1, bign_params.c:131-140)
// N x 32-bit limbs in registers, one field element per thread.
//
// Replaces the reference's zmMulCrand/zmSqrCrand (zm.c:214-253) = zzMul/zzSqr (zz_mul.c:82-159)
// + zzRedCrand (zz_red.c:71-105), zzAddMod/zzSubMod (zz_mod.c:42,120) and gfpInv (gfp.c:33-44).
// Like the reference we keep PLAIN residues and use the Crandall fold 2^(32N) = c (mod p), not
// Montgomery form: the fold costs N extra wide multiply-adds instead of N^2.
//
// Representation: "weak" residues — any value in [0, 2^(32N)) congruent to the element;
// fe_canon() brings it to [0, p) where a unique form matters (comparisons, output).
//
// Multiplication: schoolbook N x N as N^2 wide multiply-adds (mad.lo.cc/madc.hi.cc pairs that
// ptxas fuses into IMAD.WIDE with carry) split into two accumulators — products landing on even
// limb positions and products landing on odd ones — so that every carry chain runs over aligned
// 64-bit columns; the two accumulators are summed once at the end.
// Squaring: the N(N-1)/2 off-diagonal products once, doubled by a funnel shift, plus the N
// diagonal squares added by one carry chain: N(N+1)/2 wide multiply-adds.
//
// The integer pipe that executes IMAD.WIDE also executes IADD3/SHF/LOP3 (measured: they do not
// co-issue, profiles/README.md), so every carry-propagation instruction costs as much as a
// product. The final "+ carry * c" of an addition / reduction therefore only touches limb 0 and
// ripples further in a rare branch (probability ~ c / 2^32 per operation).
#pragma once
#include "gfp_asm.cuh"
#include "gfp_inv.cuh"

template <int N> struct fe { u32 v[N]; };

// c = 2^(32 N) - p
template <int N> struct fe_param;
template <> struct fe_param<8> { static constexpr u32 C = 189u; };
template <> struct fe_param<12> { static constexpr u32 C = 317u; };
template <> struct fe_param<16> { static constexpr u32 C = 569u; };

GFP_HD u32 gfp_funnel_l(u32 lo, u32 hi, int s)   // (hi:lo << s) >> 32, 0 < s < 32
{
#ifdef __CUDA_ARCH__
	return __funnelshift_l(lo, hi, s);
#else
	return (hi << s) | (lo >> (32 - s));
#endif
}

// ---------------------------------------------------------------- folding small values in
// t[0..N) += x for a small x (< 2^22), modulo p (weak). Only limb 0 is touched unless it wraps
// (rare): then the carry ripples up; if it leaves the top limb the value is now < x and a
// further + c cannot carry.
template <int N> GFP_HD void fe_fold_in(u32* t, u32 x)
{
	const u32 t0 = t[0] + x;
	t[0] = t0;
	if (t0 < x)
	{
		if (inc_n<N - 1>(t + 1, 1u))
			t[0] += fe_param<N>::C;
	}
}
// t[0..N) -= x for a small x, modulo p (weak): the mirror image
template <int N> GFP_HD void fe_fold_out(u32* t, u32 x)
{
	const u32 t0 = t[0];
	t[0] = t0 - x;
	if (t0 < x)
	{
		if (dec_n<N - 1>(t + 1, 1u))
			t[0] -= fe_param<N>::C;
	}
}

// even-position off-diagonal rows of the squaring, in carry-limb order (see fe_sqr_wide):
// step I (even, 2 <= I <= N-2) takes the even row I (if it exists) and the odd row I-1
template <int N, int I> struct fe_sqr_rows
{
	static GFP_HD void run(u32* ev, const u32* a)
	{
		if constexpr (I <= N - 4)
			mad_row<(N - 2 - I) / 2>(ev + 2 * I + 2, a + I + 2, a[I]);
		mad_row<(N - I) / 2>(ev + 2 * I, a + I + 1, a[I - 1]);
		if constexpr (I + 2 <= N - 2)
			fe_sqr_rows<N, I + 2>::run(ev, a);
	}
};

// ---------------------------------------------------------------- products
// t[0..2N) = a * b
template <int N> GFP_HD void fe_mul_wide(u32* t, const u32* a, const u32* b)
{
	constexpr int H = N / 2;
	// ev[k] sits at limb k, od[k] at limb k+1
	u32 ev[2 * N + 1], od[2 * N];
#pragma unroll
	for (int i = 0; i < 2 * N + 1; ++i) ev[i] = 0;
#pragma unroll
	for (int i = 0; i < 2 * N; ++i) od[i] = 0;
#pragma unroll
	for (int i = 0; i < N; i += 2)
	{
		// b[i], i even: even a-limbs land on even positions, odd ones on odd positions
		mad_row<H>(ev + i, a, b[i]);
		mad_row<H>(od + i, a + 1, b[i]);
		// b[i+1]: the other way round
		mad_row<H>(od + i, a, b[i + 1]);
		if (i + 1 < N - 1)
			mad_row<H>(ev + i + 2, a + 1, b[i + 1]);
		else
			mad_row_top<H>(ev + i + 2, a + 1, b[i + 1]);
	}
	// t = ev + (od << 32); the sum is a*b < 2^(64N), so nothing leaves limb 2N-1
	t[0] = ev[0];
	const u32 c = add_n<N>(t + 1, ev + 1, od);
	add_n_cin<N - 1>(t + N + 1, ev + N + 1, od + N, c);
}

// t[0..2N) = a^2
template <int N> GFP_HD void fe_sqr_wide(u32* t, const u32* a)
{
	constexpr int H = N / 2;
	u32 ev[2 * N], od[2 * N];
#pragma unroll
	for (int i = 0; i < 2 * N; ++i) ev[i] = 0, od[i] = 0;
	// products on odd positions: every (even i, odd j) pair exactly once
#pragma unroll
	for (int i = 0; i < N; i += 2)
		mad_row<H>(od + i, a + 1, a[i]);
	// products on even positions, i < j of equal parity: row(i) = a[i] * (a[i+2], a[i+4], ...)
	// starts at limb 2i+2 and drops its carry at limb i+N (i even) / i+N+1 (i odd). Rows are
	// taken in the order of that carry limb — e0, e2, o1, e4, o3, ... — so that it only ever
	// holds earlier carries (every earlier row is < 2^(32 * carry limb)) and cannot overflow.
	mad_row<H - 1>(ev + 2, a + 2, a[0]);
	fe_sqr_rows<N, 2>::run(ev, a);
	// t = 2 (ev + (od << 32)) + diagonal
	u32 s[2 * N];
	s[0] = ev[0];
	const u32 c = add_n<N>(s + 1, ev + 1, od);
	add_n_cin<N - 1>(s + N + 1, ev + N + 1, od + N, c);
#pragma unroll
	for (int k = 2 * N - 1; k > 0; --k)
		t[k] = gfp_funnel_l(s[k - 1], s[k], 1);
	t[0] = s[0] << 1;
	mad_diag<N>(t, a);
}

// r = t mod p (weak), t < 2^(64N): fold hi * c into lo twice
template <int N> GFP_HD void fe_reduce_wide(u32* r, const u32* t)
{
	constexpr int H = N / 2;
	constexpr u32 C = fe_param<N>::C;
	u32 acc[N + 1], od[N + 1], s[N + 1];
#pragma unroll
	for (int i = 0; i < N; ++i) acc[i] = t[i], od[i] = 0;
	acc[N] = 0, od[N] = 0;
	mad_row<H>(acc, t + N, C);       // positions 0, 2, ...
	mad_row<H>(od, t + N + 1, C);    // positions 1, 3, ... (od[k] at limb k+1)
	// s[1..N] = acc[1..N] + od[0..N-1]; od[N] = 0 (products < 2^42 cannot reach it)
	s[0] = acc[0];
	(void)add_n<N>(s + 1, acc + 1, od);
	// second fold: s[N] <= c + 1, so s[N] * c < 2^21
#pragma unroll
	for (int i = 0; i < N; ++i) r[i] = s[i];
	fe_fold_in<N>(r, s[N] * C);
}

// Two forms of every product:
//   fe_mul_i / fe_sqr_i — inlined bodies;
//   fe_mul / fe_sqr     — a call to one shared copy (by-value register ABI), used by the point
//                         formulas. With all ~100 call sites inlined the verify kernel was
//                         350 KB of SASS and stalled on instruction fetch (ncu r01).
template <int N> GFP_HD void fe_mul_i(fe<N>& r, const fe<N>& a, const fe<N>& b)
{
	u32 t[2 * N];
	fe_mul_wide<N>(t, a.v, b.v);
	fe_reduce_wide<N>(r.v, t);
}
template <int N> GFP_HD void fe_sqr_i(fe<N>& r, const fe<N>& a)
{
	u32 t[2 * N];
	fe_sqr_wide<N>(t, a.v);
	fe_reduce_wide<N>(r.v, t);
}
#ifndef FE_MUL_INLINE
template <int N> __host__ __device__ __noinline__ fe<N> fe_mul_fn(const fe<N> a, const fe<N> b)
{
	fe<N> r;
	fe_mul_i<N>(r, a, b);
	return r;
}
template <int N> __host__ __device__ __noinline__ fe<N> fe_sqr_fn(const fe<N> a)
{
	fe<N> r;
	fe_sqr_i<N>(r, a);
	return r;
}
template <int N> GFP_HD void fe_mul(fe<N>& r, const fe<N>& a, const fe<N>& b) { r = fe_mul_fn<N>(a, b); }
template <int N> GFP_HD void fe_sqr(fe<N>& r, const fe<N>& a) { r = fe_sqr_fn<N>(a); }
#else
template <int N> GFP_HD void fe_mul(fe<N>& r, const fe<N>& a, const fe<N>& b) { fe_mul_i<N>(r, a, b); }
template <int N> GFP_HD void fe_sqr(fe<N>& r, const fe<N>& a) { fe_sqr_i<N>(r, a); }
#endif

// ---------------------------------------------------------------- linear operations
// r = a + b (weak)
template <int N> GFP_HD void fe_add(fe<N>& r, const fe<N>& a, const fe<N>& b)
{
	u32 t[N];
	const u32 c = add_n<N>(t, a.v, b.v);
	// fold the carry: 2^(32N) = c (mod p)
	fe_fold_in<N>(t, c * fe_param<N>::C);
#pragma unroll
	for (int i = 0; i < N; ++i) r.v[i] = t[i];
}

// r = a - b (weak)
template <int N> GFP_HD void fe_sub(fe<N>& r, const fe<N>& a, const fe<N>& b)
{
	u32 t[N];
	const u32 m = sub_n<N>(t, a.v, b.v);
	// a borrow means t = a - b + 2^(32N) = a - b + c (mod p): take c back; if that borrows again
	// the value wrapped to >= 2^(32N) - c and a further - c cannot borrow
	fe_fold_out<N>(t, m & fe_param<N>::C);
#pragma unroll
	for (int i = 0; i < N; ++i) r.v[i] = t[i];
}

// r = a - b - c (weak): two borrow chains, ONE fold — every borrow stands for + 2^(32N) = + c (mod p), both are taken
// back together (a fold is six instructions with its rare-ripple guard; the point formulas are full of a - b - c)
template <int N> GFP_HD void fe_sub2(fe<N>& r, const fe<N>& a, const fe<N>& b, const fe<N>& c)
{
	u32 t[N], u[N];
	const u32 m1 = sub_n<N>(t, a.v, b.v);
	const u32 m2 = sub_n<N>(u, t, c.v);
	fe_fold_out<N>(u, ((m1 & 1u) + (m2 & 1u)) * fe_param<N>::C);
#pragma unroll
	for (int i = 0; i < N; ++i) r.v[i] = u[i];
}

// r = 3 a (weak): 2a by a funnel-shift pass, + a by one carry chain, ONE fold for the bit shifted out and the carry
template <int N> GFP_HD void fe_mul3(fe<N>& r, const fe<N>& a)
{
	u32 d[N], t[N];
	const u32 top = a.v[N - 1] >> 31;
#pragma unroll
	for (int k = N - 1; k > 0; --k)
		d[k] = gfp_funnel_l(a.v[k - 1], a.v[k], 1);
	d[0] = a.v[0] << 1;
	const u32 c = add_n<N>(t, d, a.v);
	fe_fold_in<N>(t, (top + c) * fe_param<N>::C);
#pragma unroll
	for (int i = 0; i < N; ++i) r.v[i] = t[i];
}

// r = 2^K a (weak), K = 1, 2, 3: one funnel-shift pass, the K bits shifted out fold back as * c
template <int K, int N> GFP_HD void fe_shl(fe<N>& r, const fe<N>& a)
{
	u32 t[N];
	const u32 top = a.v[N - 1] >> (32 - K);
#pragma unroll
	for (int k = N - 1; k > 0; --k)
		t[k] = gfp_funnel_l(a.v[k - 1], a.v[k], K);
	t[0] = a.v[0] << K;
	fe_fold_in<N>(t, top * fe_param<N>::C);
#pragma unroll
	for (int i = 0; i < N; ++i) r.v[i] = t[i];
}
template <int N> GFP_HD void fe_dbl(fe<N>& r, const fe<N>& a) { fe_shl<1, N>(r, a); }

// canonical form in [0, p)
template <int N> GFP_HD void fe_canon(fe<N>& a)
{
	// a >= p  <=>  a + c >= 2^(32N), and then a - p = a + c - 2^(32N)
	u32 t[N];
#pragma unroll
	for (int i = 0; i < N; ++i) t[i] = a.v[i];
	if (inc_n<N>(t, fe_param<N>::C))
	{
#pragma unroll
		for (int i = 0; i < N; ++i) a.v[i] = t[i];
	}
}

template <int N> GFP_HD bool fe_is_zero(const fe<N>& a)
{
	// 0 has two weak forms: 0 and p = 2^(32N) - c
	u32 z = a.v[0], f = a.v[0] ^ (0u - fe_param<N>::C);
#pragma unroll
	for (int i = 1; i < N; ++i) z |= a.v[i], f |= ~a.v[i];
	return z == 0 || f == 0;
}

template <int N> GFP_HD void fe_set_u32(fe<N>& r, u32 x)
{
	r.v[0] = x;
#pragma unroll
	for (int k = 1; k < N; ++k) r.v[k] = 0;
}

// raw compare of two N-limb numbers (little-endian limbs): a >= b ?
template <int N> GFP_HD bool uN_geq(const u32* a, const u32* b)
{
	u32 t[N];
	return sub_n<N>(t, a, b) == 0;
}
template <int N> GFP_HD bool uN_is_zero(const u32* a)
{
	u32 z = a[0];
#pragma unroll
	for (int i = 1; i < N; ++i) z |= a[i];
	return z == 0;
}

// r = a^(2^n) by n squarings
template <int N> GFP_HD void fe_sqr_n(fe<N>& r, const fe<N>& a, int n)
{
	r = a;
#pragma unroll 1
	for (int i = 0; i < n; ++i)
		fe_sqr<N>(r, r);
}

// r = a^(p-2) = 1/a (gfp.c:33-44 computes the same power with a sliding window).
// p - 2 = 2^(32N) - (c + 2): the top 32N - 16 bits are ones, the low 16 bits are
// 0x10000 - (c + 2) = 0xFF41 / 0xFEC1 / 0xFDC5. An addition chain on runs of ones builds
// a^(2^(32N-16) - 1), then the low 16 bits are appended bit by bit.
template <int N> __host__ __device__ __noinline__ fe<N> fe_inv_fermat_fn(const fe<N> a)
{
	fe<N> x2, x4, x8, x16, x32, x48, x64, x128, t;
	fe_sqr<N>(t, a), fe_mul<N>(x2, t, a);                  // 2^2 - 1
	fe_sqr_n<N>(t, x2, 2), fe_mul<N>(x4, t, x2);           // 2^4 - 1
	fe_sqr_n<N>(t, x4, 4), fe_mul<N>(x8, t, x4);           // 2^8 - 1
	fe_sqr_n<N>(t, x8, 8), fe_mul<N>(x16, t, x8);          // 2^16 - 1
	fe_sqr_n<N>(t, x16, 16), fe_mul<N>(x32, t, x16);       // 2^32 - 1
	fe_sqr_n<N>(t, x32, 16), fe_mul<N>(x48, t, x16);       // 2^48 - 1
	fe_sqr_n<N>(t, x32, 32), fe_mul<N>(x64, t, x32);       // 2^64 - 1
	fe_sqr_n<N>(t, x64, 64), fe_mul<N>(x128, t, x64);      // 2^128 - 1
	// t = a^(2^(32N-16) - 1): 32N - 16 = 240 / 368 / 496
	if (N == 8)
	{
		fe_sqr_n<N>(t, x128, 64), fe_mul<N>(t, t, x64);    // 2^192 - 1
		fe_sqr_n<N>(t, t, 48), fe_mul<N>(t, t, x48);       // 2^240 - 1
	}
	else if (N == 12)
	{
		fe_sqr_n<N>(t, x128, 128), fe_mul<N>(t, t, x128);  // 2^256 - 1
		fe_sqr_n<N>(t, t, 64), fe_mul<N>(t, t, x64);       // 2^320 - 1
		fe_sqr_n<N>(t, t, 48), fe_mul<N>(t, t, x48);       // 2^368 - 1
	}
	else
	{
		fe_sqr_n<N>(t, x128, 128), fe_mul<N>(t, t, x128);  // 2^256 - 1
		fe_sqr_n<N>(t, t, 128), fe_mul<N>(t, t, x128);     // 2^384 - 1
		fe_sqr_n<N>(t, t, 64), fe_mul<N>(t, t, x64);       // 2^448 - 1
		fe_sqr_n<N>(t, t, 48), fe_mul<N>(t, t, x48);       // 2^496 - 1
	}
	// append the low 16 bits of p - 2, most significant first
	const u32 low = 0x10000u - (fe_param<N>::C + 2u);
#pragma unroll 1
	for (int bit = 15; bit >= 0; --bit)
	{
		fe_sqr<N>(t, t);
		if ((low >> bit) & 1u)
			fe_mul<N>(t, t, a);
	}
	return t;
}
// r = 1/a (0 -> 0), canonical. The kernels use the division-step form (gfp_inv.cuh: a third of the power's
// instructions and short dependent chains — one warp per CTA runs it while the others wait); the power
// stays as the cross-check of the host tests. CT = true: fixed step count (secret-dependent inputs).
template <int N, bool CT> __host__ __device__ __noinline__ fe<N> fe_inv_fn(const fe<N> a)
{
	fe<N> r, c = a;
	fe_canon<N>(c);   // the other weak form of 0 (p itself) must also give 0
	inv_safegcd<N, CT>(r.v, c.v, fe_param<N>::C);
	return r;
}
#ifdef FE_INV_FERMAT   /* A/B measurement only */
template <int N, bool CT = true> GFP_HD void fe_inv(fe<N>& r, const fe<N>& a) { r = fe_inv_fermat_fn<N>(a); }
#else
template <int N, bool CT = true> GFP_HD void fe_inv(fe<N>& r, const fe<N>& a) { r = fe_inv_fn<N, CT>(a); }
#endif

// little-endian octets <-> limbs (unaligned-safe)
template <int N> GFP_HD void fe_load(fe<N>& r, const u8* p)
{
	if (((uintptr_t)p & 3) == 0)
	{
#pragma unroll
		for (int k = 0; k < N; ++k) r.v[k] = reinterpret_cast<const u32*>(p)[k];
	}
	else
	{
#pragma unroll
		for (int k = 0; k < N; ++k)
			r.v[k] = (u32)p[4 * k] | (u32)p[4 * k + 1] << 8 | (u32)p[4 * k + 2] << 16 | (u32)p[4 * k + 3] << 24;
	}
}
template <int N> GFP_HD void fe_store(u8* p, const fe<N>& a)
{
	if (((uintptr_t)p & 3) == 0)
	{
#pragma unroll
		for (int k = 0; k < N; ++k) reinterpret_cast<u32*>(p)[k] = a.v[k];
	}
	else
	{
#pragma unroll
		for (int k = 0; k < N; ++k)
			for (int b = 0; b < 4; ++b) p[4 * k + b] = (u8)(a.v[k] >> (8 * b));
	}
}
