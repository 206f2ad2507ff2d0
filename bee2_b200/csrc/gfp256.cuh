// gfp256.cuh — GF(p), p = 2^256 - 189 (bign-curve256v1), 8 x 32-bit limbs in registers.
//
// Replaces the reference's zmMulCrand/zmSqrCrand (zm.c:214-253) = zzMul/zzSqr (zz_mul.c:82-159)
// + zzRedCrand (zz_red.c:71-105), zzAddMod/zzSubMod (zz_mod.c:42,120) and gfpInv (gfp.c:33-44).
// Like the reference we keep PLAIN residues and use the Crandall fold 2^256 = 189 (mod p),
// not Montgomery form: the fold costs 8 extra wide multiply-adds instead of 64.
//
// Representation: "weak" residues — any value in [0, 2^256) congruent to the element;
// fe_canon() brings it to [0, p) where a unique form matters (comparisons, output).
//
// Multiplication: schoolbook 8x8 as 64 wide multiply-adds (mad.lo.cc/madc.hi.cc pairs that
// ptxas fuses into IMAD.WIDE with carry) split into two accumulators — products landing on
// even limb positions and products landing on odd ones — so that every carry chain runs
// over aligned 64-bit columns; the two accumulators are summed once at the end.
#pragma once
#include "common.cuh"

struct fe { u32 v[8]; };

#define FE_C 189u   // 2^256 - p

// acc[0..7] += {a0,a1,a2,a3} * b at 64-bit column steps; carry-out added into acc[8]
__device__ __forceinline__ void mad_row(u32* acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 b)
{
	asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
		"madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
		"madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
		"madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
		"madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
		"madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
		"madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
		"madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
		"addc.u32 %8, %8, 0;"
		: "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
		  "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
		: "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// same without the carry-out limb (when the bound on the partial sum proves it is zero)
__device__ __forceinline__ void mad_row_top(u32* acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 b)
{
	asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
		"madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
		"madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
		"madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
		"madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
		"madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
		"madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
		"madc.hi.u32 %7, %11, %12, %7;"
		: "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
		  "+r"(acc[6]), "+r"(acc[7])
		: "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// ---- carry chains: each chain is ONE asm statement so nothing can clobber CC.CF in between.
// Pure outputs are early-clobber ("=&r"): the statements write them before the last input is
// read, so they must never share a register with an input.
// r = a + b, returns carry (0/1)
__device__ __forceinline__ u32 add8(u32* r, const u32* a, const u32* b)
{
	u32 c;
	asm("add.cc.u32 %0, %9, %17;\n\t"
		"addc.cc.u32 %1, %10, %18;\n\t"
		"addc.cc.u32 %2, %11, %19;\n\t"
		"addc.cc.u32 %3, %12, %20;\n\t"
		"addc.cc.u32 %4, %13, %21;\n\t"
		"addc.cc.u32 %5, %14, %22;\n\t"
		"addc.cc.u32 %6, %15, %23;\n\t"
		"addc.cc.u32 %7, %16, %24;\n\t"
		"addc.u32 %8, 0, 0;"
		: "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
		: "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
		  "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
	return c;
}
// r = a - b, returns borrow mask (0 / 0xFFFFFFFF)
__device__ __forceinline__ u32 sub8(u32* r, const u32* a, const u32* b)
{
	u32 c;
	asm("sub.cc.u32 %0, %9, %17;\n\t"
		"subc.cc.u32 %1, %10, %18;\n\t"
		"subc.cc.u32 %2, %11, %19;\n\t"
		"subc.cc.u32 %3, %12, %20;\n\t"
		"subc.cc.u32 %4, %13, %21;\n\t"
		"subc.cc.u32 %5, %14, %22;\n\t"
		"subc.cc.u32 %6, %15, %23;\n\t"
		"subc.cc.u32 %7, %16, %24;\n\t"
		"subc.u32 %8, 0, 0;"
		: "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
		: "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
		  "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
	return c;
}
// r += x (one limb), returns carry
__device__ __forceinline__ u32 add8_u32(u32* r, u32 x)
{
	u32 c;
	asm("add.cc.u32 %0, %0, %9;\n\t"
		"addc.cc.u32 %1, %1, 0;\n\t"
		"addc.cc.u32 %2, %2, 0;\n\t"
		"addc.cc.u32 %3, %3, 0;\n\t"
		"addc.cc.u32 %4, %4, 0;\n\t"
		"addc.cc.u32 %5, %5, 0;\n\t"
		"addc.cc.u32 %6, %6, 0;\n\t"
		"addc.cc.u32 %7, %7, 0;\n\t"
		"addc.u32 %8, 0, 0;"
		: "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=&r"(c)
		: "r"(x));
	return c;
}
// r -= x (one limb), returns borrow mask
__device__ __forceinline__ u32 sub8_u32(u32* r, u32 x)
{
	u32 c;
	asm("sub.cc.u32 %0, %0, %9;\n\t"
		"subc.cc.u32 %1, %1, 0;\n\t"
		"subc.cc.u32 %2, %2, 0;\n\t"
		"subc.cc.u32 %3, %3, 0;\n\t"
		"subc.cc.u32 %4, %4, 0;\n\t"
		"subc.cc.u32 %5, %5, 0;\n\t"
		"subc.cc.u32 %6, %6, 0;\n\t"
		"subc.cc.u32 %7, %7, 0;\n\t"
		"subc.u32 %8, 0, 0;"
		: "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=&r"(c)
		: "r"(x));
	return c;
}

// r[0..6] = a[0..6] + b[0..6] + cin (cin = 0/1); the final carry is dropped (caller proves it 0)
__device__ __forceinline__ void add7_cin(u32* r, const u32* a, const u32* b, u32 cin)
{
	asm("{\n\t.reg .u32 t;\n\t"
		"add.cc.u32 t, %21, 0xFFFFFFFF;\n\t"   /* CF <- cin */
		"addc.cc.u32 %0, %7, %14;\n\t"
		"addc.cc.u32 %1, %8, %15;\n\t"
		"addc.cc.u32 %2, %9, %16;\n\t"
		"addc.cc.u32 %3, %10, %17;\n\t"
		"addc.cc.u32 %4, %11, %18;\n\t"
		"addc.cc.u32 %5, %12, %19;\n\t"
		"addc.u32 %6, %13, %20;\n\t}"
		: "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6])
		: "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]),
		  "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(cin));
}

// t[0..15] = a * b
__device__ __forceinline__ void fe_mul_wide(u32 (&t)[16], const fe& a, const fe& b)
{
	// ev[k] sits at limb k, od[k] at limb k+1
	u32 ev[17], od[16];
#pragma unroll
	for (int i = 0; i < 17; ++i) ev[i] = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) od[i] = 0;
#pragma unroll
	for (int i = 0; i < 8; i += 2)
	{
		// b[i], i even: even a-limbs land on even positions, odd ones on odd positions
		mad_row(ev + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
		mad_row(od + i, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
		// b[i+1]: the other way round
		mad_row(od + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1]);
		if (i + 1 < 7)
			mad_row(ev + i + 2, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
		else
			mad_row_top(ev + i + 2, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
	}
	// t = ev + (od << 32); the sum is a*b < 2^512, so nothing leaves limb 15
	t[0] = ev[0];
	const u32 c = add8(t + 1, ev + 1, od);
	add7_cin(t + 9, ev + 9, od + 8, c);
}

// r = t mod p (weak), t < 2^512: fold hi*189 into lo twice (2^256 = 189 mod p)
__device__ __forceinline__ void fe_reduce_wide(fe& r, const u32 (&t)[16])
{
	u32 acc[9], od[9], s[9];
#pragma unroll
	for (int i = 0; i < 8; ++i) acc[i] = t[i], od[i] = 0;
	acc[8] = 0, od[8] = 0;
	mad_row(acc, t[8], t[10], t[12], t[14], FE_C);   // positions 0,2,4,6
	mad_row(od, t[9], t[11], t[13], t[15], FE_C);    // positions 1,3,5,7 (od[k] at limb k+1)
	// s[1..8] = acc[1..8] + od[0..7]; od[8] = 0 (four products < 2^40 cannot reach it)
	s[0] = acc[0];
	(void)add8(s + 1, acc + 1, od);
	// second fold: s[8] <= 190, so s[8]*189 < 2^16
	const u32 c = add8_u32(s, s[8] * FE_C);
	// a carry here means the sum wrapped to a value < 2^16: one more +189 cannot carry
	s[0] += c * FE_C;
#pragma unroll
	for (int i = 0; i < 8; ++i) r.v[i] = s[i];
}

// Two forms of the multiplication:
//   fe_mul_i — inlined body (~130 SASS instructions), used inside the out-of-line point
//              operations of ecp256.cuh;
//   fe_mul   — a call to one shared copy (by-value register ABI), used everywhere else
//              (inversion chain, conversions).
// With every one of the ~100 call sites inlined the verify kernel was 350 KB of SASS and
// stalled on instruction fetch (ncu r01: 34 % "no instruction").
__device__ __forceinline__ void fe_mul_i(fe& r, const fe& a, const fe& b)
{
	u32 t[16];
	fe_mul_wide(t, a, b);
	fe_reduce_wide(r, t);
}
__device__ __forceinline__ void fe_sqr_i(fe& r, const fe& a) { fe_mul_i(r, a, a); }
#ifndef FE_MUL_INLINE
__device__ __noinline__ fe fe_mul_fn(const fe a, const fe b)
{
	fe r;
	fe_mul_i(r, a, b);
	return r;
}
__device__ __forceinline__ void fe_mul(fe& r, const fe& a, const fe& b) { r = fe_mul_fn(a, b); }
#else
__device__ __forceinline__ void fe_mul(fe& r, const fe& a, const fe& b) { fe_mul_i(r, a, b); }
#endif
__device__ __forceinline__ void fe_sqr(fe& r, const fe& a) { fe_mul(r, a, a); }

// r = a + b (weak)
__device__ __forceinline__ void fe_add(fe& r, const fe& a, const fe& b)
{
	u32 t[8];
	u32 c = add8(t, a.v, b.v);
	// fold the carry (2^256 = 189); a second carry leaves a value < 189, then +189 is safe
	c = add8_u32(t, c * FE_C);
	t[0] += c * FE_C;
#pragma unroll
	for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}

// r = a - b (weak)
__device__ __forceinline__ void fe_sub(fe& r, const fe& a, const fe& b)
{
	u32 t[8];
	u32 m = sub8(t, a.v, b.v);
	// a borrow means t = a - b + 2^256 = a - b + 189 (mod p): take 189 back; if that borrows
	// again the value wrapped to >= 2^256 - 189 and a further -189 cannot borrow
	m = sub8_u32(t, m & FE_C);
	t[0] -= m & FE_C;
#pragma unroll
	for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}

__device__ __forceinline__ void fe_dbl(fe& r, const fe& a) { fe_add(r, a, a); }

// canonical form in [0, p)
__device__ __forceinline__ void fe_canon(fe& a)
{
	// a >= p  <=>  a + 189 >= 2^256, and then a - p = a + 189 - 2^256
	u32 t[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) t[i] = a.v[i];
	if (add8_u32(t, FE_C))
	{
#pragma unroll
		for (int i = 0; i < 8; ++i) a.v[i] = t[i];
	}
}

__device__ __forceinline__ bool fe_is_zero(const fe& a)
{
	fe t = a;
	fe_canon(t);
	return (t.v[0] | t.v[1] | t.v[2] | t.v[3] | t.v[4] | t.v[5] | t.v[6] | t.v[7]) == 0;
}

__device__ __forceinline__ void fe_set_u32(fe& r, u32 x)
{
	r.v[0] = x;
#pragma unroll
	for (int k = 1; k < 8; ++k) r.v[k] = 0;
}

// raw 256-bit compare (little-endian limbs): a >= b ?
__device__ __forceinline__ bool u256_geq(const u32* a, const u32* b)
{
	u32 t[8];
	return sub8(t, a, b) == 0;
}

// r = a^(2^n) by n squarings
__device__ __forceinline__ void fe_sqr_n(fe& r, const fe& a, int n)
{
	r = a;
#pragma unroll 1
	for (int i = 0; i < n; ++i)
		fe_sqr(r, r);
}

// r = a^(p-2) = 1/a (gfp.c:33-44 computes the same power with a sliding window).
// p - 2 = 2^256 - 191 = (2^248 - 1) * 2^8 + 0x41: an addition chain on runs of ones.
__device__ __noinline__ fe fe_inv_fn(const fe a)
{
	fe x2, x4, x8, x16, x32, x64, x128, t, r;
	fe_sqr(t, a), fe_mul(x2, t, a);                 // 2^2 - 1
	fe_sqr_n(t, x2, 2), fe_mul(x4, t, x2);          // 2^4 - 1
	fe_sqr_n(t, x4, 4), fe_mul(x8, t, x4);          // 2^8 - 1
	fe_sqr_n(t, x8, 8), fe_mul(x16, t, x8);         // 2^16 - 1
	fe_sqr_n(t, x16, 16), fe_mul(x32, t, x16);      // 2^32 - 1
	fe_sqr_n(t, x32, 32), fe_mul(x64, t, x32);      // 2^64 - 1
	fe_sqr_n(t, x64, 64), fe_mul(x128, t, x64);     // 2^128 - 1
	fe_sqr_n(t, x128, 64), fe_mul(t, t, x64);       // 2^192 - 1
	fe_sqr_n(t, t, 32), fe_mul(t, t, x32);          // 2^224 - 1
	fe_sqr_n(t, t, 16), fe_mul(t, t, x16);          // 2^240 - 1
	fe_sqr_n(t, t, 8), fe_mul(t, t, x8);            // 2^248 - 1
	// append 0x41 = 0100 0001b
	fe_sqr_n(t, t, 2), fe_mul(t, t, a);             // ...01
	fe_sqr_n(t, t, 6), fe_mul(r, t, a);             // ...01000001
	return r;
}
__device__ __forceinline__ void fe_inv(fe& r, const fe& a) { r = fe_inv_fn(a); }

// little-endian octets <-> limbs (unaligned-safe)
__device__ __forceinline__ void fe_load(fe& r, const u8* p)
{
	if (((uintptr_t)p & 3) == 0)
	{
#pragma unroll
		for (int k = 0; k < 8; ++k) r.v[k] = reinterpret_cast<const u32*>(p)[k];
	}
	else
	{
#pragma unroll
		for (int k = 0; k < 8; ++k)
			r.v[k] = (u32)p[4 * k] | (u32)p[4 * k + 1] << 8 | (u32)p[4 * k + 2] << 16 | (u32)p[4 * k + 3] << 24;
	}
}
__device__ __forceinline__ void fe_store(u8* p, const fe& a)
{
	if (((uintptr_t)p & 3) == 0)
	{
#pragma unroll
		for (int k = 0; k < 8; ++k) reinterpret_cast<u32*>(p)[k] = a.v[k];
	}
	else
	{
#pragma unroll
		for (int k = 0; k < 8; ++k)
			for (int b = 0; b < 4; ++b) p[4 * k + b] = (u8)(a.v[k] >> (8 * b));
	}
}
