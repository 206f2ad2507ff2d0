// ecp.cuh — Jacobian point arithmetic on y^2 = x^3 - 3x + b over GF(2^(32N) - c), N = 8/12/16
// (the three standard bign curves all have a = p - 3, bign_params.c:42-49,86-93,141-150).
//
// Replaces the reference's ecpDblJA3 / ecpAddJ / ecpAddAJ / ecpToAJ (ecp_j.c:241-299,
// :397-497, :516-590, :104-133) and the variable-base part of ecMulA / ecAddMulA
// (ec.c:497-525, :1183-1273). x = X/Z^2, y = Y/Z^3, O <=> Z = 0. The formulas never use the
// coefficient b, exactly like the reference's.
//
// Every addition is COMPLETE in the reference's sense (ecp_j.c:416-427, :455-464): O + P,
// P + O, P + P (-> doubling) and P + (-P) (-> O) are detected and handled by rare,
// warp-divergent branches, so crafted inputs (Q = +-G, small multiples of G, ...) give the
// same affine result as the reference.
#pragma once
#include "gfp.cuh"

template <int N> struct pt { fe<N> X, Y, Z; };
// a scalar of up to 32 N bits handed BY VALUE to the out-of-line multiplication routines: the
// kernels keep no address-taken locals besides the accumulator (nvcc 12.9 was seen to overlap the
// stack slots of two live address-taken locals of bign_verify_kernel — see DESIGN.md §4.3)
template <int N> struct sc { u32 w[N]; };

#define PT_OP GFP_HD

template <int N> GFP_HD void pt_set_inf(pt<N>& R)
{
	fe_set_u32<N>(R.X, 1), fe_set_u32<N>(R.Y, 1), fe_set_u32<N>(R.Z, 0);
}
template <int N> GFP_HD void pt_set_affine(pt<N>& R, const fe<N>& x, const fe<N>& y)
{
	R.X = x, R.Y = y, fe_set_u32<N>(R.Z, 1);
}
template <int N> GFP_HD bool pt_is_inf(const pt<N>& P) { return fe_is_zero<N>(P.Z); }
template <int N> GFP_HD void fe_neg(fe<N>& r, const fe<N>& a)
{
	fe<N> z;
	fe_set_u32<N>(z, 0);
	fe_sub<N>(r, z, a);
}

// R = 2P, a = -3 (3M + 5S). Z = 0 or Y = 0 give Z3 = 0 without special casing.
// INL: products inlined instead of calls to the shared copies (-DPT_DBL_INLINE=1 uses it in the hot
// loop of pt_mul_var). Measured on B200: 40.4 M verifies/s inlined vs 46.3 M/s with calls (register
// pressure and code size outweigh the saved argument moves), so the default is calls.
template <int N, bool INL> PT_OP void pt_dbl_t(pt<N>& R, const pt<N>& P)
{
#define PM(r, a, b) (INL ? fe_mul_i<N>(r, a, b) : fe_mul<N>(r, a, b))
#define PS(r, a) (INL ? fe_sqr_i<N>(r, a) : fe_sqr<N>(r, a))
	// Order of evaluation = order of the calls (ptxas does not move them): chosen so that few values are
	// alive across each call — Z3 first (Y and Z die), then alpha (delta dies), beta (X dies), ... — because
	// whatever is alive across a call beyond the callee-saved registers is spilled around it.
	// R may alias P: a coordinate of R is written only after the last read of that coordinate of P.
	fe<N> delta, gamma, beta, alpha, t, u;
#ifdef PT_DBL_ORDER_R1   /* A/B measurement only: the order of round 1 (all four squares and products first) */
	PS(delta, P.Z);
	PS(gamma, P.Y);
	PM(beta, P.X, gamma);
	fe_sub<N>(t, P.X, delta), fe_add<N>(u, P.X, delta);
	PM(alpha, t, u);
	fe_dbl<N>(t, alpha), fe_add<N>(alpha, alpha, t);
	fe_add<N>(t, P.Y, P.Z), PS(t, t);
	fe_sub<N>(t, t, gamma), fe_sub<N>(R.Z, t, delta);
	fe_shl<2, N>(beta, beta);
	PS(t, alpha);
	fe_sub<N>(t, t, beta), fe_sub<N>(R.X, t, beta);
	PS(gamma, gamma);
	fe_shl<3, N>(gamma, gamma);
	fe_sub<N>(t, beta, R.X), PM(t, alpha, t);
	fe_sub<N>(R.Y, t, gamma);
	return;
#endif
	PS(delta, P.Z);
	PS(gamma, P.Y);
	fe_add<N>(t, P.Y, P.Z), PS(t, t);
	fe_sub2<N>(t, t, gamma, delta);                               // Z3 = (Y + Z)^2 - Y^2 - Z^2
	fe_sub<N>(u, P.X, delta), fe_add<N>(delta, P.X, delta);
	R.Z = t;
	PM(alpha, u, delta);
	fe_mul3<N>(alpha, alpha);                                     // 3 (X - Z^2)(X + Z^2)
	PM(beta, P.X, gamma);
	fe_shl<2, N>(beta, beta);                                     // 4 beta
	PS(gamma, gamma);
	fe_shl<3, N>(gamma, gamma);                                   // 8 gamma^2
	PS(t, alpha);
	fe_sub2<N>(t, t, beta, beta);                                 // X3 = alpha^2 - 8 beta
	R.X = t;
	fe_sub<N>(t, beta, t), PM(t, alpha, t);
	fe_sub<N>(R.Y, t, gamma);                                     // Y3 = alpha (4 beta - X3) - 8 gamma^2
#undef PM
#undef PS
}
template <int N> PT_OP void pt_dbl(pt<N>& R, const pt<N>& P) { pt_dbl_t<N, false>(R, P); }
#ifndef PT_DBL_INLINE
#define PT_DBL_INLINE 0
#endif
// out-of-line copy for the rare P + P branches
template <int N> __host__ __device__ __noinline__ void pt_dbl_slow(pt<N>* R, const pt<N>* P)
{
	pt<N> t = *P;
	pt_dbl<N>(t, t);
	*R = t;
}

// R = P + (x2, y2), the second point affine and finite (7M + 4S)
template <int N> PT_OP void pt_madd(pt<N>& R, const pt<N>& P, const fe<N>& x2, const fe<N>& y2)
{
	fe<N> z1z1, u2, s2, h, hh, i, j, r, v, t;
	fe_sqr<N>(z1z1, P.Z);
	fe_mul<N>(u2, x2, z1z1);
	fe_mul<N>(t, P.Z, z1z1), fe_mul<N>(s2, y2, t);
	fe_sub<N>(h, u2, P.X);
	fe_sub<N>(r, s2, P.Y);
	const bool p_inf = fe_is_zero<N>(P.Z);
	if (p_inf || fe_is_zero<N>(h))
	{
		if (p_inf)
			pt_set_affine<N>(R, x2, y2);
		else if (fe_is_zero<N>(r))
		{
			pt<N> q;
			pt_set_affine<N>(q, x2, y2);
			pt_dbl_slow<N>(&R, &q);
		}
		else
			pt_set_inf<N>(R);
		return;
	}
	// (order of the calls chosen for short live ranges, as in pt_dbl_t: Z3 first, then z1z1 and hh are dead;
	// R may alias P: Z3 is written after the last read of P.Z, X3 / Y3 after the last reads of P.X / P.Y)
	fe_dbl<N>(r, r);
	fe_sqr<N>(hh, h);
	fe_add<N>(t, P.Z, h), fe_sqr<N>(t, t);
	fe_sub2<N>(R.Z, t, z1z1, hh);                                 // Z3 = (Z1 + H)^2 - Z1Z1 - HH
	fe_shl<2, N>(i, hh);                                          // I = 4 HH
	fe_mul<N>(j, h, i);
	fe_mul<N>(v, P.X, i);
	fe_sqr<N>(t, r);
	fe_sub2<N>(t, t, j, v), fe_sub<N>(t, t, v);                   // X3 = r^2 - J - 2V
	fe_mul<N>(j, P.Y, j), fe_dbl<N>(j, j);
	R.X = t;
	fe_sub<N>(t, v, t), fe_mul<N>(t, r, t);
	fe_sub<N>(R.Y, t, j);                                         // Y3 = r (V - X3) - 2 Y1 J
}

// R = P + Q, both Jacobian (11M + 5S)
template <int N> PT_OP void pt_add(pt<N>& R, const pt<N>& P, const pt<N>& Q)
{
	fe<N> z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
	// (first everything that reads Q.X and Q.Y — the table entry sits in registers, P in memory)
	fe_sqr<N>(z1z1, P.Z);
	fe_mul<N>(u2, Q.X, z1z1);
	fe_mul<N>(t, P.Z, z1z1), fe_mul<N>(s2, Q.Y, t);
	fe_sqr<N>(z2z2, Q.Z);
	fe_mul<N>(u1, P.X, z2z2);
	fe_mul<N>(t, Q.Z, z2z2), fe_mul<N>(s1, P.Y, t);
	fe_sub<N>(h, u2, u1);
	fe_sub<N>(r, s2, s1);
	const bool p_inf = fe_is_zero<N>(P.Z), q_inf = fe_is_zero<N>(Q.Z);
	if (p_inf || q_inf || fe_is_zero<N>(h))
	{
		if (p_inf)
			R = Q;
		else if (q_inf)
			R = P;
		else if (fe_is_zero<N>(r))
			pt_dbl_slow<N>(&R, &P);
		else
			pt_set_inf<N>(R);
		return;
	}
	// (Z3 first: z1z1, z2z2 and both Z die; from here on P and Q are no longer read, so R may be written)
	fe_dbl<N>(r, r);
	fe_add<N>(t, P.Z, Q.Z), fe_sqr<N>(t, t);
	fe_sub2<N>(t, t, z1z1, z2z2);
	fe_mul<N>(R.Z, t, h);                                         // Z3 = ((Z1 + Z2)^2 - Z1Z1 - Z2Z2) H
	fe_dbl<N>(t, h), fe_sqr<N>(i, t);                             // I = (2H)^2
	fe_mul<N>(j, h, i);
	fe_mul<N>(v, u1, i);
	fe_sqr<N>(t, r);
	fe_sub2<N>(t, t, j, v), fe_sub<N>(t, t, v);                   // X3 = r^2 - J - 2V
	fe_mul<N>(j, s1, j), fe_dbl<N>(j, j);
	R.X = t;
	fe_sub<N>(t, v, t), fe_mul<N>(t, r, t);
	fe_sub<N>(R.Y, t, j);                                         // Y3 = r (V - X3) - 2 S1 J
}

// affine coordinates of a finite point, canonical residues (ecp_j.c:104-133)
template <int N> GFP_HD void pt_to_affine(fe<N>& x, fe<N>& y, const pt<N>& P)
{
	fe<N> zi, zi2;
	fe_inv<N>(zi, P.Z);
	fe_sqr<N>(zi2, zi);
	fe_mul<N>(x, P.X, zi2);
	fe_mul<N>(zi2, zi2, zi), fe_mul<N>(y, P.Y, zi2);
	fe_canon<N>(x), fe_canon<N>(y);
}
template <int N> GFP_HD void pt_to_affine_x(fe<N>& x, const pt<N>& P)
{
	fe<N> zi;
	fe_inv<N>(zi, P.Z);
	fe_sqr<N>(zi, zi);
	fe_mul<N>(x, P.X, zi);
	fe_canon<N>(x);
}

// R <- m ? P : R for an all-ones / all-zero mask m, without a branch (the regular forms below)
template <int N> GFP_HD void pt_select(pt<N>& R, const pt<N>& P, u32 m)
{
#pragma unroll
	for (int k = 0; k < N; ++k)
	{
		R.X.v[k] ^= (R.X.v[k] ^ P.X.v[k]) & m;
		R.Y.v[k] ^= (R.Y.v[k] ^ P.Y.v[k]) & m;
		R.Z.v[k] ^= (R.Z.v[k] ^ P.Z.v[k]) & m;
	}
}

// ---------------------------------------------------------------- variable-base multiplication
// acc = k * (x, y) for a scalar of nbits bits (little-endian limbs, bits above nbits must be 0).
// REGULAR signed 5-bit windows so that all lanes of a warp run the same doublings and additions
// in lock-step (the reference's width-5 wNAF, ec.c:435-484, is irregular and would diverge):
//   k = sum d_i 32^i, d_i in [-15, 16]: d_i = (window i) + carry_i, minus 32 (carry out) if > 16;
//   table {1..16}(x, y) in local memory (8 doublings + 7 mixed additions);
//   5 doublings + 1 addition per window, most significant first; nbits/5 + 1 windows, the top one
//   always has a spare bit and absorbs the last carry.
//
// CT = true is the form for SECRET scalars (bignDH, ecMulA callers such as key transport; the reference's
// ecMulA is regular on purpose, ec.c:497-525 with wwSel-style table reads, ww.c:298-310): the table entry
// is fetched by a masked scan over all 16 entries instead of T[d], a zero digit still performs the
// addition (with entry 1) and the result is dropped by a mask, the sign is applied by a mask. What remains
// data-dependent are the exceptional branches of pt_add (P = +-Q, probability ~ 2^-250 for an honest
// scalar) and the top window. CT = false (public scalars: verification) keeps the direct, cheaper form.
#define PT_WIN 5
#ifndef PT_WIN_V8
#define PT_WIN_V8 0   /* 256-bit accesses: ptxas 12.9 crashes on them inside the verification kernel */
#endif

// ---- storage of the window table T[1..16]
// win_local: a per-thread array, i.e. LOCAL memory, which the hardware interleaves over the lanes of a warp
// (word k of all 32 lanes shares one 128-byte line). Ideal when every lane reads the SAME entry — the masked
// scan of the CT form — and the worst case when every lane reads its OWN entry T[d]: a warp then touches
// ~14 different lines per word, 24 words per entry, for 3 KB of useful data: 43 KB per window and warp,
// 8.8 GB of DRAM traffic per 2^18 verifications (ncu r01 / r02: 212 x the algorithmic bytes).
template <int N> struct win_local
{
	pt<N> T[(1 << (PT_WIN - 1)) + 1];   // T[0] unused
	GFP_HD void put(int j, const pt<N>& P) { T[j] = P; }
	GFP_HD void get(pt<N>& P, int j) const { P = T[j]; }
};
// win_global: the table in a GLOBAL scratch area of the CTA, laid out for the access pattern of the direct
// form (public scalars: verification): an entry is G granules of 32 octets; granule g of entry j of thread t
// lies at area + ((j - 1) G + g) * 32 T + 32 t (T threads per CTA). A put is G fully coalesced 256-bit stores
// per warp, a get of the lane's own entry is G 256-bit loads of exactly the sectors that hold it — the table
// costs its own size in traffic (16 x 96 B written once, 25 x 96 B read per verification at N = 8).
// A thread only ever reads what it wrote itself: no fences.
template <int N> struct win_global
{
	static constexpr int G = (12 * N + 31) / 32;
	static constexpr size_t ITEM_BYTES = (size_t)(32 * G) << (PT_WIN - 1);   // per thread
	u8* base;      // the CTA's area + 32 * threadIdx.x
	u32 stride;    // 32 * blockDim.x
	GFP_HD static u32 word(const pt<N>& P, int k) { return k < N ? P.X.v[k] : k < 2 * N ? P.Y.v[k - N] : k < 3 * N ? P.Z.v[k - 2 * N] : 0u; }
	GFP_HD static void set_word(pt<N>& P, int k, u32 w)
	{
		if (k < N) P.X.v[k] = w;
		else if (k < 2 * N) P.Y.v[k - N] = w;
		else if (k < 3 * N) P.Z.v[k - 2 * N] = w;
	}
	GFP_HD void put(int j, const pt<N>& P) const
	{
		u8* e = base + (size_t)((j - 1) * G) * stride;
#pragma unroll
		for (int g = 0; g < G; ++g)
		{
#ifdef __CUDA_ARCH__
#if PT_WIN_V8
			asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
				:: "l"(e + (size_t)g * stride), "r"(word(P, 8 * g)), "r"(word(P, 8 * g + 1)), "r"(word(P, 8 * g + 2)),
				"r"(word(P, 8 * g + 3)), "r"(word(P, 8 * g + 4)), "r"(word(P, 8 * g + 5)), "r"(word(P, 8 * g + 6)),
				"r"(word(P, 8 * g + 7)) : "memory");
#else
#pragma unroll
			for (int h = 0; h < 2; ++h)
				asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};"
					:: "l"(e + (size_t)g * stride + 16 * h), "r"(word(P, 8 * g + 4 * h)), "r"(word(P, 8 * g + 4 * h + 1)),
					"r"(word(P, 8 * g + 4 * h + 2)), "r"(word(P, 8 * g + 4 * h + 3)) : "memory");
#endif
#else
			for (int k = 0; k < 8; ++k) reinterpret_cast<u32*>(e + (size_t)g * stride)[k] = word(P, 8 * g + k);
#endif
		}
	}
	GFP_HD void get(pt<N>& P, int j) const
	{
		const u8* e = base + (size_t)((j - 1) * G) * stride;
#pragma unroll
		for (int g = 0; g < G; ++g)
		{
			u32 w[8];
#ifdef __CUDA_ARCH__
#if PT_WIN_V8
			asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				: "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
				: "l"(e + (size_t)g * stride) : "memory");
#else
#pragma unroll
			for (int h = 0; h < 2; ++h)
				asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
					: "=r"(w[4 * h]), "=r"(w[4 * h + 1]), "=r"(w[4 * h + 2]), "=r"(w[4 * h + 3])
					: "l"(e + (size_t)g * stride + 16 * h) : "memory");
#endif
#else
			for (int k = 0; k < 8; ++k) w[k] = reinterpret_cast<const u32*>(e + (size_t)g * stride)[k];
#endif
#pragma unroll
			for (int k = 0; k < 8; ++k) set_word(P, 8 * g + k, w[k]);
		}
	}
};

template <int N, bool CT, class WIN> GFP_HD void pt_mul_var_w(pt<N>& acc, const u32* k, int nbits, const fe<N>& x, const fe<N>& y, WIN& T)
{
	constexpr int HALF = 1 << (PT_WIN - 1);
	{
		// T[j] = j * (x, y): 8 doublings + 7 mixed additions; an odd entry follows from the one just made
		pt<N> cur;
		pt_set_affine<N>(cur, x, y);
		T.put(1, cur);
#pragma unroll 1
		for (int j = 2; j <= HALF; ++j)
		{
			if (j & 1)
				pt_madd<N>(cur, cur, x, y);
			else
			{
				T.get(cur, j >> 1);
				pt_dbl<N>(cur, cur);
			}
			T.put(j, cur);
		}
	}
	const int nw = nbits / PT_WIN + 1;
	// carries of the recoding, least significant window first: bit i of cy = carry INTO window i
	u32 cy[(32 * N / PT_WIN + 2 + 31) / 32];
#pragma unroll
	for (int i = 0; i < (int)(sizeof(cy) / 4); ++i) cy[i] = 0;
	{
		u32 c = 0;
#pragma unroll 1
		for (int i = 0; i < nw; ++i)
		{
			const int bit = PT_WIN * i, limb = bit >> 5, sh = bit & 31;
			u32 w = limb < N ? k[limb] >> sh : 0u;
			if (sh > 32 - PT_WIN && limb + 1 < N)
				w |= k[limb + 1] << (32 - sh);
			w = (w & (2 * HALF - 1)) + c;
			cy[i >> 5] |= c << (i & 31);
			c = w > HALF ? 1u : 0u;
		}
	}
	bool first = true;
#pragma unroll 1
	for (int i = nw - 1; i >= 0; --i)
	{
		if (!first)
		{
#pragma unroll 1
			for (int s = 0; s < PT_WIN; ++s)
				pt_dbl_t<N, PT_DBL_INLINE != 0>(acc, acc);
		}
		const int bit = PT_WIN * i, limb = bit >> 5, sh = bit & 31;
		u32 w = limb < N ? k[limb] >> sh : 0u;
		if (sh > 32 - PT_WIN && limb + 1 < N)
			w |= k[limb + 1] << (32 - sh);
		w = (w & (2 * HALF - 1)) + ((cy[i >> 5] >> (i & 31)) & 1u);
		const bool neg = w > HALF;
		const u32 d = neg ? 2 * HALF - w : w;
		if constexpr (CT)
		{
			// Q = T[d ? d : 1] by a scan over the whole table; -Q by a mask
			const u32 dd = d | (u32)(d == 0);
			pt<N> Q, S;
#pragma unroll
			for (int k = 0; k < N; ++k) Q.X.v[k] = 0, Q.Y.v[k] = 0, Q.Z.v[k] = 0;
#pragma unroll 1
			for (u32 j = 1; j <= (u32)HALF; ++j)
			{
				const u32 m = 0u - (u32)(j == dd);
#pragma unroll
				for (int k = 0; k < N; ++k)
					Q.X.v[k] |= T.T[j].X.v[k] & m, Q.Y.v[k] |= T.T[j].Y.v[k] & m, Q.Z.v[k] |= T.T[j].Z.v[k] & m;
			}
			{
				fe<N> ny;
				fe_neg<N>(ny, Q.Y);
				const u32 mn = 0u - (u32)neg;
#pragma unroll
				for (int k = 0; k < N; ++k) Q.Y.v[k] ^= (Q.Y.v[k] ^ ny.v[k]) & mn;
			}
			const u32 nz = 0u - (u32)(d != 0);
			if (first)
			{
				// top window: acc = d ? Q : O (Z = 0)
				acc = Q;
#pragma unroll
				for (int k = 0; k < N; ++k) acc.Z.v[k] &= nz;
				first = false;
			}
			else
			{
				pt_add<N>(S, acc, Q);
				pt_select<N>(acc, S, nz);
			}
			continue;
		}
		if (first)
		{
			// the top window only selects (no doublings of O); it is never negative
			if (d)
				T.get(acc, (int)d);
			else
				pt_set_inf<N>(acc);
			first = false;
		}
		else if (d)
		{
			pt<N> Q;
			T.get(Q, (int)d);
			if (neg)
				fe_neg<N>(Q.Y, Q.Y);
			pt_add<N>(acc, acc, Q);
		}
	}
}

// the out-of-line forms: table in local memory (CT = true: secret scalars; CT = false: table building) ...
template <int N, bool CT = false> __host__ __device__ __noinline__ void pt_mul_var(pt<N>& acc, const sc<N> ks, int nbits,
	const fe<N> x, const fe<N> y)
{
	win_local<N> T;
	pt_mul_var_w<N, CT>(acc, ks.w, nbits, x, y, T);
}
// ... and in the caller's global scratch area (public scalars only: the verification kernel)
template <int N> __host__ __device__ __noinline__ void pt_mul_var_g(pt<N>& acc, const sc<N> ks, int nbits,
	const fe<N> x, const fe<N> y, win_global<N> T)
{
	pt_mul_var_w<N, false>(acc, ks.w, nbits, x, y, T);
}
