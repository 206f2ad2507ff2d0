// ecp256.cuh — Jacobian point arithmetic on y^2 = x^3 - 3x + b over GF(2^256 - 189).
//
// Replaces the reference's ecpDblJA3 / ecpAddJ / ecpAddAJ / ecpToAJ (ecp_j.c:241-299,
// :397-497, :516-590, :104-133). x = X/Z^2, y = Y/Z^3, O <=> Z = 0. The formulas never use
// the coefficient b, exactly like the reference's.
//
// Every addition is COMPLETE in the reference's sense (ecp_j.c:416-427, :455-464): O + P,
// P + O, P + P (-> doubling) and P + (-P) (-> O) are detected and handled by rare,
// warp-divergent branches, so crafted inputs (Q = +-G, small multiples of G, ...) give the
// same affine result as the reference.
#pragma once
#include "gfp256.cuh"

struct pt { fe X, Y, Z; };

// PT_OUTLINE: dbl / madd / add are real functions (points passed by address, i.e. through local
// memory on the otherwise idle LSU pipe) with the multiplications inlined in their bodies;
// otherwise they are inlined and every multiplication is a call.
#ifdef PT_OUTLINE
#define PT_OP __device__ __noinline__
#define PMUL fe_mul_i
#define PSQR fe_sqr_i
#else
#define PT_OP __device__ __forceinline__
#define PMUL fe_mul
#define PSQR fe_sqr
#endif

__device__ __forceinline__ void pt_set_inf(pt& R)
{
	fe_set_u32(R.X, 1), fe_set_u32(R.Y, 1), fe_set_u32(R.Z, 0);
}
__device__ __forceinline__ void pt_set_affine(pt& R, const fe& x, const fe& y)
{
	R.X = x, R.Y = y, fe_set_u32(R.Z, 1);
}
__device__ __forceinline__ bool pt_is_inf(const pt& P) { return fe_is_zero(P.Z); }

// R = 2P, a = -3 (3M + 5S). Z = 0 or Y = 0 give Z3 = 0 without special casing.
PT_OP void pt_dbl(pt& R, const pt& P)
{
	fe delta, gamma, beta, alpha, t, u;
	PSQR(delta, P.Z);
	PSQR(gamma, P.Y);
	PMUL(beta, P.X, gamma);
	fe_sub(t, P.X, delta), fe_add(u, P.X, delta);
	PMUL(alpha, t, u);
	fe_dbl(t, alpha), fe_add(alpha, alpha, t);          // 3 (X - Z^2)(X + Z^2)
	fe_add(t, P.Y, P.Z), PSQR(t, t);
	fe_sub(t, t, gamma), fe_sub(R.Z, t, delta);         // Z3 = (Y + Z)^2 - Y^2 - Z^2
	fe_dbl(beta, beta), fe_dbl(beta, beta);             // 4 beta
	PSQR(t, alpha);
	fe_sub(t, t, beta), fe_sub(R.X, t, beta);           // X3 = alpha^2 - 8 beta
	PSQR(gamma, gamma);
	fe_dbl(gamma, gamma), fe_dbl(gamma, gamma), fe_dbl(gamma, gamma);   // 8 gamma^2
	fe_sub(t, beta, R.X), PMUL(t, alpha, t);
	fe_sub(R.Y, t, gamma);                              // Y3 = alpha (4 beta - X3) - 8 gamma^2
}
// out-of-line copy for the rare P + P branches
__device__ __noinline__ void pt_dbl_slow(pt* R, const pt* P)
{
	pt t = *P;
	pt_dbl(t, t);
	*R = t;
}

// R = P + (x2, y2), the second point affine and finite (7M + 4S)
PT_OP void pt_madd(pt& R, const pt& P, const fe& x2, const fe& y2)
{
	fe z1z1, u2, s2, h, hh, i, j, r, v, t;
	PSQR(z1z1, P.Z);
	PMUL(u2, x2, z1z1);
	PMUL(t, P.Z, z1z1), PMUL(s2, y2, t);
	fe_sub(h, u2, P.X);
	fe_sub(r, s2, P.Y);
	const bool p_inf = fe_is_zero(P.Z);
	if (p_inf || fe_is_zero(h))
	{
		if (p_inf)
			pt_set_affine(R, x2, y2);
		else if (fe_is_zero(r))
		{
			pt q;
			pt_set_affine(q, x2, y2);
			pt_dbl_slow(&R, &q);
		}
		else
			pt_set_inf(R);
		return;
	}
	fe_dbl(r, r);
	PSQR(hh, h);
	fe_dbl(i, hh), fe_dbl(i, i);                        // I = 4 HH
	PMUL(j, h, i);
	PMUL(v, P.X, i);
	fe_add(t, P.Z, h), PSQR(t, t);
	fe_sub(t, t, z1z1), fe_sub(t, t, hh);               // Z3 = (Z1 + H)^2 - Z1Z1 - HH
	fe z3 = t;
	PSQR(t, r);
	fe_sub(t, t, j), fe_sub(t, t, v), fe_sub(t, t, v);  // X3 = r^2 - J - 2V
	fe x3 = t;
	fe_sub(t, v, x3), PMUL(t, r, t);
	PMUL(j, P.Y, j), fe_dbl(j, j);
	fe_sub(R.Y, t, j);                                  // Y3 = r (V - X3) - 2 Y1 J
	R.X = x3, R.Z = z3;
}

// R = P + Q, both Jacobian (11M + 5S)
PT_OP void pt_add(pt& R, const pt& P, const pt& Q)
{
	fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
	PSQR(z1z1, P.Z), PSQR(z2z2, Q.Z);
	PMUL(u1, P.X, z2z2), PMUL(u2, Q.X, z1z1);
	PMUL(t, Q.Z, z2z2), PMUL(s1, P.Y, t);
	PMUL(t, P.Z, z1z1), PMUL(s2, Q.Y, t);
	fe_sub(h, u2, u1);
	fe_sub(r, s2, s1);
	const bool p_inf = fe_is_zero(P.Z), q_inf = fe_is_zero(Q.Z);
	if (p_inf || q_inf || fe_is_zero(h))
	{
		if (p_inf)
			R = Q;
		else if (q_inf)
			R = P;
		else if (fe_is_zero(r))
			pt_dbl_slow(&R, &P);
		else
			pt_set_inf(R);
		return;
	}
	fe_dbl(r, r);
	fe_dbl(t, h), PSQR(i, t);                         // I = (2H)^2
	PMUL(j, h, i);
	PMUL(v, u1, i);
	fe_add(t, P.Z, Q.Z), PSQR(t, t);
	fe_sub(t, t, z1z1), fe_sub(t, t, z2z2);
	fe z3;
	PMUL(z3, t, h);                                   // Z3 = ((Z1 + Z2)^2 - Z1Z1 - Z2Z2) H
	PSQR(t, r);
	fe_sub(t, t, j), fe_sub(t, t, v), fe_sub(t, t, v);  // X3 = r^2 - J - 2V
	fe x3 = t;
	fe_sub(t, v, x3), PMUL(t, r, t);
	PMUL(j, s1, j), fe_dbl(j, j);
	fe_sub(R.Y, t, j);                                  // Y3 = r (V - X3) - 2 S1 J
	R.X = x3, R.Z = z3;
}

// affine coordinates of a finite point, canonical residues (ecp_j.c:104-133)
__device__ __forceinline__ void pt_to_affine(fe& x, fe& y, const pt& P)
{
	fe zi, zi2;
	fe_inv(zi, P.Z);
	fe_sqr(zi2, zi);
	fe_mul(x, P.X, zi2);
	fe_mul(zi2, zi2, zi), fe_mul(y, P.Y, zi2);
	fe_canon(x), fe_canon(y);
}
__device__ __forceinline__ void pt_to_affine_x(fe& x, const pt& P)
{
	fe zi;
	fe_inv(zi, P.Z);
	fe_sqr(zi, zi);
	fe_mul(x, P.X, zi);
	fe_canon(x);
}
