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
__device__ __forceinline__ void pt_dbl(pt& R, const pt& P)
{
	fe delta, gamma, beta, alpha, t, u;
	fe_sqr(delta, P.Z);
	fe_sqr(gamma, P.Y);
	fe_mul(beta, P.X, gamma);
	fe_sub(t, P.X, delta), fe_add(u, P.X, delta);
	fe_mul(alpha, t, u);
	fe_dbl(t, alpha), fe_add(alpha, alpha, t);          // 3 (X - Z^2)(X + Z^2)
	fe_add(t, P.Y, P.Z), fe_sqr(t, t);
	fe_sub(t, t, gamma), fe_sub(R.Z, t, delta);         // Z3 = (Y + Z)^2 - Y^2 - Z^2
	fe_dbl(beta, beta), fe_dbl(beta, beta);             // 4 beta
	fe_sqr(t, alpha);
	fe_sub(t, t, beta), fe_sub(R.X, t, beta);           // X3 = alpha^2 - 8 beta
	fe_sqr(gamma, gamma);
	fe_dbl(gamma, gamma), fe_dbl(gamma, gamma), fe_dbl(gamma, gamma);   // 8 gamma^2
	fe_sub(t, beta, R.X), fe_mul(t, alpha, t);
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
__device__ __forceinline__ void pt_madd(pt& R, const pt& P, const fe& x2, const fe& y2)
{
	fe z1z1, u2, s2, h, hh, i, j, r, v, t;
	fe_sqr(z1z1, P.Z);
	fe_mul(u2, x2, z1z1);
	fe_mul(t, P.Z, z1z1), fe_mul(s2, y2, t);
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
	fe_sqr(hh, h);
	fe_dbl(i, hh), fe_dbl(i, i);                        // I = 4 HH
	fe_mul(j, h, i);
	fe_mul(v, P.X, i);
	fe_add(t, P.Z, h), fe_sqr(t, t);
	fe_sub(t, t, z1z1), fe_sub(t, t, hh);               // Z3 = (Z1 + H)^2 - Z1Z1 - HH
	fe z3 = t;
	fe_sqr(t, r);
	fe_sub(t, t, j), fe_sub(t, t, v), fe_sub(t, t, v);  // X3 = r^2 - J - 2V
	fe x3 = t;
	fe_sub(t, v, x3), fe_mul(t, r, t);
	fe_mul(j, P.Y, j), fe_dbl(j, j);
	fe_sub(R.Y, t, j);                                  // Y3 = r (V - X3) - 2 Y1 J
	R.X = x3, R.Z = z3;
}

// R = P + Q, both Jacobian (11M + 5S)
__device__ __forceinline__ void pt_add(pt& R, const pt& P, const pt& Q)
{
	fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
	fe_sqr(z1z1, P.Z), fe_sqr(z2z2, Q.Z);
	fe_mul(u1, P.X, z2z2), fe_mul(u2, Q.X, z1z1);
	fe_mul(t, Q.Z, z2z2), fe_mul(s1, P.Y, t);
	fe_mul(t, P.Z, z1z1), fe_mul(s2, Q.Y, t);
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
	fe_dbl(t, h), fe_sqr(i, t);                         // I = (2H)^2
	fe_mul(j, h, i);
	fe_mul(v, u1, i);
	fe_add(t, P.Z, Q.Z), fe_sqr(t, t);
	fe_sub(t, t, z1z1), fe_sub(t, t, z2z2);
	fe z3;
	fe_mul(z3, t, h);                                   // Z3 = ((Z1 + Z2)^2 - Z1Z1 - Z2Z2) H
	fe_sqr(t, r);
	fe_sub(t, t, j), fe_sub(t, t, v), fe_sub(t, t, v);  // X3 = r^2 - J - 2V
	fe x3 = t;
	fe_sub(t, v, x3), fe_mul(t, r, t);
	fe_mul(j, s1, j), fe_dbl(j, j);
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
