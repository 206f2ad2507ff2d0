// host_gfp_test.cu — drives bee2_b200/csrc/gfp.cuh (and ecp.cuh) on the CPU through the portable
// twins of the carry-chain primitives, so that the field / point LOGIC (row order of the squaring,
// lazy folds, inversion chains, exceptional cases of the additions) is checked where there is no GPU.
// Protocol: one command per stdin line, "op N hex..." -> one hex result per stdout line.
// Driven by tests/test_host_gfp.py, which holds the expected values (Python integers).
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#include <sstream>
#include "../bee2_b200/csrc/ecp.cuh"

template <int N> static void from_hex(u32* v, const std::string& h, int limbs = N)
{
	for (int i = 0; i < limbs; ++i) v[i] = 0;
	int nib = 0;
	for (int i = (int)h.size() - 1; i >= 0 && nib < 8 * limbs; --i, ++nib)
	{
		const char c = h[i];
		const u32 d = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
		v[nib >> 3] |= d << (4 * (nib & 7));
	}
}
static std::string to_hex(const u32* v, int limbs)
{
	std::string s;
	char buf[16];
	for (int i = limbs - 1; i >= 0; --i) snprintf(buf, sizeof buf, "%08x", v[i]), s += buf;
	return s;
}

template <int N> static std::string run(const std::string& op, std::istringstream& in)
{
	std::string ha, hb, hc, hd;
	fe<N> a, b, r;
	if (op == "mul" || op == "add" || op == "sub")
	{
		in >> ha >> hb;
		from_hex<N>(a.v, ha), from_hex<N>(b.v, hb);
		if (op == "mul") fe_mul<N>(r, a, b);
		else if (op == "add") fe_add<N>(r, a, b);
		else fe_sub<N>(r, a, b);
		return to_hex(r.v, N);
	}
	if (op == "sub2")
	{
		in >> ha >> hb >> hc;
		fe<N> c;
		from_hex<N>(a.v, ha), from_hex<N>(b.v, hb), from_hex<N>(c.v, hc);
		fe_sub2<N>(r, a, b, c);
		return to_hex(r.v, N);
	}
	if (op == "invbatches")
	{
		// batches of 30 division steps the early-exit form needs, and the scheduled count of the fixed form
		in >> ha;
		from_hex<N>(a.v, ha);
		fe_canon<N>(a);
		int used = 0;
		inv_safegcd<N, false>(r.v, a.v, fe_param<N>::C, &used);
		return std::to_string(used) + " " + std::to_string(inv30<N>::BATCHES);
	}
	if (op == "mul3")
	{
		in >> ha;
		from_hex<N>(a.v, ha);
		fe_mul3<N>(r, a);
		return to_hex(r.v, N);
	}
	if (op == "sqr" || op == "inv" || op == "shl1" || op == "shl2" || op == "shl3" || op == "canon" || op == "iszero")
	{
		in >> ha;
		from_hex<N>(a.v, ha);
		if (op == "sqr") fe_sqr<N>(r, a);
		else if (op == "inv")
		{
			// the division-step form (fixed step count and early exit) and the power a^(p-2) must agree
			fe<N> r2, r3;
			fe_inv<N, true>(r, a), fe_inv<N, false>(r2, a);
			r3 = fe_inv_fermat_fn<N>(a), fe_canon<N>(r3);
			for (int i = 0; i < N; ++i)
				if (r.v[i] != r2.v[i] || r.v[i] != r3.v[i]) return "inv-mismatch";
		}
		else if (op == "shl1") fe_shl<1, N>(r, a);
		else if (op == "shl2") fe_shl<2, N>(r, a);
		else if (op == "shl3") fe_shl<3, N>(r, a);
		else if (op == "canon") r = a, fe_canon<N>(r);
		else return fe_is_zero<N>(a) ? "1" : "0";
		return to_hex(r.v, N);
	}
	if (op == "mulwide" || op == "sqrwide")
	{
		u32 t[2 * N];
		in >> ha;
		from_hex<N>(a.v, ha);
		if (op == "mulwide")
		{
			in >> hb;
			from_hex<N>(b.v, hb);
			fe_mul_wide<N>(t, a.v, b.v);
		}
		else
			fe_sqr_wide<N>(t, a.v);
		return to_hex(t, 2 * N);
	}
	// points: Jacobian triples X Y Z (hex), result X Y Z
	if (op == "pdbl" || op == "padd" || op == "pmadd")
	{
		pt<N> P, Q, R;
		in >> ha >> hb >> hc;
		from_hex<N>(P.X.v, ha), from_hex<N>(P.Y.v, hb), from_hex<N>(P.Z.v, hc);
		if (op == "pdbl")
			pt_dbl<N>(R, P);
		else if (op == "padd")
		{
			in >> ha >> hb >> hc;
			from_hex<N>(Q.X.v, ha), from_hex<N>(Q.Y.v, hb), from_hex<N>(Q.Z.v, hc);
			pt_add<N>(R, P, Q);
		}
		else
		{
			in >> ha >> hb;
			from_hex<N>(Q.X.v, ha), from_hex<N>(Q.Y.v, hb);
			pt_madd<N>(R, P, Q.X, Q.Y);
		}
		return to_hex(R.X.v, N) + " " + to_hex(R.Y.v, N) + " " + to_hex(R.Z.v, N);
	}
	// scalar multiplication k * (x, y): "pmul N nbits k x y" -> affine "x y" or "inf"
	if (op == "pmul")
	{
		int nbits;
		in >> nbits >> ha >> hb >> hc;
		sc<N> k;
		fe<N> x, y;
		from_hex<N>(k.w, ha), from_hex<N>(x.v, hb), from_hex<N>(y.v, hc);
		pt<N> R;
		pt_mul_var<N>(R, k, nbits, x, y);
		{
			// the regular (secret-scalar) form must land on the same point, limb for limb
			pt<N> C;
			pt_mul_var<N, true>(C, k, nbits, x, y);
			const bool same_inf = pt_is_inf<N>(R) == pt_is_inf<N>(C);
			bool same = same_inf;
			if (same && !pt_is_inf<N>(R))
				for (int i = 0; i < N; ++i)
					same = same && R.X.v[i] == C.X.v[i] && R.Y.v[i] == C.Y.v[i] && R.Z.v[i] == C.Z.v[i];
			if (!same) return "ct-mismatch";
		}
		{
			// the form with the window table in a caller-provided scratch area (win_global: the layout of the
			// verification kernel, here as thread 5 of a pretend CTA of 7 threads) must land there too
			constexpr int T = 7, tid = 5;
			static u8 area[win_global<N>::ITEM_BYTES * T + 64];
			u8* al = area + ((32 - ((uintptr_t)area & 31)) & 31);
			memset(area, 0xEE, sizeof area);
			win_global<N> W;
			W.base = al + 32 * tid, W.stride = 32 * T;
			pt<N> G;
			pt_mul_var_g<N>(G, k, nbits, x, y, W);
			bool same = pt_is_inf<N>(R) == pt_is_inf<N>(G);
			if (same && !pt_is_inf<N>(R))
				for (int i = 0; i < N; ++i)
					same = same && R.X.v[i] == G.X.v[i] && R.Y.v[i] == G.Y.v[i] && R.Z.v[i] == G.Z.v[i];
			if (!same) return "wg-mismatch";
			// only this thread's granules were written: every other 32-octet granule is untouched
			for (size_t g = 0; g < win_global<N>::ITEM_BYTES * T / 32; ++g)
				if ((int)(g % T) != tid)
					for (int b = 0; b < 32; ++b)
						if (al[32 * g + b] != 0xEE) return "wg-overrun";
		}
		if (pt_is_inf<N>(R)) return "inf";
		pt_to_affine<N>(x, y, R);
		return to_hex(x.v, N) + " " + to_hex(y.v, N);
	}
	return "?";
}

int main()
{
	std::string line;
	while (std::getline(std::cin, line))
	{
		std::istringstream in(line);
		std::string op;
		int n;
		if (!(in >> op >> n)) continue;
		std::string out = n == 8 ? run<8>(op, in) : n == 12 ? run<12>(op, in) : n == 16 ? run<16>(op, in) : "?";
		puts(out.c_str());
	}
	return 0;
}
