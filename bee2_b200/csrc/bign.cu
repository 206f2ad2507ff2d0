// bign.cu — bign (STB 34.101.45) on the three standard curves bign-curve256v1 / 384v1 / 512v1
// (levels l = 128 / 192 / 256, fields of N = 8 / 12 / 16 limbs): batch verify / sign2 /
// pubkey-calc / scalar multiplication kernels for sm_100a + C-ABI launchers.
//
// Replaces bignVerifyEc (bign_sign.c:268-347) incl. ecAddMulA (ec.c:1183-1273), ecpToAJ and
// belt-hash; bignSign2Ec (bign_sign.c:140-245) incl. bignMulBase/ecMulPreOD (bign_misc.c:115-137,
// ec.c:892-964) and beltWBL (belt_wbl.c:50-82); bignPubkeyCalc (bign_misc.c:369-412);
// ecMulA (ec.c:497-525). The same code serves bign128 / bign192 / bign256 (bign128.c, bign192.c,
// bign256.c), which only fix the level and the hash OID.
//
// Work decomposition: ONE THREAD PER ITEM (signature / key / scalar-point pair), field
// elements as N x u32 in registers. Scalar multiplication is REGULAR so that all lanes of a
// warp execute the same doublings and additions in lock-step:
//   * fixed base G: 16 / 13 / 13-bit windows over a device-resident table GTAB[GN][2^w] of
//     affine points (d + 2^w) * 2^(w i) * G ("offset windows": with the scalar k - K, K = sum 2^w 2^(w i),
//     no digit means the point at infinity; 67 / 24 / 42 MiB; generated once per process, device and level
//     by bign_gtab_kernel with the same point code) -> 16 / 30 / 40 unconditional mixed additions, no doublings;
//   * variable base Q: signed 5-bit windows, per-thread table {1..16}Q in local memory
//     (ecp.cuh pt_mul_var) -> 5 doublings + 1 addition per window.
// The reference's interleaved wNAF (ec.c:1206-1268) is irregular and would diverge.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include "ecp.cuh"
#include "belt_dev.cuh"

// CTA shape for N = 8 (bign-curve256v1): 256 threads x 3 CTAs/SM (80 registers). One field inversion is
// shared by the whole CTA (its 7 other warps wait meanwhile), so larger CTAs amortise it and more CTAs per
// SM hide it: measured 64 x 8 -> 45.8, 128 x 4 -> 49.4, 128 x 5 -> 50.7, 128 x 6 -> 48.9, 256 x 2 -> 51.3,
// 256 x 3 -> 52.2, 256 x 4 -> 51.8, 512 x 1 -> 45.0 M verifies/s (without any inversion: 54.4). The wider
// fields keep 128 threads x 4 CTAs/SM (128 registers; 2 -> 14.5 / 6.9, 3 -> 13.7 / 6.3, 4 -> 15.1 / 7.1 M
// verifies/s for N = 12 / 16 on 2^16 items) and a 12 / 16 KiB product tree per CTA.
#ifndef BIGN_THREADS
#define BIGN_THREADS 256
#endif
#ifndef BIGN_MIN_BLOCKS
#define BIGN_MIN_BLOCKS 3
#endif
#define BIGN_T(N) ((N) == 8 ? BIGN_THREADS : 128)
// resident CTAs per SM asked of the compiler: 4 x 128 threads x 128 registers for N = 8; the wider
// fields need more registers per thread
#ifndef BIGN_BLOCKS12
#define BIGN_BLOCKS12 4
#endif
#ifndef BIGN_BLOCKS16
#define BIGN_BLOCKS16 4
#endif
#define BIGN_BLOCKS(N) ((N) == 8 ? BIGN_MIN_BLOCKS : (N) == 12 ? BIGN_BLOCKS12 : BIGN_BLOCKS16)
// the signing kernel's own occupancy request (measured at l = 128: 3 CTAs/SM 242 M/s, 4 -> 229, 2 -> see DESIGN §4.3.1)
#ifndef BIGN_SIGN_BLOCKS8
#define BIGN_SIGN_BLOCKS8 BIGN_MIN_BLOCKS
#endif
#define BIGN_SIGN_BLOCKS(N) ((N) == 8 ? BIGN_SIGN_BLOCKS8 : BIGN_BLOCKS(N))
// Fixed-base window width in bits (<= 16), per field size. N = 8: 16 bits -> 16 mixed additions from a 67 MiB
// table (measured against 13 bits / 20 additions / 10 MiB: verify 52.7 -> 53.9 M/s, sign2 202 -> 218 M/s; the
// random 64-byte table reads miss L2 more often — 268 MB of extra DRAM reads per 2^18 items, far from binding).
// The wider fields keep 13 bits (24 / 42 MiB): their 16-bit tables would be 151 / 268 MiB and take ~0.5 s to build.
#ifndef BIGN_GW8
#define BIGN_GW8 16
#endif
#ifndef BIGN_GW
#define BIGN_GW 13
#endif
#define BIGN_GWN(N) ((N) == 8 ? BIGN_GW8 : BIGN_GW)
#define BIGN_GN(N) ((32 * (N) + BIGN_GWN(N) - 1) / BIGN_GWN(N))   // number of windows
#define BIGN_GE(N) (1 << BIGN_GWN(N))               // entries per window (entry 0 unused)
#define BIGN_MAX_OID 64
#define BIGN_MAX_T 64

// The device code up to and including the verification kernel is compiled twice: by bign.cu itself
// (out-of-line field products, 3 CTAs/SM at 80 registers — the issue-bound shape for full grids) and, inside
// namespace bign_lowocc, by bign_lowocc.cu (the same source with every product inlined and 1 CTA/SM — the
// latency-bound shape for grids of under ~2 warps per scheduler).
// (Only the second build is wrapped: putting bign.cu's own copy into a namespace as well changed nothing but
// the mangled names, yet ptxas laid the functions out differently and the full-grid kernel lost 3 % —
// 4.979 against 4.831 ms for 2^18 items, same box, A/B — which says how close to the instruction-cache
// edge this 166 KB kernel runs.)
struct OidArg { u8 der[BIGN_MAX_OID]; u32 len; };
struct TArg { u8 t[BIGN_MAX_T]; u32 len; };
#ifdef BIGN_LOWOCC_TU
namespace bign_lowocc {
#endif
// q and the y-coordinate of G = (0, yG), little-endian limbs
// (bign_params.c:61-73 curve256v1, :110-125 curve384v1, :169-190 curve512v1)
__constant__ u32 c_q8[8] = {0x263D6607u, 0x7E5ABF99u, 0x0DFB4DFCu, 0xD95C8ED6u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
__constant__ u32 c_yG8[8] = {0x04516A93u, 0x1E29CF18u, 0xC408F652u, 0x78913966u, 0x51D6835Du, 0x5CE4C9A3u, 0xFB16D69Fu, 0x6BF7FC3Cu};
__constant__ u32 c_q12[12] = {0xF30CA7B7u, 0x3DB7DC3Fu, 0xA6A4FF0Au, 0x8046DAE7u, 0x73AF7BBBu, 0x6CCCC403u, 0xFFFFFFFEu, 0xFFFFFFFFu,
	0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
__constant__ u32 c_yG12[12] = {0xF733C451u, 0xEA5ECB31u, 0x6B2A42F9u, 0x84403E27u, 0x66B1D355u, 0x0549E79Eu, 0xDC86FFA0u, 0x3A729A11u,
	0x432DBF89u, 0x6330117Eu, 0xA82E9E9Eu, 0x5D438224u};
__constant__ u32 c_q16[16] = {0x0D068EF1u, 0xDCFFAD49u, 0x9556DF32u, 0x361BCAE5u, 0x2E2113F4u, 0xF26BEBB0u, 0x0198004Eu, 0xB2C0092Cu,
	0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
__constant__ u32 c_yG16[16] = {0xCEEFEDBDu, 0xB792AE6Fu, 0xC94C0D04u, 0x67AA83B9u, 0xEEE82261u, 0xFF777395u, 0x0EFA6FD2u, 0x6973DDE2u,
	0x00CCCADAu, 0xD2EDF81Bu, 0xB361BCE2u, 0xB0AB41B3u, 0xA0D18FABu, 0xB182E6F7u, 0xE4037681u, 0xA826FF7Au};

// the coefficient b (bign_params.c:50-55, :94-101, :151-160): only the on-curve check of bignPubkeyVal / bignDH uses it
__constant__ u32 c_b8[8] = {0xD69C03F1u, 0xB22E7D6Bu, 0x978B9253u, 0x4CF55069u, 0xE4D8FBBEu, 0xD2C13AABu, 0x15F3A8EDu, 0x77CE6C15u};
__constant__ u32 c_b12[12] = {0x6873BF64u, 0xBCA7FC23u, 0xF3CEBD7Cu, 0x14BDE2F0u, 0xE9712E3Au, 0xA6216AF9u, 0x0FFBB196u, 0x712748BBu, 0x655D34D2u, 0x33075AABu, 0x959CEF20u, 0x3C75DFE1u};
__constant__ u32 c_b16[16] = {0xD6139C90u, 0x09346998u, 0x3A49A27Au, 0xEA862227u, 0x87ACA243u, 0x2933008Cu, 0xC4245E95u, 0x2711DCB5u, 0xDAADB088u, 0x17CE13E3u, 0xDD5D2551u, 0x5BC6A9EEu, 0x60FD5889u, 0xD88C5D6Au, 0x933B8C43u, 0x6CB45944u};

template <int N> struct bign_c;
template <> struct bign_c<8>
{
	static __device__ __forceinline__ const u32* q() { return c_q8; }
	static __device__ __forceinline__ const u32* yG() { return c_yG8; }
	static __device__ __forceinline__ const u32* b() { return c_b8; }
};
template <> struct bign_c<12>
{
	static __device__ __forceinline__ const u32* q() { return c_q12; }
	static __device__ __forceinline__ const u32* yG() { return c_yG12; }
	static __device__ __forceinline__ const u32* b() { return c_b12; }
};
template <> struct bign_c<16>
{
	static __device__ __forceinline__ const u32* q() { return c_q16; }
	static __device__ __forceinline__ const u32* yG() { return c_yG16; }
	static __device__ __forceinline__ const u32* b() { return c_b16; }
};

// device: BIGN_GN(N) * BIGN_GE entries of 8 N octets (x || y) per level; entry j = 0 unused
// one table per device the engine runs on (engine.c: per-device contexts) and level
#define BIGN_MAX_DEV 16
static std::atomic<uint4*> g_gtab[BIGN_MAX_DEV][3];

// S-box policy of the belt code inside the bign kernels (belt-hash, belt-WBL)
#ifndef BIGN_SBOX
#define BIGN_SBOX BeltT4
#endif
typedef BIGN_SBOX BignSbox;

// ---------------------------------------------------------------- small helpers
// q as a register array (constant-bank operands)
template <int N> __device__ __forceinline__ void load_q(u32* q)
{
#pragma unroll
	for (int i = 0; i < N; ++i) q[i] = bign_c<N>::q()[i];
}
template <int N> __device__ __forceinline__ bool geq_q(const u32* a)
{
	u32 q[N];
	load_q<N>(q);
	return uN_geq<N>(a, q);
}
// r = (a + b) mod q for a, b < q (zz_mod.c:42)
template <int N> __device__ __forceinline__ void modq_add(u32* r, const u32* a, const u32* b)
{
	u32 t[N], u[N], q[N];
	load_q<N>(q);
	const u32 c = add_n<N>(t, a, b);
	const u32 m = sub_n<N>(u, t, q);
	const bool take = c || m == 0;   // carried out, or t >= q
#pragma unroll
	for (int i = 0; i < N; ++i) r[i] = take ? u[i] : t[i];
}
// r = (a - b) mod 2^(32N), + q if it borrowed — zzSubMod without range assumptions (zz_mod.c:120)
template <int N> __device__ __forceinline__ void modq_sub(u32* r, const u32* a, const u32* b)
{
	u32 t[N], u[N], q[N];
	load_q<N>(q);
	const u32 m = sub_n<N>(t, a, b);
	(void)add_n<N>(u, t, q);
#pragma unroll
	for (int i = 0; i < N; ++i) r[i] = m ? u[i] : t[i];
}

// r = x mod q for x of N + N/2 + 1 limbs — the product (s0 + 2^l) d of the signature equation (zzMod,
// bign_sign.c:234). q = 2^(32N) - c with a short c (4 / 7 / 8 limbs on the three curves: the top half of q
// is all ones, bign_params.c:61-66, :110-117, :169-180), so x = lo + hi 2^(32N) = lo + hi c (mod q):
// one fold with the (N/2 + 1)-limb hi, two more with the single limb that is left, one conditional
// subtraction. Fixed shapes, fully unrolled, everything in registers (the first version — generic loops over
// local-memory arrays — was 15 % of the instructions of a signature).
template <int N> struct bign_qc { static constexpr int CL = N == 8 ? 4 : N == 12 ? 7 : 8; };
template <int N> __device__ __forceinline__ void modq_reduce(u32* r, const u32* x)
{
	constexpr int HL = N / 2 + 1, CL = bign_qc<N>::CL, PL = HL + CL, TL = (PL > N ? PL : N) + 1;
	u32 c[CL];
	{
		// c = -q mod 2^(32N): its limbs above CL are zero
		u32 cy = 1;
#pragma unroll
		for (int i = 0; i < CL; ++i)
		{
			const u64 v = (u64)(~bign_c<N>::q()[i]) + cy;
			c[i] = (u32)v, cy = (u32)(v >> 32);
		}
	}
	// P = hi * c, schoolbook: the carry of row a opens limb a + CL
	u32 P[PL];
#pragma unroll
	for (int i = 0; i < PL; ++i) P[i] = 0;
#pragma unroll
	for (int a = 0; a < HL; ++a)
	{
		u64 cy = 0;
#pragma unroll
		for (int b = 0; b < CL; ++b)
		{
			const u64 v = (u64)x[N + a] * c[b] + P[a + b] + cy;
			P[a + b] = (u32)v, cy = v >> 32;
		}
		P[a + CL] = (u32)cy;
	}
	// t = lo + P
	u32 t[TL];
	{
		u64 cy = 0;
#pragma unroll
		for (int i = 0; i < TL; ++i)
		{
			const u64 v = (u64)(i < N ? x[i] : 0u) + (i < PL ? P[i] : 0u) + cy;
			t[i] = (u32)v, cy = v >> 32;
		}
	}
	// what is left above 2^(32N) is a few bits in limb N (t < 2^(32N + 3)); after two more folds it is gone:
	// t < 2^(32N) + 7c -> h <= 1 and then lo < 8c -> lo + c < 2^(32N)
#pragma unroll
	for (int round = 0; round < 2; ++round)
	{
		const u32 h = t[N];
		u64 cy = 0;
#pragma unroll
		for (int i = 0; i < N; ++i)
		{
			const u64 v = (i < CL ? (u64)h * c[i] : 0ull) + t[i] + cy;
			t[i] = (u32)v, cy = v >> 32;
		}
		t[N] = (u32)cy;
	}
	// t < 2^(32N) = q + c: at most one subtraction of q
	{
		u32 u[N], q[N];
		load_q<N>(q);
		const u32 m = sub_n<N>(u, t, q);
#pragma unroll
		for (int i = 0; i < N; ++i) r[i] = m ? t[i] : u[i];
	}
}

// r = (t + h 2^(32N)) mod q for an N-limb t and a small h (the scalars of the fixed-base table entries)
template <int N> __device__ __forceinline__ void modq_fold_top(u32* r, const u32* t, u32 h)
{
	constexpr int CL = bign_qc<N>::CL;
	u32 c[CL], x[N + 1];
	{
		u32 cy = 1;
#pragma unroll
		for (int i = 0; i < CL; ++i)
		{
			const u64 v = (u64)(~bign_c<N>::q()[i]) + cy;
			c[i] = (u32)v, cy = (u32)(v >> 32);
		}
	}
#pragma unroll
	for (int i = 0; i < N; ++i) x[i] = t[i];
	x[N] = h;
	// h c < 2^(32 CL + 32) <= 2^(32N): after the first fold at most one bit is left above, after the second none
#pragma unroll
	for (int round = 0; round < 3; ++round)
	{
		const u32 hh = x[N];
		u64 cy = 0;
#pragma unroll
		for (int i = 0; i < N; ++i)
		{
			const u64 v = (i < CL ? (u64)hh * c[i] : 0ull) + x[i] + cy;
			x[i] = (u32)v, cy = v >> 32;
		}
		x[N] = (u32)cy;
	}
	u32 u[N], q[N];
	load_q<N>(q);
	const u32 m = sub_n<N>(u, x, q);
#pragma unroll
	for (int i = 0; i < N; ++i) r[i] = m ? x[i] : u[i];
}

template <int N> __device__ __forceinline__ void load_uN(u32* r, const u8* p)
{
	fe<N> t;
	fe_load<N>(t, p);
#pragma unroll
	for (int i = 0; i < N; ++i) r[i] = t.v[i];
}

// d_len <= 4N little-endian octets -> limbs, without dynamic indexing of the destination
template <int N> __device__ __forceinline__ void load_scalar(sc<N>& k, const u8* s, u32 d_len)
{
#pragma unroll
	for (int l = 0; l < N; ++l)
	{
		u32 w = 0;
#pragma unroll
		for (int b = 0; b < 4; ++b)
			if ((u32)(4 * l + b) < d_len)
				w |= (u32)s[4 * l + b] << (8 * b);
		k.w[l] = w;
	}
}

// ---------------------------------------------------------------- fixed-base multiplication
// acc (+)= k * G for a 32N-bit k (little-endian limbs) through the fixed-base window table.
// OFFSET WINDOWS: entry (i, d) of the table is not d 2^(w i) G but (d + 2^w) 2^(w i) G, and the scalar that is
// cut into w-bit digits is k' = k - K mod q with K = sum_i 2^w 2^(w i): then sum_i entry(i, d_i(k')) =
// (k' + K) G = k G. No digit selects the point at infinity, so every window is one unconditional mixed
// addition — no skipped zero digits (the public-scalar form of round 1), no dummy addition dropped by a masked
// select (the regular form the secret scalars used since: it kept the old and the new accumulator alive across
// the eleven product calls of the addition; that glue was 17 % of the instructions of a signature). The
// instruction stream is the same for every scalar; the reference's bignMulBase is regular too (ec.c:892-964).
// What still depends on the scalar is the ADDRESS of the table read (a cache-timing channel the reference's
// masked wwSel scan does not have; scanning 2^w entries per window is not an option here, DESIGN.md "Secret
// scalars") and the exceptional branches of the complete addition (probability ~2^-250 per addition).
// FRESH: acc is the point at infinity on entry and need not be initialised — window 0 only loads its entry.
// K (N limbs, < q) lies behind the table.
#define BIGN_GTAB_UINT4(N) ((size_t)BIGN_GN(N) * BIGN_GE(N) * ((N) / 2))
template <int N, bool FRESH = false> __device__ __noinline__ void pt_add_mul_base(pt<N>& acc, const sc<N> ks, const uint4* __restrict__ gtab)
{
	u32 k[N];
	{
		u32 K[N];
		const u32* Kp = reinterpret_cast<const u32*>(gtab + BIGN_GTAB_UINT4(N));
#pragma unroll
		for (int i = 0; i < N; ++i) K[i] = __ldg(Kp + i);
		// (k - K) mod 2^(32N), + q if it borrowed: congruent to k - K for every k < 2^(32N) (K < q)
		modq_sub<N>(k, ks.w, K);
	}
#pragma unroll 1
	for (int i = 0; i < BIGN_GN(N); ++i)
	{
		const int bit = BIGN_GWN(N) * i, limb = bit >> 5;
		const u64 w = (u64)k[limb] | (limb < N - 1 ? (u64)k[limb + 1] << 32 : 0);
		const u32 d = (u32)(w >> (bit & 31)) & (BIGN_GE(N) - 1);
		const uint4* e = gtab + ((size_t)i * BIGN_GE(N) + d) * (N / 2);
		fe<N> x, y;
#pragma unroll
		for (int j = 0; j < N / 4; ++j)
		{
			const uint4 a = __ldg(e + j), b = __ldg(e + N / 4 + j);
			x.v[4 * j] = a.x, x.v[4 * j + 1] = a.y, x.v[4 * j + 2] = a.z, x.v[4 * j + 3] = a.w;
			y.v[4 * j] = b.x, y.v[4 * j + 1] = b.y, y.v[4 * j + 2] = b.z, y.v[4 * j + 3] = b.w;
		}
		if (FRESH && i == 0)
			pt_set_affine<N>(acc, x, y);
		else
			pt_madd<N>(acc, acc, x, y);
	}
}

// ---------------------------------------------------------------- belt-hash(oid || a || b || extra)
// a, b: field-sized (4N octets) little-endian limb strings; b and extra optional
template <int N> __device__ __forceinline__ void hash_oid_ab(const BignSbox& S, u32 (&out)[8], const OidArg& oid,
	const u32* a, const u32* b, const u8* extra, u32 extra_len)
{
	// message zero-padded to whole 32-octet blocks, assembled word-wise: the DER string (zero-padded by
	// make_oid) is copied as words, everything after it is shifted in by the octet phase of its position
	u32 msg[(BIGN_MAX_OID + 8 * N + BIGN_MAX_T + 31) / 32 * 8];
	const int nw = (int)(sizeof(msg) / 4);
#pragma unroll
	for (int i = 0; i < nw; ++i) msg[i] = i < BIGN_MAX_OID / 4 ? reinterpret_cast<const u32*>(oid.der)[i] : 0u;
	u32 pos = oid.len;
	const u32 sh = 8 * (pos & 3);
	const auto append = [&](const u32* w, int nwords)
	{
		const u32 i0 = pos >> 2;
		if (sh == 0)
		{
			for (int j = 0; j < nwords; ++j) msg[i0 + j] = w[j];
		}
		else
		{
			u32 lo = msg[i0];
			for (int j = 0; j < nwords; ++j)
				msg[i0 + j] = lo | (w[j] << sh), lo = w[j] >> (32 - sh);
			msg[i0 + nwords] = lo;
		}
		pos += 4 * nwords;
	};
	append(a, N);
	if (b)
		append(b, N);
	if (extra_len)
	{
		// t is zero-padded to BIGN_MAX_T octets (TArg): whole words may be appended, the length counts octets
		append(reinterpret_cast<const u32*>(extra), (int)((extra_len + 3) >> 2));
		pos = pos - 4 * ((extra_len + 3) >> 2) + extra_len;
	}
	const belt_digest dg = belt_hash_words(S, msg, pos);
#pragma unroll
	for (int i = 0; i < 8; ++i) out[i] = dg.w[i];
}

// ---------------------------------------------------------------- block-wide inversion
// Montgomery's simultaneous inversion as a product tree in shared memory: every thread of the CTA
// hands in one z != 0 (1 if it has nothing to invert) and gets 1/z back, for ONE round of field inversions
// per CTA plus 2 (log2(CTA size) - 5) products per thread — instead of one inversion per thread, which was
// 12 % of a verification and half of a signature.
// Node i has children 2i and 2i+1, leaves at T + tid (T = CTA size). The tree stops at the level of 32
// nodes (32 .. 63): the 32 lanes of warp 0 invert one node each, in lock-step — the latency of one
// inversion, five tree levels (ten dependent products and fifteen barriers) fewer than a tree that ends in a
// single root, and 32 different operands keep the iteration on the vector pipe (with one active thread
// ptxas proves the operand warp-uniform and moves the whole inversion to the uniform datapath — UIMAD /
// ULOP3 / USHF — which runs this dependent chain slower: signing 242 -> 236 M/s, measured).
// The tree is stored word-major (word j of node i at sm[j * 2 T + i]) so that lanes hit distinct banks.
// Must be reached by ALL threads of the CTA (it synchronises).
#define BIGN_TREE_WORDS(N) (2 * BIGN_T(N) * (N))
template <int N> __device__ __forceinline__ void tree_put(u32* sm, int i, const fe<N>& a)
{
#pragma unroll
	for (int j = 0; j < N; ++j) sm[j * (2 * BIGN_T(N)) + i] = a.v[j];
}
template <int N> __device__ __forceinline__ void tree_get(fe<N>& a, const u32* sm, int i)
{
#pragma unroll
	for (int j = 0; j < N; ++j) a.v[j] = sm[j * (2 * BIGN_T(N)) + i];
}
// CT = false: the root inversion may stop early (public inputs: verification)
template <int N, bool CT = true> __device__ __noinline__ fe<N> block_inv(const fe<N> z, u32* sm)
{
	const int tid = threadIdx.x;
	fe<N> a, b;
#ifdef BIGN_FAKE_TREE   /* timing experiment only: no tree, no barriers, no inversion (wrong results) */
	return z;
#endif
	tree_put<N>(sm, BIGN_T(N) + tid, z);
	// a CTA may be launched with fewer than BIGN_T(N) threads (bign_shape: balanced grids for small
	// batches), at least BIGN_T(N) / 2: the missing leaves are 1
	if (tid + (int)blockDim.x < BIGN_T(N))
	{
		fe_set_u32<N>(a, 1);
		tree_put<N>(sm, BIGN_T(N) + tid + (int)blockDim.x, a);
	}
	__syncthreads();
	// up: products of the children
#pragma unroll 1
	for (int s = BIGN_T(N) / 2; s >= 32; s >>= 1)
	{
		if (tid < s)
		{
			tree_get<N>(a, sm, 2 * (s + tid)), tree_get<N>(b, sm, 2 * (s + tid) + 1);
			fe_mul<N>(a, a, b);
			tree_put<N>(sm, s + tid, a);
		}
		__syncthreads();
	}
#ifndef BIGN_FAKE_INV   /* timing experiment only: skips the inversion (wrong results) */
	if (tid < 32)
	{
		tree_get<N>(a, sm, 32 + tid);
		fe_inv<N, CT>(a, a);
		tree_put<N>(sm, 32 + tid, a);
	}
#endif
	__syncthreads();
	// down: 1/child = 1/parent * sibling
#pragma unroll 1
	for (int s = 64; s <= BIGN_T(N); s <<= 1)
	{
		const bool on = tid < s;
		if (on)
		{
			tree_get<N>(a, sm, (s + tid) >> 1), tree_get<N>(b, sm, (s + tid) ^ 1);
			fe_mul<N>(a, a, b);
		}
		__syncthreads();
		if (on)
			tree_put<N>(sm, s + tid, a);
		__syncthreads();
	}
	tree_get<N>(a, sm, BIGN_T(N) + tid);
	return a;
}
// affine x (and y) of a point from the inverse of its Z (ecp_j.c:104-133), canonical residues
template <int N> __device__ __forceinline__ void pt_affine_x_zi(fe<N>& x, const pt<N>& P, const fe<N>& zi)
{
	fe<N> zi2;
	fe_sqr<N>(zi2, zi);
	fe_mul<N>(x, P.X, zi2);
	fe_canon<N>(x);
}
template <int N> __device__ __forceinline__ void pt_affine_xy_zi(fe<N>& x, fe<N>& y, const pt<N>& P, const fe<N>& zi)
{
	fe<N> zi2;
	fe_sqr<N>(zi2, zi);
	fe_mul<N>(x, P.X, zi2);
	fe_mul<N>(zi2, zi2, zi), fe_mul<N>(y, P.Y, zi2);
	fe_canon<N>(x), fe_canon<N>(y);
}

// ---------------------------------------------------------------- kernels
// Table of fixed-base multiples: entry (i, j) = j * 2^(BIGN_GW i) * G, affine.
// scalar of table entry (i, j): (j + 2^w) 2^(w i) mod q — never 0 (q is a prime above 2^(w + 1))
template <int N> __device__ __forceinline__ void gtab_scalar(sc<N>& k, int i, u32 j)
{
	const int bit = BIGN_GWN(N) * i, limb = bit >> 5;
	const u64 w = (u64)(j + BIGN_GE(N)) << (bit & 31);   // 17 + 31 bits at most
	u32 t[N];
#pragma unroll
	for (int l = 0; l < N; ++l)
		t[l] = l == limb ? (u32)w : (l == limb + 1 ? (u32)(w >> 32) : 0u);
	// what does not fit the N limbs folds back: 2^(32N) = c (mod q)
	modq_fold_top<N>(k.w, t, limb + 1 == N ? (u32)(w >> 32) : 0u);
}
// K = sum_i 2^w 2^(w i) mod q, stored behind the table (pt_add_mul_base)
template <int N> __global__ void bign_gtab_k_kernel(uint4* gtab)
{
	u32 K[N];
#pragma unroll
	for (int l = 0; l < N; ++l) K[l] = 0;
	for (int i = 0; i < BIGN_GN(N); ++i)
	{
		sc<N> t;
		gtab_scalar<N>(t, i, 0);
		modq_add<N>(K, K, t.w);
	}
	u32* out = reinterpret_cast<u32*>(gtab + BIGN_GTAB_UINT4(N));
	for (int l = 0; l < N; ++l) out[l] = K[l];
}
template <int N> __global__ void __launch_bounds__(128) bign_gtab_kernel(uint4* gtab)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= BIGN_GN(N) * BIGN_GE(N))
		return;
	const int i = idx / BIGN_GE(N), j = idx % BIGN_GE(N);
	uint4* e = gtab + (size_t)idx * (N / 2);
	sc<N> k;
	gtab_scalar<N>(k, i, (u32)j);
	fe<N> gx, gy, x, y;
	fe_set_u32<N>(gx, 0);
#pragma unroll
	for (int l = 0; l < N; ++l) gy.v[l] = bign_c<N>::yG()[l];
	pt<N> acc;
	pt_mul_var<N>(acc, k, 32 * N, gx, gy);
	pt_to_affine<N>(x, y, acc);
#pragma unroll
	for (int l = 0; l < N / 4; ++l)
	{
		e[l] = make_uint4(x.v[4 * l], x.v[4 * l + 1], x.v[4 * l + 2], x.v[4 * l + 3]);
		e[N / 4 + l] = make_uint4(y.v[4 * l], y.v[4 * l + 1], y.v[4 * l + 2], y.v[4 * l + 3]);
	}
}

// bignVerifyEc per item (bign_sign.c:268-347); no = 4N octets: hash no, sig no/2 + no, pubkey 2 no.
// STAGED: the CTA's inputs (three contiguous segments: hashes, signatures, public keys of its items) are
// brought into shared memory by three TMA bulk copies (cp.async.bulk -> mbarrier) issued by one thread —
// the "TMA staging of point batches" of the north star. It pays where the inputs are PINNED HOST memory
// (bignVerifyBatch's zero-copy path): the copy engine reads them in large PCIe requests, whereas per-thread
// loads fetch one 32-byte sector per request (measured end to end, 2^18 items: 46.0 M/s with direct loads).
// The staging buffer is the memory of the product tree, which is only used after the inputs are in registers.
#define BIGN_STAGE_BYTES(N) (BIGN_T(N) * 18 * (N))   /* 4.5 no octets per item */
#define BIGN_SMEM_BYTES(N) (BIGN_STAGE_BYTES(N) > 4 * BIGN_TREE_WORDS(N) ? BIGN_STAGE_BYTES(N) : 4 * BIGN_TREE_WORDS(N))
// (BIGN_VERIFY_T8 / BIGN_VERIFY_B8: launch bounds of the l = 128 verification kernel alone, for occupancy
// experiments — e.g. 224 threads x 4 CTAs = 72 registers, to be launched with B2G_BIGN_THREADS=224)
#ifndef BIGN_VERIFY_T8
#define BIGN_VERIFY_T8 BIGN_T(8)
#endif
#ifndef BIGN_VERIFY_B8
#define BIGN_VERIFY_B8 BIGN_BLOCKS(8)
#endif
template <int N> struct bign_verify_lb { static constexpr int T = BIGN_T(N), B = BIGN_BLOCKS(N); };
template <> struct bign_verify_lb<8> { static constexpr int T = BIGN_VERIFY_T8, B = BIGN_VERIFY_B8; };
template <int N, bool STAGED> __global__ void __launch_bounds__(bign_verify_lb<N>::T, bign_verify_lb<N>::B)
bign_verify_kernel(u32* __restrict__ status, const u8* __restrict__ hashes, const u8* __restrict__ sigs,
	const u8* __restrict__ pubkeys, u64 count, const OidArg oid, const uint4* __restrict__ gtab, u8* __restrict__ wtab)
{
	constexpr int NO = 4 * N, H2 = N / 2;
	__shared__ u32 tab[BignSbox::WORDS];
	// (the direct-load build only holds the tree: 16 KB instead of 36 KB per CTA leaves more of the SM's 256 KB to L1)
	__shared__ __align__(16) u8 smem[STAGED ? BIGN_SMEM_BYTES(N) : 4 * BIGN_TREE_WORDS(N)];
	__shared__ u64 mbar;
	u32* tree = reinterpret_cast<u32*>(smem);
	BignSbox::fill(tab);
	const u64 i0 = (u64)blockIdx.x * blockDim.x;
	const u64 i = i0 + threadIdx.x;
	const u8 *p_hash = hashes + NO * i, *p_sig = sigs + (NO + NO / 2) * i, *p_pub = pubkeys + 2 * NO * i;
	// items of this CTA; a bulk copy moves multiples of 16 octets: the signature segment of a ragged last CTA
	// at l = 192 (72 octets per item, odd n) is not one — that CTA reads its items directly
	const u32 n = (u32)(count - i0 < blockDim.x ? count - i0 : blockDim.x);
	const bool staged = STAGED && ((n * (NO + NO / 2)) & 15u) == 0;
	if constexpr (!STAGED)
		__syncthreads();
	else if (staged)
	{
		u8* s_hash = smem;
		u8* s_sig = smem + BIGN_T(N) * NO;
		u8* s_pub = smem + BIGN_T(N) * (NO + NO + NO / 2);
		if (threadIdx.x == 0)
			mbar_init(&mbar, 1);
		__syncthreads();
		if (threadIdx.x == 0)
		{
			mbar_expect_tx(&mbar, n * (4 * NO + NO / 2));
			bulk_g2s(s_hash, hashes + NO * i0, n * NO, &mbar);
			bulk_g2s(s_sig, sigs + (NO + NO / 2) * i0, n * (NO + NO / 2), &mbar);
			bulk_g2s(s_pub, pubkeys + 2 * NO * i0, n * 2 * NO, &mbar);
		}
		mbar_wait(&mbar, 0);
		p_hash = s_hash + NO * threadIdx.x, p_sig = s_sig + (NO + NO / 2) * threadIdx.x, p_pub = s_pub + 2 * NO * threadIdx.x;
	}
	else
		__syncthreads();
	const BignSbox S(tab);
	// every thread stays until the block-wide inversion; `live` = still computing, `st` = verdict so far
	bool live = i < count;
	u32 st = B2G_OK;
	fe<N> qx, qy;
	u32 s0[H2], s1[N], H[N];
	pt<N> R;
	if (live)
	{
		fe_load<N>(qx, p_pub), fe_load<N>(qy, p_pub + NO);
		load_uN<N>(s1, p_sig + NO / 2);
		load_uN<N>(H, p_hash);
		{
			const u8* p = p_sig;
#pragma unroll
			for (int k = 0; k < H2; ++k)
				s0[k] = (u32)p[4 * k] | (u32)p[4 * k + 1] << 8 | (u32)p[4 * k + 2] << 16 | (u32)p[4 * k + 3] << 24;
		}
	}
	if (staged)
		__syncthreads();   // the staging buffer becomes the product tree: nobody may still be reading it
	if (live)
	{
		u32 Hq[N];
		// Q.x, Q.y < p else BAD_PUBKEY (qrFrom, :306-311); no on-curve check in the reference
		{
			fe<N> cx = qx, cy = qy;
			fe_canon<N>(cx), fe_canon<N>(cy);
			bool same = true;
#pragma unroll
			for (int k = 0; k < N; ++k) same &= cx.v[k] == qx.v[k] && cy.v[k] == qy.v[k];
			if (!same)
				st = B2G_BAD_PUBKEY, live = false;
		}
		// s1 < q else BAD_SIG (:313-318)
		if (live && geq_q<N>(s1))
			st = B2G_BAD_SIG, live = false;
		// H >= q -> H - q, once (:320-326); s1 <- (s1 + H) mod q
#pragma unroll
		for (int k = 0; k < N; ++k) Hq[k] = H[k];
		if (geq_q<N>(H))
		{
			u32 q[N];
			load_q<N>(q);
			(void)sub_n<N>(Hq, H, q);
		}
		if (live)
			modq_add<N>(s1, s1, Hq);
	}
	if (live)
	{
		// R <- (s0 + 2^l) Q + s1 G   (:329-336)
		{
			sc<N> k5;
#pragma unroll
			for (int k = 0; k < N; ++k) k5.w[k] = k < H2 ? s0[k] : (k == H2 ? 1u : 0u);
			// the window table {1..16}Q lives in this CTA's part of the launch's scratch area (ecp.cuh win_global)
#ifdef BIGN_WTAB_LOCAL   /* A/B measurement only: the round-1 form, table in local memory */
			pt_mul_var<N>(R, k5, 16 * N + 1, qx, qy);
#else
			win_global<N> W;
			W.base = wtab + i0 * win_global<N>::ITEM_BYTES + 32u * threadIdx.x, W.stride = 32u * blockDim.x;
			pt_mul_var_g<N>(R, k5, 16 * N + 1, qx, qy, W);
#endif
		}
		{
			sc<N> ks;
#pragma unroll
			for (int k = 0; k < N; ++k) ks.w[k] = s1[k];
			pt_add_mul_base<N>(R, ks, gtab);   // public scalar, same regular form
		}
		if (pt_is_inf<N>(R))
			st = B2G_BAD_SIG, live = false;
	}
	fe<N> z;
	if (live)
		z = R.Z;
	else
		fe_set_u32<N>(z, 1);
	const fe<N> zi = block_inv<N, false>(z, tree);
	if (live)
	{
		fe<N> rx;
		pt_affine_x_zi<N>(rx, R, zi);
		// s0 == belt-hash(oid || R.x || H) mod 2^l ? (:339-343)
		u32 hv[8];
		hash_oid_ab<N>(S, hv, oid, rx.v, H, (const u8*)0, 0);
		bool ok = true;
#pragma unroll
		for (int k = 0; k < H2; ++k) ok &= hv[k] == s0[k];
		st = ok ? B2G_OK : B2G_BAD_SIG;
	}
	if (i < count)
		status[i] = st;
}

#ifdef BIGN_LOWOCC_TU
}   // namespace bign_lowocc
using namespace bign_lowocc;

// ---------------------------------------------------------------- bign_lowocc.cu: launcher of the inlined build
extern "C" u32 b2g_bign_lowocc_upload_tables(const u8 H[256]) { return belt_upload_H(H); }
extern "C" u32 b2g_bign_verify8_lowocc(void* d_status, const OidArg* oid, const void* d_hashes, const void* d_sigs,
	const void* d_pubkeys, size_t count, const void* gtab, void* wtab, u32 grid, u32 threads, int staged, void* stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (staged)
		bign_verify_kernel<8, true><<<grid, threads, 0, st>>>((u32*)d_status, (const u8*)d_hashes, (const u8*)d_sigs,
			(const u8*)d_pubkeys, count, *oid, (const uint4*)gtab, (u8*)wtab);
	else
		bign_verify_kernel<8, false><<<grid, threads, 0, st>>>((u32*)d_status, (const u8*)d_hashes, (const u8*)d_sigs,
			(const u8*)d_pubkeys, count, *oid, (const uint4*)gtab, (u8*)wtab);
	b2g_note_launch();
	return b2g_check_launch("bign_verify_kernel (low-occupancy build)");
}
#else
// belt-WBL encryption of NB = 2, 3, 4 blocks: 2 NB rounds (belt_wbl.c:50-82, round reset :203).
// Round: S = r_1 ^ ... ^ r_{NB-1}; r <- (r_2, ..., r_NB ^ E(S) ^ <round>, S)
template <int NB> __device__ __forceinline__ void wbl(const BignSbox& S, u32 (&r)[4 * NB], const u32 (&key)[8])
{
#pragma unroll 1
	for (u32 round = 1; round <= 2 * NB; ++round)
	{
		u32 s[4];
#pragma unroll
		for (int w = 0; w < 4; ++w)
		{
			s[w] = r[w];
#pragma unroll
			for (int b = 1; b < NB - 1; ++b) s[w] ^= r[4 * b + w];
		}
		u32 a = s[0], b = s[1], c = s[2], d = s[3];
		belt_encr(S, a, b, c, d, key);
		a ^= round;   // <round> as a 64-bit LE word into the low half of the block
		const u32 e[4] = {a, b, c, d};
#pragma unroll
		for (int w = 0; w < 4 * (NB - 1); ++w) r[w] = r[w + 4];
#pragma unroll
		for (int w = 0; w < 4; ++w) r[4 * (NB - 2) + w] ^= e[w], r[4 * (NB - 1) + w] = s[w];
	}
}

// bignSign2Ec per item (bign_sign.c:140-245).
// STAGED (hashes, private keys and signatures 16-byte aligned, no caller-supplied nonces): the CTA's two input
// segments arrive by TMA bulk copies and its signatures leave by one bulk store, through the memory that holds
// the product tree in between — with pinned host buffers bignSign2Batch then runs zero-copy, and the private
// keys never rest in device memory.
#define BIGN_SIGN_SMEM_BYTES(N) (BIGN_T(N) * 8 * (N) > 4 * BIGN_TREE_WORDS(N) ? BIGN_T(N) * 8 * (N) : 4 * BIGN_TREE_WORDS(N))
template <int N, bool STAGED> __global__ void __launch_bounds__(BIGN_T(N), BIGN_SIGN_BLOCKS(N))
bign_sign2_kernel(u32* __restrict__ status, u8* __restrict__ sigs, const u8* __restrict__ hashes,
	const u8* __restrict__ privkeys, u64 count, const OidArg oid, const TArg targ,
	const uint4* __restrict__ gtab, const u8* __restrict__ nonces)
{
	constexpr int NO = 4 * N, H2 = N / 2, SO = NO + NO / 2;
	__shared__ u32 tab[BignSbox::WORDS];
	__shared__ __align__(16) u8 smem[BIGN_SIGN_SMEM_BYTES(N)];
	__shared__ u64 mbar;
	u32* tree = reinterpret_cast<u32*>(smem);
	BignSbox::fill(tab);
	const u64 i0 = (u64)blockIdx.x * blockDim.x;
	const u64 i = i0 + threadIdx.x;
	const u32 n = (u32)(count - i0 < blockDim.x ? count - i0 : blockDim.x);   // items of this CTA
	// bulk copies move multiples of 16 octets: l = 192 items are 48 / 72 octets, a ragged odd tail goes direct
	const bool staged = STAGED && ((n * SO) & 15u) == 0;
	const u8 *p_hash = hashes + NO * i, *p_key = privkeys + NO * i;
	if (staged)
	{
		if (threadIdx.x == 0)
			mbar_init(&mbar, 1);
		__syncthreads();
		if (threadIdx.x == 0)
		{
			mbar_expect_tx(&mbar, 2 * n * NO);
			bulk_g2s(smem, hashes + NO * i0, n * NO, &mbar);
			bulk_g2s(smem + BIGN_T(N) * NO, privkeys + NO * i0, n * NO, &mbar);
		}
		mbar_wait(&mbar, 0);
		p_hash = smem + NO * threadIdx.x, p_key = smem + BIGN_T(N) * NO + NO * threadIdx.x;
	}
	else
		__syncthreads();
	const BignSbox S(tab);
	bool live = i < count;
	u32 st = B2G_OK;
	u32 d[N], H[N], k[N];
	pt<N> R;
	if (live)
	{
		load_uN<N>(d, p_key);
		load_uN<N>(H, p_hash);
	}
	if (staged)
	{
		__syncthreads();   // everybody has its inputs: wipe the staged private keys, the buffer becomes the tree
		for (u32 w = threadIdx.x; w < BIGN_T(N) * NO / 4; w += blockDim.x)
			reinterpret_cast<u32*>(smem + BIGN_T(N) * NO)[w] = 0;
		__syncthreads();
	}
	if (live)
	{
		// 0 < d < q else BAD_PRIVKEY (:189-194)
		if (uN_is_zero<N>(d) || geq_q<N>(d))
			st = B2G_BAD_PRIVKEY, live = false;
	}
	if (live)
	{
		if (nonces)
		{
			// bignSign (bign_sign.c:27-125): the one-time key was drawn by the caller's generator
			load_uN<N>(k, nonces + NO * i);
		}
		else
		{
			// theta <- belt-hash(oid || d || t); k <- H; k <- WBL_theta(k) until 0 < k < q (:198-218)
			u32 theta[8];
			hash_oid_ab<N>(S, theta, oid, d, (const u32*)0, targ.t, targ.len);
#pragma unroll
			for (int j = 0; j < N; ++j) k[j] = H[j];
			do
				wbl<N / 4>(S, k, theta);
			while (uN_is_zero<N>(k) || geq_q<N>(k));
		}
		// R <- k G (:219-224)
		{
			sc<N> ks;
#pragma unroll
			for (int j = 0; j < N; ++j) ks.w[j] = k[j];
			pt_add_mul_base<N, true>(R, ks, gtab);   // the one-time key is secret: the form is regular
		}
		if (pt_is_inf<N>(R))
			st = B2G_BAD_PARAMS, live = false;
	}
	fe<N> z;
	if (live)
		z = R.Z;
	else
		fe_set_u32<N>(z, 1);
	const fe<N> zi = block_inv<N>(z, tree);
	if (staged)
		__syncthreads();   // every thread has read its leaf: the tree's memory becomes the output segment
	if (live)
	{
		fe<N> rx;
		pt_affine_x_zi<N>(rx, R, zi);
		// s0 <- belt-hash(oid || R.x || H) mod 2^l (:226-229)
		u32 hv[8];
		hash_oid_ab<N>(S, hv, oid, rx.v, H, (const u8*)0, 0);
		// s1 <- (k - (s0 + 2^l) d - H) mod q (:231-238)
		u32 prod[N + H2 + 1];
		{
			// prod = s0 d + d 2^l: H2 rows of products, then d added at limb H2
#pragma unroll
			for (int j = 0; j < N + H2 + 1; ++j) prod[j] = 0;
#pragma unroll
			for (int a = 0; a < H2; ++a)
			{
				u64 carry = 0;
#pragma unroll
				for (int b = 0; b < N; ++b)
				{
					const u64 t = (u64)hv[a] * d[b] + prod[a + b] + carry;
					prod[a + b] = (u32)t, carry = t >> 32;
				}
				prod[a + N] = (u32)carry;
			}
			u64 carry = 0;
#pragma unroll
			for (int b = 0; b < N; ++b)
			{
				const u64 t = (u64)d[b] + prod[H2 + b] + carry;
				prod[H2 + b] = (u32)t, carry = t >> 32;
			}
			prod[H2 + N] = (u32)carry;
		}
		u32 s1[N];
		modq_reduce<N>(s1, prod);
		modq_sub<N>(s1, k, s1);
		modq_sub<N>(s1, s1, H);   // H as is, not reduced first (zzSubMod, :237-238)
		if (staged)
		{
			// the signature goes to the CTA's output segment in shared memory (the tree is done with)
			u32* o = reinterpret_cast<u32*>(smem + SO * threadIdx.x);
#pragma unroll
			for (int j = 0; j < H2; ++j) o[j] = hv[j];
#pragma unroll
			for (int j = 0; j < N; ++j) o[H2 + j] = s1[j];
		}
		else
		{
			u8* o = sigs + SO * i;
			for (int j = 0; j < NO / 2; ++j) o[j] = (u8)(hv[j >> 2] >> (8 * (j & 3)));
			for (int j = 0; j < NO; ++j) o[NO / 2 + j] = (u8)(s1[j >> 2] >> (8 * (j & 3)));
		}
	}
	if (staged)
	{
		// one bulk store for the whole CTA when every item signed; otherwise only the successful items are
		// written, each by its own thread (a failed item must leave the caller's buffer untouched, :240-245)
		bulk_store_fence();
		const int all_ok = __syncthreads_and(i >= count || st == B2G_OK);
		if (all_ok)
		{
			if (threadIdx.x == 0)
			{
				bulk_s2g(sigs + SO * i0, smem, n * SO);
				bulk_store_wait();
			}
		}
		else if (live)
		{
			const u32* src = reinterpret_cast<const u32*>(smem + SO * threadIdx.x);
			u32* dst = reinterpret_cast<u32*>(sigs + SO * i);   // 4-byte aligned: the base is 16-byte aligned
#pragma unroll
			for (int j = 0; j < SO / 4; ++j) dst[j] = src[j];
		}
	}
	if (i < count)
		status[i] = st;
}

// bignPubkeyCalc per item (bign_misc.c:369-412): Q = d G, 0 < d < q
template <int N> __global__ void __launch_bounds__(BIGN_T(N), BIGN_BLOCKS(N))
bign_pubkey_kernel(u32* __restrict__ status, u8* __restrict__ pubkeys, const u8* __restrict__ privkeys,
	u64 count, const uint4* __restrict__ gtab)
{
	constexpr int NO = 4 * N;
	__shared__ u32 tree[BIGN_TREE_WORDS(N)];
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	bool live = i < count;
	u32 st = B2G_OK;
	pt<N> R;
	if (live)
	{
		u32 d[N];
		load_uN<N>(d, privkeys + NO * i);
		if (uN_is_zero<N>(d) || geq_q<N>(d))
			st = B2G_BAD_PRIVKEY, live = false;
		else
		{
			sc<N> ks;
#pragma unroll
			for (int j = 0; j < N; ++j) ks.w[j] = d[j];
			pt_add_mul_base<N, true>(R, ks, gtab);   // the private key is secret: the form is regular
			if (pt_is_inf<N>(R))
				st = B2G_BAD_PARAMS, live = false;
		}
	}
	fe<N> z;
	if (live)
		z = R.Z;
	else
		fe_set_u32<N>(z, 1);
	const fe<N> zi = block_inv<N>(z, tree);
	if (live)
	{
		fe<N> x, y;
		pt_affine_xy_zi<N>(x, y, R, zi);
		fe_store<N>(pubkeys + 2 * NO * i, x), fe_store<N>(pubkeys + 2 * NO * i + NO, y);
	}
	if (i < count)
		status[i] = st;
}

// ecMulA per item (ec.c:497-525): b = d * a, affine in/out; ok = 0 iff the result is O.
// ecAddMulA with the base point (ec.c:1183-1273) when kbase != 0: b = d * a + k * G.
template <int N> __global__ void __launch_bounds__(BIGN_T(N), BIGN_BLOCKS(N))
ecp_mul_kernel(u8* __restrict__ out, int* __restrict__ ok, const u8* __restrict__ pts,
	const u8* __restrict__ scalars, u32 d_len, const u8* __restrict__ kbase, u64 count,
	const uint4* __restrict__ gtab)
{
	constexpr int NO = 4 * N;
	__shared__ u32 tree[BIGN_TREE_WORDS(N)];
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	bool live = i < count;
	pt<N> R;
	if (live)
	{
		fe<N> x, y;
		fe_load<N>(x, pts + 2 * NO * i), fe_load<N>(y, pts + 2 * NO * i + NO);
		sc<N> k;
		load_scalar<N>(k, scalars + (u64)d_len * i, d_len);
		// ecMulA / ecAddMulA callers may pass secret scalars (bignDH, key transport): regular forms
		pt_mul_var<N, true>(R, k, (int)(8 * d_len), x, y);
		if (kbase)
		{
			sc<N> kg;
			load_uN<N>(kg.w, kbase + NO * i);
			pt_add_mul_base<N>(R, kg, gtab);
		}
		live = !pt_is_inf<N>(R);
	}
	fe<N> z;
	if (live)
		z = R.Z;
	else
		fe_set_u32<N>(z, 1);
	const fe<N> zi = block_inv<N>(z, tree);
	if (live)
	{
		fe<N> x, y;
		pt_affine_xy_zi<N>(x, y, R, zi);
		fe_store<N>(out + 2 * NO * i, x), fe_store<N>(out + 2 * NO * i + NO, y);
	}
	if (i < count)
		ok[i] = live ? 1 : 0;
}

// bignPubkeyVal (bign_misc.c:317-352) and bignDH (bign_misc.c:437-500) per item:
//   status BAD_PRIVKEY unless 0 < d < q (DH only), BAD_PUBKEY unless x, y < p and y^2 = x^3 - 3x + b
//   (qrFrom + ecpIsOnA), then K = d Q (BAD_PARAMS if O) and out = K.x || K.y, 2 no octets.
template <int N> __global__ void __launch_bounds__(BIGN_T(N), BIGN_BLOCKS(N))
bign_dh_kernel(u32* __restrict__ status, u8* __restrict__ out, const u8* __restrict__ privkeys,
	const u8* __restrict__ pubkeys, u64 count, u32 validate_only)
{
	constexpr int NO = 4 * N;
	__shared__ u32 tree[BIGN_TREE_WORDS(N)];
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	bool live = i < count;
	u32 st = B2G_OK;
	pt<N> R;
	if (live)
	{
		fe<N> x, y;
		sc<N> d;
		fe_load<N>(x, pubkeys + 2 * NO * i), fe_load<N>(y, pubkeys + 2 * NO * i + NO);
		if (!validate_only)
		{
			load_uN<N>(d.w, privkeys + NO * i);
			if (uN_is_zero<N>(d.w) || geq_q<N>(d.w))
				st = B2G_BAD_PRIVKEY, live = false;
		}
		if (live)
		{
			fe<N> cx = x, cy = y, l, r, t, bb;
			fe_canon<N>(cx), fe_canon<N>(cy);
			bool ok = true;
#pragma unroll
			for (int k = 0; k < N; ++k) ok &= cx.v[k] == x.v[k] && cy.v[k] == y.v[k];
			// y^2 == x^3 - 3x + b ?
#pragma unroll
			for (int k = 0; k < N; ++k) bb.v[k] = bign_c<N>::b()[k];
			fe_sqr<N>(l, y);
			fe_sqr<N>(t, x), fe_mul<N>(r, t, x);
			fe_dbl<N>(t, x), fe_add<N>(t, t, x);
			fe_sub<N>(r, r, t), fe_add<N>(r, r, bb);
			fe_sub<N>(t, l, r);
			ok &= fe_is_zero<N>(t);
			if (!ok)
				st = B2G_BAD_PUBKEY, live = false;
		}
		if (live && !validate_only)
		{
			pt_mul_var<N, true>(R, d, 32 * N, x, y);   // the private key is secret: regular form
			if (pt_is_inf<N>(R))
				st = B2G_BAD_PARAMS, live = false;
		}
	}
	if (!validate_only)
	{
		fe<N> z;
		if (live)
			z = R.Z;
		else
			fe_set_u32<N>(z, 1);
		const fe<N> zi = block_inv<N>(z, tree);
		if (live)
		{
			fe<N> x, y;
			pt_affine_xy_zi<N>(x, y, R, zi);
			fe_store<N>(out + 2 * NO * i, x), fe_store<N>(out + 2 * NO * i + NO, y);
		}
	}
	if (i < count)
		status[i] = st;
}

// sum of k affine points (the tail of the drop-in ecAddMulA: the k products d_i a_i come from
// ecp_mul_kernel, items with ok_in[i] == 0 are the point at infinity). One thread; k is small.
template <int N> __global__ void ecp_sum_kernel(u8* __restrict__ out, int* __restrict__ ok_out,
	const u8* __restrict__ pts, const int* __restrict__ ok_in, u32 k)
{
	constexpr int NO = 4 * N;
	if (blockIdx.x || threadIdx.x)
		return;
	pt<N> R;
	pt_set_inf<N>(R);
#pragma unroll 1
	for (u32 i = 0; i < k; ++i)
	{
		if (!ok_in[i])
			continue;
		fe<N> x, y;
		fe_load<N>(x, pts + 2 * NO * i), fe_load<N>(y, pts + 2 * NO * i + NO);
		pt_madd<N>(R, R, x, y);
	}
	if (pt_is_inf<N>(R))
	{
		ok_out[0] = 0;
		return;
	}
	fe<N> x, y;
	pt_to_affine<N>(x, y, R);
	fe_store<N>(out, x), fe_store<N>(out + NO, y);
	ok_out[0] = 1;
}

// ---------------------------------------------------------------- launchers (C ABI)
// q and yG are static constants, GTAB is built lazily; only the belt S-box needs uploading
extern "C" u32 b2g_bign_lowocc_upload_tables(const u8 H[256]);   // bign_lowocc.cu
extern "C" u32 b2g_bign_upload_tables(const u8 H[256])
{
	const u32 e = belt_upload_H(H);
	return e ? e : b2g_bign_lowocc_upload_tables(H);   // the second build has its own copy of the S-box
}

template <int N> static u32 bign_build_gtab(cudaStream_t st, uint4** out)
{
	uint4* p = 0;
	const size_t entries = (size_t)BIGN_GN(N) * BIGN_GE(N);
	if (cudaMalloc(&p, entries * 8 * N + 4 * N) != cudaSuccess)   /* + K behind the table */
		return b2g_check_launch("cudaMalloc(gtab)");
	bign_gtab_kernel<N><<<(u32)((entries + 127) / 128), 128, 0, st>>>(p);
	b2g_note_launch();
	bign_gtab_k_kernel<N><<<1, 1, 0, st>>>(p);
	b2g_note_launch();
	u32 e = b2g_check_launch("bign_gtab_kernel");
	if (e)
	{
		cudaFree(p);
		return e;
	}
	// the table must be complete before any other stream reads it
	if (cudaStreamSynchronize(st) != cudaSuccess)
	{
		e = b2g_check_launch("sync(gtab)");
		cudaFree(p);
		return e ? e : B2G_ERR_CUDA;
	}
	*out = p;
	return B2G_OK;
}
// the device entry points may be called from several host threads: build each table once
// (published with release / read with acquire; lives until the process exits)
template <int N> static u32 bign_ensure_gtab(cudaStream_t st, const uint4** out)
{
	static std::mutex mu;
	std::atomic<uint4*>& slot = g_gtab[b2g_cur_dev() & (BIGN_MAX_DEV - 1)][N / 4 - 2];
	uint4* p = slot.load(std::memory_order_acquire);
	if (!p)
	{
		std::lock_guard<std::mutex> lock(mu);
		p = slot.load(std::memory_order_relaxed);
		if (!p)
		{
			const u32 e = bign_build_gtab<N>(st, &p);
			if (e) return e;
			slot.store(p, std::memory_order_release);
		}
	}
	*out = p;
	return B2G_OK;
}

static u32 make_oid(OidArg& o, const u8* der, size_t len)
{
	if (len > BIGN_MAX_OID)
		return 119u;   // ERR_NOT_IMPLEMENTED: longer OIDs are not staged into kernel arguments
	for (size_t i = 0; i < BIGN_MAX_OID; ++i) o.der[i] = i < len ? der[i] : 0;
	o.len = (u32)len;
	return B2G_OK;
}

// CTA size for a batch of `count` items. Large batches: BIGN_T(N) threads (one inversion per CTA amortised
// over the most items). Small batches (a shard of a strong-scaled job): the time is the load of the busiest
// SM, ceil(CTAs / SMs) x threads, so try smaller CTAs too and take the smallest load, with a measured
// penalty for the shorter amortisation (128-thread CTAs are ~6 % slower per item than 256-thread ones).
// 2^17 items on 148 SMs: 512 CTAs of 256 put 4 CTAs = 1024 threads on some SMs (1.70x of one GPU's rate
// at N = 2, measured), 586 CTAs of 224 put at most 896 (ideal 886).
template <int N> static inline u32 bign_threads(size_t count)
{
	const u64 sms = (u64)b2g_sm_count();
	{
		// measurement switch: a fixed CTA size (a multiple of 32 in [BIGN_T / 2, BIGN_T])
		static const char* fix = getenv("B2G_BIGN_THREADS");
		const u32 t = fix ? (u32)atoi(fix) : 0;
		if (t >= BIGN_T(N) / 2 && t <= BIGN_T(N) && t % 32 == 0)
			return t;
	}
	u32 best_t = BIGN_T(N);
	double best = 0;
	for (u32 t = BIGN_T(N); t >= BIGN_T(N) / 2; t -= 32)
	{
		const u64 ctas = (count + t - 1) / t;
		const u64 per_sm = (ctas + sms - 1) / sms;
		const double cost = (double)(per_sm * t) * (1.0 + 0.06 * (double)(BIGN_T(N) - t) / (BIGN_T(N) / 2));
		if (best == 0 || cost < best * 0.995)
			best = cost, best_t = t;
	}
	return best_t;
}
template <int N> static inline u32 bign_grid(size_t count, u32 threads) { return (u32)((count + threads - 1) / threads); }

extern "C" u32 b2g_bign_verify8_lowocc(void* d_status, const OidArg* oid, const void* d_hashes, const void* d_sigs,
	const void* d_pubkeys, size_t count, const void* gtab, void* wtab, u32 grid, u32 threads, int staged, void* stream);
extern "C" u32 b2g_bign_lowocc_upload_tables(const u8 H[256]);

// Scratch for the per-thread window tables of a verification launch (ecp.cuh win_global): stream-ordered
// allocations from the device's default pool, which is told to keep what it has (no trip to the driver per
// launch after the first). Concurrent launches on different streams each get their own area.
#ifndef BIGN_WTAB_CHUNK
#define BIGN_WTAB_CHUNK ((size_t)1 << 20)   /* items per launch at most: 1.5 / 2.5 / 3 GiB of scratch */
#endif
static u32 wtab_alloc(void** p, size_t bytes, cudaStream_t st)
{
	static std::atomic<bool> tuned[BIGN_MAX_DEV];
	const int dev = b2g_cur_dev();
	if (!tuned[dev & (BIGN_MAX_DEV - 1)].exchange(true))
	{
		cudaMemPool_t pool;
		unsigned long long keep = ~0ull;
		if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
			(void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		(void)cudaGetLastError();
	}
	if (cudaMallocAsync(p, bytes, st) != cudaSuccess)
		return b2g_check_launch("cudaMallocAsync(bign window tables)");
	return B2G_OK;
}

template <int N> static u32 verify_launch(void* d_status, const OidArg& oid, const void* d_hashes, const void* d_sigs,
	const void* d_pubkeys, size_t count, cudaStream_t st)
{
	constexpr size_t NO = 4 * N;
	const uint4* gtab;
	u32 e = bign_ensure_gtab<N>(st, &gtab);
	if (e) return e;
	// TMA staging needs 16-byte aligned segments (every CTA's first item then is: item sizes are multiples
	// of 16 and so is every CTA size); B2G_NO_STAGING=1 keeps the per-thread loads (A/B measurement)
	// Staging pays where the inputs are (pinned) HOST memory — the zero-copy path of bignVerifyBatch: 46.0 -> 51.3 M/s.
	// For inputs resident in HBM it gains nothing at l = 128 (56.7 against 56.9 M/s) and costs at the wider fields
	// (l = 192: 15.2 against 15.7, l = 256: 6.3 against 7.2 M/s): its 36 KB buffer per CTA, against 16 KB for the
	// tree alone, is taken from the L1 that serves the call frames. B2G_FORCE_STAGING=1 stages device inputs too.
	static const bool no_staging = getenv("B2G_NO_STAGING") != 0, force_staging = getenv("B2G_FORCE_STAGING") != 0;
	bool staged = !no_staging && (((uintptr_t)d_hashes | (uintptr_t)d_sigs | (uintptr_t)d_pubkeys) & 15) == 0;
	if (staged && !force_staging)
	{
		cudaPointerAttributes pa;
		staged = cudaPointerGetAttributes(&pa, d_hashes) == cudaSuccess && pa.type == cudaMemoryTypeHost;
		(void)cudaGetLastError();
	}
	const size_t chunk = count < BIGN_WTAB_CHUNK ? count : BIGN_WTAB_CHUNK;
	// every launch covers at most chunk items rounded up to whole CTAs of at most BIGN_T(N) threads
	void* wtab = 0;
	if ((e = wtab_alloc(&wtab, (chunk + BIGN_T(N)) * win_global<N>::ITEM_BYTES, st)))
		return e;
	for (size_t off = 0; off < count && !e; off += chunk)
	{
		const size_t n = count - off < chunk ? count - off : chunk;
		const u32 threads = bign_threads<N>(n), grid = bign_grid<N>(n, threads);
		u32* p_st = (u32*)d_status + off;
		const u8 *p_h = (const u8*)d_hashes + NO * off, *p_s = (const u8*)d_sigs + (NO + NO / 2) * off,
			*p_q = (const u8*)d_pubkeys + 2 * NO * off;
		// A small grid (a shard of a strong-scaled batch: 2^15 items = one CTA per SM, under two warps per
		// scheduler) is latency-bound, not issue-bound: the build with every field product inlined gives the
		// scheduler independent chains to interleave. Measured, 2^15 / 2^16 / 2^17 / 2^18 items: out-of-line
		// products 0.978 / 1.537 / 2.650 / 4.835 ms, inlined 0.813 / 1.605 / 3.216 / 5.729 ms.
		if (N == 8 && n <= (size_t)b2g_sm_count() * 320 && !getenv("B2G_NO_LOWOCC"))
		{
			e = b2g_bign_verify8_lowocc(p_st, &oid, p_h, p_s, p_q, n, gtab, wtab, grid, threads, staged, st);
			continue;
		}
		if (staged)
			bign_verify_kernel<N, true><<<grid, threads, 0, st>>>(p_st, p_h, p_s, p_q, n, oid, gtab, (u8*)wtab);
		else
			bign_verify_kernel<N, false><<<grid, threads, 0, st>>>(p_st, p_h, p_s, p_q, n, oid, gtab, (u8*)wtab);
		b2g_note_launch();
		e = b2g_check_launch("bign_verify_kernel");
	}
	if (cudaFreeAsync(wtab, st) != cudaSuccess && !e)
		e = b2g_check_launch("cudaFreeAsync(bign window tables)");
	return e;
}

// l = 128 / 192 / 256 -> dispatch on the limb count
#define BIGN_DISPATCH(l, CALL) \
	((l) == 128 ? CALL(8) : (l) == 192 ? CALL(12) : (l) == 256 ? CALL(16) : 119u)

extern "C" u32 b2g_bignVerifyBatchL_dev(size_t l, void* d_status, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_sigs, const void* d_pubkeys, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	OidArg oid;
	if ((e = make_oid(oid, oid_der, oid_len))) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) verify_launch<N>(d_status, oid, d_hashes, d_sigs, d_pubkeys, count, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}
extern "C" u32 b2g_bignVerifyBatch_dev(void* d_status, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_sigs, const void* d_pubkeys, size_t count, void* stream)
{
	return b2g_bignVerifyBatchL_dev(128, d_status, oid_der, oid_len, d_hashes, d_sigs, d_pubkeys, count, stream);
}

template <int N> static u32 sign2_launch(void* d_status, void* d_sigs, const OidArg& oid, const TArg& ta,
	const void* d_hashes, const void* d_privkeys, size_t count, cudaStream_t st, const void* d_nonces = 0)
{
	const uint4* gtab;
	u32 e = bign_ensure_gtab<N>(st, &gtab);
	if (e) return e;
	static const bool no_staging = getenv("B2G_NO_STAGING") != 0;
	const bool staged = !no_staging && !d_nonces &&
		(((uintptr_t)d_hashes | (uintptr_t)d_privkeys | (uintptr_t)d_sigs) & 15) == 0;
	const u32 threads = bign_threads<N>(count), grid = bign_grid<N>(count, threads);
	if (staged)
		bign_sign2_kernel<N, true><<<grid, threads, 0, st>>>((u32*)d_status, (u8*)d_sigs,
			(const u8*)d_hashes, (const u8*)d_privkeys, count, oid, ta, gtab, (const u8*)d_nonces);
	else
		bign_sign2_kernel<N, false><<<grid, threads, 0, st>>>((u32*)d_status, (u8*)d_sigs,
			(const u8*)d_hashes, (const u8*)d_privkeys, count, oid, ta, gtab, (const u8*)d_nonces);
	b2g_note_launch();
	return b2g_check_launch("bign_sign2_kernel");
}

extern "C" u32 b2g_bignSign2BatchL_t_dev(size_t l, void* d_status, void* d_sigs, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, const void* t, size_t t_len, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	OidArg oid;
	TArg ta;
	if ((e = make_oid(oid, oid_der, oid_len))) return e;
	if (t_len > BIGN_MAX_T) return 119u;
	if (l != 128 && l != 192 && l != 256) return 119u;
	for (size_t i = 0; i < BIGN_MAX_T; ++i) ta.t[i] = (t && i < t_len) ? ((const u8*)t)[i] : 0;
	ta.len = t ? (u32)t_len : 0;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) sign2_launch<N>(d_status, d_sigs, oid, ta, d_hashes, d_privkeys, count, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}
extern "C" u32 b2g_bignSign2Batch_t_dev(void* d_status, void* d_sigs, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, const void* t, size_t t_len, void* stream)
{
	return b2g_bignSign2BatchL_t_dev(128, d_status, d_sigs, oid_der, oid_len, d_hashes, d_privkeys, count, t, t_len, stream);
}
extern "C" u32 b2g_bignSign2Batch_dev(void* d_status, void* d_sigs, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, void* stream)
{
	return b2g_bignSign2BatchL_t_dev(128, d_status, d_sigs, oid_der, oid_len, d_hashes, d_privkeys, count, 0, 0, stream);
}

template <int N> static u32 pubkey_launch(void* d_status, void* d_pubkeys, const void* d_privkeys, size_t count, cudaStream_t st)
{
	const uint4* gtab;
	u32 e = bign_ensure_gtab<N>(st, &gtab);
	if (e) return e;
	bign_pubkey_kernel<N><<<bign_grid<N>(count, bign_threads<N>(count)), bign_threads<N>(count), 0, st>>>((u32*)d_status, (u8*)d_pubkeys,
		(const u8*)d_privkeys, count, gtab);
	b2g_note_launch();
	return b2g_check_launch("bign_pubkey_kernel");
}

extern "C" u32 b2g_bignPubkeyCalcBatchL_dev(size_t l, void* d_status, void* d_pubkeys, const void* d_privkeys,
	size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) pubkey_launch<N>(d_status, d_pubkeys, d_privkeys, count, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}
extern "C" u32 b2g_bignPubkeyCalcBatch_dev(void* d_status, void* d_pubkeys, const void* d_privkeys,
	size_t count, void* stream)
{
	return b2g_bignPubkeyCalcBatchL_dev(128, d_status, d_pubkeys, d_privkeys, count, stream);
}

template <int N> static u32 mul_launch(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	size_t count, cudaStream_t st)
{
	ecp_mul_kernel<N><<<bign_grid<N>(count, bign_threads<N>(count)), bign_threads<N>(count), 0, st>>>((u8*)d_b, (int*)d_ok,
		(const u8*)d_a, (const u8*)d_d, (u32)d_len, (const u8*)0, count, (const uint4*)0);
	b2g_note_launch();
	return b2g_check_launch("ecp_mul_kernel");
}

extern "C" u32 b2g_ecMulABatchL_dev(size_t l, void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (d_len == 0 || d_len > l / 4) return B2G_BAD_INPUT;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_ok & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) mul_launch<N>(d_b, d_ok, d_a, d_d, d_len, count, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}
extern "C" u32 b2g_ecMulABatch_dev(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	size_t count, void* stream)
{
	return b2g_ecMulABatchL_dev(128, d_b, d_ok, d_a, d_d, d_len, count, stream);
}

template <int N> static u32 addmul_launch(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	const void* d_k, size_t count, cudaStream_t st)
{
	const uint4* gtab;
	u32 e = bign_ensure_gtab<N>(st, &gtab);
	if (e) return e;
	if (!d_k) return B2G_BAD_INPUT;
	ecp_mul_kernel<N><<<bign_grid<N>(count, bign_threads<N>(count)), bign_threads<N>(count), 0, st>>>((u8*)d_b, (int*)d_ok, (const u8*)d_a,
		(const u8*)d_d, (u32)d_len, (const u8*)d_k, count, gtab);
	b2g_note_launch();
	return b2g_check_launch("ecp_mul_kernel(+G)");
}

extern "C" u32 b2g_ecAddMulABatchL_dev(size_t l, void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	const void* d_k, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (d_len == 0 || d_len > l / 4) return B2G_BAD_INPUT;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_ok & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) addmul_launch<N>(d_b, d_ok, d_a, d_d, d_len, d_k, count, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}
extern "C" u32 b2g_ecAddMulABatch_dev(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	const void* d_k, size_t count, void* stream)
{
	return b2g_ecAddMulABatchL_dev(128, d_b, d_ok, d_a, d_d, d_len, d_k, count, stream);
}

template <int N> static u32 dh_launch(void* d_status, void* d_out, const void* d_privkeys, const void* d_pubkeys,
	size_t count, u32 validate_only, cudaStream_t st)
{
	bign_dh_kernel<N><<<bign_grid<N>(count, bign_threads<N>(count)), bign_threads<N>(count), 0, st>>>((u32*)d_status, (u8*)d_out,
		(const u8*)d_privkeys, (const u8*)d_pubkeys, count, validate_only);
	b2g_note_launch();
	return b2g_check_launch("bign_dh_kernel");
}

// bignDH on a batch: d_out receives K.x || K.y (l/2 octets per item) for items with status 0
extern "C" u32 b2g_bignDHBatchL_dev(size_t l, void* d_status, void* d_out, const void* d_privkeys,
	const void* d_pubkeys, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) dh_launch<N>(d_status, d_out, d_privkeys, d_pubkeys, count, 0u, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}

// bignPubkeyVal on a batch
extern "C" u32 b2g_bignPubkeyValBatchL_dev(size_t l, void* d_status, const void* d_pubkeys, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) dh_launch<N>(d_status, (void*)0, (const void*)0, d_pubkeys, count, 1u, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}

// bignSign on a batch: like sign2 but the one-time keys k_i (l/4 octets each, 0 < k_i < q, drawn by the
// caller's generator) come from d_nonces
extern "C" u32 b2g_bignSignBatchL_k_dev(size_t l, void* d_status, void* d_sigs, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, const void* d_nonces, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	OidArg oid;
	TArg ta;
	if ((e = make_oid(oid, oid_der, oid_len))) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (!d_nonces) return B2G_BAD_INPUT;
	for (size_t i = 0; i < BIGN_MAX_T; ++i) ta.t[i] = 0;
	ta.len = 0;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) sign2_launch<N>(d_status, d_sigs, oid, ta, d_hashes, d_privkeys, count, st, d_nonces)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}

template <int N> static u32 sum_launch(void* d_out, void* d_ok_out, const void* d_pts, const void* d_ok_in, size_t k,
	cudaStream_t st)
{
	ecp_sum_kernel<N><<<1, 32, 0, st>>>((u8*)d_out, (int*)d_ok_out, (const u8*)d_pts, (const int*)d_ok_in, (u32)k);
	b2g_note_launch();
	return b2g_check_launch("ecp_sum_kernel");
}

// d_out (l/2 octets) <- sum of the k affine points d_pts[i] with d_ok_in[i] != 0; d_ok_out[0] = 0 iff the sum is O
extern "C" u32 b2g_ecSumL_dev(size_t l, void* d_out, void* d_ok_out, const void* d_pts, const void* d_ok_in,
	size_t k, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (l != 128 && l != 192 && l != 256) return 119u;
	if (((uintptr_t)d_ok_out & 3) || ((uintptr_t)d_ok_in & 3) || k > 0xFFFFFFFFu) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) sum_launch<N>(d_out, d_ok_out, d_pts, d_ok_in, k, st)
	return BIGN_DISPATCH(l, CALL);
#undef CALL
}
#endif   // !BIGN_LOWOCC_TU
