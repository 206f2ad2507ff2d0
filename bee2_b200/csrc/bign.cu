// bign.cu — bign (STB 34.101.45) on bign-curve256v1: batch verify / sign2 / pubkey-calc /
// scalar multiplication kernels for sm_100a + C-ABI launchers.
//
// Replaces bignVerifyEc (bign_sign.c:268-347) incl. ecAddMulA (ec.c:1183-1273), ecpToAJ and
// belt-hash; bignSign2Ec (bign_sign.c:140-245) incl. bignMulBase/ecMulPreOD (bign_misc.c:115-137,
// ec.c:892-964) and beltWBL (belt_wbl.c:50-82); bignPubkeyCalc (bign_misc.c:369-412);
// ecMulA (ec.c:497-525).
//
// Work decomposition: ONE THREAD PER ITEM (signature / key / scalar-point pair), field
// elements as 8 x u32 in registers. Scalar multiplication is REGULAR so that all lanes of a
// warp execute the same doublings and additions in lock-step:
//   * fixed base G: BIGN_GW-bit windows (13) over a device-resident table GTAB[20][8192] of
//     affine multiples j * 2^(13 i) * G (10 MiB, L2-resident; generated once per process by
//     bign_gtab_kernel with the same point code) -> 20 mixed additions, no doublings;
//   * variable base Q: 4-bit windows, per-thread table {1..15}Q in local memory
//     -> 4 doublings + 1 addition per nibble.
// The reference's interleaved wNAF (ec.c:1206-1268) is irregular and would diverge.
#include <mutex>
#include "ecp256.cuh"
#include "belt_dev.cuh"

#define BIGN_THREADS 128
#ifndef BIGN_MIN_BLOCKS
#define BIGN_MIN_BLOCKS 4
#endif
#ifndef BIGN_GW
#define BIGN_GW 13                                  // fixed-base window width in bits (<= 16)
#endif
#define BIGN_GN ((256 + BIGN_GW - 1) / BIGN_GW)     // number of windows
#define BIGN_GE (1 << BIGN_GW)                      // entries per window (entry 0 unused)
#define BIGN_MAX_OID 64
#define BIGN_MAX_T 64

// q, little-endian limbs (bign_params.c:61-66)
__constant__ u32 c_q[8] = {0x263D6607u, 0x7E5ABF99u, 0x0DFB4DFCu, 0xD95C8ED6u,
	0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
// y-coordinate of G = (0, yG), little-endian limbs (bign_params.c:68-73)
__constant__ u32 c_yG[8] = {0x04516A93u, 0x1E29CF18u, 0xC408F652u, 0x78913966u,
	0x51D6835Du, 0x5CE4C9A3u, 0xFB16D69Fu, 0x6BF7FC3Cu};

static uint4* g_gtab;          // device: BIGN_GN * BIGN_GE entries of 64 octets (x || y); entry j = 0 unused

struct OidArg { u8 der[BIGN_MAX_OID]; u32 len; };
// a scalar handed BY VALUE to the out-of-line multiplication routines: the kernels keep no
// address-taken locals besides the accumulator (nvcc 12.9 was seen to overlap the stack slots of
// two live address-taken locals of bign_verify_kernel — see DESIGN.md §4.3)
struct sc256 { u32 w[8]; };
struct TArg { u8 t[BIGN_MAX_T]; u32 len; };

// ---------------------------------------------------------------- small helpers
// r = (a + b) mod q for a, b < q (zz_mod.c:42)
__device__ __forceinline__ void modq_add(u32* r, const u32* a, const u32* b)
{
	u32 t[8], u[8];
	const u32 c = add8(t, a, b);
	const u32 m = sub8(u, t, c_q);
	const bool take = c || m == 0;   // carried out, or t >= q
#pragma unroll
	for (int i = 0; i < 8; ++i) r[i] = take ? u[i] : t[i];
}
// r = (a - b) mod 2^256, + q if it borrowed — zzSubMod without range assumptions (zz_mod.c:120)
__device__ __forceinline__ void modq_sub(u32* r, const u32* a, const u32* b)
{
	u32 t[8], u[8];
	const u32 m = sub8(t, a, b);
	(void)add8(u, t, c_q);
#pragma unroll
	for (int i = 0; i < 8; ++i) r[i] = m ? u[i] : t[i];
}
__device__ __forceinline__ bool u256_is_zero(const u32* a)
{
	return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0;
}

// x (n limbs, n <= 16) mod q, q = 2^256 - c with c < 2^128: fold hi*c into lo until hi = 0.
// Plain 64-bit loops; runs once per signature.
__device__ __noinline__ void modq_reduce(u32* r, const u32* x, int n)
{
	u32 cur[16], c[4];
#pragma unroll
	for (int i = 0; i < 16; ++i) cur[i] = i < n ? x[i] : 0;
	// c = 2^256 - q: the low four limbs of -q
	{
		u64 b = 0;
		for (int i = 0; i < 4; ++i)
		{
			const u64 d = (u64)0 - c_q[i] - b;
			c[i] = (u32)d, b = (d >> 32) & 1;
		}
	}
	for (int round = 0; round < 8; ++round)
	{
		u32 hi[8], nxt[16];
		bool any = false;
		for (int i = 0; i < 8; ++i) hi[i] = cur[8 + i], any |= hi[i] != 0;
		if (!any) break;
		for (int i = 0; i < 16; ++i) nxt[i] = i < 8 ? cur[i] : 0;
		for (int i = 0; i < 8; ++i)
		{
			u64 carry = 0;
			for (int j = 0; j < 4; ++j)
			{
				const u64 t = (u64)hi[i] * c[j] + nxt[i + j] + carry;
				nxt[i + j] = (u32)t, carry = t >> 32;
			}
			for (int k = i + 4; carry && k < 16; ++k)
			{
				const u64 t = (u64)nxt[k] + carry;
				nxt[k] = (u32)t, carry = t >> 32;
			}
		}
		for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
	}
	// now cur < 2^256 (+ tiny); bring into [0, q)
	for (int guard = 0; guard < 4 && u256_geq(cur, c_q); ++guard)
	{
		u32 t[8];
		(void)sub8(t, cur, c_q);
		for (int i = 0; i < 8; ++i) cur[i] = t[i];
	}
	for (int i = 0; i < 8; ++i) r[i] = cur[i];
}

__device__ __forceinline__ void load_u256(u32* r, const u8* p)
{
	fe t;
	fe_load(t, p);
#pragma unroll
	for (int i = 0; i < 8; ++i) r[i] = t.v[i];
}

// d_len <= 32 little-endian octets -> limbs, without dynamic indexing of the destination
__device__ __forceinline__ void load_scalar(sc256& k, const u8* s, u32 d_len)
{
#pragma unroll
	for (int l = 0; l < 8; ++l)
	{
		u32 w = 0;
#pragma unroll
		for (int b = 0; b < 4; ++b)
			if ((u32)(4 * l + b) < d_len)
				w |= (u32)s[4 * l + b] << (8 * b);
		k.w[l] = w;
	}
}

// ---------------------------------------------------------------- scalar multiplication
// acc += k * G for a 256-bit k (little-endian limbs) through the fixed-base window table
__device__ __noinline__ void pt_add_mul_base(pt& acc, const sc256 ks, const uint4* __restrict__ gtab)
{
	const u32* k = ks.w;
#pragma unroll 1
	for (int i = 0; i < BIGN_GN; ++i)
	{
		const int bit = BIGN_GW * i, limb = bit >> 5;
		const u64 w = (u64)k[limb] | (limb < 7 ? (u64)k[limb + 1] << 32 : 0);
		const u32 d = (u32)(w >> (bit & 31)) & (BIGN_GE - 1);
		if (d)
		{
			const uint4* e = gtab + ((size_t)(i * BIGN_GE + (int)d) << 2);
			const uint4 a0 = __ldg(e), a1 = __ldg(e + 1), a2 = __ldg(e + 2), a3 = __ldg(e + 3);
			fe x, y;
			x.v[0] = a0.x, x.v[1] = a0.y, x.v[2] = a0.z, x.v[3] = a0.w;
			x.v[4] = a1.x, x.v[5] = a1.y, x.v[6] = a1.z, x.v[7] = a1.w;
			y.v[0] = a2.x, y.v[1] = a2.y, y.v[2] = a2.z, y.v[3] = a2.w;
			y.v[4] = a3.x, y.v[5] = a3.y, y.v[6] = a3.z, y.v[7] = a3.w;
			pt_madd(acc, acc, x, y);
		}
	}
}

// acc = k * (x, y) for a scalar of nbits bits (little-endian limbs; nbits multiple of 4,
// bits above nbits ignored), 4-bit fixed windows, most significant first
__device__ __noinline__ void pt_mul_var(pt& acc, const sc256 ks, int nbits, const fe x, const fe y)
{
	const u32* k = ks.w;
	pt T[16];   // T[j] = j * (x, y); T[0] unused. Dynamic indexing -> local memory.
	pt_set_affine(T[1], x, y);
#pragma unroll 1
	for (int j = 2; j < 16; ++j)
	{
		if (j & 1)
			pt_madd(T[j], T[j - 1], x, y);
		else
			pt_dbl(T[j], T[j >> 1]);
	}
	// the top window only selects (no doublings of O)
	{
		const int i = nbits / 4 - 1;
		const u32 d = (k[i >> 3] >> (4 * (i & 7))) & 15u;
		if (d)
			acc = T[d];
		else
			pt_set_inf(acc);
	}
#pragma unroll 1
	for (int i = nbits / 4 - 2; i >= 0; --i)
	{
#pragma unroll 1
		for (int s = 0; s < 4; ++s)
			pt_dbl(acc, acc);
		const u32 d = (k[i >> 3] >> (4 * (i & 7))) & 15u;
		if (d)
			pt_add(acc, acc, T[d]);
	}
}

// ---------------------------------------------------------------- belt-hash(oid || a || b)
__device__ __forceinline__ void hash_oid_2x32(const BeltSmallT& S, u32 (&out)[8], const OidArg& oid,
	const u32* a, const u32* b, const u8* extra, u32 extra_len)
{
	// message = oid || a (32) || [b (32)] || extra, zero-padded to whole 32-octet blocks
	u32 msg[(BIGN_MAX_OID + 64 + BIGN_MAX_T + 31) / 32 * 8];
	u8* m8 = reinterpret_cast<u8*>(msg);
	const int nw = (int)(sizeof(msg) / 4);
	for (int i = 0; i < nw; ++i) msg[i] = 0;
	u32 pos = 0;
	for (u32 i = 0; i < oid.len; ++i) m8[pos++] = oid.der[i];
	for (u32 i = 0; i < 32; ++i) m8[pos++] = (u8)(a[i >> 2] >> (8 * (i & 3)));
	if (b)
		for (u32 i = 0; i < 32; ++i) m8[pos++] = (u8)(b[i >> 2] >> (8 * (i & 3)));
	for (u32 i = 0; i < extra_len; ++i) m8[pos++] = extra[i];
	const belt_digest dg = belt_hash_words(S, msg, pos);
#pragma unroll
	for (int i = 0; i < 8; ++i) out[i] = dg.w[i];
}

// ---------------------------------------------------------------- kernels
// Table of fixed-base multiples: entry (i, j) = j * 2^(BIGN_GW i) * G, affine.
__global__ void __launch_bounds__(BIGN_THREADS) bign_gtab_kernel(uint4* gtab)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= BIGN_GN * BIGN_GE)
		return;
	const int i = idx / BIGN_GE, j = idx % BIGN_GE;
	uint4* e = gtab + ((size_t)idx << 2);
	if (j == 0)
	{
		e[0] = e[1] = e[2] = e[3] = make_uint4(0, 0, 0, 0);
		return;
	}
	// k = j << (BIGN_GW * i); bits past 2^256 cannot occur for the digits a 256-bit scalar has,
	// such entries are never read
	sc256 k = {{0, 0, 0, 0, 0, 0, 0, 0}};
	{
		const int bit = BIGN_GW * i, limb = bit >> 5;
		const u64 w = (u64)j << (bit & 31);
#pragma unroll
		for (int l = 0; l < 8; ++l)
			k.w[l] = l == limb ? (u32)w : (l == limb + 1 ? (u32)(w >> 32) : 0u);
	}
	fe gx, gy, x, y;
	fe_set_u32(gx, 0);
#pragma unroll
	for (int l = 0; l < 8; ++l) gy.v[l] = c_yG[l];
	pt acc;
	pt_mul_var(acc, k, 256, gx, gy);
	pt_to_affine(x, y, acc);
	e[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
	e[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
	e[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
	e[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
}

// bignVerifyEc per item (bign_sign.c:268-347, l = 128)
__global__ void __launch_bounds__(BIGN_THREADS, BIGN_MIN_BLOCKS)
bign_verify_kernel(u32* __restrict__ status, const u8* __restrict__ hashes, const u8* __restrict__ sigs,
	const u8* __restrict__ pubkeys, u64 count, const OidArg oid, const uint4* __restrict__ gtab)
{
	__shared__ u32 tab[256];
	BeltSmallT::fill(tab);
	__syncthreads();
	const BeltSmallT S(tab);
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	fe qx, qy;
	u32 s0[5], s1[8], H[8], Hq[8];
	fe_load(qx, pubkeys + 64 * i), fe_load(qy, pubkeys + 64 * i + 32);
	load_u256(s1, sigs + 48 * i + 16);
	load_u256(H, hashes + 32 * i);
	{
		const u8* p = sigs + 48 * i;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			s0[k] = (u32)p[4 * k] | (u32)p[4 * k + 1] << 8 | (u32)p[4 * k + 2] << 16 | (u32)p[4 * k + 3] << 24;
		s0[4] = 1;   // s0 + 2^l (bign_sign.c:329-330)
	}
	// Q.x, Q.y < p else BAD_PUBKEY (qrFrom, :306-311); no on-curve check in the reference
	{
		fe cx = qx, cy = qy;
		fe_canon(cx), fe_canon(cy);
		bool same = true;
#pragma unroll
		for (int k = 0; k < 8; ++k) same &= cx.v[k] == qx.v[k] && cy.v[k] == qy.v[k];
		if (!same)
		{
			status[i] = B2G_BAD_PUBKEY;
			return;
		}
	}
	// s1 < q else BAD_SIG (:313-318)
	if (u256_geq(s1, c_q))
	{
		status[i] = B2G_BAD_SIG;
		return;
	}
	// H >= q -> H - q, once (:320-326); s1 <- (s1 + H) mod q
#pragma unroll
	for (int k = 0; k < 8; ++k) Hq[k] = H[k];
	if (u256_geq(H, c_q))
		(void)sub8(Hq, H, c_q);
	modq_add(s1, s1, Hq);
	// R <- (s0 + 2^128) Q + s1 G   (:332-336)
	pt R;
	{
		// 129-bit scalar: the top bit is always 1 -> start from Q and consume 32 nibbles
		sc256 k5 = {{s0[0], s0[1], s0[2], s0[3], 1u, 0u, 0u, 0u}};
		pt_mul_var(R, k5, 132, qx, qy);
	}
	{
		sc256 ks;
#pragma unroll
		for (int k = 0; k < 8; ++k) ks.w[k] = s1[k];
		pt_add_mul_base(R, ks, gtab);
	}
	if (pt_is_inf(R))
	{
		status[i] = B2G_BAD_SIG;
		return;
	}
	fe rx;
	pt_to_affine_x(rx, R);
	// s0 == belt-hash(oid || R.x || H) mod 2^l ? (:339-343)
	u32 hv[8];
	hash_oid_2x32(S, hv, oid, rx.v, H, (const u8*)0, 0);
	const bool ok = hv[0] == s0[0] && hv[1] == s0[1] && hv[2] == s0[2] && hv[3] == s0[3];
	status[i] = ok ? B2G_OK : B2G_BAD_SIG;
}

// belt-WBL encryption of exactly 32 octets: 2n = 4 rounds (belt_wbl.c:50-82)
__device__ __forceinline__ void wbl32(const BeltSmallT& S, u32 (&r)[8], const u32 (&key)[8])
{
#pragma unroll 1
	for (u32 round = 1; round <= 4; ++round)
	{
		u32 a = r[0], b = r[1], c = r[2], d = r[3];
		belt_encr(S, a, b, c, d, key);
		a ^= round;   // <round> as a 64-bit LE word into the low half of the block
		const u32 n0 = r[4] ^ a, n1 = r[5] ^ b, n2 = r[6] ^ c, n3 = r[7] ^ d;
		r[4] = r[0], r[5] = r[1], r[6] = r[2], r[7] = r[3];
		r[0] = n0, r[1] = n1, r[2] = n2, r[3] = n3;
	}
}

// bignSign2Ec per item (bign_sign.c:140-245, l = 128)
__global__ void __launch_bounds__(BIGN_THREADS, BIGN_MIN_BLOCKS)
bign_sign2_kernel(u32* __restrict__ status, u8* __restrict__ sigs, const u8* __restrict__ hashes,
	const u8* __restrict__ privkeys, u64 count, const OidArg oid, const TArg targ,
	const uint4* __restrict__ gtab)
{
	__shared__ u32 tab[256];
	BeltSmallT::fill(tab);
	__syncthreads();
	const BeltSmallT S(tab);
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	u32 d[8], H[8], k[8], theta[8];
	load_u256(d, privkeys + 32 * i);
	load_u256(H, hashes + 32 * i);
	// 0 < d < q else BAD_PRIVKEY (:189-194)
	if (u256_is_zero(d) || u256_geq(d, c_q))
	{
		status[i] = B2G_BAD_PRIVKEY;
		return;
	}
	// theta <- belt-hash(oid || d || t); k <- H; k <- WBL_theta(k) until 0 < k < q (:198-218)
	hash_oid_2x32(S, theta, oid, d, (const u32*)0, targ.t, targ.len);
#pragma unroll
	for (int j = 0; j < 8; ++j) k[j] = H[j];
	do
		wbl32(S, k, theta);
	while (u256_is_zero(k) || u256_geq(k, c_q));
	// R <- k G (:219-224)
	pt R;
	pt_set_inf(R);
	{
		sc256 ks;
#pragma unroll
		for (int j = 0; j < 8; ++j) ks.w[j] = k[j];
		pt_add_mul_base(R, ks, gtab);
	}
	if (pt_is_inf(R))
	{
		status[i] = B2G_BAD_PARAMS;
		return;
	}
	fe rx;
	pt_to_affine_x(rx, R);
	// s0 <- belt-hash(oid || R.x || H) mod 2^l (:226-229)
	u32 hv[8];
	hash_oid_2x32(S, hv, oid, rx.v, H, (const u8*)0, 0);
	// s1 <- (k - (s0 + 2^l) d - H) mod q (:231-238)
	u32 prod[13];
	{
		const u32 s0w[5] = {hv[0], hv[1], hv[2], hv[3], 1u};
		for (int j = 0; j < 13; ++j) prod[j] = 0;
		for (int a = 0; a < 5; ++a)
		{
			u64 carry = 0;
			for (int b = 0; b < 8; ++b)
			{
				const u64 t = (u64)s0w[a] * d[b] + prod[a + b] + carry;
				prod[a + b] = (u32)t, carry = t >> 32;
			}
			prod[a + 8] = (u32)carry;
		}
	}
	u32 s1[8];
	modq_reduce(s1, prod, 13);
	modq_sub(s1, k, s1);
	modq_sub(s1, s1, H);
	u8* o = sigs + 48 * i;
	for (int j = 0; j < 16; ++j) o[j] = (u8)(hv[j >> 2] >> (8 * (j & 3)));
	for (int j = 0; j < 32; ++j) o[16 + j] = (u8)(s1[j >> 2] >> (8 * (j & 3)));
	status[i] = B2G_OK;
}

// bignPubkeyCalc per item (bign_misc.c:369-412): Q = d G, 0 < d < q
__global__ void __launch_bounds__(BIGN_THREADS, BIGN_MIN_BLOCKS)
bign_pubkey_kernel(u32* __restrict__ status, u8* __restrict__ pubkeys, const u8* __restrict__ privkeys,
	u64 count, const uint4* __restrict__ gtab)
{
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	u32 d[8];
	load_u256(d, privkeys + 32 * i);
	if (u256_is_zero(d) || u256_geq(d, c_q))
	{
		status[i] = B2G_BAD_PRIVKEY;
		return;
	}
	pt R;
	pt_set_inf(R);
	{
		sc256 ks;
#pragma unroll
		for (int j = 0; j < 8; ++j) ks.w[j] = d[j];
		pt_add_mul_base(R, ks, gtab);
	}
	if (pt_is_inf(R))
	{
		status[i] = B2G_BAD_PARAMS;
		return;
	}
	fe x, y;
	pt_to_affine(x, y, R);
	fe_store(pubkeys + 64 * i, x), fe_store(pubkeys + 64 * i + 32, y);
	status[i] = B2G_OK;
}

// ecMulA per item (ec.c:497-525): b = d * a, affine in/out; ok = 0 iff the result is O
__global__ void __launch_bounds__(BIGN_THREADS, BIGN_MIN_BLOCKS)
ecp_mul_kernel(u8* __restrict__ out, int* __restrict__ ok, const u8* __restrict__ pts,
	const u8* __restrict__ scalars, u32 d_len, u64 count)
{
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	fe x, y;
	fe_load(x, pts + 64 * i), fe_load(y, pts + 64 * i + 32);
	sc256 k;
	load_scalar(k, scalars + (u64)d_len * i, d_len);
	pt R;
	pt_mul_var(R, k, (int)(8 * d_len), x, y);
	if (pt_is_inf(R))
	{
		ok[i] = 0;
		return;
	}
	pt_to_affine(x, y, R);
	fe_store(out + 64 * i, x), fe_store(out + 64 * i + 32, y);
	ok[i] = 1;
}

// ecAddMulA with the base point (ec.c:1183-1273): b = d * a + k * G; ok = 0 iff the result is O
__global__ void __launch_bounds__(BIGN_THREADS, BIGN_MIN_BLOCKS)
ecp_addmul_kernel(u8* __restrict__ out, int* __restrict__ ok, const u8* __restrict__ pts,
	const u8* __restrict__ scalars, u32 d_len, const u8* __restrict__ kbase, u64 count,
	const uint4* __restrict__ gtab)
{
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	fe x, y;
	fe_load(x, pts + 64 * i), fe_load(y, pts + 64 * i + 32);
	sc256 k, kg;
	load_scalar(k, scalars + (u64)d_len * i, d_len);
	load_u256(kg.w, kbase + 32 * i);
	pt R;
	pt_mul_var(R, k, (int)(8 * d_len), x, y);
	pt_add_mul_base(R, kg, gtab);
	if (pt_is_inf(R))
	{
		ok[i] = 0;
		return;
	}
	pt_to_affine(x, y, R);
	fe_store(out + 64 * i, x), fe_store(out + 64 * i + 32, y);
	ok[i] = 1;
}

// ---------------------------------------------------------------- launchers (C ABI)
// q and yG are static constants, GTAB is built lazily; only the belt S-box needs uploading
extern "C" u32 b2g_bign_upload_tables(const u8 H[256]) { return belt_upload_H(H); }

static u32 bign_build_gtab(cudaStream_t st);
// the device entry points may be called from several host threads: build the table once
static u32 bign_ensure_gtab(cudaStream_t st)
{
	static std::mutex mu;
	if (g_gtab)
		return B2G_OK;
	std::lock_guard<std::mutex> lock(mu);
	return g_gtab ? B2G_OK : bign_build_gtab(st);
}
static u32 bign_build_gtab(cudaStream_t st)
{
	uint4* p = 0;
	if (cudaMalloc(&p, (size_t)BIGN_GN * BIGN_GE * 64) != cudaSuccess)
		return b2g_check_launch("cudaMalloc(gtab)");
	bign_gtab_kernel<<<(BIGN_GN * BIGN_GE + BIGN_THREADS - 1) / BIGN_THREADS, BIGN_THREADS, 0, st>>>(p);
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
	g_gtab = p;
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

static inline u32 bign_grid(size_t count) { return (u32)((count + BIGN_THREADS - 1) / BIGN_THREADS); }

extern "C" u32 b2g_bignVerifyBatch_dev(void* d_status, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_sigs, const void* d_pubkeys, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	OidArg oid;
	if ((e = make_oid(oid, oid_der, oid_len))) return e;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
	if ((e = bign_ensure_gtab(st))) return e;
	bign_verify_kernel<<<bign_grid(count), BIGN_THREADS, 0, st>>>((u32*)d_status, (const u8*)d_hashes,
		(const u8*)d_sigs, (const u8*)d_pubkeys, count, oid, g_gtab);
	b2g_note_launch();
	return b2g_check_launch("bign_verify_kernel");
}

extern "C" u32 b2g_bignSign2Batch_t_dev(void* d_status, void* d_sigs, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, const void* t, size_t t_len, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	OidArg oid;
	TArg ta;
	if ((e = make_oid(oid, oid_der, oid_len))) return e;
	if (t_len > BIGN_MAX_T) return 119u;
	for (size_t i = 0; i < BIGN_MAX_T; ++i) ta.t[i] = (t && i < t_len) ? ((const u8*)t)[i] : 0;
	ta.len = t ? (u32)t_len : 0;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
	if ((e = bign_ensure_gtab(st))) return e;
	bign_sign2_kernel<<<bign_grid(count), BIGN_THREADS, 0, st>>>((u32*)d_status, (u8*)d_sigs,
		(const u8*)d_hashes, (const u8*)d_privkeys, count, oid, ta, g_gtab);
	b2g_note_launch();
	return b2g_check_launch("bign_sign2_kernel");
}

extern "C" u32 b2g_bignSign2Batch_dev(void* d_status, void* d_sigs, const u8 oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, void* stream)
{
	return b2g_bignSign2Batch_t_dev(d_status, d_sigs, oid_der, oid_len, d_hashes, d_privkeys, count, 0, 0, stream);
}

extern "C" u32 b2g_bignPubkeyCalcBatch_dev(void* d_status, void* d_pubkeys, const void* d_privkeys,
	size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_status & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
	if ((e = bign_ensure_gtab(st))) return e;
	bign_pubkey_kernel<<<bign_grid(count), BIGN_THREADS, 0, st>>>((u32*)d_status, (u8*)d_pubkeys,
		(const u8*)d_privkeys, count, g_gtab);
	b2g_note_launch();
	return b2g_check_launch("bign_pubkey_kernel");
}

extern "C" u32 b2g_ecMulABatch_dev(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (d_len == 0 || d_len > 32) return B2G_BAD_INPUT;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_ok & 3) return B2G_BAD_INPUT;
	ecp_mul_kernel<<<bign_grid(count), BIGN_THREADS, 0, (cudaStream_t)stream>>>((u8*)d_b, (int*)d_ok,
		(const u8*)d_a, (const u8*)d_d, (u32)d_len, count);
	b2g_note_launch();
	return b2g_check_launch("ecp_mul_kernel");
}

extern "C" u32 b2g_ecAddMulABatch_dev(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	const void* d_k, size_t count, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (d_len == 0 || d_len > 32) return B2G_BAD_INPUT;
	if (count == 0) return B2G_OK;
	if ((uintptr_t)d_ok & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
	if ((e = bign_ensure_gtab(st))) return e;
	ecp_addmul_kernel<<<bign_grid(count), BIGN_THREADS, 0, st>>>((u8*)d_b, (int*)d_ok, (const u8*)d_a,
		(const u8*)d_d, (u32)d_len, (const u8*)d_k, count, g_gtab);
	b2g_note_launch();
	return b2g_check_launch("ecp_addmul_kernel");
}
