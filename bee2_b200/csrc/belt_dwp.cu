// belt_dwp.cu — the belt-DWP authentication tag (STB 34.101.31 "data wrap") for sm_100a.
//
// Replaces the block loop of beltDWPStepI / StepA / StepG (belt_dwp.c:73-207): a Horner
// evaluation t <- (t ^ B_i) * r over GF(2^128), r = E_K(E_K(iv)), t_0 = beltH()[0..16), over the
// open data, the critical data (both zero-padded to whole blocks) and the length block, followed
// by mac = E_K(t)[0..8). Encryption itself is the belt-CTR kernel of belt.cu.
//
// The Horner chain is sequential in the reference; here it is a polynomial evaluation
//     t_N = t_0 r^N  ^  sum_i B_i r^(N - i)          (blocks B_0 .. B_(N-1))
// cut into one contiguous chunk of 32 K blocks per WARP, interleaved over its lanes: lane L takes
// blocks L, L + 32, L + 64, ... of the chunk, so the 32 lanes of a load read 512 contiguous octets
// (round 1 gave every THREAD a contiguous chunk: 32 different 128-byte lines per load, 32 LSU
// wavefronts instead of 4 — a fifth of that kernel's time, profiles/README.md r02). A lane runs
// Horner in R = r^32 (multiplication by the fixed R through conflict-free 4-bit window tables in
// shared memory), then its value is weighted with r^(1..32) (a 33-entry table of small powers), the
// lanes are XOR-reduced by warp shuffles, the warp's value is weighted with r^(blocks after the
// chunk) = R^q r^s (square-and-multiply in R: squarings are bit spreads) and one atomicXor per warp
// and word adds it to the total. A one-thread kernel encrypts the sum.
#include "belt_dev.cuh"
#include "gf128.cuh"

#define DWP_THREADS 256
#define DWP_MIN_CHUNK 32

struct DwpArgs
{
	const u8* open;    // src2: open (authenticated only) data
	u64 n2;
	const u8* crit;    // ciphertext
	u64 n1;
	u64 nI, nA, N;     // blocks of open data, of critical data, total (+1 length block)
	u64 K;             // blocks per thread
	BeltKey key;
	uint4 s;           // E_K(iv)
	u32 r_is_s;        // belt-CHE: r = s = E_K(iv) (belt_che.c:54-57); belt-DWP: r = E_K(s)
	u32* acc;          // 4 words, zeroed: XOR of the weighted chunk values
	uint4 t0;          // streaming form: running value t the chain starts from ...
	u32 t0_given;      // ... instead of t_0 = beltH()[0..16)
	u32 no_len;        // streaming form: no length block after the data
};

// block i (0-based) of the sequence open || critical || length, zero-padded
__device__ __forceinline__ gf128 dwp_block(const DwpArgs& a, u64 i)
{
	gf128 b;
	if (i >= a.nI + a.nA)
	{
		const u64 l0 = a.n2 << 3, l1 = a.n1 << 3;   // belt_dwp.c:185-187: |I| || |A| in bits
		b.w[0] = (u32)l0, b.w[1] = (u32)(l0 >> 32), b.w[2] = (u32)l1, b.w[3] = (u32)(l1 >> 32);
		return b;
	}
	const bool in_open = i < a.nI;
	const u8* base = in_open ? a.open : a.crit;
	const u64 off = 16 * (in_open ? i : i - a.nI);
	const u64 total = in_open ? a.n2 : a.n1;
	const u8* p = base + off;
	if (off + 16 <= total && ((uintptr_t)p & 15) == 0)
	{
		const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(p));   // a warp reads 512 contiguous octets
		b.w[0] = v.x, b.w[1] = v.y, b.w[2] = v.z, b.w[3] = v.w;
		return b;
	}
	const u32 valid = (u32)(total - off < 16 ? total - off : 16);
#pragma unroll
	for (int k = 0; k < 4; ++k)
	{
		u32 w = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if ((u32)(4 * k + j) < valid)
				w |= (u32)p[4 * k + j] << (8 * j);
		b.w[k] = w;
	}
	return b;
}

__global__ void __launch_bounds__(DWP_THREADS) belt_dwp_mac_kernel(const DwpArgs a)
{
	extern __shared__ __align__(1024) u8 sm[];
	u8* tab = sm;                                             // GF_TAB_BYTES
	u32* sbox = reinterpret_cast<u32*>(sm + GF_TAB_BYTES);    // 256 words
	u32* rsh = sbox + 256;                                    // 4 words: r
	BeltSmallT::fill(sbox);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		// r <- E_K(s) (belt_dwp.c:52-55)
		const BeltSmallT S(sbox);
		u32 x0 = a.s.x, x1 = a.s.y, x2 = a.s.z, x3 = a.s.w;
		if (!a.r_is_s)
			belt_encr(S, x0, x1, x2, x3, a.key.k);
		rsh[0] = x0, rsh[1] = x1, rsh[2] = x2, rsh[3] = x3;
	}
	__syncthreads();
	gf128 r;
#pragma unroll
	for (int i = 0; i < 4; ++i) r.w[i] = rsh[i];
	// small powers r^0 .. r^32 (thread k computes r^k by square-and-multiply with the generic product)
	gf128* pw = reinterpret_cast<gf128*>(rsh + 4);
	if (threadIdx.x <= 32)
	{
		gf128 acc = {{1, 0, 0, 0}};
		for (int bit = 5; bit >= 0; --bit)
		{
			acc = gf_sqr(acc);
			if (threadIdx.x >> bit & 1u) acc = gf_mul(acc, r);
		}
		pw[threadIdx.x] = acc;
	}
	// tables of R = r^32
	gf128 R = r;
#pragma unroll
	for (int i = 0; i < 5; ++i) R = gf_sqr(R);
	gf_tab_build(tab, R);
	__syncthreads();

	const u32 lane = threadIdx.x & 31u;
	const u64 warp = ((u64)blockIdx.x * DWP_THREADS + threadIdx.x) >> 5;
	const u64 b0 = warp * 32 * a.K;                    // the warp's chunk [b0, b1)
	gf128 w = {{0, 0, 0, 0}};
	if (b0 < a.N)
	{
		const u64 b1 = b0 + 32 * a.K < a.N ? b0 + 32 * a.K : a.N;
		gf128 acc = {{0, 0, 0, 0}};
		u64 last = 0;
		bool any = false;
#pragma unroll 1
		for (u64 i = b0 + lane; i < b1; i += 32)
		{
			gf128 blk = dwp_block(a, i);
			if (i == 0)
			{
				// t_0 = beltH()[0..16) (belt_dwp.c:60), or the running t of a streaming state: the chain
				// (t ^ B_0) r ... = t_0 r^N ^ sum B_i r^(N-i), so t_0 simply joins block 0
				const u32* H32 = reinterpret_cast<const u32*>(c_beltH);
				gf128 t0;
#pragma unroll
				for (int k = 0; k < 4; ++k) t0.w[k] = H32[k];
				if (a.t0_given)
					t0.w[0] = a.t0.x, t0.w[1] = a.t0.y, t0.w[2] = a.t0.z, t0.w[3] = a.t0.w;
				blk = gf_xor(blk, t0);
			}
			acc = any ? gf_xor(gf_mul_tab(tab, acc), blk) : blk;   // acc <- acc R ^ B
			any = true, last = i;
		}
		// the lane's value counts from its last block: weight r^(b1 - last), 1 <= b1 - last <= 32
		if (any)
			w = gf_mul(acc, pw[b1 - last]);
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			u32 v = w.w[k];
#pragma unroll
			for (int d = 16; d; d >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, d);
			w.w[k] = v;
		}
		// the warp's value counts from b1: weight r^(N - b1) = R^q r^s
		const u64 e = a.N - b1;
		if (lane == 0 && e)
		{
			const gf128 Rq = gf_pow_r(tab, e >> 5);
			w = gf_mul(w, Rq);
			if (e & 31)
				w = gf_mul(w, pw[e & 31]);
		}
	}
	if (lane == 0)
	{
#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (w.w[k])
				atomicXor(a.acc + k, w.w[k]);
	}
}

// mac <- E_K(t)[0..8) (belt_dwp.c:196-206)
__global__ void belt_dwp_fin_kernel(u8* mac, const u32* acc, const BeltKey key)
{
	__shared__ u32 sbox[256];
	BeltSmallT::fill(sbox);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		const BeltSmallT S(sbox);
		u32 x0 = acc[0], x1 = acc[1], x2 = acc[2], x3 = acc[3];
		belt_encr(S, x0, x1, x2, x3, key.k);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			mac[i] = (u8)(x0 >> (8 * i)), mac[4 + i] = (u8)(x1 >> (8 * i));
	}
}

extern "C" u32 b2g_beltdwp_upload_tables(const u8 H[256])
{
	const u32 e = belt_upload_H(H);
	if (e) return e;
	if (cudaFuncSetAttribute(belt_dwp_mac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
			GF_TAB_BYTES + 2048) != cudaSuccess)
		return b2g_check_launch("cudaFuncSetAttribute(belt_dwp)");
	return B2G_OK;
}

// d_mac (8 octets) <- DWP tag of (open data d_open[n2], critical data d_crit[n1]) under the
// expanded key and ctr0 = E_K(iv); d_scratch: 16 octets of device scratch
static u32 dwp_mac_launch(void* d_mac, const void* d_crit, size_t n1, const void* d_open, size_t n2,
	const u32 key[8], const u32 ctr0[4], void* d_scratch, void* stream, u32 r_is_s);

extern "C" u32 b2g_beltDWPMac_dev(void* d_mac, const void* d_crit, size_t n1, const void* d_open, size_t n2,
	const u32 key[8], const u32 ctr0[4], void* d_scratch, void* stream)
{
	return dwp_mac_launch(d_mac, d_crit, n1, d_open, n2, key, ctr0, d_scratch, stream, 0);
}

// belt-CHE tag: same polynomial MAC with r = s0 = E_K(iv) (belt_che.c:218-239)
extern "C" u32 b2g_beltCHEMac_dev(void* d_mac, const void* d_crit, size_t n1, const void* d_open, size_t n2,
	const u32 key[8], const u32 s0[4], void* d_scratch, void* stream)
{
	return dwp_mac_launch(d_mac, d_crit, n1, d_open, n2, key, s0, d_scratch, stream, 1);
}

static u32 dwp_mac_launch(void* d_mac, const void* d_crit, size_t n1, const void* d_open, size_t n2,
	const u32 key[8], const u32 ctr0[4], void* d_scratch, void* stream, u32 r_is_s)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if ((uintptr_t)d_scratch & 3) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
	DwpArgs a;
	a.open = (const u8*)d_open, a.n2 = n2, a.crit = (const u8*)d_crit, a.n1 = n1;
	a.nI = ((u64)n2 + 15) / 16, a.nA = ((u64)n1 + 15) / 16, a.N = a.nI + a.nA + 1;
	a.t0 = make_uint4(0, 0, 0, 0), a.t0_given = 0, a.no_len = 0;
	for (int i = 0; i < 8; ++i) a.key.k[i] = key[i];
	a.s = make_uint4(ctr0[0], ctr0[1], ctr0[2], ctr0[3]);
	a.acc = (u32*)d_scratch;
	a.r_is_s = r_is_s;
	// one chunk per thread; at most 2 CTAs per SM worth of threads, at least DWP_MIN_CHUNK blocks each
	const u64 tmax = (u64)b2g_sm_count() * 2 * DWP_THREADS;
	u64 K = (a.N + tmax - 1) / tmax;
	if (K < DWP_MIN_CHUNK) K = DWP_MIN_CHUNK;
	a.K = K;
	const u64 nwarps = (a.N + 32 * K - 1) / (32 * K);   // one chunk of 32 K blocks per warp
	const u32 grid = (u32)((nwarps * 32 + DWP_THREADS - 1) / DWP_THREADS);
	if (cudaMemsetAsync(d_scratch, 0, 16, st) != cudaSuccess)
		return b2g_check_launch("cudaMemsetAsync(dwp)");
	belt_dwp_mac_kernel<<<grid, DWP_THREADS, GF_TAB_BYTES + 2048, st>>>(a);
	b2g_note_launch();
	if ((e = b2g_check_launch("belt_dwp_mac_kernel"))) return e;
	belt_dwp_fin_kernel<<<1, 32, 0, st>>>((u8*)d_mac, (const u32*)d_scratch, a.key);
	b2g_note_launch();
	return b2g_check_launch("belt_dwp_fin_kernel");
}

// Streaming form (beltDWPStepI / StepA / StepG, belt_dwp.c:73-207; belt-CHE twins belt_che.c:122-239):
// t <- Horner(t; blocks) with the given r, over nbytes octets (zero-padded to whole blocks), no length
// block, no final encryption. d_t receives the 16 octets of the new t.
extern "C" u32 b2g_beltPolyAbsorb_dev(void* d_t, const void* d_blocks, size_t nbytes, const u32 r[4],
	const u32 t0[4], void* d_scratch, void* stream)
{
	u32 e = b2g_ensure_device();
	if (e) return e;
	if (((uintptr_t)d_scratch & 3) || ((uintptr_t)d_t & 3)) return B2G_BAD_INPUT;
	cudaStream_t st = (cudaStream_t)stream;
	DwpArgs a;
	a.open = 0, a.n2 = 0, a.crit = (const u8*)d_blocks, a.n1 = nbytes;
	a.nI = 0, a.nA = ((u64)nbytes + 15) / 16, a.N = a.nA;
	for (int i = 0; i < 8; ++i) a.key.k[i] = 0;
	a.s = make_uint4(r[0], r[1], r[2], r[3]);
	a.r_is_s = 1;
	a.t0 = make_uint4(t0[0], t0[1], t0[2], t0[3]), a.t0_given = 1, a.no_len = 1;
	a.acc = (u32*)d_scratch;
	if (a.N == 0)
	{
		if (cudaMemcpyAsync(d_t, t0, 16, cudaMemcpyHostToDevice, st) != cudaSuccess)
			return b2g_check_launch("cudaMemcpyAsync(poly t)");
		return B2G_OK;
	}
	const u64 tmax = (u64)b2g_sm_count() * 2 * DWP_THREADS;
	u64 K = (a.N + tmax - 1) / tmax;
	if (K < DWP_MIN_CHUNK) K = DWP_MIN_CHUNK;
	a.K = K;
	const u64 nwarps = (a.N + 32 * K - 1) / (32 * K);
	const u32 grid = (u32)((nwarps * 32 + DWP_THREADS - 1) / DWP_THREADS);
	if (cudaMemsetAsync(d_scratch, 0, 16, st) != cudaSuccess)
		return b2g_check_launch("cudaMemsetAsync(poly)");
	belt_dwp_mac_kernel<<<grid, DWP_THREADS, GF_TAB_BYTES + 2048, st>>>(a);
	b2g_note_launch();
	if ((e = b2g_check_launch("belt_dwp_mac_kernel(absorb)"))) return e;
	if (cudaMemcpyAsync(d_t, d_scratch, 16, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
		return b2g_check_launch("cudaMemcpyAsync(poly t)");
	return B2G_OK;
}
