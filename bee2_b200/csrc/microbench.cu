// microbench.cu — measured integer-issue peaks of the device (denominators for the
// "issue roofline" bench.py reports next to the HBM roofline: bash/belt/bign are bound by
// the ALU / FMA-integer / shared-memory pipes).
//
// Each kernel runs ILP independent dependency chains per thread of one instruction kind,
// 1024 threads x (2 x SM count) CTAs, and reports lane-operations per second from CUDA
// events. The multipliers / shift amounts come from kernel arguments so ptxas cannot fold
// or strength-reduce the chains.
#include "common.cuh"

#define MB_ILP 8
#define MB_UNROLL 32

enum { MB_LOP3 = 0, MB_SHF = 1, MB_PRMT = 2, MB_IADD3 = 3, MB_IMAD = 4, MB_IMAD_WIDE = 5, MB_LDS = 6, MB_MIX_LOP3_IMADW = 7,
	MB_MIX_LOP3_IMAD = 8, MB_MIX_LOP3_FFMA = 9, MB_IMAD_HI = 10, MB_MIX_LOP3_LDS = 11, MB_FFMA = 12, MB_DFMA = 13, MB_MIX_DFMA_IMADW = 14, MB_MIX_IMAD_IMADW = 15,
	MB_MIX_IADDX_IMADW = 16, MB_MADC_ROW = 17, MB_ADDC_ROW = 18, MB_MADWIDE_ROW = 19, MB_MIX_MADC_ADDC = 20, MB_MIX_MADC_2ADDC = 21, MB_MIX_MADWIDE_ADDC = 22 };
#define MB_IS_MIX(k) ((k) == MB_MIX_LOP3_IMADW || (k) == MB_MIX_LOP3_IMAD || (k) == MB_MIX_LOP3_FFMA || (k) == MB_MIX_LOP3_LDS || \
	(k) == MB_MIX_DFMA_IMADW || (k) == MB_MIX_IMAD_IMADW || (k) == MB_MIX_IADDX_IMADW)

template <int KIND>
__global__ void __launch_bounds__(1024) mb_kernel(u32* out, u32 iters, u32 a, u32 b, u32 sh)
{
	__shared__ u32 sm[32 * 64];
	u32 x[MB_ILP], y[MB_ILP], z[MB_ILP], v[MB_ILP];
	u64 w[MB_ILP], w2[MB_ILP];
	u32 p0[MB_ILP], p1[MB_ILP], p2[MB_ILP], p3[MB_ILP];
	float f[MB_ILP];
	double d[MB_ILP];
	const double da = 1.0 + 1e-9 * (a & 3), db = 1e-3 * (b & 7);
	const float fa = __uint_as_float(0x3F800001u + (a & 1)), fb = __uint_as_float(b & 0x007FFFFFu);
#pragma unroll
	for (int i = 0; i < MB_ILP; ++i)
	{
		x[i] = (KIND == MB_LDS || KIND == MB_MIX_LOP3_LDS) ? ((threadIdx.x + i) & 63u) << 5 : threadIdx.x * 2654435761u + i * a;
		w[i] = ((u64)x[i] << 32) | (b + i);
		y[i] = x[i] ^ b, f[i] = (float)(threadIdx.x + i), d[i] = (double)(threadIdx.x + i);
		z[i] = x[i] + a, v[i] = y[i] + b, w2[i] = w[i] ^ a;
		p0[i] = x[i] * 3u, p1[i] = y[i] * 5u, p2[i] = x[i] * 7u, p3[i] = y[i] * 9u;
	}
	for (u32 i = threadIdx.x; i < 32 * 64; i += blockDim.x)
		sm[i] = (i * 7 + a) & (63u << 5);   // next index: keeps the lane's own bank (multiple of 32 words)
	__syncthreads();
	const u32 lane = threadIdx.x & 31;
#pragma unroll 1
	for (u32 it = 0; it < iters; ++it)
	{
#pragma unroll
		for (int u = 0; u < MB_UNROLL; ++u)
		{
#pragma unroll
			for (int i = 0; i < MB_ILP; ++i)
			{
				if (KIND == MB_LOP3)
					asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
				else if (KIND == MB_SHF)
					asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(sh));
				else if (KIND == MB_PRMT)
					asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(sh));
				else if (KIND == MB_IADD3)
					asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
				else if (KIND == MB_IMAD)
					asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
				else if (KIND == MB_IMAD_WIDE)
					asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(a));
				else if (KIND == MB_LDS)
					x[i] = sm[x[i] + lane];
				else if (KIND == MB_MIX_LOP3_IMADW)
				{
					asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
					asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(sh), "r"(a));
				}
				else if (KIND == MB_MIX_LOP3_IMAD)
				{
					asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
					asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
				}
				else if (KIND == MB_MIX_LOP3_FFMA)
				{
					asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
					asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
				}
				else if (KIND == MB_FFMA)
					asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
				else if (KIND == MB_DFMA)
					asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
				else if (KIND == MB_MIX_DFMA_IMADW)
				{
					asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
					asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(sh), "r"(a));
				}
				else if (KIND == MB_MIX_IMAD_IMADW)
				{
					asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
					asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(sh), "r"(a));
				}
				else if (KIND == MB_MIX_IADDX_IMADW)
				{
					asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
					asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(sh), "r"(a));
				}
				else if (KIND == MB_MADC_ROW)
				{
					// the carry-chained row of gfp_asm.cuh mad_row<4>: 8 wide multiply-adds linked through CC.CF
					asm volatile("mad.lo.cc.u32 %0, %4, %8, %0;\n\tmadc.hi.cc.u32 %1, %4, %8, %1;\n\t"
						"madc.lo.cc.u32 %2, %5, %8, %2;\n\tmadc.hi.cc.u32 %3, %5, %8, %3;\n\t"
						"madc.lo.cc.u32 %0, %6, %8, %0;\n\tmadc.hi.cc.u32 %1, %6, %8, %1;\n\t"
						"madc.lo.cc.u32 %2, %7, %8, %2;\n\tmadc.hi.cc.u32 %3, %7, %8, %3;\n\t"
						"madc.lo.cc.u32 %0, %4, %7, %0;\n\tmadc.hi.cc.u32 %1, %4, %7, %1;\n\t"
						"madc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.cc.u32 %3, %5, %6, %3;\n\t"
						"madc.lo.cc.u32 %0, %6, %5, %0;\n\tmadc.hi.cc.u32 %1, %6, %5, %1;\n\t"
						"madc.lo.cc.u32 %2, %7, %4, %2;\n\tmadc.hi.u32 %3, %7, %4, %3;"
						: "+r"(x[i]), "+r"(y[i]), "+r"(z[i]), "+r"(v[i]) : "r"(a), "r"(b), "r"(sh), "r"(a ^ b), "r"(b + i));
				}
				else if (KIND == MB_MADWIDE_ROW)
				{
					// the same 8 products as independent plain wide multiply-adds (no carry links)
					asm volatile("mad.wide.u32 %0, %2, %6, %0;\n\tmad.wide.u32 %1, %3, %6, %1;\n\t"
						"mad.wide.u32 %0, %4, %6, %0;\n\tmad.wide.u32 %1, %5, %6, %1;\n\t"
						"mad.wide.u32 %0, %2, %5, %0;\n\tmad.wide.u32 %1, %3, %4, %1;\n\t"
						"mad.wide.u32 %0, %4, %3, %0;\n\tmad.wide.u32 %1, %5, %2, %1;"
						: "+l"(w[i]), "+l"(w2[i]) : "r"(a), "r"(b), "r"(sh), "r"(a ^ b), "r"(b + i));
				}
				else if (KIND == MB_ADDC_ROW)
				{
					asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %4;\n\taddc.cc.u32 %3, %3, %5;\n\t"
						"addc.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.cc.u32 %2, %2, %5;\n\taddc.u32 %3, %3, %4;"
						: "+r"(x[i]), "+r"(y[i]), "+r"(z[i]), "+r"(v[i]) : "r"(a), "r"(b));
				}
				else if (KIND == MB_MIX_MADC_ADDC)
				{
					asm volatile("mad.lo.cc.u32 %0, %4, %8, %0;\n\tmadc.hi.cc.u32 %1, %4, %8, %1;\n\t"
						"madc.lo.cc.u32 %2, %5, %8, %2;\n\tmadc.hi.cc.u32 %3, %5, %8, %3;\n\t"
						"madc.lo.cc.u32 %0, %6, %8, %0;\n\tmadc.hi.cc.u32 %1, %6, %8, %1;\n\t"
						"madc.lo.cc.u32 %2, %7, %8, %2;\n\tmadc.hi.cc.u32 %3, %7, %8, %3;\n\t"
						"madc.lo.cc.u32 %0, %4, %7, %0;\n\tmadc.hi.cc.u32 %1, %4, %7, %1;\n\t"
						"madc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.cc.u32 %3, %5, %6, %3;\n\t"
						"madc.lo.cc.u32 %0, %6, %5, %0;\n\tmadc.hi.cc.u32 %1, %6, %5, %1;\n\t"
						"madc.lo.cc.u32 %2, %7, %4, %2;\n\tmadc.hi.u32 %3, %7, %4, %3;"
						: "+r"(x[i]), "+r"(y[i]), "+r"(z[i]), "+r"(v[i]) : "r"(a), "r"(b), "r"(sh), "r"(a ^ b), "r"(b + i));
					asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %4;\n\taddc.cc.u32 %3, %3, %5;\n\t"
						"addc.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.cc.u32 %2, %2, %5;\n\taddc.u32 %3, %3, %4;"
						: "+r"(p0[i]), "+r"(p1[i]), "+r"(p2[i]), "+r"(p3[i]) : "r"(a), "r"(b));
				}
				else if (KIND == MB_MIX_MADC_2ADDC)
				{
					asm volatile("mad.lo.cc.u32 %0, %4, %8, %0;\n\tmadc.hi.cc.u32 %1, %4, %8, %1;\n\t"
						"madc.lo.cc.u32 %2, %5, %8, %2;\n\tmadc.hi.cc.u32 %3, %5, %8, %3;\n\t"
						"madc.lo.cc.u32 %0, %6, %8, %0;\n\tmadc.hi.cc.u32 %1, %6, %8, %1;\n\t"
						"madc.lo.cc.u32 %2, %7, %8, %2;\n\tmadc.hi.cc.u32 %3, %7, %8, %3;\n\t"
						"madc.lo.cc.u32 %0, %4, %7, %0;\n\tmadc.hi.cc.u32 %1, %4, %7, %1;\n\t"
						"madc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.cc.u32 %3, %5, %6, %3;\n\t"
						"madc.lo.cc.u32 %0, %6, %5, %0;\n\tmadc.hi.cc.u32 %1, %6, %5, %1;\n\t"
						"madc.lo.cc.u32 %2, %7, %4, %2;\n\tmadc.hi.u32 %3, %7, %4, %3;"
						: "+r"(x[i]), "+r"(y[i]), "+r"(z[i]), "+r"(v[i]) : "r"(a), "r"(b), "r"(sh), "r"(a ^ b), "r"(b + i));
					asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %4;\n\taddc.cc.u32 %3, %3, %5;\n\t"
						"addc.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.cc.u32 %2, %2, %5;\n\taddc.u32 %3, %3, %4;"
						: "+r"(p0[i]), "+r"(p1[i]), "+r"(p2[i]), "+r"(p3[i]) : "r"(a), "r"(b));
					asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %4;\n\taddc.cc.u32 %3, %3, %5;\n\t"
						"addc.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.cc.u32 %2, %2, %5;\n\taddc.u32 %3, %3, %4;"
						: "+r"(p2[i]), "+r"(p3[i]), "+r"(p0[i]), "+r"(p1[i]) : "r"(a), "r"(b));
				}
				else if (KIND == MB_MIX_MADWIDE_ADDC)
				{
					asm volatile("mad.wide.u32 %0, %2, %6, %0;\n\tmad.wide.u32 %1, %3, %6, %1;\n\t"
						"mad.wide.u32 %0, %4, %6, %0;\n\tmad.wide.u32 %1, %5, %6, %1;\n\t"
						"mad.wide.u32 %0, %2, %5, %0;\n\tmad.wide.u32 %1, %3, %4, %1;\n\t"
						"mad.wide.u32 %0, %4, %3, %0;\n\tmad.wide.u32 %1, %5, %2, %1;"
						: "+l"(w[i]), "+l"(w2[i]) : "r"(a), "r"(b), "r"(sh), "r"(a ^ b), "r"(b + i));
					asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %4;\n\taddc.cc.u32 %3, %3, %5;\n\t"
						"addc.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.cc.u32 %2, %2, %5;\n\taddc.u32 %3, %3, %4;"
						: "+r"(p0[i]), "+r"(p1[i]), "+r"(p2[i]), "+r"(p3[i]) : "r"(a), "r"(b));
				}
				else if (KIND == MB_IMAD_HI)
					asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
				else if (KIND == MB_MIX_LOP3_LDS)
				{
					asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(a), "r"(b));
					x[i] = sm[x[i] + lane];
				}
			}
		}
	}
	u32 acc = 0;
#pragma unroll
	for (int i = 0; i < MB_ILP; ++i)
		acc ^= p0[i] ^ p1[i] ^ p2[i] ^ p3[i] ^ x[i] ^ y[i] ^ z[i] ^ v[i] ^ (u32)w2[i] ^ (u32)(w2[i] >> 32) ^ __float_as_uint(f[i]) ^ (u32)__double2hiint(d[i]) ^ (u32)__double2loint(d[i]) ^ (u32)w[i] ^ (u32)(w[i] >> 32);
	if (acc == 0x12345678u)
		out[0] = acc;
}

template <int KIND> static double mb_run(u32 iters, u32* d_out)
{
	const int grid = 2 * b2g_sm_count();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	const u32 sh = KIND == MB_PRMT ? 0x2103u : 7u;
	mb_kernel<KIND><<<grid, 1024>>>(d_out, 4, 0x9E3779B1u, 0x85EBCA6Bu, sh);   // warm-up
	cudaEventRecord(e0);
	mb_kernel<KIND><<<grid, 1024>>>(d_out, iters, 0x9E3779B1u, 0x85EBCA6Bu, sh);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0), cudaEventDestroy(e1);
	b2g_note_launch(), b2g_note_launch();
	if (b2g_check_launch("mb_kernel") || ms <= 0)
		return -1.0;
	const double per_thread = (double)iters * MB_UNROLL * MB_ILP * (KIND == MB_MIX_IADDX_IMADW ? 3 : (KIND == MB_MADC_ROW || KIND == MB_ADDC_ROW || KIND == MB_MADWIDE_ROW) ? 8 : (KIND == MB_MIX_MADC_ADDC || KIND == MB_MIX_MADWIDE_ADDC) ? 16 : KIND == MB_MIX_MADC_2ADDC ? 24 : MB_IS_MIX(KIND) ? 2 : 1);
	return per_thread * 1024.0 * grid / (ms * 1e-3);
}

// lane-operations per second of one instruction kind on the whole chip (or < 0 on error)
extern "C" double b2g_microbench(int kind, unsigned iters)
{
	if (b2g_ensure_device())
		return -1.0;
	u32* d_out = 0;
	if (cudaMalloc(&d_out, 64) != cudaSuccess)
		return -1.0;
	double r = -1.0;
	if (iters == 0)
		iters = 2000;
	switch (kind)
	{
	case MB_LOP3: r = mb_run<MB_LOP3>(iters, d_out); break;
	case MB_SHF: r = mb_run<MB_SHF>(iters, d_out); break;
	case MB_PRMT: r = mb_run<MB_PRMT>(iters, d_out); break;
	case MB_IADD3: r = mb_run<MB_IADD3>(iters, d_out); break;
	case MB_IMAD: r = mb_run<MB_IMAD>(iters, d_out); break;
	case MB_IMAD_WIDE: r = mb_run<MB_IMAD_WIDE>(iters, d_out); break;
	case MB_LDS: r = mb_run<MB_LDS>(iters, d_out); break;
	case MB_MIX_LOP3_IMADW: r = mb_run<MB_MIX_LOP3_IMADW>(iters, d_out); break;
	case MB_MIX_LOP3_IMAD: r = mb_run<MB_MIX_LOP3_IMAD>(iters, d_out); break;
	case MB_MIX_LOP3_FFMA: r = mb_run<MB_MIX_LOP3_FFMA>(iters, d_out); break;
	case MB_IMAD_HI: r = mb_run<MB_IMAD_HI>(iters, d_out); break;
	case MB_MIX_LOP3_LDS: r = mb_run<MB_MIX_LOP3_LDS>(iters, d_out); break;
	case MB_FFMA: r = mb_run<MB_FFMA>(iters, d_out); break;
	case MB_DFMA: r = mb_run<MB_DFMA>(iters, d_out); break;
	case MB_MADC_ROW: r = mb_run<MB_MADC_ROW>(iters / 4, d_out); break;
	case MB_ADDC_ROW: r = mb_run<MB_ADDC_ROW>(iters / 4, d_out); break;
	case MB_MADWIDE_ROW: r = mb_run<MB_MADWIDE_ROW>(iters / 4, d_out); break;
	case MB_MIX_MADC_ADDC: r = mb_run<MB_MIX_MADC_ADDC>(iters / 8, d_out); break;
	case MB_MIX_MADC_2ADDC: r = mb_run<MB_MIX_MADC_2ADDC>(iters / 8, d_out); break;
	case MB_MIX_MADWIDE_ADDC: r = mb_run<MB_MIX_MADWIDE_ADDC>(iters / 8, d_out); break;
	case MB_MIX_DFMA_IMADW: r = mb_run<MB_MIX_DFMA_IMADW>(iters, d_out); break;
	case MB_MIX_IMAD_IMADW: r = mb_run<MB_MIX_IMAD_IMADW>(iters, d_out); break;
	case MB_MIX_IADDX_IMADW: r = mb_run<MB_MIX_IADDX_IMADW>(iters, d_out); break;
	default: break;
	}
	cudaFree(d_out);
	return r;
}
