// common.cuh — shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

// err_t values mirrored from include/bee2_b200.h (reference include/bee2/core/err.h)
#define B2G_OK 0u
#define B2G_BAD_INPUT 109u
#define B2G_BAD_PARAMS 502u
#define B2G_BAD_PRIVKEY 504u
#define B2G_BAD_PUBKEY 505u
#define B2G_BAD_SIG 510u
#define B2G_ERR_NO_DEVICE 9001u
#define B2G_ERR_CUDA 9002u

extern "C" {
// engine.c
int b2g_sm_count(void);
int b2g_cur_dev(void);     // CUDA ordinal selected by the last b2g_ensure_device() on this thread
void b2g_note_launch(void);
u32 b2g_check_launch(const char* what);   // cudaGetLastError -> err_t, records text
u32 b2g_ensure_device(void);
}

__device__ __forceinline__ u32 rotl32(u32 x, int n) { return __funnelshift_l(x, x, n); }

// 128-bit streaming global accesses (data touched once: keep it out of L1)
__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ void stg_stream(uint4* p, uint4 v)
{
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
		:: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit streaming load (LDG.E.256, new on sm_100): 32 octets per thread in one request, p 32-byte aligned
__device__ __forceinline__ void ldg_stream256(const uint4* p, uint4& lo, uint4& hi)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p));
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier (cp.async.bulk; SASS UBLKCP) ----
// src / dst 16-byte aligned, bytes a multiple of 16. `src` may be pinned HOST memory (zero-copy input):
// the copy engine then reads it in large PCIe requests instead of one 32-byte sector per load instruction.
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(arrivals) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, u32 bytes, u64* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity)
{
	asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}"
		:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// shared -> global bulk store (cp.async.bulk, bulk-group completion). All threads that wrote the source
// call bulk_store_fence() and then synchronise; ONE thread then calls bulk_s2g() and waits in
// bulk_store_wait() before the shared buffer is reused or the CTA exits.
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* dst_global, const void* src_smem, u32 bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
		:: "l"(dst_global), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
