/*
 * engine.c — host-side runtime of the batch engine (C): device bring-up, constant
 * tables, error mapping, pinned/device allocators and the pipelined workspace used by
 * the host-pointer entry points. No cryptographic computation happens here.
 */
#include "engine.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static pthread_mutex_t g_init_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_mutex_t g_api_mu = PTHREAD_MUTEX_INITIALIZER;
static int g_ready;          /* 1 after a successful bring-up */
static int g_want_dev = -1;  /* device requested through b2g_init */
static int g_sms;
static volatile u64 g_launches;
static __thread char g_err[384];
static b2g_slot g_slots[B2G_NSLOT];

const char* b2g_last_error(void) { return g_err; }
u64 b2g_launch_count(void) { return g_launches; }
void b2g_note_launch(void) { __sync_fetch_and_add(&g_launches, 1); }
int b2g_sm_count(void) { return g_sms ? g_sms : 148; }
void b2g_lock(void) { pthread_mutex_lock(&g_api_mu); }
void b2g_unlock(void) { pthread_mutex_unlock(&g_api_mu); }

u32 b2g_cuda_fail(cudaError_t e, const char* what)
{
	snprintf(g_err, sizeof g_err, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
	if (e == cudaErrorMemoryAllocation)
		return ERR_OUTOFMEMORY;
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
		return ERR_B2G_NO_DEVICE;
	return ERR_B2G_CUDA;
}

u32 b2g_check_launch(const char* what)
{
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? ERR_OK : b2g_cuda_fail(e, what);
}

void b2g_die(const char* fn, u32 code)
{
	fprintf(stderr, "bee2_b200: %s failed on the GPU path (err %u: %s); there is no CPU fallback\n",
		fn, code, g_err);
	abort();
}

/* belt S-box from its definition: H[10] = 0, H[(11 + x) % 256] = 0x8E * z^(116 x) in
   GF(2)[z]/(z^8 + z^7 + z^6 + z + 1) — the reference keeps it as a literal table
   (belt_block.c:43-60) and regenerates it this way in test/crypto/belt_test.c:27-57. */
static octet g_H[256];
static void gen_beltH(void)
{
	unsigned x, i;
	g_H[10] = 0, g_H[11] = 0x8E;
	for (x = 12; x < 266; ++x)
	{
		unsigned t = g_H[(x - 1) % 256];
		for (i = 0; i < 116; ++i)
			t = (t >> 1) | ((unsigned)__builtin_parity(t & 0x63) << 7);
		g_H[x % 256] = (octet)t;
	}
}
const octet* beltH(void)
{
	if (g_H[11] != 0x8E)
		gen_beltH();
	return g_H;
}

u32 b2g_ensure_device(void)
{
	cudaError_t e;
	int n = 0, dev = 0, i;
	u32 code;
	if (g_ready)
		return ERR_OK;
	pthread_mutex_lock(&g_init_mu);
	if (g_ready)
	{
		pthread_mutex_unlock(&g_init_mu);
		return ERR_OK;
	}
	code = ERR_OK;
	e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
	{
		if (e == cudaSuccess)
			snprintf(g_err, sizeof g_err, "no CUDA device");
		else
			b2g_cuda_fail(e, "cudaGetDeviceCount");
		(void)cudaGetLastError();
		pthread_mutex_unlock(&g_init_mu);
		return ERR_B2G_NO_DEVICE;
	}
	if (g_want_dev >= 0)
	{
		if ((e = cudaSetDevice(g_want_dev)) != cudaSuccess)
			code = b2g_cuda_fail(e, "cudaSetDevice");
	}
	if (!code && (e = cudaGetDevice(&dev)) != cudaSuccess)
		code = b2g_cuda_fail(e, "cudaGetDevice");
	if (!code && (e = cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
		code = b2g_cuda_fail(e, "cudaDeviceGetAttribute");
	if (!code)
		code = b2g_belt_upload_tables(beltH());
	if (!code)
		code = b2g_beltdwp_upload_tables(beltH());
	if (!code)
		code = b2g_bash_upload_tables();
	if (!code)
		code = b2g_bign_upload_tables(beltH());
	for (i = 0; !code && i < B2G_NSLOT; ++i)
		if ((e = cudaStreamCreateWithFlags(&g_slots[i].stream, cudaStreamNonBlocking)) != cudaSuccess)
			code = b2g_cuda_fail(e, "cudaStreamCreate");
	if (!code)
		g_ready = 1;
	pthread_mutex_unlock(&g_init_mu);
	return code;
}

err_t b2g_init(int device)
{
	if (g_ready)
	{
		int cur = -1;
		cudaGetDevice(&cur);
		return (device < 0 || cur == device) ? ERR_OK : ERR_BAD_INPUT;
	}
	g_want_dev = device;
	return b2g_ensure_device();
}

b2g_slot* b2g_slot_get(int i) { return &g_slots[i % B2G_NSLOT]; }

u32 b2g_slot_buf(b2g_slot* s, int which, size_t bytes, void** out)
{
	cudaError_t e;
	if (bytes == 0)
		bytes = 16;
	if (s->cap[which] < bytes)
	{
		/* the slot's stream may still be using the old buffer */
		if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess)
			return b2g_cuda_fail(e, "cudaStreamSynchronize");
		if (s->buf[which])
			cudaFree(s->buf[which]);
		s->buf[which] = 0, s->cap[which] = 0;
		bytes = (bytes + 0xFFFFF) & ~(size_t)0xFFFFF;
		if ((e = cudaMalloc(&s->buf[which], bytes)) != cudaSuccess)
			return b2g_cuda_fail(e, "cudaMalloc(workspace)");
		s->cap[which] = bytes;
	}
	*out = s->buf[which];
	return ERR_OK;
}

size_t b2g_chunk_units(size_t unit_bytes, size_t target_bytes)
{
	size_t u = target_bytes / (unit_bytes ? unit_bytes : 1);
	/* multiple of 148 SMs x 1024 threads keeps grids whole; at least one */
	if (u > 148 * 1024)
		u -= u % (148 * 1024);
	return u ? u : 1;
}

void* b2g_host_alloc(size_t bytes)
{
	void* p = 0;
	if (b2g_ensure_device())
		return 0;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess)
	{
		(void)cudaGetLastError();
		return 0;
	}
	return p;
}
void b2g_host_free(void* p) { if (p) cudaFreeHost(p); }

void* b2g_dev_alloc(size_t bytes)
{
	void* p = 0;
	if (b2g_ensure_device())
		return 0;
	if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess)
	{
		(void)cudaGetLastError();
		return 0;
	}
	return p;
}
void b2g_dev_free(void* p) { if (p) cudaFree(p); }

err_t b2g_memcpy_h2d(void* d, const void* h, size_t n)
{
	cudaError_t e = cudaMemcpy(d, h, n, cudaMemcpyHostToDevice);
	return e == cudaSuccess ? ERR_OK : b2g_cuda_fail(e, "cudaMemcpy(H2D)");
}
err_t b2g_memcpy_d2h(void* h, const void* d, size_t n)
{
	cudaError_t e = cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost);
	return e == cudaSuccess ? ERR_OK : b2g_cuda_fail(e, "cudaMemcpy(D2H)");
}
err_t b2g_sync(void)
{
	cudaError_t e = cudaDeviceSynchronize();
	return e == cudaSuccess ? ERR_OK : b2g_cuda_fail(e, "cudaDeviceSynchronize");
}
