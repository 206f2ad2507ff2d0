/*
 * engine.c — host-side runtime of the batch engine (C): device bring-up, constant
 * tables, error mapping, pinned/device allocators and the pipelined workspace used by
 * the host-pointer entry points. No cryptographic computation happens here.
 */
#define _GNU_SOURCE
#include "engine.h"
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static pthread_mutex_t g_init_mu = PTHREAD_MUTEX_INITIALIZER;
/* One context per CUDA device the engine has been brought up on: constant tables, two pipeline
   slots (stream + growable buffers) and the mutex that serialises host-pointer calls on that
   device. Indexed by CUDA ordinal. */
typedef struct
{
	int dev;
	volatile int ready;
	int sms;
	pthread_mutex_t mu;
	b2g_slot slots[B2G_NSLOT];
} b2g_ctx;
static b2g_ctx g_ctx[B2G_MAX_DEV];
static volatile int g_primary = -1;     /* device of host-pointer calls from threads whose current device is not ours */
static int g_set[B2G_MAX_DEV];          /* devices the Batch entry points shard over (b2g_init_devices) */
static volatile int g_nset;
static volatile u64 g_launches;
static __thread char g_err[384];
static __thread b2g_ctx* t_ctx;         /* context selected by the last b2g_ensure_device() on this thread */
static __thread int t_bound = -1;       /* fan-out workers are pinned to one device */

const char* b2g_last_error(void) { return g_err; }
u64 b2g_launch_count(void) { return g_launches; }
void b2g_note_launch(void) { __sync_fetch_and_add(&g_launches, 1); }
int b2g_sm_count(void) { return t_ctx && t_ctx->sms ? t_ctx->sms : 148; }
int b2g_cur_dev(void) { return t_ctx ? t_ctx->dev : 0; }
int b2g_device_count(void) { return g_nset > 0 ? g_nset : 1; }
void b2g_lock(void)
{
	if (!t_ctx && b2g_ensure_device())
		return;
	pthread_mutex_lock(&t_ctx->mu);
}
void b2g_unlock(void)
{
	if (t_ctx)
		pthread_mutex_unlock(&t_ctx->mu);
}

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
	fprintf(stderr, "bee2_b200: %s failed on the GPU path (err %u: %s); no stock libbee2 behind this library "
		"to forward to (see INTEGRATION.md, overlay mode)\n", fn, code, g_err);
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
static pthread_once_t g_H_once = PTHREAD_ONCE_INIT;
const octet* beltH(void)
{
	pthread_once(&g_H_once, gen_beltH);
	return g_H;
}

/* bring the engine up on device `dev` (tables, streams); the device is current on return */
static u32 ctx_bring_up(int dev)
{
	b2g_ctx* c = &g_ctx[dev];
	cudaError_t e;
	u32 code = ERR_OK;
	int i;
	if (c->ready)
		return ERR_OK;
	pthread_mutex_lock(&g_init_mu);
	if (!c->ready)
	{
		if ((e = cudaSetDevice(dev)) != cudaSuccess)
			code = b2g_cuda_fail(e, "cudaSetDevice");
		if (!code && (e = cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
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
			if (!c->slots[i].stream &&
				(e = cudaStreamCreateWithFlags(&c->slots[i].stream, cudaStreamNonBlocking)) != cudaSuccess)
				code = b2g_cuda_fail(e, "cudaStreamCreate");
		if (!code)
		{
			pthread_mutex_init(&c->mu, 0);
			c->dev = dev;
			if (g_primary < 0)
				g_primary = dev;
			if (g_nset == 0)
				g_set[0] = dev, g_nset = 1;
			__sync_synchronize();
			c->ready = 1;
		}
	}
	pthread_mutex_unlock(&g_init_mu);
	return code;
}

/* Select the device this thread's call runs on, make it current and make sure the engine is up on
   it: a fan-out worker's pinned device; else the thread's current CUDA device if the engine is up
   there (one process per GPU under torchrun, or a caller that drives several devices itself);
   else the primary device (b2g_init, or the device current at first use). The CUDA current device
   is per host thread, so this runs at the start of EVERY entry point. */
u32 b2g_ensure_device(void)
{
	cudaError_t e;
	int n = 0, cur = 0, want;
	u32 code;
	if ((e = cudaGetDevice(&cur)) != cudaSuccess)
	{
		if ((e = cudaGetDeviceCount(&n)) != cudaSuccess || n <= 0)
		{
			if (e == cudaSuccess)
				snprintf(g_err, sizeof g_err, "no CUDA device");
			else
				b2g_cuda_fail(e, "cudaGetDeviceCount");
			(void)cudaGetLastError();
			return ERR_B2G_NO_DEVICE;
		}
		cur = 0;
	}
	if (t_bound >= 0)
		want = t_bound;
	else if (cur >= 0 && cur < B2G_MAX_DEV && g_ctx[cur].ready)
		want = cur;
	else
		want = g_primary >= 0 ? g_primary : cur;
	if (want < 0 || want >= B2G_MAX_DEV)
	{
		snprintf(g_err, sizeof g_err, "device ordinal %d out of range", want);
		return ERR_B2G_NO_DEVICE;
	}
	if (!g_ctx[want].ready)
	{
		if ((code = ctx_bring_up(want)))
		{
			(void)cudaGetLastError();
			return code;
		}
	}
	else if (cur != want && (e = cudaSetDevice(want)) != cudaSuccess)
		return b2g_cuda_fail(e, "cudaSetDevice");
	t_ctx = &g_ctx[want];
	return ERR_OK;
}

err_t b2g_init(int device)
{
	cudaError_t e;
	int n = 0;
	err_t code;
	if (device < 0)
		return b2g_ensure_device();
	if ((e = cudaGetDeviceCount(&n)) != cudaSuccess || n <= 0)
	{
		if (e == cudaSuccess)
			snprintf(g_err, sizeof g_err, "no CUDA device");
		else
			b2g_cuda_fail(e, "cudaGetDeviceCount");
		(void)cudaGetLastError();
		return ERR_B2G_NO_DEVICE;
	}
	if (device >= n || device >= B2G_MAX_DEV)
		return ERR_BAD_INPUT;
	if ((code = ctx_bring_up(device)))
		return code;
	if ((e = cudaSetDevice(device)) != cudaSuccess)
		return b2g_cuda_fail(e, "cudaSetDevice");
	t_ctx = &g_ctx[device];
	return ERR_OK;
}

/* In-process multi-device mode: bring the engine up on devices 0..n-1 (n <= 0: all) and let the
   host-pointer *Batch entry points shard their units over them (b2g_fanout). */
err_t b2g_init_devices(int n)
{
	cudaError_t e;
	int have = 0, d, back = -1;
	err_t code = ERR_OK;
	if ((e = cudaGetDeviceCount(&have)) != cudaSuccess || have <= 0)
	{
		if (e == cudaSuccess)
			snprintf(g_err, sizeof g_err, "no CUDA device");
		else
			b2g_cuda_fail(e, "cudaGetDeviceCount");
		(void)cudaGetLastError();
		return ERR_B2G_NO_DEVICE;
	}
	if (n <= 0 || n > have)
		n = have;
	if (n > B2G_MAX_DEV)
		n = B2G_MAX_DEV;
	cudaGetDevice(&back);
	for (d = 0; d < n && !code; ++d)
		code = ctx_bring_up(d);
	if (!code)
	{
		pthread_mutex_lock(&g_init_mu);
		for (d = 0; d < n; ++d)
			g_set[d] = d;
		g_nset = n;
		pthread_mutex_unlock(&g_init_mu);
	}
	if (back >= 0)
		cudaSetDevice(back);
	if (!code)
		code = b2g_ensure_device();
	return code;
}

/* ---- fan-out of a batch over the device set: contiguous shares, one host thread per device ---- */
typedef struct
{
	b2g_shard_fn fn;
	void* arg;
	size_t first, n;
	int dev;
	u32 code;
	char err[sizeof g_err];
} fan_job;

static void* fan_worker(void* p)
{
	fan_job* j = (fan_job*)p;
	t_bound = j->dev;
	j->code = b2g_ensure_device();
	if (!j->code)
		j->code = j->fn(j->arg, j->first, j->n);
	if (j->code)
		memcpy(j->err, g_err, sizeof g_err);
	return 0;
}

u32 b2g_fanout(size_t count, size_t grain, b2g_shard_fn fn, void* arg)
{
	fan_job jobs[B2G_MAX_DEV];
	pthread_t th[B2G_MAX_DEV];
	int k = g_nset, i, started = 0;
	size_t per, off = 0;
	u32 code = ERR_OK;
	if (grain == 0)
		grain = 1;
	if ((size_t)k > count / grain)
		k = (int)(count / grain);
	if (k <= 1 || t_bound >= 0)
		return fn(arg, 0, count);
	per = (count + (size_t)k - 1) / (size_t)k;
	per = (per + 255) & ~(size_t)255;          /* whole CTAs per device */
	for (i = 0; i < k && off < count; ++i)
	{
		jobs[i].fn = fn, jobs[i].arg = arg, jobs[i].dev = g_set[i], jobs[i].code = ERR_OK, jobs[i].err[0] = 0;
		jobs[i].first = off, jobs[i].n = count - off < per ? count - off : per;
		off += jobs[i].n;
		if (pthread_create(&th[i], 0, fan_worker, &jobs[i]) != 0)
		{
			/* could not start a thread: run this share here, on its device */
			const int keep = t_bound;
			fan_worker(&jobs[i]);
			t_bound = keep;
			th[i] = 0;
		}
		++started;
	}
	for (i = 0; i < started; ++i)
	{
		if (th[i])
			pthread_join(th[i], 0);
		if (jobs[i].code && !code)
			code = jobs[i].code, memcpy(g_err, jobs[i].err, sizeof g_err);
	}
	/* the caller's own selection may have been a different device: re-select */
	(void)b2g_ensure_device();
	return code;
}

b2g_slot* b2g_slot_get(int i)
{
	if (!t_ctx)
		(void)b2g_ensure_device();
	return &t_ctx->slots[i % B2G_NSLOT];
}

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

void b2g_slot_wipe(b2g_slot* s)
{
	int i;
	cudaStreamSynchronize(s->stream);
	for (i = 0; i < B2G_NBUF; ++i)
		if (s->buf[i] && s->cap[i])
			cudaMemsetAsync(s->buf[i], 0, s->cap[i], s->stream);
	cudaStreamSynchronize(s->stream);
	(void)cudaGetLastError();
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
	/* portable: usable from every device context of the fan-out set */
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess)
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

/* ---- CUDA IPC: a device buffer of one process mapped into its peers (one process per GPU) ---- */
err_t b2g_ipc_export(octet handle[64], void* dptr)
{
	cudaIpcMemHandle_t h;
	cudaError_t e;
	if (!handle || !dptr)
		return ERR_BAD_INPUT;
	if ((e = cudaIpcGetMemHandle(&h, dptr)) != cudaSuccess)
		return b2g_cuda_fail(e, "cudaIpcGetMemHandle");
	memcpy(handle, &h, sizeof h <= 64 ? sizeof h : 64);
	return ERR_OK;
}
err_t b2g_ipc_open(void** dptr, const octet handle[64])
{
	cudaIpcMemHandle_t h;
	cudaError_t e;
	if (!handle || !dptr)
		return ERR_BAD_INPUT;
	memcpy(&h, handle, sizeof h <= 64 ? sizeof h : 64);
	if ((e = cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess)
		return b2g_cuda_fail(e, "cudaIpcOpenMemHandle");
	return ERR_OK;
}
err_t b2g_ipc_close(void* dptr)
{
	cudaError_t e = cudaIpcCloseMemHandle(dptr);
	return e == cudaSuccess ? ERR_OK : b2g_cuda_fail(e, "cudaIpcCloseMemHandle");
}
/* raw async copies on a caller's stream (bench.py's copy-roof measurement) */
err_t b2g_memcpy_async(void* dst, const void* src, size_t n, int to_device, void* stream)
{
	cudaError_t e = cudaMemcpyAsync(dst, src, n, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
		(cudaStream_t)stream);
	return e == cudaSuccess ? ERR_OK : b2g_cuda_fail(e, "cudaMemcpyAsync");
}

/* ---- overlay mode: forwarding to the stock libbee2 behind this library ---- */
static volatile u64 g_forwards;
static volatile long g_cpu_below = -1;     /* -1: not read from the environment yet */

void* b2g_stock(const char* name)
{
	/* RTLD_NEXT: the next definition of `name` after THIS library in the search order, i.e. the
	   stock libbee2 the application (also) links or that this library was preloaded in front of;
	   NULL when there is none. A process that dlopen()s this library (Python, a plugin host) has no
	   such order: it names the stock library in B2G_STOCK_LIB and gets a private copy of it. */
	static void* handle;
	static int tried;
	void* f = dlsym(RTLD_NEXT, name);
	if (!f)
	{
		if (!tried)
		{
			const char* path = getenv("B2G_STOCK_LIB");
			pthread_mutex_lock(&g_init_mu);
			if (!tried && path && *path)
				handle = dlopen(path, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
			__sync_synchronize();
			tried = 1;
			pthread_mutex_unlock(&g_init_mu);
		}
		if (handle)
			f = dlsym(handle, name);
	}
	(void)dlerror();
	return f;
}
void b2g_note_forward(const char* fn)
{
	static int trace = -1;
	__sync_fetch_and_add(&g_forwards, 1);
	if (trace < 0)
	{
		const char* e = getenv("B2G_TRACE_FORWARD");
		trace = e && *e && *e != '0';
	}
	if (trace)
		fprintf(stderr, "b2g-forward %s\n", fn);
}
u64 b2g_forward_count(void) { return g_forwards; }
void b2g_warn_forward(const char* fn, u32 code)
{
	static int said;
	b2g_note_forward(fn);
	if (!__sync_lock_test_and_set(&said, 1))
		fprintf(stderr, "bee2_b200: %s cannot run on the GPU path (err %u: %s); forwarding such calls to the stock "
			"libbee2 behind this library (said once)\n", fn, code, g_err);
}
void b2g_set_cpu_below(size_t bytes) { g_cpu_below = (long)bytes; }
int b2g_route_small(size_t bytes)
{
	long t = g_cpu_below;
	if (t < 0)
	{
		const char* e = getenv("B2G_CPU_BELOW");
		t = e && *e ? strtol(e, 0, 0) : 0;
		if (t < 0)
			t = 0;
		g_cpu_below = t;
	}
	return t > 0 && bytes < (size_t)t;
}
int b2g_has_stock(void) { return b2g_stock("bashHash") != 0; }

int b2g_ct_eq(const void* a, const void* b, size_t n)
{
	const volatile octet* x = (const volatile octet*)a;
	const volatile octet* y = (const volatile octet*)b;
	octet d = 0;
	size_t i;
	for (i = 0; i < n; ++i)
		d |= x[i] ^ y[i];
	return d == 0;
}
void b2g_wipe(void* p, size_t n)
{
	volatile octet* v = (volatile octet*)p;
	while (n--)
		*v++ = 0;
}
