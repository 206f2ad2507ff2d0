/*
 * cpu_harness.c — pthread fan-out over a CPU implementation of the hot path.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and --impl reference, and
 * bulk differential tests). It dlopen()s either oracle/_ref/libbee2ref_*.so (the unmodified
 * reference: symbols bashHash, beltCTR, beltECBEncr, bign128Verify, bign128Sign2,
 * bign128PubkeyCalc) or oracle/_ref/libbee2oracle.so (our restatement: orc_* symbols) and
 * runs T threads, each over a disjoint contiguous slice of the same batch. Every entry
 * point returns the wall-clock seconds of the parallel section (CLOCK_MONOTONIC), or a
 * negative number on failure.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef uint32_t (*bash_hash_fn)(uint8_t*, size_t, const void*, size_t);
typedef uint32_t (*belt_ctr_fn)(void*, const void*, size_t, const uint8_t*, size_t, const uint8_t*);
typedef uint32_t (*belt_ecb_fn)(void*, const void*, size_t, const uint8_t*, size_t);
typedef uint32_t (*belt_dwp_fn)(void*, uint8_t*, const void*, size_t, const void*, size_t, const uint8_t*, size_t, const uint8_t*);
typedef uint32_t (*ref_verify_fn)(const uint8_t*, const uint8_t*, const uint8_t*);
typedef uint32_t (*ref_sign2_fn)(uint8_t*, const uint8_t*, const uint8_t*, const void*, size_t);
typedef uint32_t (*ref_pubkey_fn)(uint8_t*, const uint8_t*);
typedef uint32_t (*orc_verify_fn)(const uint8_t*, size_t, const uint8_t*, const uint8_t*, const uint8_t*);
typedef uint32_t (*orc_sign2_fn)(uint8_t*, const uint8_t*, size_t, const uint8_t*, const uint8_t*, const void*, size_t);

static const uint8_t OID[11] = {0x06, 0x09, 0x2A, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x1F, 0x51};

enum { K_BASH, K_CTR, K_ECB_MK, K_VERIFY, K_SIGN2, K_PUBKEY, K_DWP };

typedef struct
{
	int kind, is_port;
	void* fn;
	size_t first, count;       /* units of this thread */
	/* bash */
	const uint8_t* msgs; size_t msg_len, stride, l; uint8_t* out;
	/* belt */
	const uint8_t* key; const uint8_t* iv; uint8_t* dst; const uint8_t* src; size_t unit_bytes;
	const uint8_t* keys;
	/* bign */
	const uint8_t *hashes, *sigs, *pubkeys, *privkeys; uint32_t* status; uint8_t* sig_out; uint8_t* pub_out;
	int failed;
} job_t;

static void* worker(void* arg)
{
	job_t* j = (job_t*)arg;
	size_t i;
	switch (j->kind)
	{
	case K_BASH:
		for (i = j->first; i < j->first + j->count; ++i)
			if (((bash_hash_fn)j->fn)(j->out + i * (j->l / 4), j->l, j->msgs + i * j->stride, j->msg_len))
				j->failed = 1;
		break;
	case K_CTR:
		/* one independent beltCTR call per thread over its slice (iv tweaked per thread) */
		{
			uint8_t iv[16];
			memcpy(iv, j->iv, 16);
			iv[0] ^= (uint8_t)j->first, iv[1] ^= (uint8_t)(j->first >> 8);
			if (j->count && ((belt_ctr_fn)j->fn)(j->dst + j->first * j->unit_bytes,
					j->src ? j->src + j->first * j->unit_bytes : j->dst + j->first * j->unit_bytes,
					j->count * j->unit_bytes, j->key, 32, iv))
				j->failed = 1;
		}
		break;
	case K_DWP:
		/* one independent beltDWPWrap per thread over its slice, 16 octets of open data */
		{
			uint8_t iv[16], mac[8];
			memcpy(iv, j->iv, 16);
			iv[0] ^= (uint8_t)j->first, iv[1] ^= (uint8_t)(j->first >> 8);
			if (j->count && ((belt_dwp_fn)j->fn)(j->dst + j->first * j->unit_bytes, mac,
					j->dst + j->first * j->unit_bytes, j->count * j->unit_bytes, j->iv, 16, j->key, 32, iv))
				j->failed = 1;
		}
		break;
	case K_ECB_MK:
		for (i = j->first; i < j->first + j->count; ++i)
			if (((belt_ecb_fn)j->fn)(j->dst + 16 * i, j->dst + 16 * i, 16, j->keys + 32 * i, 32))
				j->failed = 1;
		break;
	case K_VERIFY:
		for (i = j->first; i < j->first + j->count; ++i)
			j->status[i] = j->is_port ?
				((orc_verify_fn)j->fn)(OID, sizeof OID, j->hashes + 32 * i, j->sigs + 48 * i, j->pubkeys + 64 * i) :
				((ref_verify_fn)j->fn)(j->hashes + 32 * i, j->sigs + 48 * i, j->pubkeys + 64 * i);
		break;
	case K_SIGN2:
		for (i = j->first; i < j->first + j->count; ++i)
			j->status[i] = j->is_port ?
				((orc_sign2_fn)j->fn)(j->sig_out + 48 * i, OID, sizeof OID, j->hashes + 32 * i, j->privkeys + 32 * i, 0, 0) :
				((ref_sign2_fn)j->fn)(j->sig_out + 48 * i, j->hashes + 32 * i, j->privkeys + 32 * i, 0, 0);
		break;
	case K_PUBKEY:
		for (i = j->first; i < j->first + j->count; ++i)
			j->status[i] = ((ref_pubkey_fn)j->fn)(j->pub_out + 64 * i, j->privkeys + 32 * i);
		break;
	}
	return 0;
}

static double now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double fan_out(job_t proto, size_t units, int threads)
{
	pthread_t* th;
	job_t* jobs;
	double t0, t1;
	int t, failed = 0;
	if (threads < 1) threads = 1;
	if ((size_t)threads > units && units) threads = (int)units;
	th = (pthread_t*)calloc((size_t)threads, sizeof *th);
	jobs = (job_t*)calloc((size_t)threads, sizeof *jobs);
	if (!th || !jobs) return -1;
	for (t = 0; t < threads; ++t)
	{
		jobs[t] = proto;
		jobs[t].first = units * (size_t)t / (size_t)threads;
		jobs[t].count = units * (size_t)(t + 1) / (size_t)threads - jobs[t].first;
	}
	t0 = now();
	for (t = 0; t < threads; ++t)
		pthread_create(&th[t], 0, worker, &jobs[t]);
	for (t = 0; t < threads; ++t)
		pthread_join(th[t], 0), failed |= jobs[t].failed;
	t1 = now();
	free(th), free(jobs);
	return failed ? -2 : t1 - t0;
}

static void* sym(const char* libpath, const char* name)
{
	void* h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
	return h ? dlsym(h, name) : 0;
}

double harness_bash(const char* libpath, int is_port, uint8_t* out, size_t l, const uint8_t* msgs,
	size_t msg_len, size_t stride, size_t count, int threads)
{
	job_t j;
	memset(&j, 0, sizeof j);
	j.kind = K_BASH, j.is_port = is_port;
	if (!(j.fn = sym(libpath, is_port ? "orc_bashHash" : "bashHash"))) return -3;
	j.out = out, j.l = l, j.msgs = msgs, j.msg_len = msg_len, j.stride = stride;
	return fan_out(j, count, threads);
}

/* dst[0..units*unit_bytes) <- src ^ keystream; src may be NULL (in place over dst) */
double harness_belt_ctr(const char* libpath, int is_port, uint8_t* dst, const uint8_t* src, size_t unit_bytes,
	size_t units, const uint8_t key[32], const uint8_t iv[16], int threads)
{
	job_t j;
	memset(&j, 0, sizeof j);
	j.kind = K_CTR, j.is_port = is_port;
	if (!(j.fn = sym(libpath, is_port ? "orc_beltCTR" : "beltCTR"))) return -3;
	j.dst = dst, j.src = src, j.unit_bytes = unit_bytes, j.key = key, j.iv = iv;
	return fan_out(j, units, threads);
}

double harness_belt_dwp(const char* libpath, int is_port, uint8_t* buf, size_t unit_bytes, size_t units,
	const uint8_t key[32], const uint8_t iv[16], int threads)
{
	job_t j;
	memset(&j, 0, sizeof j);
	j.kind = K_DWP, j.is_port = is_port;
	if (!(j.fn = sym(libpath, is_port ? "orc_beltDWPWrap" : "beltDWPWrap"))) return -3;
	j.dst = buf, j.unit_bytes = unit_bytes, j.key = key, j.iv = iv;
	return fan_out(j, units, threads);
}

double harness_belt_ecb_multikey(const char* libpath, int is_port, uint8_t* blocks, const uint8_t* keys32,
	size_t count, int threads)
{
	job_t j;
	memset(&j, 0, sizeof j);
	j.kind = K_ECB_MK, j.is_port = is_port;
	if (!(j.fn = sym(libpath, is_port ? "orc_beltECBEncr" : "beltECBEncr"))) return -3;
	j.dst = blocks, j.keys = keys32;
	return fan_out(j, count, threads);
}

double harness_bign_verify(const char* libpath, int is_port, uint32_t* status, const uint8_t* hashes,
	const uint8_t* sigs, const uint8_t* pubkeys, size_t count, int threads)
{
	job_t j;
	memset(&j, 0, sizeof j);
	j.kind = K_VERIFY, j.is_port = is_port;
	if (!(j.fn = sym(libpath, is_port ? "orc_bignVerify128" : "bign128Verify"))) return -3;
	if (!is_port && count)   /* build the lazily-created global curve before the threads race for it */
		(void)((ref_verify_fn)j.fn)(hashes, sigs, pubkeys);
	j.status = status, j.hashes = hashes, j.sigs = sigs, j.pubkeys = pubkeys;
	return fan_out(j, count, threads);
}

double harness_bign_sign2(const char* libpath, int is_port, uint32_t* status, uint8_t* sigs,
	const uint8_t* hashes, const uint8_t* privkeys, size_t count, int threads)
{
	job_t j;
	memset(&j, 0, sizeof j);
	j.kind = K_SIGN2, j.is_port = is_port;
	if (!(j.fn = sym(libpath, is_port ? "orc_bignSign2_128" : "bign128Sign2"))) return -3;
	if (!is_port && count)
	{
		uint8_t tmp[48];
		(void)((ref_sign2_fn)j.fn)(tmp, hashes, privkeys, 0, 0);
	}
	j.status = status, j.sig_out = sigs, j.hashes = hashes, j.privkeys = privkeys;
	return fan_out(j, count, threads);
}

double harness_bign_pubkey(const char* libpath, int is_port, uint32_t* status, uint8_t* pubkeys,
	const uint8_t* privkeys, size_t count, int threads)
{
	job_t j;
	memset(&j, 0, sizeof j);
	j.kind = K_PUBKEY, j.is_port = is_port;
	if (!(j.fn = sym(libpath, is_port ? "orc_bignPubkeyCalc128" : "bign128PubkeyCalc"))) return -3;
	if (!is_port && count)
	{
		uint8_t tmp[64];
		(void)((ref_pubkey_fn)j.fn)(tmp, privkeys);
	}
	j.status = status, j.pub_out = pubkeys, j.privkeys = privkeys;
	return fan_out(j, count, threads);
}
