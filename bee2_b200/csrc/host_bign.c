/*
 * host_bign.c — host side (C) of the bign path: the reference's bign.h surface for
 * bignParamsStd / bignVerify / bignSign2 / bignPubkeyCalc (bign_params.c:197-280,
 * bign_sign.c:247-260, :349-361, bign_misc.c:369-412) and the host-pointer batch entry
 * points. Argument checks (parameter block, OID DER syntax) mirror the reference; all
 * field, curve and hash arithmetic runs in bign.cu on the device.
 */
#include "engine.h"
#include <string.h>
#include <stdlib.h>
#include <stdarg.h>

#define CU(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { code = b2g_cuda_fail(e_, what); goto done; } } while (0)

/* The three standard curves, tables B.1-B.3 of STB 34.101.45 (bign_params.c:33-190):
   bign-curve256v1 / 384v1 / 512v1 = levels l = 128 / 192 / 256; p = 2^(2l) - c, a = p - 3 */
static const octet curve256v1_b[32] = {
	0xF1, 0x03, 0x9C, 0xD6, 0x6B, 0x7D, 0x2E, 0xB2, 0x53, 0x92, 0x8B, 0x97, 0x69, 0x50, 0xF5, 0x4C,
	0xBE, 0xFB, 0xD8, 0xE4, 0xAB, 0x3A, 0xC1, 0xD2, 0xED, 0xA8, 0xF3, 0x15, 0x15, 0x6C, 0xCE, 0x77};
static const octet curve256v1_q[32] = {
	0x07, 0x66, 0x3D, 0x26, 0x99, 0xBF, 0x5A, 0x7E, 0xFC, 0x4D, 0xFB, 0x0D, 0xD6, 0x8E, 0x5C, 0xD9,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
static const octet curve256v1_yG[32] = {
	0x93, 0x6A, 0x51, 0x04, 0x18, 0xCF, 0x29, 0x1E, 0x52, 0xF6, 0x08, 0xC4, 0x66, 0x39, 0x91, 0x78,
	0x5D, 0x83, 0xD6, 0x51, 0xA3, 0xC9, 0xE4, 0x5C, 0x9F, 0xD6, 0x16, 0xFB, 0x3C, 0xFC, 0xF7, 0x6B};
static const octet curve256v1_seed[8] = {0x5E, 0x38, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00};
static const octet curve384v1_b[48] = {
	0x64, 0xBF, 0x73, 0x68, 0x23, 0xFC, 0xA7, 0xBC, 0x7C, 0xBD, 0xCE, 0xF3, 0xF0, 0xE2, 0xBD, 0x14,
	0x3A, 0x2E, 0x71, 0xE9, 0xF9, 0x6A, 0x21, 0xA6, 0x96, 0xB1, 0xFB, 0x0F, 0xBB, 0x48, 0x27, 0x71,
	0xD2, 0x34, 0x5D, 0x65, 0xAB, 0x5A, 0x07, 0x33, 0x20, 0xEF, 0x9C, 0x95, 0xE1, 0xDF, 0x75, 0x3C};
static const octet curve384v1_q[48] = {
	0xB7, 0xA7, 0x0C, 0xF3, 0x3F, 0xDC, 0xB7, 0x3D, 0x0A, 0xFF, 0xA4, 0xA6, 0xE7, 0xDA, 0x46, 0x80,
	0xBB, 0x7B, 0xAF, 0x73, 0x03, 0xC4, 0xCC, 0x6C, 0xFE, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
static const octet curve384v1_yG[48] = {
	0x51, 0xC4, 0x33, 0xF7, 0x31, 0xCB, 0x5E, 0xEA, 0xF9, 0x42, 0x2A, 0x6B, 0x27, 0x3E, 0x40, 0x84,
	0x55, 0xD3, 0xB1, 0x66, 0x9E, 0xE7, 0x49, 0x05, 0xA0, 0xFF, 0x86, 0xDC, 0x11, 0x9A, 0x72, 0x3A,
	0x89, 0xBF, 0x2D, 0x43, 0x7E, 0x11, 0x30, 0x63, 0x9E, 0x9E, 0x2E, 0xA8, 0x24, 0x82, 0x43, 0x5D};
static const octet curve384v1_seed[8] = {0x23, 0xAF, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00};
static const octet curve512v1_b[64] = {
	0x90, 0x9C, 0x13, 0xD6, 0x98, 0x69, 0x34, 0x09, 0x7A, 0xA2, 0x49, 0x3A, 0x27, 0x22, 0x86, 0xEA,
	0x43, 0xA2, 0xAC, 0x87, 0x8C, 0x00, 0x33, 0x29, 0x95, 0x5E, 0x24, 0xC4, 0xB5, 0xDC, 0x11, 0x27,
	0x88, 0xB0, 0xAD, 0xDA, 0xE3, 0x13, 0xCE, 0x17, 0x51, 0x25, 0x5D, 0xDD, 0xEE, 0xA9, 0xC6, 0x5B,
	0x89, 0x58, 0xFD, 0x60, 0x6A, 0x5D, 0x8C, 0xD8, 0x43, 0x8C, 0x3B, 0x93, 0x44, 0x59, 0xB4, 0x6C};
static const octet curve512v1_q[64] = {
	0xF1, 0x8E, 0x06, 0x0D, 0x49, 0xAD, 0xFF, 0xDC, 0x32, 0xDF, 0x56, 0x95, 0xE5, 0xCA, 0x1B, 0x36,
	0xF4, 0x13, 0x21, 0x2E, 0xB0, 0xEB, 0x6B, 0xF2, 0x4E, 0x00, 0x98, 0x01, 0x2C, 0x09, 0xC0, 0xB2,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
static const octet curve512v1_yG[64] = {
	0xBD, 0xED, 0xEF, 0xCE, 0x6F, 0xAE, 0x92, 0xB7, 0x04, 0x0D, 0x4C, 0xC9, 0xB9, 0x83, 0xAA, 0x67,
	0x61, 0x22, 0xE8, 0xEE, 0x95, 0x73, 0x77, 0xFF, 0xD2, 0x6F, 0xFA, 0x0E, 0xE2, 0xDD, 0x73, 0x69,
	0xDA, 0xCA, 0xCC, 0x00, 0x1B, 0xF8, 0xED, 0xD2, 0xE2, 0xBC, 0x61, 0xB3, 0xB3, 0x41, 0xAB, 0xB0,
	0xAB, 0x8F, 0xD1, 0xA0, 0xF7, 0xE6, 0x82, 0xB1, 0x81, 0x76, 0x03, 0xE4, 0x7A, 0xFF, 0x26, 0xA8};
static const octet curve512v1_seed[8] = {0xAE, 0x17, 0x02, 0x00, 0x00, 0x00, 0x00, 0x00};

typedef struct
{
	const char* name;
	size_t l;
	octet p0, p1;          /* the two low octets of p (all others are 0xFF) */
	const octet *b, *q, *yG, *seed;
} std_curve;
static const std_curve std_curves[3] = {
	{"1.2.112.0.2.0.34.101.45.3.1", 128, 0x43, 0xFF, curve256v1_b, curve256v1_q, curve256v1_yG, curve256v1_seed},
	{"1.2.112.0.2.0.34.101.45.3.2", 192, 0xC3, 0xFE, curve384v1_b, curve384v1_q, curve384v1_yG, curve384v1_seed},
	{"1.2.112.0.2.0.34.101.45.3.3", 256, 0xC7, 0xFD, curve512v1_b, curve512v1_q, curve512v1_yG, curve512v1_seed},
};

static void std_fill(bign_params* params, const std_curve* c)
{
	const size_t no = c->l / 4;
	memset(params, 0, sizeof *params);
	params->l = c->l;
	memset(params->p, 0xFF, no), params->p[0] = c->p0, params->p[1] = c->p1;
	memset(params->a, 0xFF, no), params->a[0] = (octet)(c->p0 - 3), params->a[1] = c->p1;
	memcpy(params->b, c->b, no);
	memcpy(params->q, c->q, no);
	memcpy(params->yG, c->yG, no);
	memcpy(params->seed, c->seed, 8);
}

err_t bignParamsStd(bign_params* params, const char* name)
{
	size_t i;
	if (!params)
		return ERR_BAD_INPUT;
	memset(params, 0, sizeof *params);
	for (i = 0; name && i < 3; ++i)
		if (strcmp(name, std_curves[i].name) == 0)
		{
			std_fill(params, &std_curves[i]);
			return ERR_OK;
		}
	return ERR_FILE_NOT_FOUND;
}

static int is_zero(const octet* p, size_t n)
{
	octet acc = 0;
	while (n--) acc |= *p++;
	return acc == 0;
}

/* bignParamsCheck (bign_params.c:244-280), then: is it the curve this engine implements? */
static err_t params_check(const bign_params* params)
{
	bign_params std;
	size_t no;
	if (!params)
		return ERR_BAD_INPUT;
	if (2 * params->l % 64)
		return ERR_NOT_IMPLEMENTED;
	no = (2 * params->l + 7) / 8;
	if (no == 0 || no > 64)
		return ERR_BAD_PARAMS;
	if (!(params->p[0] % 4 == 3 && params->q[0] % 2 == 1 &&
		params->p[no - 1] >= 128 && params->q[no - 1] >= 128 &&
		is_zero(params->p + no, 64 - no) && !is_zero(params->a, no) && !is_zero(params->b, no) &&
		is_zero(params->a + no, 64 - no) && is_zero(params->b + no, 64 - no) &&
		is_zero(params->q + no, 64 - no) && is_zero(params->yG + no, 64 - no)))
		return ERR_BAD_PARAMS;
	if (params->l % 64)
		return ERR_NOT_IMPLEMENTED;
	if (params->l != 128 && params->l != 192 && params->l != 256)
		return ERR_BAD_PARAMS;
	/* GPU path: the three standard curves (seed is not used by sign/verify) */
	{
		size_t i;
		for (i = 0; i < 3; ++i)
			if (params->l == std_curves[i].l)
			{
				std_fill(&std, &std_curves[i]);
				if (memcmp(params->p, std.p, 64) || memcmp(params->a, std.a, 64) || memcmp(params->b, std.b, 64) ||
					memcmp(params->q, std.q, 64) || memcmp(params->yG, std.yG, 64))
					return ERR_NOT_IMPLEMENTED;
			}
	}
	return ERR_OK;
}

/* Overlay routing of the single-item drop-ins: a parameter block that is structurally valid but not one
   of the three standard curves (the reference accepts any, bign_params.c:197-280), and — when the
   caller asked for it (b2g_set_cpu_below) — every single-item call, go to the stock libbee2 behind us. */
#define BIGN_ROUTE(name, ...) do { \
	if (params && (params_check(params) == ERR_NOT_IMPLEMENTED || b2g_route_small(1))) \
		B2G_STOCK_R(name, __VA_ARGS__); } while (0)

/* DER OBJECT IDENTIFIER syntax as accepted by oidFromDER(0, der, len) != SIZE_MAX
   (oid.c:94-101, der.c:114-151, :193-232, :921-975) */
static int oid_der_valid(const octet* der, size_t count)
{
	size_t l, hdr, pos;
	u32 val = 0;
	if (count == (size_t)-1 || !der || count < 2 || der[0] != 0x06)
		return 0;
	if (der[1] == 128 || der[1] == 255)
		return 0;
	if (der[1] < 128)
		l = der[1], hdr = 2;
	else
	{
		size_t r = der[1] - 128, i;
		hdr = 2 + r;
		if (count < hdr || r > sizeof(size_t) || der[2] == 0 || (r == 1 && der[2] < 128))
			return 0;
		for (l = 0, i = 0; i < r; ++i)
			l = l << 8 | der[2 + i];
		if (l == (size_t)-1)
			return 0;
	}
	if (hdr + l != count)
		return 0;
	for (pos = 0; pos < l; ++pos)
	{
		const octet o = der[hdr + pos];
		if (val & 0xFE000000u)
			return 0;
		if (val == 0 && o == 128)
			return 0;
		val = val << 7 | (o & 127u);
		if ((o & 128) == 0)
			val = 0;
	}
	return 1;
}

/* upload `n` octets per item for `count` items into slot buffer `which` */
static err_t stage_in(b2g_slot* sl, int which, const void* host, size_t bytes, void** dev)
{
	err_t code;
	cudaError_t e;
	if ((code = b2g_slot_buf(sl, which, bytes, dev)))
		return code;
	if ((e = cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, sl->stream)) != cudaSuccess)
		return b2g_cuda_fail(e, "H2D(bign)");
	return ERR_OK;
}

/* Device-visible address of a PINNED host buffer (b2g_host_alloc, cudaHostAlloc, cudaHostRegister), or NULL
   for pageable memory. The bign kernels touch every input octet exactly once (144 B read, 4 B written per
   verification against ~5 ms of arithmetic per 2^18 items, i.e. 7.5 GB/s), so with pinned buffers they read
   and write host memory directly over PCIe — no staging copy, nothing to overlap, nothing left on the device. */
static void* pinned_dev_ptr(const void* host)
{
	struct cudaPointerAttributes at;
	if (!host || cudaPointerGetAttributes(&at, host) != cudaSuccess)
	{
		(void)cudaGetLastError();
		return 0;
	}
	return at.type == cudaMemoryTypeHost ? at.devicePointer : 0;
}

/* items per pipeline stage of bignVerifyBatch; B2G_BIGN_CHUNK overrides it (tuning knob) */
static size_t verify_chunk(void)
{
	const char* e = getenv("B2G_BIGN_CHUNK");
	if (e && *e)
	{
		const unsigned long v = strtoul(e, 0, 10);
		if (v >= 128)
			return (size_t)v;
	}
	/* 2^16 items per stage: the upload of stage i+1 and the status download of stage i-1 run under the
	   kernel of stage i. (Round 1 used 2^18 — no overlap at the config size — because smaller grids of
	   256-thread CTAs lost more to wave quantisation, 35 M/s at 2^16; the launcher now sizes the CTAs of a
	   small grid so that every SM gets the same load, bign.cu bign_threads.) */
	return (size_t)1 << 16;
}

static err_t verify_batch_1(err_t* status, const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet* hashes, const octet* sigs, const octet* pubkeys, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_h, *d_s, *d_p, *d_st;
	size_t no;
	if ((code = params_check(params)))
		return code;
	no = params->l / 4;
	if (count && (!status || !hashes || !sigs || !pubkeys))
		return ERR_BAD_INPUT;
	if (!oid_der_valid(oid_der, oid_len))
		return ERR_BAD_OID;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	{
		/* all four buffers pinned: one launch straight on host memory (zero-copy) */
		void *z_h = pinned_dev_ptr(hashes), *z_s = pinned_dev_ptr(sigs), *z_p = pinned_dev_ptr(pubkeys),
			*z_st = pinned_dev_ptr(status);
		if (z_h && z_s && z_p && z_st && !getenv("B2G_NO_ZEROCOPY"))
		{
			if ((code = b2g_bignVerifyBatchL_dev(params->l, z_st, oid_der, oid_len, z_h, z_s, z_p, count, s0->stream)))
				goto done;
			CU(cudaStreamSynchronize(s0->stream), "sync(bign verify, zero-copy)");
			goto done;
		}
	}
	/* chunks of verify_chunk() items alternate between the two workspace slots, so the H2D copy of
	   one chunk overlaps the kernel of the previous one */
	{
		const size_t chunk = verify_chunk();
		size_t off, c;
		for (off = 0, c = 0; off < count; off += chunk, ++c)
		{
			const size_t n = count - off < chunk ? count - off : chunk;
			b2g_slot* sl = b2g_slot_get((int)c);
			if ((code = stage_in(sl, 0, hashes + no * off, no * n, &d_h)) ||
				(code = stage_in(sl, 1, sigs + (no + no / 2) * off, (no + no / 2) * n, &d_s)) ||
				(code = stage_in(sl, 2, pubkeys + 2 * no * off, 2 * no * n, &d_p)) ||
				(code = b2g_slot_buf(sl, 3, 4 * n, &d_st)))
				goto done;
			if ((code = b2g_bignVerifyBatchL_dev(params->l, d_st, oid_der, oid_len, d_h, d_s, d_p, n, sl->stream)))
				goto done;
			CU(cudaMemcpyAsync(status + off, d_st, 4 * n, cudaMemcpyDeviceToHost, sl->stream), "D2H(bign status)");
		}
		CU(cudaStreamSynchronize(s0->stream), "sync(bign verify)");
		CU(cudaStreamSynchronize(s1->stream), "sync(bign verify)");
	}
done:
	if (code)
		cudaStreamSynchronize(s0->stream), cudaStreamSynchronize(s1->stream);
	b2g_unlock();
	return code;
}

/* in-process multi-device mode: contiguous shares of the batch, one device each */
typedef struct
{
	err_t* status;
	const bign_params* params;
	const octet* oid_der;
	size_t oid_len;
	const octet *hashes, *sigs, *pubkeys;
} verify_args;
static u32 verify_shard(void* arg, size_t first, size_t n)
{
	const verify_args* a = (const verify_args*)arg;
	const size_t no = a->params->l / 4;
	return verify_batch_1(a->status + first, a->params, a->oid_der, a->oid_len, a->hashes + no * first,
		a->sigs + (no + no / 2) * first, a->pubkeys + 2 * no * first, n);
}
err_t bignVerifyBatch(err_t* status, const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet* hashes, const octet* sigs, const octet* pubkeys, size_t count)
{
	verify_args a = {status, params, oid_der, oid_len, hashes, sigs, pubkeys};
	err_t code;
	if (b2g_device_count() <= 1 || count < 2 * 4096)
		return verify_batch_1(status, params, oid_der, oid_len, hashes, sigs, pubkeys, count);
	if ((code = params_check(params)))
		return code;
	if (!status || !hashes || !sigs || !pubkeys)
		return ERR_BAD_INPUT;
	return b2g_fanout(count, 4096, verify_shard, &a);
}

err_t bignVerify(const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet sig[], const octet pubkey[])
{
	BIGN_ROUTE(bignVerify, params, oid_der, oid_len, hash, sig, pubkey);
	err_t st = ERR_BAD_SIG, code;
	/* order of checks: params (bign_sign.c:355-356), buffers, OID (:286-290) */
	if ((code = params_check(params)))
		return code;
	if (!hash || !sig || !pubkey)
		return ERR_BAD_INPUT;
	if ((code = bignVerifyBatch(&st, params, oid_der, oid_len, hash, sig, pubkey, 1)))
	{
		if (code == ERR_NOT_IMPLEMENTED)   /* an OID longer than 64 octets */
			B2G_STOCK_R(bignVerify, params, oid_der, oid_len, hash, sig, pubkey);
		return code;
	}
	return st;
}

static err_t sign2_batch(err_t* status, octet* sigs, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* privkeys, size_t count, const void* t, size_t t_len)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_h, *d_k, *d_sig, *d_st;
	size_t no, so;
	if ((code = params_check(params)))
		return code;
	no = params->l / 4, so = no + no / 2;
	if (count && (!status || !sigs || !hashes || !privkeys))
		return ERR_BAD_INPUT;
	if (!oid_der_valid(oid_der, oid_len))
		return ERR_BAD_OID;
	if (t_len && !t)
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	{
		/* all four buffers pinned: one launch straight on host memory (zero-copy; the kernel stages its inputs
		   and its signatures with TMA bulk copies and writes a signature only where the item signed) — the
		   private keys then never rest in device memory */
		void *z_h = pinned_dev_ptr(hashes), *z_k = pinned_dev_ptr(privkeys), *z_sig = pinned_dev_ptr(sigs),
			*z_st = pinned_dev_ptr(status);
		if (z_h && z_k && z_sig && z_st && !getenv("B2G_NO_ZEROCOPY"))
		{
			if ((code = b2g_bignSign2BatchL_t_dev(params->l, z_st, z_sig, oid_der, oid_len, z_h, z_k, count, t, t_len, s0->stream)))
				goto done;
			CU(cudaStreamSynchronize(s0->stream), "sync(bign sign2, zero-copy)");
			goto done;
		}
	}
	if ((code = stage_in(s0, 0, hashes, no * count, &d_h)) || (code = stage_in(s0, 1, privkeys, no * count, &d_k)) ||
		(code = b2g_slot_buf(s0, 2, so * count, &d_sig)) || (code = b2g_slot_buf(s1, 0, 4 * count, &d_st)))
		goto done;
	if ((code = b2g_bignSign2BatchL_t_dev(params->l, d_st, d_sig, oid_der, oid_len, d_h, d_k, count, t, t_len, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(status, d_st, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign status)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign sign2)");
	{
		/* only items that signed successfully overwrite the caller's sig buffer */
		size_t i, all_ok = 1;
		for (i = 0; i < count; ++i)
			all_ok &= status[i] == ERR_OK;
		if (all_ok)
			CU(cudaMemcpy(sigs, d_sig, so * count, cudaMemcpyDeviceToHost), "D2H(bign sigs)");
		else
			for (i = 0; i < count; ++i)
				if (status[i] == ERR_OK)
					CU(cudaMemcpy(sigs + so * i, (octet*)d_sig + so * i, so, cudaMemcpyDeviceToHost), "D2H(bign sig)");
	}
	/* the private keys were staged on the device: wipe them */
	CU(cudaMemsetAsync(d_k, 0, no * count, s0->stream), "memset(bign keys)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign sign2)");
done:
	if (code)
		b2g_slot_wipe(s0);   /* failure path: nothing staged (keys, nonces, shared points) stays behind */
	b2g_unlock();
	return code;
}

typedef struct
{
	err_t* status;
	octet* sigs;
	const bign_params* params;
	const octet* oid_der;
	size_t oid_len;
	const octet *hashes, *privkeys;
} sign2_args;
static u32 sign2_shard(void* arg, size_t first, size_t n)
{
	const sign2_args* a = (const sign2_args*)arg;
	const size_t no = a->params->l / 4;
	return sign2_batch(a->status + first, a->sigs + (no + no / 2) * first, a->params, a->oid_der, a->oid_len,
		a->hashes + no * first, a->privkeys + no * first, n, 0, 0);
}
err_t bignSign2Batch(err_t* status, octet* sigs, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* privkeys, size_t count)
{
	sign2_args a = {status, sigs, params, oid_der, oid_len, hashes, privkeys};
	err_t code;
	if (b2g_device_count() <= 1 || count < 2 * 4096)
		return sign2_batch(status, sigs, params, oid_der, oid_len, hashes, privkeys, count, 0, 0);
	if ((code = params_check(params)))
		return code;
	if (!status || !sigs || !hashes || !privkeys)
		return ERR_BAD_INPUT;
	return b2g_fanout(count, 4096, sign2_shard, &a);
}

err_t bignSign2(octet sig[], const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet privkey[], const void* t, size_t t_len)
{
	BIGN_ROUTE(bignSign2, sig, params, oid_der, oid_len, hash, privkey, t, t_len);
	err_t st = ERR_BAD_INPUT, code;
	if ((code = params_check(params)))
		return code;
	if (!hash || !privkey || !sig)
		return ERR_BAD_INPUT;
	if ((code = sign2_batch(&st, sig, params, oid_der, oid_len, hash, privkey, 1, t, t ? t_len : 0)))
	{
		if (code == ERR_NOT_IMPLEMENTED)   /* OID or t longer than 64 octets (bign_sign.c:198-206 takes any) */
			B2G_STOCK_R(bignSign2, sig, params, oid_der, oid_len, hash, privkey, t, t_len);
		return code;
	}
	return st;
}

static err_t pubkey_batch_1(err_t* status, octet* pubkeys, const bign_params* params,
	const octet* privkeys, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_k, *d_p, *d_st;
	size_t no;
	if ((code = params_check(params)))
		return code;
	no = params->l / 4;
	if (count && (!status || !pubkeys || !privkeys))
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, privkeys, no * count, &d_k)) || (code = b2g_slot_buf(s0, 1, 2 * no * count, &d_p)) ||
		(code = b2g_slot_buf(s1, 0, 4 * count, &d_st)))
		goto done;
	CU(cudaMemsetAsync(d_p, 0, 2 * no * count, s0->stream), "memset(bign pubkeys)");
	if ((code = b2g_bignPubkeyCalcBatchL_dev(params->l, d_st, d_p, d_k, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(status, d_st, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign status)");
	CU(cudaMemcpyAsync(pubkeys, d_p, 2 * no * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign pubkeys)");
	CU(cudaMemsetAsync(d_k, 0, no * count, s0->stream), "memset(bign keys)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign pubkey)");
done:
	if (code)
		b2g_slot_wipe(s0);   /* failure path: nothing staged (keys, nonces, shared points) stays behind */
	b2g_unlock();
	return code;
}

typedef struct
{
	err_t* status;
	octet* pubkeys;
	const bign_params* params;
	const octet* privkeys;
} pubkey_args;
static u32 pubkey_shard(void* arg, size_t first, size_t n)
{
	const pubkey_args* a = (const pubkey_args*)arg;
	const size_t no = a->params->l / 4;
	return pubkey_batch_1(a->status + first, a->pubkeys + 2 * no * first, a->params, a->privkeys + no * first, n);
}
err_t bignPubkeyCalcBatch(err_t* status, octet* pubkeys, const bign_params* params,
	const octet* privkeys, size_t count)
{
	pubkey_args a = {status, pubkeys, params, privkeys};
	err_t code;
	if (b2g_device_count() <= 1 || count < 2 * 4096)
		return pubkey_batch_1(status, pubkeys, params, privkeys, count);
	if ((code = params_check(params)))
		return code;
	if (!status || !pubkeys || !privkeys)
		return ERR_BAD_INPUT;
	return b2g_fanout(count, 4096, pubkey_shard, &a);
}

err_t bignPubkeyCalc(octet pubkey[], const bign_params* params, const octet privkey[])
{
	BIGN_ROUTE(bignPubkeyCalc, pubkey, params, privkey);
	err_t st = ERR_BAD_INPUT, code;
	octet out[128];
	if ((code = params_check(params)))
		return code;
	if (!pubkey || !privkey)
		return ERR_BAD_INPUT;
	if ((code = bignPubkeyCalcBatch(&st, out, params, privkey, 1)))
		return code;
	if (st == ERR_OK)
		memcpy(pubkey, out, params->l / 2);
	return st;
}

err_t ecMulABatchL(size_t l, octet* b, int* ok, const octet* a, const octet* d, size_t d_len, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_a, *d_d, *d_b, *d_ok;
	const size_t po = l / 2;   /* octets per affine point */
	if (l != 128 && l != 192 && l != 256)
		return ERR_NOT_IMPLEMENTED;
	if (count && (!b || !ok || !a || !d))
		return ERR_BAD_INPUT;
	if (d_len == 0 || d_len > l / 4)
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, a, po * count, &d_a)) || (code = stage_in(s0, 1, d, d_len * count, &d_d)) ||
		(code = b2g_slot_buf(s0, 2, po * count, &d_b)) || (code = b2g_slot_buf(s1, 0, 4 * count, &d_ok)))
		goto done;
	CU(cudaMemsetAsync(d_b, 0, po * count, s0->stream), "memset(ecMulA out)");
	if ((code = b2g_ecMulABatchL_dev(l, d_b, d_ok, d_a, d_d, d_len, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(b, d_b, po * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecMulA)");
	CU(cudaMemcpyAsync(ok, d_ok, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecMulA ok)");
	CU(cudaStreamSynchronize(s0->stream), "sync(ecMulA)");
done:
	if (code)
		b2g_slot_wipe(s0);   /* failure path: nothing staged (keys, nonces, shared points) stays behind */
	b2g_unlock();
	return code;
}

err_t ecMulABatch(octet* b, int* ok, const octet* a, const octet* d, size_t d_len, size_t count)
{
	return ecMulABatchL(128, b, ok, a, d, d_len, count);
}

/* b_i = d_i * a_i + k_i * G (ecAddMulA with the base point as second term, ec.c:1183-1273) */
err_t ecAddMulABatchL(size_t l, octet* b, int* ok, const octet* a, const octet* d, size_t d_len, const octet* k,
	size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_a, *d_d, *d_k, *d_b, *d_ok;
	const size_t po = l / 2;
	if (l != 128 && l != 192 && l != 256)
		return ERR_NOT_IMPLEMENTED;
	if (count && (!b || !ok || !a || !d || !k))
		return ERR_BAD_INPUT;
	if (d_len == 0 || d_len > l / 4)
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, a, po * count, &d_a)) || (code = stage_in(s0, 1, d, d_len * count, &d_d)) ||
		(code = stage_in(s0, 2, k, po / 2 * count, &d_k)) || (code = b2g_slot_buf(s1, 0, po * count, &d_b)) ||
		(code = b2g_slot_buf(s1, 1, 4 * count, &d_ok)))
		goto done;
	CU(cudaMemsetAsync(d_b, 0, po * count, s0->stream), "memset(ecAddMulA out)");
	if ((code = b2g_ecAddMulABatchL_dev(l, d_b, d_ok, d_a, d_d, d_len, d_k, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(b, d_b, po * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecAddMulA)");
	CU(cudaMemcpyAsync(ok, d_ok, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecAddMulA ok)");
	CU(cudaStreamSynchronize(s0->stream), "sync(ecAddMulA)");
done:
	if (code)
		b2g_slot_wipe(s0);   /* failure path: nothing staged (keys, nonces, shared points) stays behind */
	b2g_unlock();
	return code;
}

err_t ecAddMulABatch(octet* b, int* ok, const octet* a, const octet* d, size_t d_len, const octet* k,
	size_t count)
{
	return ecAddMulABatchL(128, b, ok, a, d, d_len, k, count);
}

/* ---------------------------------------------------------------- key generation / validation / DH
   (bign_misc.c:182-300, :317-352, :437-515) */

/* d <-R {1, ..., p - 1}: zzRandNZMod over the FIELD modulus exactly as bignKeypairGenEc calls it
   (bign_misc.c:213 hands ec->f->mod, zz_mod.c:463-484): no octets from rng per attempt, at most
   B_PER_IMPOSSIBLE + 1 attempts. For d in [q, p) (probability ~2^-l) the engine then reports
   ERR_BAD_PRIVKEY where the reference would go on with d mod q. */
/* a <-R {1, ..., mod - 1} (zzRandNZMod, zz_mod.c:463-484; mod has 8 no bits here) */
static int rand_nz_mod(octet d[], const octet mod[], size_t no, gen_i rng, void* rng_state)
{
	size_t tries = 64 + 1, i;
	while (tries--)
	{
		int zero = 1, less = 0;
		rng(d, no, rng_state);
		for (i = 0; i < no; ++i)
			zero &= d[i] == 0;
		for (i = no; i-- > 0;)
			if (d[i] != mod[i])
			{
				less = d[i] < mod[i];
				break;
			}
		if (!zero && less)
			return 1;
	}
	return 0;
}

err_t bignKeypairGenBatch(octet* privkeys, octet* pubkeys, const bign_params* params, gen_i rng,
	void* rng_state, size_t count)
{
	err_t code, *st;
	size_t no, i;
	if ((code = params_check(params)))
		return code;
	if (count && (!privkeys || !pubkeys))
		return ERR_BAD_INPUT;
	if (!rng)
		return ERR_BAD_RNG;
	no = params->l / 4;
	/* the private keys are drawn in item order, like `count` successive bignKeypairGen calls */
	for (i = 0; i < count; ++i)
		if (!rand_nz_mod(privkeys + no * i, params->p, no, rng, rng_state))
			return ERR_BAD_RNG;
	if (!count)
		return ERR_OK;
	if (!(st = (err_t*)malloc(count * sizeof *st)))
		return ERR_OUTOFMEMORY;
	code = bignPubkeyCalcBatch(st, pubkeys, params, privkeys, count);
	for (i = 0; !code && i < count; ++i)
		code = st[i];
	free(st);
	return code;
}

err_t bignKeypairGen(octet privkey[], octet pubkey[], const bign_params* params, gen_i rng, void* rng_state)
{
	BIGN_ROUTE(bignKeypairGen, privkey, pubkey, params, rng, rng_state);
	err_t code;
	if ((code = params_check(params)))
		return code;
	if (!privkey || !pubkey)
		return ERR_BAD_INPUT;
	return bignKeypairGenBatch(privkey, pubkey, params, rng, rng_state, 1);
}

/* status[i] = ERR_OK / ERR_BAD_PRIVKEY / ERR_BAD_PARAMS / ERR_BAD_PUBKEY (Q != d G) */
err_t bignKeypairValBatch(err_t* status, const bign_params* params, const octet* privkeys,
	const octet* pubkeys, size_t count)
{
	err_t code;
	octet* calc;
	size_t no, i;
	if ((code = params_check(params)))
		return code;
	if (count && (!status || !privkeys || !pubkeys))
		return ERR_BAD_INPUT;
	if (!count)
		return ERR_OK;
	no = params->l / 4;
	if (!(calc = (octet*)malloc(2 * no * count)))
		return ERR_OUTOFMEMORY;
	if (!(code = bignPubkeyCalcBatch(status, calc, params, privkeys, count)))
		for (i = 0; i < count; ++i)
			if (status[i] == ERR_OK && memcmp(calc + 2 * no * i, pubkeys + 2 * no * i, 2 * no) != 0)
				status[i] = ERR_BAD_PUBKEY;
	free(calc);
	return code;
}

err_t bignKeypairVal(const bign_params* params, const octet privkey[], const octet pubkey[])
{
	BIGN_ROUTE(bignKeypairVal, params, privkey, pubkey);
	err_t st = ERR_BAD_INPUT, code;
	if ((code = params_check(params)))
		return code;
	if (!privkey || !pubkey)
		return ERR_BAD_INPUT;
	if ((code = bignKeypairValBatch(&st, params, privkey, pubkey, 1)))
		return code;
	return st;
}

/* status[i] = ERR_OK / ERR_BAD_PUBKEY: coordinates < p and the point on the curve (ecpIsOnA) */
err_t bignPubkeyValBatch(err_t* status, const bign_params* params, const octet* pubkeys, size_t count)
{
	err_t code;
	b2g_slot* s0;
	void *d_p, *d_st;
	size_t no;
	if ((code = params_check(params)))
		return code;
	if (count && (!status || !pubkeys))
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	no = params->l / 4;
	b2g_lock();
	s0 = b2g_slot_get(0);
	if ((code = stage_in(s0, 0, pubkeys, 2 * no * count, &d_p)) || (code = b2g_slot_buf(s0, 1, 4 * count, &d_st)))
		goto done;
	if ((code = b2g_bignPubkeyValBatchL_dev(params->l, d_st, d_p, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(status, d_st, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign status)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign pubkey val)");
done:
	if (code)
		b2g_slot_wipe(s0);   /* failure path: nothing staged (keys, nonces, shared points) stays behind */
	b2g_unlock();
	return code;
}

err_t bignPubkeyVal(const bign_params* params, const octet pubkey[])
{
	BIGN_ROUTE(bignPubkeyVal, params, pubkey);
	err_t st = ERR_BAD_INPUT, code;
	if ((code = params_check(params)))
		return code;
	if (!pubkey)
		return ERR_BAD_INPUT;
	if ((code = bignPubkeyValBatch(&st, params, pubkey, 1)))
		return code;
	return st;
}

/* keys + key_len i <- the first key_len octets of (K.x || K.y), K = d_i Q_i, for items with status 0 */
err_t bignDHBatch(err_t* status, octet* keys, const bign_params* params, const octet* privkeys,
	const octet* pubkeys, size_t key_len, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_k, *d_p, *d_out, *d_st;
	octet* full = 0;
	size_t no, i;
	if ((code = params_check(params)))
		return code;
	if (count && (!status || !privkeys || !pubkeys || (key_len && !keys)))
		return ERR_BAD_INPUT;
	no = params->l / 4;
	if (key_len > 2 * no)
		return ERR_BAD_SHAREDKEY;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	if (!(full = (octet*)malloc(2 * no * count)))
		return ERR_OUTOFMEMORY;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, privkeys, no * count, &d_k)) || (code = stage_in(s0, 1, pubkeys, 2 * no * count, &d_p)) ||
		(code = b2g_slot_buf(s0, 2, 2 * no * count, &d_out)) || (code = b2g_slot_buf(s1, 0, 4 * count, &d_st)))
		goto done;
	CU(cudaMemsetAsync(d_out, 0, 2 * no * count, s0->stream), "memset(bign dh out)");
	if ((code = b2g_bignDHBatchL_dev(params->l, d_st, d_out, d_k, d_p, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(status, d_st, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign status)");
	CU(cudaMemcpyAsync(full, d_out, 2 * no * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign dh)");
	/* private keys and shared points were staged on the device: wipe them */
	CU(cudaMemsetAsync(d_k, 0, no * count, s0->stream), "memset(bign keys)");
	CU(cudaMemsetAsync(d_out, 0, 2 * no * count, s0->stream), "memset(bign dh out)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign dh)");
	for (i = 0; i < count; ++i)
		if (status[i] == ERR_OK)
			memcpy(keys + key_len * i, full + 2 * no * i, key_len);
done:
	if (code)
		b2g_slot_wipe(s0);   /* failure path: nothing staged (keys, nonces, shared points) stays behind */
	b2g_unlock();
	if (full)
	{
		memset(full, 0, 2 * no * count);
		free(full);
	}
	return code;
}

err_t bignDH(octet key[], const bign_params* params, const octet privkey[], const octet pubkey[], size_t key_len)
{
	BIGN_ROUTE(bignDH, key, params, privkey, pubkey, key_len);
	err_t st = ERR_BAD_INPUT, code;
	if ((code = params_check(params)))
		return code;
	if (!privkey || !pubkey || (key_len && !key))
		return ERR_BAD_INPUT;
	if ((code = bignDHBatch(&st, key, params, privkey, pubkey, key_len, 1)))
		return code;
	return st;
}

/* ---------------------------------------------------------------- bignSign (bign_sign.c:27-138) */
/* status[i] = what bignSign would return for item i. One-time keys are drawn on the host with the
   caller's generator in item order and only for items whose private key is valid — the order in
   which `count` successive bignSign calls would consume the generator (:78-91). */
err_t bignSignBatch(err_t* status, octet* sigs, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* privkeys, gen_i rng, void* rng_state, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_h, *d_k, *d_n, *d_sig, *d_st;
	octet* nonces;
	size_t no, so, i, j;
	if ((code = params_check(params)))
		return code;
	no = params->l / 4, so = no + no / 2;
	if (count && (!status || !sigs || !hashes || !privkeys))
		return ERR_BAD_INPUT;
	if (!oid_der_valid(oid_der, oid_len))
		return ERR_BAD_OID;
	if (!rng)
		return ERR_BAD_RNG;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	if (!(nonces = (octet*)calloc(count, no)))
		return ERR_OUTOFMEMORY;
	for (i = 0; i < count; ++i)
	{
		/* 0 < d < q ? (the device repeats the check and sets the status) */
		const octet* d = privkeys + no * i;
		int zero = 1, less = 0;
		for (j = 0; j < no; ++j)
			zero &= d[j] == 0;
		for (j = no; j-- > 0;)
			if (d[j] != params->q[j])
			{
				less = d[j] < params->q[j];
				break;
			}
		if (zero || !less)
			nonces[no * i] = 1;   /* a placeholder: the item ends with ERR_BAD_PRIVKEY */
		else if (!rand_nz_mod(nonces + no * i, params->q, no, rng, rng_state))
		{
			memset(nonces, 0, count * no), free(nonces);
			return ERR_BAD_RNG;
		}
	}
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, hashes, no * count, &d_h)) || (code = stage_in(s0, 1, privkeys, no * count, &d_k)) ||
		(code = stage_in(s0, 3, nonces, no * count, &d_n)) ||
		(code = b2g_slot_buf(s0, 2, so * count, &d_sig)) || (code = b2g_slot_buf(s1, 0, 4 * count, &d_st)))
		goto done;
	if ((code = b2g_bignSignBatchL_k_dev(params->l, d_st, d_sig, oid_der, oid_len, d_h, d_k, d_n, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(status, d_st, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign status)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign sign)");
	{
		/* only items that signed successfully overwrite the caller's sig buffer */
		size_t all_ok = 1;
		for (i = 0; i < count; ++i)
			all_ok &= status[i] == ERR_OK;
		if (all_ok)
			CU(cudaMemcpyAsync(sigs, d_sig, so * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign sigs)");
		else
			for (i = 0; i < count; ++i)
				if (status[i] == ERR_OK)
					CU(cudaMemcpyAsync(sigs + so * i, (octet*)d_sig + so * i, so, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign sig)");
	}
	/* private and one-time keys were staged on the device: wipe them */
	CU(cudaMemsetAsync(d_k, 0, no * count, s0->stream), "memset(bign keys)");
	CU(cudaMemsetAsync(d_n, 0, no * count, s0->stream), "memset(bign nonces)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign sign)");
done:
	if (code)
		b2g_slot_wipe(s0);   /* failure path: nothing staged (keys, nonces, shared points) stays behind */
	b2g_unlock();
	memset(nonces, 0, count * no), free(nonces);
	return code;
}

err_t bignSign(octet sig[], const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet privkey[], gen_i rng, void* rng_state)
{
	BIGN_ROUTE(bignSign, sig, params, oid_der, oid_len, hash, privkey, rng, rng_state);
	err_t st = ERR_BAD_INPUT, code;
	if (oid_len > 64)                      /* not staged into kernel arguments: before the generator is touched */
		B2G_STOCK_R(bignSign, sig, params, oid_der, oid_len, hash, privkey, rng, rng_state);
	if ((code = params_check(params)))
		return code;
	if (!hash || !privkey || !sig)
		return ERR_BAD_INPUT;
	if ((code = bignSignBatch(&st, sig, params, oid_der, oid_len, hash, privkey, rng, rng_state, 1)))
		return code;   /* (a too-long OID is caught before the generator is touched: see bignSignBatch) */
	return st;
}

/* ---------------------------------------------------------------- fixed-level wrappers
   bign128.c:96-185, bign192.c, bign256.c: the standard curve of the level and the OID of its hash
   algorithm (belt-hash / bash384 / bash512). The reference keeps a lazily created curve per level
   (bign128.c:34-88); here the parameter block is a constant and the per-level tables live on the device. */
static const octet oid_belt_hash[] = {0x06, 0x09, 0x2A, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x1F, 0x51};
static const octet oid_bash384[] = {0x06, 0x09, 0x2A, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x4D, 0x0C};
static const octet oid_bash512[] = {0x06, 0x09, 0x2A, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x4D, 0x0D};

#define BIGN_LEVEL_WRAPPERS(PFX, IDX, OID)                                                                   \
	static const bign_params* PFX##_params(bign_params* p) { std_fill(p, &std_curves[IDX]); return p; }     \
	err_t PFX##KeypairGen(octet privkey[], octet pubkey[], gen_i rng, void* rng_state)                       \
	{ bign_params p; return bignKeypairGen(privkey, pubkey, PFX##_params(&p), rng, rng_state); }             \
	err_t PFX##KeypairVal(const octet privkey[], const octet pubkey[])                                       \
	{ bign_params p; return bignKeypairVal(PFX##_params(&p), privkey, pubkey); }                             \
	err_t PFX##PubkeyVal(const octet pubkey[])                                                               \
	{ bign_params p; return bignPubkeyVal(PFX##_params(&p), pubkey); }                                       \
	err_t PFX##PubkeyCalc(octet pubkey[], const octet privkey[])                                             \
	{ bign_params p; return bignPubkeyCalc(pubkey, PFX##_params(&p), privkey); }                             \
	err_t PFX##DH(octet key[], const octet privkey[], const octet pubkey[], size_t key_len)                  \
	{ bign_params p; return bignDH(key, PFX##_params(&p), privkey, pubkey, key_len); }                       \
	err_t PFX##Sign(octet sig[], const octet hash[], const octet privkey[], gen_i rng, void* rng_state)      \
	{ bign_params p; return bignSign(sig, PFX##_params(&p), OID, sizeof(OID), hash, privkey, rng, rng_state); } \
	err_t PFX##Sign2(octet sig[], const octet hash[], const octet privkey[], const void* t, size_t t_len)    \
	{ bign_params p; return bignSign2(sig, PFX##_params(&p), OID, sizeof(OID), hash, privkey, t, t_len); }   \
	err_t PFX##Verify(const octet hash[], const octet sig[], const octet pubkey[])                           \
	{ bign_params p; return bignVerify(PFX##_params(&p), OID, sizeof(OID), hash, sig, pubkey); }

BIGN_LEVEL_WRAPPERS(bign128, 0, oid_belt_hash)
BIGN_LEVEL_WRAPPERS(bign192, 1, oid_bash384)
BIGN_LEVEL_WRAPPERS(bign256, 2, oid_bash512)

/* ---------------------------------------------------------------- ecMulA behind the reference's ec_o
   (include/bee2/math/ec.h:540-571, :892-901; qr.h:317-338; obj.h:53-58). The reference dispatches every
   point operation through the vtable of `ec`; here only the DATA at the head of the two descriptions are
   read — field size and modulus, A, B — to recognise one of the three standard bign curves, and the whole
   multiplication runs on the device. Fields over a Crandall prime keep plain residues (zm.c:270-310), so
   the word arrays of a 64-bit little-endian build are the octet strings the kernels take. Any other curve:
   there is no CPU path — the call fails loudly. */
typedef struct { size_t keep, p_count, o_count; } b2g_obj_hdr;
typedef struct
{
	b2g_obj_hdr hdr;
	const u64* mod;
	const u64* unity;
	const void* params;
	size_t n, no;
	/* function pointers follow */
} b2g_qr_view;
typedef struct
{
	b2g_obj_hdr hdr;
	const b2g_qr_view* f;
	const u64 *A, *B, *base, *order;
	const void* pre;
	size_t d;
	u64 cofactor;
	/* function pointers follow */
} b2g_ec_view;

static size_t ec_std_level(const void* ec_)
{
	const b2g_ec_view* ec = (const b2g_ec_view*)ec_;
	bign_params std;
	size_t i;
	if (!ec || !ec->f || !ec->f->mod || !ec->A || !ec->B)
		return 0;
	for (i = 0; i < 3; ++i)
	{
		const size_t no = std_curves[i].l / 4;
		if (ec->f->n != no / 8 || ec->f->no != no)
			continue;
		std_fill(&std, &std_curves[i]);
		if (!memcmp(ec->f->mod, std.p, no) && !memcmp(ec->A, std.a, no) && !memcmp(ec->B, std.b, no))
			return std_curves[i].l;
	}
	return 0;
}

size_t ecMulA_deep(size_t n, size_t ec_d, size_t ec_deep, size_t m)
{
	/* overlay mode: a call may be forwarded to stock libbee2, whose ecMulA works in the caller's stack */
	B2G_STOCK_R(ecMulA_deep, n, ec_d, ec_deep, m);
	return 0;   /* no host scratch: the caller's `stack` is not used */
}

/* b <- d a, FALSE iff the result is O (ec.c:497-525); a, b: 2n words affine; d: m words, m <= n */
bool_t ecMulA(u64 b[], const u64 a[], const void* ec, const u64 d[], size_t m, void* stack)
{
	const size_t l = ec_std_level(ec);
	octet out[128];
	int ok = 0;
	err_t code;
	/* a curve other than the three standard bign ones, or a scalar longer than the field (ec.c:497-525
	   takes any ec_o and any m): stock libbee2 behind this library, else there is no path */
	if (!l || !a || !b || !d || m == 0 || 8 * m > l / 4)
		B2G_FAIL_R(l ? ERR_BAD_INPUT : ERR_NOT_IMPLEMENTED, ecMulA, b, a, ec, d, m, stack);
	B2G_SMALL_R(1, ecMulA, b, a, ec, d, m, stack);
	if ((code = ecMulABatchL(l, out, &ok, (const octet*)a, (const octet*)d, 8 * m, 1)))
		B2G_FAIL_R(code, ecMulA, b, a, ec, d, m, stack);
	if (ok)
		memcpy(b, out, l / 2);
	b2g_wipe(out, sizeof out);   /* the product may be a shared secret (bignDH, key transport) */
	return ok ? 1 : 0;
}

/* b <- sum_{i<k} d_i a_i for the k (a_i, d_i, m_i) triples that follow k (ec.h:1176-1190, ec.c:1183-1273),
   FALSE iff the sum is O. The k products run as one batch of ecp_mul_kernel, a one-thread kernel adds
   them. Same curve recognition and limits as ecMulA (m_i <= n words, k <= 64). */
/* the variadic call rebuilt for stock libbee2 (k <= 8 triples; the reference's own callers use 2..4) */
static int addmul_stock(bool_t* ret, u64 b[], const void* ec, void* stack, size_t k, const u64* const a[],
	const u64* const d[], const size_t m[])
{
	__typeof__(&ecAddMulA) f = B2G_STOCK_FN(ecAddMulA);
	if (!f || k == 0 || k > 8)
		return 0;
	b2g_note_forward("ecAddMulA");
#define T_(i) a[i], d[i], m[i]
	switch (k)
	{
	case 1: *ret = f(b, ec, stack, k, T_(0)); break;
	case 2: *ret = f(b, ec, stack, k, T_(0), T_(1)); break;
	case 3: *ret = f(b, ec, stack, k, T_(0), T_(1), T_(2)); break;
	case 4: *ret = f(b, ec, stack, k, T_(0), T_(1), T_(2), T_(3)); break;
	case 5: *ret = f(b, ec, stack, k, T_(0), T_(1), T_(2), T_(3), T_(4)); break;
	case 6: *ret = f(b, ec, stack, k, T_(0), T_(1), T_(2), T_(3), T_(4), T_(5)); break;
	case 7: *ret = f(b, ec, stack, k, T_(0), T_(1), T_(2), T_(3), T_(4), T_(5), T_(6)); break;
	default: *ret = f(b, ec, stack, k, T_(0), T_(1), T_(2), T_(3), T_(4), T_(5), T_(6), T_(7)); break;
	}
#undef T_
	return 1;
}

bool_t ecAddMulA(u64 b[], const void* ec, void* stack, size_t k, ...)
{
	const size_t l = ec_std_level(ec);
	const size_t no = l / 4;
	octet pts[64 * 128], scal[64 * 64], out[128];
	const u64 *va[8], *vd[8];
	size_t vm[8];
	b2g_slot* sl = 0;
	void *d_a, *d_d, *d_p, *d_ok, *d_out;
	int ok = 0, foreign = !l || !b || k == 0 || k > 64;
	bool_t ret = 0;
	err_t code;
	size_t i;
	va_list ap;
	memset(scal, 0, sizeof scal);
	va_start(ap, k);
	for (i = 0; i < k; ++i)
	{
		const u64* a = va_arg(ap, const u64*);
		const u64* d = va_arg(ap, const u64*);
		const size_t m = va_arg(ap, size_t);
		if (i < 8)
			va[i] = a, vd[i] = d, vm[i] = m;
		if (!a || !d || 8 * m > no)
			foreign = 1;
		else if (!foreign && i < 64)
		{
			memcpy(pts + 2 * no * i, a, 2 * no);
			memcpy(scal + no * i, d, 8 * m);
		}
	}
	va_end(ap);
	/* another curve, a scalar longer than the field, more than 64 terms: stock libbee2, else no path */
	if (foreign || b2g_route_small(1) || (code = b2g_ensure_device()))
	{
		if (addmul_stock(&ret, b, ec, stack, k, va, vd, vm))
		{
			b2g_wipe(scal, sizeof scal);
			return ret;
		}
		if (!foreign && !b2g_ensure_device())
			goto gpu;                  /* small-call routing asked for stock, but there is none */
		b2g_wipe(scal, sizeof scal);
		b2g_die("ecAddMulA (not a standard bign curve, m > n or k > 64, and no stock libbee2 behind)",
			l ? ERR_BAD_INPUT : ERR_NOT_IMPLEMENTED);
	}
gpu:
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = stage_in(sl, 0, pts, 2 * no * k, &d_a)) || (code = stage_in(sl, 1, scal, no * k, &d_d)) ||
		(code = b2g_slot_buf(sl, 2, 2 * no * k, &d_p)) || (code = b2g_slot_buf(sl, 3, 4 * k + 8 + 2 * no, &d_ok)))
		goto done;
	d_out = (octet*)d_ok + 4 * k + 8;
	if ((code = b2g_ecMulABatchL_dev(l, d_p, d_ok, d_a, d_d, no, k, sl->stream)) ||
		(code = b2g_ecSumL_dev(l, d_out, (octet*)d_ok + 4 * k, d_p, d_ok, k, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(&ok, (octet*)d_ok + 4 * k, 4, cudaMemcpyDeviceToHost, sl->stream), "D2H(ecAddMulA ok)");
	CU(cudaMemcpyAsync(out, d_out, 2 * no, cudaMemcpyDeviceToHost, sl->stream), "D2H(ecAddMulA)");
	/* the scalars may be secret: wipe the staged copy */
	CU(cudaMemsetAsync(d_d, 0, no * k, sl->stream), "memset(ecAddMulA scalars)");
	CU(cudaStreamSynchronize(sl->stream), "sync(ecAddMulA)");
done:
	if (code)
		b2g_slot_wipe(sl);
	b2g_unlock();
	b2g_wipe(scal, sizeof scal);
	if (code)
	{
		if (addmul_stock(&ret, b, ec, stack, k, va, vd, vm))
			return ret;
		b2g_die("ecAddMulA", code);
	}
	if (ok)
		memcpy(b, out, 2 * no);
	b2g_wipe(out, sizeof out);
	return ok ? 1 : 0;
}
