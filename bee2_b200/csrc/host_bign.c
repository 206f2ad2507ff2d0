/*
 * host_bign.c — host side (C) of the bign path: the reference's bign.h surface for
 * bignParamsStd / bignVerify / bignSign2 / bignPubkeyCalc (bign_params.c:197-280,
 * bign_sign.c:247-260, :349-361, bign_misc.c:369-412) and the host-pointer batch entry
 * points. Argument checks (parameter block, OID DER syntax) mirror the reference; all
 * field, curve and hash arithmetic runs in bign.cu on the device.
 */
#include "engine.h"
#include <string.h>

#define CU(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { code = b2g_cuda_fail(e_, what); goto done; } } while (0)

/* bign-curve256v1 = level l = 128, table B.1 of STB 34.101.45 (bign_params.c:33-73) */
static const char curve256v1_name[] = "1.2.112.0.2.0.34.101.45.3.1";
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

err_t bignParamsStd(bign_params* params, const char* name)
{
	if (!params)
		return ERR_BAD_INPUT;
	memset(params, 0, sizeof *params);
	if (!name || strcmp(name, curve256v1_name) != 0)
		return ERR_FILE_NOT_FOUND;
	params->l = 128;
	memset(params->p, 0xFF, 32), params->p[0] = 0x43;   /* p = 2^256 - 189 */
	memset(params->a, 0xFF, 32), params->a[0] = 0x40;   /* a = p - 3 */
	memcpy(params->b, curve256v1_b, 32);
	memcpy(params->q, curve256v1_q, 32);
	memcpy(params->yG, curve256v1_yG, 32);
	memcpy(params->seed, curve256v1_seed, 8);
	return ERR_OK;
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
	/* GPU path: bign-curve256v1 only (seed is not used by sign/verify) */
	bignParamsStd(&std, curve256v1_name);
	if (params->l != 128 || memcmp(params->p, std.p, 64) || memcmp(params->a, std.a, 64) ||
		memcmp(params->b, std.b, 64) || memcmp(params->q, std.q, 64) || memcmp(params->yG, std.yG, 64))
		return ERR_NOT_IMPLEMENTED;
	return ERR_OK;
}

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

err_t bignVerifyBatch(err_t* status, const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet* hashes, const octet* sigs, const octet* pubkeys, size_t count)
{
	err_t code;
	b2g_slot *s0 = b2g_slot_get(0), *s1 = b2g_slot_get(1);
	void *d_h, *d_s, *d_p, *d_st;
	if ((code = params_check(params)))
		return code;
	if (count && (!status || !hashes || !sigs || !pubkeys))
		return ERR_BAD_INPUT;
	if (!oid_der_valid(oid_der, oid_len))
		return ERR_BAD_OID;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	/* chunks of 2^16 items alternate between the two workspace slots, so the H2D copy of one
	   chunk overlaps the kernel of the previous one */
	{
		const size_t chunk = (size_t)1 << 16;
		size_t off, c;
		for (off = 0, c = 0; off < count; off += chunk, ++c)
		{
			const size_t n = count - off < chunk ? count - off : chunk;
			b2g_slot* sl = b2g_slot_get((int)c);
			if ((code = stage_in(sl, 0, hashes + 32 * off, 32 * n, &d_h)) ||
				(code = stage_in(sl, 1, sigs + 48 * off, 48 * n, &d_s)) ||
				(code = stage_in(sl, 2, pubkeys + 64 * off, 64 * n, &d_p)) ||
				(code = b2g_slot_buf(sl, 3, 4 * n, &d_st)))
				goto done;
			if ((code = b2g_bignVerifyBatch_dev(d_st, oid_der, oid_len, d_h, d_s, d_p, n, sl->stream)))
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

err_t bignVerify(const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet sig[], const octet pubkey[])
{
	err_t st = ERR_BAD_SIG, code;
	/* order of checks: params (bign_sign.c:355-356), buffers, OID (:286-290) */
	if ((code = params_check(params)))
		return code;
	if (!hash || !sig || !pubkey)
		return ERR_BAD_INPUT;
	if ((code = bignVerifyBatch(&st, params, oid_der, oid_len, hash, sig, pubkey, 1)))
		return code;
	return st;
}

static err_t sign2_batch(err_t* status, octet* sigs, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* privkeys, size_t count, const void* t, size_t t_len)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_h, *d_k, *d_sig, *d_st;
	if ((code = params_check(params)))
		return code;
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
	if ((code = stage_in(s0, 0, hashes, 32 * count, &d_h)) || (code = stage_in(s0, 1, privkeys, 32 * count, &d_k)) ||
		(code = b2g_slot_buf(s0, 2, 48 * count, &d_sig)) || (code = b2g_slot_buf(s1, 0, 4 * count, &d_st)))
		goto done;
	if ((code = b2g_bignSign2Batch_t_dev(d_st, d_sig, oid_der, oid_len, d_h, d_k, count, t, t_len, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(status, d_st, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign status)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign sign2)");
	{
		/* only items that signed successfully overwrite the caller's sig buffer */
		size_t i, all_ok = 1;
		for (i = 0; i < count; ++i)
			all_ok &= status[i] == ERR_OK;
		if (all_ok)
			CU(cudaMemcpy(sigs, d_sig, 48 * count, cudaMemcpyDeviceToHost), "D2H(bign sigs)");
		else
			for (i = 0; i < count; ++i)
				if (status[i] == ERR_OK)
					CU(cudaMemcpy(sigs + 48 * i, (octet*)d_sig + 48 * i, 48, cudaMemcpyDeviceToHost), "D2H(bign sig)");
	}
	/* the private keys were staged on the device: wipe them */
	CU(cudaMemsetAsync(d_k, 0, 32 * count, s0->stream), "memset(bign keys)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign sign2)");
done:
	if (code)
		cudaStreamSynchronize(s0->stream);
	b2g_unlock();
	return code;
}

err_t bignSign2Batch(err_t* status, octet* sigs, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* privkeys, size_t count)
{
	return sign2_batch(status, sigs, params, oid_der, oid_len, hashes, privkeys, count, 0, 0);
}

err_t bignSign2(octet sig[], const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet privkey[], const void* t, size_t t_len)
{
	err_t st = ERR_BAD_INPUT, code;
	if ((code = params_check(params)))
		return code;
	if (!hash || !privkey || !sig)
		return ERR_BAD_INPUT;
	if ((code = sign2_batch(&st, sig, params, oid_der, oid_len, hash, privkey, 1, t, t ? t_len : 0)))
		return code;
	return st;
}

err_t bignPubkeyCalcBatch(err_t* status, octet* pubkeys, const bign_params* params,
	const octet* privkeys, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_k, *d_p, *d_st;
	if ((code = params_check(params)))
		return code;
	if (count && (!status || !pubkeys || !privkeys))
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, privkeys, 32 * count, &d_k)) || (code = b2g_slot_buf(s0, 1, 64 * count, &d_p)) ||
		(code = b2g_slot_buf(s1, 0, 4 * count, &d_st)))
		goto done;
	CU(cudaMemsetAsync(d_p, 0, 64 * count, s0->stream), "memset(bign pubkeys)");
	if ((code = b2g_bignPubkeyCalcBatch_dev(d_st, d_p, d_k, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(status, d_st, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign status)");
	CU(cudaMemcpyAsync(pubkeys, d_p, 64 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bign pubkeys)");
	CU(cudaMemsetAsync(d_k, 0, 32 * count, s0->stream), "memset(bign keys)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bign pubkey)");
done:
	if (code)
		cudaStreamSynchronize(s0->stream);
	b2g_unlock();
	return code;
}

err_t bignPubkeyCalc(octet pubkey[], const bign_params* params, const octet privkey[])
{
	err_t st = ERR_BAD_INPUT, code;
	octet out[64];
	if ((code = params_check(params)))
		return code;
	if (!pubkey || !privkey)
		return ERR_BAD_INPUT;
	if ((code = bignPubkeyCalcBatch(&st, out, params, privkey, 1)))
		return code;
	if (st == ERR_OK)
		memcpy(pubkey, out, 64);
	return st;
}

err_t ecMulABatch(octet* b, int* ok, const octet* a, const octet* d, size_t d_len, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_a, *d_d, *d_b, *d_ok;
	if (count && (!b || !ok || !a || !d))
		return ERR_BAD_INPUT;
	if (d_len == 0 || d_len > 32)
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, a, 64 * count, &d_a)) || (code = stage_in(s0, 1, d, d_len * count, &d_d)) ||
		(code = b2g_slot_buf(s0, 2, 64 * count, &d_b)) || (code = b2g_slot_buf(s1, 0, 4 * count, &d_ok)))
		goto done;
	CU(cudaMemsetAsync(d_b, 0, 64 * count, s0->stream), "memset(ecMulA out)");
	if ((code = b2g_ecMulABatch_dev(d_b, d_ok, d_a, d_d, d_len, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(b, d_b, 64 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecMulA)");
	CU(cudaMemcpyAsync(ok, d_ok, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecMulA ok)");
	CU(cudaStreamSynchronize(s0->stream), "sync(ecMulA)");
done:
	if (code)
		cudaStreamSynchronize(s0->stream);
	b2g_unlock();
	return code;
}

/* b_i = d_i * a_i + k_i * G (ecAddMulA with the base point as second term, ec.c:1183-1273) */
err_t ecAddMulABatch(octet* b, int* ok, const octet* a, const octet* d, size_t d_len, const octet* k,
	size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_a, *d_d, *d_k, *d_b, *d_ok;
	if (count && (!b || !ok || !a || !d || !k))
		return ERR_BAD_INPUT;
	if (d_len == 0 || d_len > 32)
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = stage_in(s0, 0, a, 64 * count, &d_a)) || (code = stage_in(s0, 1, d, d_len * count, &d_d)) ||
		(code = stage_in(s0, 2, k, 32 * count, &d_k)) || (code = b2g_slot_buf(s1, 0, 64 * count, &d_b)) ||
		(code = b2g_slot_buf(s1, 1, 4 * count, &d_ok)))
		goto done;
	CU(cudaMemsetAsync(d_b, 0, 64 * count, s0->stream), "memset(ecAddMulA out)");
	if ((code = b2g_ecAddMulABatch_dev(d_b, d_ok, d_a, d_d, d_len, d_k, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(b, d_b, 64 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecAddMulA)");
	CU(cudaMemcpyAsync(ok, d_ok, 4 * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(ecAddMulA ok)");
	CU(cudaStreamSynchronize(s0->stream), "sync(ecAddMulA)");
done:
	if (code)
		cudaStreamSynchronize(s0->stream);
	b2g_unlock();
	return code;
}
