/*
 * host_belt.c — host side (C) of the belt path: the reference's belt.h surface for the
 * block function, ECB and CTR (belt_block.c:72-373, belt_ecb.c:44-158, belt_ctr.c:27-135,
 * belt_hash.c:174-190) plus the host-pointer batch entry points. State structs keep the
 * reference's layouts (belt_ecb.c:44-48, belt_lcl.h:135-141). Key formatting and
 * buffer bookkeeping happen here; every block encryption happens on the device.
 */
#include "engine.h"
#include <string.h>

typedef struct { u32 key[8]; octet block[16]; } belt_ecb_st;
typedef struct { u32 key[8]; u32 ctr[4]; octet block[16]; size_t reserved; } belt_ctr_st;

#define CU(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { code = b2g_cuda_fail(e_, what); goto done; } } while (0)
#define CHUNK_BYTES ((size_t)32 << 20)

static int key_len_ok(size_t len) { return len == 16 || len == 24 || len == 32; }

static u32 sync_all(void)
{
	int c;
	u32 code = ERR_OK;
	for (c = 0; c < B2G_NSLOT; ++c)
	{
		cudaError_t e = cudaStreamSynchronize(b2g_slot_get(c)->stream);
		if (e != cudaSuccess && !code)
			code = b2g_cuda_fail(e, "cudaStreamSynchronize");
	}
	return code;
}

/* ---------------------------------------------------------------- key schedule (belt_block.c:72-106) */
void beltKeyExpand(octet key_[32], const octet key[], size_t len)
{
	size_t i;
	memmove(key_, key, len);
	if (len == 16)
		memcpy(key_ + 16, key_, 16);
	else if (len == 24)
		for (i = 0; i < 8; ++i)
			key_[24 + i] = key_[i] ^ key_[8 + i] ^ key_[16 + i];
}

void beltKeyExpand2(u32 key_[8], const octet key[], size_t len)
{
	memcpy(key_, key, len);   /* little-endian host: u32From is a copy (u32.c:233-244) */
	if (len == 16)
		key_[4] = key_[0], key_[5] = key_[1], key_[6] = key_[2], key_[7] = key_[3];
	else if (len == 24)
		/* word-wise, as STB 34.101.31 defines it; note that the reference's octet variant above
		   combines 8-octet halves instead (belt_block.c:82-85 vs :101-104) — both are mirrored */
		key_[6] = key_[0] ^ key_[1] ^ key_[2], key_[7] = key_[3] ^ key_[4] ^ key_[5];
}

/* ---------------------------------------------------------------- single blocks */
static err_t blocks_small(void* blocks, size_t n, const u32 key[8], int decrypt)
{
	err_t code;
	b2g_slot* sl;
	void* d;
	if ((code = b2g_ensure_device()))
		return code;
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, 16 * n, &d)))
		goto done;
	CU(cudaMemcpyAsync(d, blocks, 16 * n, cudaMemcpyHostToDevice, sl->stream), "H2D(belt block)");
	if ((code = b2g_beltECB_dev(d, d, n, key, decrypt, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(blocks, d, 16 * n, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt block)");
	CU(cudaStreamSynchronize(sl->stream), "sync(belt block)");
done:
	b2g_unlock();
	return code;
}

void beltBlockEncr(octet block[16], const u32 key[8])
{
	err_t code = blocks_small(block, 1, key, 0);
	if (code) b2g_die("beltBlockEncr", code);
}
void beltBlockEncr2(u32 block[4], const u32 key[8])
{
	err_t code = blocks_small(block, 1, key, 0);
	if (code) b2g_die("beltBlockEncr2", code);
}
void beltBlockEncr3(u32* a, u32* b, u32* c, u32* d, const u32 key[8])
{
	u32 t[4];
	err_t code;
	t[0] = *a, t[1] = *b, t[2] = *c, t[3] = *d;
	if ((code = blocks_small(t, 1, key, 0))) b2g_die("beltBlockEncr3", code);
	*a = t[0], *b = t[1], *c = t[2], *d = t[3];
}
void beltBlockDecr(octet block[16], const u32 key[8])
{
	err_t code = blocks_small(block, 1, key, 1);
	if (code) b2g_die("beltBlockDecr", code);
}
void beltBlockDecr2(u32 block[4], const u32 key[8])
{
	err_t code = blocks_small(block, 1, key, 1);
	if (code) b2g_die("beltBlockDecr2", code);
}
void beltBlockDecr3(u32* a, u32* b, u32* c, u32* d, const u32 key[8])
{
	u32 t[4];
	err_t code;
	t[0] = *a, t[1] = *b, t[2] = *c, t[3] = *d;
	if ((code = blocks_small(t, 1, key, 1))) b2g_die("beltBlockDecr3", code);
	*a = t[0], *b = t[1], *c = t[2], *d = t[3];
}

/* ---------------------------------------------------------------- ECB */
size_t beltECB_keep(void) { return sizeof(belt_ecb_st); }

void beltECBStart(void* state, const octet key[], size_t len)
{
	beltKeyExpand2(((belt_ecb_st*)state)->key, key, len);
}

/* whole blocks dest <- E/D(src), pipelined in chunks over two streams */
static err_t ecb_run(octet* dest, const octet* src, size_t nblocks, const u32 key[8], int decrypt)
{
	err_t code = ERR_OK;
	const size_t chunk = b2g_chunk_units(16, CHUNK_BYTES);
	size_t off, c;
	if (!nblocks)
		return ERR_OK;
	if ((code = b2g_ensure_device()))
		return code;
	b2g_lock();
	for (off = 0, c = 0; off < nblocks; off += chunk, ++c)
	{
		const size_t n = nblocks - off < chunk ? nblocks - off : chunk;
		b2g_slot* sl = b2g_slot_get((int)c);
		void* d;
		if ((code = b2g_slot_buf(sl, 0, 16 * n, &d)))
			goto done;
		CU(cudaMemcpyAsync(d, src + 16 * off, 16 * n, cudaMemcpyHostToDevice, sl->stream), "H2D(belt ecb)");
		if ((code = b2g_beltECB_dev(d, d, n, key, decrypt, sl->stream)))
			goto done;
		CU(cudaMemcpyAsync(dest + 16 * off, d, 16 * n, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt ecb)");
	}
	code = sync_all();
done:
	if (code)
		sync_all();
	b2g_unlock();
	return code;
}

/* full blocks, then ciphertext stealing for a ragged tail (belt_ecb.c:76-84, :100-108) */
static err_t ecb_step(octet* buf, size_t count, const u32 key[8], int decrypt)
{
	const size_t full = count / 16, r = count % 16;
	err_t code = ecb_run(buf, buf, full, key, decrypt);
	if (!code && r)
	{
		octet t[16];
		octet* last = buf + 16 * (full - 1);
		memcpy(t, last + 16, r);
		memcpy(t + r, last + r, 16 - r);
		code = blocks_small(t, 1, key, decrypt);
		memcpy(last + 16, last, r);
		memcpy(last, t, 16);
	}
	return code;
}

void beltECBStepE(void* buf, size_t count, void* state)
{
	err_t code = ecb_step((octet*)buf, count, ((belt_ecb_st*)state)->key, 0);
	if (code) b2g_die("beltECBStepE", code);
}
void beltECBStepD(void* buf, size_t count, void* state)
{
	err_t code = ecb_step((octet*)buf, count, ((belt_ecb_st*)state)->key, 1);
	if (code) b2g_die("beltECBStepD", code);
}

static err_t ecb_oneshot(void* dest, const void* src, size_t count, const octet key[], size_t len, int decrypt)
{
	u32 k[8];
	if (count < 16 || !key_len_ok(len) || !src || !key || !dest)
		return ERR_BAD_INPUT;
	beltKeyExpand2(k, key, len);
	memmove(dest, src, count);
	return ecb_step((octet*)dest, count, k, decrypt);
}
err_t beltECBEncr(void* dest, const void* src, size_t count, const octet key[], size_t len)
{
	return ecb_oneshot(dest, src, count, key, len, 0);
}
err_t beltECBDecr(void* dest, const void* src, size_t count, const octet key[], size_t len)
{
	return ecb_oneshot(dest, src, count, key, len, 1);
}

err_t beltECBEncrBatch(void* blocks, const octet* keys32, size_t count)
{
	err_t code = ERR_OK;
	const size_t chunk = b2g_chunk_units(48, CHUNK_BYTES);
	size_t off, c;
	if (count && (!blocks || !keys32))
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	for (off = 0, c = 0; off < count; off += chunk, ++c)
	{
		const size_t n = count - off < chunk ? count - off : chunk;
		b2g_slot* sl = b2g_slot_get((int)c);
		void *d_b, *d_k;
		if ((code = b2g_slot_buf(sl, 0, 16 * n, &d_b)) || (code = b2g_slot_buf(sl, 1, 32 * n, &d_k)))
			goto done;
		CU(cudaMemcpyAsync(d_k, keys32 + 32 * off, 32 * n, cudaMemcpyHostToDevice, sl->stream), "H2D(belt keys)");
		CU(cudaMemcpyAsync(d_b, (octet*)blocks + 16 * off, 16 * n, cudaMemcpyHostToDevice, sl->stream), "H2D(belt blocks)");
		if ((code = b2g_beltECBEncrBatch_dev(d_b, d_k, n, sl->stream)))
			goto done;
		CU(cudaMemcpyAsync((octet*)blocks + 16 * off, d_b, 16 * n, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt blocks)");
	}
	code = sync_all();
done:
	if (code)
		sync_all();
	b2g_unlock();
	return code;
}

/* ---------------------------------------------------------------- CTR */
size_t beltCTR_keep(void) { return sizeof(belt_ctr_st); }

/* ctr <- ctr + n as a 128-bit little-endian integer (n applications of belt_ctr.c:27-35) */
static void ctr_add(u32 ctr[4], u64 n)
{
	u64 lo = ((u64)ctr[1] << 32 | ctr[0]) + n;
	u64 hi = ((u64)ctr[3] << 32 | ctr[2]) + (lo < n ? 1 : 0);
	ctr[0] = (u32)lo, ctr[1] = (u32)(lo >> 32), ctr[2] = (u32)hi, ctr[3] = (u32)(hi >> 32);
}

void beltCTRStart(void* state, const octet key[], size_t len, const octet iv[16])
{
	belt_ctr_st* st = (belt_ctr_st*)state;
	err_t code;
	beltKeyExpand2(st->key, key, len);
	memcpy(st->ctr, iv, 16);
	if ((code = blocks_small(st->ctr, 1, st->key, 0)))
		b2g_die("beltCTRStart", code);
	st->reserved = 0;
}

/* dest[0..count) <- src[0..count) ^ keystream(key, ctr0) (src == NULL: keystream only).
   If last_ks != NULL and count % 16 != 0 it receives the whole last keystream block. */
static err_t ctr_run(octet* dest, const octet* src, size_t count, const u32 key[8], const u32 ctr0[4],
	octet last_ks[16])
{
	err_t code = ERR_OK;
	const size_t chunk = b2g_chunk_units(16, CHUNK_BYTES) * 16;   /* bytes, multiple of 16 */
	size_t off, c;
	if (!count)
		return ERR_OK;
	if ((code = b2g_ensure_device()))
		return code;
	b2g_lock();
	for (off = 0, c = 0; off < count; off += chunk, ++c)
	{
		const size_t n = count - off < chunk ? count - off : chunk;
		const size_t padded = (n + 15) & ~(size_t)15;
		b2g_slot* sl = b2g_slot_get((int)c);
		void* d;
		if ((code = b2g_slot_buf(sl, 0, padded, &d)))
			goto done;
		if (src)
		{
			if (padded != n)
				CU(cudaMemsetAsync((octet*)d + padded - 16, 0, 16, sl->stream), "memset(belt ctr)");
			CU(cudaMemcpyAsync(d, src + off, n, cudaMemcpyHostToDevice, sl->stream), "H2D(belt ctr)");
		}
		/* the padded tail is produced on the device buffer; only n octets go back */
		if ((code = b2g_beltCTR_dev(d, src ? d : 0, padded, key, ctr0, off / 16, sl->stream)))
			goto done;
		CU(cudaMemcpyAsync(dest + off, d, n, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt ctr)");
		if (padded != n && last_ks)
			CU(cudaMemcpyAsync(last_ks, (octet*)d + padded - 16, 16, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt ctr tail)");
	}
	code = sync_all();
done:
	if (code)
		sync_all();
	b2g_unlock();
	return code;
}

void beltCTRStepE(void* buf, size_t count, void* state)
{
	belt_ctr_st* st = (belt_ctr_st*)state;
	octet* p = (octet*)buf;
	octet last[16];
	size_t i, r;
	err_t code;
	/* reserve of keystream octets left from the previous call (belt_ctr.c:70-83) */
	if (st->reserved)
	{
		const size_t take = st->reserved < count ? st->reserved : count;
		for (i = 0; i < take; ++i)
			p[i] ^= st->block[16 - st->reserved + i];
		st->reserved -= take, p += take, count -= take;
	}
	if (!count)
		return;
	if ((code = ctr_run(p, p, count, st->key, st->ctr, last)))
		b2g_die("beltCTRStepE", code);
	ctr_add(st->ctr, (count + 15) / 16);
	r = count % 16;
	if (r)
	{
		/* last[] = data^ks on [0,r) and pure keystream on [r,16): keep the unused part */
		memcpy(st->block + r, last + r, 16 - r);
		st->reserved = 16 - r;
	}
}

err_t beltCTR(void* dest, const void* src, size_t count, const octet key[], size_t len,
	const octet iv[16])
{
	belt_ctr_st st;
	err_t code;
	if (!key_len_ok(len) || (count && (!src || !dest)) || !key || !iv)
		return ERR_BAD_INPUT;
	beltKeyExpand2(st.key, key, len);
	memcpy(st.ctr, iv, 16);
	if ((code = blocks_small(st.ctr, 1, st.key, 0)))
		return code;
	return ctr_run((octet*)dest, (const octet*)src, count, st.key, st.ctr, 0);
}

err_t beltCTRKeystream(void* dest, size_t count, const octet key[], size_t len, const octet iv[16])
{
	belt_ctr_st st;
	err_t code;
	if (!key_len_ok(len) || (count && !dest) || !key || !iv)
		return ERR_BAD_INPUT;
	beltKeyExpand2(st.key, key, len);
	memcpy(st.ctr, iv, 16);
	if ((code = blocks_small(st.ctr, 1, st.key, 0)))
		return code;
	return ctr_run((octet*)dest, 0, count, st.key, st.ctr, 0);
}

/* ---------------------------------------------------------------- belt-hash */
err_t beltHashBatch(octet* hashes, const void* msgs, size_t msg_len, size_t stride, size_t count)
{
	err_t code;
	size_t pitch, chunk, off, c;
	int contiguous;
	if (count && (!hashes || (msg_len && !msgs) || (count > 1 && stride < msg_len)))
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	contiguous = (stride == msg_len || count == 1) && msg_len % 4 == 0;
	pitch = contiguous ? msg_len : (msg_len + 15) & ~(size_t)15;
	chunk = b2g_chunk_units(pitch + 32, CHUNK_BYTES);
	b2g_lock();
	for (off = 0, c = 0; off < count; off += chunk, ++c)
	{
		const size_t n = count - off < chunk ? count - off : chunk;
		b2g_slot* sl = b2g_slot_get((int)c);
		const octet* src = (const octet*)msgs + off * stride;
		void *d_in, *d_out;
		if ((code = b2g_slot_buf(sl, 0, n * pitch, &d_in)) || (code = b2g_slot_buf(sl, 1, n * 32, &d_out)))
			goto done;
		if (msg_len)
		{
			if (contiguous)
				CU(cudaMemcpyAsync(d_in, src, n * msg_len, cudaMemcpyHostToDevice, sl->stream), "H2D(belt msgs)");
			else
				CU(cudaMemcpy2DAsync(d_in, pitch, src, n > 1 ? stride : msg_len, msg_len, n,
					cudaMemcpyHostToDevice, sl->stream), "H2D2D(belt msgs)");
		}
		if ((code = b2g_beltHashBatch_dev(d_out, d_in, msg_len, pitch, n, sl->stream)))
			goto done;
		CU(cudaMemcpyAsync(hashes + 32 * off, d_out, 32 * n, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt digests)");
	}
	code = sync_all();
done:
	if (code)
		sync_all();
	b2g_unlock();
	return code;
}

err_t beltHash(octet hash[32], const void* src, size_t count)
{
	if ((count && !src) || !hash)
		return ERR_BAD_INPUT;
	return beltHashBatch(hash, src, count, count, 1);
}

/* ---------------------------------------------------------------- belt-DWP (belt_dwp.c:250-330) */
/* One-shot AEAD on whole buffers: the data stay on the device between the CTR pass (belt.cu) and
   the tag pass (belt_dwp.cu). */
static err_t dwp_run(void* dest, octet mac_out[8], const void* src1, size_t count1, const void* src2,
	size_t count2, const octet key[], size_t len, const octet iv[16], const octet* expect_mac, int che)
{
	/* che: belt-CHE (belt_che.c:257-330) — LFSR counter and r = E_K(iv); else belt-DWP */
	err_t (*crypt)(void*, const void*, size_t, const u32*, const u32*, u64, void*) =
		che ? b2g_beltCHE_dev : b2g_beltCTR_dev;
	err_t (*tag)(void*, const void*, size_t, const void*, size_t, const u32*, const u32*, void*, void*) =
		che ? b2g_beltCHEMac_dev : b2g_beltDWPMac_dev;
	err_t code;
	belt_ctr_st st;
	b2g_slot* sl;
	void *d_crit, *d_open, *d_small;
	octet mac[8];
	if (!key_len_ok(len) || (count1 && (!src1 || !dest)) || (count2 && !src2) || !key || !iv)
		return ERR_BAD_INPUT;
	beltKeyExpand2(st.key, key, len);
	memcpy(st.ctr, iv, 16);
	if ((code = blocks_small(st.ctr, 1, st.key, 0)))   /* s = E_K(iv) */
		return code;
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, (count1 + 15) & ~(size_t)15, &d_crit)) ||
		(code = b2g_slot_buf(sl, 1, count2, &d_open)) || (code = b2g_slot_buf(sl, 2, 64, &d_small)))
		goto done;
	if (count2)
		CU(cudaMemcpyAsync(d_open, src2, count2, cudaMemcpyHostToDevice, sl->stream), "H2D(dwp open)");
	if (!expect_mac)
	{
		/* wrap: encrypt, then authenticate the ciphertext (belt_dwp.c:277-282). The critical data
		   move in chunks that alternate between the two streams — upload, encrypt (counter offset =
		   chunk offset) and download of one chunk overlap the neighbours' transfers in the other PCIe
		   direction — while the ciphertext stays resident for ONE tag launch over the whole buffer. */
		b2g_slot* s2 = b2g_slot_get(1);
		const size_t chunk = b2g_chunk_units(16, CHUNK_BYTES) * 16;
		size_t off, c;
		cudaEvent_t ev;
		for (off = 0, c = 0; off < count1; off += chunk, ++c)
		{
			const size_t n = count1 - off < chunk ? count1 - off : chunk;
			cudaStream_t stc = (c & 1) ? s2->stream : sl->stream;
			octet* d = (octet*)d_crit + off;
			CU(cudaMemcpyAsync(d, (const octet*)src1 + off, n, cudaMemcpyHostToDevice, stc), "H2D(dwp data)");
			if ((code = crypt(d, d, n, st.key, st.ctr, off / 16, stc)))
				goto done;
			CU(cudaMemcpyAsync((octet*)dest + off, d, n, cudaMemcpyDeviceToHost, stc), "D2H(dwp data)");
		}
		if (c > 1)
		{
			/* the tag pass (first stream) reads what the second stream encrypted */
			CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate(dwp)");
			if (cudaEventRecord(ev, s2->stream) != cudaSuccess || cudaStreamWaitEvent(sl->stream, ev, 0) != cudaSuccess)
			{
				code = b2g_check_launch("event(dwp)");
				cudaEventDestroy(ev);
				if (!code) code = ERR_B2G_CUDA;
				goto done;
			}
			cudaEventDestroy(ev);
		}
		if ((code = tag(d_small, d_crit, count1, d_open, count2, st.key, st.ctr,
				(octet*)d_small + 16, sl->stream)))
			goto done;
		CU(cudaMemcpyAsync(mac, d_small, 8, cudaMemcpyDeviceToHost, sl->stream), "D2H(dwp mac)");
		CU(cudaStreamSynchronize(sl->stream), "sync(dwp)");
		CU(cudaStreamSynchronize(s2->stream), "sync(dwp)");
		memcpy(mac_out, mac, 8);
	}
	else
	{
		/* unwrap: check the tag first; a wrong tag leaves dest untouched (belt_dwp.c:316-324) */
		if (count1)
			CU(cudaMemcpyAsync(d_crit, src1, count1, cudaMemcpyHostToDevice, sl->stream), "H2D(dwp data)");
		if ((code = tag(d_small, d_crit, count1, d_open, count2, st.key, st.ctr,
				(octet*)d_small + 16, sl->stream)))
			goto done;
		CU(cudaMemcpyAsync(mac, d_small, 8, cudaMemcpyDeviceToHost, sl->stream), "D2H(dwp mac)");
		CU(cudaStreamSynchronize(sl->stream), "sync(dwp)");
		if (memcmp(mac, expect_mac, 8) != 0)
		{
			code = ERR_BAD_MAC;
			goto done;
		}
		if (count1)
		{
			if ((code = crypt(d_crit, d_crit, count1, st.key, st.ctr, 0, sl->stream)))
				goto done;
			CU(cudaMemcpyAsync(dest, d_crit, count1, cudaMemcpyDeviceToHost, sl->stream), "D2H(dwp data)");
			CU(cudaStreamSynchronize(sl->stream), "sync(dwp)");
		}
	}
done:
	if (code && code != ERR_BAD_MAC)
		cudaStreamSynchronize(sl->stream), cudaStreamSynchronize(b2g_slot_get(1)->stream);
	b2g_unlock();
	return code;
}

err_t beltDWPWrap(void* dest, octet mac[8], const void* src1, size_t count1, const void* src2,
	size_t count2, const octet key[], size_t len, const octet iv[16])
{
	if (!mac)
		return ERR_BAD_INPUT;
	return dwp_run(dest, mac, src1, count1, src2, count2, key, len, iv, 0, 0);
}

err_t beltDWPUnwrap(void* dest, const void* src1, size_t count1, const void* src2, size_t count2,
	const octet mac[8], const octet key[], size_t len, const octet iv[16])
{
	octet unused[8];
	if (!mac)
		return ERR_BAD_INPUT;
	return dwp_run(dest, unused, src1, count1, src2, count2, key, len, iv, mac, 0);
}

err_t beltCHEWrap(void* dest, octet mac[8], const void* src1, size_t count1, const void* src2,
	size_t count2, const octet key[], size_t len, const octet iv[16])
{
	if (!mac)
		return ERR_BAD_INPUT;
	return dwp_run(dest, mac, src1, count1, src2, count2, key, len, iv, 0, 1);
}

err_t beltCHEUnwrap(void* dest, const void* src1, size_t count1, const void* src2, size_t count2,
	const octet mac[8], const octet key[], size_t len, const octet iv[16])
{
	octet unused[8];
	if (!mac)
		return ERR_BAD_INPUT;
	return dwp_run(dest, unused, src1, count1, src2, count2, key, len, iv, mac, 1);
}
