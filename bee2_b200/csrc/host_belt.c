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
	err_t code;
	B2G_SMALL_V(16, beltBlockEncr, block, key);
	if ((code = blocks_small(block, 1, key, 0))) B2G_FAIL_V(code, beltBlockEncr, block, key);
}
void beltBlockEncr2(u32 block[4], const u32 key[8])
{
	err_t code;
	B2G_SMALL_V(16, beltBlockEncr2, block, key);
	if ((code = blocks_small(block, 1, key, 0))) B2G_FAIL_V(code, beltBlockEncr2, block, key);
}
void beltBlockEncr3(u32* a, u32* b, u32* c, u32* d, const u32 key[8])
{
	u32 t[4];
	err_t code;
	B2G_SMALL_V(16, beltBlockEncr3, a, b, c, d, key);
	t[0] = *a, t[1] = *b, t[2] = *c, t[3] = *d;
	if ((code = blocks_small(t, 1, key, 0))) B2G_FAIL_V(code, beltBlockEncr3, a, b, c, d, key);
	*a = t[0], *b = t[1], *c = t[2], *d = t[3];
}
void beltBlockDecr(octet block[16], const u32 key[8])
{
	err_t code;
	B2G_SMALL_V(16, beltBlockDecr, block, key);
	if ((code = blocks_small(block, 1, key, 1))) B2G_FAIL_V(code, beltBlockDecr, block, key);
}
void beltBlockDecr2(u32 block[4], const u32 key[8])
{
	err_t code;
	B2G_SMALL_V(16, beltBlockDecr2, block, key);
	if ((code = blocks_small(block, 1, key, 1))) B2G_FAIL_V(code, beltBlockDecr2, block, key);
}
void beltBlockDecr3(u32* a, u32* b, u32* c, u32* d, const u32 key[8])
{
	u32 t[4];
	err_t code;
	B2G_SMALL_V(16, beltBlockDecr3, a, b, c, d, key);
	t[0] = *a, t[1] = *b, t[2] = *c, t[3] = *d;
	if ((code = blocks_small(t, 1, key, 1))) B2G_FAIL_V(code, beltBlockDecr3, a, b, c, d, key);
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
	B2G_PREFLIGHT_V(beltECBStepE, buf, count, state);
	err_t code = ecb_step((octet*)buf, count, ((belt_ecb_st*)state)->key, 0);
	if (code) b2g_die("beltECBStepE", code);
}
void beltECBStepD(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltECBStepD, buf, count, state);
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
	B2G_SMALL_R(count, beltECBEncr, dest, src, count, key, len);
	return ecb_oneshot(dest, src, count, key, len, 0);
}
err_t beltECBDecr(void* dest, const void* src, size_t count, const octet key[], size_t len)
{
	B2G_SMALL_R(count, beltECBDecr, dest, src, count, key, len);
	return ecb_oneshot(dest, src, count, key, len, 1);
}

static err_t ecb_batch_1(void* blocks, const octet* keys32, size_t count)
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

typedef struct
{
	octet* blocks;
	const octet* keys32;
} ecbk_args;
static u32 ecbk_shard(void* arg, size_t first, size_t n)
{
	const ecbk_args* a = (const ecbk_args*)arg;
	return ecb_batch_1(a->blocks + 16 * first, a->keys32 + 32 * first, n);
}
err_t beltECBEncrBatch(void* blocks, const octet* keys32, size_t count)
{
	ecbk_args a = {(octet*)blocks, keys32};
	if (b2g_device_count() <= 1 || count < ((size_t)1 << 20))
		return ecb_batch_1(blocks, keys32, count);
	if (!blocks || !keys32)
		return ERR_BAD_INPUT;
	return b2g_fanout(count, (size_t)1 << 19, ecbk_shard, &a);
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
	B2G_PREFLIGHT_V(beltCTRStart, state, key, len, iv);
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
static err_t ctr_run1(octet* dest, const octet* src, size_t count, const u32 key[8], const u32 ctr0[4],
	octet last_ks[16], u64 first_block)
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
		if ((code = b2g_beltCTR_dev(d, src ? d : 0, padded, key, ctr0, first_block + off / 16, sl->stream)))
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

/* in-process multi-device mode: device g produces counter blocks [first, first + n) of the stream */
typedef struct
{
	octet* dest;
	const octet* src;
	size_t count;
	const u32 *key, *ctr0;
	octet* last_ks;
} ctr_args;
static u32 ctr_shard(void* arg, size_t first, size_t n)
{
	const ctr_args* a = (const ctr_args*)arg;
	const size_t off = 16 * first;
	const size_t bytes = a->count - off < 16 * n ? a->count - off : 16 * n;
	return ctr_run1(a->dest + off, a->src ? a->src + off : 0, bytes, a->key, a->ctr0,
		off + bytes == a->count ? a->last_ks : 0, first);
}
static err_t ctr_run(octet* dest, const octet* src, size_t count, const u32 key[8], const u32 ctr0[4],
	octet last_ks[16])
{
	ctr_args a = {dest, src, count, key, ctr0, last_ks};
	if (b2g_device_count() <= 1 || count < ((size_t)32 << 20))
		return ctr_run1(dest, src, count, key, ctr0, last_ks, 0);
	return b2g_fanout((count + 15) / 16, (size_t)1 << 20, ctr_shard, &a);
}

void beltCTRStepE(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltCTRStepE, buf, count, state);
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
	B2G_SMALL_R(count, beltCTR, dest, src, count, key, len, iv);
	beltKeyExpand2(st.key, key, len);
	memcpy(st.ctr, iv, 16);
	if (!(code = blocks_small(st.ctr, 1, st.key, 0)))
		code = ctr_run((octet*)dest, (const octet*)src, count, st.key, st.ctr, 0);
	b2g_wipe(&st, sizeof st);
	return code;
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
	B2G_SMALL_R(count, beltHash, hash, src, count);
	return beltHashBatch(hash, src, count, count, 1);
}

/* ---------------------------------------------------------------- belt-hash, streaming forms
   (belt_hash.c:27-158). State layout = belt_hash_st without the compression stack. The chain of
   sigma-compressions runs on the device (b2g_beltHashStep_dev); the host buffers the ragged tail and
   counts bits. */
typedef struct
{
	u32 ls[8];             /* [4] length in bits || [4] s */
	u32 s1[4];
	u32 h[8];
	u32 h1[8];
	octet block[32];
	size_t filled;
} belt_hash_st;

size_t beltHash_keep(void) { return sizeof(belt_hash_st); }

void beltHashStart(void* state)
{
	belt_hash_st* st = (belt_hash_st*)state;
	memset(st->ls, 0, sizeof st->ls);
	memcpy(st->h, beltH(), 32);
	st->filled = 0;
}

/* sh = s || h (12 words) <- after head32 (optional) || data[0..nbytes) (whole blocks) [|| length block] */
static err_t hash_chain(u32 s[4], u32 h[8], const octet* head32, const octet* data, size_t nbytes,
	int final, const u32 len[4])
{
	err_t code;
	b2g_slot* sl;
	void *d, *d_state;
	u32 sh[12];
	const size_t total = nbytes + (head32 ? 32 : 0);
	if (!total && !final)
		return ERR_OK;
	if ((code = b2g_ensure_device()))
		return code;
	memcpy(sh, s, 16), memcpy(sh + 4, h, 32);
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, total ? total : 32, &d)) || (code = b2g_slot_buf(sl, 2, 64, &d_state)))
		goto done;
	if (head32)
		CU(cudaMemcpyAsync(d, head32, 32, cudaMemcpyHostToDevice, sl->stream), "H2D(hash head)");
	if (nbytes)
		CU(cudaMemcpyAsync((octet*)d + (head32 ? 32 : 0), data, nbytes, cudaMemcpyHostToDevice, sl->stream), "H2D(hash data)");
	CU(cudaMemcpyAsync(d_state, sh, 48, cudaMemcpyHostToDevice, sl->stream), "H2D(hash state)");
	if ((code = b2g_beltHashStep_dev(d_state, d, total / 32, final, len, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(sh, d_state, 48, cudaMemcpyDeviceToHost, sl->stream), "D2H(hash state)");
	CU(cudaStreamSynchronize(sl->stream), "sync(hash)");
	memcpy(s, sh, 16), memcpy(h, sh + 4, 32);
done:
	if (code)
		cudaStreamSynchronize(sl->stream);
	b2g_unlock();
	return code;
}

void beltHashStepH(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltHashStepH, buf, count, state);
	belt_hash_st* st = (belt_hash_st*)state;
	const octet* p = (const octet*)buf;
	const octet* head = 0;
	size_t full;
	err_t code;
	/* 128-bit length in bits += 8 count (beltBlockAddBitSizeU32, belt_lcl.c:25-50) */
	{
		u64 lo = (u64)st->ls[0] | (u64)st->ls[1] << 32, hi = (u64)st->ls[2] | (u64)st->ls[3] << 32;
		const u64 add_lo = (u64)count << 3, add_hi = (u64)count >> 61;
		lo += add_lo;
		hi += add_hi + (lo < add_lo);
		st->ls[0] = (u32)lo, st->ls[1] = (u32)(lo >> 32), st->ls[2] = (u32)hi, st->ls[3] = (u32)(hi >> 32);
	}
	if (st->filled)
	{
		if (count < 32 - st->filled)
		{
			memcpy(st->block + st->filled, p, count);
			st->filled += count;
			return;
		}
		memcpy(st->block + st->filled, p, 32 - st->filled);
		count -= 32 - st->filled, p += 32 - st->filled;
		st->filled = 0;
		head = st->block;
	}
	full = count & ~(size_t)31;
	if ((code = hash_chain(st->ls + 4, st->h, head, p, full, 0, st->ls)))
		b2g_die("beltHashStepH", code);
	p += full, count -= full;
	if (count)
		memcpy(st->block, p, st->filled = count);
}

/* h1 <- the digest of what has been absorbed; s and h stay as they are (belt_hash.c:103-121) */
static void hash_step_g(belt_hash_st* st, const char* who)
{
	octet last[32];
	err_t code;
	memcpy(st->s1, st->ls + 4, 16);
	memcpy(st->h1, st->h, 32);
	if (st->filled)
	{
		memcpy(last, st->block, st->filled);
		memset(last + st->filled, 0, 32 - st->filled);
	}
	/* the padded block moves s, the length block is len || that s; then s is restored */
	if ((code = hash_chain(st->ls + 4, st->h1, st->filled ? last : 0, 0, 0, 1, st->ls)))
		b2g_die(who, code);
	memcpy(st->ls + 4, st->s1, 16);
}

void beltHashStepG(octet hash[32], void* state)
{
	B2G_PREFLIGHT_V(beltHashStepG, hash, state);
	belt_hash_st* st = (belt_hash_st*)state;
	hash_step_g(st, "beltHashStepG");
	memcpy(hash, st->h1, 32);
}
void beltHashStepG2(octet hash[], size_t hash_len, void* state)
{
	B2G_PREFLIGHT_V(beltHashStepG2, hash, hash_len, state);
	belt_hash_st* st = (belt_hash_st*)state;
	hash_step_g(st, "beltHashStepG2");
	memcpy(hash, st->h1, hash_len < 32 ? hash_len : 32);
}
bool_t beltHashStepV(const octet hash[32], void* state)
{
	B2G_PREFLIGHT_R(beltHashStepV, hash, state);
	belt_hash_st* st = (belt_hash_st*)state;
	hash_step_g(st, "beltHashStepV");
	return b2g_ct_eq(hash, st->h1, 32);
}
bool_t beltHashStepV2(const octet hash[], size_t hash_len, void* state)
{
	B2G_PREFLIGHT_R(beltHashStepV2, hash, hash_len, state);
	belt_hash_st* st = (belt_hash_st*)state;
	hash_step_g(st, "beltHashStepV2");
	return b2g_ct_eq(hash, st->h1, hash_len < 32 ? hash_len : 32);
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
		if (!b2g_ct_eq(mac, expect_mac, 8))
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

/* ---------------------------------------------------------------- belt-DWP / belt-CHE, streaming forms
   (belt_dwp.c:27-207, belt_che.c:27-239). State layouts follow the reference's belt_dwp_st /
   belt_che_st field for field (without the multiplication stack, which lives on the device here). The
   block chain t <- (t ^ B) * r of every StepI / StepA / StepG call runs on the device
   (b2g_beltPolyAbsorb_dev); the host side only buffers ragged tails and counts bits. */
typedef struct
{
	belt_ctr_st ctr[1];
	u32 r[4], t[4], t1[4];
	u32 len[4];            /* bits of open data (low half) || bits of critical data (high half) */
	octet block[16];
	size_t filled;
} belt_dwp_st;

typedef struct
{
	u32 key[8];
	u32 s[4];
	u32 r[4], t[4], t1[4];
	u32 len[4];
	octet block[16];       /* pending authenticated data */
	octet block1[16];      /* keystream block */
	size_t filled;
	size_t reserved;
} belt_che_st;

/* the part both states share once located: r, t, t1, len, block, filled */
typedef struct { u32 *r, *t, *t1, *len; octet* block; size_t* filled; const u32* key; } aead_view;

static aead_view dwp_view(belt_dwp_st* st)
{
	aead_view v = {st->r, st->t, st->t1, st->len, st->block, &st->filled, st->ctr->key};
	return v;
}
static aead_view che_view(belt_che_st* st)
{
	aead_view v = {st->r, st->t, st->t1, st->len, st->block, &st->filled, st->key};
	return v;
}

/* t <- Horner(t; head16 (optional) || data[0..nbytes)), nbytes a multiple of 16 */
static err_t poly_absorb(u32 t[4], const u32 r[4], const octet* head16, const octet* data, size_t nbytes)
{
	err_t code;
	b2g_slot* sl;
	void *d, *d_small;
	const size_t total = nbytes + (head16 ? 16 : 0);
	if (!total)
		return ERR_OK;
	if ((code = b2g_ensure_device()))
		return code;
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, total, &d)) || (code = b2g_slot_buf(sl, 2, 64, &d_small)))
		goto done;
	if (head16)
		CU(cudaMemcpyAsync(d, head16, 16, cudaMemcpyHostToDevice, sl->stream), "H2D(poly head)");
	if (nbytes)
		CU(cudaMemcpyAsync((octet*)d + (head16 ? 16 : 0), data, nbytes, cudaMemcpyHostToDevice, sl->stream), "H2D(poly data)");
	if ((code = b2g_beltPolyAbsorb_dev(d_small, d, total, r, t, (octet*)d_small + 16, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(t, d_small, 16, cudaMemcpyDeviceToHost, sl->stream), "D2H(poly t)");
	CU(cudaStreamSynchronize(sl->stream), "sync(poly)");
done:
	if (code)
		cudaStreamSynchronize(sl->stream);
	b2g_unlock();
	return code;
}

/* half <- half + 8 count mod 2^64 (beltHalfBlockAddBitSizeW, belt_lcl.c:53-98) */
static void add_bit_size(u32 half[2], size_t count)
{
	u64 v = (u64)half[0] | (u64)half[1] << 32;
	v += (u64)count << 3;
	half[0] = (u32)v, half[1] = (u32)(v >> 32);
}

/* the common tail of StepI / StepA (belt_dwp.c:81-113, :131-163): complete a pending block, run the
   whole blocks, keep the ragged rest */
static void aead_absorb(aead_view v, const void* buf, size_t count, const char* who)
{
	const octet* p = (const octet*)buf;
	const octet* head = 0;
	size_t full;
	err_t code;
	if (*v.filled)
	{
		if (count < 16 - *v.filled)
		{
			memcpy(v.block + *v.filled, p, count);
			*v.filled += count;
			return;
		}
		memcpy(v.block + *v.filled, p, 16 - *v.filled);
		count -= 16 - *v.filled, p += 16 - *v.filled;
		*v.filled = 0;
		head = v.block;
	}
	full = count & ~(size_t)15;
	if ((code = poly_absorb(v.t, v.r, head, p, full)))
		b2g_die(who, code);
	p += full, count -= full;
	if (count)
		memcpy(v.block, p, *v.filled = count);
}

static void aead_step_i(aead_view v, const void* buf, size_t count, const char* who)
{
	add_bit_size(v.len, count);
	aead_absorb(v, buf, count, who);
}

static void aead_step_a(aead_view v, const void* buf, size_t count, const char* who)
{
	/* first non-empty fragment of critical data: close the open data with zeros (belt_dwp.c:121-131) */
	if (count && v.len[2] == 0 && v.len[3] == 0 && *v.filled)
	{
		err_t code;
		memset(v.block + *v.filled, 0, 16 - *v.filled);
		if ((code = poly_absorb(v.t, v.r, v.block, 0, 0)))
			b2g_die(who, code);
		*v.filled = 0;
	}
	add_bit_size(v.len + 2, count);
	aead_absorb(v, buf, count, who);
}

/* t1 <- E_K(((t [^ padded block] * r) ^ len) * r); t itself is kept (belt_dwp.c:172-196) */
static void aead_step_g(aead_view v, const char* who)
{
	octet tail[32];
	size_t n = 0;
	err_t code;
	if (*v.filled)
	{
		memcpy(tail, v.block, *v.filled);
		memset(tail + *v.filled, 0, 16 - *v.filled);
		n = 16;
	}
	memcpy(tail + n, v.len, 16), n += 16;
	memcpy(v.t1, v.t, 16);
	if ((code = poly_absorb(v.t1, v.r, 0, tail, n)) || (code = blocks_small(v.t1, 1, v.key, 0)))
		b2g_die(who, code);
}

size_t beltDWP_keep(void) { return sizeof(belt_dwp_st); }

void beltDWPStart(void* state, const octet key[], size_t len, const octet iv[16])
{
	B2G_PREFLIGHT_V(beltDWPStart, state, key, len, iv);
	belt_dwp_st* st = (belt_dwp_st*)state;
	err_t code;
	beltCTRStart(st->ctr, key, len, iv);
	/* r <- E_K(s), t <- beltH()[0..16) (belt_dwp.c:50-60) */
	memcpy(st->r, st->ctr->ctr, 16);
	if ((code = blocks_small(st->r, 1, st->ctr->key, 0)))
		b2g_die("beltDWPStart", code);
	memcpy(st->t, beltH(), 16);
	memset(st->len, 0, sizeof st->len);
	st->filled = 0;
}
void beltDWPStepE(void* buf, size_t count, void* state) { 	B2G_PREFLIGHT_V(beltDWPStepE, buf, count, state);
beltCTRStepE(buf, count, state); }
void beltDWPStepD(void* buf, size_t count, void* state) { 	B2G_PREFLIGHT_V(beltDWPStepD, buf, count, state);
beltCTRStepE(buf, count, state); }
void beltDWPStepI(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltDWPStepI, buf, count, state);
	aead_step_i(dwp_view((belt_dwp_st*)state), buf, count, "beltDWPStepI");
}
void beltDWPStepA(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltDWPStepA, buf, count, state);
	aead_step_a(dwp_view((belt_dwp_st*)state), buf, count, "beltDWPStepA");
}
void beltDWPStepG(octet mac[8], void* state)
{
	B2G_PREFLIGHT_V(beltDWPStepG, mac, state);
	belt_dwp_st* st = (belt_dwp_st*)state;
	aead_step_g(dwp_view(st), "beltDWPStepG");
	memcpy(mac, st->t1, 8);
}
bool_t beltDWPStepV(const octet mac[8], void* state)
{
	B2G_PREFLIGHT_R(beltDWPStepV, mac, state);
	belt_dwp_st* st = (belt_dwp_st*)state;
	aead_step_g(dwp_view(st), "beltDWPStepV");
	return b2g_ct_eq(mac, st->t1, 8);
}

/* ---- belt-CHE: the counter is the LFSR s <- s x ^ 1 over GF(2^128) (belt_che.c:86-88, belt_lcl.c:99-108) */
/* c <- a b mod x^128 + x^7 + x^2 + x + 1, bit i of the little-endian 128-bit integer = coefficient of x^i */
static void gf128_mul_host(u32 c[4], const u32 a[4], const u32 b[4])
{
	u32 acc[4] = {0, 0, 0, 0}, v[4];
	int i, k;
	memcpy(v, a, 16);
	for (i = 0; i < 128; ++i)
	{
		if (b[i >> 5] >> (i & 31) & 1)
			for (k = 0; k < 4; ++k) acc[k] ^= v[k];
		{
			const u32 top = v[3] >> 31;
			v[3] = v[3] << 1 | v[2] >> 31, v[2] = v[2] << 1 | v[1] >> 31, v[1] = v[1] << 1 | v[0] >> 31;
			v[0] = v[0] << 1 ^ (top ? 0x87u : 0u);
		}
	}
	memcpy(c, acc, 16);
}
/* s <- the LFSR state n steps later: the map s -> s x ^ 1 iterated n times is s -> s x^n ^ c_n; maps
   compose as (A, c) o (B, d) = (A B, A d ^ c), so n is consumed by square-and-multiply */
static void che_advance(u32 s[4], u64 n)
{
	u32 A[4] = {1, 0, 0, 0}, c[4] = {0, 0, 0, 0};     /* identity */
	u32 B[4] = {2, 0, 0, 0}, d[4] = {1, 0, 0, 0};     /* one step: s x ^ 1 */
	u32 t[4];
	int k;
	for (; n; n >>= 1)
	{
		if (n & 1)
		{
			/* (A, c) <- (B, d) o (A, c) */
			gf128_mul_host(t, B, c);
			for (k = 0; k < 4; ++k) c[k] = t[k] ^ d[k];
			gf128_mul_host(A, A, B);
		}
		/* (B, d) <- (B, d) o (B, d) */
		gf128_mul_host(t, B, d);
		for (k = 0; k < 4; ++k) d[k] ^= t[k];
		gf128_mul_host(B, B, B);
	}
	gf128_mul_host(t, A, s);
	for (k = 0; k < 4; ++k) s[k] = t[k] ^ c[k];
}

size_t beltCHE_keep(void) { return sizeof(belt_che_st); }

void beltCHEStart(void* state, const octet key[], size_t len, const octet iv[16])
{
	B2G_PREFLIGHT_V(beltCHEStart, state, key, len, iv);
	belt_che_st* st = (belt_che_st*)state;
	err_t code;
	beltKeyExpand2(st->key, key, len);
	/* r <- s <- E_K(iv), t <- beltH()[0..16) (belt_che.c:52-62) */
	memcpy(st->r, iv, 16);
	if ((code = blocks_small(st->r, 1, st->key, 0)))
		b2g_die("beltCHEStart", code);
	memcpy(st->s, st->r, 16);
	memcpy(st->t, beltH(), 16);
	memset(st->len, 0, sizeof st->len);
	st->reserved = 0;
	st->filled = 0;
}

void beltCHEStepE(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltCHEStepE, buf, count, state);
	belt_che_st* st = (belt_che_st*)state;
	octet* p = (octet*)buf;
	b2g_slot* sl;
	void* d;
	size_t i, padded, rem;
	err_t code;
	/* reserve of keystream octets left from the previous call (belt_che.c:70-82) */
	if (st->reserved)
	{
		const size_t take = st->reserved < count ? st->reserved : count;
		for (i = 0; i < take; ++i)
			p[i] ^= st->block1[16 - st->reserved + i];
		st->reserved -= take, p += take, count -= take;
	}
	if (!count)
		return;
	if ((code = b2g_ensure_device()))
		b2g_die("beltCHEStepE", code);
	padded = (count + 15) & ~(size_t)15, rem = count % 16;
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, padded, &d)))
		goto done;
	if (rem)
		CU(cudaMemsetAsync((octet*)d + padded - 16, 0, 16, sl->stream), "memset(belt che)");
	CU(cudaMemcpyAsync(d, p, count, cudaMemcpyHostToDevice, sl->stream), "H2D(belt che)");
	if ((code = b2g_beltCHE_dev(d, d, padded, st->key, st->s, 0, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(p, d, count, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt che)");
	if (rem)
		CU(cudaMemcpyAsync(st->block1, (octet*)d + padded - 16, 16, cudaMemcpyDeviceToHost, sl->stream), "D2H(belt che tail)");
	CU(cudaStreamSynchronize(sl->stream), "sync(belt che)");
done:
	if (code)
		cudaStreamSynchronize(sl->stream);
	b2g_unlock();
	if (code)
		b2g_die("beltCHEStepE", code);
	che_advance(st->s, padded / 16);
	if (rem)
		st->reserved = 16 - rem;   /* block1[rem..16) = unused keystream of the last block (its data part was 0) */
}
void beltCHEStepD(void* buf, size_t count, void* state) { 	B2G_PREFLIGHT_V(beltCHEStepD, buf, count, state);
beltCHEStepE(buf, count, state); }
void beltCHEStepI(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltCHEStepI, buf, count, state);
	aead_step_i(che_view((belt_che_st*)state), buf, count, "beltCHEStepI");
}
void beltCHEStepA(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(beltCHEStepA, buf, count, state);
	aead_step_a(che_view((belt_che_st*)state), buf, count, "beltCHEStepA");
}
void beltCHEStepG(octet mac[8], void* state)
{
	B2G_PREFLIGHT_V(beltCHEStepG, mac, state);
	belt_che_st* st = (belt_che_st*)state;
	aead_step_g(che_view(st), "beltCHEStepG");
	memcpy(mac, st->t1, 8);
}
bool_t beltCHEStepV(const octet mac[8], void* state)
{
	B2G_PREFLIGHT_R(beltCHEStepV, mac, state);
	belt_che_st* st = (belt_che_st*)state;
	aead_step_g(che_view(st), "beltCHEStepV");
	return b2g_ct_eq(mac, st->t1, 8);
}

err_t beltDWPWrap(void* dest, octet mac[8], const void* src1, size_t count1, const void* src2,
	size_t count2, const octet key[], size_t len, const octet iv[16])
{
	if (!mac)
		return ERR_BAD_INPUT;
	B2G_SMALL_R(count1 + count2, beltDWPWrap, dest, mac, src1, count1, src2, count2, key, len, iv);
	return dwp_run(dest, mac, src1, count1, src2, count2, key, len, iv, 0, 0);
}

err_t beltDWPUnwrap(void* dest, const void* src1, size_t count1, const void* src2, size_t count2,
	const octet mac[8], const octet key[], size_t len, const octet iv[16])
{
	octet unused[8];
	if (!mac)
		return ERR_BAD_INPUT;
	B2G_SMALL_R(count1 + count2, beltDWPUnwrap, dest, src1, count1, src2, count2, mac, key, len, iv);
	return dwp_run(dest, unused, src1, count1, src2, count2, key, len, iv, mac, 0);
}

err_t beltCHEWrap(void* dest, octet mac[8], const void* src1, size_t count1, const void* src2,
	size_t count2, const octet key[], size_t len, const octet iv[16])
{
	if (!mac)
		return ERR_BAD_INPUT;
	B2G_SMALL_R(count1 + count2, beltCHEWrap, dest, mac, src1, count1, src2, count2, key, len, iv);
	return dwp_run(dest, mac, src1, count1, src2, count2, key, len, iv, 0, 1);
}

err_t beltCHEUnwrap(void* dest, const void* src1, size_t count1, const void* src2, size_t count2,
	const octet mac[8], const octet key[], size_t len, const octet iv[16])
{
	octet unused[8];
	if (!mac)
		return ERR_BAD_INPUT;
	B2G_SMALL_R(count1 + count2, beltCHEUnwrap, dest, src1, count1, src2, count2, mac, key, len, iv);
	return dwp_run(dest, unused, src1, count1, src2, count2, key, len, iv, mac, 1);
}
