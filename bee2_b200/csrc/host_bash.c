/*
 * host_bash.c — host side (C) of the bash path: the reference's bash.h surface
 * (bash_hash.c:25-137, bash_f.c:26-44) and the host-pointer batch entry points, all
 * implemented by staging buffers to the GPU and launching the sm_100a kernels in bash.cu.
 * Buffer bookkeeping only — every bash-f evaluation happens on the device.
 */
#include "engine.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

const char bash_platform[] = "BASH_CUDA_SM100A";

/* same layout as the reference's bash_hash_st (bash_hash.c:25-31); bashF_deep() == 0 */
typedef struct
{
	octet s[192];
	octet s1[192];
	size_t buf_len;
	size_t pos;
} bash_hash_st;

size_t bashF_deep(void) { return 0; }
size_t bashHash_keep(void) { return sizeof(bash_hash_st); }

#define CU(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { code = b2g_cuda_fail(e_, what); goto done; } } while (0)

err_t bashFBatch(octet* blocks, size_t count)
{
	err_t code;
	b2g_slot* sl;
	void* d;
	if (count && !blocks)
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, 192 * count, &d)))
		goto done;
	CU(cudaMemcpyAsync(d, blocks, 192 * count, cudaMemcpyHostToDevice, sl->stream), "H2D(bashF)");
	if ((code = b2g_bashFBatch_dev(d, count, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(blocks, d, 192 * count, cudaMemcpyDeviceToHost, sl->stream), "D2H(bashF)");
	CU(cudaStreamSynchronize(sl->stream), "sync(bashF)");
done:
	b2g_unlock();
	return code;
}

void bashF(octet block[192], void* stack)
{
	err_t code;
	B2G_SMALL_V(192, bashF, block, stack);
	if ((code = bashFBatch(block, 1)))
		B2G_FAIL_V(code, bashF, block, stack);
}

static err_t hash_batch_1(octet* hashes, size_t l, const void* msgs, size_t msg_len, size_t stride,
	size_t count)
{
	err_t code;
	size_t hl, pitch, chunk, done_units, c;
	int contiguous;
	if (l == 0 || l % 16 != 0 || l > 256)
		return ERR_BAD_PARAMS;
	if (count && (!hashes || (msg_len && !msgs) || (count > 1 && stride < msg_len)))
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	hl = l / 4;
	contiguous = (stride == msg_len || count == 1) && msg_len % 16 == 0;
	pitch = contiguous ? msg_len : (msg_len + 15) & ~(size_t)15;
	chunk = b2g_chunk_units(pitch + hl, (size_t)64 << 20);
	b2g_lock();
	for (done_units = 0, c = 0; done_units < count; done_units += chunk, ++c)
	{
		const size_t n = count - done_units < chunk ? count - done_units : chunk;
		b2g_slot* sl = b2g_slot_get((int)c);
		const octet* src = (const octet*)msgs + done_units * stride;
		void *d_in, *d_out;
		if ((code = b2g_slot_buf(sl, 0, n * pitch, &d_in)) || (code = b2g_slot_buf(sl, 1, n * hl, &d_out)))
			goto done;
		if (msg_len)
		{
			if (contiguous)
				CU(cudaMemcpyAsync(d_in, src, n * msg_len, cudaMemcpyHostToDevice, sl->stream), "H2D(bash msgs)");
			else
				CU(cudaMemcpy2DAsync(d_in, pitch, src, n > 1 ? stride : msg_len, msg_len, n,
					cudaMemcpyHostToDevice, sl->stream), "H2D2D(bash msgs)");
		}
		if ((code = b2g_bashHashBatch_dev(d_out, l, d_in, msg_len, pitch, n, sl->stream)))
			goto done;
		CU(cudaMemcpyAsync(hashes + done_units * hl, d_out, n * hl, cudaMemcpyDeviceToHost, sl->stream), "D2H(bash digests)");
	}
	for (c = 0; c < B2G_NSLOT; ++c)
		CU(cudaStreamSynchronize(b2g_slot_get((int)c)->stream), "sync(bash)");
done:
	if (code)
		for (c = 0; c < B2G_NSLOT; ++c)
			cudaStreamSynchronize(b2g_slot_get((int)c)->stream);
	b2g_unlock();
	return code;
}

/* in-process multi-device mode: contiguous shares of the messages, one device each */
typedef struct
{
	octet* hashes;
	size_t l;
	const octet* msgs;
	size_t msg_len, stride;
} hashb_args;
static u32 hashb_shard(void* arg, size_t first, size_t n)
{
	const hashb_args* a = (const hashb_args*)arg;
	return hash_batch_1(a->hashes + first * (a->l / 4), a->l, a->msgs + first * a->stride, a->msg_len, a->stride, n);
}
err_t bashHashBatch(octet* hashes, size_t l, const void* msgs, size_t msg_len, size_t stride,
	size_t count)
{
	hashb_args a = {hashes, l, (const octet*)msgs, msg_len, stride};
	/* shares of at least ~16 MiB of message data */
	const size_t grain = ((size_t)16 << 20) / (msg_len ? msg_len : 1) + 256;
	if (b2g_device_count() <= 1 || count < 2 * grain)
		return hash_batch_1(hashes, l, msgs, msg_len, stride, count);
	if (l == 0 || l % 16 != 0 || l > 256)
		return ERR_BAD_PARAMS;
	if (!hashes || (msg_len && !msgs) || stride < msg_len)
		return ERR_BAD_INPUT;
	return b2g_fanout(count, grain, hashb_shard, &a);
}

/* ragged batch: message i = data[offsets[i] .. offsets[i] + lens[i]) — many files in one buffer */
err_t bashHashBatchV(octet* hashes, size_t l, const void* data, size_t data_len, const u64* offsets,
	const u64* lens, size_t count)
{
	err_t code;
	b2g_slot *s0, *s1;
	void *d_data, *d_off, *d_len, *d_out;
	size_t i, hl;
	if (l == 0 || l % 16 != 0 || l > 256)
		return ERR_BAD_PARAMS;
	if (count && (!hashes || !offsets || !lens || (data_len && !data)))
		return ERR_BAD_INPUT;
	for (i = 0; i < count; ++i)
		if (offsets[i] > data_len || lens[i] > data_len - offsets[i])
			return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	hl = l / 4;
	b2g_lock();
	s0 = b2g_slot_get(0), s1 = b2g_slot_get(1);
	if ((code = b2g_slot_buf(s0, 0, data_len, &d_data)) || (code = b2g_slot_buf(s0, 1, 8 * count, &d_off)) ||
		(code = b2g_slot_buf(s0, 2, 8 * count, &d_len)) || (code = b2g_slot_buf(s1, 0, hl * count, &d_out)))
		goto done;
	if (data_len)
		CU(cudaMemcpyAsync(d_data, data, data_len, cudaMemcpyHostToDevice, s0->stream), "H2D(bash data)");
	CU(cudaMemcpyAsync(d_off, offsets, 8 * count, cudaMemcpyHostToDevice, s0->stream), "H2D(bash offsets)");
	CU(cudaMemcpyAsync(d_len, lens, 8 * count, cudaMemcpyHostToDevice, s0->stream), "H2D(bash lens)");
	if ((code = b2g_bashHashBatchV_dev(d_out, l, d_data, d_off, d_len, count, s0->stream)))
		goto done;
	CU(cudaMemcpyAsync(hashes, d_out, hl * count, cudaMemcpyDeviceToHost, s0->stream), "D2H(bash digests)");
	CU(cudaStreamSynchronize(s0->stream), "sync(bash ragged)");
done:
	if (code)
		cudaStreamSynchronize(s0->stream);
	b2g_unlock();
	return code;
}

err_t bashHash(octet hash[], size_t l, const void* src, size_t count)
{
	if (l == 0 || l % 16 != 0 || l > 256)
		return ERR_BAD_PARAMS;
	if ((count && !src) || !hash)
		return ERR_BAD_INPUT;
	B2G_SMALL_R(count, bashHash, hash, l, src, count);
	return bashHashBatch(hash, l, src, count, count, 1);
}

void bashHashStart(void* state, size_t l)
{
	bash_hash_st* st = (bash_hash_st*)state;
	memset(st->s, 0, sizeof st->s);
	st->s[192 - 8] = (octet)(l / 4);
	st->buf_len = 192 - l / 2;
	st->pos = 0;
}

/* absorb: run F on the (already merged) state, then `nfull` further rate blocks from buf */
static err_t sponge_absorb(bash_hash_st* st, const octet* buf, size_t nfull)
{
	err_t code;
	b2g_slot* sl;
	void *d_state, *d_msg;
	const size_t bytes = nfull * st->buf_len, l = 2 * (192 - st->buf_len);
	if ((code = b2g_ensure_device()))
		return code;
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, bytes, &d_msg)) || (code = b2g_slot_buf(sl, 1, 192, &d_state)))
		goto done;
	CU(cudaMemcpyAsync(d_state, st->s, 192, cudaMemcpyHostToDevice, sl->stream), "H2D(bash state)");
	if (bytes)
		CU(cudaMemcpyAsync(d_msg, buf, bytes, cudaMemcpyHostToDevice, sl->stream), "H2D(bash data)");
	if ((code = b2g_bashSponge_dev(d_msg, bytes, bytes, 1, l, d_state, d_state, 0, 0, 1, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(st->s, d_state, 192, cudaMemcpyDeviceToHost, sl->stream), "D2H(bash state)");
	CU(cudaStreamSynchronize(sl->stream), "sync(bash absorb)");
done:
	b2g_unlock();
	return code;
}

void bashHashStepH(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashHashStepH, buf, count, state);
	bash_hash_st* st = (bash_hash_st*)state;
	const octet* p = (const octet*)buf;
	size_t nfull;
	const size_t pos0 = st->pos;
	err_t code;
	if (count < st->buf_len - st->pos)
	{
		memcpy(st->s + st->pos, p, count);
		st->pos += count;
		return;
	}
	memcpy(st->s + st->pos, p, st->buf_len - st->pos);
	p += st->buf_len - st->pos, count -= st->buf_len - st->pos;
	nfull = count / st->buf_len;
	if ((code = sponge_absorb(st, p, nfull)))
	{
		/* the device result is copied into the state only on success: s[0..pos0) is still the
		   pending input, so the stock code (same state layout) can take the whole call over */
		st->pos = pos0;
		B2G_FAIL_V(code, bashHashStepH, buf, (p - (const octet*)buf) + count, state);
	}
	p += nfull * st->buf_len, count -= nfull * st->buf_len;
	st->pos = count;
	if (count)
		memcpy(st->s, p, count);
}

static err_t sponge_final(bash_hash_st* st)
{
	memcpy(st->s1, st->s, 192);
	memset(st->s1 + st->pos, 0, st->buf_len - st->pos);
	st->s1[st->pos] = 0x40;
	return bashFBatch(st->s1, 1);
}

void bashHashStepG(octet hash[], size_t hash_len, void* state)
{
	B2G_PREFLIGHT_V(bashHashStepG, hash, hash_len, state);
	bash_hash_st* st = (bash_hash_st*)state;
	err_t code;
	if ((code = sponge_final(st)))
		B2G_FAIL_V(code, bashHashStepG, hash, hash_len, state);   /* s is untouched: stock finishes it */
	memmove(hash, st->s1, hash_len);
}

bool_t bashHashStepV(const octet hash[], size_t hash_len, void* state)
{
	B2G_PREFLIGHT_R(bashHashStepV, hash, hash_len, state);
	bash_hash_st* st = (bash_hash_st*)state;
	err_t code;
	if ((code = sponge_final(st)))
		B2G_FAIL_R(code, bashHashStepV, hash, hash_len, state);
	return b2g_ct_eq(hash, st->s1, hash_len);
}

/* ---------------------------------------------------------------- bash-prg (bash_prg.c:54-385) */
/* same layout as the reference's bash_prg_st; every bash-f runs on the device */
typedef struct
{
	size_t l;
	size_t d;
	octet s[192];
	size_t buf_len;
	size_t pos;
	octet t[192];
} bash_prg_st;

#define PRG_NULL 0x01
#define PRG_KEY 0x05
#define PRG_DATA 0x09
#define PRG_TEXT 0x0D
#define PRG_OUT 0x11

u32 b2g_bashPrgBlocks_dev(void* d_states, void* d_data, size_t stride, size_t nblocks, size_t buf_len,
	int mode, int pre_f, size_t count, void* stream);

size_t bashPrg_keep(void) { return sizeof(bash_prg_st); }

/* bash-f on the state, then `nfull` whole blocks of `mode` over buf (in place) */
static void prg_blocks(bash_prg_st* st, octet* buf, size_t nfull, int mode, const char* fn)
{
	err_t code;
	b2g_slot* sl;
	void *d_state, *d_data;
	const size_t bytes = nfull * st->buf_len;
	if ((code = b2g_ensure_device()))
		b2g_die(fn, code);
	b2g_lock();
	sl = b2g_slot_get(0);
	if ((code = b2g_slot_buf(sl, 0, bytes, &d_data)) || (code = b2g_slot_buf(sl, 1, 192, &d_state)))
		goto done;
	CU(cudaMemcpyAsync(d_state, st->s, 192, cudaMemcpyHostToDevice, sl->stream), "H2D(prg state)");
	if (bytes && mode != 1)
		CU(cudaMemcpyAsync(d_data, buf, bytes, cudaMemcpyHostToDevice, sl->stream), "H2D(prg data)");
	if ((code = b2g_bashPrgBlocks_dev(d_state, d_data, bytes, nfull, st->buf_len, mode, 1, 1, sl->stream)))
		goto done;
	CU(cudaMemcpyAsync(st->s, d_state, 192, cudaMemcpyDeviceToHost, sl->stream), "D2H(prg state)");
	if (bytes && mode != 0)
		CU(cudaMemcpyAsync(buf, d_data, bytes, cudaMemcpyDeviceToHost, sl->stream), "D2H(prg data)");
	CU(cudaStreamSynchronize(sl->stream), "sync(prg)");
done:
	b2g_unlock();
	if (code)
		b2g_die(fn, code);
}

static void prg_commit(bash_prg_st* st, octet code)   /* bash_prg.c:92-105 */
{
	st->s[st->pos] ^= code;
	st->s[st->buf_len] ^= 0x80;
	prg_blocks(st, 0, 0, 0, "bashPrgCommit");
	st->pos = 0;
}

void bashPrgStart(void* state, size_t l, size_t d, const octet ann[], size_t ann_len, const octet key[],
	size_t key_len)
{
	B2G_PREFLIGHT_V(bashPrgStart, state, l, d, ann, ann_len, key, key_len);
	bash_prg_st* st = (bash_prg_st*)state;
	st->pos = 1 + ann_len + key_len;
	memset(st->s, 0, 192);
	st->s[0] = (octet)(ann_len * 4 + key_len / 4);
	if (ann_len) memcpy(st->s + 1, ann, ann_len);
	if (key_len) memcpy(st->s + 1 + ann_len, key, key_len);
	st->s[192 - 8] = (octet)(l / 4 + d);
	st->buf_len = key_len ? (192 - l * (2 + d) / 16) : (192 - d * l / 4);
	st->l = l, st->d = d;
}

void bashPrgRestart(const octet ann[], size_t ann_len, const octet key[], size_t key_len, void* state)
{
	B2G_PREFLIGHT_V(bashPrgRestart, ann, ann_len, key, key_len, state);
	bash_prg_st* st = (bash_prg_st*)state;
	size_t i;
	if (key_len)
	{
		prg_commit(st, PRG_KEY);
		st->buf_len = 192 - st->l * (2 + st->d) / 16;
	}
	else
		prg_commit(st, PRG_NULL);
	st->pos = 1 + ann_len + key_len;
	st->s[0] ^= (octet)(ann_len * 4 + key_len / 4);
	for (i = 0; i < ann_len; ++i) st->s[1 + i] ^= ann[i];
	for (i = 0; i < key_len; ++i) st->s[1 + ann_len + i] ^= key[i];
}

/* the octets that do not complete a block are merged into the host-side state (bookkeeping);
   mode as in bash_prg_kernel */
static void prg_bytes(bash_prg_st* st, octet* buf, size_t n, int mode)
{
	octet* s = st->s + st->pos;
	size_t i;
	for (i = 0; i < n; ++i)
		switch (mode)
		{
		case 0: s[i] ^= buf[i]; break;
		case 1: buf[i] = s[i]; break;
		case 2: s[i] ^= buf[i], buf[i] = s[i]; break;
		default: buf[i] ^= s[i], s[i] ^= buf[i]; break;
		}
}

static void prg_step(bash_prg_st* st, octet* buf, size_t count, int mode, const char* fn)
{
	size_t nfull;
	if (count < st->buf_len - st->pos)
	{
		prg_bytes(st, buf, count, mode);
		st->pos += count;
		return;
	}
	prg_bytes(st, buf, st->buf_len - st->pos, mode);
	buf += st->buf_len - st->pos, count -= st->buf_len - st->pos;
	st->pos = 0;
	nfull = count / st->buf_len;
	prg_blocks(st, buf, nfull, mode, fn);
	buf += nfull * st->buf_len, count -= nfull * st->buf_len;
	if (count)
		prg_bytes(st, buf, count, mode);   /* into s[0..count): pos is still 0 here */
	st->pos = count;
}

void bashPrgAbsorbStart(void* state) { 	B2G_PREFLIGHT_V(bashPrgAbsorbStart, state);
prg_commit((bash_prg_st*)state, PRG_DATA); }
void bashPrgAbsorbStep(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgAbsorbStep, buf, count, state);
	/* absorb never writes to buf */
	prg_step((bash_prg_st*)state, (octet*)(size_t)buf, count, 0, "bashPrgAbsorbStep");
}
void bashPrgAbsorb(const void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgAbsorb, buf, count, state);
	bashPrgAbsorbStart(state);
	bashPrgAbsorbStep(buf, count, state);
}
void bashPrgSqueezeStart(void* state) { 	B2G_PREFLIGHT_V(bashPrgSqueezeStart, state);
prg_commit((bash_prg_st*)state, PRG_OUT); }
void bashPrgSqueezeStep(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgSqueezeStep, buf, count, state);
	prg_step((bash_prg_st*)state, (octet*)buf, count, 1, "bashPrgSqueezeStep");
}
void bashPrgSqueeze(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgSqueeze, buf, count, state);
	bashPrgSqueezeStart(state);
	bashPrgSqueezeStep(buf, count, state);
}
void bashPrgEncrStart(void* state) { 	B2G_PREFLIGHT_V(bashPrgEncrStart, state);
prg_commit((bash_prg_st*)state, PRG_TEXT); }
void bashPrgEncrStep(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgEncrStep, buf, count, state);
	prg_step((bash_prg_st*)state, (octet*)buf, count, 2, "bashPrgEncrStep");
}
void bashPrgEncr(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgEncr, buf, count, state);
	bashPrgEncrStart(state);
	bashPrgEncrStep(buf, count, state);
}
void bashPrgDecrStart(void* state) { 	B2G_PREFLIGHT_V(bashPrgDecrStart, state);
prg_commit((bash_prg_st*)state, PRG_TEXT); }
void bashPrgDecrStep(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgDecrStep, buf, count, state);
	prg_step((bash_prg_st*)state, (octet*)buf, count, 3, "bashPrgDecrStep");
}
void bashPrgDecr(void* buf, size_t count, void* state)
{
	B2G_PREFLIGHT_V(bashPrgDecr, buf, count, state);
	bashPrgDecrStart(state);
	bashPrgDecrStep(buf, count, state);
}
void bashPrgRatchet(void* state)   /* bash_prg.c:374-385 */
{
	B2G_PREFLIGHT_V(bashPrgRatchet, state);
	bash_prg_st* st = (bash_prg_st*)state;
	size_t i;
	memcpy(st->t, st->s, 192);
	prg_commit(st, PRG_NULL);
	for (i = 0; i < 192; ++i) st->s[i] ^= st->t[i];
}

/* ---------------------------------------------------------------- many files (the bsum case)
   cmd/bsum/bsum.c:142-200 hashes one file after another through bashHashStepH with a 4 KiB buffer.
   Here whole files are packed into a pinned staging buffer and every buffer-full goes through ONE ragged
   batch launch (bashHashBatchV: H2D at PCIe rate, one thread per file); a file larger than the staging
   buffer is streamed through bashHashStart / StepH / StepG in buffer-sized pieces.
   status[i]: ERR_OK, ERR_FILE_OPEN (203) or ERR_FILE_READ (207); digest i at hashes + i*(l/4). */
#define FILES_STAGE ((size_t)64 << 20)
#define FILES_MAX_BATCH 65536

static err_t files_flush(octet* hashes, size_t hl, size_t l, const octet* stage, size_t used,
	const u64* offsets, const u64* lens, const size_t* index, size_t n)
{
	err_t code;
	octet* out;
	size_t i;
	if (!n)
		return ERR_OK;
	if (!(out = (octet*)malloc(hl * n)))
		return ERR_OUTOFMEMORY;
	code = bashHashBatchV(out, l, stage, used, offsets, lens, n);
	for (i = 0; !code && i < n; ++i)
		memcpy(hashes + hl * index[i], out + hl * i, hl);
	free(out);
	return code;
}

err_t bashHashFiles(octet* hashes, err_t* status, size_t l, const char* const paths[], size_t count)
{
	err_t code = ERR_OK;
	octet* stage = 0;
	u64 *offsets = 0, *lens = 0;
	size_t* index = 0;
	size_t i, n = 0, used = 0, hl;
	if (l == 0 || l % 16 != 0 || l > 256)
		return ERR_BAD_PARAMS;
	if (count && (!hashes || !status || !paths))
		return ERR_BAD_INPUT;
	if ((code = b2g_ensure_device()))
		return code;
	if (!count)
		return ERR_OK;
	hl = l / 4;
	if (cudaHostAlloc((void**)&stage, FILES_STAGE, cudaHostAllocDefault) != cudaSuccess)
		return b2g_cuda_fail(cudaGetLastError(), "cudaHostAlloc(files)");
	offsets = (u64*)malloc(8 * FILES_MAX_BATCH), lens = (u64*)malloc(8 * FILES_MAX_BATCH);
	index = (size_t*)malloc(sizeof(size_t) * FILES_MAX_BATCH);
	if (!offsets || !lens || !index)
	{
		code = ERR_OUTOFMEMORY;
		goto done;
	}
	for (i = 0; i < count && !code; ++i)
	{
		FILE* f = paths[i] ? fopen(paths[i], "rb") : 0;
		long size = -1;
		size_t got, start;
		if (!f)
		{
			status[i] = ERR_FILE_OPEN;
			continue;
		}
		if (fseek(f, 0, SEEK_END) == 0)
			size = ftell(f), rewind(f);
		if (size < 0 || (size_t)size > FILES_STAGE)
		{
			/* larger than the staging buffer (or not seekable): stream it through the sponge state in
			   buffer-sized pieces — sequential by construction. The staged files go first. */
			octet state[sizeof(bash_hash_st)];
			int bad = 0;
			if ((code = files_flush(hashes, hl, l, stage, used, offsets, lens, index, n)))
			{
				fclose(f);
				break;
			}
			n = 0, used = 0;
			bashHashStart(state, l);
			do
			{
				got = fread(stage, 1, FILES_STAGE, f);
				if (ferror(f))
				{
					bad = 1;
					break;
				}
				bashHashStepH(stage, got, state);
			}
			while (got == FILES_STAGE);
			if (bad)
				status[i] = ERR_FILE_READ;
			else
				bashHashStepG(hashes + hl * i, hl, state), status[i] = ERR_OK;
			fclose(f);
			continue;
		}
		/* an 8-aligned start lets the kernel use 64-bit loads for this message */
		start = (used + 7) & ~(size_t)7;
		if (start + (size_t)size > FILES_STAGE || n == FILES_MAX_BATCH)
		{
			if ((code = files_flush(hashes, hl, l, stage, used, offsets, lens, index, n)))
			{
				fclose(f);
				break;
			}
			n = 0, used = 0, start = 0;
		}
		got = fread(stage + start, 1, (size_t)size, f);
		if (ferror(f))
		{
			status[i] = ERR_FILE_READ;
			fclose(f);
			continue;
		}
		fclose(f);
		offsets[n] = start, lens[n] = got, index[n] = i, ++n;
		used = start + got;
		status[i] = ERR_OK;
	}
	if (!code)
		code = files_flush(hashes, hl, l, stage, used, offsets, lens, index, n);
done:
	free(offsets), free(lens), free(index);
	cudaFreeHost(stage);
	return code;
}
