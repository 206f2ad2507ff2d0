/*
 * bee2_oracle.c — plain-C CPU restatement of the bee2 hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see bee2_oracle.h). Written from the algorithm descriptions
 * in SURVEY.md Appendix A and the cited reference lines; no reference source is copied
 * (tables such as the belt S-box and the bash round constants are regenerated from
 * their defining recurrences). Little-endian host assumed (x86-64 / aarch64).
 *
 * Parity: PINNED by tests/test_oracle_kat.py (STB annex vectors from the reference's
 * test/crypto/{bash,belt,bign}_test.c) and by differential runs against
 * oracle/_ref/libbee2ref_64.so (the unmodified reference compiled by oracle/Makefile).
 */
#include "bee2_oracle.h"
#include <string.h>
#include <stdlib.h>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef unsigned __int128 u128;

static inline u64 rotl64(u64 x, unsigned n) { return (x << n) | (x >> (64 - n)); }
static inline u32 rotl32(u32 x, unsigned n) { return (x << n) | (x >> (32 - n)); }
static inline u32 ld32(const u8* p) { u32 v; memcpy(&v, p, 4); return v; }
static inline void st32(u8* p, u32 v) { memcpy(p, &v, 4); }

/* ======================================================================= bash */

/* bash_f64.c:32-44 (S-box), :50-57 (constants), :100-112 (P), :126-133 (rotations) */
void orc_bashF(u8 block[192])
{
	u64 s[24], n[24], c = 0x3BF5080AC8BA94B1ull;
	unsigned rot[8][4] = {{8, 53, 14, 1}};
	int t, j, x;
	for (j = 1; j < 8; ++j)
		for (x = 0; x < 4; ++x)
			rot[j][x] = 7 * rot[j - 1][x] % 64;
	memcpy(s, block, 192);
	for (t = 0; t < 24; ++t)
	{
		for (j = 0; j < 8; ++j)
		{
			u64 w0 = s[j], w1 = s[8 + j], w2 = s[16 + j], t1, t2, u0, u1, u2;
			t2 = rotl64(w0, rot[j][0]);
			w0 ^= w1 ^ w2;
			t1 = w1 ^ rotl64(w0, rot[j][1]);
			w1 = t1 ^ t2;
			w2 ^= rotl64(w2, rot[j][2]) ^ rotl64(t1, rot[j][3]);
			u0 = ~w2 | w1, u1 = w0 | w2, u2 = w0 & w1;
			s[j] = w0 ^ u0, s[8 + j] = w1 ^ u1, s[16 + j] = w2 ^ u2;
		}
		for (x = 0; x < 24; ++x)
		{
			int from = x < 8 ? 8 + (x + 2 * (x & 1) + 7) % 8 :
				x < 16 ? 8 + (x ^ 1) : (5 * x + 6) % 8;
			n[x] = s[from];
		}
		memcpy(s, n, sizeof s);
		s[23] ^= c;
		c = (c >> 1) ^ ((c & 1) ? 0xDC2BE1997FE0D8AEull : 0);
	}
	memcpy(block, s, 192);
}

void orc_bashHashStart(orc_bash_st* st, size_t l)
{
	memset(st->s, 0, 192);
	st->s[184] = (u8)(l / 4);
	st->rate = 192 - l / 2;
	st->pos = 0;
}

void orc_bashHashStepH(const void* buf, size_t n, orc_bash_st* st)
{
	const u8* p = (const u8*)buf;
	while (n)
	{
		size_t take = st->rate - st->pos;
		if (take > n) take = n;
		memcpy(st->s + st->pos, p, take);   /* overwrite, not XOR: bash_hash.c:59,64,71 */
		st->pos += take, p += take, n -= take;
		if (st->pos == st->rate)
			orc_bashF(st->s), st->pos = 0;
	}
}

void orc_bashHashStepG(u8* hash, size_t hash_len, const orc_bash_st* st)
{
	u8 s1[192];
	memcpy(s1, st->s, 192);
	memset(s1 + st->pos, 0, st->rate - st->pos);  /* pos == 0 -> a whole extra block, :95-99 */
	s1[st->pos] = 0x40;
	orc_bashF(s1);
	memcpy(hash, s1, hash_len);
}

u32 orc_bashHash(u8* hash, size_t l, const void* src, size_t n)
{
	orc_bash_st st;
	if (l == 0 || l % 16 != 0 || l > 256)
		return ORC_BAD_PARAMS;
	orc_bashHashStart(&st, l);
	orc_bashHashStepH(src, n, &st);
	orc_bashHashStepG(hash, l / 4, &st);
	return ORC_OK;
}

/* ======================================================================= belt */

static u8 H_[256];
static int H_ready;

/* S-box from its definition (test/crypto/belt_test.c:27-57): H[10]=0,
   H[(11+x)%256] = 0x8E * z^(116x) in GF(2)[z]/(z^8+z^7+z^6+z+1). */
const u8* orc_beltH(void)
{
	if (!H_ready)
	{
		unsigned x, i;
		H_[10] = 0, H_[11] = 0x8E;
		for (x = 12; x < 266; ++x)
		{
			unsigned t = H_[(x - 1) % 256];
			for (i = 0; i < 116; ++i)
				t = (t >> 1) | ((unsigned)__builtin_parity(t & 0x63) << 7);
			H_[x % 256] = (u8)t;
		}
		H_ready = 1;
	}
	return H_;
}

static inline u32 G(u32 x, unsigned r)
{
	const u8* H = orc_beltH();
	u32 v = (u32)H[x & 255] | (u32)H[x >> 8 & 255] << 8 | (u32)H[x >> 16 & 255] << 16 | (u32)H[x >> 24] << 24;
	return rotl32(v, r);
}

void orc_beltKeyExpand2(u32 k[8], const u8* key, size_t len)
{
	size_t i;
	for (i = 0; i < len / 4; ++i) k[i] = ld32(key + 4 * i);
	if (len == 16)
		for (i = 0; i < 4; ++i) k[4 + i] = k[i];
	else if (len == 24)
		k[6] = k[0] ^ k[1] ^ k[2], k[7] = k[3] ^ k[4] ^ k[5];
}

/* belt_block.c:231-269 */
void orc_beltBlockEncr2(u32 blk[4], const u32 k[8])
{
	u32 a = blk[0], b = blk[1], c = blk[2], d = blk[3], e, t;
	unsigned i;
	for (i = 1; i <= 8; ++i)
	{
		const unsigned o = 7 * i - 7;
		b ^= G(a + k[(o + 0) % 8], 5);
		c ^= G(d + k[(o + 1) % 8], 21);
		a -= G(b + k[(o + 2) % 8], 13);
		e = G(b + c + k[(o + 3) % 8], 21) ^ i;
		b += e, c -= e;
		d += G(c + k[(o + 4) % 8], 13);
		b ^= G(a + k[(o + 5) % 8], 21);
		c ^= G(d + k[(o + 6) % 8], 5);
		t = a, a = b, b = t;
		t = c, c = d, d = t;
		t = b, b = c, c = t;
	}
	blk[0] = b, blk[1] = d, blk[2] = a, blk[3] = c;
}

/* belt_block.c:243,284-295 */
void orc_beltBlockDecr2(u32 blk[4], const u32 k[8])
{
	u32 a = blk[0], b = blk[1], c = blk[2], d = blk[3], e, t;
	unsigned i;
	for (i = 8; i >= 1; --i)
	{
		const unsigned o = 7 * i - 1;
		b ^= G(a + k[(o - 0) % 8], 5);
		c ^= G(d + k[(o - 1) % 8], 21);
		a -= G(b + k[(o - 2) % 8], 13);
		e = G(b + c + k[(o - 3) % 8], 21) ^ i;
		b += e, c -= e;
		d += G(c + k[(o - 4) % 8], 13);
		b ^= G(a + k[(o - 5) % 8], 21);
		c ^= G(d + k[(o - 6) % 8], 5);
		t = a, a = b, b = t;
		t = c, c = d, d = t;
		t = a, a = d, d = t;
	}
	blk[0] = c, blk[1] = a, blk[2] = d, blk[3] = b;
}

static void blk_encr(u8 b[16], const u32 k[8])
{
	u32 w[4];
	memcpy(w, b, 16), orc_beltBlockEncr2(w, k), memcpy(b, w, 16);
}
static void blk_decr(u8 b[16], const u32 k[8])
{
	u32 w[4];
	memcpy(w, b, 16), orc_beltBlockDecr2(w, k), memcpy(b, w, 16);
}

void orc_beltCTRStart(orc_belt_ctr_st* st, const u8* key, size_t len, const u8 iv[16])
{
	orc_beltKeyExpand2(st->key, key, len);
	memcpy(st->ctr, iv, 16);
	orc_beltBlockEncr2(st->ctr, st->key);
	st->reserved = 0;
}

static void ctr_next(orc_belt_ctr_st* st)
{
	/* 128-bit little-endian increment, belt_ctr.c:27-35 */
	if (++st->ctr[0] == 0 && ++st->ctr[1] == 0 && ++st->ctr[2] == 0) ++st->ctr[3];
	memcpy(st->block, st->ctr, 16);
	blk_encr(st->block, st->key);
}

void orc_beltCTRStepE(void* buf, size_t n, orc_belt_ctr_st* st)
{
	u8* p = (u8*)buf;
	size_t i;
	while (n)
	{
		size_t take;
		if (!st->reserved)
			ctr_next(st), st->reserved = 16;
		take = st->reserved < n ? st->reserved : n;
		for (i = 0; i < take; ++i)
			p[i] ^= st->block[16 - st->reserved + i];
		st->reserved -= take, p += take, n -= take;
	}
}

u32 orc_beltCTR(void* dst, const void* src, size_t n, const u8* key, size_t len, const u8 iv[16])
{
	orc_belt_ctr_st st;
	if (len != 16 && len != 24 && len != 32)
		return ORC_BAD_INPUT;
	orc_beltCTRStart(&st, key, len, iv);
	memmove(dst, src, n);
	orc_beltCTRStepE(dst, n, &st);
	return ORC_OK;
}

/* full blocks then ciphertext stealing, belt_ecb.c:62-110 */
static u32 ecb(void* dst, const void* src, size_t n, const u8* key, size_t len, int enc)
{
	u32 k[8];
	u8* p = (u8*)dst;
	size_t full, r, i;
	if (n < 16 || (len != 16 && len != 24 && len != 32))
		return ORC_BAD_INPUT;
	orc_beltKeyExpand2(k, key, len);
	memmove(dst, src, n);
	full = n / 16, r = n % 16;
	for (i = 0; i < full; ++i)
		enc ? blk_encr(p + 16 * i, k) : blk_decr(p + 16 * i, k);
	if (r)
	{
		u8 t[16];
		u8* last = p + 16 * (full - 1);
		memcpy(t, last + 16, r);
		memcpy(t + r, last + r, 16 - r);
		enc ? blk_encr(t, k) : blk_decr(t, k);
		memcpy(last + 16, last, r);
		memcpy(last, t, 16);
	}
	return ORC_OK;
}
u32 orc_beltECBEncr(void* d, const void* s, size_t n, const u8* key, size_t len) { return ecb(d, s, n, key, len, 1); }
u32 orc_beltECBDecr(void* d, const void* s, size_t n, const u8* key, size_t len) { return ecb(d, s, n, key, len, 0); }

void orc_beltECBEncrMultiKey(u8* blocks, const u8* keys32, size_t count)
{
	size_t i;
	for (i = 0; i < count; ++i)
	{
		u32 k[8];
		orc_beltKeyExpand2(k, keys32 + 32 * i, 32);
		blk_encr(blocks + 16 * i, k);
	}
}

/* sigma1/sigma2 compression, belt_compr.c:27-87. h: 8 words, X: 8 words; s may be NULL. */
static void belt_compress(u32 s[4], u32 h[8], const u32 X[8])
{
	u32 S[4], k1[8], k2[8], y0[4], y1[4];
	int i;
	for (i = 0; i < 4; ++i) S[i] = h[i] ^ h[4 + i];
	orc_beltBlockEncr2(S, X);
	for (i = 0; i < 4; ++i) S[i] ^= h[i] ^ h[4 + i];
	if (s)
		for (i = 0; i < 4; ++i) s[i] ^= S[i];
	for (i = 0; i < 4; ++i)
		k1[i] = S[i], k1[4 + i] = h[4 + i], k2[i] = ~S[i], k2[4 + i] = h[i];
	memcpy(y0, X, 16), memcpy(y1, X + 4, 16);
	orc_beltBlockEncr2(y0, k1);
	orc_beltBlockEncr2(y1, k2);
	for (i = 0; i < 4; ++i)
		h[i] = y0[i] ^ X[i], h[4 + i] = y1[i] ^ X[4 + i];
}

/* belt_hash.c:43-190 */
void orc_beltHash(u8 hash[32], const void* src, size_t n)
{
	const u8* p = (const u8*)src;
	u32 ls[8] = {0}, h[8], X[8];
	u64 bits_lo = (u64)n << 3, bits_hi = (u64)n >> 61;
	size_t i;
	memcpy(h, orc_beltH(), 32);
	for (i = 0; i + 32 <= n; i += 32)
		memcpy(X, p + i, 32), belt_compress(ls + 4, h, X);
	if (i < n)
	{
		memset(X, 0, 32), memcpy(X, p + i, n - i);
		belt_compress(ls + 4, h, X);
	}
	ls[0] = (u32)bits_lo, ls[1] = (u32)(bits_lo >> 32), ls[2] = (u32)bits_hi, ls[3] = (u32)(bits_hi >> 32);
	belt_compress(0, h, ls);
	memcpy(hash, h, 32);
}

/* wide-block encryption of count = 32, 48 or 64 bytes: 2n rounds over n = count/16 blocks
   (belt_wbl.c:50-82; the round counter restarts on every call, :199-207) */
static void belt_wbl(u8* buf, size_t count, const u32 key[8])
{
	const size_t n = count / 16;
	u64 round;
	for (round = 1; round <= 2 * n; ++round)
	{
		u8 s[16], e[16];
		size_t i, j;
		/* s <- r1 + ... + r_{n-1} */
		memcpy(s, buf, 16);
		for (j = 1; j + 1 < n; ++j)
			for (i = 0; i < 16; ++i) s[i] ^= buf[16 * j + i];
		/* r <- ShLo^128(r), r* <- s */
		memmove(buf, buf + 16, count - 16);
		memcpy(buf + count - 16, s, 16);
		/* r*_before_shift += E(s) + <round> */
		memcpy(e, s, 16);
		blk_encr(e, key);
		for (i = 0; i < 8; ++i) e[i] ^= (u8)(round >> (8 * i));
		for (i = 0; i < 16; ++i) buf[count - 32 + i] ^= e[i];
	}
}

/* ======================================================================= GF(p), p = 2^(64 n) - c */
/* The three standard bign curves (bign_params.c:36-73, :78-125, :131-190): level l = 128 / 192 /
   256, n = l/32 words of 64 bits, p = 2^(2l) - c with c = 189 / 317 / 569 (Crandall reduction
   zz_red.c:71-105), a = p - 3, G = (0, yG). */

#define FE_MAXW 8
typedef struct { u64 w[FE_MAXW]; } fe;           /* words above n are kept 0 */
typedef struct { int n; size_t no; fe p, q, yG, b; } lvl;

static const u8 Q128_LE[32] = { /* q of bign-curve256v1 */
	0x07, 0x66, 0x3D, 0x26, 0x99, 0xBF, 0x5A, 0x7E, 0xFC, 0x4D, 0xFB, 0x0D, 0xD6, 0x8E, 0x5C, 0xD9,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
static const u8 YG128_LE[32] = { /* base point G = (0, yG) */
	0x93, 0x6A, 0x51, 0x04, 0x18, 0xCF, 0x29, 0x1E, 0x52, 0xF6, 0x08, 0xC4, 0x66, 0x39, 0x91, 0x78,
	0x5D, 0x83, 0xD6, 0x51, 0xA3, 0xC9, 0xE4, 0x5C, 0x9F, 0xD6, 0x16, 0xFB, 0x3C, 0xFC, 0xF7, 0x6B};
static const u8 Q192_LE[48] = { /* q of bign-curve384v1 */
	0xB7, 0xA7, 0x0C, 0xF3, 0x3F, 0xDC, 0xB7, 0x3D, 0x0A, 0xFF, 0xA4, 0xA6, 0xE7, 0xDA, 0x46, 0x80,
	0xBB, 0x7B, 0xAF, 0x73, 0x03, 0xC4, 0xCC, 0x6C, 0xFE, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
static const u8 YG192_LE[48] = { /* base point G = (0, yG) */
	0x51, 0xC4, 0x33, 0xF7, 0x31, 0xCB, 0x5E, 0xEA, 0xF9, 0x42, 0x2A, 0x6B, 0x27, 0x3E, 0x40, 0x84,
	0x55, 0xD3, 0xB1, 0x66, 0x9E, 0xE7, 0x49, 0x05, 0xA0, 0xFF, 0x86, 0xDC, 0x11, 0x9A, 0x72, 0x3A,
	0x89, 0xBF, 0x2D, 0x43, 0x7E, 0x11, 0x30, 0x63, 0x9E, 0x9E, 0x2E, 0xA8, 0x24, 0x82, 0x43, 0x5D};
static const u8 Q256_LE[64] = { /* q of bign-curve512v1 */
	0xF1, 0x8E, 0x06, 0x0D, 0x49, 0xAD, 0xFF, 0xDC, 0x32, 0xDF, 0x56, 0x95, 0xE5, 0xCA, 0x1B, 0x36,
	0xF4, 0x13, 0x21, 0x2E, 0xB0, 0xEB, 0x6B, 0xF2, 0x4E, 0x00, 0x98, 0x01, 0x2C, 0x09, 0xC0, 0xB2,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
	0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
static const u8 YG256_LE[64] = { /* base point G = (0, yG) */
	0xBD, 0xED, 0xEF, 0xCE, 0x6F, 0xAE, 0x92, 0xB7, 0x04, 0x0D, 0x4C, 0xC9, 0xB9, 0x83, 0xAA, 0x67,
	0x61, 0x22, 0xE8, 0xEE, 0x95, 0x73, 0x77, 0xFF, 0xD2, 0x6F, 0xFA, 0x0E, 0xE2, 0xDD, 0x73, 0x69,
	0xDA, 0xCA, 0xCC, 0x00, 0x1B, 0xF8, 0xED, 0xD2, 0xE2, 0xBC, 0x61, 0xB3, 0xB3, 0x41, 0xAB, 0xB0,
	0xAB, 0x8F, 0xD1, 0xA0, 0xF7, 0xE6, 0x82, 0xB1, 0x81, 0x76, 0x03, 0xE4, 0x7A, 0xFF, 0x26, 0xA8};

static const u8 B128_LE[32] = { /* coefficient b of bign-curve256v1 */
	0xF1, 0x03, 0x9C, 0xD6, 0x6B, 0x7D, 0x2E, 0xB2, 0x53, 0x92, 0x8B, 0x97, 0x69, 0x50, 0xF5, 0x4C,
	0xBE, 0xFB, 0xD8, 0xE4, 0xAB, 0x3A, 0xC1, 0xD2, 0xED, 0xA8, 0xF3, 0x15, 0x15, 0x6C, 0xCE, 0x77};
static const u8 B192_LE[48] = { /* coefficient b of bign-curve384v1 */
	0x64, 0xBF, 0x73, 0x68, 0x23, 0xFC, 0xA7, 0xBC, 0x7C, 0xBD, 0xCE, 0xF3, 0xF0, 0xE2, 0xBD, 0x14,
	0x3A, 0x2E, 0x71, 0xE9, 0xF9, 0x6A, 0x21, 0xA6, 0x96, 0xB1, 0xFB, 0x0F, 0xBB, 0x48, 0x27, 0x71,
	0xD2, 0x34, 0x5D, 0x65, 0xAB, 0x5A, 0x07, 0x33, 0x20, 0xEF, 0x9C, 0x95, 0xE1, 0xDF, 0x75, 0x3C};
static const u8 B256_LE[64] = { /* coefficient b of bign-curve512v1 */
	0x90, 0x9C, 0x13, 0xD6, 0x98, 0x69, 0x34, 0x09, 0x7A, 0xA2, 0x49, 0x3A, 0x27, 0x22, 0x86, 0xEA,
	0x43, 0xA2, 0xAC, 0x87, 0x8C, 0x00, 0x33, 0x29, 0x95, 0x5E, 0x24, 0xC4, 0xB5, 0xDC, 0x11, 0x27,
	0x88, 0xB0, 0xAD, 0xDA, 0xE3, 0x13, 0xCE, 0x17, 0x51, 0x25, 0x5D, 0xDD, 0xEE, 0xA9, 0xC6, 0x5B,
	0x89, 0x58, 0xFD, 0x60, 0x6A, 0x5D, 0x8C, 0xD8, 0x43, 0x8C, 0x3B, 0x93, 0x44, 0x59, 0xB4, 0x6C};
static fe fe_from_n(const u8* b, size_t no) { fe r; memset(&r, 0, sizeof r); memcpy(r.w, b, no); return r; }
static const lvl* level(size_t l)
{
	static lvl L[3];
	static int ready;
	if (!ready)
	{
		static const size_t ls[3] = {128, 192, 256};
		static const u64 cs[3] = {189, 317, 569};
		static const u8* const qs[3] = {Q128_LE, Q192_LE, Q256_LE};
		static const u8* const ys[3] = {YG128_LE, YG192_LE, YG256_LE};
		static const u8* const bs[3] = {B128_LE, B192_LE, B256_LE};
		int i, j;
		for (i = 0; i < 3; ++i)
		{
			L[i].n = (int)(ls[i] / 32), L[i].no = ls[i] / 4;
			memset(&L[i].p, 0, sizeof(fe));
			for (j = 0; j < L[i].n; ++j) L[i].p.w[j] = ~0ull;
			L[i].p.w[0] -= cs[i] - 1;
			L[i].q = fe_from_n(qs[i], L[i].no), L[i].yG = fe_from_n(ys[i], L[i].no), L[i].b = fe_from_n(bs[i], L[i].no);
		}
		ready = 1;
	}
	return l == 128 ? &L[0] : l == 192 ? &L[1] : l == 256 ? &L[2] : 0;
}

static fe fe_from(const lvl* L, const u8* b) { return fe_from_n(b, L->no); }
static void fe_to(const lvl* L, u8* b, fe a) { memcpy(b, a.w, L->no); }
static int fe_cmp(fe a, fe b)
{
	int i;
	for (i = FE_MAXW - 1; i >= 0; --i)
		if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
	return 0;
}
static int fe_is0(fe a)
{
	u64 z = 0; int i;
	for (i = 0; i < FE_MAXW; ++i) z |= a.w[i];
	return z == 0;
}
/* n-word add / sub, returning the carry / borrow out of word n-1 */
static u64 raw_add(const lvl* L, fe* r, fe a, fe b)
{
	u128 c = 0; int i;
	memset(r, 0, sizeof *r);
	for (i = 0; i < L->n; ++i) c += (u128)a.w[i] + b.w[i], r->w[i] = (u64)c, c >>= 64;
	return (u64)c;
}
static u64 raw_sub(const lvl* L, fe* r, fe a, fe b)
{
	u64 br = 0; int i;
	memset(r, 0, sizeof *r);
	for (i = 0; i < L->n; ++i)
	{
		u128 d = (u128)a.w[i] - b.w[i] - br;
		r->w[i] = (u64)d, br = (u64)(d >> 64) & 1;
	}
	return br;
}
/* (a + b) mod m, a, b < m (zz_mod.c:42) */
static fe addmod(const lvl* L, fe a, fe b, fe m)
{
	fe r, t;
	u64 c = raw_add(L, &r, a, b);
	if (c || fe_cmp(r, m) >= 0) raw_sub(L, &t, r, m), r = t;
	return r;
}
/* a - b mod 2^(64n), + m if it borrowed (zz_mod.c:120) */
static fe submod(const lvl* L, fe a, fe b, fe m)
{
	fe r, t;
	if (raw_sub(L, &r, a, b)) raw_add(L, &t, r, m), r = t;
	return r;
}
static fe fp_add(const lvl* L, fe a, fe b) { return addmod(L, a, b, L->p); }
static fe fp_sub(const lvl* L, fe a, fe b) { return submod(L, a, b, L->p); }

/* r[0..na+nb) = a * b (zz_mul.c:82-120) */
static void mul_wide(u64* r, const u64* a, int na, const u64* b, int nb)
{
	int i, j;
	memset(r, 0, 8 * (size_t)(na + nb));
	for (i = 0; i < na; ++i)
	{
		u64 carry = 0;
		for (j = 0; j < nb; ++j)
		{
			u128 t = (u128)a[i] * b[j] + r[i + j] + carry;
			r[i + j] = (u64)t, carry = (u64)(t >> 64);
		}
		r[i + nb] = carry;
	}
}

/* x (2n words) mod m, m = 2^(64n) - c: fold hi * c into lo until hi = 0, then subtract m while >= m */
static fe fold_mod(const lvl* L, const u64* x, fe m)
{
	const int n = L->n;
	u64 cur[2 * FE_MAXW], hc[2 * FE_MAXW];
	fe lo, hi, t, c, z;
	int i;
	memset(&z, 0, sizeof z);
	raw_sub(L, &c, z, m);                /* c = 2^(64n) - m */
	memcpy(cur, x, 16 * (size_t)n);
	for (;;)
	{
		memset(&lo, 0, sizeof lo), memset(&hi, 0, sizeof hi);
		memcpy(lo.w, cur, 8 * (size_t)n), memcpy(hi.w, cur + n, 8 * (size_t)n);
		if (fe_is0(hi)) break;
		mul_wide(hc, hi.w, n, c.w, n);
		{
			u128 acc = 0;
			for (i = 0; i < 2 * n; ++i)
			{
				acc += (u128)hc[i] + (i < n ? lo.w[i] : 0);
				cur[i] = (u64)acc, acc >>= 64;
			}
		}
	}
	while (fe_cmp(lo, m) >= 0) raw_sub(L, &t, lo, m), lo = t;
	return lo;
}

static fe fp_mul(const lvl* L, fe a, fe b)
{
	u64 r[2 * FE_MAXW];
	mul_wide(r, a.w, L->n, b.w, L->n);
	return fold_mod(L, r, L->p);
}
static fe fp_sqr(const lvl* L, fe a) { return fp_mul(L, a, a); }
/* a^(p-2), gfp.c:33-44 */
static fe fp_inv(const lvl* L, fe a)
{
	fe e = L->p, r;
	int i;
	memset(&r, 0, sizeof r), r.w[0] = 1;
	e.w[0] -= 2;
	for (i = 64 * L->n - 1; i >= 0; --i)
	{
		r = fp_sqr(L, r);
		if (e.w[i / 64] >> (i % 64) & 1) r = fp_mul(L, r, a);
	}
	return r;
}

void orc_gfpMul(u8 c[32], const u8 a[32], const u8 b[32])
{
	const lvl* L = level(128);
	fe_to(L, c, fp_mul(L, fe_from(L, a), fe_from(L, b)));
}
void orc_gfpInv(u8 c[32], const u8 a[32])
{
	const lvl* L = level(128);
	fe_to(L, c, fp_inv(L, fe_from(L, a)));
}

/* ======================================================================= curve y^2 = x^3 - 3x + b */

typedef struct { fe X, Y, Z; } pt;  /* Jacobian, O <=> Z = 0 (ecp_j.c) */

static pt pt_inf(void)
{
	pt R;
	memset(&R, 0, sizeof R), R.X.w[0] = R.Y.w[0] = 1;
	return R;
}

static pt pt_dbl(const lvl* L, pt P)
{
	pt R;
	fe d, g, bt, al, t, u;
	if (fe_is0(P.Z) || fe_is0(P.Y)) return pt_inf();
	d = fp_sqr(L, P.Z), g = fp_sqr(L, P.Y), bt = fp_mul(L, P.X, g);
	t = fp_sub(L, P.X, d), u = fp_add(L, P.X, d), al = fp_mul(L, t, u);
	al = fp_add(L, fp_add(L, al, al), al);                /* 3(X - Z^2)(X + Z^2), a = -3 */
	t = fp_add(L, bt, bt), t = fp_add(L, t, t);           /* 4 beta */
	R.X = fp_sub(L, fp_sqr(L, al), fp_add(L, t, t));
	u = fp_add(L, P.Y, P.Z), R.Z = fp_sub(L, fp_sub(L, fp_sqr(L, u), g), d);
	g = fp_sqr(L, g), g = fp_add(L, g, g), g = fp_add(L, g, g), g = fp_add(L, g, g);  /* 8 gamma^2 */
	R.Y = fp_sub(L, fp_mul(L, al, fp_sub(L, t, R.X)), g);
	return R;
}

static pt pt_add(const lvl* L, pt P, pt Q)
{
	pt R;
	fe z1z1, z2z2, u1, u2, s1, s2, h, r, hh, hhh, v;
	if (fe_is0(P.Z)) return Q;
	if (fe_is0(Q.Z)) return P;
	z1z1 = fp_sqr(L, P.Z), z2z2 = fp_sqr(L, Q.Z);
	u1 = fp_mul(L, P.X, z2z2), u2 = fp_mul(L, Q.X, z1z1);
	s1 = fp_mul(L, P.Y, fp_mul(L, Q.Z, z2z2)), s2 = fp_mul(L, Q.Y, fp_mul(L, P.Z, z1z1));
	h = fp_sub(L, u2, u1), r = fp_sub(L, s2, s1);
	if (fe_is0(h))
		return fe_is0(r) ? pt_dbl(L, P) : pt_inf();
	hh = fp_sqr(L, h), hhh = fp_mul(L, h, hh), v = fp_mul(L, u1, hh);
	R.X = fp_sub(L, fp_sub(L, fp_sqr(L, r), hhh), fp_add(L, v, v));
	R.Y = fp_sub(L, fp_mul(L, r, fp_sub(L, v, R.X)), fp_mul(L, s1, hhh));
	R.Z = fp_mul(L, fp_mul(L, P.Z, Q.Z), h);
	return R;
}

/* scalar given as little-endian octets of any length */
static pt pt_mul(const lvl* L, pt A, const u8* d, size_t d_len)
{
	pt R = pt_inf();
	long i;
	for (i = (long)d_len * 8 - 1; i >= 0; --i)
	{
		R = pt_dbl(L, R);
		if (d[i / 8] >> (i % 8) & 1) R = pt_add(L, R, A);
	}
	return R;
}

static int pt_to_affine(const lvl* L, fe* x, fe* y, pt P)
{
	fe zi, zi2;
	if (fe_is0(P.Z)) return 0;
	zi = fp_inv(L, P.Z), zi2 = fp_sqr(L, zi);
	*x = fp_mul(L, P.X, zi2), *y = fp_mul(L, P.Y, fp_mul(L, zi2, zi));
	return 1;
}

static pt pt_affine(fe x, fe y)
{
	pt P;
	P.X = x, P.Y = y, memset(&P.Z, 0, sizeof P.Z), P.Z.w[0] = 1;
	return P;
}
static pt pt_base(const lvl* L)
{
	fe zero;
	memset(&zero, 0, sizeof zero);
	return pt_affine(zero, L->yG);
}

/* ecMulA (ec.c:497-525) on the standard curve of level l: b = d * a, 0 iff the result is O */
int orc_ecMulA(size_t l, u8* b, const u8* a, const u8* d, size_t d_len)
{
	const lvl* L = level(l);
	fe x, y;
	if (!L) return 0;
	if (!pt_to_affine(L, &x, &y, pt_mul(L, pt_affine(fe_from(L, a), fe_from(L, a + L->no)), d, d_len)))
		return 0;
	fe_to(L, b, x), fe_to(L, b + L->no, y);
	return 1;
}
int orc_ecMulA128(u8 b[64], const u8 a[64], const u8* d, size_t d_len) { return orc_ecMulA(128, b, a, d, d_len); }

/* bign_sign.c:268-347; no = l/4 octets: hash no, sig no/2 + no, pubkey 2 no */
u32 orc_bignVerify(size_t l, const u8* oid_der, size_t oid_len, const u8* hash, const u8* sig, const u8* pubkey)
{
	const lvl* L = level(l);
	size_t no;
	fe Qx, Qy, s1, Hh, t, x, y;
	u8 s0[33], s1b[64], buf[128 + 128], hv[32];
	pt R;
	if (!L) return 119u;
	no = L->no;
	if (oid_len > 128) return ORC_BAD_INPUT;
	Qx = fe_from(L, pubkey), Qy = fe_from(L, pubkey + no), s1 = fe_from(L, sig + no / 2), Hh = fe_from(L, hash);
	if (fe_cmp(Qx, L->p) >= 0 || fe_cmp(Qy, L->p) >= 0) return ORC_BAD_PUBKEY;
	if (fe_cmp(s1, L->q) >= 0) return ORC_BAD_SIG;
	if (fe_cmp(Hh, L->q) >= 0) raw_sub(L, &t, Hh, L->q), Hh = t;
	s1 = addmod(L, s1, Hh, L->q);
	memcpy(s0, sig, no / 2), s0[no / 2] = 1;
	fe_to(L, s1b, s1);
	R = pt_add(L, pt_mul(L, pt_base(L), s1b, no), pt_mul(L, pt_affine(Qx, Qy), s0, no / 2 + 1));
	if (!pt_to_affine(L, &x, &y, R)) return ORC_BAD_SIG;
	memcpy(buf, oid_der, oid_len), fe_to(L, buf + oid_len, x), memcpy(buf + oid_len + no, hash, no);
	orc_beltHash(hv, buf, oid_len + 2 * no);
	return memcmp(hv, sig, no / 2) == 0 ? ORC_OK : ORC_BAD_SIG;
}
u32 orc_bignVerify128(const u8* oid_der, size_t oid_len, const u8 hash[32], const u8 sig[48], const u8 pubkey[64])
{
	return orc_bignVerify(128, oid_der, oid_len, hash, sig, pubkey);
}

/* bign_misc.c:369-412 */
u32 orc_bignPubkeyCalc(size_t l, u8* pubkey, const u8* privkey)
{
	const lvl* L = level(l);
	fe d, x, y;
	if (!L) return 119u;
	d = fe_from(L, privkey);
	if (fe_is0(d) || fe_cmp(d, L->q) >= 0) return ORC_BAD_PRIVKEY;
	if (!pt_to_affine(L, &x, &y, pt_mul(L, pt_base(L), privkey, L->no))) return ORC_BAD_PARAMS;
	fe_to(L, pubkey, x), fe_to(L, pubkey + L->no, y);
	return ORC_OK;
}
u32 orc_bignPubkeyCalc128(u8 pubkey[64], const u8 privkey[32]) { return orc_bignPubkeyCalc(128, pubkey, privkey); }

/* bign_misc.c:317-352: coordinates < p and y^2 = x^3 - 3x + b (ecpIsOnA, ecp_j.c) */
u32 orc_bignPubkeyVal(size_t l, const u8* pubkey)
{
	const lvl* L = level(l);
	fe x, y, lhs, rhs, t;
	if (!L) return 119u;
	x = fe_from(L, pubkey), y = fe_from(L, pubkey + L->no);
	if (fe_cmp(x, L->p) >= 0 || fe_cmp(y, L->p) >= 0) return ORC_BAD_PUBKEY;
	lhs = fp_sqr(L, y);
	t = fp_add(L, fp_add(L, x, x), x);
	rhs = fp_add(L, fp_sub(L, fp_mul(L, fp_sqr(L, x), x), t), L->b);
	return fe_cmp(lhs, rhs) == 0 ? ORC_OK : ORC_BAD_PUBKEY;
}

/* bign_misc.c:437-500: key <- the first key_len octets of (K.x || K.y), K = d Q */
u32 orc_bignDH(size_t l, u8* key, const u8* privkey, const u8* pubkey, size_t key_len)
{
	const lvl* L = level(l);
	fe d, x, y;
	u8 xy[128];
	u32 code;
	if (!L) return 119u;
	if (key_len > 2 * L->no) return 507u;
	d = fe_from(L, privkey);
	if (fe_is0(d) || fe_cmp(d, L->q) >= 0) return ORC_BAD_PRIVKEY;
	if ((code = orc_bignPubkeyVal(l, pubkey))) return code;
	if (!pt_to_affine(L, &x, &y, pt_mul(L, pt_affine(fe_from(L, pubkey), fe_from(L, pubkey + L->no)), privkey, L->no)))
		return ORC_BAD_PARAMS;
	fe_to(L, xy, x), fe_to(L, xy + L->no, y);
	memcpy(key, xy, key_len);
	return ORC_OK;
}

/* the part bignSign and bignSign2 share once the one-time key k is known (bign_sign.c:92-122, :219-243):
   R = k G, s0 = belt-hash(oid || R.x || H)[0..no/2), s1 = (k - (s0 + 2^l) d - H) mod q */
static u32 sign_with_k(const lvl* L, u8* sig, const u8* oid_der, size_t oid_len, const u8* hash, fe d, const u8* kb)
{
	const size_t no = L->no;
	fe k = fe_from(L, kb), x, y, s0d, s1, Hh, s0w;
	u8* buf;
	u8 hv[32];
	u64 prod[2 * FE_MAXW];
	if (!pt_to_affine(L, &x, &y, pt_mul(L, pt_base(L), kb, no))) return ORC_BAD_PARAMS;
	buf = (u8*)malloc(oid_len + 2 * no + 1);
	if (!buf) return 110u;
	memcpy(buf, oid_der, oid_len), fe_to(L, buf + oid_len, x), memcpy(buf + oid_len + no, hash, no);
	orc_beltHash(hv, buf, oid_len + 2 * no);
	free(buf);
	memcpy(sig, hv, no / 2);
	memset(&s0w, 0, sizeof s0w), memcpy(s0w.w, hv, no / 2), s0w.w[L->n / 2] = 1;
	mul_wide(prod, s0w.w, L->n, d.w, L->n);
	s0d = fold_mod(L, prod, L->q);
	s1 = submod(L, k, s0d, L->q);
	Hh = fe_from(L, hash);           /* not reduced first: bign_sign.c:236-237 */
	s1 = submod(L, s1, Hh, L->q);
	fe_to(L, sig + no / 2, s1);
	return ORC_OK;
}

/* bign_sign.c:27-125 with the one-time key k (0 < k < q, what zzRandNZMod drew from the generator) given */
u32 orc_bignSignK(size_t l, u8* sig, const u8* oid_der, size_t oid_len, const u8* hash,
	const u8* privkey, const u8* k)
{
	const lvl* L = level(l);
	fe d;
	if (!L) return 119u;
	d = fe_from(L, privkey);
	if (fe_is0(d) || fe_cmp(d, L->q) >= 0) return ORC_BAD_PRIVKEY;
	return sign_with_k(L, sig, oid_der, oid_len, hash, d, k);
}

/* bign_sign.c:140-245 */
u32 orc_bignSign2(size_t l, u8* sig, const u8* oid_der, size_t oid_len, const u8* hash,
	const u8* privkey, const void* t, size_t t_len)
{
	const lvl* L = level(l);
	size_t no;
	fe d, k;
	u8* buf;
	u8 theta[32], kb[64];
	u32 tk[8];
	if (!L) return 119u;
	no = L->no;
	d = fe_from(L, privkey);
	if (fe_is0(d) || fe_cmp(d, L->q) >= 0) return ORC_BAD_PRIVKEY;
	buf = (u8*)malloc(oid_len + no + t_len + 1);
	if (!buf) return 110u;
	/* theta = belt-hash(oid || d || t) */
	memcpy(buf, oid_der, oid_len), memcpy(buf + oid_len, privkey, no);
	if (t) memcpy(buf + oid_len + no, t, t_len);
	orc_beltHash(theta, buf, oid_len + no + (t ? t_len : 0));
	free(buf);
	orc_beltKeyExpand2(tk, theta, 32);
	/* k = H; k = WBL(k) until 0 < k < q */
	memcpy(kb, hash, no);
	do belt_wbl(kb, no, tk), k = fe_from(L, kb);
	while (fe_is0(k) || fe_cmp(k, L->q) >= 0);
	return sign_with_k(L, sig, oid_der, oid_len, hash, d, kb);
}
u32 orc_bignSign2_128(u8 sig[48], const u8* oid_der, size_t oid_len, const u8 hash[32],
	const u8 privkey[32], const void* t, size_t t_len)
{
	return orc_bignSign2(128, sig, oid_der, oid_len, hash, privkey, t, t_len);
}

/* ======================================================================= belt-DWP (belt_dwp.c:45-330) */

/* c = a * b in GF(2)[x]/(x^128 + x^7 + x^2 + x + 1); bit i of the 128-bit LE integer is the
   coefficient of x^i (belt_lcl.c:119-132: ppMul + ppRedBelt, pp_red.c:129-141) */
static void gf128_mul(u64 c[2], const u64 a[2], const u64 b[2])
{
	u64 r0 = 0, r1 = 0, v0 = b[0], v1 = b[1];
	int i;
	for (i = 0; i < 128; ++i)
	{
		if (a[i / 64] >> (i % 64) & 1)
			r0 ^= v0, r1 ^= v1;
		{
			/* v <- v * x mod f */
			const u64 top = v1 >> 63;
			v1 = v1 << 1 | v0 >> 63;
			v0 = v0 << 1 ^ (top ? 0x87 : 0);
		}
	}
	c[0] = r0, c[1] = r1;
}

/* mac state after absorbing `n` octets of one data class (zero-padded to whole blocks) */
static void dwp_absorb(u64 t[2], const u64 r[2], const u8* p, size_t n)
{
	while (n)
	{
		u8 blk[16] = {0};
		u64 x[2];
		const size_t take = n < 16 ? n : 16;
		memcpy(blk, p, take), memcpy(x, blk, 16);
		t[0] ^= x[0], t[1] ^= x[1];
		gf128_mul(t, t, r);
		p += take, n -= take;
	}
}

static void dwp_mac(u8 mac[8], const u32 key[8], const u32 s[4], const u8* crit, size_t n1, const u8* open, size_t n2)
{
	u64 r[2], t[2], len[2];
	u32 w[4];
	memcpy(w, s, 16);
	orc_beltBlockEncr2(w, key);          /* r = E_K(s), s = E_K(iv) (belt_dwp.c:52-55) */
	memcpy(r, w, 16);
	memcpy(t, orc_beltH(), 16);
	dwp_absorb(t, r, open, n2);
	dwp_absorb(t, r, crit, n1);
	len[0] = (u64)n2 << 3, len[1] = (u64)n1 << 3;
	t[0] ^= len[0], t[1] ^= len[1];
	gf128_mul(t, t, r);
	memcpy(w, t, 16);
	orc_beltBlockEncr2(w, key);
	memcpy(mac, w, 8);
}

/* belt_dwp.c:250-287: dest = E(src1), mac over (src2 open, dest critical) */
u32 orc_beltDWPWrap(void* dest, u8 mac[8], const void* src1, size_t n1, const void* src2, size_t n2,
	const u8* key, size_t len, const u8 iv[16])
{
	orc_belt_ctr_st st;
	u32 s[4];
	if (len != 16 && len != 24 && len != 32)
		return ORC_BAD_INPUT;
	orc_beltCTRStart(&st, key, len, iv);
	memcpy(s, st.ctr, 16);
	memmove(dest, src1, n1);
	orc_beltCTRStepE(dest, n1, &st);
	dwp_mac(mac, st.key, s, (const u8*)dest, n1, (const u8*)src2, n2);
	return ORC_OK;
}

/* belt_dwp.c:289-330: returns 511 (ERR_BAD_MAC) without touching dest when the tag differs */
u32 orc_beltDWPUnwrap(void* dest, const void* src1, size_t n1, const void* src2, size_t n2,
	const u8 mac[8], const u8* key, size_t len, const u8 iv[16])
{
	orc_belt_ctr_st st;
	u32 s[4];
	u8 m[8];
	if (len != 16 && len != 24 && len != 32)
		return ORC_BAD_INPUT;
	orc_beltCTRStart(&st, key, len, iv);
	memcpy(s, st.ctr, 16);
	dwp_mac(m, st.key, s, (const u8*)src1, n1, (const u8*)src2, n2);
	if (memcmp(m, mac, 8) != 0)
		return 511u;
	memmove(dest, src1, n1);
	orc_beltCTRStepE(dest, n1, &st);
	return ORC_OK;
}

/* ======================================================================= belt-CHE (belt_che.c:48-330) */

/* keystream of CHE: s_0 = E_K(iv); s_j = s_{j-1} * x ^ 1 (belt_lcl.c:99-108, belt_che.c:89);
   block j (j >= 1) of the gamma is E_K(s_j) */
static void che_crypt(u8* buf, size_t n, const u32 key[8], const u32 s0[4])
{
	u64 s[2];
	memcpy(s, s0, 16);
	while (n)
	{
		u32 w[4];
		u8 g[16];
		const size_t take = n < 16 ? n : 16;
		size_t i;
		const u64 top = s[1] >> 63;
		s[1] = s[1] << 1 | s[0] >> 63;
		s[0] = (s[0] << 1 ^ (top ? 0x87 : 0)) ^ 1;
		memcpy(w, s, 16);
		orc_beltBlockEncr2(w, key);
		memcpy(g, w, 16);
		for (i = 0; i < take; ++i) buf[i] ^= g[i];
		buf += take, n -= take;
	}
}

static void che_mac(u8 mac[8], const u32 key[8], const u32 r32[4], const u8* crit, size_t n1, const u8* open, size_t n2)
{
	u64 r[2], t[2];
	u32 w[4];
	memcpy(r, r32, 16);                   /* r = E_K(iv) (belt_che.c:54-57) */
	memcpy(t, orc_beltH(), 16);
	dwp_absorb(t, r, open, n2);
	dwp_absorb(t, r, crit, n1);
	t[0] ^= (u64)n2 << 3, t[1] ^= (u64)n1 << 3;
	gf128_mul(t, t, r);
	memcpy(w, t, 16);
	orc_beltBlockEncr2(w, key);
	memcpy(mac, w, 8);
}

u32 orc_beltCHEWrap(void* dest, u8 mac[8], const void* src1, size_t n1, const void* src2, size_t n2,
	const u8* key, size_t len, const u8 iv[16])
{
	u32 k[8], r[4];
	if (len != 16 && len != 24 && len != 32)
		return ORC_BAD_INPUT;
	orc_beltKeyExpand2(k, key, len);
	memcpy(r, iv, 16);
	orc_beltBlockEncr2(r, k);
	memmove(dest, src1, n1);
	che_crypt((u8*)dest, n1, k, r);
	che_mac(mac, k, r, (const u8*)dest, n1, (const u8*)src2, n2);
	return ORC_OK;
}

u32 orc_beltCHEUnwrap(void* dest, const void* src1, size_t n1, const void* src2, size_t n2,
	const u8 mac[8], const u8* key, size_t len, const u8 iv[16])
{
	u32 k[8], r[4];
	u8 m[8];
	if (len != 16 && len != 24 && len != 32)
		return ORC_BAD_INPUT;
	orc_beltKeyExpand2(k, key, len);
	memcpy(r, iv, 16);
	orc_beltBlockEncr2(r, k);
	che_mac(m, k, r, (const u8*)src1, n1, (const u8*)src2, n2);
	if (memcmp(m, mac, 8) != 0)
		return 511u;
	memmove(dest, src1, n1);
	che_crypt((u8*)dest, n1, k, r);
	return ORC_OK;
}

/* ======================================================================= belt-DWP / belt-CHE, streaming
   (belt_dwp.c:45-207, belt_che.c:48-239): one state for both modes; `che` selects the LFSR counter
   and r = E_K(iv) (belt-CHE) instead of the incrementing counter and r = E_K(E_K(iv)) (belt-DWP). */

void orc_beltAEADStart(orc_belt_aead_st* st, int che, const u8* key, size_t len, const u8 iv[16])
{
	u32 w[4];
	memset(st, 0, sizeof *st);
	st->che = che;
	orc_beltKeyExpand2(st->key, key, len);
	memcpy(w, iv, 16);
	orc_beltBlockEncr2(w, st->key);                 /* s = E_K(iv) */
	memcpy(st->s, w, 16);
	if (!che)
		orc_beltBlockEncr2(w, st->key);             /* DWP: r = E_K(s) (belt_dwp.c:52-55); CHE: r = s */
	memcpy(st->r, w, 16);
	memcpy(st->t, orc_beltH(), 16);
}

/* StepE == StepD: XOR with the keystream, block by block, keeping the unused rest of the last block */
void orc_beltAEADStepE(void* buf, size_t n, orc_belt_aead_st* st)
{
	u8* p = (u8*)buf;
	while (n)
	{
		if (!st->reserved)
		{
			u32 w[4];
			if (st->che)
			{
				/* s <- s x ^ 1 (belt_che.c:86-88) */
				u64 v[2];
				u64 top;
				memcpy(v, st->s, 16);
				top = v[1] >> 63;
				v[1] = v[1] << 1 | v[0] >> 63;
				v[0] = (v[0] << 1 ^ (top ? 0x87 : 0)) ^ 1;
				memcpy(st->s, v, 16);
			}
			else
			{
				/* s <- s + 1 as a 128-bit little-endian integer (belt_ctr.c:27-35) */
				int i;
				for (i = 0; i < 4 && ++st->s[i] == 0; ++i)
					;
			}
			memcpy(w, st->s, 16);
			orc_beltBlockEncr2(w, st->key);
			memcpy(st->ks, w, 16);
			st->reserved = 16;
		}
		*p++ ^= st->ks[16 - st->reserved];
		--st->reserved, --n;
	}
}

static void aead_block(orc_belt_aead_st* st, u64 t[2], const u8 blk[16])
{
	u64 x[2];
	memcpy(x, blk, 16);
	t[0] ^= x[0], t[1] ^= x[1];
	gf128_mul(t, t, st->r);
}

static void aead_absorb(orc_belt_aead_st* st, const u8* p, size_t n)
{
	while (n)
	{
		const size_t take = n < 16 - st->filled ? n : 16 - st->filled;
		memcpy(st->block + st->filled, p, take);
		st->filled += take, p += take, n -= take;
		if (st->filled == 16)
			aead_block(st, st->t, st->block), st->filled = 0;
	}
}

void orc_beltAEADStepI(const void* buf, size_t n, orc_belt_aead_st* st)
{
	st->len[0] += (u64)n << 3;
	aead_absorb(st, (const u8*)buf, n);
}

void orc_beltAEADStepA(const void* buf, size_t n, orc_belt_aead_st* st)
{
	/* first non-empty critical fragment closes the open data with zeros (belt_dwp.c:121-131) */
	if (n && st->len[1] == 0 && st->filled)
	{
		memset(st->block + st->filled, 0, 16 - st->filled);
		aead_block(st, st->t, st->block), st->filled = 0;
	}
	st->len[1] += (u64)n << 3;
	aead_absorb(st, (const u8*)buf, n);
}

/* the state itself is not changed: more data may follow (belt_dwp.c:172-196) */
void orc_beltAEADStepG(u8 mac[8], const orc_belt_aead_st* cst)
{
	orc_belt_aead_st* st = (orc_belt_aead_st*)cst;
	u64 t1[2];
	u32 w[4];
	u8 blk[16];
	memcpy(t1, st->t, 16);
	if (st->filled)
	{
		memset(blk, 0, 16), memcpy(blk, st->block, st->filled);
		aead_block(st, t1, blk);
	}
	memcpy(blk, st->len, 16);
	aead_block(st, t1, blk);
	memcpy(w, t1, 16);
	orc_beltBlockEncr2(w, st->key);
	memcpy(mac, w, 8);
}

/* ======================================================================= bash-prg (bash_prg.c:56-385) */
/* programmable sponge automaton; own state layout (the semantics of bash_prg_st, :54-62) */

static void prg_commit(orc_bash_prg_st* st, u8 code)   /* bash_prg.c:92-105 */
{
	st->s[st->pos] ^= code;
	st->s[st->buf_len] ^= 0x80;
	orc_bashF(st->s);
	st->pos = 0;
}

void orc_bashPrgStart(orc_bash_prg_st* st, size_t l, size_t d, const u8* ann, size_t ann_len,
	const u8* key, size_t key_len)
{
	memset(st->s, 0, 192);
	st->pos = 1 + ann_len + key_len;
	st->s[0] = (u8)(ann_len * 4 + key_len / 4);
	if (ann_len) memcpy(st->s + 1, ann, ann_len);
	if (key_len) memcpy(st->s + 1 + ann_len, key, key_len);
	st->s[184] = (u8)(l / 4 + d);
	st->buf_len = key_len ? 192 - l * (2 + d) / 16 : 192 - d * l / 4;
	st->l = l, st->d = d;
}

void orc_bashPrgRestart(const u8* ann, size_t ann_len, const u8* key, size_t key_len, orc_bash_prg_st* st)
{
	size_t i;
	if (key_len)
		prg_commit(st, 0x05), st->buf_len = 192 - st->l * (2 + st->d) / 16;
	else
		prg_commit(st, 0x01);
	st->pos = 1 + ann_len + key_len;
	st->s[0] ^= (u8)(ann_len * 4 + key_len / 4);
	for (i = 0; i < ann_len; ++i) st->s[1 + i] ^= ann[i];
	for (i = 0; i < key_len; ++i) st->s[1 + ann_len + i] ^= key[i];
}

/* mode: 0 absorb (s ^= in), 1 squeeze (out = s), 2 encr (s ^= buf, buf = s), 3 decr (buf ^= s, s ^= buf) */
static void prg_step(orc_bash_prg_st* st, u8* buf, size_t n, int mode)
{
	while (n)
	{
		size_t take = st->buf_len - st->pos, i;
		if (take > n) take = n;
		for (i = 0; i < take; ++i)
		{
			u8* s = st->s + st->pos + i;
			switch (mode)
			{
			case 0: *s ^= buf[i]; break;
			case 1: buf[i] = *s; break;
			case 2: *s ^= buf[i], buf[i] = *s; break;
			default: buf[i] ^= *s, *s ^= buf[i]; break;
			}
		}
		st->pos += take, buf += take, n -= take;
		if (st->pos == st->buf_len)
			orc_bashF(st->s), st->pos = 0;
	}
}

void orc_bashPrgAbsorbStart(orc_bash_prg_st* st) { prg_commit(st, 0x09); }
void orc_bashPrgAbsorbStep(const void* buf, size_t n, orc_bash_prg_st* st) { prg_step(st, (u8*)(size_t)buf, n, 0); }
void orc_bashPrgSqueezeStart(orc_bash_prg_st* st) { prg_commit(st, 0x11); }
void orc_bashPrgSqueezeStep(void* buf, size_t n, orc_bash_prg_st* st) { prg_step(st, (u8*)buf, n, 1); }
void orc_bashPrgEncrStart(orc_bash_prg_st* st) { prg_commit(st, 0x0D); }
void orc_bashPrgEncrStep(void* buf, size_t n, orc_bash_prg_st* st) { prg_step(st, (u8*)buf, n, 2); }
void orc_bashPrgDecrStart(orc_bash_prg_st* st) { prg_commit(st, 0x0D); }
void orc_bashPrgDecrStep(void* buf, size_t n, orc_bash_prg_st* st) { prg_step(st, (u8*)buf, n, 3); }
void orc_bashPrgRatchet(orc_bash_prg_st* st)           /* bash_prg.c:374-385 */
{
	u8 t[192];
	int i;
	memcpy(t, st->s, 192);
	prg_commit(st, 0x01);
	for (i = 0; i < 192; ++i) st->s[i] ^= t[i];
}
