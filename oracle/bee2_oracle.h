/*
 * bee2_oracle.h — CPU restatement of the bee2 hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the checker, not the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The shipped library
 * (bee2_b200/csrc) never links, loads or calls anything in oracle/.
 *
 * Every function names the reference file:line (agievich/bee2 @ d9e689a0) it follows.
 * Parity is PINNED: tests/test_oracle_kat.py checks this file against every STB annex
 * vector the reference's own tests hold for the path (tests/golden/kat.json) and, where
 * oracle/_ref/libbee2ref_64.so is present, against the unmodified reference on seeded
 * random inputs.
 */
#ifndef BEE2_ORACLE_H
#define BEE2_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* err_t values, include/bee2/core/err.h:72,132,180,184,186,196 */
#define ORC_OK 0u
#define ORC_BAD_INPUT 109u
#define ORC_BAD_PARAMS 502u
#define ORC_BAD_PRIVKEY 504u
#define ORC_BAD_PUBKEY 505u
#define ORC_BAD_SIG 510u

/* ---- bash (STB 34.101.77) ---- */
void orc_bashF(uint8_t block[192]);                                         /* bash_f64.c:142-187 */
uint32_t orc_bashHash(uint8_t* hash, size_t l, const void* src, size_t n);  /* bash_hash.c:118-137 */
/* streaming: state = 192 s | 8 rate | 8 pos (own layout; bash_hash.c:25-31 semantics) */
typedef struct { uint8_t s[192]; size_t rate; size_t pos; } orc_bash_st;
void orc_bashHashStart(orc_bash_st* st, size_t l);                          /* bash_hash.c:38-50 */
void orc_bashHashStepH(const void* buf, size_t n, orc_bash_st* st);         /* bash_hash.c:52-79 */
void orc_bashHashStepG(uint8_t* hash, size_t hash_len, const orc_bash_st*); /* bash_hash.c:81-109 */

/* programmable automaton bash-prg (bash_prg.c:56-385) */
typedef struct { size_t l, d; uint8_t s[192]; size_t buf_len, pos; } orc_bash_prg_st;
void orc_bashPrgStart(orc_bash_prg_st* st, size_t l, size_t d, const uint8_t* ann, size_t ann_len,
	const uint8_t* key, size_t key_len);
void orc_bashPrgRestart(const uint8_t* ann, size_t ann_len, const uint8_t* key, size_t key_len, orc_bash_prg_st* st);
void orc_bashPrgAbsorbStart(orc_bash_prg_st* st);
void orc_bashPrgAbsorbStep(const void* buf, size_t n, orc_bash_prg_st* st);
void orc_bashPrgSqueezeStart(orc_bash_prg_st* st);
void orc_bashPrgSqueezeStep(void* buf, size_t n, orc_bash_prg_st* st);
void orc_bashPrgEncrStart(orc_bash_prg_st* st);
void orc_bashPrgEncrStep(void* buf, size_t n, orc_bash_prg_st* st);
void orc_bashPrgDecrStart(orc_bash_prg_st* st);
void orc_bashPrgDecrStep(void* buf, size_t n, orc_bash_prg_st* st);
void orc_bashPrgRatchet(orc_bash_prg_st* st);

/* ---- belt (STB 34.101.31) ---- */
const uint8_t* orc_beltH(void);                                             /* belt_block.c:43-65 */
void orc_beltKeyExpand2(uint32_t key_[8], const uint8_t* key, size_t len);  /* belt_block.c:88-106 */
void orc_beltBlockEncr2(uint32_t block[4], const uint32_t key[8]);          /* belt_block.c:324-328 */
void orc_beltBlockDecr2(uint32_t block[4], const uint32_t key[8]);          /* belt_block.c:362-366 */
typedef struct { uint32_t key[8]; uint32_t ctr[4]; uint8_t block[16]; size_t reserved; } orc_belt_ctr_st; /* belt_lcl.h:135-141 */
void orc_beltCTRStart(orc_belt_ctr_st* st, const uint8_t* key, size_t len, const uint8_t iv[16]); /* belt_ctr.c:55-64 */
void orc_beltCTRStepE(void* buf, size_t n, orc_belt_ctr_st* st);            /* belt_ctr.c:66-111 */
uint32_t orc_beltCTR(void* dst, const void* src, size_t n, const uint8_t* key, size_t len, const uint8_t iv[16]); /* belt_ctr.c:113-135 */
uint32_t orc_beltECBEncr(void* dst, const void* src, size_t n, const uint8_t* key, size_t len);  /* belt_ecb.c:112-134 */
uint32_t orc_beltECBDecr(void* dst, const void* src, size_t n, const uint8_t* key, size_t len);  /* belt_ecb.c:136-158 */
void orc_beltHash(uint8_t hash[32], const void* src, size_t n);             /* belt_hash.c:174-190 */
/* belt-DWP AEAD (belt_dwp.c:250-330); Unwrap returns 511 (ERR_BAD_MAC) on a wrong tag */
uint32_t orc_beltDWPWrap(void* dest, uint8_t mac[8], const void* src1, size_t n1, const void* src2, size_t n2,
	const uint8_t* key, size_t len, const uint8_t iv[16]);
uint32_t orc_beltDWPUnwrap(void* dest, const void* src1, size_t n1, const void* src2, size_t n2,
	const uint8_t mac[8], const uint8_t* key, size_t len, const uint8_t iv[16]);
/* key-agility batch: block i under key i (config 5) */
void orc_beltECBEncrMultiKey(uint8_t* blocks, const uint8_t* keys32, size_t count);

/* ---- belt-DWP / belt-CHE streaming (belt_dwp.c:45-207, belt_che.c:48-239); che = 0: DWP, 1: CHE ---- */
typedef struct
{
	uint32_t key[8], s[4];
	uint64_t r[2], t[2], len[2];
	uint8_t block[16];
	size_t filled;
	uint8_t ks[16];
	size_t reserved;
	int che;
} orc_belt_aead_st;
void orc_beltAEADStart(orc_belt_aead_st* st, int che, const uint8_t* key, size_t len, const uint8_t iv[16]);
void orc_beltAEADStepE(void* buf, size_t n, orc_belt_aead_st* st);      /* also StepD */
void orc_beltAEADStepI(const void* buf, size_t n, orc_belt_aead_st* st);
void orc_beltAEADStepA(const void* buf, size_t n, orc_belt_aead_st* st);
void orc_beltAEADStepG(uint8_t mac[8], const orc_belt_aead_st* st);

/* ---- bign on bign-curve256v1 (STB 34.101.45, l = 128) ---- */
uint32_t orc_bignVerify128(const uint8_t* oid_der, size_t oid_len, const uint8_t hash[32],
	const uint8_t sig[48], const uint8_t pubkey[64]);                       /* bign_sign.c:268-347 */
uint32_t orc_bignSign2_128(uint8_t sig[48], const uint8_t* oid_der, size_t oid_len,
	const uint8_t hash[32], const uint8_t privkey[32], const void* t, size_t t_len); /* bign_sign.c:140-245 */
uint32_t orc_bignPubkeyCalc128(uint8_t pubkey[64], const uint8_t privkey[32]); /* bign_misc.c:369-412 */
/* d * A on the curve; returns 1, or 0 when the result is the point at infinity. ec.c:497-525 */
int orc_ecMulA128(uint8_t b[64], const uint8_t a[64], const uint8_t* d, size_t d_len);
/* the same on the standard curve of level l = 128 / 192 / 256 (bign-curve256v1 / 384v1 / 512v1,
   bign_params.c:36-190); no = l/4 octets: hash no, sig no/2 + no, privkey no, pubkey 2 no */
uint32_t orc_bignVerify(size_t l, const uint8_t* oid_der, size_t oid_len, const uint8_t* hash,
	const uint8_t* sig, const uint8_t* pubkey);
uint32_t orc_bignSign2(size_t l, uint8_t* sig, const uint8_t* oid_der, size_t oid_len,
	const uint8_t* hash, const uint8_t* privkey, const void* t, size_t t_len);
uint32_t orc_bignPubkeyCalc(size_t l, uint8_t* pubkey, const uint8_t* privkey);
/* bignSign (bign_sign.c:27-125) with the one-time key the generator produced */
uint32_t orc_bignSignK(size_t l, uint8_t* sig, const uint8_t* oid_der, size_t oid_len,
	const uint8_t* hash, const uint8_t* privkey, const uint8_t* k);
int orc_ecMulA(size_t l, uint8_t* b, const uint8_t* a, const uint8_t* d, size_t d_len);
uint32_t orc_bignPubkeyVal(size_t l, const uint8_t* pubkey);                        /* bign_misc.c:317-352 */
uint32_t orc_bignDH(size_t l, uint8_t* key, const uint8_t* privkey, const uint8_t* pubkey, size_t key_len); /* :437-500 */
/* field helpers exposed for unit tests of the device field layer (zm.c:214-253, gfp.c:33-44) */
void orc_gfpMul(uint8_t c[32], const uint8_t a[32], const uint8_t b[32]);
void orc_gfpInv(uint8_t c[32], const uint8_t a[32]);

#ifdef __cplusplus
}
#endif
#endif
