/*
 * bee2_b200.h — C ABI of the B200-native batch engine for the bee2 hot path.
 *
 * Three groups of entry points, all `extern "C"`, plain pointers and sizes:
 *
 *  1. DROP-IN symbols: same names, prototypes, state layouts and err_t behaviour as
 *     the reference headers (include/bee2/crypto/{bash,belt,bign}.h). A program that
 *     links libbee2_b200.so in place of libbee2 for these symbols gets the GPU path.
 *     Each declaration cites the reference declaration it replaces.
 *  2. BATCH symbols (`...Batch`): host pointers in/out, many independent units per
 *     call — what the reference-side binding calls for throughput.
 *  3. DEVICE symbols (`b2g_*_dev`): the same work on buffers already resident in HBM,
 *     on a caller-supplied CUDA stream (NULL = default stream). Used by bench.py and
 *     by multi-GPU drivers that keep data on the device.
 *
 * Every function here runs on the GPU. There is no CPU fallback: if no CUDA device
 * is usable, `err_t` functions return ERR_B2G_NO_DEVICE and `void` functions abort()
 * with a message on stderr.
 *
 * Conventions (identical to the reference): all multi-byte integers are
 * little-endian octet strings; the caller owns every buffer; streaming states are
 * caller-allocated flat structs of exactly X_keep() bytes and may be memcpy'd; every
 * function is re-entrant on distinct states (the engine serialises device access with
 * an internal mutex).
 */
#ifndef BEE2_B200_H
#define BEE2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- basic types: include/bee2/defs.h:269,372,441,463 ---- */
typedef uint8_t octet;
typedef uint32_t u32;
typedef uint64_t u64;
typedef uint64_t word;
typedef int bool_t;
typedef u32 err_t;

/* ---- error codes: include/bee2/core/err.h:72,74,92,112,132,138,180,184,186,190,196 ---- */
#define ERR_OK 0u
#define ERR_BAD_INPUT 109u
#define ERR_OUTOFMEMORY 110u
#define ERR_NOT_IMPLEMENTED 119u
#define ERR_FILE_NOT_FOUND 202u
#define ERR_FILE_OPEN 203u
#define ERR_FILE_READ 207u
#define ERR_BAD_OID 301u
#define ERR_BAD_RNG 304u
#define ERR_BAD_PARAMS 502u
#define ERR_BAD_PRIVKEY 504u
#define ERR_BAD_PUBKEY 505u
#define ERR_BAD_SHAREDKEY 507u
#define ERR_BAD_SIG 510u
#define ERR_BAD_MAC 511u
/* engine-specific (outside the reference's ranges) */
#define ERR_B2G_NO_DEVICE 9001u   /* no usable CUDA device / driver */
#define ERR_B2G_CUDA 9002u        /* a CUDA call failed; see b2g_last_error() */

/* ======================================================================= engine */
/* Bring the engine up on CUDA device `device` and make it current on the calling thread
   (< 0: the thread's current device). The first device initialised is the PRIMARY one: host-pointer
   calls from threads whose current CUDA device is not an initialised one run there. Idempotent;
   may be called for several devices (a caller that drives each device from its own thread). */
err_t b2g_init(int device);
/* In-process multi-device mode: bring the engine up on devices 0..n-1 (n <= 0: all visible) and
   shard the units of the host-pointer *Batch entry points (bashHashBatch, beltCTRKeystream,
   beltECBEncrBatch, bignVerifyBatch, bignSign2Batch, bignPubkeyCalcBatch) over them — one host
   thread and stream set per device, contiguous shares, results byte-identical to one device. */
err_t b2g_init_devices(int n);
/* Number of devices the Batch entry points shard over (1 unless b2g_init_devices was called). */
int b2g_device_count(void);
/* Human-readable text of the last CUDA failure on this thread ("" if none). */
const char* b2g_last_error(void);
/* Number of SMs of the active device (grid sizing is a multiple of it). */
int b2g_sm_count(void);
/* Pinned host memory for full-rate PCIe transfers through the Batch entry points. */
void* b2g_host_alloc(size_t bytes);
void b2g_host_free(void* p);
/* Plain device memory helpers for callers without their own allocator. */
void* b2g_dev_alloc(size_t bytes);
void b2g_dev_free(void* p);
err_t b2g_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes);
err_t b2g_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes);
err_t b2g_sync(void);
/* One process per GPU (torchrun): map a b2g_dev_alloc'ed buffer of another rank into this process
   (CUDA IPC, peer access over NVLink enabled on open). A `b2g_*_dev` launcher given such a pointer
   as its output writes straight into the peer's HBM — compute and the final gather in ONE kernel. */
err_t b2g_ipc_export(octet handle[64], void* dptr);
err_t b2g_ipc_open(void** dptr, const octet handle[64]);
err_t b2g_ipc_close(void* dptr);
/* cudaMemcpyAsync on a caller's stream (to_device != 0: H2D, else D2H) */
err_t b2g_memcpy_async(void* dst, const void* src, size_t n, int to_device, void* stream);
/* Overlay mode. With a stock libbee2 BEHIND this library in the symbol search order (link with
   `-lbee2_b200 -lbee2`, or LD_PRELOAD this library into a bee2 application) the drop-in entry
   points forward to it — dlsym(RTLD_NEXT, name) — for inputs the GPU path does not cover
   (non-standard bign_params, a generic ec_o, scalars longer than the field, OID / t > 64 octets),
   for one-shot calls whose payload is below `bytes` (default 0 = never; also B2G_CPU_BELOW in the
   environment), and when the GPU path of a `void` function fails (instead of abort()). A process that
   dlopen()s this library names the stock library in B2G_STOCK_LIB instead (loaded privately). */
void b2g_set_cpu_below(size_t bytes);
int b2g_has_stock(void);           /* 1 if a stock libbee2 is reachable behind this library */
u64 b2g_forward_count(void);       /* calls forwarded to it so far */
/* How many kernels this library has launched in this process (for gpu_launches). */
u64 b2g_launch_count(void);
/* Measured issue peak of one instruction kind, lane-operations per second on the whole chip:
   0 LOP3, 1 SHF, 2 PRMT, 3 IADD, 4 IMAD, 5 IMAD.WIDE, 6 LDS.32 (conflict-free), 10 IMAD.HI, 12 FFMA;
   co-issue mixes (two independent chains, both counted): 7 LOP3+IMAD.WIDE, 8 LOP3+IMAD,
   9 LOP3+FFMA, 11 LOP3+LDS.32. iters = 0 picks a default. < 0 on failure.
   Denominators of the issue roofline. */
double b2g_microbench(int kind, unsigned iters);

/* ======================================================================= bash (STB 34.101.77) */
/* drop-in: include/bee2/crypto/bash.h:127-139 (bash_f64.c:174-192) */
extern const char bash_platform[];                       /* bash_f.c:28-43 -> "BASH_CUDA_SM100A" */
size_t bashF_deep(void);
void bashF(octet block[192], void* stack);
/* drop-in: bash.h:152-225 (bash_hash.c:25-137); state layout = bash_hash_st */
size_t bashHash_keep(void);
void bashHashStart(void* state, size_t l);
void bashHashStepH(const void* buf, size_t count, void* state);
void bashHashStepG(octet hash[], size_t hash_len, void* state);
bool_t bashHashStepV(const octet hash[], size_t hash_len, void* state);
err_t bashHash(octet hash[], size_t l, const void* src, size_t count);
/* drop-in: bash.h (bash_prg.c:54-385) — the programmable sponge automaton; state layout =
   bash_prg_st. A single automaton is sequential by construction: these calls run its
   permutations on the device one after another (correct, latency-bound). */
size_t bashPrg_keep(void);
void bashPrgStart(void* state, size_t l, size_t d, const octet ann[], size_t ann_len,
	const octet key[], size_t key_len);
void bashPrgRestart(const octet ann[], size_t ann_len, const octet key[], size_t key_len, void* state);
void bashPrgAbsorbStart(void* state);
void bashPrgAbsorbStep(const void* buf, size_t count, void* state);
void bashPrgAbsorb(const void* buf, size_t count, void* state);
void bashPrgSqueezeStart(void* state);
void bashPrgSqueezeStep(void* buf, size_t count, void* state);
void bashPrgSqueeze(void* buf, size_t count, void* state);
void bashPrgEncrStart(void* state);
void bashPrgEncrStep(void* buf, size_t count, void* state);
void bashPrgEncr(void* buf, size_t count, void* state);
void bashPrgDecrStart(void* state);
void bashPrgDecrStep(void* buf, size_t count, void* state);
void bashPrgDecr(void* buf, size_t count, void* state);
void bashPrgRatchet(void* state);
/* device: `count` automata (192-octet states, 8-aligned) each run bash-f (if pre_f) and then
   nblocks whole buf_len-octet blocks of one command over its data at d_data + i*stride, in place:
   mode 0 absorb, 1 squeeze, 2 encr, 3 decr */
err_t b2g_bashPrgBlocks_dev(void* d_states, void* d_data, size_t stride, size_t nblocks,
	size_t buf_len, int mode, int pre_f, size_t count, void* stream);
/* batch: `count` messages of msg_len octets, message i at msgs + i*stride; digest i
   (l/4 octets) at hashes + i*(l/4). Same checks as bashHash. */
err_t bashHashBatch(octet* hashes, size_t l, const void* msgs, size_t msg_len,
	size_t stride, size_t count);
/* ragged batch (the many-files case of cmd/bsum/bsum.c:142-200): message i is
   data[offsets[i] .. offsets[i] + lens[i]); digest i at hashes + i*(l/4). Sort by length for
   best warp efficiency (a warp runs as long as its longest message). */
err_t bashHashBatchV(octet* hashes, size_t l, const void* data, size_t data_len, const u64* offsets,
	const u64* lens, size_t count);
/* batch: bashF on `count` independent 192-octet states, in place */
err_t bashFBatch(octet* blocks, size_t count);
/* device */
/* the bsum case (cmd/bsum/bsum.c:142-200): digest of each of `count` files; whole files are staged in a
   pinned buffer and hashed by ragged batch launches, files above 64 MiB are streamed.
   status[i] = ERR_OK / ERR_FILE_OPEN / ERR_FILE_READ */
err_t bashHashFiles(octet* hashes, err_t* status, size_t l, const char* const paths[], size_t count);
err_t b2g_bashHashBatch_dev(void* d_hashes, size_t l, const void* d_msgs, size_t msg_len,
	size_t stride, size_t count, void* stream);
err_t b2g_bashHashBatchV_dev(void* d_hashes, size_t l, const void* d_data, const void* d_offsets,
	const void* d_lens, size_t count, void* stream);
err_t b2g_bashFBatch_dev(void* d_blocks, size_t count, void* stream);

/* ======================================================================= belt (STB 34.101.31) */
/* drop-in: include/bee2/crypto/belt.h:148-257 (belt_block.c) */
const octet* beltH(void);
void beltKeyExpand(octet key_[32], const octet key[], size_t len);
void beltKeyExpand2(u32 key_[8], const octet key[], size_t len);
void beltBlockEncr(octet block[16], const u32 key[8]);
void beltBlockEncr2(u32 block[4], const u32 key[8]);
void beltBlockEncr3(u32* a, u32* b, u32* c, u32* d, const u32 key[8]);
void beltBlockDecr(octet block[16], const u32 key[8]);
void beltBlockDecr2(u32 block[4], const u32 key[8]);
void beltBlockDecr3(u32* a, u32* b, u32* c, u32* d, const u32 key[8]);
/* drop-in: belt.h:401-487 (belt_ecb.c:44-158); state layout = belt_ecb_st */
size_t beltECB_keep(void);
void beltECBStart(void* state, const octet key[], size_t len);
void beltECBStepE(void* buf, size_t count, void* state);
void beltECBStepD(void* buf, size_t count, void* state);
err_t beltECBEncr(void* dest, const void* src, size_t count, const octet key[], size_t len);
err_t beltECBDecr(void* dest, const void* src, size_t count, const octet key[], size_t len);
/* drop-in: belt.h:692-743 (belt_ctr.c:55-135); state layout = belt_ctr_st (belt_lcl.h:135-141) */
size_t beltCTR_keep(void);
void beltCTRStart(void* state, const octet key[], size_t len, const octet iv[16]);
void beltCTRStepE(void* buf, size_t count, void* state);
#define beltCTRStepD beltCTRStepE
err_t beltCTR(void* dest, const void* src, size_t count, const octet key[], size_t len,
	const octet iv[16]);
/* drop-in: belt.h (belt_hash.c:27-190): one-shot and streaming (hash-and-continue) forms; state layout =
   belt_hash_st without the compression stack */
err_t beltHash(octet hash[32], const void* src, size_t count);
size_t beltHash_keep(void);
void beltHashStart(void* state);
void beltHashStepH(const void* buf, size_t count, void* state);
void beltHashStepG(octet hash[32], void* state);
void beltHashStepG2(octet hash[], size_t hash_len, void* state);
bool_t beltHashStepV(const octet hash[32], void* state);
bool_t beltHashStepV2(const octet hash[], size_t hash_len, void* state);
/* drop-in: belt.h:984-1030 (belt_dwp.c:250-330) — authenticated encryption of (critical src1,
   open src2): dest = CTR(src1), mac = 8-octet tag; Unwrap returns ERR_BAD_MAC and leaves dest
   untouched when the tag differs. Whole buffers are staged on the device. */
err_t beltDWPWrap(void* dest, octet mac[8], const void* src1, size_t count1, const void* src2,
	size_t count2, const octet key[], size_t len, const octet iv[16]);
err_t beltDWPUnwrap(void* dest, const void* src1, size_t count1, const void* src2, size_t count2,
	const octet mac[8], const octet key[], size_t len, const octet iv[16]);
/* drop-in: belt.h (belt_che.c:257-330) — belt-CHE: same interface as DWP; the gamma runs over
   the LFSR counter s <- s*x ^ 1 and the tag uses r = E_K(iv) */
err_t beltCHEWrap(void* dest, octet mac[8], const void* src1, size_t count1, const void* src2,
	size_t count2, const octet key[], size_t len, const octet iv[16]);
err_t beltCHEUnwrap(void* dest, const void* src1, size_t count1, const void* src2, size_t count2,
	const octet mac[8], const octet key[], size_t len, const octet iv[16]);
/* drop-in: belt.h:850-983 (belt_dwp.c:27-207) and the belt-CHE twins (belt_che.c:27-239) — the
   streaming forms: StepE / StepD transform critical data in place, StepI absorbs open data, StepA
   absorbs critical data (ciphertext), StepG / StepV produce / check the tag without closing the
   state. Field order of the states follows belt_dwp_st / belt_che_st (the multiplication stack
   of the reference is not needed: the block chain runs on the device). */
size_t beltDWP_keep(void);
void beltDWPStart(void* state, const octet key[], size_t len, const octet iv[16]);
void beltDWPStepE(void* buf, size_t count, void* state);
void beltDWPStepI(const void* buf, size_t count, void* state);
void beltDWPStepA(const void* buf, size_t count, void* state);
void beltDWPStepD(void* buf, size_t count, void* state);
void beltDWPStepG(octet mac[8], void* state);
bool_t beltDWPStepV(const octet mac[8], void* state);
size_t beltCHE_keep(void);
void beltCHEStart(void* state, const octet key[], size_t len, const octet iv[16]);
void beltCHEStepE(void* buf, size_t count, void* state);
void beltCHEStepI(const void* buf, size_t count, void* state);
void beltCHEStepA(const void* buf, size_t count, void* state);
void beltCHEStepD(void* buf, size_t count, void* state);
void beltCHEStepG(octet mac[8], void* state);
bool_t beltCHEStepV(const octet mac[8], void* state);
/* batch: pure keystream (beltCTR of zeros) */
err_t beltCTRKeystream(void* dest, size_t count, const octet key[], size_t len,
	const octet iv[16]);
/* batch: key agility — block i (16 octets, in place) under 32-octet key i */
err_t beltECBEncrBatch(void* blocks, const octet* keys32, size_t count);
/* batch: belt-hash of `count` messages of msg_len octets (message i at msgs+i*stride) */
err_t beltHashBatch(octet* hashes, const void* msgs, size_t msg_len, size_t stride, size_t count);
/* device: dest[0..count) = src[0..count) XOR keystream blocks first_block.. of the stream
   whose encrypted iv is ctr0 = E_K(iv) (as produced by beltCTRStart); src may be NULL
   (keystream only) or equal to dest. key = expanded key (beltKeyExpand2). */
err_t b2g_beltCTR_dev(void* d_dest, const void* d_src, size_t count, const u32 key[8],
	const u32 ctr0[4], u64 first_block, void* stream);
/* device: d_mac[8] = belt-DWP tag over (d_open[n2] open, d_crit[n1] critical = ciphertext);
   d_scratch = 16 octets of device scratch (4-aligned) */
err_t b2g_beltDWPMac_dev(void* d_mac, const void* d_crit, size_t n1, const void* d_open, size_t n2,
	const u32 key[8], const u32 ctr0[4], void* d_scratch, void* stream);
/* device: belt-CHE data pass (gamma block j = E_K(s_(first_block+j+1)), s0 = E_K(iv)) and tag */
err_t b2g_beltCHE_dev(void* d_dest, const void* d_src, size_t count, const u32 key[8],
	const u32 s0[4], u64 first_block, void* stream);
err_t b2g_beltCHEMac_dev(void* d_mac, const void* d_crit, size_t n1, const void* d_open, size_t n2,
	const u32 key[8], const u32 s0[4], void* d_scratch, void* stream);
err_t b2g_beltECB_dev(void* d_dest, const void* d_src, size_t nblocks, const u32 key[8],
	int decrypt, void* stream);
err_t b2g_beltECBEncrBatch_dev(void* d_blocks, const void* d_keys32, size_t count, void* stream);
/* out of place: d_dst may be another device's HBM mapped with b2g_ipc_open (gather fused into the kernel) */
err_t b2g_beltECBEncrBatch2_dev(void* d_dst, const void* d_src, const void* d_keys32, size_t count, void* stream);
err_t b2g_beltHashBatch_dev(void* d_hashes, const void* d_msgs, size_t msg_len, size_t stride,
	size_t count, void* stream);

/* ======================================================================= bign (STB 34.101.45) */
/* include/bee2/crypto/bign.h:65-74 */
typedef struct
{
	size_t l;
	octet p[64];
	octet a[64];
	octet b[64];
	octet q[64];
	octet yG[64];
	octet seed[8];
} bign_params;

/* drop-in: bign.h (bign_params.c:197-236): "1.2.112.0.2.0.34.101.45.3.1" / ".3.2" / ".3.3"
   (bign-curve256v1 / 384v1 / 512v1, l = 128 / 192 / 256); other names -> ERR_FILE_NOT_FOUND */
err_t bignParamsStd(bign_params* params, const char* name);
/* drop-in: bign.h:370-402 (bign_sign.c:349-361, :247-260). With no = l/4 octets: hash no,
   sig no/2 + no, privkey no, pubkey 2 no. Parameter blocks other than the three standard
   curves return ERR_NOT_IMPLEMENTED (after the reference's structural checks,
   bign_params.c:244-280). These also serve bign128/192/256{Verify,Sign2,PubkeyCalc}
   (bign128.c:177-185, bign192.c, bign256.c), which only fix the level and the hash OID. */
err_t bignVerify(const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet sig[], const octet pubkey[]);
/* GPU-path limits of the bign entry points (they are staged as kernel arguments): the DER of the hash OID
   and bignSign2's optional `t` up to 64 octets each; the three standard parameter blocks. Beyond that the
   call goes to the stock libbee2 behind this library when there is one (overlay mode, see b2g_set_cpu_below)
   and returns ERR_NOT_IMPLEMENTED otherwise — the reference (bign_sign.c:198-206) takes any length. */
err_t bignSign2(octet sig[], const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet privkey[], const void* t, size_t t_len);
err_t bignPubkeyCalc(octet pubkey[], const bign_params* params, const octet privkey[]);
/* drop-in: bign.h (bign_misc.c:182-300, :317-352, :437-515). gen_i is the reference's generator
   callback (defs.h:520-524); the private key is drawn on the host exactly as the reference draws it
   (zzRandNZMod, no octets per attempt), the public key is computed on the device. */
typedef void (*gen_i)(void* buf, size_t count, void* state);
err_t bignKeypairGen(octet privkey[], octet pubkey[], const bign_params* params, gen_i rng, void* rng_state);
err_t bignKeypairVal(const bign_params* params, const octet privkey[], const octet pubkey[]);
err_t bignPubkeyVal(const bign_params* params, const octet pubkey[]);
err_t bignDH(octet key[], const bign_params* params, const octet privkey[], const octet pubkey[], size_t key_len);
/* drop-in: bign.h bignSign (bign_sign.c:27-138) — the randomised signature; the one-time key comes
   from the caller's generator (zzRandNZMod over q), everything else runs on the device */
err_t bignSign(octet sig[], const bign_params* params, const octet oid_der[], size_t oid_len,
	const octet hash[], const octet privkey[], gen_i rng, void* rng_state);
err_t bignSignBatch(err_t* status, octet* sigs, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* privkeys, gen_i rng, void* rng_state, size_t count);
/* drop-in: bign128.h / bign192.h / bign256.h (bign128.c:96-185, bign192.c, bign256.c): the fixed-level
   forms — standard curve of the level, OID of belt-hash / bash384 / bash512; with no = 32 / 48 / 64 octets:
   privkey, hash: no; pubkey: 2 no; sig: no/2 + no */
err_t bign128KeypairGen(octet privkey[], octet pubkey[], gen_i rng, void* rng_state);
err_t bign128KeypairVal(const octet privkey[], const octet pubkey[]);
err_t bign128PubkeyVal(const octet pubkey[]);
err_t bign128PubkeyCalc(octet pubkey[], const octet privkey[]);
err_t bign128DH(octet key[], const octet privkey[], const octet pubkey[], size_t key_len);
err_t bign128Sign(octet sig[], const octet hash[], const octet privkey[], gen_i rng, void* rng_state);
err_t bign128Sign2(octet sig[], const octet hash[], const octet privkey[], const void* t, size_t t_len);
err_t bign128Verify(const octet hash[], const octet sig[], const octet pubkey[]);
err_t bign192KeypairGen(octet privkey[], octet pubkey[], gen_i rng, void* rng_state);
err_t bign192KeypairVal(const octet privkey[], const octet pubkey[]);
err_t bign192PubkeyVal(const octet pubkey[]);
err_t bign192PubkeyCalc(octet pubkey[], const octet privkey[]);
err_t bign192DH(octet key[], const octet privkey[], const octet pubkey[], size_t key_len);
err_t bign192Sign(octet sig[], const octet hash[], const octet privkey[], gen_i rng, void* rng_state);
err_t bign192Sign2(octet sig[], const octet hash[], const octet privkey[], const void* t, size_t t_len);
err_t bign192Verify(const octet hash[], const octet sig[], const octet pubkey[]);
err_t bign256KeypairGen(octet privkey[], octet pubkey[], gen_i rng, void* rng_state);
err_t bign256KeypairVal(const octet privkey[], const octet pubkey[]);
err_t bign256PubkeyVal(const octet pubkey[]);
err_t bign256PubkeyCalc(octet pubkey[], const octet privkey[]);
err_t bign256DH(octet key[], const octet privkey[], const octet pubkey[], size_t key_len);
err_t bign256Sign(octet sig[], const octet hash[], const octet privkey[], gen_i rng, void* rng_state);
err_t bign256Sign2(octet sig[], const octet hash[], const octet privkey[], const void* t, size_t t_len);
err_t bign256Verify(const octet hash[], const octet sig[], const octet pubkey[]);
/* batch forms: keys are drawn in item order; status[i] is what the one-shot call would return */
err_t bignKeypairGenBatch(octet* privkeys, octet* pubkeys, const bign_params* params, gen_i rng,
	void* rng_state, size_t count);
err_t bignKeypairValBatch(err_t* status, const bign_params* params, const octet* privkeys,
	const octet* pubkeys, size_t count);
err_t bignPubkeyValBatch(err_t* status, const bign_params* params, const octet* pubkeys, size_t count);
err_t bignDHBatch(err_t* status, octet* keys, const bign_params* params, const octet* privkeys,
	const octet* pubkeys, size_t key_len, size_t count);
/* batch: item i uses hashes + no i, sigs + (no + no/2) i, pubkeys + 2 no i (l = 128: 32, 48, 64
   octets per item); status[i] = the err_t the
   reference's bignVerify would return for that item. Return value: ERR_OK when the batch
   ran (look at status[]), else the parameter/OID/device error. */
err_t bignVerifyBatch(err_t* status, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* sigs, const octet* pubkeys, size_t count);
err_t bignSign2Batch(err_t* status, octet* sigs, const bign_params* params, const octet oid_der[],
	size_t oid_len, const octet* hashes, const octet* privkeys, size_t count);
err_t bignPubkeyCalcBatch(err_t* status, octet* pubkeys, const bign_params* params,
	const octet* privkeys, size_t count);
/* batch ecMulA on bign-curve256v1 (ec.c:497-525): b_i = d_i * a_i; scalars are d_len octets
   each (LE, <= 32); ok[i] = 0 iff the result is the point at infinity. */
err_t ecMulABatch(octet* b, int* ok, const octet* a, const octet* d, size_t d_len, size_t count);
/* batch ecAddMulA with the base point (ec.c:1183-1273): b_i = d_i * a_i + k_i * G, k_i 32 octets */
err_t ecAddMulABatch(octet* b, int* ok, const octet* a, const octet* d, size_t d_len, const octet* k,
	size_t count);
/* drop-in: ec.h:892-901 (ec.c:497-525) for an `ec_o` built by the reference (bignEcCreate, ecpCreateJ over
   gfpCreate) that describes one of the three standard bign curves: only the data at the head of the
   descriptions are read (field modulus, A, B), the multiplication runs on the device; m <= n words.
   Other curves abort() — there is no CPU path. `ec` is the reference's `const ec_o*`, `word` = u64. */
bool_t ecMulA(u64 b[], const u64 a[], const void* ec, const u64 d[], size_t m, void* stack);
size_t ecMulA_deep(size_t n, size_t ec_d, size_t ec_deep, size_t m);
/* drop-in: ec.h:1176-1190 (ec.c:1183-1273): b <- sum of d_i a_i over the k triples
   (const word a_i[], const word d_i[], size_t m_i) that follow k; same recognition as ecMulA, k <= 64 */
bool_t ecAddMulA(u64 b[], const void* ec, void* stack, size_t k, ...);
/* the same on the standard curve of level l = 128 / 192 / 256: points l/2 octets, d_len <= l/4,
   k_i l/4 octets */
err_t ecMulABatchL(size_t l, octet* b, int* ok, const octet* a, const octet* d, size_t d_len, size_t count);
err_t ecAddMulABatchL(size_t l, octet* b, int* ok, const octet* a, const octet* d, size_t d_len,
	const octet* k, size_t count);
/* device */
err_t b2g_bignVerifyBatch_dev(void* d_status, const octet oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_sigs, const void* d_pubkeys, size_t count, void* stream);
err_t b2g_bignSign2Batch_dev(void* d_status, void* d_sigs, const octet oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, void* stream);
err_t b2g_bignPubkeyCalcBatch_dev(void* d_status, void* d_pubkeys, const void* d_privkeys,
	size_t count, void* stream);
err_t b2g_ecMulABatch_dev(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	size_t count, void* stream);
err_t b2g_ecAddMulABatch_dev(void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	const void* d_k, size_t count, void* stream);
/* device, any level (the names above are l = 128) */
err_t b2g_bignVerifyBatchL_dev(size_t l, void* d_status, const octet oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_sigs, const void* d_pubkeys, size_t count, void* stream);
err_t b2g_bignSign2BatchL_t_dev(size_t l, void* d_status, void* d_sigs, const octet oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, const void* t, size_t t_len, void* stream);
err_t b2g_bignPubkeyCalcBatchL_dev(size_t l, void* d_status, void* d_pubkeys, const void* d_privkeys,
	size_t count, void* stream);
err_t b2g_bignSignBatchL_k_dev(size_t l, void* d_status, void* d_sigs, const octet oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, const void* d_nonces, size_t count, void* stream);
err_t b2g_bignDHBatchL_dev(size_t l, void* d_status, void* d_out, const void* d_privkeys,
	const void* d_pubkeys, size_t count, void* stream);
err_t b2g_bignPubkeyValBatchL_dev(size_t l, void* d_status, const void* d_pubkeys, size_t count, void* stream);
err_t b2g_ecSumL_dev(size_t l, void* d_out, void* d_ok_out, const void* d_pts, const void* d_ok_in,
	size_t k, void* stream);
err_t b2g_ecMulABatchL_dev(size_t l, void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	size_t count, void* stream);
err_t b2g_ecAddMulABatchL_dev(size_t l, void* d_b, void* d_ok, const void* d_a, const void* d_d, size_t d_len,
	const void* d_k, size_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BEE2_B200_H */
