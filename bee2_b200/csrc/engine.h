/* engine.h — internal host-side plumbing shared by host_*.c (C, not exported as API). */
#ifndef B2G_ENGINE_H
#define B2G_ENGINE_H

#include <cuda_runtime_api.h>
#include "../../include/bee2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* selects this thread's device (see engine.c), makes it current, lazily brings the engine up on it
   (constant tables, streams). 0 or an err_t. Every entry point starts with it. */
u32 b2g_ensure_device(void);
void b2g_note_launch(void);
u32 b2g_check_launch(const char* what);
u32 b2g_cuda_fail(cudaError_t e, const char* what);   /* records text, returns ERR_B2G_CUDA (or OOM) */
void b2g_die(const char* fn, u32 code);              /* void drop-ins: print + abort() */

/* The lock of the device selected by the last b2g_ensure_device() on this thread: host-pointer
   entry points on one device share its workspace. */
void b2g_lock(void);
void b2g_unlock(void);
#define B2G_MAX_DEV 16
int b2g_cur_dev(void);   /* CUDA ordinal selected by the last b2g_ensure_device() on this thread */
/* Shard `count` units over the device set (b2g_init_devices): fn(arg, first, n) runs once per
   device on its own host thread with that device selected; shares are contiguous, multiples of 256
   units, none smaller than `grain`. One device (or a small batch): a plain call on this thread. */
typedef u32 (*b2g_shard_fn)(void* arg, size_t first, size_t n);
u32 b2g_fanout(size_t count, size_t grain, b2g_shard_fn fn, void* arg);

/* Pipelined workspace: NSLOT slots, each a stream + growable device buffers. */
#define B2G_NSLOT 2
#define B2G_NBUF 4
typedef struct
{
	cudaStream_t stream;
	void* buf[B2G_NBUF];
	size_t cap[B2G_NBUF];
} b2g_slot;
b2g_slot* b2g_slot_get(int i);
/* device buffer `which` of slot `s` with at least `bytes` capacity (grown if needed) */
u32 b2g_slot_buf(b2g_slot* s, int which, size_t bytes, void** out);

/* waits for the slot's stream, then zeroes every device buffer of the slot (failure paths of the
   entry points that stage secrets) */
void b2g_slot_wipe(b2g_slot* s);

/* kernels' device-level launchers that are not part of the public header */
u32 b2g_belt_upload_tables(const octet H[256]);
u32 b2g_bash_upload_tables(void);
u32 b2g_beltdwp_upload_tables(const octet H[256]);
u32 b2g_bign_upload_tables(const octet H[256]);
u32 b2g_bignSign2Batch_t_dev(void* d_status, void* d_sigs, const octet oid_der[], size_t oid_len,
	const void* d_hashes, const void* d_privkeys, size_t count, const void* t, size_t t_len, void* stream);
u32 b2g_bashSponge_dev(const void* d_msgs, size_t stride, size_t msg_len, size_t count,
	size_t l, const void* d_states_in, void* d_states_out, void* d_hashes, size_t hash_len,
	int pre_f, void* stream);

u32 b2g_beltPolyAbsorb_dev(void* d_t, const void* d_blocks, size_t nbytes, const u32 r[4], const u32 t0[4],
	void* d_scratch, void* stream);

u32 b2g_beltHashStep_dev(void* d_state, const void* d_data, size_t nblocks, int final, const u32 len[4], void* stream);

/* units per pipeline chunk so that one chunk moves about `target_bytes` */
size_t b2g_chunk_units(size_t unit_bytes, size_t target_bytes);

/* ---- overlay mode (SURVEY.md §8b "Packaging consequence") ----------------
   When a stock libbee2 sits BEHIND this library in the process's symbol search order (link order
   `-lbee2_b200 -lbee2`, or LD_PRELOAD=libbee2_b200.so), the drop-in entry points forward to it —
   through dlsym(RTLD_NEXT, name), never to oracle/ — in three cases:
     1. inputs the GPU path does not cover (non-standard bign_params, a generic ec_o, m > n words,
        an OID or `t` longer than 64 octets);
     2. payloads below the routing threshold (b2g_set_cpu_below / B2G_CPU_BELOW, default 0 = off):
        a 13-byte bashHash or a single beltBlockEncr is latency-bound on any GPU;
     3. a `void` drop-in whose GPU path failed (instead of abort()).
   Without a stock library behind, 1 returns ERR_NOT_IMPLEMENTED and 3 aborts, as before. */
/* constant-time comparison of digests / MACs (the reference's memEq is regular too, mem.c) */
int b2g_ct_eq(const void* a, const void* b, size_t n);
/* zero a host buffer that held key material (not optimised away) */
void b2g_wipe(void* p, size_t n);
void* b2g_stock(const char* name);
int b2g_route_small(size_t bytes);
void b2g_note_forward(const char* fn);   /* counted; B2G_TRACE_FORWARD=1 prints each name on stderr */
/* a call handed to stock libbee2 because the GPU path could NOT run it (not by routing policy):
   counted, and said once per process on stderr — the engine never degrades silently */
void b2g_warn_forward(const char* fn, u32 code);
#define B2G_STOCK_FN(name) __extension__({ static void* f_; static int tried_; \
	if (!tried_) { f_ = b2g_stock(#name); __sync_synchronize(); tried_ = 1; } (__typeof__(&name))f_; })
/* forward a small call: for functions returning a value / void */
#define B2G_SMALL_R(bytes, name, ...) do { if (b2g_route_small(bytes)) { __typeof__(&name) fs_ = B2G_STOCK_FN(name); \
	if (fs_) { b2g_note_forward(#name); return fs_(__VA_ARGS__); } } } while (0)
#define B2G_SMALL_V(bytes, name, ...) do { if (b2g_route_small(bytes)) { __typeof__(&name) fs_ = B2G_STOCK_FN(name); \
	if (fs_) { b2g_note_forward(#name); fs_(__VA_ARGS__); return; } } } while (0)
/* forward unconditionally if a stock library exists (unsupported input) */
#define B2G_STOCK_R(name, ...) do { __typeof__(&name) fs_ = B2G_STOCK_FN(name); \
	if (fs_) { b2g_note_forward(#name); return fs_(__VA_ARGS__); } } while (0)
/* the GPU path of a void drop-in failed: stock if there is one, else abort */
#define B2G_FAIL_V(code, name, ...) do { __typeof__(&name) fs_ = B2G_STOCK_FN(name); \
	if (fs_) { b2g_warn_forward(#name, code); fs_(__VA_ARGS__); return; } b2g_die(#name, code); } while (0)
/* first line of a void drop-in: no usable device -> the whole call goes to stock libbee2, untouched */
#define B2G_PREFLIGHT_V(name, ...) do { const u32 pc_ = b2g_ensure_device(); if (pc_) B2G_FAIL_V(pc_, name, __VA_ARGS__); } while (0)
#define B2G_PREFLIGHT_R(name, ...) do { const u32 pc_ = b2g_ensure_device(); if (pc_) B2G_FAIL_R(pc_, name, __VA_ARGS__); } while (0)
#define B2G_FAIL_R(code, name, ...) do { __typeof__(&name) fs_ = B2G_STOCK_FN(name); \
	if (fs_) { b2g_note_forward(#name); return fs_(__VA_ARGS__); } b2g_die(#name, code); } while (0)

#ifdef __cplusplus
}
#endif
#endif
