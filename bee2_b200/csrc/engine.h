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

#ifdef __cplusplus
}
#endif
#endif
