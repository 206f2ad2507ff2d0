// bign_lowocc.cu — the LATENCY-BOUND build of the bign verification kernel (l = 128).
//
// Same source as bign.cu (it is included below), compiled with every field product inlined
// (FE_MUL_INLINE) and one CTA per SM (255 registers, no spills): for grids of under ~2 warps per
// scheduler — a shard of 2^18 / 8 signatures on one B200 is 147 CTAs, one per SM — the kernel is bound
// by the latency of each warp's dependent carry chains, and the inlined body lets the scheduler
// interleave the independent products of a point formula. Measured (B200, 2^15 items): 0.813 ms
// against 0.978 ms for the out-of-line build; at full grids the out-of-line build wins (4.84 vs
// 5.73 ms for 2^18), so bign.cu's launcher picks by grid size (verify_launch).
#define FE_MUL_INLINE 1
#define BIGN_MIN_BLOCKS 1
#define BIGN_LOWOCC_TU 1
#include "bign.cu"
