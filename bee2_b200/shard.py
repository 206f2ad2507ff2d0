"""Multi-GPU plumbing for the batch engine: one process per GPU, units sharded by contiguous
index ranges, no data-path collective. The only collectives are a broadcast of the
small shared parameters (belt key + iv, bign OID) from rank 0 and an optional final gather of the
outputs; both go through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

Nothing here computes: the functions only decide WHICH units a rank owns and move bytes.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, last) of `total` units owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError((rank, world))
    return total * rank // world, total * (rank + 1) // world


def ctr_add(ctr_words: np.ndarray, n: int) -> np.ndarray:
    """ctr + n as a 128-bit little-endian integer in four u32 words: the counter state a
    belt-CTR stream has after n blocks (n applications of beltBlockIncU32, belt_ctr.c:27-35)."""
    v = (int.from_bytes(np.asarray(ctr_words, dtype=np.uint32).tobytes(), "little") + n) % (1 << 128)
    return np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint32).copy()


def ctr_shard(total_bytes: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(first_block, byte_offset, nbytes) of this rank's slice of one belt-CTR stream of
    total_bytes octets: whole 16-octet blocks, the ragged tail goes to the last rank."""
    nblocks = (total_bytes + 15) // 16
    b0, b1 = shard_range(nblocks, rank, world)
    off = 16 * b0
    end = min(16 * b1, total_bytes)
    return b0, off, max(0, end - off)


def broadcast_bytes(data: Optional[bytes], nbytes: int, device=None, src: int = 0) -> bytes:
    """Rank `src` sends `data` (nbytes octets) to every rank; a no-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert data is not None and len(data) == nbytes
        return bytes(data)
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t = torch.from_numpy(np.frombuffer(bytes(data), dtype=np.uint8).copy()).to(t.device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy().tobytes()


def gather_concat(local, dst: int = 0):
    """Final gather of per-rank outputs (torch tensors, possibly of different lengths along dim 0)
    to rank `dst`, concatenated in rank order. Other ranks get None."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: List = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def max_over_ranks(seconds: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
