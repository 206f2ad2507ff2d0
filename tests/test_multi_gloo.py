"""N > 1 path on CPU: two gloo ranks shard one belt-CTR stream / one bash batch / one verify batch
with bee2_b200.shard, compute their shards with the ORACLE (test stand-in for the kernels), gather,
and must reproduce the single-rank bytes exactly (results identical for any G)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch
    import torch.distributed as dist
    import _oracle as o
    from bee2_b200 import shard

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- shared parameters come from rank 0 only
        secret = bytes(range(1, 49)) if rank == 0 else None
        kiv = shard.broadcast_bytes(secret, 48)
        key, iv = kiv[:32], kiv[32:]
        # ---- belt-CTR: one stream of 1000 blocks + 5 octets, split by counter offset
        total = 16 * 1000 + 5
        first, off, n = shard.ctr_shard(total, rank, world)
        st = o.port().orc_beltCTRStart
        state = (C.c_ubyte * 80)()
        st(state, key, C.c_size_t(32), iv)
        words = np.frombuffer(bytes(state)[32:48], dtype=np.uint32)
        adv = shard.ctr_add(words, first)                       # counter state after `first` blocks
        C.memmove(C.addressof(state) + 32, adv.ctypes.data, 16)
        buf = (C.c_ubyte * max(n, 1))()
        o.port().orc_beltCTRStepE(buf, C.c_size_t(n), state)
        mine = torch.from_numpy(np.frombuffer(bytes(buf)[:n], dtype=np.uint8).copy())
        whole = shard.gather_concat(mine)
        # ---- bash-512 batch of 37 messages x 200 octets
        rng = np.random.default_rng(0)
        msgs = rng.integers(0, 256, (37, 200), dtype=np.uint8)
        a, b_ = shard.shard_range(37, rank, world)
        dig = shard.gather_concat(torch.from_numpy(o.bashHashBatch(256, msgs[a:b_])))
        t = shard.max_over_ranks(1.0 + rank)
        if rank == 0:
            ok_ctr = whole.numpy().tobytes() == o.beltCTR(bytes(total), key, iv)
            ok_bash = np.array_equal(dig.numpy(), o.bashHashBatch(256, msgs))
            q.put((ok_ctr, ok_bash, t, len(whole)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_ranks_reproduce_single_rank_bytes(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ok_ctr, ok_bash, t, n = res
    assert ok_ctr and ok_bash and n == 16005
    assert t == float(world)          # max over ranks of (1 + rank)


def test_shard_arithmetic():
    from bee2_b200 import shard
    for total in (0, 1, 7, 1 << 26, (1 << 18) + 3):
        for world in (1, 2, 3, 8):
            rs = [shard.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    # ragged CTR stream: the tail goes to the last rank, offsets are whole blocks
    parts = [shard.ctr_shard(16 * 10 + 3, r, 4) for r in range(4)]
    assert sum(p[2] for p in parts) == 163 and all(p[1] == 16 * p[0] for p in parts)
    c = shard.ctr_add(np.array([0xFFFFFFFF, 0xFFFFFFFF, 0, 0], dtype=np.uint32), 1)
    assert list(c) == [0, 0, 1, 0]
    c = shard.ctr_add(np.array([0xFFFFFFFF] * 4, dtype=np.uint32), 2)
    assert list(c) == [1, 0, 0, 0]
