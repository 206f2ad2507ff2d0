"""Multi-GPU parity on hardware (results byte-identical for G in {1, 2, 4, 8}).

Two ways to reach several GPUs are covered, both through the C ABI:
  * the in-process multi-device mode (`b2g_init_devices`): the host-pointer *Batch calls shard their
    units over the devices, one host thread per device — compared with the one-device result and with
    the oracle / the unmodified reference;
  * one process per GPU (what bench.py does under torchrun): a 2-rank NCCL job in which each rank
    computes its shard and stores it straight into rank 0's HBM through a CUDA-IPC mapping.
Also: the engine on a device != 0 driven from a thread that did not initialise it (the CUDA current
device is per host thread). Skipped when the box has one GPU."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import _oracle as o
import bee2_b200 as b

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = o.beltH()


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs at least two GPUs")


@pytest.fixture
def all_devices():
    assert b.b2g_init(0) == 0
    assert b.b2g_init_devices(0) == 0
    n = b.b2g_device_count()
    yield n
    assert b.b2g_init_devices(1) == 0
    assert b.b2g_device_count() == 1


@needs2
def test_batch_calls_shard_over_all_devices_byte_identical(all_devices):
    n_dev = all_devices
    assert n_dev == _ngpu()
    rng = np.random.default_rng(11)
    # bash-512: 2^15 messages x 4 KiB (128 MiB: every device gets a share)
    msgs = rng.integers(0, 256, (1 << 15, 4096), dtype=np.uint8)
    multi = b.bashHashBatch(256, msgs)
    # belt-CTR keystream, 256 MiB + a ragged tail
    nks = (256 << 20) + 5
    ks_multi = b.beltCTRKeystream(nks, H[128:160], H[192:208])
    # belt-ECB key agility, 2^21 (key, block) pairs
    blocks = rng.integers(0, 256, (1 << 21, 16), dtype=np.uint8)
    keys = rng.integers(0, 256, (1 << 21, 32), dtype=np.uint8)
    ecb_multi = b.beltECBEncrBatch(blocks, keys)
    # bign: keys, signatures, verification with 1/7 corrupted
    cnt = 1 << 15
    params = b.bignParamsStd()
    priv = rng.integers(0, 256, (cnt, 32), dtype=np.uint8)
    priv[:, 31] &= 0x7F
    hashes = rng.integers(0, 256, (cnt, 32), dtype=np.uint8)
    st_p, pub_multi = b.bignPubkeyCalcBatch(params, priv)
    st_s, sig_multi = b.bignSign2Batch(params, b.OID_BELT_HASH_DER, hashes, priv)
    bad = sig_multi.copy()
    bad[::7, 3] ^= 0x10
    st_multi = b.bignVerifyBatch(params, b.OID_BELT_HASH_DER, hashes, bad, pub_multi)
    assert not st_p.any() and not st_s.any()

    # the same calls on ONE device
    assert b.b2g_init_devices(1) == 0
    assert np.array_equal(multi, b.bashHashBatch(256, msgs))
    assert ks_multi == b.beltCTRKeystream(nks, H[128:160], H[192:208])
    assert np.array_equal(ecb_multi, b.beltECBEncrBatch(blocks, keys))
    st1, pub1 = b.bignPubkeyCalcBatch(params, priv)
    st2, sig1 = b.bignSign2Batch(params, b.OID_BELT_HASH_DER, hashes, priv)
    assert np.array_equal(pub_multi, pub1) and np.array_equal(sig_multi, sig1)
    assert np.array_equal(st_multi, b.bignVerifyBatch(params, b.OID_BELT_HASH_DER, hashes, bad, pub1))
    assert (st_multi[::7] == b.ERR_BAD_SIG).all() and int((st_multi == 0).sum()) == cnt - len(st_multi[::7])

    # and against the checkers on samples that hit every device's share
    idx = np.linspace(0, msgs.shape[0] - 1, 4 * n_dev + 1).astype(int)
    assert np.array_equal(multi[idx], o.bashHashBatch(256, msgs[idx]))
    head = 1 << 16
    assert ks_multi[:head] == o.beltCTR(bytes(head), H[128:160], H[192:208])
    R = o.ref()
    if R is not None:
        import ctypes as C
        want = np.zeros(nks, dtype=np.uint8)
        src = np.zeros(nks, dtype=np.uint8)
        assert R.beltCTR(want.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), C.c_size_t(nks),
                         H[128:160], C.c_size_t(32), H[192:208]) == 0
        assert ks_multi == want.tobytes()
    idx = np.linspace(0, blocks.shape[0] - 1, 4 * n_dev + 1).astype(int)
    assert np.array_equal(ecb_multi[idx], o.beltECBEncrMultiKey(blocks[idx], keys[idx]))
    for i in np.linspace(0, cnt - 1, 2 * n_dev + 1).astype(int):
        assert o.bignPubkeyCalc(priv[i].tobytes()) == (0, pub_multi[i].tobytes())
        assert o.bignSign2(hashes[i].tobytes(), priv[i].tobytes()) == (0, sig_multi[i].tobytes())
        assert o.bignVerify(hashes[i].tobytes(), bad[i].tobytes(), pub_multi[i].tobytes()) == st_multi[i]


_OTHER_DEVICE = textwrap.dedent("""
    import sys, threading
    import numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    import bee2_b200 as b, _oracle as o
    assert b.b2g_init(1) == 0
    out = {{}}
    def work():
        rng = np.random.default_rng(3)
        msgs = rng.integers(0, 256, (512, 1000), dtype=np.uint8)
        out["bash"] = np.array_equal(b.bashHashBatch(256, msgs), o.bashHashBatch(256, msgs))
        H = o.beltH()
        out["ctr"] = b.beltCTRKeystream(1 << 16, H[128:160], H[192:208]) == o.beltCTR(bytes(1 << 16), H[128:160], H[192:208])
        params = b.bignParamsStd()
        priv = rng.integers(0, 256, (64, 32), dtype=np.uint8); priv[:, 31] &= 0x7F
        hashes = rng.integers(0, 256, (64, 32), dtype=np.uint8)
        st, pub = b.bignPubkeyCalcBatch(params, priv)
        st2, sig = b.bignSign2Batch(params, b.OID_BELT_HASH_DER, hashes, priv)
        st3 = b.bignVerifyBatch(params, b.OID_BELT_HASH_DER, hashes, sig, pub)
        out["bign"] = (not st.any() and not st2.any() and not st3.any()
                       and o.bignSign2(hashes[5].tobytes(), priv[5].tobytes()) == (0, sig[5].tobytes()))
    ts = [threading.Thread(target=work) for _ in range(2)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert out == {{"bash": True, "ctr": True, "bign": True}}, out
    print("OK")
""")


@needs2
def test_engine_on_device_1_from_threads_that_did_not_initialise_it():
    r = subprocess.run([sys.executable, "-c", _OTHER_DEVICE.format(root=ROOT)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


_TWO_RANKS = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    import bee2_b200 as b, _oracle as o
    from bee2_b200 import shard
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    assert b.b2g_init(rank) == 0
    L = b.lib(); stream = torch.cuda.current_stream().cuda_stream
    H = o.beltH()
    # one NCCL broadcast of key || iv, then every rank's kernel writes its shard into rank 0's HBM
    kiv = shard.broadcast_bytes(H[128:160] + H[192:208] if rank == 0 else None, 48, device=dev)
    st = b.BeltCTR(kiv[:32], kiv[32:])
    total = 1 << 22                                  # blocks
    lo, hi = shard.shard_range(total, rank, world)
    gptr = L.b2g_dev_alloc(total * 16) if rank == 0 else None
    handle = shard.broadcast_bytes(b.b2g_ipc_export(gptr) if rank == 0 else None, 64, device=dev)
    peer = gptr if rank == 0 else b.b2g_ipc_open(handle)
    b.beltCTR_dev(peer + 16 * lo, 0, 16 * (hi - lo), st.key_words, st.ctr_words, lo, stream)
    # bash-512 digests the same way
    g = torch.Generator(device=dev).manual_seed(1)
    msgs = torch.randint(0, 256, (1 << 12, 4096), dtype=torch.uint8, device=dev, generator=g)
    lo2, hi2 = shard.shard_range(1 << 12, rank, world)
    hptr = L.b2g_dev_alloc((1 << 12) * 64) if rank == 0 else None
    handle2 = shard.broadcast_bytes(b.b2g_ipc_export(hptr) if rank == 0 else None, 64, device=dev)
    peer2 = hptr if rank == 0 else b.b2g_ipc_open(handle2)
    b.bashHashBatch_dev(peer2 + 64 * lo2, 256, msgs.data_ptr() + 4096 * lo2, 4096, 4096, hi2 - lo2, stream)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        ks = np.zeros(total * 16, dtype=np.uint8)
        assert L.b2g_memcpy_d2h(ks.ctypes.data, gptr, ks.size) == 0
        assert ks.tobytes() == b.beltCTRKeystream(total * 16, kiv[:32], kiv[32:])          # one-GPU run
        assert ks[: 1 << 16].tobytes() == o.beltCTR(bytes(1 << 16), kiv[:32], kiv[32:])      # oracle
        assert ks[-(1 << 12):].tobytes() == o.beltCTR(bytes(total * 16), kiv[:32], kiv[32:])[-(1 << 12):]
        dg = np.zeros(((1 << 12), 64), dtype=np.uint8)
        assert L.b2g_memcpy_d2h(dg.ctypes.data, hptr, dg.size) == 0
        hm = msgs.cpu().numpy()
        assert np.array_equal(dg, b.bashHashBatch(256, hm))
        idx = [0, 1, 2047, 2048, 4095]
        assert np.array_equal(dg[idx], o.bashHashBatch(256, hm[idx]))
        print("OK")
    dist.barrier()
    if rank != 0:
        b.b2g_ipc_close(peer); b.b2g_ipc_close(peer2)
    dist.barrier()
    dist.destroy_process_group()
""")


@needs2
def test_two_ranks_fused_gather_into_rank0_over_cuda_ipc(tmp_path):
    script = tmp_path / "two_ranks.py"
    script.write_text(_TWO_RANKS.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29653", str(script)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
