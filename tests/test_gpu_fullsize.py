"""GPU parity at BASELINE.json's FULL sizes through size-independent properties (the oracle only
sees samples): configs 2 (belt-CTR 1 GiB), 3 (bash-512, 2^20 x 4 KiB), 4 (bign verify, 2^18) and
5 (belt-ECB, 2^26 keys/blocks). Device-resident buffers, device-level C-ABI entry points."""
import numpy as np
import pytest

import _oracle as o
import bee2_b200 as b

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
H = o.beltH()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def test_config2_belt_ctr_1GiB():
    n = 1 << 30
    st = b.BeltCTR(H[128:160], H[192:208])
    key, ctr = st.key_words, st.ctr_words
    ks = torch.empty(n, dtype=torch.uint8, device="cuda")
    b.beltCTR_dev(ks.data_ptr(), 0, n, key, ctr, 0, _stream())
    torch.cuda.synchronize()
    # the head of the stream against the port AND the unmodified reference
    head = 1 << 20
    assert ks[:head].cpu().numpy().tobytes() == o.beltCTR(bytes(head), H[128:160], H[192:208])
    if o.ref() is not None:
        assert ks[:head].cpu().numpy().tobytes() == o.ref_beltCTR(bytes(head), H[128:160], H[192:208])
    # windows DEEP inside the stream (middle, last MiB, across the 2^32-block... the low counter word carries
    # at block 2^32 - ctr0, far beyond 1 GiB, so the carry case is covered by test_gpu_belt's counter-wrap test):
    # keystream block j is E_K(s + j + 1) with s = E_K(iv) as a 128-bit LE integer (belt_ctr.c:27-35, :55-111),
    # computed here with the checkers' ECB on explicitly built counter blocks
    s0 = int.from_bytes(o.beltECBEncr(H[192:208], H[128:160]), "little")
    for first in ((n >> 5) + 12345, (n >> 4) - (1 << 12)):            # block indices: mid-stream, the last 64 KiB
        cnt_blocks = 1 << 12
        ctrs = b"".join(((s0 + first + j + 1) % (1 << 128)).to_bytes(16, "little") for j in range(cnt_blocks))
        want = o.beltECBEncr(ctrs, H[128:160])
        assert ks[16 * first:16 * (first + cnt_blocks)].cpu().numpy().tobytes() == want
        if o.ref() is not None:
            assert o.ref_beltECBEncr(ctrs, H[128:160]) == want
    # 8-way sharding by block offset reproduces the same stream (what N ranks compute)
    part = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
    for r in (0, 3, 7):
        b.beltCTR_dev(part.data_ptr(), 0, n // 8, key, ctr, r * (n // 128), _stream())
        torch.cuda.synchronize()
        assert torch.equal(part, ks[r * (n // 8):(r + 1) * (n // 8)])
    del part
    # keystream blocks are pairwise distinct on a large sample (a permutation of distinct counters)
    sample = ks[: 1 << 24].view(torch.int64).view(-1, 2)
    assert torch.unique(sample, dim=0).shape[0] == sample.shape[0]
    # involution on data
    g = torch.Generator(device="cuda").manual_seed(5)
    data = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    chk = int(data.view(torch.int64).sum().item())
    b.beltCTR_dev(data.data_ptr(), data.data_ptr(), n, key, ctr, 0, _stream())
    torch.cuda.synchronize()
    assert int(data.view(torch.int64).sum().item()) != chk
    b.beltCTR_dev(data.data_ptr(), data.data_ptr(), n, key, ctr, 0, _stream())
    torch.cuda.synchronize()
    assert int(data.view(torch.int64).sum().item()) == chk


def test_config3_bash512_2pow20_messages():
    cnt, n = 1 << 20, 4096
    g = torch.Generator(device="cuda").manual_seed(1)
    msgs = torch.randint(0, 256, (cnt, n), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.empty((cnt, 64), dtype=torch.uint8, device="cuda")
    b.bashHashBatch_dev(out.data_ptr(), 256, msgs.data_ptr(), n, n, cnt, _stream())
    torch.cuda.synchronize()
    idx = np.random.default_rng(0).integers(0, cnt, 96)
    host = out.cpu().numpy()
    for i in list(idx) + [0, cnt - 1]:
        m = msgs[int(i)].cpu().numpy().tobytes()
        assert host[i].tobytes() == o.bashHash(256, m)
        if o.ref() is not None and i % 3 == 0:
            assert host[i].tobytes() == o.ref_bashHash(256, m)
    # duplicates hash alike, single-bit changes do not: copy message 0 over message 1, flip a bit in 2
    msgs[1] = msgs[0]
    msgs[2, 4095] ^= 1
    out2 = torch.empty((8, 64), dtype=torch.uint8, device="cuda")
    b.bashHashBatch_dev(out2.data_ptr(), 256, msgs.data_ptr(), n, n, 8, _stream())
    torch.cuda.synchronize()
    assert torch.equal(out2[1], out2[0]) and torch.equal(out2[0], out[0]) and not torch.equal(out2[2], out[2])
    assert torch.equal(out2[3:], out[3:8])
    # all 2^20 digests distinct
    assert torch.unique(out.view(torch.int64), dim=0).shape[0] == cnt - 0


def test_config4_bign_verify_2pow18():
    n = 1 << 18
    rng = np.random.default_rng(2)
    p = b.bignParamsStd()
    priv = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    priv[:, 31] &= 0x7F
    hashes = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    assert not st.any()
    st, sigs = b.bignSign2Batch(p, b.OID_BELT_HASH_DER, hashes, priv)
    assert not st.any()
    # sign -> verify round trip over the whole batch, then 1/16 corrupted in four different ways
    assert not b.bignVerifyBatch(p, b.OID_BELT_HASH_DER, hashes, sigs, pubs).any()
    want = np.zeros(n, dtype=np.uint32)
    k = np.arange(0, n, 16)
    sigs[k[0::4], 3] ^= 0x10
    want[k[0::4]] = 510                                # s0 changed
    sigs[k[1::4], 16:] = 0xFF
    want[k[1::4]] = 510                                # s1 >= q
    pubs[k[2::4], :32] = 0xFF
    want[k[2::4]] = 505                                # Qx >= p
    hashes[k[3::4], 7] ^= 1
    want[k[3::4]] = 510                                # other message
    got = b.bignVerifyBatch(p, b.OID_BELT_HASH_DER, hashes, sigs, pubs)
    assert np.array_equal(got, want)
    for i in rng.integers(0, n, 48):
        assert got[i] == o.bignVerify(hashes[i].tobytes(), sigs[i].tobytes(), pubs[i].tobytes())
    # 2^15 items of the batch (every 8th: half of them corrupted, all four kinds) against the UNMODIFIED
    # reference's bign128Verify, fanned out over the host threads by oracle/cpu_harness.c
    if o.ref() is not None:
        import ctypes as C
        import os
        sel = np.arange(0, n, 8)
        hs, ss, ps = (np.ascontiguousarray(x[sel]) for x in (hashes, sigs, pubs))
        harness = C.CDLL(os.path.join(o.REF_DIR, "libcpuharness.so"))
        harness.harness_bign_verify.restype = C.c_double
        st_ref = np.zeros(len(sel), dtype=np.uint32)
        dt = harness.harness_bign_verify(os.path.join(o.REF_DIR, "libbee2ref_64.so").encode(), 0,
                                         st_ref.ctypes.data_as(C.c_void_p), hs.ctypes.data_as(C.c_void_p),
                                         ss.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p),
                                         C.c_size_t(len(sel)), len(os.sched_getaffinity(0)))
        assert dt > 0 and np.array_equal(st_ref, got[sel])
        assert {0, 505, 510} <= set(int(x) for x in st_ref)


def test_config5_belt_ecb_2pow26_keys():
    cnt = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(3)
    keys = torch.randint(0, 256, (cnt, 32), dtype=torch.uint8, device="cuda", generator=g)
    blocks = torch.randint(0, 256, (cnt, 16), dtype=torch.uint8, device="cuda", generator=g)
    idx = np.concatenate([np.random.default_rng(1).integers(0, cnt, 500), [0, cnt - 1]])
    tidx = torch.from_numpy(idx).cuda()
    pk, pb = keys[tidx].cpu().numpy(), blocks[tidx].cpu().numpy()
    b.beltECBEncrBatch_dev(blocks.data_ptr(), keys.data_ptr(), cnt, _stream())
    torch.cuda.synchronize()
    enc = blocks[tidx].cpu().numpy()
    assert np.array_equal(enc, o.beltECBEncrMultiKey(pb, pk))
    if o.ref() is not None:
        for j in range(0, len(idx), 25):
            assert enc[j].tobytes() == o.ref_beltECBEncr(pb[j].tobytes(), pk[j].tobytes())
    # same key + same block -> same ciphertext, wherever it sits in the batch
    keys[5] = keys[cnt - 7]
    blocks[5] = 7
    blocks[cnt - 7] = 7
    b.beltECBEncrBatch_dev(blocks.data_ptr(), keys.data_ptr(), cnt, _stream())
    torch.cuda.synchronize()
    assert torch.equal(blocks[5], blocks[cnt - 7])


def test_bign_verify_beyond_one_launch_chunk():
    """More items than one verification launch takes (bign.cu BIGN_WTAB_CHUNK = 2^20: the per-thread window
    tables live in a stream-ordered scratch area sized per launch): 2^20 + 777 items are verified in two
    launches; the batch is 2^12 distinct (hash, signature, key) triples, one in nine corrupted and each checked
    against the oracle or the unmodified reference, tiled over the whole count."""
    base, n = 1 << 12, (1 << 20) + 777
    rng = np.random.default_rng(20)
    p = b.bignParamsStd()
    priv = rng.integers(0, 256, (base, 32), dtype=np.uint8)
    priv[:, 31] &= 0x7F
    hashes = rng.integers(0, 256, (base, 32), dtype=np.uint8)
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    st2, sigs = b.bignSign2Batch(p, b.OID_BELT_HASH_DER, hashes, priv)
    assert not st.any() and not st2.any()
    sigs[::9, 5] ^= 0x20
    want = b.bignVerifyBatch(p, b.OID_BELT_HASH_DER, hashes, sigs, pubs)
    for i in range(0, base, 64):
        assert want[i] == o.bignVerify(hashes[i].tobytes(), sigs[i].tobytes(), pubs[i].tobytes())
    assert (want[::9] == 510).all() and int((want == 0).sum()) == base - len(range(0, base, 9))
    reps = (n + base - 1) // base
    big = [np.ascontiguousarray(np.tile(x, (reps, 1))[:n]) for x in (hashes, sigs, pubs)]
    got = b.bignVerifyBatch(p, b.OID_BELT_HASH_DER, *big)
    assert np.array_equal(got, np.tile(want, reps)[:n])
