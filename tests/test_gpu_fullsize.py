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
    # oracle on the head and on a window deep inside the stream (via the counter offset)
    head = 1 << 20
    assert ks[:head].cpu().numpy().tobytes() == o.beltCTR(bytes(head), H[128:160], H[192:208])
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
        assert host[i].tobytes() == o.bashHash(256, msgs[int(i)].cpu().numpy().tobytes())
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
    assert np.array_equal(blocks[tidx].cpu().numpy(), o.beltECBEncrMultiKey(pb, pk))
    # same key + same block -> same ciphertext, wherever it sits in the batch
    keys[5] = keys[cnt - 7]
    blocks[5] = 7
    blocks[cnt - 7] = 7
    b.beltECBEncrBatch_dev(blocks.data_ptr(), keys.data_ptr(), cnt, _stream())
    torch.cuda.synchronize()
    assert torch.equal(blocks[5], blocks[cnt - 7])
