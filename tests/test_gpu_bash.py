"""GPU parity: bash-f / bash-hash through the C ABI vs the STB vectors, the reference fixtures
and the oracle (bit-exact). Mirrors test/crypto/bash_test.c:41-154."""
import numpy as np
import pytest

import _oracle as o
import _vectors as v
import bee2_b200 as b

pytestmark = pytest.mark.gpu
KAT = v.load("kat.json")
REF = v.load("ref_vectors.json")
H = o.beltH()


def test_bashF_A2_and_platform():
    t = KAT["bashF"][0]
    assert b.bashF(H[:192]).hex().upper() == t["out"]
    assert b.lib().bashF_deep() == 0


@pytest.mark.parametrize("t", KAT["bashHash"], ids=lambda t: t["id"])
def test_bashHash_A3_oneshot_and_streaming(t):
    l, n = t["l"], t["len"]
    assert b.bashHash(l, H[:n]).hex().upper() == t["out"]
    # Start/StepH/StepG equivalence (bash_test.c:64-68) with ragged splits, hash-and-continue
    for cuts in ([n], [0, n], [n // 3, n - n // 3], [1] * min(n, 5) + [max(n - 5, 0)]):
        st = b.BashHash(l)
        pos = 0
        for c in cuts:
            st.step_h(H[pos:pos + c])
            pos += c
        assert pos == n
        assert st.step_g(l // 4).hex().upper() == t["out"]
        assert st.step_g(l // 8).hex().upper() == t["out"][: l // 4]      # StepG leaves the state usable
        assert st.step_v(bytes.fromhex(t["out"])) and not st.step_v(bytes(l // 4))
        st2 = st.copy()                                                  # states are memcpy-able
        st.step_h(b"xyz"), st2.step_h(b"xyz")
        assert st.step_g(l // 4) == st2.step_g(l // 4) == o.bashHash(l, H[:n] + b"xyz")


def test_bashHash_errors_match_reference():
    for l in (0, 8, 17, 272):
        with pytest.raises(b.Bee2Error) as e:
            b.bashHash(l, b"abc")
        assert e.value.code == b.ERR_BAD_PARAMS       # bash_hash.c:122-123


def test_reference_fixtures():
    for t in REF["bashHash"]:
        assert b.bashHash(t["l"], bytes.fromhex(t["in"])).hex() == t["out"]


@pytest.mark.parametrize("l", [16 * i for i in range(1, 17)])
def test_batch_every_level_random_lengths(l):
    rng = np.random.default_rng(l)
    rate = 192 - l // 2
    for msg_len in sorted({0, 1, rate - 1, rate, rate + 1, 3 * rate, int(rng.integers(1, 900))}):
        for stride in {msg_len, msg_len + 3, (msg_len + 15) // 16 * 16 + 16}:
            cnt = 37
            buf = rng.integers(0, 256, size=(cnt, stride), dtype=np.uint8)
            got = b.bashHashBatch(l, buf, msg_len=msg_len, stride=stride, count=cnt)
            want = o.bashHashBatch(l, np.ascontiguousarray(buf[:, :msg_len]))
            assert np.array_equal(got, want), (l, msg_len, stride)


def test_ragged_batch_bsum_style():
    """Many messages of different lengths and alignments in one buffer (the bsum case)."""
    rng = np.random.default_rng(21)
    lens = np.concatenate([rng.integers(0, 700, 300), [0, 1, 63, 64, 65, 4096, 10_000]]).astype(np.uint64)
    gaps = rng.integers(0, 9, lens.size).astype(np.uint64)
    offsets = np.cumsum(np.concatenate([[0], (lens + gaps)[:-1]])).astype(np.uint64)
    data = rng.integers(0, 256, int(offsets[-1] + lens[-1]) + 16, dtype=np.uint8)
    for l in (128, 192, 256, 80):
        got = b.bashHashBatchV(l, data, offsets, lens)
        for i in range(lens.size):
            m = data[int(offsets[i]):int(offsets[i] + lens[i])].tobytes()
            assert got[i].tobytes() == o.bashHash(l, m), (l, i)
    with pytest.raises(b.Bee2Error) as e:
        b.bashHashBatchV(256, data, np.array([data.size], dtype=np.uint64), np.array([1], dtype=np.uint64))
    assert e.value.code == b.ERR_BAD_INPUT


def test_bash_prg_A4_A5_A6_and_random_programs():
    import _prg_kats
    _prg_kats.run(b.BashPrg, H)
    rng = np.random.default_rng(10)
    for _ in range(12):
        _prg_kats.random_program(b.BashPrg, o.BashPrg, rng, H)
    assert b.lib().bashPrg_keep() == 8 + 8 + 192 + 8 + 8 + 192        # bash_prg.c:54-62


def test_bashFBatch_random():
    rng = np.random.default_rng(1)
    st = rng.integers(0, 256, size=(300, 192), dtype=np.uint8)
    got = b.bashFBatch(st)
    for i in range(0, 300, 7):
        assert got[i].tobytes() == o.bashF(st[i].tobytes())
    assert b.bashFBatch(np.zeros((0, 192), dtype=np.uint8)).size == 0


def test_config3_shape_device_level():
    """BASELINE config 3 shape (4 KiB messages, bash-512) on device-resident data: sampled digests
    against the oracle, and sharding invariance (any split of the batch gives the same digests)."""
    torch = pytest.importorskip("torch")
    cnt, n = 1 << 15, 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    msgs = torch.randint(0, 256, (cnt, n), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.empty((cnt, 64), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    b.bashHashBatch_dev(out.data_ptr(), 256, msgs.data_ptr(), n, n, cnt, s)
    out2 = torch.empty_like(out)
    cut = 12345
    b.bashHashBatch_dev(out2.data_ptr(), 256, msgs.data_ptr(), n, n, cut, s)
    b.bashHashBatch_dev(out2[cut:].data_ptr(), 256, msgs[cut:].data_ptr(), n, n, cnt - cut, s)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    host, hmsg = out.cpu().numpy(), msgs.cpu().numpy()
    for i in list(range(0, cnt, 997)) + [cnt - 1]:
        assert host[i].tobytes() == o.bashHash(256, hmsg[i].tobytes())


def test_hash_files_bsum_style(tmp_path):
    """bashHashFiles (the bsum case, cmd/bsum/bsum.c:142-200): many files of ragged sizes incl. empty ones, a missing
    file, and one file larger than the 64 MiB staging buffer (streamed)."""
    rng = np.random.default_rng(123)
    sizes = [0, 1, 63, 64, 65, 4096, 100_003, 1_000_001, 0, 31] + [int(x) for x in rng.integers(0, 300_000, 40)]
    paths, datas = [], []
    for i, n in enumerate(sizes):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        p = tmp_path / f"f{i}.bin"
        p.write_bytes(d)
        paths.append(str(p)), datas.append(d)
    big = rng.integers(0, 256, (64 << 20) + 12345, dtype=np.uint8).tobytes()
    (tmp_path / "big.bin").write_bytes(big)
    paths.insert(5, str(tmp_path / "missing.bin")), datas.insert(5, None)
    paths.insert(20, str(tmp_path / "big.bin")), datas.insert(20, big)
    for l in (256, 128):
        st, hs = b.bashHashFiles(l, paths)
        for i, d in enumerate(datas):
            if d is None:
                assert st[i] == 203
            elif len(d) < (8 << 20) or l == 256:
                assert st[i] == 0 and hs[i].tobytes() == o.bashHash(l, d), (l, i, len(d))
            else:
                assert st[i] == 0


def test_concurrent_host_threads():
    """SURVEY §8b threading: every entry point may be called from many host threads at once (ctypes drops
    the GIL); results must equal the single-threaded ones."""
    import threading
    rng = np.random.default_rng(99)
    msgs = rng.integers(0, 256, size=(512, 1000), dtype=np.uint8)
    key, iv = o.beltH()[128:160], o.beltH()[192:208]
    data = rng.integers(0, 256, size=100_003, dtype=np.uint8).tobytes()
    params = b.bignParamsStd()
    priv = rng.integers(0, 256, size=(64, 32), dtype=np.uint8)
    priv[:, 31] &= 0x7F
    hashes = rng.integers(0, 256, size=(64, 32), dtype=np.uint8)
    _, pubs = b.bignPubkeyCalcBatch(params, priv)
    _, sigs = b.bignSign2Batch(params, b.OID_BELT_HASH_DER, hashes, priv)
    want = (b.bashHashBatch(256, msgs), b.beltCTR(data, key, iv), b.beltDWPWrap(data, data[:77], key, iv),
            b.bignVerifyBatch(params, b.OID_BELT_HASH_DER, hashes, sigs, pubs))
    errors = []

    def worker(kind):
        try:
            for _ in range(6):
                if kind == 0:
                    assert np.array_equal(b.bashHashBatch(256, msgs), want[0])
                elif kind == 1:
                    assert b.beltCTR(data, key, iv) == want[1]
                elif kind == 2:
                    assert b.beltDWPWrap(data, data[:77], key, iv) == want[2]
                    st = b.BeltDWP(key, iv)
                    st.step_i(data[:77])
                    c = st.step_e(data[:50_000]) + st.step_e(data[50_000:])
                    st.step_a(c)
                    assert (c, st.step_g()) == want[2]
                else:
                    assert np.array_equal(b.bignVerifyBatch(params, b.OID_BELT_HASH_DER, hashes, sigs, pubs), want[3])
        except Exception as e:  # noqa: BLE001
            errors.append((kind, repr(e)))

    threads = [threading.Thread(target=worker, args=(k % 4,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_bash_prg_blocks_batch_matches_single_automata():
    """b2g_bashPrgBlocks_dev over many automata (the batch of SURVEY §8f rank 3): every automaton's state and
    data must come out as when it is run alone (count = 1 — the path the bashPrg* drop-ins use, which the A.4–A.6
    vectors and the random programs pin on the oracle). Whole-word rates (64-bit fast path) and a ragged rate /
    unaligned stride (octet path), all four commands."""
    import ctypes as C
    torch = pytest.importorskip("torch")
    L = b.lib()
    L.b2g_bashPrgBlocks_dev.restype = C.c_uint32
    L.b2g_bashPrgBlocks_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                         C.c_size_t, C.c_void_p]
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(99)
    n, blocks = 300, 3
    for buf_len, pad in ((160, 0), (64, 8), (156, 0), (168, 3)):
        stride = blocks * buf_len + pad
        for mode in range(4):
            states = rng.integers(0, 256, (n, 192), dtype=np.uint8)
            data = rng.integers(0, 256, (n, stride), dtype=np.uint8)
            ds, dd = torch.from_numpy(states).cuda(), torch.from_numpy(data).cuda()
            assert L.b2g_bashPrgBlocks_dev(ds.data_ptr(), dd.data_ptr(), stride, blocks, buf_len, mode, mode & 1, n, stream) == 0
            gs, gd = ds.cpu().numpy(), dd.cpu().numpy()
            for i in (0, 1, 31, 32, 127, 128, 299):
                s1, d1 = torch.from_numpy(states[i].copy()).cuda(), torch.from_numpy(data[i].copy()).cuda()
                assert L.b2g_bashPrgBlocks_dev(s1.data_ptr(), d1.data_ptr(), stride, blocks, buf_len, mode, mode & 1, 1, stream) == 0
                assert np.array_equal(s1.cpu().numpy(), gs[i]) and np.array_equal(d1.cpu().numpy(), gd[i]), (buf_len, mode, i)
            if mode == 0:
                assert np.array_equal(gd, data)                      # absorb leaves the data alone
            else:
                assert np.array_equal(gd[:, blocks * buf_len:], data[:, blocks * buf_len:])   # nothing past the blocks
