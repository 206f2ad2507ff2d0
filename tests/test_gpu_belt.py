"""GPU parity: belt block / ECB / CTR / hash through the C ABI vs the STB vectors, the reference
fixtures and the oracle (bit-exact). Mirrors test/crypto/belt_test.c:174-215, :288-339, :423-448, :593-631."""
import numpy as np
import pytest

import _oracle as o
import _vectors as v
import bee2_b200 as b

pytestmark = pytest.mark.gpu
KAT = v.load("kat.json")
REF = v.load("ref_vectors.json")
H = o.beltH()
R = lambda e: v.resolve(e, H, o.beltHash)  # noqa: E731


def test_beltH():
    assert b.beltH() == H


def test_block_A1_A4_all_flavours():
    for t in KAT["beltBlock"]:
        f = b.beltBlockEncr if t["op"] == "encr" else b.beltBlockDecr
        for flavour in (1, 2, 3):
            assert f(R(t["in"]), R(t["key"]), flavour).hex().upper() == t["out"]
    t = KAT["beltBlock"][0]
    assert b.beltBlockDecr(bytes.fromhex(t["out"]), R(t["key"])) == R(t["in"])
    # 16- and 24-octet keys go through beltKeyExpand2 (belt_block.c:88-106)
    for klen in (16, 24):
        assert b.beltBlockEncr(H[:16], H[128:128 + klen]) == o.beltBlockEncr(H[:16], H[128:128 + klen])


def test_zerosum():
    x = np.zeros((128, 4), dtype=np.uint32)
    x[:, 0] = KAT["beltZerosum"]["x"]
    st = b.BeltECB(bytes(32))
    buf = x.view(np.uint8).reshape(-1).copy()
    st.step_e(buf)
    acc = np.bitwise_xor.reduce(x ^ buf.view(np.uint32).reshape(128, 4), axis=0)
    assert not acc.any()


def test_ecb_A9_A10_split_calls_and_stealing():
    for t in KAT["beltECB"]:
        src, key = R(t["in"]), R(t["key"])
        one = (b.beltECBEncr if t["op"] == "encr" else b.beltECBDecr)(src, key)
        assert one.hex().upper() == t["out"], t["id"]
        st = b.BeltECB(key)
        buf = np.frombuffer(src, dtype=np.uint8).copy()
        pos = 0
        for c in t["split"]:
            (st.step_e if t["op"] == "encr" else st.step_d)(buf[pos:pos + c])
            pos += c
        assert buf.tobytes() == one
    with pytest.raises(b.Bee2Error) as e:
        b.beltECBEncr(H[:15], H[128:160])
    assert e.value.code == b.ERR_BAD_INPUT             # belt_ecb.c:117-118


def test_ctr_A15_A16_split_calls():
    for t in KAT["beltCTR"]:
        src, key, iv = R(t["in"]), R(t["key"]), R(t["iv"])
        assert b.beltCTR(src, key, iv).hex().upper() == t["out"], t["id"]
        st = b.BeltCTR(key, iv)
        buf = np.frombuffer(src, dtype=np.uint8).copy()
        pos = 0
        for c in t["split"]:
            st.step_e(buf[pos:pos + c])
            pos += c
        assert buf.tobytes().hex().upper() == t["out"]
    with pytest.raises(b.Bee2Error) as e:
        b.beltCTR(H[:16], H[128:145], H[192:208])
    assert e.value.code == b.ERR_BAD_INPUT


@pytest.mark.parametrize("mode", ["DWP", "CHE"])
def test_dwp_che_A19_A20_and_reference_fixtures(mode):
    wrap, unwrap = getattr(b, f"belt{mode}Wrap"), getattr(b, f"belt{mode}Unwrap")
    for t in KAT[f"belt{mode}"]:
        src, op, key, iv = R(t["in"]), R(t["open"]), R(t["key"]), R(t["iv"])
        if t["op"] == "wrap":
            out, mac = wrap(src, op, key, iv)
            assert out.hex().upper() == t["out"] and mac.hex().upper() == t["mac"]
        else:
            code, out = unwrap(src, op, bytes.fromhex(t["mac"]), key, iv)
            assert code == 0 and out.hex().upper() == t["out"]
            bad = bytearray(bytes.fromhex(t["mac"]))
            bad[3] ^= 1
            assert unwrap(src, op, bad, key, iv) == (b.ERR_BAD_MAC, None)     # belt_dwp.c:318-322
    for t in REF[f"belt{mode}"]:
        a = [bytes.fromhex(t[k]) for k in ("in", "open", "key", "iv")]
        out, mac = wrap(*a)
        assert out.hex() == t["out"] and mac.hex() == t["mac"]
        assert unwrap(out, a[1], mac, a[2], a[3]) == (0, a[0])


@pytest.mark.parametrize("mode", ["DWP", "CHE"])
def test_dwp_che_random_sizes_vs_oracle(mode):
    """Sizes that exercise one thread, many threads, many CTAs and ragged tails of both data classes."""
    rng = np.random.default_rng(31)
    wrap, unwrap = getattr(b, f"belt{mode}Wrap"), getattr(b, f"belt{mode}Unwrap")
    owrap = getattr(o, f"belt{mode}Wrap")
    for n1, n2 in [(0, 0), (5, 0), (0, 5), (511, 513), (16 * 40, 16 * 3), (100_001, 33), (7, 200_003), (3_000_017, 1_000_003)]:
        key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
        iv = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        a = rng.integers(0, 256, n1, dtype=np.uint8).tobytes()
        op = rng.integers(0, 256, n2, dtype=np.uint8).tobytes()
        want = owrap(a, op, key, iv)
        assert wrap(a, op, key, iv) == want, (n1, n2)
        assert unwrap(want[0], op, want[1], key, iv) == (0, a)
        if n2:
            assert unwrap(want[0], op[:-1] + bytes([op[-1] ^ 1]), want[1], key, iv)[0] == b.ERR_BAD_MAC


@pytest.mark.parametrize("mode", ["DWP", "CHE"])
def test_dwp_che_streaming(mode):
    """The drop-in streaming forms (belt_test.c:474-560): A.19 incremental sequences, then random
    programs of split StepI / StepE / StepA / StepG calls against the oracle's sequential chain."""
    key, iv, steps, want_buf, want_mac = v.aead_incremental_program(mode, H)
    st = b.BeltDWP(key, iv, mode)
    buf, tags = v.run_aead_program(st, steps)
    assert buf.hex().upper() == want_buf and tags[-1].hex().upper() == want_mac
    assert st.step_v(tags[-1]) and not st.step_v(tags[0])
    assert (buf, tags[-1]) == getattr(b, f"belt{mode}Wrap")(H[:len(buf)], H[16:48], key, iv)
    rng = np.random.default_rng(41)
    rb = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()  # noqa: E731
    for trial in range(12):
        key, iv = rb(int(rng.choice([16, 24, 32]))), rb(16)
        g, w = b.BeltDWP(key, iv, mode), o.BeltDWP(key, iv, mode)
        for _ in range(int(rng.integers(0, 4))):
            d = rb(int(rng.choice([0, 1, 5, 16, 17, 40, 100, 5000])))
            g.step_i(d), w.step_i(d)
            if rng.random() < 0.3:
                assert g.step_g() == w.step_g()
        for _ in range(int(rng.integers(1, 5))):
            d = rb(int(rng.choice([0, 1, 7, 16, 33, 64, 129, 70001])))
            c = g.step_e(d)
            assert c == w.step_e(d)
            g.step_a(c), w.step_a(c)
            if rng.random() < 0.4:
                assert g.step_g() == w.step_g()
        mac = w.step_g()
        assert g.step_g() == mac and g.step_v(mac)
        assert not g.step_v(bytes([mac[0] ^ 1]) + mac[1:])


@pytest.mark.parametrize("mode", ["DWP", "CHE"])
def test_dwp_che_multi_chunk_pipeline(mode):
    """More than two 32 MiB pipeline stages with a ragged tail: the chunks are encrypted on alternating
    streams from their own counter offsets and authenticated by one tag launch over the whole buffer."""
    rng = np.random.default_rng(77)
    n1, n2 = (70 << 20) + 13, 4097
    key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    iv = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    a = rng.integers(0, 256, n1, dtype=np.uint8).tobytes()
    op = rng.integers(0, 256, n2, dtype=np.uint8).tobytes()
    want = getattr(o, f"belt{mode}Wrap")(a, op, key, iv)
    got = getattr(b, f"belt{mode}Wrap")(a, op, key, iv)
    assert got[1] == want[1] and got[0] == want[0]
    assert getattr(b, f"belt{mode}Unwrap")(want[0], op, want[1], key, iv) == (0, a)


def test_che_sharded_by_block_offset():
    """belt-CHE gamma from any block offset equals the slice of the whole stream (multi-GPU sharding)."""
    torch = pytest.importorskip("torch")
    key = b.beltKeyExpand2(H[128:160])
    s0 = np.frombuffer(b.beltBlockEncr(H[192:208], H[128:160]), dtype=np.uint32).copy()
    n = 1 << 24
    s = torch.cuda.current_stream().cuda_stream
    whole = torch.zeros(n, dtype=torch.uint8, device="cuda")
    b.beltCHE_dev(whole.data_ptr(), 0, n, key, s0, 0, s)
    torch.cuda.synchronize()
    assert whole[: 1 << 16].cpu().numpy().tobytes() == o.beltCHEWrap(bytes(1 << 16), b"", H[128:160], H[192:208])[0]
    for off_blocks, m in ((1, 4096), (12345, 1 << 20), ((n // 16) - 100, 1600)):
        part = torch.zeros(m, dtype=torch.uint8, device="cuda")
        b.beltCHE_dev(part.data_ptr(), 0, m, key, s0, off_blocks, s)
        torch.cuda.synchronize()
        assert torch.equal(part, whole[16 * off_blocks:16 * off_blocks + m])


def test_hash_A23():
    for t in KAT["beltHash"]:
        assert b.beltHash(R(t["in"])).hex().upper() == t["out"]


def test_reference_fixtures():
    for t in REF["beltCTR"]:
        assert b.beltCTR(bytes.fromhex(t["in"]), bytes.fromhex(t["key"]), bytes.fromhex(t["iv"])).hex() == t["out"]
    for t in REF["beltECB"]:
        assert b.beltECBEncr(bytes.fromhex(t["in"]), bytes.fromhex(t["key"])).hex() == t["out"]
        assert b.beltECBDecr(bytes.fromhex(t["out"]), bytes.fromhex(t["key"])).hex() == t["in"]
    for t in REF["beltHash"]:
        assert b.beltHash(bytes.fromhex(t["in"])).hex() == t["out"]


def test_ctr_random_lengths_and_streaming():
    rng = np.random.default_rng(11)
    for n in [1, 15, 16, 17, 255, 4096, 100_003, (1 << 20) + 5]:
        key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
        iv = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        src = rng.integers(0, 256, n, dtype=np.uint8)
        want = o.beltCTR(src.tobytes(), key, iv)
        assert b.beltCTR(src.tobytes(), key, iv) == want
        assert b.beltCTRKeystream(n, key, iv) == o.beltCTR(bytes(n), key, iv)
        st, buf, pos = b.BeltCTR(key, iv), src.copy(), 0
        while pos < n:                      # ragged steps exercise the keystream reserve
            c = min(int(rng.integers(1, max(2, n // 3))), n - pos)
            st.step_e(buf[pos:pos + c])
            pos += c
        assert buf.tobytes() == want
    # counter carry across 2^32 and 2^64 (belt_ctr.c:27-35)
    key, n = bytes(range(32)), 16 * 40
    ks = b.BeltCTR(key, bytes(16))
    want = o.beltCTR(bytes(n), key, bytes(16))
    out = np.zeros(n, dtype=np.uint8)
    ks.step_e(out)
    assert out.tobytes() == want


def test_ctr_counter_wrap_device_level():
    torch = pytest.importorskip("torch")
    key = np.arange(8, dtype=np.uint32)
    for ctr in ([0xFFFFFFF0, 0, 0, 0], [0xFFFFFFF8, 0xFFFFFFFF, 7, 0], [0xFFFFFFFE, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF]):
        c = np.array(ctr, dtype=np.uint32)
        n = 16 * 64
        out = torch.zeros(n, dtype=torch.uint8, device="cuda")
        b.beltCTR_dev(out.data_ptr(), 0, n, key, c, 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        # oracle: encrypt the incremented counters one by one
        val = int.from_bytes(c.tobytes(), "little")
        want = b"".join(o.beltBlockEncr(((val + j + 1) % (1 << 128)).to_bytes(16, "little"), key.tobytes())
                        for j in range(64))
        assert out.cpu().numpy().tobytes() == want


def test_ecb_multikey_batch_config5_shape():
    rng = np.random.default_rng(3)
    cnt = 20_011
    blocks = rng.integers(0, 256, (cnt, 16), dtype=np.uint8)
    keys = rng.integers(0, 256, (cnt, 32), dtype=np.uint8)
    assert np.array_equal(b.beltECBEncrBatch(blocks, keys), o.beltECBEncrMultiKey(blocks, keys))


def test_hash_streaming():
    """belt_test.c:593-621 (A.23-1/2/3 with StepG, StepV, StepV2, StepG2 between StepH calls), then random
    splits: the digest taken after any prefix equals the one-shot hash of that prefix and hashing goes on."""
    st = b.BeltHash()
    st.step_h(H[:13])
    assert st.step_g().hex().upper() == "ABEF9725D4C5A83597A367D14494CC2542F20F659DDFECC961A3EC550CBA8C75"
    st = b.BeltHash()
    st.step_h(H[:32])
    d = bytes.fromhex("749E4C3653AECE5E48DB4761227742EB6DBE13F4A80F7BEFF1A9CF8D10EE7786")
    assert st.step_v(d) and st.step_v(d[:13]) and not st.step_v(d[:12] + b"\0")
    st = b.BeltHash()
    st.step_h(H[:11])
    assert st.step_g(32) == o.beltHash(H[:11])
    st.step_h(H[11:48])
    assert st.step_v(bytes.fromhex("9D02EE446FB6A29FE5C982D4B13AF9D3E90861BC4CEF27CF306BFB0B174A154A"))
    rng = np.random.default_rng(61)
    for trial in range(6):
        data = rng.integers(0, 256, int(rng.choice([0, 31, 32, 33, 1000, 100_003])), dtype=np.uint8).tobytes()
        st, pos = b.BeltHash(), 0
        while True:
            assert st.step_g() == o.beltHash(data[:pos])
            if pos == len(data):
                break
            n = min(len(data) - pos, int(rng.choice([0, 1, 5, 31, 32, 64, 97, 4096, 50_000])))
            st.step_h(data[pos:pos + n])
            pos += n
        assert st.step_g(7) == o.beltHash(data)[:7]


def test_hash_batch_random():
    rng = np.random.default_rng(4)
    for n in (0, 1, 31, 32, 33, 75, 96, 1001):
        msgs = rng.integers(0, 256, (50, n), dtype=np.uint8)
        got = b.beltHashBatch(msgs)
        for i in range(50):
            assert got[i].tobytes() == o.beltHash(msgs[i].tobytes())


def test_hash_batch_large_persistent_kernel():
    """Batches of at least half a wave run the persistent shape (bank-replicated tables, grid-stride over the
    messages): ragged count, aligned and unaligned message lengths / strides, sampled against the oracle and — all
    of them — against the small-batch kernel (the same messages hashed in slices of 1000)."""
    rng = np.random.default_rng(44)
    for n, mlen in ((160_001, 96), (80_000, 203)):
        msgs = rng.integers(0, 256, (n, mlen), dtype=np.uint8)
        got = b.beltHashBatch(msgs)
        for i in list(range(0, n, n // 40)) + [n - 1]:
            assert got[i].tobytes() == o.beltHash(msgs[i].tobytes())
        for lo in range(0, n, 20_000):
            assert np.array_equal(b.beltHashBatch(msgs[lo:lo + 1000]), got[lo:lo + 1000])


def test_config2_shape_device_level():
    """BASELINE config 2 shape on device-resident buffers (256 MiB here; bench.py runs the full
    1 GiB): keystream equals the oracle on the head, offsets are consistent, and E(E(x)) = x."""
    torch = pytest.importorskip("torch")
    n = 1 << 28
    st = b.BeltCTR(H[128:160], H[192:208])
    key, ctr = st.key_words, st.ctr_words
    s = torch.cuda.current_stream().cuda_stream
    ks = torch.empty(n, dtype=torch.uint8, device="cuda")
    b.beltCTR_dev(ks.data_ptr(), 0, n, key, ctr, 0, s)
    head = 1 << 22
    torch.cuda.synchronize()
    assert ks[:head].cpu().numpy().tobytes() == o.beltCTR(bytes(head), H[128:160], H[192:208])
    # any shard computed with its block offset equals the same slice of the whole stream
    off_blocks, m = (n // 16) // 3, 1 << 20
    part = torch.empty(m, dtype=torch.uint8, device="cuda")
    b.beltCTR_dev(part.data_ptr(), 0, m, key, ctr, off_blocks, s)
    torch.cuda.synchronize()
    assert torch.equal(part, ks[16 * off_blocks:16 * off_blocks + m])
    # involution on data, in place
    g = torch.Generator(device="cuda").manual_seed(9)
    data = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    ref = data.clone()
    b.beltCTR_dev(data.data_ptr(), data.data_ptr(), n, key, ctr, 0, s)
    torch.cuda.synchronize()
    assert torch.equal(data, ref ^ ks)
    b.beltCTR_dev(data.data_ptr(), data.data_ptr(), n, key, ctr, 0, s)
    torch.cuda.synchronize()
    assert torch.equal(data, ref)
