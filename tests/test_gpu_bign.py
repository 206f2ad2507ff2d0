"""GPU parity: bign on bign-curve256v1 through the C ABI vs the STB vectors, the reference fixtures
and the oracle (bit-exact statuses / signatures / points). Mirrors test/crypto/bign_test.c:303-457 and
the differential style of test/math/ec_test.c:342-470."""
import ctypes as C
import os

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
OID = b.OID_BELT_HASH_DER
Q = v.Q
P = (1 << 256) - 189


def A(x):
    return np.frombuffer(bytes(x), dtype=np.uint8).copy()


def test_params_std_and_checks():
    p = b.bignParamsStd()
    assert p.l == 128 and bytes(p.p)[:32] == P.to_bytes(32, "little") and bytes(p.q)[:32] == Q.to_bytes(32, "little")
    with pytest.raises(b.Bee2Error) as e:
        b.bignParamsStd("1.2.3")
    assert e.value.code == b.ERR_FILE_NOT_FOUND
    z = bytes(32)
    bad = b.bignParamsStd()
    bad.l = 100
    assert b.bignVerify(bad, OID, z, bytes(48), bytes(64)) == b.ERR_NOT_IMPLEMENTED     # bign_params.c:251-252
    bad = b.bignParamsStd()
    bad.p[0] = 0x41
    assert b.bignVerify(bad, OID, z, bytes(48), bytes(64)) == b.ERR_BAD_PARAMS          # p mod 4 != 3
    other = b.bignParamsStd()
    other.b[0] ^= 2                                                                     # a valid-looking other curve
    assert b.bignVerify(other, OID, z, bytes(48), bytes(64)) == b.ERR_NOT_IMPLEMENTED
    p = b.bignParamsStd()
    assert b.bignVerify(p, b"\x06\x02\x80\x01", z, bytes(48), bytes(64)) == b.ERR_BAD_OID
    assert b.bignVerify(p, b"\x05\x00", z, bytes(48), bytes(64)) == b.ERR_BAD_OID


def test_G1_G2_G3_and_negatives():
    p, k = b.bignParamsStd(), KAT["bign"]
    priv, pub = bytes.fromhex(k["privkey"]), bytes.fromhex(k["pubkey"])
    assert b.bignPubkeyCalc(p, priv) == pub                                    # G.1 (bign_test.c:320-327)
    for t in k["verify"]:
        h, sig = R(t["hash"]), bytearray(bytes.fromhex(t["sig"]))
        assert b.bignVerify(p, OID, h, sig, pub) == b.ERR_OK
        sig[0] ^= 1
        assert b.bignVerify(p, OID, h, sig, pub) == b.ERR_BAD_SIG              # bign_test.c:349-351
        sig[0] ^= 1
        bad = bytearray(pub)
        bad[0] ^= 1
        assert b.bignVerify(p, OID, h, sig, bad) == o.bignVerify(h, sig, bad) != b.ERR_OK


def test_G6_G7_sign2():
    p, k = b.bignParamsStd(), KAT["bign"]
    priv, pub = bytes.fromhex(k["privkey"]), bytes.fromhex(k["pubkey"])
    for t in k["sign2_nonce"]:
        h = R(t["hash"])
        sig = b.bignSign2(p, OID, h, priv, R(t["t"]))
        assert v.sign2_nonce(sig, priv, h).hex().upper() == t["k"], t["id"]
        assert (0, sig) == o.bignSign2(h, priv, R(t["t"]))
        assert b.bignVerify(p, OID, h, sig, pub) == b.ERR_OK
    with pytest.raises(b.Bee2Error) as e:
        b.bignSign2(p, OID, bytes(32), bytes(32))                              # d = 0
    assert e.value.code == b.ERR_BAD_PRIVKEY
    with pytest.raises(b.Bee2Error) as e:
        b.bignSign2(p, OID, bytes(32), Q.to_bytes(32, "little"))               # d = q
    assert e.value.code == b.ERR_BAD_PRIVKEY


def test_reference_fixtures():
    p = b.bignParamsStd()
    for t in REF["bign"]:
        priv, pub, h, sig = (bytes.fromhex(t[k]) for k in ("privkey", "pubkey", "hash", "sig"))
        tt = bytes.fromhex(t["t"]) if t["t"] else None
        assert b.bignPubkeyCalc(p, priv) == pub
        assert b.bignSign2(p, OID, h, priv, tt) == sig
        assert b.bignVerify(p, OID, h, sig, pub) == t["verify"]
        bad = t["bad"]
        assert b.bignVerify(p, OID, bytes.fromhex(bad["hash"]), bytes.fromhex(bad["sig"]),
                            bytes.fromhex(bad["pubkey"])) == bad["verify"]


# ---------------------------------------------------------------- levels 192 / 256
CURVE_Q = {192: 0xfffffffffffffffffffffffffffffffffffffffffffffffe6cccc40373af7bbb8046dae7a6a4ff0a3db7dc3ff30ca7b7,
           256: 0xffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffb2c0092c0198004ef26bebb02e2113f4361bcae59556df32dcffad490d068ef1}


@pytest.mark.parametrize("l", [192, 256])
def test_levels_params_and_fixtures(l):
    """bign-curve384v1 / 512v1 against outputs of the unmodified reference (tests/golden/ref_vectors.json
    "bignL"): key calculation, deterministic signatures, verification verdicts incl. corrupted inputs."""
    p = b.bignParamsStd(b.BIGN_CURVES[l])
    no, oid = l // 4, o.OIDS[l]
    assert p.l == l and bytes(p.q)[:no] == CURVE_Q[l].to_bytes(no, "little") and bytes(p.q)[no:] == bytes(64 - no)
    if o.ref() is not None:
        rp = o.ref_params(l)
        for f in ("p", "a", "b", "q", "yG", "seed"):
            assert bytes(getattr(p, f)) == bytes(getattr(rp, f)), f
    other = b.bignParamsStd(b.BIGN_CURVES[l])
    other.b[0] ^= 2
    assert b.bignVerify(other, oid, bytes(no), bytes(no + no // 2), bytes(2 * no)) == b.ERR_NOT_IMPLEMENTED
    for t in [t for t in REF["bignL"] if t["l"] == l]:
        priv, pub, h, sig = (bytes.fromhex(t[k]) for k in ("privkey", "pubkey", "hash", "sig"))
        tt = bytes.fromhex(t["t"]) if t["t"] else None
        assert b.bignPubkeyCalc(p, priv) == pub
        assert b.bignSign2(p, oid, h, priv, tt) == sig
        assert b.bignVerify(p, oid, h, sig, pub) == t["verify"]
        bad = t["bad"]
        assert b.bignVerify(p, oid, bytes.fromhex(bad["hash"]), bytes.fromhex(bad["sig"]),
                            bytes.fromhex(bad["pubkey"])) == bad["verify"]
    with pytest.raises(b.Bee2Error) as e:
        b.bignSign2(p, oid, bytes(no), CURVE_Q[l].to_bytes(no, "little"))       # d = q
    assert e.value.code == b.ERR_BAD_PRIVKEY


@pytest.mark.parametrize("l", [192, 256])
def test_levels_batch_random_vs_oracle(l):
    rng = np.random.default_rng(100 + l)
    p, n = b.bignParamsStd(b.BIGN_CURVES[l]), 200
    no, oid, q = l // 4, o.OIDS[l], CURVE_Q[l]
    priv = rng.integers(0, 256, (n, no), dtype=np.uint8)
    priv[:, no - 1] &= 0x7F
    for i, d in enumerate([1, 2, q - 1, q - 2, 1 << l, (1 << l) + 1]):          # exceptional keys: Q = +-G, ...
        priv[i] = A(d.to_bytes(no, "little"))
    hashes = rng.integers(0, 256, (n, no), dtype=np.uint8)
    hashes[::5, no // 2:] = 0xFF                                               # H >= q path
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    assert (st == 0).all()
    st, sigs = b.bignSign2Batch(p, oid, hashes, priv)
    assert (st == 0).all()
    for i in list(range(6)) + list(range(6, n, 25)):
        assert o.bignPubkeyCalc(priv[i].tobytes(), l) == (0, pubs[i].tobytes())
        assert o.bignSign2(hashes[i].tobytes(), priv[i].tobytes(), None, oid, l) == (0, sigs[i].tobytes())
    assert (b.bignVerifyBatch(p, oid, hashes, sigs, pubs) == 0).all()
    for i in range(n):
        kind = int(rng.integers(0, 18))
        if kind == 0:
            sigs[i, int(rng.integers(0, no // 2))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            sigs[i, no // 2 + int(rng.integers(0, no))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:
            hashes[i, int(rng.integers(0, no))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 3:
            pubs[i, int(rng.integers(0, 2 * no))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 4:
            sigs[i, no // 2:] = 0xFF
        elif kind == 5:
            pubs[i, (0 if i & 1 else no):][:no] = 0xFF
    got = b.bignVerifyBatch(p, oid, hashes, sigs, pubs)
    want = np.array([o.bignVerify(hashes[i].tobytes(), sigs[i].tobytes(), pubs[i].tobytes(), oid, l) for i in range(n)],
                    dtype=np.uint32)
    assert np.array_equal(got, want)
    assert {0, 505, 510} <= set(int(x) for x in got)


@pytest.mark.parametrize("l", [192, 256])
def test_levels_ecMulA_vs_oracle(l):
    rng = np.random.default_rng(200 + l)
    p, n = b.bignParamsStd(b.BIGN_CURVES[l]), 24
    no, q = l // 4, CURVE_Q[l]
    priv = rng.integers(0, 256, (n, no), dtype=np.uint8)
    priv[:, no - 1] &= 0x7F
    _, pts = b.bignPubkeyCalcBatch(p, priv)
    for d_len in (no, no // 2 + 1, 3):
        sc = rng.integers(0, 256, (n, d_len), dtype=np.uint8)
        sc[0] = 0
        sc[1] = 0
        sc[1, 0] = 1
        got, ok = b.ecMulABatch(pts, sc, l)
        for i in range(n):
            want_ok, want = o.ecMulA(pts[i].tobytes(), sc[i].tobytes(), l)
            assert ok[i] == want_ok
            if want_ok:
                assert got[i].tobytes() == want
        assert ok[0] == 0 and got[1].tobytes() == pts[1].tobytes()
    sc = np.stack([A(q.to_bytes(no, "little")), A((q + 1).to_bytes(no, "little"))])
    got, ok = b.ecMulABatch(pts[:2].copy(), sc, l)
    assert ok[0] == 0 and ok[1] == 1 and got[1].tobytes() == pts[1].tobytes()
    # d A + k G with A = a G: the result is ((d a + k) mod q) G, incl. d A = +-k G
    a = [1, 1] + [int.from_bytes(priv[i].tobytes(), "little") for i in range(2, 8)]
    d = [5, 7] + [int(rng.integers(1, 1 << 62)) for _ in range(6)]
    k = [5, q - 7] + [int.from_bytes(rng.integers(0, 256, no, dtype=np.uint8).tobytes(), "little") % q for _ in range(6)]
    _, apts = b.bignPubkeyCalcBatch(p, np.stack([A(x.to_bytes(no, "little")) for x in a]))
    got, ok = b.ecAddMulABatch(apts, np.stack([A(x.to_bytes(no, "little")) for x in d]),
                               np.stack([A(x.to_bytes(no, "little")) for x in k]), l)
    for i in range(8):
        e = (d[i] * a[i] + k[i]) % q
        if e == 0:
            assert ok[i] == 0
        else:
            assert ok[i] == 1 and (0, got[i].tobytes()) == o.bignPubkeyCalc(e.to_bytes(no, "little"), l)


@pytest.mark.parametrize("l", [128, 192, 256])
def test_keypair_gen_val_dh_fixtures(l):
    """bignKeypairGen / KeypairVal / PubkeyVal / DH (bign_misc.c:182-515) against outputs of the unmodified
    reference (ref_vectors.json "bignMisc"), incl. rejected generator draws and invalid public keys."""
    p = b.bignParamsStd(b.BIGN_CURVES[l])
    no = l // 4
    t = [t for t in REF["bignMisc"] if t["l"] == l][0]
    priv, pub, used = b.bignKeypairGenBatch(p, bytes.fromhex(t["stream"]), 1)
    assert (priv[0].tobytes().hex(), pub[0].tobytes().hex(), used) == (t["privkey"], t["pubkey"], t["used"])
    priv_b, pub_b = bytes.fromhex(t["privkey"]), bytes.fromhex(t["pubkey"])
    for d in t["dh"]:
        code, key = b.bignDH(p, bytes.fromhex(d["privkey"]), pub_b, d["key_len"])
        assert code == d["code"] and (code != 0 or key.hex() == d["key"])
    assert b.bignDH(p, priv_b, pub_b, 2 * no + 1)[0] == b.ERR_BAD_SHAREDKEY
    for v_ in t["val"]:
        pk = bytes.fromhex(v_["pubkey"])
        assert b.bignPubkeyVal(p, pk) == v_["pubkey_val"], v_["case"]
        assert b.bignKeypairVal(p, priv_b, pk) == v_["keypair_val"], v_["case"]
        assert b.bignDH(p, priv_b, pk, no)[0] == v_["dh"], v_["case"]


@pytest.mark.parametrize("l", [128, 192, 256])
def test_sign_with_generator_and_level_wrappers(l):
    """bignSign (bign_sign.c:27-138) with the caller's generator: reference fixture (rejected draws incl.),
    a batch against the oracle given the same one-time keys, and the fixed-level bign128/192/256 wrappers."""
    import ctypes as C
    p = b.bignParamsStd(b.BIGN_CURVES[l])
    no, oid = l // 4, o.OIDS[l]
    t = [t for t in REF["bignMisc"] if t["l"] == l][0]
    priv, pub, sg = bytes.fromhex(t["privkey"]), bytes.fromhex(t["pubkey"]), t["sign"]
    st, sigs, used = b.bignSignBatch(p, oid, A(bytes.fromhex(sg["hash"])), A(priv), bytes.fromhex(sg["stream"]))
    assert (int(st[0]), sigs[0].tobytes().hex(), used) == (0, sg["sig"], sg["used"])
    rng = np.random.default_rng(400 + l)
    n = 40
    privs = rng.integers(0, 256, (n, no), dtype=np.uint8)
    privs[:, no - 1] &= 0x7F
    privs[5] = 0                                                   # ERR_BAD_PRIVKEY: consumes no generator octets
    hashes = rng.integers(0, 256, (n, no), dtype=np.uint8)
    ks = rng.integers(0, 256, (n - 1, no), dtype=np.uint8)
    ks[:, no - 1] &= 0x7F
    st, sigs, used = b.bignSignBatch(p, oid, hashes, privs, ks.tobytes())
    assert used == (n - 1) * no and st[5] == b.ERR_BAD_PRIVKEY
    j = 0
    for i in range(n):
        if i == 5:
            continue
        assert o.bignSignK(hashes[i].tobytes(), privs[i].tobytes(), ks[j].tobytes(), oid, l) == (0, sigs[i].tobytes())
        j += 1
    # fixed-level wrappers through the C ABI
    L, pre = b.lib(), f"bign{l}"
    buf = C.create_string_buffer(2 * no)
    assert getattr(L, pre + "PubkeyCalc")(buf, priv) == 0 and buf.raw == pub
    assert getattr(L, pre + "PubkeyVal")(pub) == 0 and getattr(L, pre + "KeypairVal")(priv, pub) == 0
    sig = C.create_string_buffer(no + no // 2)
    h = bytes.fromhex(sg["hash"])
    assert getattr(L, pre + "Sign2")(sig, h, priv, None, 0) == 0
    assert (0, sig.raw) == o.bignSign2(h, priv, None, oid, l)
    assert getattr(L, pre + "Verify")(h, sig.raw, pub) == 0
    assert getattr(L, pre + "Verify")(h, sig.raw[:-1] + bytes([sig.raw[-1] ^ 1]), pub) == b.ERR_BAD_SIG
    key = C.create_string_buffer(no)
    assert getattr(L, pre + "DH")(key, priv, pub, no) == 0 and (0, key.raw) == o.bignDH(priv, pub, no, l)


@pytest.mark.skipif(o.ref() is None, reason="oracle/_ref/libbee2ref_64.so not built")
@pytest.mark.parametrize("l", [128, 192, 256])
def test_ecMulA_dropin_on_reference_ec_object(l):
    """The drop-in ecMulA (ec.h:892-901) takes an ec_o BUILT BY THE UNMODIFIED REFERENCE (bignEcCreate) and
    returns what the reference's own ecMulA returns on the same object: points as n-word arrays, scalars of
    m <= n words, FALSE for the point at infinity."""
    import ctypes as C
    R, L = o.ref(), b.lib()
    ec = C.c_void_p()
    assert R.bignEcCreate(C.byref(ec), C.byref(o.ref_params(l))) == 0
    try:
        no, n = l // 4, l // 32
        R.ecMulA_deep.restype = C.c_size_t
        L.ecMulA.restype = C.c_int
        L.ecMulA.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        R.ecMulA.restype = C.c_int
        R.ecMulA.argtypes = L.ecMulA.argtypes
        stack = C.create_string_buffer(1 << 20)
        rng = np.random.default_rng(500 + l)
        p = b.bignParamsStd(b.BIGN_CURVES[l])
        priv = rng.integers(0, 256, (6, no), dtype=np.uint8)
        priv[:, no - 1] &= 0x7F
        _, pts = b.bignPubkeyCalcBatch(p, priv)
        q = Q if l == 128 else CURVE_Q[l]
        for i in range(6):
            for m in (n, n // 2 + 1, 1):
                d = rng.integers(0, 256, 8 * m, dtype=np.uint8).tobytes()
                if i == 0 and m == n:
                    d = q.to_bytes(no, "little")               # q A = O -> FALSE
                if i == 1 and m == 1:
                    d = bytes(8)                                # 0 A = O -> FALSE
                a = pts[i].tobytes()
                got, want = C.create_string_buffer(2 * no), C.create_string_buffer(2 * no)
                r1 = L.ecMulA(got, a, ec, d, m, None)
                r2 = R.ecMulA(want, a, ec, d, m, stack)
                assert r1 == r2, (i, m)
                if r2:
                    assert got.raw == want.raw, (i, m)
        assert L.ecMulA_deep(n, 3, 0, n) == 0
        # ecAddMulA (variadic): the verification's own call shape s1 G + s0 Q (bign_sign.c:332), a 3-term sum,
        # a sum that cancels to O
        L.ecAddMulA.restype = C.c_int
        R.ecAddMulA.restype = C.c_int
        base = bytes(no) + bytes(p.yG)[:no]
        sz, vp = C.c_size_t, C.c_void_p

        def both(terms):
            args = []
            keep = []
            for a, d in terms:
                ba, bd = C.create_string_buffer(a, 2 * no), C.create_string_buffer(d, len(d))
                keep += [ba, bd]
                args += [C.cast(ba, vp), C.cast(bd, vp), sz(len(d) // 8)]
            got, want = C.create_string_buffer(2 * no), C.create_string_buffer(2 * no)
            r1 = L.ecAddMulA(C.cast(got, vp), ec, vp(None), sz(len(terms)), *args)
            r2 = R.ecAddMulA(C.cast(want, vp), ec, C.cast(stack, vp), sz(len(terms)), *args)
            assert r1 == r2 and (not r2 or got.raw == want.raw)
            return r1
        s1 = rng.integers(0, 256, no, dtype=np.uint8).tobytes()
        s0 = rng.integers(0, 256, no // 2, dtype=np.uint8).tobytes() + (1).to_bytes(8, "little")
        assert both([(base, s1), (pts[2].tobytes(), s0)]) == 1
        assert both([(pts[0].tobytes(), s1), (pts[1].tobytes(), s0), (pts[2].tobytes(), s1[:8])]) == 1
        neg = (q - 5).to_bytes(no, "little")
        assert both([(pts[3].tobytes(), (5).to_bytes(8, "little")), (pts[3].tobytes(), neg)]) == 0
    finally:
        R.bignEcClose(ec)


@pytest.mark.parametrize("l", [128, 256])
def test_dh_batch_vs_oracle(l):
    """d_i Q_j = d_j Q_i (both sides of the exchange agree), keys equal the oracle's, statuses item by item."""
    rng = np.random.default_rng(300 + l)
    p, n, no = b.bignParamsStd(b.BIGN_CURVES[l]), 96, l // 4
    stream = rng.integers(0, 256, 2 * n * no, dtype=np.uint8).tobytes()
    priv, pub, used = b.bignKeypairGenBatch(p, stream, n)
    assert used <= len(stream) and (b.bignKeypairValBatch(p, priv, pub) == 0).all()
    assert (b.bignPubkeyValBatch(p, pub) == 0).all()
    peer = np.roll(pub, 1, axis=0).copy()
    st1, k1 = b.bignDHBatch(p, priv, peer, 2 * no)
    st2, k2 = b.bignDHBatch(p, np.roll(priv, 1, axis=0).copy(), pub, 2 * no)
    assert (st1 == 0).all() and (st2 == 0).all() and np.array_equal(k1, k2)
    for i in range(0, n, 12):
        assert o.bignDH(priv[i].tobytes(), peer[i].tobytes(), 2 * no, l) == (0, k1[i].tobytes())
    # corrupted items: off-curve points, d = 0, d = q, x >= p
    q = (CURVE_Q[l] if l != 128 else Q)
    bad_pub, bad_priv = peer.copy(), priv.copy()
    bad_pub[::7, 5] ^= 1
    bad_priv[3] = 0
    bad_priv[10] = A(q.to_bytes(no, "little"))
    bad_pub[20, :no] = 0xFF
    st, keys = b.bignDHBatch(p, bad_priv, bad_pub, no)
    want = [o.bignDH(bad_priv[i].tobytes(), bad_pub[i].tobytes(), no, l) for i in range(n)]
    assert [int(x) for x in st] == [w[0] for w in want]
    assert all(keys[i].tobytes() == want[i][1] for i in range(n) if want[i][0] == 0)
    assert {0, 504, 505} <= set(int(x) for x in st)
    assert [int(x) for x in b.bignPubkeyValBatch(p, bad_pub)] == [o.bignPubkeyVal(bad_pub[i].tobytes(), l) for i in range(n)]
    kv = b.bignKeypairValBatch(p, bad_priv, pub)
    assert kv[3] == 504 and kv[10] == 504 and kv[0] == 0


def _corrupt(rng, hashes, sigs, pubs):
    """~1/3 of the items get one of the SURVEY §8d corruptions."""
    n = hashes.shape[0]
    for i in range(n):
        kind = int(rng.integers(0, 18))
        if kind == 0:
            sigs[i, int(rng.integers(0, 16))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            sigs[i, 16 + int(rng.integers(0, 32))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:
            hashes[i, int(rng.integers(0, 32))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 3:
            pubs[i, int(rng.integers(0, 64))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 4:
            sigs[i, 16:] = 0xFF                       # s1 >= q
        elif kind == 5:
            pubs[i, (0 if i & 1 else 32):][:32] = 0xFF  # Qx or Qy >= p


def test_verify_batch_random_vs_oracle():
    rng = np.random.default_rng(2)
    p, n = b.bignParamsStd(), 600
    priv = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    priv[:, 31] &= 0x7F
    hashes = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    hashes[::5, 16:] = 0xFF                                                    # H >= q path (bign_sign.c:320-326)
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    assert (st == 0).all()
    st, sigs = b.bignSign2Batch(p, OID, hashes, priv)
    assert (st == 0).all()
    for i in range(0, n, 40):
        assert o.bignPubkeyCalc(priv[i].tobytes()) == (0, pubs[i].tobytes())
        assert o.bignSign2(hashes[i].tobytes(), priv[i].tobytes()) == (0, sigs[i].tobytes())
    _corrupt(rng, hashes, sigs, pubs)
    got = b.bignVerifyBatch(p, OID, hashes, sigs, pubs)
    want = np.array([o.bignVerify(hashes[i].tobytes(), sigs[i].tobytes(), pubs[i].tobytes()) for i in range(n)],
                    dtype=np.uint32)
    assert np.array_equal(got, want)
    assert {0, 505, 510} <= set(int(x) for x in got)


def test_exceptional_points():
    """Keys that make the two halves of s1 G + s0 Q collide (Q = +-G, +-2G, ...) — SURVEY §7 hard parts."""
    p = b.bignParamsStd()
    rng = np.random.default_rng(8)
    ds = [1, 2, 3, Q - 1, Q - 2, 1 << 128, (1 << 128) + 1, Q - (1 << 128)]
    priv = np.stack([A(d.to_bytes(32, "little")) for d in ds])
    hashes = rng.integers(0, 256, (len(ds), 32), dtype=np.uint8)
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    assert (st == 0).all()
    G = bytes(32) + bytes(b.bignParamsStd().yG)[:32]
    assert pubs[0].tobytes() == G
    negG = bytes(32) + (P - int.from_bytes(G[32:], "little")).to_bytes(32, "little")
    assert pubs[3].tobytes() == negG
    st, sigs = b.bignSign2Batch(p, OID, hashes, priv)
    assert (st == 0).all()
    for i in range(len(ds)):
        assert o.bignSign2(hashes[i].tobytes(), priv[i].tobytes()) == (0, sigs[i].tobytes())
        assert o.bignPubkeyCalc(priv[i].tobytes()) == (0, pubs[i].tobytes())
    assert (b.bignVerifyBatch(p, OID, hashes, sigs, pubs) == 0).all()
    # R = O: s1 G + s0 Q with Q = G needs s1 + H = -(s0 + 2^128) mod q  -> BAD_SIG (bign_sign.c:332-336)
    s0 = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    hh = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    hq = int.from_bytes(hh, "little") % Q if int.from_bytes(hh, "little") >= Q else int.from_bytes(hh, "little")
    s1 = (-(int.from_bytes(s0, "little") + (1 << 128)) - hq) % Q
    sig = s0 + s1.to_bytes(32, "little")
    assert b.bignVerify(p, OID, hh, sig, G) == o.bignVerify(hh, sig, G) == b.ERR_BAD_SIG
    # same scalars against -G: R = 2 s0' G etc. — just has to agree with the oracle
    assert b.bignVerify(p, OID, hh, sig, negG) == o.bignVerify(hh, sig, negG)
    # all-zero inputs: Q = (0,0) is not on the curve; statuses still agree
    assert b.bignVerify(p, OID, bytes(32), bytes(48), bytes(64)) == o.bignVerify(bytes(32), bytes(48), bytes(64))


def test_ecMulA_batch_vs_oracle():
    rng = np.random.default_rng(6)
    p, n = b.bignParamsStd(), 48
    priv = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    priv[:, 31] &= 0x7F
    _, pts = b.bignPubkeyCalcBatch(p, priv)
    for d_len in (32, 16, 2):
        sc = rng.integers(0, 256, (n, d_len), dtype=np.uint8)
        sc[0] = 0                                        # 0 * A = O -> ok = 0 (ec.c:497-525 returns FALSE)
        sc[1] = 0
        sc[1, 0] = 1                                     # 1 * A = A
        got, ok = b.ecMulABatch(pts, sc)
        for i in range(n):
            want_ok, want = o.ecMulA(pts[i].tobytes(), sc[i].tobytes())
            assert ok[i] == want_ok
            if want_ok:
                assert got[i].tobytes() == want
        assert ok[0] == 0 and got[1].tobytes() == pts[1].tobytes()
    # q * A = O and (q + 1) * A = A: scalars may exceed the group order
    sc = np.stack([A(Q.to_bytes(32, "little")), A((Q + 1).to_bytes(32, "little"))])
    got, ok = b.ecMulABatch(pts[:2].copy(), sc)
    assert ok[0] == 0 and ok[1] == 1 and got[1].tobytes() == pts[1].tobytes()


def test_ecAddMulA_batch_vs_oracle():
    """d A + k G (ec.c:1183-1273) with A = a G of known a: the result must be ((d a + k) mod q) G,
    including the collisions d A = +-k G (result 2kG resp. O) that need the complete addition."""
    rng = np.random.default_rng(12)
    p, n = b.bignParamsStd(), 64
    a = [int.from_bytes(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), "little") % (Q - 1) + 1 for _ in range(n)]
    d = [int.from_bytes(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), "little") % Q for _ in range(n)]
    k = [int.from_bytes(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), "little") % Q for _ in range(n)]
    a[0], d[0], k[0] = 1, 5, 5                    # d A = k G  -> doubling branch
    a[1], d[1], k[1] = 1, 7, Q - 7                # d A = -k G -> O
    a[2], d[2], k[2] = 3, 0, 9                    # d = 0 -> k G
    a[3], d[3], k[3] = 3, 9, 0                    # k = 0 -> d A
    a[4], d[4], k[4] = 2, 0, 0                    # O
    _, A_ = b.bignPubkeyCalcBatch(p, np.stack([A(x.to_bytes(32, "little")) for x in a]))
    dm = np.stack([A(x.to_bytes(32, "little")) for x in d])
    km = np.stack([A(x.to_bytes(32, "little")) for x in k])
    got, ok = b.ecAddMulABatch(A_, dm, km)
    for i in range(n):
        s = (d[i] * a[i] + k[i]) % Q
        if s == 0:
            assert ok[i] == 0, i
        else:
            assert ok[i] == 1 and (0, got[i].tobytes()) == o.bignPubkeyCalc(s.to_bytes(32, "little")), i


@pytest.mark.skipif(o.ref() is None, reason="oracle/_ref/libbee2ref_64.so not on this machine")
def test_config4_shape_vs_reference_harness():
    """A 2^13 slice of BASELINE config 4 (2^18 sigs): every status equals the unmodified
    reference's bign128Verify, run multi-threaded through oracle/cpu_harness.c."""
    rng = np.random.default_rng(4)
    p, n = b.bignParamsStd(), 1 << 13
    priv = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    priv[:, 31] &= 0x7F
    hashes = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    _, pubs = b.bignPubkeyCalcBatch(p, priv)
    st, sigs = b.bignSign2Batch(p, OID, hashes, priv)
    assert (st == 0).all()
    _corrupt(rng, hashes, sigs, pubs)
    got = b.bignVerifyBatch(p, OID, hashes, sigs, pubs)
    hs = C.CDLL(os.path.join(o.REF_DIR, "libcpuharness.so"))
    hs.harness_bign_verify.restype = C.c_double
    want = np.zeros(n, dtype=np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    dt = hs.harness_bign_verify(os.path.join(o.REF_DIR, "libbee2ref_64.so").encode(), 0, vp(want), vp(hashes),
                                vp(sigs), vp(pubs), C.c_size_t(n), os.cpu_count() or 1)
    assert dt > 0
    assert np.array_equal(got, want)
    assert (got == 0).sum() > n // 2


@pytest.mark.parametrize("l", [128, 192, 256])
def test_verify_ragged_counts_staged_and_pinned(l):
    """Odd and ragged batch sizes through every input path of the verification kernel: pageable host buffers
    (staging copies), pinned host buffers (zero-copy, TMA bulk copies straight from host memory), device
    buffers, and a misaligned device view (per-thread loads). At l = 192 a signature is 72 octets, so a
    ragged last CTA with an odd item count cannot be a bulk copy (multiples of 16 octets) and reads directly."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(7 * l)
    p, no, oid = b.bignParamsStd(b.BIGN_CURVES[l]), l // 4, o.OIDS[l]
    nmax = 515
    priv = rng.integers(0, 256, (nmax, no), dtype=np.uint8)
    priv[:, no - 1] &= 0x7F
    hashes = rng.integers(0, 256, (nmax, no), dtype=np.uint8)
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    st2, sigs = b.bignSign2Batch(p, oid, hashes, priv)
    assert not st.any() and not st2.any()
    sigs[::3, 1] ^= 4
    want_all = np.array([o.bignVerify(hashes[i].tobytes(), sigs[i].tobytes(), pubs[i].tobytes(), oid, l)
                         for i in range(nmax)], dtype=np.uint32)
    assert {0, 510} <= set(int(x) for x in want_all)
    stream = torch.cuda.current_stream().cuda_stream
    L = b.lib()
    ko = np.frombuffer(oid, dtype=np.uint8)
    for n in (1, 2, 3, 31, 127, 129, 255, 257, 383, 515):
        want = want_all[:n]
        # pageable host buffers
        assert np.array_equal(b.bignVerifyBatch(p, oid, hashes[:n], sigs[:n], pubs[:n]), want), (l, n, "pageable")
        # pinned host buffers: zero-copy
        ph, ps, pp = (torch.from_numpy(x[:n].copy()).pin_memory() for x in (hashes, sigs, pubs))
        pst = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        got = b.bignVerifyBatch(p, oid, ph.numpy(), ps.numpy(), pp.numpy(), status=pst)
        assert np.array_equal(got, want), (l, n, "pinned")
        # device buffers, aligned and shifted by one octet
        for shift in (0, 1):
            bufs = []
            for x in (hashes, sigs, pubs):
                t = torch.zeros(x[:n].size + 16, dtype=torch.uint8, device="cuda")
                t[shift:shift + x[:n].size] = torch.from_numpy(x[:n].reshape(-1)).cuda()
                bufs.append(t)
            dst = torch.full((n,), 77, dtype=torch.int32, device="cuda")
            assert L.b2g_bignVerifyBatchL_dev(l, dst.data_ptr(), ko.ctypes.data, len(oid), bufs[0].data_ptr() + shift,
                                              bufs[1].data_ptr() + shift, bufs[2].data_ptr() + shift, n, stream) == 0
            assert np.array_equal(dst.cpu().numpy().view(np.uint32), want), (l, n, "device", shift)


@pytest.mark.parametrize("l", [128, 192, 256])
def test_sign2_ragged_counts_staged_and_pinned(l):
    """Every input path of the signing kernel at odd and ragged batch sizes: pageable host buffers (staging
    copies), pinned host buffers (zero-copy: inputs by TMA bulk copies, signatures by one bulk store per CTA),
    device buffers aligned (staged) and shifted by one octet (direct loads). Items that fail (d = 0, d >= q)
    must leave their signature slot untouched (bign_sign.c:189-194), also inside a CTA whose other items
    succeed — the kernel then drops the bulk store and writes item by item."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11 * l)
    p, no, oid = b.bignParamsStd(b.BIGN_CURVES[l]), l // 4, o.OIDS[l]
    so, nmax = 3 * no // 2, 515
    priv = rng.integers(0, 256, (nmax, no), dtype=np.uint8)
    priv[:, no - 1] &= 0x7F
    hashes = rng.integers(0, 256, (nmax, no), dtype=np.uint8)
    want_sig = np.zeros((nmax, so), dtype=np.uint8)
    step = 5 if l == 128 else 40
    checked = sorted(set(range(0, nmax, step)) | {1, 2, 30, 126, 128, 254, 256, 382, 514})
    for i in checked:
        rc, s = o.bignSign2(hashes[i].tobytes(), priv[i].tobytes(), None, oid, l)
        assert rc == 0
        want_sig[i] = A(s)
    stream = torch.cuda.current_stream().cuda_stream
    L = b.lib()
    ko = np.frombuffer(oid, dtype=np.uint8)
    for bad in ((), (0, 300)):
        pv = priv.copy()
        for j, i in enumerate(bad):
            pv[i] = 0 if j == 0 else 0xFF
        for n in (1, 2, 3, 31, 127, 129, 255, 257, 383, 515):
            ok = [i for i in checked if i < n and i not in bad]
            dead = [i for i in bad if i < n]

            def check(st, sg, tag):
                assert all(st[i] == 0 for i in range(n) if i not in bad), (l, n, tag)
                assert all(st[i] == b.ERR_BAD_PRIVKEY for i in dead), (l, n, tag)
                assert np.array_equal(sg[ok], want_sig[ok]), (l, n, tag)
                assert all((sg[i] == 0xA5).all() for i in dead), (l, n, tag, "failed item was written")

            # pageable host buffers
            sg = np.full((n, so), 0xA5, dtype=np.uint8)
            st, sg = b.bignSign2Batch(p, oid, hashes[:n], pv[:n], sigs=sg)
            check(st, sg, "pageable")
            # pinned host buffers: zero-copy
            ph, pk = (torch.from_numpy(x[:n].copy()).pin_memory() for x in (hashes, pv))
            psg = torch.full((n, so), 0xA5, dtype=torch.uint8).pin_memory()
            pst = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
            st, sg = b.bignSign2Batch(p, oid, ph.numpy(), pk.numpy(), status=pst, sigs=psg.numpy())
            check(st, sg, "pinned")
            # device buffers, aligned and shifted by one octet
            for shift in (0, 1):
                bufs = []
                for x in (hashes, pv):
                    t = torch.zeros(x[:n].size + 16, dtype=torch.uint8, device="cuda")
                    t[shift:shift + x[:n].size] = torch.from_numpy(x[:n].reshape(-1)).cuda()
                    bufs.append(t)
                dsg = torch.full((n * so + 16,), 0xA5, dtype=torch.uint8, device="cuda")
                dst = torch.full((n,), 77, dtype=torch.int32, device="cuda")
                assert L.b2g_bignSign2BatchL_t_dev(l, dst.data_ptr(), dsg.data_ptr() + shift, ko.ctypes.data, len(oid),
                                                   bufs[0].data_ptr() + shift, bufs[1].data_ptr() + shift, n, None, 0,
                                                   stream) == 0
                sg = dsg.cpu().numpy()[shift:shift + n * so].reshape(n, so)
                assert (dsg.cpu().numpy()[shift + n * so:] == 0xA5).all() and (dsg.cpu().numpy()[:shift] == 0xA5).all()
                check(dst.cpu().numpy().view(np.uint32), sg, ("device", shift))


def test_verify_forced_staging_on_device_buffers():
    """The launcher stages (TMA bulk copies) only inputs that lie in host memory; B2G_FORCE_STAGING=1 makes it
    stage device-resident inputs too (the switch is read once per process, hence the subprocess). Both
    builds of the kernel must give the statuses of the default path on ragged sizes."""
    import subprocess
    import sys
    script = r'''
import numpy as np, torch, sys
sys.path.insert(0, "tests")
import bee2_b200 as b, _oracle as o
rng = np.random.default_rng(5)
out = []
for l in (128, 192, 256):
    p, no, oid = b.bignParamsStd(b.BIGN_CURVES[l]), l // 4, o.OIDS[l]
    n = 700
    priv = rng.integers(0, 256, (n, no), dtype=np.uint8); priv[:, no - 1] &= 0x7F
    hashes = rng.integers(0, 256, (n, no), dtype=np.uint8)
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    st2, sigs = b.bignSign2Batch(p, oid, hashes, priv)
    assert not st.any() and not st2.any()
    sigs[::3, 1] ^= 4
    ko = np.frombuffer(oid, dtype=np.uint8)
    stream = torch.cuda.current_stream().cuda_stream
    for m in (1, 255, 256, 257, 700):
        d = [torch.from_numpy(x[:m].copy()).cuda() for x in (hashes, sigs, pubs)]
        dst = torch.full((m,), 77, dtype=torch.int32, device="cuda")
        assert b.lib().b2g_bignVerifyBatchL_dev(l, dst.data_ptr(), ko.ctypes.data, len(oid), d[0].data_ptr(), d[1].data_ptr(),
                                                d[2].data_ptr(), m, stream) == 0
        out.append(dst.cpu().numpy().view(np.uint32).tobytes().hex())
print("\n".join(out))
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runs = []
    for force in ("", "1"):
        env = dict(os.environ)
        env.pop("B2G_FORCE_STAGING", None)
        if force:
            env["B2G_FORCE_STAGING"] = force
        r = subprocess.run([sys.executable, "-c", script], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        runs.append(r.stdout.split())
    assert runs[0] == runs[1] and len(runs[0]) == 15
    first = np.frombuffer(bytes.fromhex(runs[0][4]), dtype=np.uint32)      # l = 128, 700 items
    assert (first[::3] == 510).all() and (np.delete(first, np.arange(0, 700, 3)) == 0).all()
