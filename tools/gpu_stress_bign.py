"""One-off stress (run from the repo root on a GPU box): bign verify / sign2 / pubkey / DH on large random batches of all
three levels, every verification status compared with the unmodified reference (oracle/_ref).
    python tools/gpu_stress_bign.py
"""
import sys, numpy as np, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import _oracle as o, bee2_b200 as b
assert b.b2g_init(0) == 0
for l, n in ((128, 1 << 18), (192, 1 << 17), (256, 1 << 16)):
    p = b.bignParamsStd(b.BIGN_CURVES[l]); no = l // 4; oid = o.OIDS[l]
    rng = np.random.default_rng(7000 + l)
    priv = rng.integers(0, 256, (n, no), dtype=np.uint8); priv[:, no - 1] &= 0x7F
    hashes = rng.integers(0, 256, (n, no), dtype=np.uint8)
    st, pub = b.bignPubkeyCalcBatch(p, priv); assert not st.any()
    st, sig = b.bignSign2Batch(p, oid, hashes, priv); assert not st.any()
    bad = np.arange(0, n, 9); sig[bad, bad % (no + no // 2)] ^= 1 << (bad % 8).astype(np.uint8)
    got = b.bignVerifyBatch(p, oid, hashes, sig, pub)
    t0 = time.time()
    want = np.array([o.ref_bignVerify(hashes[i].tobytes(), sig[i].tobytes(), pub[i].tobytes(), oid, l) for i in range(n)], dtype=np.uint32)
    assert np.array_equal(got, want), (l, np.nonzero(got != want)[0][:10])
    # signatures themselves and public keys vs the reference on a stride
    for i in range(0, n, 97):
        assert o.ref_bignSign2(hashes[i].tobytes(), priv[i].tobytes(), None, oid, l)[1] == sig[i].tobytes() or i in bad
        assert o.ref_bignPubkeyCalc(priv[i].tobytes(), l) == (0, pub[i].tobytes())
    # DH both ways
    peer = np.roll(pub, 1, axis=0).copy()
    st, k = b.bignDHBatch(p, priv, peer, 2 * no); assert not st.any()
    for i in range(0, n, 53):
        assert o.ref_bignDH(priv[i].tobytes(), peer[i].tobytes(), 2 * no, l) == (0, k[i].tobytes())
    print(f"l={l}: {n} verifies == reference ({int((want == 0).sum())} ok, {int((want != 0).sum())} rejected), "
          f"sign2/pubkey/DH samples equal; ref time {time.time() - t0:.1f} s", flush=True)
