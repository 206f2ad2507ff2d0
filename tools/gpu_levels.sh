#!/bin/bash
# device-resident verify / sign2 rates for the three bign levels (2^16 items each)
python - <<'PY'
import numpy as np, torch, bee2_b200 as b, sys
sys.path.insert(0, "tests")
import _oracle as o
assert b.b2g_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
n = 1 << 16
for l in (128, 192, 256):
    p = b.bignParamsStd(b.BIGN_CURVES[l]); no = l // 4; oid = o.OIDS[l]
    rng = np.random.default_rng(l)
    priv = rng.integers(0, 256, (n, no), dtype=np.uint8); priv[:, no - 1] &= 0x7F
    hashes = rng.integers(0, 256, (n, no), dtype=np.uint8)
    st, pub = b.bignPubkeyCalcBatch(p, priv); assert not st.any()
    st, sig = b.bignSign2Batch(p, oid, hashes, priv); assert not st.any()
    dh, ds, dp, dk = (torch.from_numpy(x).cuda() for x in (hashes, sig, pub, priv))
    dst = torch.empty(n, dtype=torch.int32, device="cuda"); dsig = torch.zeros_like(ds)
    L = b.lib(); ko = np.frombuffer(oid, dtype=np.uint8)
    def verify(): assert L.b2g_bignVerifyBatchL_dev(l, dst.data_ptr(), ko.ctypes.data, len(oid), dh.data_ptr(), ds.data_ptr(), dp.data_ptr(), n, stream) == 0
    def sign(): assert L.b2g_bignSign2BatchL_t_dev(l, dst.data_ptr(), dsig.data_ptr(), ko.ctypes.data, len(oid), dh.data_ptr(), dk.data_ptr(), n, None, 0, stream) == 0
    for name, fn in (("verify", verify), ("sign2", sign)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        assert not dst.cpu().numpy().any()
        print(f"l={l} {name}: {ms:.3f} ms per 2^16 = {n / ms / 1e3:.2f} M/s")
PY
