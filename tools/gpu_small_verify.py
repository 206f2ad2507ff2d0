#!/usr/bin/env python3
"""Device-resident bign verify rate at small batch sizes (a shard of a strong-scaled job): one line per size.
   BEE2_B200_LIB selects a variant build."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bee2_b200 as b

assert b.b2g_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
OID = bytes.fromhex("06092A7000020022651F51")
sizes = [int(a) for a in sys.argv[1:]] or [1 << 15, 1 << 16, 1 << 17, 1 << 18]
nmax = max(sizes + [1 << 18])
rng = np.random.default_rng(2)
priv = rng.integers(0, 256, (nmax, 32), dtype=np.uint8)
priv[:, 31] &= 0x7F
hashes = rng.integers(0, 256, (nmax, 32), dtype=np.uint8)
p = b.bignParamsStd()
st, pubs = b.bignPubkeyCalcBatch(p, priv)
st2, sigs = b.bignSign2Batch(p, OID, hashes, priv)
assert not st.any() and not st2.any()
sigs[::16, 5] ^= 1
d_h, d_s, d_p = (torch.from_numpy(x).cuda() for x in (hashes, sigs, pubs))
d_st = torch.empty(nmax, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for n in sizes:
    fn = lambda: b.bignVerifyBatch_dev(d_st.data_ptr(), OID, d_h.data_ptr(), d_s.data_ptr(), d_p.data_ptr(), n, stream)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    stc = d_st[:n].cpu().numpy()
    assert (stc[::16] == 510).all() and int((stc == 0).sum()) == n - len(stc[::16])
    ms = float(np.median(ts))
    out.append(f"n={n}: {ms:.3f} ms = {n / ms / 1e3:.2f} M/s")
print("; ".join(out))
