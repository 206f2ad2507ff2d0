#!/usr/bin/env python3
"""Device-resident time of the l = 128 verification and signing kernels at 2^18 items WITHOUT checking results
(for timing experiments with builds that skip work, e.g. -DBIGN_FAKE_TREE). BEE2_B200_LIB selects the build."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bee2_b200 as b

assert b.b2g_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
OID = bytes.fromhex("06092A7000020022651F51")
n = 1 << 18
rng = np.random.default_rng(2)
priv = rng.integers(0, 256, (n, 32), dtype=np.uint8)
priv[:, 31] &= 0x7F
hashes = rng.integers(0, 256, (n, 32), dtype=np.uint8)
sigs = rng.integers(0, 256, (n, 48), dtype=np.uint8)
sigs[:, 47] &= 0x7F
pubs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
pubs[:, 31] &= 0x7F
pubs[:, 63] &= 0x7F
d_h, d_s, d_p, d_k = (torch.from_numpy(x).cuda() for x in (hashes, sigs, pubs, priv))
d_st = torch.empty(n, dtype=torch.int32, device="cuda")
d_sig = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ko = np.frombuffer(OID, dtype=np.uint8)
L = b.lib()
fns = {"verify": lambda: L.b2g_bignVerifyBatchL_dev(128, d_st.data_ptr(), ko.ctypes.data, len(OID), d_h.data_ptr(), d_s.data_ptr(), d_p.data_ptr(), n, stream),
       "sign2": lambda: L.b2g_bignSign2BatchL_t_dev(128, d_st.data_ptr(), d_sig.data_ptr(), ko.ctypes.data, len(OID), d_h.data_ptr(), d_k.data_ptr(), n, None, 0, stream)}
out = []
for name, fn in fns.items():
    for _ in range(3):
        assert fn() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    out.append(f"{name}: {ms:.3f} ms = {n / ms / 1e3:.2f} M/s")
print("; ".join(out))
