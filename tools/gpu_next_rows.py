#!/usr/bin/env python3
"""Measured rates of the SURVEY §8(f) "next" rows that bench.py's contract line does not carry: belt-hash
batch, ragged bash batch, bash-prg batch, bign public-key / key-pair generation / Diffie-Hellman / ecMulA
batches and the levels l = 192 / 256. Device-resident (CUDA events, 3 warm-ups, median of 7) where a
`b2g_*_dev` entry exists, else the host-pointer call; beside each the UNMODIFIED reference
(oracle/_ref/libbee2ref_64.so through tests/_oracle.py) on ONE host thread over a small sample of the
same inputs, whose outputs are compared with the engine's. One JSON document on stdout:
    python tools/gpu_next_rows.py > gpurun_out/next_rows.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bee2_b200 as b
import _oracle as o

assert b.b2g_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
L = b.lib()
rows = []


def dev_time(fn, reps=7):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def host_time(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def cpu_rate(fn, n_units, budget=0.5):
    """units/s of the reference on one thread: fn(i) handles unit i; at least 8 units, about `budget` seconds"""
    done, t0 = 0, time.perf_counter()
    while done < n_units and (done < 8 or time.perf_counter() - t0 < budget):
        fn(done)
        done += 1
    return done / (time.perf_counter() - t0), done


def row(name, unit, value, how, cpu=None, cpu_n=0, note=None):
    r = {"row": name, "value": value, "unit": unit, "how": how}
    if cpu is not None:
        r["reference_1_thread"] = cpu
        r["reference_sample_units"] = cpu_n
        r["ratio_to_1_thread"] = value / cpu
    if note:
        r["note"] = note
    rows.append(r)
    print(json.dumps(r), file=sys.stderr)


rng = np.random.default_rng(77)
have_ref = o.ref() is not None

# ---------------------------------------------------------------- belt-hash batch (§8f rank 2)
n, mlen = 1 << 20, 1024
msgs = torch.randint(0, 256, (n, mlen), dtype=torch.uint8, device="cuda")
out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
t = dev_time(lambda: b.beltHashBatch_dev(out.data_ptr(), msgs.data_ptr(), mlen, mlen, n, stream))
hm, ho = msgs[:64].cpu().numpy(), out[:64].cpu().numpy()
if have_ref:
    c, cn = cpu_rate(lambda i: o.ref_beltHash(hm[i % 64].tobytes()), 1 << 30)
    assert all(o.ref_beltHash(hm[i].tobytes()) == ho[i].tobytes() for i in range(64))
    row("belt-hash batch, 2^20 messages x 1 KiB", "GB/s", n * mlen / t / 1e9, "device-resident, b2g_beltHashBatch_dev", c * mlen / 1e9, cn)
else:
    row("belt-hash batch, 2^20 messages x 1 KiB", "GB/s", n * mlen / t / 1e9, "device-resident, b2g_beltHashBatch_dev")
del msgs, out

# ---------------------------------------------------------------- ragged bash batch (§8f rank 3, the bsum case)
n = 1 << 16
lens = np.sort(rng.integers(0, 8192, n).astype(np.uint64))
offs = np.zeros(n, dtype=np.uint64)
offs[1:] = np.cumsum((lens[:-1] + 7) & ~np.uint64(7))
total = int(offs[-1] + lens[-1])
data = rng.integers(0, 256, total, dtype=np.uint8)
pd = b.pinned_empty(total)
pd[:] = data
t = host_time(lambda: b.bashHashBatchV(256, pd, offs, lens))
got = b.bashHashBatchV(256, pd, offs, lens)
if have_ref:
    idx = list(range(0, n, n // 48))
    assert all(o.ref_bashHash(256, data[int(offs[i]):int(offs[i] + lens[i])].tobytes()) == got[i].tobytes() for i in idx)
    c, cn = cpu_rate(lambda i: o.ref_bashHash(256, data[int(offs[idx[i % 48]]):int(offs[idx[i % 48]] + lens[idx[i % 48]])].tobytes()), 1 << 30)
    mean_len = float(np.mean([lens[i] for i in idx]))
    row("bash-512 ragged batch (bashHashBatchV), 2^16 messages of 0..8 KiB, sorted by length", "GB/s", int(lens.sum()) / t / 1e9,
        "host-pointer call, pinned input (copies inside)", c * mean_len / 1e9, cn, "reference: BASH_64 build, one thread")
else:
    row("bash-512 ragged batch (bashHashBatchV), 2^16 messages of 0..8 KiB, sorted by length", "GB/s", int(lens.sum()) / t / 1e9,
        "host-pointer call, pinned input (copies inside)")

# ---------------------------------------------------------------- bash-prg batch (§8f rank 3)
n, blocks = 1 << 16, 16
p0 = b.BashPrg(128, 1, b"", b"")
st0 = np.asarray(p0.state)                  # bash_prg_st: l, d (2 x size_t), s[192], buf_len, pos, t[192]
buf_len = int(st0[208:216].view(np.uint64)[0])
d_states = torch.from_numpy(np.tile(st0[16:208], (n, 1))).cuda()
d_data = torch.randint(0, 256, (n, blocks * buf_len), dtype=torch.uint8, device="cuda")
import ctypes as C
L.b2g_bashPrgBlocks_dev.restype = C.c_uint32
L.b2g_bashPrgBlocks_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_void_p]


def prg():
    assert L.b2g_bashPrgBlocks_dev(d_states.data_ptr(), d_data.data_ptr(), blocks * buf_len, blocks, buf_len, 2, 1, n, stream) == 0


t = dev_time(prg)
row(f"bash-prg batch: 2^16 automata (l = 128, d = 1) x 16 blocks of {buf_len} octets, encrypt", "GB/s", n * blocks * buf_len / t / 1e9,
    "device-resident, b2g_bashPrgBlocks_dev")
del d_states, d_data

# ---------------------------------------------------------------- bign: fixed-base and variable-base batches
for l in (128, 192, 256):
    p, no, oid = b.bignParamsStd(b.BIGN_CURVES[l]), l // 4, o.OIDS[l]
    n = 1 << 16
    priv = rng.integers(0, 256, (n, no), dtype=np.uint8)
    priv[:, no - 1] &= 0x7F
    hashes = rng.integers(0, 256, (n, no), dtype=np.uint8)
    st, pubs = b.bignPubkeyCalcBatch(p, priv)
    st2, sigs = b.bignSign2Batch(p, oid, hashes, priv)
    assert not st.any() and not st2.any()
    d_k, d_h, d_s, d_q = (torch.from_numpy(x).cuda() for x in (priv, hashes, sigs, pubs))
    d_st = torch.empty(n, dtype=torch.int32, device="cuda")
    d_pub = torch.empty(n * 2 * no, dtype=torch.uint8, device="cuda")
    ko = np.frombuffer(oid, dtype=np.uint8)
    if l == 128:
        t = dev_time(lambda: b.bignPubkeyCalcBatch_dev(d_st.data_ptr(), d_pub.data_ptr(), d_k.data_ptr(), n, stream))
        c = cn = None
        if have_ref:
            c, cn = cpu_rate(lambda i: o.ref_bignPubkeyCalc(priv[i].tobytes(), l), n)
            assert all(o.ref_bignPubkeyCalc(priv[i].tobytes(), l) == (0, pubs[i].tobytes()) for i in range(8))
        row("bignPubkeyCalc batch (k G, regular form), l = 128, 2^16 keys", "keys/s", n / t, "device-resident, b2g_bignPubkeyCalcBatch_dev", c, cn or 0)
        # key pairs from the caller's generator (drawn on the host exactly as zzRandNZMod does), then the same kernel
        stream_bytes = rng.integers(0, 256, 40 * n, dtype=np.uint8).tobytes()
        t = host_time(lambda: b.bignKeypairGenBatch(p, stream_bytes, n))
        row("bignKeypairGen batch, l = 128, 2^16 pairs (generator calls on the host)", "pairs/s", n / t, "host-pointer call",
            note="bound by the generator: here a Python ctypes callback per draw (gen_i, defs.h:520-524), not by the device")
        # Diffie-Hellman: d Q, regular variable-base ladder with masked table scan
        other = np.roll(pubs, 1, axis=0).copy()
        t = host_time(lambda: b.bignDHBatch(p, priv, other, 32))
        stc, keys = b.bignDHBatch(p, priv, other, 32)
        c = cn = None
        if have_ref:
            c, cn = cpu_rate(lambda i: o.ref_bignDH(priv[i].tobytes(), other[i].tobytes(), 32, l), n)
            assert all(o.ref_bignDH(priv[i].tobytes(), other[i].tobytes(), 32, l) == (0, keys[i].tobytes()) for i in range(8))
        row("bignDH batch (d Q, regular form), l = 128, 2^16 pairs", "keys/s", n / t, "host-pointer call (copies inside)", c, cn or 0)
    else:
        tv = dev_time(lambda: L.b2g_bignVerifyBatchL_dev(l, d_st.data_ptr(), ko.ctypes.data, len(oid), d_h.data_ptr(), d_s.data_ptr(), d_q.data_ptr(), n, stream), reps=5)
        assert not d_st.cpu().numpy().any()
        d_sig = torch.empty(n * 3 * no // 2, dtype=torch.uint8, device="cuda")
        ts = dev_time(lambda: L.b2g_bignSign2BatchL_t_dev(l, d_st.data_ptr(), d_sig.data_ptr(), ko.ctypes.data, len(oid), d_h.data_ptr(), d_k.data_ptr(), n, None, 0, stream), reps=5)
        cv = cs = cnv = cns = None
        if have_ref:
            cv, cnv = cpu_rate(lambda i: o.ref_bignVerify(hashes[i].tobytes(), sigs[i].tobytes(), pubs[i].tobytes(), oid, l), n)
            cs, cns = cpu_rate(lambda i: o.ref_bignSign2(hashes[i].tobytes(), priv[i].tobytes(), None, oid, l), n)
            assert all(o.ref_bignSign2(hashes[i].tobytes(), priv[i].tobytes(), None, oid, l) == (0, sigs[i].tobytes()) for i in range(4))
        row(f"bign verify batch, l = {l} (bign{l}), 2^16 signatures", "verifies/s", n / tv, "device-resident, b2g_bignVerifyBatchL_dev", cv, cnv or 0)
        row(f"bign sign2 batch, l = {l} (bign{l}), 2^16 signatures", "signatures/s", n / ts, "device-resident, b2g_bignSign2BatchL_t_dev", cs, cns or 0)

print(json.dumps({"what": "SURVEY §8(f) rows outside bench.py's contract line, one B200", "reference": "unmodified bee2 (oracle/_ref/libbee2ref_64.so), ONE host thread, ctypes call per unit",
                  "rows": rows}, indent=1))
