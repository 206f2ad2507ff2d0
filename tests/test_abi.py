"""The drop-in boundary without a GPU: libbee2_b200.so loads, exports every symbol that
include/bee2_b200.h declares, keeps the reference's state sizes, and FAILS LOUDLY (no CPU
fallback) when no CUDA device is usable. No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import bee2_b200 as b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bee2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.strip().startswith("#"))
    src = re.sub(r"typedef\s+[^;{]*\(\s*\*\s*\w+\s*\)\s*\([^;]*\)\s*;", "", src)      # function-pointer typedefs (gen_i)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src))
    names |= set(re.findall(r"extern\s+const\s+char\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[", src))
    return names - {"defined", "sizeof"}


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert {"bashHash", "bashHashStepH", "beltCTRStart", "beltCTRStepE", "beltECBEncr", "bignVerify", "bignSign2",
            "bashHashBatch", "beltECBEncrBatch", "bignVerifyBatch", "ecMulABatch", "b2g_beltCTR_dev",
            "b2g_bignVerifyBatch_dev", "bash_platform"} <= names
    out = subprocess.run(["nm", "-D", "--defined-only", b.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    missing = sorted(names - exported)
    assert not missing, missing


def test_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", b.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "bee2ref" not in out and "bee2oracle" not in out
    syms = subprocess.run(["nm", "-D", b.lib_path()], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_state_sizes_match_reference_layouts():
    L = b.lib()
    assert L.bashHash_keep() == 192 + 192 + 2 * C.sizeof(C.c_size_t)      # bash_hash.c:25-31
    assert L.beltCTR_keep() == 32 + 16 + 16 + C.sizeof(C.c_size_t)        # belt_lcl.h:135-141
    assert L.beltECB_keep() == 32 + 16                                     # belt_ecb.c:44-48
    assert L.bashF_deep() == 0
    assert C.sizeof(b.BignParams) == 8 + 5 * 64 + 8                        # bign.h:65-74
    assert (C.c_char * 17).in_dll(L, "bash_platform").value == b"BASH_CUDA_SM100A"


def test_host_side_argument_checks_need_no_device():
    # parameter / OID / level checks come before any device work, in the reference's order
    with pytest.raises(b.Bee2Error) as e:
        b.bashHash(7, b"abc")
    assert e.value.code == b.ERR_BAD_PARAMS
    p = b.bignParamsStd()
    assert p.l == 128
    bad = b.bignParamsStd()
    bad.q[0] ^= 1                                  # q even
    assert b.bignVerify(bad, b.OID_BELT_HASH_DER, bytes(32), bytes(48), bytes(64)) == b.ERR_BAD_PARAMS
    bad = b.bignParamsStd()
    bad.l = 192                                    # another (valid-looking) level: not on the GPU path
    assert b.bignVerify(bad, b.OID_BELT_HASH_DER, bytes(32), bytes(48), bytes(64)) in (b.ERR_BAD_PARAMS, b.ERR_NOT_IMPLEMENTED)
    for der in (b"", b"\x06", b"\x06\x01", b"\x06\x02\x80\x01", b"\x05\x01\x00", b"\x06\x81\x01\x2a", b"\x06\x01\x2a\x00"):
        assert b.bignVerify(p, der, bytes(32), bytes(48), bytes(64)) == b.ERR_BAD_OID, der
    # key expansion is host-side formatting (belt_block.c:88-106)
    k = np.frombuffer(bytes(range(24)), dtype=np.uint32)
    e = b.beltKeyExpand2(bytes(range(24)))
    assert e[6] == k[0] ^ k[1] ^ k[2] and e[7] == k[3] ^ k[4] ^ k[5]
    e = b.beltKeyExpand2(bytes(range(16)))
    assert (e[4:] == e[:4]).all()
    H = b.beltH()
    assert H[:4].hex() == "b194bac8" and sorted(H) == list(range(256))


def test_fails_loudly_without_cuda():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(b.Bee2Error) as e:
        b.bashHash(256, b"abc")
    assert e.value.code == b.ERR_B2G_NO_DEVICE
    with pytest.raises(b.Bee2Error) as e:
        b.beltCTR(bytes(32), bytes(32), bytes(16))
    assert e.value.code == b.ERR_B2G_NO_DEVICE
    assert b.bignVerify(b.bignParamsStd(), b.OID_BELT_HASH_DER, bytes(32), bytes(48), bytes(64)) == b.ERR_B2G_NO_DEVICE


_NO_DEVICE_OVERLAY = r"""
import ctypes as C, os, sys
root = sys.argv[1]
# a dlopen()ing process has no link order: the stock library is named in B2G_STOCK_LIB (engine.c: b2g_stock);
# binaries that LINK both libraries / preload the engine are covered on the GPU box (test_gpu_reftests.py)
os.environ["B2G_STOCK_LIB"] = os.path.join(root, "oracle", "_ref", "libbee2ref_64.so")
eng = C.CDLL(os.path.join(root, "bee2_b200", "libbee2_b200.so"))
ref = C.CDLL(os.path.join(root, "oracle", "_ref", "libbee2ref_64.so"))
eng.b2g_has_stock.restype = C.c_int
eng.b2g_forward_count.restype = C.c_uint64
eng.beltH.restype = C.c_void_p
assert eng.b2g_has_stock() == 1
H = bytes((C.c_ubyte * 256).from_address(eng.beltH()))
# err_t entry points never degrade: no device -> ERR_B2G_NO_DEVICE, nothing forwarded
out = (C.c_ubyte * 32)()
eng.bashHash.restype = C.c_uint32
assert eng.bashHash(out, C.c_size_t(128), H, C.c_size_t(13)) == 9001
assert eng.b2g_forward_count() == 0
# a void drop-in cannot report: with stock libbee2 behind it is handed over (round 1: abort())
st = (C.c_uint64 * 24).from_buffer_copy(H[:192])
eng.bashF(st, None)
assert eng.b2g_forward_count() == 1
ref.bashF.restype = None
st2 = (C.c_uint64 * 24).from_buffer_copy(H[:192])
ref.bashF(st2, None)
assert bytes(st) == bytes(st2) and bytes(st) != H[:192]
print("OK")
"""


def test_no_device_void_dropin_goes_to_stock_and_err_t_fails_loudly():
    """CPU box: with a stock libbee2 BEHIND the engine a `void` drop-in is forwarded (and says so on stderr),
    an `err_t` entry point still returns ERR_B2G_NO_DEVICE — no silent CPU path."""
    import subprocess
    import sys
    ref = os.path.join(ROOT, "oracle", "_ref", "libbee2ref_64.so")
    if not os.path.exists(ref):
        pytest.skip("reference library not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: the no-device path cannot be exercised")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, "-c", _NO_DEVICE_OVERLAY, ROOT], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    assert "forwarding such calls to the stock libbee2" in r.stderr
