"""ctypes binding of include/bee2_b200.h (one Python function per C entry point).

Host-pointer functions take/return ``bytes`` / numpy ``uint8`` arrays; ``*_dev`` functions
take raw device addresses (``int``, e.g. ``torch.Tensor.data_ptr()``) and a CUDA stream
handle (``int``, e.g. ``torch.cuda.current_stream().cuda_stream``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

__all__ = [
    "ERR_OK", "ERR_BAD_INPUT", "ERR_OUTOFMEMORY", "ERR_NOT_IMPLEMENTED", "ERR_FILE_NOT_FOUND",
    "ERR_BAD_OID", "ERR_BAD_PARAMS", "ERR_BAD_PRIVKEY", "ERR_BAD_PUBKEY", "ERR_BAD_SIG", "ERR_BAD_MAC",
    "ERR_B2G_NO_DEVICE", "ERR_B2G_CUDA", "Bee2Error", "BignParams", "lib", "lib_path",
    "b2g_init", "b2g_init_devices", "b2g_device_count", "b2g_ipc_export", "b2g_ipc_open", "b2g_ipc_close", "b2g_last_error", "b2g_sm_count", "b2g_launch_count", "b2g_sync", "b2g_microbench",
    "bashF", "bashHash", "bashHashBatch", "bashHashBatchV", "bashHashFiles", "bashFBatch", "BashHash",
    "beltH", "beltKeyExpand2", "beltBlockEncr", "beltBlockDecr", "beltECBEncr", "beltECBDecr",
    "beltECBEncrBatch", "BeltECB", "BeltCTR", "beltCTR", "beltCTRKeystream", "beltHash", "beltHashBatch",
    "beltDWPWrap", "beltDWPUnwrap", "BeltDWP", "BeltHash", "beltDWPMac_dev", "beltCHEWrap", "beltCHEUnwrap", "beltCHE_dev",
    "bignParamsStd", "bignVerify", "bignVerifyBatch", "bignSign2", "bignSign2Batch",
    "bignPubkeyCalc", "bignPubkeyCalcBatch", "ecMulABatch", "ecAddMulABatch", "OID_BELT_HASH_DER",
    "bashHashBatch_dev", "bashFBatch_dev", "beltCTR_dev", "beltECB_dev", "beltECBEncrBatch_dev",
    "beltHashBatch_dev", "bignVerifyBatch_dev", "bignSign2Batch_dev", "bignPubkeyCalcBatch_dev",
    "ecMulABatch_dev", "bignVerifyBatchL_dev", "bignKeypairGenBatch", "bignKeypairValBatch", "bignPubkeyValBatch",
    "bignDHBatch", "bignSignBatch", "GEN_I", "bignDH", "bignPubkeyVal", "bignKeypairVal", "ERR_BAD_RNG", "ERR_BAD_SHAREDKEY", "pinned_empty", "BIGN_CURVES",
]

ERR_OK = 0
ERR_BAD_INPUT = 109
ERR_OUTOFMEMORY = 110
ERR_NOT_IMPLEMENTED = 119
ERR_FILE_NOT_FOUND = 202
ERR_BAD_OID = 301
ERR_BAD_RNG = 304
ERR_BAD_PARAMS = 502
ERR_BAD_PRIVKEY = 504
ERR_BAD_PUBKEY = 505
ERR_BAD_SHAREDKEY = 507
ERR_BAD_SIG = 510
ERR_BAD_MAC = 511
ERR_B2G_NO_DEVICE = 9001
ERR_B2G_CUDA = 9002

# DER of OID 1.2.112.0.2.0.34.101.31.81 (belt-hash), the hash OID used by the reference's
# bign tests (test/crypto/bign_test.c:296-300) and by bign128 (bign128.c:151-153)
OID_BELT_HASH_DER = bytes.fromhex("06092A7000020022651F51")


# names of the standard parameter blocks by level l (bign_params.c:197-236)
BIGN_CURVES = {128: "1.2.112.0.2.0.34.101.45.3.1", 192: "1.2.112.0.2.0.34.101.45.3.2", 256: "1.2.112.0.2.0.34.101.45.3.3"}


class Bee2Error(RuntimeError):
    def __init__(self, fn: str, code: int):
        self.code = code
        super().__init__(f"{fn} -> err_t {code} ({b2g_last_error()})")


class BignParams(C.Structure):
    """include/bee2/crypto/bign.h:65-74"""
    _fields_ = [("l", C.c_size_t), ("p", C.c_ubyte * 64), ("a", C.c_ubyte * 64), ("b", C.c_ubyte * 64),
                ("q", C.c_ubyte * 64), ("yG", C.c_ubyte * 64), ("seed", C.c_ubyte * 8)]


_LIB: Optional[C.CDLL] = None


def lib_path() -> str:
    # BEE2_B200_LIB: an alternative build of the same library (kernel-variant experiments)
    return os.environ.get("BEE2_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbee2_b200.so")


def lib() -> C.CDLL:
    """Load libbee2_b200.so (built by ``__graft_entry__.build()`` / ``make -C bee2_b200/csrc``)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                               " — bee2_b200 has no CPU fallback")
        L = C.CDLL(path)
        _declare(L)
        _LIB = L
    return _LIB


def _declare(L: C.CDLL) -> None:
    sz, vp, u32, u64, ci = C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "b2g_init": (u32, [ci]), "b2g_last_error": (C.c_char_p, []), "b2g_sm_count": (ci, []),
        "b2g_host_alloc": (vp, [sz]), "b2g_host_free": (None, [vp]), "b2g_dev_alloc": (vp, [sz]),
        "b2g_dev_free": (None, [vp]), "b2g_memcpy_h2d": (u32, [vp, vp, sz]), "b2g_memcpy_d2h": (u32, [vp, vp, sz]),
        "b2g_sync": (u32, []), "b2g_launch_count": (u64, []), "b2g_microbench": (C.c_double, [ci, C.c_uint]),
        "b2g_init_devices": (u32, [ci]), "b2g_device_count": (ci, []),
        "b2g_ipc_export": (u32, [vp, vp]), "b2g_ipc_open": (u32, [vp, vp]), "b2g_ipc_close": (u32, [vp]),
        "b2g_memcpy_async": (u32, [vp, vp, sz, ci, vp]),
        "bashF_deep": (sz, []), "bashF": (None, [vp, vp]), "bashHash_keep": (sz, []),
        "bashHashStart": (None, [vp, sz]), "bashHashStepH": (None, [vp, sz, vp]),
        "bashHashStepG": (None, [vp, sz, vp]), "bashHashStepV": (ci, [vp, sz, vp]),
        "bashHash": (u32, [vp, sz, vp, sz]), "bashHashBatch": (u32, [vp, sz, vp, sz, sz, sz]),
        "bashPrg_keep": (sz, []), "bashHashFiles": (u32, [vp, vp, sz, vp, sz]),
        "bashFBatch": (u32, [vp, sz]), "bashHashBatchV": (u32, [vp, sz, vp, sz, vp, vp, sz]),
        "b2g_bashHashBatchV_dev": (u32, [vp, sz, vp, vp, vp, sz, vp]),
        "b2g_bashHashBatch_dev": (u32, [vp, sz, vp, sz, sz, sz, vp]), "b2g_bashFBatch_dev": (u32, [vp, sz, vp]),
        "beltH": (vp, []), "beltKeyExpand": (None, [vp, vp, sz]), "beltKeyExpand2": (None, [vp, vp, sz]),
        "beltBlockEncr": (None, [vp, vp]), "beltBlockEncr2": (None, [vp, vp]),
        "beltBlockEncr3": (None, [vp, vp, vp, vp, vp]), "beltBlockDecr": (None, [vp, vp]),
        "beltBlockDecr2": (None, [vp, vp]), "beltBlockDecr3": (None, [vp, vp, vp, vp, vp]),
        "beltECB_keep": (sz, []), "beltECBStart": (None, [vp, vp, sz]), "beltECBStepE": (None, [vp, sz, vp]),
        "beltECBStepD": (None, [vp, sz, vp]), "beltECBEncr": (u32, [vp, vp, sz, vp, sz]),
        "beltECBDecr": (u32, [vp, vp, sz, vp, sz]),
        "beltCTR_keep": (sz, []), "beltCTRStart": (None, [vp, vp, sz, vp]), "beltCTRStepE": (None, [vp, sz, vp]),
        "beltCTR": (u32, [vp, vp, sz, vp, sz, vp]), "beltHash": (u32, [vp, vp, sz]),
        "beltCTRKeystream": (u32, [vp, sz, vp, sz, vp]), "beltECBEncrBatch": (u32, [vp, vp, sz]),
        "beltHashBatch": (u32, [vp, vp, sz, sz, sz]),
        "beltDWPWrap": (u32, [vp, vp, vp, sz, vp, sz, vp, sz, vp]),
        "beltHash_keep": (sz, []), "beltHashStart": (None, [vp]), "beltHashStepH": (None, [vp, sz, vp]),
        "beltHashStepG": (None, [vp, vp]), "beltHashStepG2": (None, [vp, sz, vp]),
        "beltHashStepV": (ci, [vp, vp]), "beltHashStepV2": (ci, [vp, sz, vp]),
        **{f"belt{m}_keep": (sz, []) for m in ("DWP", "CHE")},
        **{f"belt{m}Start": (None, [vp, vp, sz, vp]) for m in ("DWP", "CHE")},
        **{f"belt{m}Step{x}": (None, [vp, sz, vp]) for m in ("DWP", "CHE") for x in "EDIA"},
        **{f"belt{m}StepG": (None, [vp, vp]) for m in ("DWP", "CHE")},
        **{f"belt{m}StepV": (ci, [vp, vp]) for m in ("DWP", "CHE")},
        "beltDWPUnwrap": (u32, [vp, vp, sz, vp, sz, vp, vp, sz, vp]),
        "beltCHEWrap": (u32, [vp, vp, vp, sz, vp, sz, vp, sz, vp]),
        "beltCHEUnwrap": (u32, [vp, vp, sz, vp, sz, vp, vp, sz, vp]),
        "b2g_beltCHE_dev": (u32, [vp, vp, sz, vp, vp, u64, vp]),
        "b2g_beltCHEMac_dev": (u32, [vp, vp, sz, vp, sz, vp, vp, vp, vp]),
        "b2g_beltDWPMac_dev": (u32, [vp, vp, sz, vp, sz, vp, vp, vp, vp]),
        "b2g_beltCTR_dev": (u32, [vp, vp, sz, vp, vp, u64, vp]), "b2g_beltECB_dev": (u32, [vp, vp, sz, vp, ci, vp]),
        "b2g_beltECBEncrBatch_dev": (u32, [vp, vp, sz, vp]), "b2g_beltECBEncrBatch2_dev": (u32, [vp, vp, vp, sz, vp]), "b2g_beltHashBatch_dev": (u32, [vp, vp, sz, sz, sz, vp]),
        "bignParamsStd": (u32, [vp, C.c_char_p]), "bignVerify": (u32, [vp, vp, sz, vp, vp, vp]),
        "bignSign2": (u32, [vp, vp, vp, sz, vp, vp, vp, sz]), "bignPubkeyCalc": (u32, [vp, vp, vp]),
        "bignVerifyBatch": (u32, [vp, vp, vp, sz, vp, vp, vp, sz]),
        "bignSign2Batch": (u32, [vp, vp, vp, vp, sz, vp, vp, sz]),
        "bignPubkeyCalcBatch": (u32, [vp, vp, vp, vp, sz]), "ecMulABatch": (u32, [vp, vp, vp, vp, sz, sz]),
        "b2g_bignVerifyBatch_dev": (u32, [vp, vp, sz, vp, vp, vp, sz, vp]),
        "b2g_bignSign2Batch_dev": (u32, [vp, vp, vp, sz, vp, vp, sz, vp]),
        "b2g_bignPubkeyCalcBatch_dev": (u32, [vp, vp, vp, sz, vp]),
        "b2g_ecMulABatch_dev": (u32, [vp, vp, vp, vp, sz, sz, vp]),
        "ecAddMulABatch": (u32, [vp, vp, vp, vp, sz, vp, sz]),
        "b2g_ecAddMulABatch_dev": (u32, [vp, vp, vp, vp, sz, vp, sz, vp]),
        "bignKeypairGen": (u32, [vp, vp, vp, vp, vp]), "bignKeypairGenBatch": (u32, [vp, vp, vp, vp, vp, sz]),
        "bignKeypairVal": (u32, [vp, vp, vp]), "bignKeypairValBatch": (u32, [vp, vp, vp, vp, sz]),
        "bignPubkeyVal": (u32, [vp, vp]), "bignPubkeyValBatch": (u32, [vp, vp, vp, sz]),
        "bignSign": (u32, [vp, vp, vp, sz, vp, vp, vp, vp]), "bignSignBatch": (u32, [vp, vp, vp, vp, sz, vp, vp, vp, vp, sz]),
        **{f"bign{lv}{fn}": sig_ for lv in (128, 192, 256) for fn, sig_ in (
            ("KeypairGen", (u32, [vp, vp, vp, vp])), ("KeypairVal", (u32, [vp, vp])), ("PubkeyVal", (u32, [vp])),
            ("PubkeyCalc", (u32, [vp, vp])), ("DH", (u32, [vp, vp, vp, sz])), ("Sign", (u32, [vp, vp, vp, vp, vp])),
            ("Sign2", (u32, [vp, vp, vp, vp, sz])), ("Verify", (u32, [vp, vp, vp])))},
        "bignDH": (u32, [vp, vp, vp, vp, sz]), "bignDHBatch": (u32, [vp, vp, vp, vp, vp, sz, sz]),
        "b2g_bignDHBatchL_dev": (u32, [sz, vp, vp, vp, vp, sz, vp]),
        "b2g_bignPubkeyValBatchL_dev": (u32, [sz, vp, vp, sz, vp]),
        "ecMulABatchL": (u32, [sz, vp, vp, vp, vp, sz, sz]), "ecAddMulABatchL": (u32, [sz, vp, vp, vp, vp, sz, vp, sz]),
        "b2g_bignVerifyBatchL_dev": (u32, [sz, vp, vp, sz, vp, vp, vp, sz, vp]),
        "b2g_bignSign2BatchL_t_dev": (u32, [sz, vp, vp, vp, sz, vp, vp, sz, vp, sz, vp]),
        "b2g_bignPubkeyCalcBatchL_dev": (u32, [sz, vp, vp, vp, sz, vp]),
        "b2g_ecMulABatchL_dev": (u32, [sz, vp, vp, vp, vp, sz, sz, vp]),
        "b2g_ecAddMulABatchL_dev": (u32, [sz, vp, vp, vp, vp, sz, vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args


# ------------------------------------------------------------------ helpers
def _buf(data) -> Tuple[object, int, int]:
    """(keepalive, address, nbytes) for bytes / bytearray / numpy array (None -> NULL)."""
    if data is None:
        return None, 0, 0
    if isinstance(data, np.ndarray):
        if not data.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return data, data.ctypes.data, data.nbytes
    if isinstance(data, (bytes, bytearray, memoryview)):
        a = np.frombuffer(data, dtype=np.uint8)
        return (a, data), a.ctypes.data if a.size else 0, a.size
    raise TypeError(type(data))


def _out(n: int) -> np.ndarray:
    return np.zeros(max(n, 0), dtype=np.uint8)


def _chk(fn: str, code: int) -> None:
    if code != ERR_OK:
        raise Bee2Error(fn, code)


def pinned_empty(nbytes: int) -> np.ndarray:
    """uint8 array over page-locked host memory (b2g_host_alloc), for full-rate PCIe copies.
    The allocation lives until the process exits."""
    p = lib().b2g_host_alloc(nbytes)
    if not p:
        raise Bee2Error("b2g_host_alloc", ERR_OUTOFMEMORY)
    _PINNED.append(p)
    return np.ctypeslib.as_array((C.c_ubyte * max(nbytes, 1)).from_address(p))[:nbytes]


_PINNED: list = []


# ------------------------------------------------------------------ engine
def b2g_init(device: int = -1) -> int:
    return lib().b2g_init(device)


def b2g_init_devices(n: int = 0) -> int:
    """In-process multi-device mode: the *Batch calls shard over devices 0..n-1 (n <= 0: all)."""
    return lib().b2g_init_devices(n)


def b2g_device_count() -> int:
    return lib().b2g_device_count()


def b2g_ipc_export(dptr: int) -> bytes:
    """CUDA IPC handle (64 octets) of a b2g_dev_alloc'ed buffer, for a peer rank's b2g_ipc_open."""
    h = (C.c_ubyte * 64)()
    code = lib().b2g_ipc_export(h, dptr)
    if code:
        raise Bee2Error("b2g_ipc_export", code)
    return bytes(h)


def b2g_ipc_open(handle: bytes) -> int:
    """Map a peer rank's buffer into this process; returns the device pointer."""
    p = C.c_void_p()
    hb = (C.c_ubyte * 64).from_buffer_copy(handle)
    code = lib().b2g_ipc_open(C.byref(p), hb)
    if code:
        raise Bee2Error("b2g_ipc_open", code)
    return int(p.value)


def b2g_ipc_close(dptr: int) -> None:
    lib().b2g_ipc_close(dptr)


def b2g_last_error() -> str:
    try:
        return (lib().b2g_last_error() or b"").decode(errors="replace")
    except Exception:  # library not loadable
        return ""


def b2g_sm_count() -> int:
    return lib().b2g_sm_count()


def b2g_launch_count() -> int:
    return int(lib().b2g_launch_count())


def b2g_microbench(kind: int, iters: int = 0) -> float:
    """lane-ops/s of one instruction kind (see include/bee2_b200.h)"""
    return float(lib().b2g_microbench(kind, iters))


def b2g_sync() -> None:
    _chk("b2g_sync", lib().b2g_sync())


# ------------------------------------------------------------------ bash
def bashF(block: bytes) -> bytes:
    """bash.h:133-139"""
    a = np.frombuffer(bytes(block), dtype=np.uint8).copy()
    assert a.size == 192
    lib().bashF(a.ctypes.data, None)
    return a.tobytes()


def bashHash(l: int, src: bytes) -> bytes:
    """bash.h:218-225; raises Bee2Error with the reference's err_t on bad arguments."""
    k, p, n = _buf(src)
    out = _out(l // 4 if 0 < l <= 256 else 64)
    _chk("bashHash", lib().bashHash(out.ctypes.data, l, p, n))
    return out[: l // 4].tobytes()


def bashHashBatch(l: int, msgs: np.ndarray, msg_len: Optional[int] = None, stride: Optional[int] = None,
                  count: Optional[int] = None) -> np.ndarray:
    """msgs: uint8 array [count, stride] (or flat with explicit msg_len/stride/count) -> [count, l/4]."""
    if msgs.ndim == 2 and msg_len is None:
        count, stride = msgs.shape
        msg_len = stride
    assert msg_len is not None and stride is not None and count is not None
    k, p, n = _buf(msgs)
    out = np.zeros((count, l // 4), dtype=np.uint8)
    _chk("bashHashBatch", lib().bashHashBatch(out.ctypes.data, l, p, msg_len, stride, count))
    return out


def bashHashBatchV(l: int, data: np.ndarray, offsets: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """ragged batch: message i = data[offsets[i]:offsets[i]+lens[i]] -> digests [count, l/4]"""
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint64)
    out = np.zeros((off.size, max(l // 4, 1)), dtype=np.uint8)
    k, p, n = _buf(data)
    _chk("bashHashBatchV", lib().bashHashBatchV(out.ctypes.data, l, p, n, off.ctypes.data, ln.ctypes.data, off.size))
    return out


def bashFBatch(blocks: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(blocks, dtype=np.uint8).copy()
    _chk("bashFBatch", lib().bashFBatch(a.ctypes.data, a.size // 192))
    return a


class BashHash:
    """bashHashStart/StepH/StepG/StepV over a caller-owned state blob (bash.h:152-216)."""

    def __init__(self, l: int):
        self.L = lib()
        self.state = np.zeros(self.L.bashHash_keep(), dtype=np.uint8)
        self.L.bashHashStart(self.state.ctypes.data, l)

    def step_h(self, data: bytes) -> None:
        k, p, n = _buf(data)
        self.L.bashHashStepH(p, n, self.state.ctypes.data)

    def step_g(self, hash_len: int) -> bytes:
        out = _out(hash_len)
        self.L.bashHashStepG(out.ctypes.data, hash_len, self.state.ctypes.data)
        return out.tobytes()

    def step_v(self, h: bytes) -> bool:
        k, p, n = _buf(h)
        return bool(self.L.bashHashStepV(p, n, self.state.ctypes.data))

    def copy(self) -> "BashHash":
        o = object.__new__(BashHash)
        o.L, o.state = self.L, self.state.copy()   # states are memcpy-able (belt.h:90-96)
        return o


# ------------------------------------------------------------------ belt
def beltH() -> bytes:
    p = lib().beltH()
    return bytes((C.c_ubyte * 256).from_address(p))


def beltKeyExpand2(key: bytes) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint32)
    k, p, n = _buf(key)
    lib().beltKeyExpand2(out.ctypes.data, p, n)
    return out


def beltBlockEncr(block: bytes, key: bytes, flavour: int = 1) -> bytes:
    """flavour 1/2/3 = beltBlockEncr / Encr2 / Encr3 (belt.h:193-225)"""
    ke = beltKeyExpand2(key)
    a = np.frombuffer(bytes(block), dtype=np.uint8).copy()
    L = lib()
    if flavour == 1:
        L.beltBlockEncr(a.ctypes.data, ke.ctypes.data)
    elif flavour == 2:
        L.beltBlockEncr2(a.ctypes.data, ke.ctypes.data)
    else:
        w = a.view(np.uint32)
        L.beltBlockEncr3(w[0:].ctypes.data, w[1:].ctypes.data, w[2:].ctypes.data, w[3:].ctypes.data, ke.ctypes.data)
    return a.tobytes()


def beltBlockDecr(block: bytes, key: bytes, flavour: int = 1) -> bytes:
    ke = beltKeyExpand2(key)
    a = np.frombuffer(bytes(block), dtype=np.uint8).copy()
    L = lib()
    if flavour == 1:
        L.beltBlockDecr(a.ctypes.data, ke.ctypes.data)
    elif flavour == 2:
        L.beltBlockDecr2(a.ctypes.data, ke.ctypes.data)
    else:
        w = a.view(np.uint32)
        L.beltBlockDecr3(w[0:].ctypes.data, w[1:].ctypes.data, w[2:].ctypes.data, w[3:].ctypes.data, ke.ctypes.data)
    return a.tobytes()


def _ecb(fn: str, src: bytes, key: bytes) -> bytes:
    k1, p, n = _buf(src)
    k2, kp, kn = _buf(key)
    out = _out(n)
    _chk(fn, getattr(lib(), fn)(out.ctypes.data, p, n, kp, kn))
    return out.tobytes()


def beltECBEncr(src: bytes, key: bytes) -> bytes:
    return _ecb("beltECBEncr", src, key)


def beltECBDecr(src: bytes, key: bytes) -> bytes:
    return _ecb("beltECBDecr", src, key)


def beltECBEncrBatch(blocks: np.ndarray, keys32: np.ndarray) -> np.ndarray:
    """blocks [count,16] in, keys [count,32] -> encrypted blocks [count,16] (key agility, config 5)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).copy()
    k = np.ascontiguousarray(keys32, dtype=np.uint8)
    _chk("beltECBEncrBatch", lib().beltECBEncrBatch(b.ctypes.data, k.ctypes.data, b.size // 16))
    return b


class BeltECB:
    """beltECBStart/StepE/StepD (belt.h:411-455)"""

    def __init__(self, key: bytes):
        self.L = lib()
        self.state = np.zeros(self.L.beltECB_keep(), dtype=np.uint8)
        k, p, n = _buf(key)
        self.L.beltECBStart(self.state.ctypes.data, p, n)

    def step_e(self, buf: np.ndarray) -> None:
        self.L.beltECBStepE(buf.ctypes.data, buf.size, self.state.ctypes.data)

    def step_d(self, buf: np.ndarray) -> None:
        self.L.beltECBStepD(buf.ctypes.data, buf.size, self.state.ctypes.data)


class BeltCTR:
    """beltCTRStart/StepE(=StepD) (belt.h:702-724); buffers are transformed in place."""

    def __init__(self, key: bytes, iv: bytes):
        self.L = lib()
        self.state = np.zeros(self.L.beltCTR_keep(), dtype=np.uint8)
        k, p, n = _buf(key)
        k2, ivp, _ = _buf(iv)
        self.L.beltCTRStart(self.state.ctypes.data, p, n, ivp)

    def step_e(self, buf: np.ndarray) -> None:
        self.L.beltCTRStepE(buf.ctypes.data, buf.size, self.state.ctypes.data)

    step_d = step_e

    @property
    def key_words(self) -> np.ndarray:
        return self.state[:32].view(np.uint32).copy()

    @property
    def ctr_words(self) -> np.ndarray:
        return self.state[32:48].view(np.uint32).copy()


class BeltHash:
    """beltHashStart/StepH/StepG/StepG2/StepV/StepV2 (belt.h, belt_hash.c:27-158)."""

    def __init__(self):
        self.L = lib()
        self.state = np.zeros(self.L.beltHash_keep(), dtype=np.uint8)
        self.L.beltHashStart(self.state.ctypes.data)

    def step_h(self, data: bytes) -> None:
        k, p, n = _buf(data)
        self.L.beltHashStepH(p, n, self.state.ctypes.data)

    def step_g(self, hash_len: int = 32) -> bytes:
        out = _out(32)
        if hash_len == 32:
            self.L.beltHashStepG(out.ctypes.data, self.state.ctypes.data)
        else:
            self.L.beltHashStepG2(out.ctypes.data, hash_len, self.state.ctypes.data)
        return out.tobytes()[:hash_len]

    def step_v(self, digest: bytes) -> bool:
        k, p, n = _buf(digest)
        if n == 32:
            return bool(self.L.beltHashStepV(p, self.state.ctypes.data))
        return bool(self.L.beltHashStepV2(p, n, self.state.ctypes.data))


class BeltDWP:
    """beltDWPStart/StepE/StepI/StepA/StepD/StepG/StepV (belt.h:850-983); ``mode="CHE"`` binds the
    belt-CHE twins. Step methods take bytes and return bytes where data are transformed."""

    def __init__(self, key: bytes, iv: bytes, mode: str = "DWP"):
        self.L, self.pre = lib(), f"belt{mode}"
        self.state = np.zeros(getattr(self.L, self.pre + "_keep")(), dtype=np.uint8)
        k, p, n = _buf(key)
        k2, ivp, _ = _buf(iv)
        getattr(self.L, self.pre + "Start")(self.state.ctypes.data, p, n, ivp)

    def _xform(self, which: str, data: bytes) -> bytes:
        buf = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        getattr(self.L, self.pre + which)(buf.ctypes.data if buf.size else 0, buf.size, self.state.ctypes.data)
        return buf.tobytes()

    def step_e(self, data: bytes) -> bytes:
        return self._xform("StepE", data)

    def step_d(self, data: bytes) -> bytes:
        return self._xform("StepD", data)

    def step_i(self, data: bytes) -> None:
        k, p, n = _buf(data)
        getattr(self.L, self.pre + "StepI")(p, n, self.state.ctypes.data)

    def step_a(self, data: bytes) -> None:
        k, p, n = _buf(data)
        getattr(self.L, self.pre + "StepA")(p, n, self.state.ctypes.data)

    def step_g(self) -> bytes:
        mac = _out(8)
        getattr(self.L, self.pre + "StepG")(mac.ctypes.data, self.state.ctypes.data)
        return mac.tobytes()

    def step_v(self, mac: bytes) -> bool:
        k, p, n = _buf(mac)
        return bool(getattr(self.L, self.pre + "StepV")(p, self.state.ctypes.data))


def beltCTR(src, key: bytes, iv: bytes, out: Optional[np.ndarray] = None):
    """belt.h:736-743. ``src``/``out`` may be numpy arrays (e.g. pinned) to avoid copies."""
    k1, p, n = _buf(src)
    k2, kp, kn = _buf(key)
    k3, ivp, _ = _buf(iv)
    o = out if out is not None else _out(n)
    _chk("beltCTR", lib().beltCTR(o.ctypes.data, p, n, kp, kn, ivp))
    return o if out is not None else o.tobytes()


def beltCTRKeystream(count: int, key: bytes, iv: bytes, out: Optional[np.ndarray] = None):
    k2, kp, kn = _buf(key)
    k3, ivp, _ = _buf(iv)
    o = out if out is not None else _out(count)
    _chk("beltCTRKeystream", lib().beltCTRKeystream(o.ctypes.data, count, kp, kn, ivp))
    return o if out is not None else o.tobytes()


def beltCHEWrap(src1, src2, key: bytes, iv: bytes):
    """belt-CHE wrap (belt_che.c:257-285) — returns (ciphertext bytes, mac[8])"""
    return beltDWPWrap(src1, src2, key, iv, fn="beltCHEWrap")


def beltCHEUnwrap(src1, src2, mac: bytes, key: bytes, iv: bytes):
    return beltDWPUnwrap(src1, src2, mac, key, iv, fn="beltCHEUnwrap")


def beltCHE_dev(d_dest: int, d_src: int, count: int, key_words: np.ndarray, s0_words: np.ndarray,
                first_block: int = 0, stream: int = 0) -> None:
    _chk("b2g_beltCHE_dev", lib().b2g_beltCHE_dev(d_dest, d_src, count, key_words.ctypes.data, s0_words.ctypes.data,
                                                  first_block, stream))


def beltDWPWrap(src1, src2, key: bytes, iv: bytes, fn: str = "beltDWPWrap"):
    """belt.h:984-1007 — returns (ciphertext bytes, mac[8])"""
    k1, p1, n1 = _buf(src1)
    k2, p2, n2 = _buf(src2)
    kk, kp, kn = _buf(key)
    ki, ivp, _ = _buf(iv)
    out, mac = _out(n1), _out(8)
    _chk(fn, getattr(lib(), fn)(out.ctypes.data, mac.ctypes.data, p1, n1, p2, n2, kp, kn, ivp))
    return out.tobytes(), mac.tobytes()


def beltDWPUnwrap(src1, src2, mac: bytes, key: bytes, iv: bytes, fn: str = "beltDWPUnwrap"):
    """belt.h:1009-1030 — returns (err_t, plaintext bytes or None)"""
    k1, p1, n1 = _buf(src1)
    k2, p2, n2 = _buf(src2)
    km, mp, _ = _buf(mac)
    kk, kp, kn = _buf(key)
    ki, ivp, _ = _buf(iv)
    out = _out(n1)
    code = getattr(lib(), fn)(out.ctypes.data, p1, n1, p2, n2, mp, kp, kn, ivp)
    return code, (out.tobytes() if code == ERR_OK else None)


def beltDWPMac_dev(d_mac: int, d_crit: int, n1: int, d_open: int, n2: int, key_words: np.ndarray,
                   ctr_words: np.ndarray, d_scratch: int, stream: int = 0) -> None:
    _chk("b2g_beltDWPMac_dev", lib().b2g_beltDWPMac_dev(d_mac, d_crit, n1, d_open, n2, key_words.ctypes.data,
                                                        ctr_words.ctypes.data, d_scratch, stream))


def beltHash(src: bytes) -> bytes:
    k, p, n = _buf(src)
    out = _out(32)
    _chk("beltHash", lib().beltHash(out.ctypes.data, p, n))
    return out.tobytes()


def beltHashBatch(msgs: np.ndarray) -> np.ndarray:
    count, stride = msgs.shape
    k, p, n = _buf(msgs)
    out = np.zeros((count, 32), dtype=np.uint8)
    _chk("beltHashBatch", lib().beltHashBatch(out.ctypes.data, p, stride, stride, count))
    return out


# ------------------------------------------------------------------ bign
def bignParamsStd(name: str = "1.2.112.0.2.0.34.101.45.3.1") -> BignParams:
    p = BignParams()
    _chk("bignParamsStd", lib().bignParamsStd(C.addressof(p), name.encode()))
    return p


def bignVerify(params: BignParams, oid_der: bytes, hash_: bytes, sig: bytes, pubkey: bytes) -> int:
    """bign.h:392-402 — returns the err_t (ERR_OK / ERR_BAD_SIG / ERR_BAD_PUBKEY / ...)."""
    ks = [_buf(x) for x in (oid_der, hash_, sig, pubkey)]
    return lib().bignVerify(C.addressof(params), ks[0][1], ks[0][2], ks[1][1], ks[2][1], ks[3][1])


def bignVerifyBatch(params: BignParams, oid_der: bytes, hashes: np.ndarray, sigs: np.ndarray,
                    pubkeys: np.ndarray, status: Optional[np.ndarray] = None) -> np.ndarray:
    """``status`` may be a caller-provided uint32 array (e.g. pinned) to avoid a pageable download."""
    count = hashes.size // (params.l // 4)
    if status is None:
        status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    ko = _buf(oid_der)
    _chk("bignVerifyBatch", lib().bignVerifyBatch(status.ctypes.data, C.addressof(params), ko[1], ko[2],
                                                  hashes.ctypes.data, sigs.ctypes.data, pubkeys.ctypes.data, count))
    return status


def bignSign2(params: BignParams, oid_der: bytes, hash_: bytes, privkey: bytes, t: Optional[bytes] = None) -> bytes:
    ks = [_buf(x) for x in (oid_der, hash_, privkey, t)]
    sig = _out(3 * params.l // 8)
    _chk("bignSign2", lib().bignSign2(sig.ctypes.data, C.addressof(params), ks[0][1], ks[0][2], ks[1][1], ks[2][1],
                                      ks[3][1], ks[3][2]))
    return sig.tobytes()


def bignSign2Batch(params: BignParams, oid_der: bytes, hashes: np.ndarray, privkeys: np.ndarray,
                   status: Optional[np.ndarray] = None, sigs: Optional[np.ndarray] = None):
    count = hashes.size // (params.l // 4)
    if status is None:
        status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    if sigs is None:
        sigs = np.zeros((count, 3 * params.l // 8), dtype=np.uint8)
    ko = _buf(oid_der)
    _chk("bignSign2Batch", lib().bignSign2Batch(status.ctypes.data, sigs.ctypes.data, C.addressof(params), ko[1], ko[2],
                                                hashes.ctypes.data, privkeys.ctypes.data, count))
    return status, sigs


def bignPubkeyCalc(params: BignParams, privkey: bytes) -> bytes:
    k = _buf(privkey)
    out = _out(params.l // 2)
    _chk("bignPubkeyCalc", lib().bignPubkeyCalc(out.ctypes.data, C.addressof(params), k[1]))
    return out.tobytes()


def bignPubkeyCalcBatch(params: BignParams, privkeys: np.ndarray):
    count = privkeys.size // (params.l // 4)
    status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    pub = np.zeros((count, params.l // 2), dtype=np.uint8)
    _chk("bignPubkeyCalcBatch", lib().bignPubkeyCalcBatch(status.ctypes.data, pub.ctypes.data, C.addressof(params),
                                                          privkeys.ctypes.data, count))
    return status, pub


GEN_I = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p)


def _rng_callback(stream: bytes):
    """A gen_i (defs.h:520-524) that hands out the octets of `stream` in order."""
    pos = [0]

    def fn(buf, count, state):
        chunk = stream[pos[0]:pos[0] + count]
        assert len(chunk) == count, "rng stream exhausted"
        C.memmove(buf, chunk, count)
        pos[0] += count
    return GEN_I(fn), pos


def bignKeypairGenBatch(params: BignParams, rng_stream: bytes, count: int):
    """`count` key pairs from the octets of rng_stream (consumed as the reference's generator calls
    would consume them); returns (privkeys [count,l/4], pubkeys [count,l/2], octets consumed)."""
    no = params.l // 4
    priv = np.zeros((count, no), dtype=np.uint8)
    pub = np.zeros((count, 2 * no), dtype=np.uint8)
    cb, pos = _rng_callback(bytes(rng_stream))
    _chk("bignKeypairGenBatch", lib().bignKeypairGenBatch(priv.ctypes.data, pub.ctypes.data, C.addressof(params),
                                                          C.cast(cb, C.c_void_p), None, count))
    return priv, pub, pos[0]


def bignSignBatch(params: BignParams, oid_der: bytes, hashes: np.ndarray, privkeys: np.ndarray, rng_stream: bytes):
    """bignSign on a batch with the generator octets taken from rng_stream; returns (status, sigs, consumed)."""
    no = params.l // 4
    count = hashes.size // no
    status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    sigs = np.zeros((count, no + no // 2), dtype=np.uint8)
    ko = _buf(oid_der)
    cb, pos = _rng_callback(bytes(rng_stream))
    _chk("bignSignBatch", lib().bignSignBatch(status.ctypes.data, sigs.ctypes.data, C.addressof(params), ko[1], ko[2],
                                              hashes.ctypes.data, privkeys.ctypes.data, C.cast(cb, C.c_void_p), None, count))
    return status, sigs, pos[0]


def bignKeypairValBatch(params: BignParams, privkeys: np.ndarray, pubkeys: np.ndarray) -> np.ndarray:
    count = privkeys.size // (params.l // 4)
    status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    _chk("bignKeypairValBatch", lib().bignKeypairValBatch(status.ctypes.data, C.addressof(params), privkeys.ctypes.data,
                                                          pubkeys.ctypes.data, count))
    return status


def bignPubkeyValBatch(params: BignParams, pubkeys: np.ndarray) -> np.ndarray:
    count = pubkeys.size // (params.l // 2)
    status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    _chk("bignPubkeyValBatch", lib().bignPubkeyValBatch(status.ctypes.data, C.addressof(params), pubkeys.ctypes.data, count))
    return status


def bignDHBatch(params: BignParams, privkeys: np.ndarray, pubkeys: np.ndarray, key_len: int):
    count = privkeys.size // (params.l // 4)
    status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    keys = np.zeros((count, key_len), dtype=np.uint8)
    _chk("bignDHBatch", lib().bignDHBatch(status.ctypes.data, keys.ctypes.data, C.addressof(params), privkeys.ctypes.data,
                                          pubkeys.ctypes.data, key_len, count))
    return status, keys


def bignDH(params: BignParams, privkey: bytes, pubkey: bytes, key_len: int):
    """bign.h bignDH — returns (err_t, key)."""
    ks = [_buf(x) for x in (privkey, pubkey)]
    key = _out(max(key_len, 1))
    code = lib().bignDH(key.ctypes.data, C.addressof(params), ks[0][1], ks[1][1], key_len)
    return code, key.tobytes()[:key_len]


def bignPubkeyVal(params: BignParams, pubkey: bytes) -> int:
    k = _buf(pubkey)
    return lib().bignPubkeyVal(C.addressof(params), k[1])


def bignKeypairVal(params: BignParams, privkey: bytes, pubkey: bytes) -> int:
    ks = [_buf(x) for x in (privkey, pubkey)]
    return lib().bignKeypairVal(C.addressof(params), ks[0][1], ks[1][1])


def ecMulABatch(points: np.ndarray, scalars: np.ndarray, l: int = 128):
    """points [count,l/2], scalars [count,d_len] -> (results [count,l/2], ok [count]) on the standard curve of level l"""
    count, d_len = scalars.shape
    out = np.zeros((count, l // 2), dtype=np.uint8)
    ok = np.zeros(count, dtype=np.int32)
    _chk("ecMulABatchL", lib().ecMulABatchL(l, out.ctypes.data, ok.ctypes.data, points.ctypes.data, scalars.ctypes.data,
                                            d_len, count))
    return out, ok


def ecAddMulABatch(points: np.ndarray, scalars: np.ndarray, kbase: np.ndarray, l: int = 128):
    """d_i * A_i + k_i * G: points [count,l/2], scalars [count,d_len], kbase [count,l/4]"""
    count, d_len = scalars.shape
    out = np.zeros((count, l // 2), dtype=np.uint8)
    ok = np.zeros(count, dtype=np.int32)
    _chk("ecAddMulABatchL", lib().ecAddMulABatchL(l, out.ctypes.data, ok.ctypes.data, points.ctypes.data,
                                                  scalars.ctypes.data, d_len, kbase.ctypes.data, count))
    return out, ok


# ------------------------------------------------------------------ device-pointer level
def bashHashBatch_dev(d_hashes: int, l: int, d_msgs: int, msg_len: int, stride: int, count: int, stream: int = 0) -> None:
    _chk("b2g_bashHashBatch_dev", lib().b2g_bashHashBatch_dev(d_hashes, l, d_msgs, msg_len, stride, count, stream))


def bashFBatch_dev(d_blocks: int, count: int, stream: int = 0) -> None:
    _chk("b2g_bashFBatch_dev", lib().b2g_bashFBatch_dev(d_blocks, count, stream))


def beltCTR_dev(d_dest: int, d_src: int, count: int, key_words: np.ndarray, ctr_words: np.ndarray,
                first_block: int = 0, stream: int = 0) -> None:
    _chk("b2g_beltCTR_dev", lib().b2g_beltCTR_dev(d_dest, d_src, count, key_words.ctypes.data, ctr_words.ctypes.data,
                                                  first_block, stream))


def beltECB_dev(d_dest: int, d_src: int, nblocks: int, key_words: np.ndarray, decrypt: bool = False, stream: int = 0) -> None:
    _chk("b2g_beltECB_dev", lib().b2g_beltECB_dev(d_dest, d_src, nblocks, key_words.ctypes.data, int(decrypt), stream))


def beltECBEncrBatch_dev(d_blocks: int, d_keys32: int, count: int, stream: int = 0) -> None:
    _chk("b2g_beltECBEncrBatch_dev", lib().b2g_beltECBEncrBatch_dev(d_blocks, d_keys32, count, stream))


def beltHashBatch_dev(d_hashes: int, d_msgs: int, msg_len: int, stride: int, count: int, stream: int = 0) -> None:
    _chk("b2g_beltHashBatch_dev", lib().b2g_beltHashBatch_dev(d_hashes, d_msgs, msg_len, stride, count, stream))


def bignVerifyBatchL_dev(l: int, d_status: int, oid_der: bytes, d_hashes: int, d_sigs: int, d_pubkeys: int, count: int,
                         stream: int = 0) -> None:
    ko = _buf(oid_der)
    _chk("b2g_bignVerifyBatchL_dev", lib().b2g_bignVerifyBatchL_dev(l, d_status, ko[1], ko[2], d_hashes, d_sigs, d_pubkeys,
                                                                  count, stream))


def bignVerifyBatch_dev(d_status: int, oid_der: bytes, d_hashes: int, d_sigs: int, d_pubkeys: int, count: int,
                        stream: int = 0) -> None:
    ko = _buf(oid_der)
    _chk("b2g_bignVerifyBatch_dev", lib().b2g_bignVerifyBatch_dev(d_status, ko[1], ko[2], d_hashes, d_sigs, d_pubkeys,
                                                                  count, stream))


def bignSign2Batch_dev(d_status: int, d_sigs: int, oid_der: bytes, d_hashes: int, d_privkeys: int, count: int,
                       stream: int = 0) -> None:
    ko = _buf(oid_der)
    _chk("b2g_bignSign2Batch_dev", lib().b2g_bignSign2Batch_dev(d_status, d_sigs, ko[1], ko[2], d_hashes, d_privkeys,
                                                                count, stream))


def bignPubkeyCalcBatch_dev(d_status: int, d_pubkeys: int, d_privkeys: int, count: int, stream: int = 0) -> None:
    _chk("b2g_bignPubkeyCalcBatch_dev", lib().b2g_bignPubkeyCalcBatch_dev(d_status, d_pubkeys, d_privkeys, count, stream))


def ecMulABatch_dev(d_b: int, d_ok: int, d_a: int, d_d: int, d_len: int, count: int, stream: int = 0) -> None:
    _chk("b2g_ecMulABatch_dev", lib().b2g_ecMulABatch_dev(d_b, d_ok, d_a, d_d, d_len, count, stream))


def bashHashFiles(l: int, paths):
    """bsum-style: digests of the files in `paths` -> (status [count] uint32, hashes [count, l/4])."""
    count = len(paths)
    arr = (C.c_char_p * max(count, 1))(*[os.fsencode(p) for p in paths])
    status = np.full(count, 0xFFFFFFFF, dtype=np.uint32)
    hashes = np.zeros((count, l // 4), dtype=np.uint8)
    _chk("bashHashFiles", lib().bashHashFiles(hashes.ctypes.data, status.ctypes.data, l, C.cast(arr, C.c_void_p), count))
    return status, hashes


# ------------------------------------------------------------------ bash-prg (bash.h, bash_prg.c)
class BashPrg:
    """bashPrgStart/Restart/Absorb/Squeeze/Encr/Decr/Ratchet over a caller-owned state blob."""
    _prefix = "bashPrg"

    def __init__(self, l: int, d: int, ann: bytes = b"", key: bytes = b"", _lib=None, _keep=None):
        self.L = _lib or lib()
        self.state = np.zeros(_keep or self.L.bashPrg_keep(), dtype=np.uint8)
        # every argument is an explicit ctypes object: these entry points are called without argtypes (the
        # oracle's orc_bashPrg* twins share this class), where a bare Python int would be cut to 32 bits
        self._call("Start", self._sp(), C.c_size_t(l), C.c_size_t(d), bytes(ann), C.c_size_t(len(ann)),
                   bytes(key), C.c_size_t(len(key)))

    def _call(self, name, *args):
        return getattr(self.L, self._prefix + name)(*args)

    def _sp(self):
        return C.c_void_p(self.state.ctypes.data)

    def restart(self, ann: bytes = b"", key: bytes = b"") -> None:
        self._call("Restart", bytes(ann), C.c_size_t(len(ann)), bytes(key), C.c_size_t(len(key)), self._sp())

    def _inout(self, name, data: bytes) -> bytes:
        buf = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        self._call(name, C.c_void_p(buf.ctypes.data), C.c_size_t(buf.size), self._sp())
        return buf.tobytes()

    def absorb_start(self): self._call("AbsorbStart", self._sp())
    def absorb_step(self, data: bytes): self._inout("AbsorbStep", data)
    def absorb(self, data: bytes): self.absorb_start(); self.absorb_step(data)
    def squeeze_start(self): self._call("SqueezeStart", self._sp())
    def squeeze_step(self, n: int) -> bytes: return self._inout("SqueezeStep", bytes(n))
    def squeeze(self, n: int) -> bytes: self.squeeze_start(); return self.squeeze_step(n)
    def encr_start(self): self._call("EncrStart", self._sp())
    def encr_step(self, data: bytes) -> bytes: return self._inout("EncrStep", data)
    def encr(self, data: bytes) -> bytes: self.encr_start(); return self.encr_step(data)
    def decr_start(self): self._call("DecrStart", self._sp())
    def decr_step(self, data: bytes) -> bytes: return self._inout("DecrStep", data)
    def decr(self, data: bytes) -> bytes: self.decr_start(); return self.decr_step(data)
    def ratchet(self): self._call("Ratchet", self._sp())

    def copy(self):
        o = object.__new__(type(self))
        o.L, o.state = self.L, self.state.copy()
        return o


__all__ += ["BashPrg"]
