"""Loaders for the CHECKERS used by the tests: oracle/_ref/libbee2oracle.so (our C restatement)
and, when present, oracle/_ref/libbee2ref_64.so (the unmodified reference built by oracle/Makefile).
Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
OID = bytes.fromhex("06092A7000020022651F51")
sz = C.c_size_t

_port = None
_ref = None


def port():
    global _port
    if _port is None:
        L = C.CDLL(os.path.join(REF_DIR, "libbee2oracle.so"))
        L.orc_beltH.restype = C.c_void_p
        for n in ("orc_beltCHEWrap", "orc_beltCHEUnwrap", "orc_beltDWPWrap", "orc_beltDWPUnwrap", "orc_bashHash", "orc_beltCTR", "orc_beltECBEncr", "orc_beltECBDecr", "orc_bignVerify128",
                  "orc_bignSign2_128", "orc_bignPubkeyCalc128", "orc_bignVerify", "orc_bignSign2", "orc_bignPubkeyCalc", "orc_bignPubkeyVal", "orc_bignDH", "orc_bignSignK"):
            getattr(L, n).restype = C.c_uint32
        L.orc_ecMulA128.restype = C.c_int
        L.orc_ecMulA.restype = C.c_int
        _port = L
    return _port


def ref():
    """The unmodified reference, or None when it was not built (no /root/reference at build time)."""
    global _ref
    if _ref is None:
        p = os.path.join(REF_DIR, "libbee2ref_64.so")
        if not os.path.exists(p):
            return None
        L = C.CDLL(p)
        L.beltH.restype = C.c_void_p
        for n in ("beltCHEWrap", "beltCHEUnwrap", "beltDWPWrap", "beltDWPUnwrap", "bashHash", "beltCTR", "beltECBEncr", "beltECBDecr", "beltHash", "bignVerify", "bignSign2",
                  "bignParamsStd", "bignPubkeyCalc"):
            getattr(L, n).restype = C.c_uint32
        _ref = L
    return _ref


def beltH() -> bytes:
    return bytes((C.c_ubyte * 256).from_address(port().orc_beltH()))


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return a


# ---- oracle (port) wrappers, bytes in / bytes out
def bashF(block: bytes) -> bytes:
    b = C.create_string_buffer(bytes(block), 192)
    port().orc_bashF(b)
    return b.raw


def bashHash(l: int, src: bytes) -> bytes:
    out = C.create_string_buffer(64)
    code = port().orc_bashHash(out, sz(l), bytes(src), sz(len(src)))
    assert code == 0, code
    return out.raw[: l // 4]


def bashHashBatch(l: int, msgs: np.ndarray) -> np.ndarray:
    out = np.zeros((msgs.shape[0], l // 4), dtype=np.uint8)
    for i in range(msgs.shape[0]):
        out[i] = np.frombuffer(bashHash(l, msgs[i].tobytes()), dtype=np.uint8)
    return out


def beltBlockEncr(block: bytes, key: bytes) -> bytes:
    k = (C.c_uint32 * 8)()
    port().orc_beltKeyExpand2(k, bytes(key), sz(len(key)))
    b = C.create_string_buffer(bytes(block), 16)
    port().orc_beltBlockEncr2(b, k)
    return b.raw


def beltBlockDecr(block: bytes, key: bytes) -> bytes:
    k = (C.c_uint32 * 8)()
    port().orc_beltKeyExpand2(k, bytes(key), sz(len(key)))
    b = C.create_string_buffer(bytes(block), 16)
    port().orc_beltBlockDecr2(b, k)
    return b.raw


def beltCTR(src: bytes, key: bytes, iv: bytes) -> bytes:
    out = C.create_string_buffer(max(len(src), 1))
    assert port().orc_beltCTR(out, bytes(src), sz(len(src)), bytes(key), sz(len(key)), bytes(iv)) == 0
    return out.raw[: len(src)]


def beltECBEncr(src: bytes, key: bytes) -> bytes:
    out = C.create_string_buffer(max(len(src), 1))
    assert port().orc_beltECBEncr(out, bytes(src), sz(len(src)), bytes(key), sz(len(key))) == 0
    return out.raw[: len(src)]


def beltECBDecr(src: bytes, key: bytes) -> bytes:
    out = C.create_string_buffer(max(len(src), 1))
    assert port().orc_beltECBDecr(out, bytes(src), sz(len(src)), bytes(key), sz(len(key))) == 0
    return out.raw[: len(src)]


def beltECBEncrMultiKey(blocks: np.ndarray, keys: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(blocks, dtype=np.uint8).copy()
    k = np.ascontiguousarray(keys, dtype=np.uint8)
    port().orc_beltECBEncrMultiKey(_p(b), _p(k), sz(b.size // 16))
    return b


def beltCHEWrap(src1: bytes, src2: bytes, key: bytes, iv: bytes):
    return beltDWPWrap(src1, src2, key, iv, fn="orc_beltCHEWrap")


def beltCHEUnwrap(src1: bytes, src2: bytes, mac: bytes, key: bytes, iv: bytes):
    return beltDWPUnwrap(src1, src2, mac, key, iv, fn="orc_beltCHEUnwrap")


def beltDWPWrap(src1: bytes, src2: bytes, key: bytes, iv: bytes, fn="orc_beltDWPWrap"):
    d, m = C.create_string_buffer(max(len(src1), 1)), C.create_string_buffer(8)
    code = getattr(port(), fn)(d, m, bytes(src1), sz(len(src1)), bytes(src2), sz(len(src2)), bytes(key), sz(len(key)), bytes(iv))
    assert code == 0
    return d.raw[: len(src1)], m.raw


def beltDWPUnwrap(src1: bytes, src2: bytes, mac: bytes, key: bytes, iv: bytes, fn="orc_beltDWPUnwrap"):
    d = C.create_string_buffer(max(len(src1), 1))
    code = getattr(port(), fn)(d, bytes(src1), sz(len(src1)), bytes(src2), sz(len(src2)), bytes(mac), bytes(key),
                                    sz(len(key)), bytes(iv))
    return code, (d.raw[: len(src1)] if code == 0 else None)


class _AeadSt(C.Structure):
    _fields_ = [("key", C.c_uint32 * 8), ("s", C.c_uint32 * 4), ("r", C.c_uint64 * 2), ("t", C.c_uint64 * 2),
                ("len", C.c_uint64 * 2), ("block", C.c_ubyte * 16), ("filled", C.c_size_t), ("ks", C.c_ubyte * 16),
                ("reserved", C.c_size_t), ("che", C.c_int)]


class BeltDWP:
    """The oracle's streaming belt-DWP / belt-CHE behind the interface of bee2_b200.BeltDWP."""

    def __init__(self, key: bytes, iv: bytes, mode: str = "DWP"):
        self.L, self.st = port(), _AeadSt()
        self.L.orc_beltAEADStart(C.byref(self.st), 1 if mode == "CHE" else 0, bytes(key), sz(len(key)), bytes(iv))

    def step_e(self, data: bytes) -> bytes:
        buf = C.create_string_buffer(bytes(data), max(len(data), 1))
        self.L.orc_beltAEADStepE(buf, sz(len(data)), C.byref(self.st))
        return buf.raw[: len(data)]

    step_d = step_e

    def step_i(self, data: bytes) -> None:
        self.L.orc_beltAEADStepI(bytes(data), sz(len(data)), C.byref(self.st))

    def step_a(self, data: bytes) -> None:
        self.L.orc_beltAEADStepA(bytes(data), sz(len(data)), C.byref(self.st))

    def step_g(self) -> bytes:
        mac = C.create_string_buffer(8)
        self.L.orc_beltAEADStepG(mac, C.byref(self.st))
        return mac.raw

    def step_v(self, mac: bytes) -> bool:
        return self.step_g() == bytes(mac)


def beltHash(src: bytes) -> bytes:
    out = C.create_string_buffer(32)
    port().orc_beltHash(out, bytes(src), sz(len(src)))
    return out.raw


# OIDs the reference's fixed-level wrappers use: belt-hash (bign128.c:151-153), bash384 (bign192.c:151-153),
# bash512 (bign256.c:151-153)
OIDS = {128: OID, 192: bytes.fromhex("06092A7000020022654D0C"), 256: bytes.fromhex("06092A7000020022654D0D")}
CURVES = {128: b"1.2.112.0.2.0.34.101.45.3.1", 192: b"1.2.112.0.2.0.34.101.45.3.2", 256: b"1.2.112.0.2.0.34.101.45.3.3"}


def bignVerify(hash_: bytes, sig: bytes, pub: bytes, oid: bytes = OID, l: int = 128) -> int:
    return port().orc_bignVerify(sz(l), oid, sz(len(oid)), bytes(hash_), bytes(sig), bytes(pub))


def bignSign2(hash_: bytes, priv: bytes, t: bytes = None, oid: bytes = OID, l: int = 128):
    sig = C.create_string_buffer(3 * l // 8)
    code = port().orc_bignSign2(sz(l), sig, oid, sz(len(oid)), bytes(hash_), bytes(priv), t, sz(len(t) if t else 0))
    return code, sig.raw


def bignSignK(hash_: bytes, priv: bytes, k: bytes, oid: bytes = OID, l: int = 128):
    sig = C.create_string_buffer(3 * l // 8)
    code = port().orc_bignSignK(sz(l), sig, oid, sz(len(oid)), bytes(hash_), bytes(priv), bytes(k))
    return code, sig.raw


def bignPubkeyCalc(priv: bytes, l: int = 128):
    pub = C.create_string_buffer(l // 2)
    code = port().orc_bignPubkeyCalc(sz(l), pub, bytes(priv))
    return code, pub.raw


def bignPubkeyVal(pub: bytes, l: int = 128) -> int:
    return port().orc_bignPubkeyVal(sz(l), bytes(pub))


def bignDH(priv: bytes, pub: bytes, key_len: int, l: int = 128):
    key = C.create_string_buffer(max(key_len, 1))
    code = port().orc_bignDH(sz(l), key, bytes(priv), bytes(pub), sz(key_len))
    return code, key.raw[:key_len] if code == 0 else b""


def ecMulA(a: bytes, d: bytes, l: int = 128):
    out = C.create_string_buffer(l // 2)
    ok = port().orc_ecMulA(sz(l), out, bytes(a), bytes(d), sz(len(d)))
    return ok, out.raw


# ---- reference wrappers (only where ref() is not None)
class RefParams(C.Structure):
    _fields_ = [("l", sz), ("p", C.c_ubyte * 64), ("a", C.c_ubyte * 64), ("b", C.c_ubyte * 64),
                ("q", C.c_ubyte * 64), ("yG", C.c_ubyte * 64), ("seed", C.c_ubyte * 8)]


_rp = {}


def ref_params(l: int = 128):
    if l not in _rp:
        _rp[l] = RefParams()
        assert ref().bignParamsStd(C.byref(_rp[l]), CURVES[l]) == 0
    return _rp[l]


def ref_bashHash(l: int, src: bytes) -> bytes:
    out = C.create_string_buffer(64)
    assert ref().bashHash(out, sz(l), bytes(src), sz(len(src))) == 0
    return out.raw[: l // 4]


def ref_beltCTR(src: bytes, key: bytes, iv: bytes) -> bytes:
    out = C.create_string_buffer(max(len(src), 1))
    assert ref().beltCTR(out, bytes(src), sz(len(src)), bytes(key), sz(len(key)), bytes(iv)) == 0
    return out.raw[: len(src)]


def ref_beltECBEncr(src: bytes, key: bytes) -> bytes:
    out = C.create_string_buffer(max(len(src), 1))
    assert ref().beltECBEncr(out, bytes(src), sz(len(src)), bytes(key), sz(len(key))) == 0
    return out.raw[: len(src)]


def ref_beltHash(src: bytes) -> bytes:
    out = C.create_string_buffer(32)
    assert ref().beltHash(out, bytes(src), sz(len(src))) == 0
    return out.raw


def ref_bignVerify(hash_: bytes, sig: bytes, pub: bytes, oid: bytes = OID, l: int = 128) -> int:
    return ref().bignVerify(C.byref(ref_params(l)), oid, sz(len(oid)), bytes(hash_), bytes(sig), bytes(pub))


def ref_bignSign2(hash_: bytes, priv: bytes, t: bytes = None, oid: bytes = OID, l: int = 128):
    sig = C.create_string_buffer(3 * l // 8)
    code = ref().bignSign2(sig, C.byref(ref_params(l)), oid, sz(len(oid)), bytes(hash_), bytes(priv), t,
                           sz(len(t) if t else 0))
    return code, sig.raw


def ref_bignPubkeyVal(pub: bytes, l: int = 128) -> int:
    return ref().bignPubkeyVal(C.byref(ref_params(l)), bytes(pub))


def ref_bignKeypairVal(priv: bytes, pub: bytes, l: int = 128) -> int:
    return ref().bignKeypairVal(C.byref(ref_params(l)), bytes(priv), bytes(pub))


def ref_bignDH(priv: bytes, pub: bytes, key_len: int, l: int = 128):
    key = C.create_string_buffer(max(key_len, 1))
    code = ref().bignDH(key, C.byref(ref_params(l)), bytes(priv), bytes(pub), sz(key_len))
    return code, key.raw[:key_len] if code == 0 else b""


def ref_bignKeypairGen(stream: bytes, l: int = 128):
    """bignKeypairGen of the reference fed from `stream`; returns (code, priv, pub, octets consumed)."""
    pos = [0]

    def fn(buf, count, state):
        chunk = stream[pos[0]:pos[0] + count]
        C.memmove(buf, chunk, count)
        pos[0] += count
    cb = C.CFUNCTYPE(None, C.c_void_p, sz, C.c_void_p)(fn)
    priv, pub = C.create_string_buffer(l // 4), C.create_string_buffer(l // 2)
    code = ref().bignKeypairGen(priv, pub, C.byref(ref_params(l)), cb, None)
    return code, priv.raw, pub.raw, pos[0]


def ref_bignSign(hash_: bytes, priv: bytes, stream: bytes, oid: bytes = OID, l: int = 128):
    """bignSign of the reference fed from `stream`; returns (code, sig, octets consumed)."""
    pos = [0]

    def fn(buf, count, state):
        C.memmove(buf, stream[pos[0]:pos[0] + count], count)
        pos[0] += count
    cb = C.CFUNCTYPE(None, C.c_void_p, sz, C.c_void_p)(fn)
    sig = C.create_string_buffer(3 * l // 8)
    code = ref().bignSign(sig, C.byref(ref_params(l)), oid, sz(len(oid)), bytes(hash_), bytes(priv), cb, None)
    return code, sig.raw, pos[0]


def ref_bignPubkeyCalc(priv: bytes, l: int = 128):
    pub = C.create_string_buffer(l // 2)
    code = ref().bignPubkeyCalc(pub, C.byref(ref_params(l)), bytes(priv))
    return code, pub.raw


def BashPrg(l, d, ann=b"", key=b""):
    """The oracle's bash-prg automaton behind the same Python interface as bee2_b200.BashPrg."""
    import bee2_b200.api as api

    class _OrcPrg(api.BashPrg):
        _prefix = "orc_bashPrg"
    return _OrcPrg(l, d, ann, key, _lib=port(), _keep=8 + 8 + 192 + 8 + 8)


def RefBashPrg(l, d, ann=b"", key=b""):
    import bee2_b200.api as api
    ref().bashPrg_keep.restype = C.c_size_t

    class _RefPrg(api.BashPrg):
        _prefix = "bashPrg"
    return _RefPrg(l, d, ann, key, _lib=ref(), _keep=ref().bashPrg_keep())
