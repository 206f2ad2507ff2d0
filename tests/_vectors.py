"""Golden fixtures shared by the CPU (oracle) and GPU (product) parity tests."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def resolve(expr, H: bytes, belt_hash=None) -> bytes:
    """'H[a:b]' -> slice of beltH(); 'beltHash(H[a:b])' -> its belt-hash; None -> None; else hex."""
    if expr is None:
        return None
    m = re.fullmatch(r"beltHash\((.*)\)", expr)
    if m:
        return belt_hash(resolve(m.group(1), H))
    m = re.fullmatch(r"H\[(\d+):(\d+)\]", expr)
    if m:
        return H[int(m.group(1)):int(m.group(2))]
    return bytes.fromhex(expr)


Q = int.from_bytes(bytes.fromhex("07663D2699BF5A7EFC4DFB0DD68E5CD9" + "FF" * 16), "little")


def sign2_nonce(sig: bytes, priv: bytes, h: bytes) -> bytes:
    """k = s1 + (s0 + 2^128) d + H mod q (test/crypto/bign_test.c:430-440)"""
    s0 = int.from_bytes(sig[:16], "little") + (1 << 128)
    s1 = int.from_bytes(sig[16:], "little")
    d = int.from_bytes(priv, "little")
    return ((s1 + s0 * d + int.from_bytes(h, "little")) % Q).to_bytes(32, "little")


def aead_incremental_program(mode, H):
    """The incremental A.19-1 / A.19-2 sequences of test/crypto/belt_test.c:474-520: (key, iv, steps,
    expected ciphertext hex, expected tag hex); steps are (op, data) with op in E/I/A/G, the A steps
    take the ciphertext produced by the E steps (marked by slices of "buf")."""
    key, iv = H[128:160], H[192:208]
    if mode == "DWP":
        return key, iv, [("E", H[:7]), ("E", H[7:16]), ("I", H[16:30]), ("G", None), ("I", H[30:48]), ("G", None),
                         ("A", (0, 12)), ("G", None), ("A", (12, 16)), ("G", None)], \
            "52C9AF96FF50F64435FC43DEF56BD797", "3B2E0AEB2B91854B"
    return key, iv, [("E", H[:11]), ("E", H[11:15]), ("I", H[16:30]), ("G", None), ("I", H[30:48]), ("G", None),
                     ("A", (0, 12)), ("G", None), ("A", (12, 15)), ("G", None)], \
        "BF3DAEAF5D18D2BCC30EA62D2E70A4", "548622B844123FF7"


def run_aead_program(st, steps):
    """Drive a BeltDWP-like object; returns (ciphertext, list of tags)."""
    buf, tags = b"", []
    for op, data in steps:
        if op == "E":
            buf += st.step_e(data)
        elif op == "I":
            st.step_i(data)
        elif op == "A":
            st.step_a(buf[data[0]:data[1]] if isinstance(data, tuple) else data)
        else:
            tags.append(st.step_g())
    return buf, tags
