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
