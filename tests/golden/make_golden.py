#!/usr/bin/env python3
"""Generates tests/golden/kat.json and tests/golden/ref_vectors.json.

kat.json        — the STB annex known-answer vectors the reference's own tests hold for the
                  hot path, transcribed with the file:line they come from. Inputs are cut from
                  the belt S-box table beltH() exactly as in the reference tests.
ref_vectors.json — seeded random inputs with the outputs of the UNMODIFIED reference
                  (oracle/_ref/libbee2ref_64.so, built from /root/reference by oracle/Makefile).
                  Needs this container (the reference does not exist on the GPU box); the
                  committed JSON travels instead.

Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as o  # noqa: E402

# ---------------------------------------------------------------- STB annex vectors
# "H[a:b]" means beltH()[a:b]
KAT = {
    "source": "agievich/bee2 @ d9e689a0, test/crypto/{bash,belt,bign}_test.c",
    "bashF": [  # bash_test.c:41-57 (A.2)
        {"id": "A.2", "in": "H[0:192]", "out":
         "8FE727775EA7F140B95BB6A200CBB28C7F0809C0C0BC68B7DC5AEDC841BD94E403630C301FC255DF5B67DB53EF65E376"
         "E8A4D797A6172F2271BA48093173D329C3502AC946767326A2891971392D3F7089959F5D61621238655975E00E2132A0"
         "D5018CEEDB17731CCD88FC50151D37C0D4A3359506AEDC2E6109511E7703AFBB014642348D8568AA1A5D9868C4C7E6DF"
         "A756B1690C7C2608A2DC136F5997AB8FBB3F4D9F033C87CA6070E117F099C4094972ACD9D976214B7CED8E3F8B6E058E"},
    ],
    "bashHash": [  # bash_test.c:58-154 (A.3.1 - A.3.11); l is the security level
        {"id": "A.3.1", "l": 128, "len": 0, "out": "114C3DFAE373D9BCBC3602D6386F2D6A2059BA1BF9048DBAA5146A6CB775709D"},
        {"id": "A.3.2", "l": 128, "len": 127, "out": "3D7F4EFA00E9BA33FEED259986567DCF5C6D12D51057A968F14F06CC0F905961"},
        {"id": "A.3.3", "l": 128, "len": 128, "out": "D7F428311254B8B2D00F7F9EEFBD8F3025FA87C4BABD1BDDBE87E35B7AC80DD6"},
        {"id": "A.3.4", "l": 128, "len": 135, "out": "1393FA1B65172F2D18946AEAE576FA1CF54FDD354A0CB2974A997DC4865D3100"},
        {"id": "A.3.5", "l": 192, "len": 95, "out":
         "64334AF830D33F63E9ACDFA184E32522103FFF5C6860110A2CD369EDBC04387C501D8F92F749AE4DE15A8305C353D64D"},
        {"id": "A.3.6", "l": 192, "len": 96, "out":
         "D06EFBC16FD6C0880CBFC6A4E3D65AB101FA82826934190FAABEBFBFFEDE93B22B85EA72A7FB3147A133A5A8FEBD8320"},
        {"id": "A.3.7", "l": 192, "len": 108, "out":
         "FF763296571E2377E71A1538070CC0DE88888606F32EEE6B082788D246686B00FC05A17405C5517699DA44B7EF5F55AB"},
        {"id": "A.3.8", "l": 256, "len": 63, "out":
         "2A66C87C189C12E255239406123BDEDBF19955EAF0808B2AD705E249220845E2"
         "0F4786FB6765D0B5C48984B1B16556EF19EA8192B985E4233D9C09508D6339E7"},
        {"id": "A.3.9", "l": 256, "len": 64, "out":
         "07ABBF8580E7E5A321E9B940F667AE209E2952CEF557978AE743DB086BAB4885"
         "B708233C3F5541DF8AAFC3611482FDE498E58B3379A6622DAC2664C9C118A162"},
        {"id": "A.3.10", "l": 256, "len": 127, "out":
         "526073918F97928E9D15508385F42F03ADE3211A23900A30131F8A1E3E1EE21C"
         "C09D13CFF6981101235D895746A4643F0AA62B0A7BC98A269E4507A257F0D4EE"},
        {"id": "A.3.11", "l": 256, "len": 192, "out":
         "8724C7FF8A2A83F22E38CB9763777B96A70ABA3444F214C763D93CD6D19FCFDE"
         "6C3D3931857C4FF6CCCD49BD99852FE9EAA7495ECCDD96B571E0EDCF47F89768"},
    ],
    "beltBlock": [  # belt_test.c:178-215
        {"id": "A.1", "op": "encr", "in": "H[0:16]", "key": "H[128:160]", "out": "69CCA1C93557C9E3D66BC3E0FA88FA6E"},
        {"id": "A.4", "op": "decr", "in": "H[64:80]", "key": "H[160:192]", "out": "0DC5300600CAB840B38448E5E993F421"},
    ],
    "beltECB": [  # belt_test.c:288-339 (A.9-1/2 encrypt, A.10-1/2 decrypt; -2 use ciphertext stealing)
        {"id": "A.9-1", "op": "encr", "in": "H[0:48]", "key": "H[128:160]", "split": [32, 16], "out":
         "69CCA1C93557C9E3D66BC3E0FA88FA6E5F23102EF109710775017F73806DA9DC46FB2ED2CE771F26DCB5E5D1569F9AB0"},
        {"id": "A.9-2", "op": "encr", "in": "H[0:47]", "key": "H[128:160]", "split": [16, 31], "out":
         "69CCA1C93557C9E3D66BC3E0FA88FA6E36F00CFED6D1CA1498C12798F4BEB2075F23102EF109710775017F73806DA9"},
        {"id": "A.10-1", "op": "decr", "in": "H[64:112]", "key": "H[160:192]", "split": [16, 32], "out":
         "0DC5300600CAB840B38448E5E993F421E55A239F2AB5C5D5FDB6E81B40938E2A54120CA3E6E19C7AD750FC3531DAEAB7"},
        {"id": "A.10-2", "op": "decr", "in": "H[64:100]", "key": "H[160:192]", "split": [16, 20], "out":
         "0DC5300600CAB840B38448E5E993F4215780A6E2B69EAFBB258726D7B6718523E55A239F"},
    ],
    "beltCTR": [  # belt_test.c:423-448 (A.15, A.16) with the reference's split calls
        {"id": "A.15", "in": "H[0:48]", "key": "H[128:160]", "iv": "H[192:208]", "split": [15, 7, 26], "out":
         "52C9AF96FF50F64435FC43DEF56BD797D5B5B1FF79FB41257AB9CDF6E63E81F8F00341473EAE409833622DE05213773A"},
        {"id": "A.16", "in": "H[64:108]", "key": "H[160:192]", "iv": "H[208:224]", "split": [11, 5, 28], "out":
         "DF181ED008A20F43DCBBB93650DAD34B389CDEE5826D40E2D4BD80F49A93F5D212F6333166456F169043CC5F"},
    ],
    "beltHash": [  # belt_test.c:593-631 (A.23-1/2/3)
        {"id": "A.23-1", "in": "H[0:13]", "out": "ABEF9725D4C5A83597A367D14494CC2542F20F659DDFECC961A3EC550CBA8C75"},
        {"id": "A.23-2", "in": "H[0:32]", "out": "749E4C3653AECE5E48DB4761227742EB6DBE13F4A80F7BEFF1A9CF8D10EE7786"},
        {"id": "A.23-3", "in": "H[0:48]", "out": "9D02EE446FB6A29FE5C982D4B13AF9D3E90861BC4CEF27CF306BFB0B174A154A"},
    ],
    "beltDWP": [  # belt_test.c:473-497 (A.19-1 wrap), :523-543 (A.20-1 unwrap)
        {"id": "A.19-1", "op": "wrap", "in": "H[0:16]", "open": "H[16:48]", "key": "H[128:160]", "iv": "H[192:208]",
         "out": "52C9AF96FF50F64435FC43DEF56BD797", "mac": "3B2E0AEB2B91854B"},
        {"id": "A.20-1", "op": "unwrap", "in": "H[64:80]", "open": "H[80:112]", "key": "H[160:192]", "iv": "H[208:224]",
         "out": "DF181ED008A20F43DCBBB93650DAD34B", "mac": "6A2C2C94C4150DC0"},
    ],
    "beltCHE": [  # belt_test.c:498-522 (A.19-2 wrap), :544-564 (A.20-2 unwrap)
        {"id": "A.19-2", "op": "wrap", "in": "H[0:15]", "open": "H[16:48]", "key": "H[128:160]", "iv": "H[192:208]",
         "out": "BF3DAEAF5D18D2BCC30EA62D2E70A4", "mac": "548622B844123FF7"},
        {"id": "A.20-2", "op": "unwrap", "in": "H[64:84]", "open": "H[80:112]", "key": "H[160:192]", "iv": "H[208:224]",
         "out": "2BABF43EB37B5398A9068F31A3C758B762F44AA9", "mac": "7D9D4F59D40D197D"},
    ],
    "beltZerosum": {  # belt_test.c:69-110: XOR over i of X_i ^ Belt_0(X_i) = 0, X_i = (x_i,0,0,0)
        "x": [15014, 124106, 166335, 206478, 313245, 366839, 455597, 502723, 535141, 625112, 659461, 752253, 801048,
              897899, 943850, 1041695, 1101266, 1170856, 1217537, 1248520, 1366084, 1421171, 1448429, 1514215, 1573855,
              1701341, 1738016, 1781705, 1837300, 1948449, 1999650, 2089289, 2117830, 2175758, 2249930, 2358928,
              2404262, 2447467, 2552783, 2556713, 2678348, 2705770, 2808011, 2827994, 2948039, 2995213, 3029188,
              3096649, 3170243, 3230306, 3285991, 3350691, 3457162, 3500592, 3539783, 3636611, 3735543, 3752463,
              3814136, 3875630, 3935109, 4002291, 4088401, 4129247, 4257830, 4266427, 4352389, 4397389, 4470348,
              4531932, 4598961, 4691323, 4747531, 4839756, 4900773, 4958368, 5021928, 5099836, 5164752, 5214964,
              5269476, 5356247, 5391667, 5496861, 5561223, 5601750, 5700311, 5761736, 5812345, 5856838, 5956987,
              5966502, 6059392, 6104328, 6193021, 6233226, 6311341, 6369016, 6475468, 6540894, 6598453, 6666092,
              6711620, 6804478, 6834201, 6932158, 6971325, 7059579, 7089192, 7188715, 7245095, 7325355, 7367748,
              7426778, 7475903, 7599231, 7643174, 7722266, 7747291, 7832837, 7887591, 7942192, 8043937, 8108261,
              8169299, 8233361, 8305861, 8367181]},
    "bign": {  # bign_test.c:303-357 (G.1, G.2 + negatives), :388-400 (G.3), :417-457 (G.6, G.7)
        "oid_der": "06092A7000020022651F51",
        "privkey": "1F66B5B84B7339674533F0329C74F21834281FED0732429E0C79235FC273E269",
        "pubkey": "BD1A5650179D79E03FCEE49D4C2BD5DDF54CE46D0CF11E4FF87BF7A890857FD0"
                  "7AC6A60361E8C8173491686D461B2826190C2EDA5909054A9AB84D2AB9D99A90",
        "verify": [
            {"id": "G.2", "hash": "beltHash(H[0:13])", "sig":
             "E36B7F0377AE4C524027C387FADF1B20CE72F1530B71F2B5FD3A8C584FE2E1AED20082E30C8AF65011F4FB54649DFD3D"},
            {"id": "G.3", "hash": "beltHash(H[0:48])", "sig":
             "47A63C8B9C936E94B5FAB3D9CBD78366290F3210E163EEC8DB4E921E8479D4138F112CC23E6DCE65EC5FF21DF4231C28"},
        ],
        "sign2_nonce": [  # k recovered as s1 + (s0 + 2^128) d + H mod q
            {"id": "G.6", "hash": "beltHash(H[0:13])", "t": None,
             "k": "829614D8411DBBC4E1F2471A4004586440FD8C9553FAB6A1A45CE417AE97111E"},
            {"id": "G.7", "hash": "beltHash(H[0:48])", "t": "H[192:215]",
             "k": "7ADC8713283EBFA547A2AD9CDFB245AE0F7B968DF0F91CB785D1F932A3583107"},
        ],
    },
}


def main():
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(KAT, f, indent=1)
    print("wrote kat.json")
    if o.ref() is None:
        print("oracle/_ref/libbee2ref_64.so absent: ref_vectors.json not regenerated")
        return
    rng = np.random.default_rng(20261017)
    rb = lambda n: rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()  # noqa: E731
    out = {"source": "unmodified reference via oracle/_ref/libbee2ref_64.so, seed 20261017", "bashHash": [],
           "beltCTR": [], "beltECB": [], "beltHash": [], "bign": []}
    for l, n in [(128, 0), (128, 1), (128, 127), (128, 128), (128, 129), (128, 1000), (192, 95), (192, 96), (192, 97),
                 (192, 500), (256, 63), (256, 64), (256, 65), (256, 4096), (256, 4097), (16, 10), (32, 200), (64, 361),
                 (112, 137), (240, 333), (176, 87), (176, 88), (176, 89)]:
        m = rb(n)
        out["bashHash"].append({"l": l, "in": m.hex(), "out": o.ref_bashHash(l, m).hex()})
    for klen, n in [(32, 0), (32, 1), (32, 15), (32, 16), (32, 17), (32, 1000), (16, 100), (24, 333), (32, 4096)]:
        k, iv, m = rb(klen), rb(16), rb(n)
        out["beltCTR"].append({"key": k.hex(), "iv": iv.hex(), "in": m.hex(), "out": o.ref_beltCTR(m, k, iv).hex()})
    for klen, n in [(32, 16), (32, 17), (32, 31), (32, 32), (32, 33), (16, 100), (24, 333), (32, 1024)]:
        k, m = rb(klen), rb(n)
        out["beltECB"].append({"key": k.hex(), "in": m.hex(), "out": o.ref_beltECBEncr(m, k).hex()})
    out["beltDWP"] = []
    for klen, n1, n2 in [(32, 0, 0), (32, 1, 0), (32, 0, 1), (32, 15, 17), (32, 16, 16), (16, 33, 5), (24, 1000, 77),
                         (32, 4096, 4096), (32, 5000, 3)]:
        k, iv, a, b_ = rb(klen), rb(16), rb(n1), rb(n2)
        import ctypes as C
        d, m = C.create_string_buffer(max(n1, 1)), C.create_string_buffer(8)
        assert o.ref().beltDWPWrap(d, m, a, C.c_size_t(n1), b_, C.c_size_t(n2), k, C.c_size_t(klen), iv) == 0
        out["beltDWP"].append({"key": k.hex(), "iv": iv.hex(), "in": a.hex(), "open": b_.hex(),
                               "out": d.raw[:n1].hex(), "mac": m.raw.hex()})
        assert o.ref().beltCHEWrap(d, m, a, C.c_size_t(n1), b_, C.c_size_t(n2), k, C.c_size_t(klen), iv) == 0
        out.setdefault("beltCHE", []).append({"key": k.hex(), "iv": iv.hex(), "in": a.hex(), "open": b_.hex(),
                                              "out": d.raw[:n1].hex(), "mac": m.raw.hex()})
    for n in [0, 1, 31, 32, 33, 64, 75, 1000]:
        m = rb(n)
        out["beltHash"].append({"in": m.hex(), "out": o.ref_beltHash(m).hex()})
    for i in range(24):
        d = bytearray(rb(32))
        d[31] &= 0x7F
        h = rb(32)
        t = None if i % 3 else rb(5 + i)
        code, pub = o.ref_bignPubkeyCalc(bytes(d))
        assert code == 0
        code, sig = o.ref_bignSign2(h, bytes(d), t)
        assert code == 0
        rec = {"privkey": bytes(d).hex(), "pubkey": pub.hex(), "hash": h.hex(), "t": t.hex() if t else None,
               "sig": sig.hex(), "verify": o.ref_bignVerify(h, sig, pub)}
        # one corrupted variant per record, with the reference's verdict
        bad_sig, bad_pub, bad_h = bytearray(sig), bytearray(pub), bytearray(h)
        kind = i % 6
        if kind == 0:
            bad_sig[i % 16] ^= 1 << (i % 8)
        elif kind == 1:
            bad_sig[16 + i % 32] ^= 1 << (i % 8)
        elif kind == 2:
            bad_h[i % 32] ^= 0x80
        elif kind == 3:
            bad_pub[i % 64] ^= 1
        elif kind == 4:
            bad_sig[16:48] = b"\xff" * 32          # s1 >= q
        else:
            bad_pub[0:32] = b"\xff" * 32           # Qx >= p
        rec["bad"] = {"hash": bytes(bad_h).hex(), "sig": bytes(bad_sig).hex(), "pubkey": bytes(bad_pub).hex(),
                      "verify": o.ref_bignVerify(bytes(bad_h), bytes(bad_sig), bytes(bad_pub))}
        out["bign"].append(rec)
    # levels 192 / 256 (bign-curve384v1 / 512v1) with the OIDs of bign192.c / bign256.c; an own generator so
    # that the records above keep their values
    rng2 = np.random.default_rng(20261018)
    rb2 = lambda n: rng2.integers(0, 256, size=n, dtype=np.uint8).tobytes()  # noqa: E731
    out["bignL"] = []
    for l in (192, 256):
        no, oid = l // 4, o.OIDS[l]
        for i in range(10):
            d = bytearray(rb2(no))
            d[no - 1] &= 0x7F
            h = rb2(no) if i != 7 else b"\xff" * no      # i = 7: H >= q
            t = None if i % 3 else rb2(5 + i)
            code, pub = o.ref_bignPubkeyCalc(bytes(d), l)
            assert code == 0
            code, sig = o.ref_bignSign2(h, bytes(d), t, oid, l)
            assert code == 0
            rec = {"l": l, "privkey": bytes(d).hex(), "pubkey": pub.hex(), "hash": h.hex(), "t": t.hex() if t else None,
                   "sig": sig.hex(), "verify": o.ref_bignVerify(h, sig, pub, oid, l)}
            assert rec["verify"] == 0
            bad_sig, bad_pub, bad_h = bytearray(sig), bytearray(pub), bytearray(h)
            kind = i % 6
            if kind == 0:
                bad_sig[i % (no // 2)] ^= 1 << (i % 8)
            elif kind == 1:
                bad_sig[no // 2 + i % no] ^= 1 << (i % 8)
            elif kind == 2:
                bad_h[i % no] ^= 0x80
            elif kind == 3:
                bad_pub[i % (2 * no)] ^= 1
            elif kind == 4:
                bad_sig[no // 2:] = b"\xff" * no        # s1 >= q
            else:
                bad_pub[0:no] = b"\xff" * no            # Qx >= p
            rec["bad"] = {"hash": bytes(bad_h).hex(), "sig": bytes(bad_sig).hex(), "pubkey": bytes(bad_pub).hex(),
                          "verify": o.ref_bignVerify(bytes(bad_h), bytes(bad_sig), bytes(bad_pub), oid, l)}
            out["bignL"].append(rec)
    # key generation / validation / Diffie-Hellman (bign_misc.c), all three levels; own generator again
    rng3 = np.random.default_rng(20261019)
    rb3 = lambda n: rng3.integers(0, 256, size=n, dtype=np.uint8).tobytes()  # noqa: E731
    out["bignMisc"] = []
    for l in (128, 192, 256):
        no = l // 4
        # generator stream: an all-ones draw (>= p, rejected), an all-zero draw (rejected), then real octets
        stream = b"\xff" * no + bytes(no) + rb3(3 * no)
        code, priv, pub, used = o.ref_bignKeypairGen(stream, l)
        assert code == 0 and used == 3 * no
        rec = {"l": l, "stream": stream.hex(), "privkey": priv.hex(), "pubkey": pub.hex(), "used": used, "dh": [], "val": []}
        for i in range(4):
            u = bytearray(rb3(no))
            u[no - 1] &= 0x7F
            if i == 3:
                u = bytearray(no)                       # d = 0 -> ERR_BAD_PRIVKEY
            key_len = [2 * no, no, 7, no + 9][i]
            code, key = o.ref_bignDH(bytes(u), pub, key_len, l)
            rec["dh"].append({"privkey": bytes(u).hex(), "key_len": key_len, "code": code, "key": key.hex()})
        offcurve = bytearray(pub)
        offcurve[1] ^= 4
        big = b"\xff" * no + pub[no:]
        for name, pk in (("ok", pub), ("offcurve", bytes(offcurve)), ("x>=p", big)):
            rec["val"].append({"case": name, "pubkey": pk.hex(), "pubkey_val": o.ref_bignPubkeyVal(pk, l),
                               "keypair_val": o.ref_bignKeypairVal(priv, pk, l),
                               "dh": o.ref_bignDH(priv, pk, no, l)[0]})
        # bignSign with a generator stream whose first draws are rejected (>= q, zero)
        h = rb3(no)
        sstream = b"\xff" * no + bytes(no) + rb3(2 * no)
        code, sig, used = o.ref_bignSign(h, priv, sstream, o.OIDS[l], l)
        assert code == 0 and used == 3 * no and o.ref_bignVerify(h, sig, pub, o.OIDS[l], l) == 0
        rec["sign"] = {"hash": h.hex(), "stream": sstream.hex(), "sig": sig.hex(), "used": used}
        out["bignMisc"].append(rec)
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote ref_vectors.json")


if __name__ == "__main__":
    main()
