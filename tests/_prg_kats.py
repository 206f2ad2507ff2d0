"""The bash-prg programs of STB 34.101.77 Annex A.4 - A.6 as the reference runs them
(test/crypto/bash_test.c:155-294), parameterised by the automaton class under test."""


def run(P, H):
    st = P(256, 2, b"", H[:32])                                   # A.4.alpha
    st.absorb(H[32:32 + 95])
    st.ratchet()
    h16 = st.squeeze(16)
    assert h16.hex().upper() == "71CC358A0D5082173DE04803F7E905CB"
    st = P(128, 1, H[128:144], h16)                               # A.4.beta
    st1 = st.copy()
    ct = st.encr(H[160:183])
    assert ct.hex().upper() == "51ED3B28D345FFD1AD22815B86ECC17C278C8FE8920214"
    st = P(128, 1, H[128:144], h16)
    assert st.decr(ct) == H[160:183]
    st1.restart(H[144:148], b"")                                  # A.4.gamma
    st = st1.copy()
    ct = st1.encr(H[160:183])
    assert ct.hex().upper() == "28FE0998BFC010F13B260685A27AFB36CCF580F753521B"
    assert st.decr(ct) == H[160:183]
    for n, want in ((0, "36FA075EC15721F250B9A641A8CB99A333A9EE7BA8586D0646CBAC3686C03DF3"),     # A.5.1-4
                    (127, "C930FF427307420DA6E4182969AA1FFC3310179B8A0EDB3E20BEC285B568BA17"),
                    (128, "92AD1402C2007191F2F7CFAD6A2F8807BB0C50F73DFF95EF1B8AF08504D54007"),
                    (150, "48DB61832CA1009003BC0D8BDE67893A9DC683C48A5BC23AC884EB4613B480A6")):
        st = P(128, 2)
        st.absorb(H[:n])
        assert st.squeeze(32).hex().upper() == want, n
    st = P(128, 2)                                                # A.5.4 in steps
    st.absorb_start()
    st.absorb_step(b""), st.absorb_step(H[:50]), st.absorb_step(H[50:100]), st.absorb_step(H[100:150])
    st.squeeze_start()
    assert (st.squeeze_step(13) + st.squeeze_step(19)).hex().upper() == want
    for n, want in ((143, "6166032D6713D401A6BC687CCFFF2E603287143A84C78D2C62C71551E0E2FB2AF6B799EE33B5DECD7F62F190B1FBB052"),
                    (144, "8D84C82ECD0AB6468CC451CFC5EEB3B298DFD381D200DA69FBED5AE67D26BAD5C727E2652A225BF465993043039E338B"),
                    (150, "47529F9D499AB6AB8AD72B1754C90C39E7DA237BEB16CDFC00FE87934F5AFC1101862DFA50560F062A4DAC859CC13DBC")):
        st = P(192, 1)                                            # A.5.5-7
        st.absorb(H[:n])
        assert st.squeeze(48).hex().upper() == want, n
    st = P(256, 1, H[:16], H[32:64])                              # A.6.encr
    st.absorb(H[64:64 + 49])
    ct = st.encr(bytes(192))
    assert ct.hex().upper() == (
        "690673766C3E848CAC7C05169FFB7B7751E52A011040E5602573FAF991044A00"
        "4329EEF7BED8E6875830A91854D1BD2EDC6FC2FF37851DBAC249DF400A0549EA"
        "2E0C811D499E1FF1E5E32FAE7F0532FA4051D0F9E300D9B1DBF119AC8CFFC48D"
        "D3CBF1CA0DBA5DD97481C88DF0BE412785E40988B31585537948B80F5A9C49E0"
        "8DD684A7DCA871C380DFDC4C4DFBE61F50D2D0FBD24D8B9D32974A347247D001"
        "BAD5B168440025693967E77394DC088B0ECCFA8D291BA13D44F60B06E2EDB351")
    tag = st.squeeze(32)
    assert tag.hex().upper() == "CDE5AF6EF9A14B7D0C191B869A6343ED6A4E9AAB4EE00A579E9E682D0EC051E3"
    st = P(256, 1, H[:16], H[32:64])                              # A.6.decr
    st.absorb(H[64:64 + 49])
    st.decr_start()
    pt = b"".join(st.decr_step(ct[i:i + 32]) for i in range(0, 192, 32))
    assert pt == bytes(192)
    st.squeeze_start()
    assert st.squeeze_step(14) + st.squeeze_step(18) == tag


def random_program(P, Q, rng, H):
    """The same random command sequence on two automata; every output must agree."""
    l, d = int(rng.choice([128, 192, 256])), int(rng.choice([1, 2]))
    keyed = bool(rng.integers(0, 2))
    key = rng.integers(0, 256, 32, dtype="uint8").tobytes() if keyed else b""
    ann = rng.integers(0, 256, 4 * int(rng.integers(0, 5)), dtype="uint8").tobytes()
    a, b = P(l, d, ann, key), Q(l, d, ann, key)
    for _ in range(12):
        op = int(rng.integers(0, 6 if keyed else 3))
        n = int(rng.integers(0, 700))
        data = rng.integers(0, 256, n, dtype="uint8").tobytes()
        if op == 0:
            a.absorb(data), b.absorb(data)
        elif op == 1:
            assert a.squeeze(n) == b.squeeze(n)
        elif op == 2:
            a.ratchet(), b.ratchet()
        elif op == 3:
            assert a.encr(data) == b.encr(data)
        elif op == 4:
            assert a.decr(data) == b.decr(data)
        else:
            a.restart(ann, key), b.restart(ann, key)
    assert a.squeeze(64) == b.squeeze(64)
