"""Boundary acceptance: the reference's OWN tests (test/crypto/{bash,belt,bign,bign128,bign192,bign256}_test.c,
test/math/{ec,ecp}_test.c — compiled from /root/reference by oracle/Makefile, driver oracle/reftests_main.c)
run in a process where libbee2_b200.so stands in front of the unmodified reference library. Every symbol
this library exports is served by the GPU path, also for the reference's own internals (bignKeyWrap ->
ecMulA, beltCBC -> beltBlockEncr, ...); the rest (DER, brng, other belt modes) is the reference's.

Also the overlay routing of SURVEY.md §8b: inputs the GPU path does not cover (a generic ec_o, scalars
longer than the field) and — on request — small one-shot calls go to the stock library behind
(dlsym(RTLD_NEXT)), never to oracle/."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "reftests_b200")
LINE = re.compile(r"^(\w+)Test: (OK|Err)\s+\(gpu launches (\d+), forwarded to stock (\d+)\)", re.M)


def _run(names, env=None, timeout=1500):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([BIN] + list(names), capture_output=True, text=True, timeout=timeout, env=e)
    res = {m.group(1): (m.group(2), int(m.group(3)), int(m.group(4))) for m in LINE.finditer(r.stdout)}
    return r, res


def test_reftests_binary_binds_the_engine_first():
    """CPU check: the acceptance binary exists (built where /root/reference is) and resolves the hot-path
    symbols to libbee2_b200.so, which precedes the reference library in its search order."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/reftests_b200 not built (no /root/reference at build time)")
    out = subprocess.run(["readelf", "-d", BIN], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert "libbee2_b200.so" in needed and "libbee2ref_64.so" in needed
    assert needed.index("libbee2_b200.so") < needed.index("libbee2ref_64.so")


@pytest.mark.gpu
def test_reference_own_tests_pass_on_the_gpu_path():
    assert os.path.exists(BIN), "oracle/_ref/reftests_b200 missing: run __graft_entry__.build() where /root/reference exists"
    r, res = _run(["bash", "belt", "bign", "bign128", "bign192", "bign256"], env={"B2G_TRACE_FORWARD": "1"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "stock libbee2 behind: yes" in r.stdout
    for name in ("bash", "belt", "bign", "bign128", "bign192", "bign256"):
        verdict, launches, forwards = res[name]
        assert verdict == "OK", r.stdout
        assert launches > 0, f"{name}Test launched no kernel"
    # nothing in the bash / belt / bign128-256 tests is outside the GPU path's coverage: no call may have gone
    # to the CPU library. bign_test.c also validates parameter blocks (bignParamsVal -> ecMulA by the group
    # order + cofactor words, m > n): only such scalar multiplications may be forwarded.
    forwarded = set(re.findall(r"^b2g-forward (\w+)", r.stderr, re.M))
    assert forwarded <= {"ecMulA", "ecAddMulA", "ecMulA_deep"}, forwarded
    for name in ("bash", "belt", "bign128", "bign192", "bign256"):
        assert res[name][2] == 0, f"{name}Test forwarded {res[name][2]} calls to stock"


@pytest.mark.gpu
def test_unsupported_inputs_are_forwarded_to_stock_not_aborted():
    """ec_test.c / ecp_test.c drive ecMulA / ecAddMulA with scalars longer than the field (m = n + 1, 5 words)
    and, in ecp_test.c, on other curves: those calls must reach the stock library through RTLD_NEXT (round 1
    aborted the process), the in-range ones on the standard curve stay on the GPU."""
    r, res = _run(["ec", "ecp"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    for name in ("ec", "ecp"):
        assert res[name][0] == "OK", r.stdout
    assert res["ec"][2] > 0, "no call was forwarded"
    assert res["ec"][1] + res["ecp"][1] > 0, "no call ran on the GPU"


@pytest.mark.gpu
def test_small_call_routing_to_stock():
    """b2g_set_cpu_below / B2G_CPU_BELOW: one-shot calls with a payload below the threshold (a 13-byte
    bashHash, one beltBlockEncr, a single bignVerify) go to the CPU library; results are the same."""
    r, res = _run(["bash", "belt", "bign128"], env={"B2G_CPU_BELOW": str(1 << 20)})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    for name in ("bash", "belt", "bign128"):
        assert res[name][0] == "OK", r.stdout
        assert res[name][2] > 0, f"{name}Test: nothing was routed to stock"


@pytest.mark.gpu
def test_overlay_demo_same_output_stock_linked_preloaded():
    """examples/overlay_demo.c, a plain bee2 application: built against stock libbee2 only, built with the
    engine in front of it, and the stock build with the engine LD_PRELOADed — identical digests, signatures
    and verdicts (incl. the foreign-curve call the engine forwards), and the accelerated runs did launch
    kernels."""
    d = os.path.join(ROOT, "examples", "_build")
    stock, linked = os.path.join(d, "demo_stock"), os.path.join(d, "demo_linked")
    assert os.path.exists(stock) and os.path.exists(linked), "examples/_build missing: run __graft_entry__.build()"
    a = subprocess.run([stock], capture_output=True, text=True, timeout=300)
    b_ = subprocess.run([linked], capture_output=True, text=True, timeout=300)
    env = dict(os.environ, LD_PRELOAD=os.path.join(ROOT, "bee2_b200", "libbee2_b200.so"))
    c = subprocess.run([stock], capture_output=True, text=True, timeout=300, env=env)
    assert a.returncode == 0 and b_.returncode == 0 and c.returncode == 0, (a.stderr, b_.stderr, c.stderr)
    assert "stock libbee2 only" in a.stderr
    assert a.stdout == b_.stdout == c.stdout, (a.stdout, b_.stdout, c.stdout)
    for r in (b_, c):
        m = re.search(r"engine present: (\d+) kernel launches, (\d+) calls forwarded", r.stderr)
        assert m and int(m.group(1)) > 0 and int(m.group(2)) >= 1, r.stderr
