#!/usr/bin/env python3
"""Static SASS evidence for the dominant kernels: per kernel the opcode histogram of the
whole function and of each loop body (backward branch), from `cuobjdump -sass` of the built objects.
    python tools/sass_hist.py > profiles/r02_sass_opcodes.json
The per-unit counts DESIGN.md quotes (LOP3/SHF per 6 bash rounds, LDS/PRMT per two belt blocks, wide MADs per
field product) are the loop-body / function rows of this file."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "bee2_b200", "csrc", "build")
INS = re.compile(r"^\s*/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")


def classify(op):
    base = op.split(".")[0]
    if base == "IMAD":
        if "WIDE" in op:
            return "IMAD.WIDE.X" if ".X" in op else "IMAD.WIDE"
        if "MOV" in op:
            return "IMAD.MOV"
        if ".X" in op:
            return "IMAD.X"
        if "IADD" in op:
            return "IMAD.IADD"
        if "SHL" in op:
            return "IMAD.SHL"
        return "IMAD"
    if base == "IADD3" and ".X" in op:
        return "IADD3.X"
    if base in ("LDS", "LDG", "STG", "STS", "LDL", "STL"):
        m = re.search(r"\.(32|64|128|256)", op)
        return base + ("." + m.group(1) if m else "")
    return base


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = INS.match(line)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2), m.group(3)))
    return res


def hist(ins):
    c = collections.Counter(classify(op) for _, op, _ in ins)
    return dict(sorted(c.items(), key=lambda kv: -kv[1]))


def loops(ins):
    out = []
    for a, op, args in ins:
        if op.startswith("BRA"):
            m = re.search(r"(0x[0-9a-f]+)", args)
            if m and int(m.group(1), 16) < a:
                out.append((int(m.group(1), 16), a))
    return out


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def report(obj, pattern, min_body=64):
    rep = {}
    for name, ins in functions(os.path.join(OBJ, obj)).items():
        d = demangle(name)
        if not re.search(pattern, d):
            continue
        entry = {"instructions": len(ins), "opcodes": hist(ins), "loops": []}
        for lo, hi in loops(ins):
            body = [i for i in ins if lo <= i[0] <= hi]
            if len(body) >= min_body:
                entry["loops"].append({"from": hex(lo), "to": hex(hi), "instructions": len(body), "opcodes": hist(body)})
        # out-of-line callees (CALL.REL targets inside the same function text): field products of bign
        calls = collections.Counter()
        for a, op, args in ins:
            if op.startswith("CALL"):
                m = re.search(r"(0x[0-9a-f]+)", args)
                if m:
                    calls[int(m.group(1), 16)] += 1
        rets = sorted(a for a, op, _ in ins if op.startswith("RET"))
        sub = []
        for tgt, n in calls.most_common(4):
            end = next((r for r in rets if r >= tgt), None)
            if end is None:
                continue
            body = [i for i in ins if tgt <= i[0] <= end]
            sub.append({"entry": hex(tgt), "call_sites": n, "instructions": len(body), "opcodes": hist(body)})
        if sub:
            entry["callees"] = sub
        rep[d] = entry
    return rep


if __name__ == "__main__":
    doc = {
        "how": "cuobjdump -sass bee2_b200/csrc/build/*.o (nvcc 12.9, -O3, sm_100a), tools/sass_hist.py; static counts",
        "bash": report("bash.o", r"bash_sponge_kernel<8, 16>|bash_f_kernel"),
        "belt": report("belt.o", r"belt_ctr_kernel|belt_ecb_kernel<false, true>"),
        "belt_dwp": report("belt_dwp.o", r"belt_dwp_mac_kernel|belt_dwp_fused"),
        "bign": report("bign.o", r"bign_verify_kernel<8, true>|bign_sign2_kernel<8>"),
    }
    json.dump(doc, sys.stdout, indent=1)
    print()
