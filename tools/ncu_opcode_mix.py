#!/usr/bin/env python3
"""Executed SASS opcode mix of one kernel from an `ncu --set full --import-source on` capture:
    python tools/ncu_opcode_mix.py gpurun_out/ncu_r02_bign_verify.ncu-rep ITEMS
(per-item = warp instructions x 32 / ITEMS, the convention of profiles/r01_bign_opcode_mix.json)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from sass_hist import classify  # noqa: E402

rep, items = sys.argv[1], int(sys.argv[2])
if rep.endswith(".csv"):            # already exported on the GPU box: ncu -i x.ncu-rep --page source --csv
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernel = rows[0][1]
hdr = rows[1]
i_src, i_exec, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, samples, total = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
    if not m:
        continue
    n = int(r[i_exec] or 0)
    op = classify(m.group(1))
    mix[op] += n
    samples[op] += int(r[i_samp] or 0)
    total += n
doc = {"source": f"ncu --set full --import-source on, source page (SASS), {kernel}, {items} items",
       "warp_instructions_executed": total, "instructions_per_item": total * 32 / items,
       "by_opcode_per_item": {k: round(v * 32 / items, 1) for k, v in mix.most_common(24)},
       "stall_samples_by_opcode": dict(samples.most_common(12))}
json.dump(doc, sys.stdout, indent=1)
print()
