#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import bee2_b200 as b, json
assert b.b2g_init(0)==0
names={0:"lop3",4:"imad",5:"imad_wide",7:"lop3+imad_wide",15:"imad+imad_wide",16:"2iadd.x+imad_wide",17:"imad_wide.cc chain (wide MADs)",18:"iadd3.cc chain",19:"imad_wide x8 plain",13:"dfma",20:"8 cc-wide + 8 addc",21:"8 cc-wide + 16 addc",22:"8 plain wide + 8 addc"}
out={}
for k,n in names.items():
    out[n]=b.b2g_microbench(k,0)/1e12
    print(n, round(out[n],2))
json.dump(out,open("gpurun_out/micro.json","w"),indent=1)
PY
