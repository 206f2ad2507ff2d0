#!/bin/bash
mkdir -p gpurun_out
for c in 16384 32768 65536 75776 131072 151552 262144; do
B2G_BIGN_CHUNK=$c python - <<PY
import numpy as np, time, torch, bee2_b200 as b
OID=b.OID_BELT_HASH_DER
assert b.b2g_init(0)==0
units=1<<18
rng=np.random.default_rng(3)
priv=rng.integers(0,256,(units,32),dtype=np.uint8); priv[:,31]&=0x7F
hashes=rng.integers(0,256,(units,32),dtype=np.uint8)
params=b.bignParamsStd()
st1,pubs=b.bignPubkeyCalcBatch(params,priv); st2,sigs=b.bignSign2Batch(params,OID,hashes,priv)
ph,ps,pp=(torch.from_numpy(x).pin_memory() for x in (hashes,sigs,pubs))
nh,ns,npb=ph.numpy(),ps.numpy(),pp.numpy()
for _ in range(2): st=b.bignVerifyBatch(params,OID,nh,ns,npb)
assert not st.any()
t0=time.perf_counter()
for _ in range(5): b.bignVerifyBatch(params,OID,nh,ns,npb)
t=(time.perf_counter()-t0)/5
print("chunk $c: %.3f ms  %.1f M/s"%(t*1e3, units/t/1e6))
PY
done
