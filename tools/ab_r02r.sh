#!/bin/bash
# programmatic dependent launch at small batch (latency-bound steps), with the pair kernels restricted to > 1 wave
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
for pdl in 0 1; do
for b in 8 16; do
WSR_PDL=$pdl $T 240 python bench.py --batch $b --steps 50 --no-cpu --no-extras --no-e2e > $O/r02r_b${b}_pdl$pdl.json 2> $O/r02r_b${b}_pdl$pdl.err
done
done
python - <<'PY'
import json
for f in ("r02r_b8_pdl0","r02r_b8_pdl1","r02r_b16_pdl0","r02r_b16_pdl1"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "no result", e)
PY
