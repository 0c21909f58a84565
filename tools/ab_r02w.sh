#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
for b in 8 16; do for m in 0 1 2; do
WSR_FUSE_GN=$m $T 240 python bench.py --batch $b --steps 50 --no-cpu --no-extras --no-e2e > $O/r02w_b${b}_m$m.json 2> $O/r02w_b${b}_m$m.err
done; done
python - <<'PY'
import json
for b in (8,16):
  for m in (0,1,2):
    f="r02w_b%d_m%d"%(b,m)
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "no result", e)
PY
