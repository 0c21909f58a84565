#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_entrypoints_gpu.py -x -q -m gpu 2>&1 | tail -4
for b in 8 4 16; do
$T 240 python bench.py --batch $b --steps 50 --no-cpu --no-extras --no-e2e > $O/r02u_b${b}.json 2> $O/r02u_b${b}.err
WSR_PDL=0 $T 240 python bench.py --batch $b --steps 50 --no-cpu --no-extras --no-e2e > $O/r02u_b${b}_nopdl.json 2> $O/r02u_b${b}_nopdl.err
done
python - <<'PY'
import json
for b in (8,4,16):
  for sfx in ("","_nopdl"):
    f="r02u_b%d%s"%(b,sfx)
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "no result", e)
PY
