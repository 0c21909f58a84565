#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 400 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
$T 400 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
for sk in 1 0; do
for b in 8 4; do
WSR_SPLITK=$sk $T 240 python bench.py --batch $b --steps 50 --no-cpu --no-extras --no-e2e > $O/r02ag_b${b}_sk$sk.json 2> /dev/null
done
WSR_SPLITK=$sk $T 300 python bench.py --workload train --steps 20 --warmup 5 > $O/r02ag_train_sk$sk.json 2> /dev/null
done
python - <<'PY'
import json
for n in ("b8","b4","train"):
  for sk in (1,0):
    f="r02ag_%s_sk%d"%(n,sk)
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3))
    except Exception as e: print(f, "no result", e)
PY
