#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests/test_train_gpu.py tests/test_entrypoints_gpu.py -x -q -m gpu 2>&1 | tail -4
for h in 1 0; do
WSR_HFCA_STREAM=$h $T 300 python bench.py --workload train --steps 20 --warmup 5 > $O/r02ab_train_h$h.json 2> $O/r02ab_train_h$h.err
done
python - <<'PY'
import json
for f in ("r02ab_train_h1","r02ab_train_h0"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["value"], d.get("loss"))
    except Exception as e: print(f, "no result", e)
PY
