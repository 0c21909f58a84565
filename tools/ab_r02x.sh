#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests/test_train_gpu.py tests/test_entrypoints_gpu.py -x -q -m gpu 2>&1 | tail -4
for w in 1 0; do
WSR_WGRAD_STREAM=$w $T 300 python bench.py --workload train --steps 20 --warmup 5 > $O/r02x_train_w$w.json 2> $O/r02x_train_w$w.err
WSR_WGRAD_STREAM=$w $T 300 python bench.py --workload train --train-batch 8 --steps 20 --warmup 5 > $O/r02x_train8_w$w.json 2> $O/r02x_train8_w$w.err
done
python - <<'PY'
import json
for f in ("r02x_train_w1","r02x_train_w0","r02x_train8_w1","r02x_train8_w0"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["value"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "no result", e)
PY
tail -3 $O/r02x_train_w1.err
