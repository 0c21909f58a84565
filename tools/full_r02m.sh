#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > $O/r02o_pytest.log
tail -3 $O/r02o_pytest.log
$T 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --profile-ops > $O/r02o_b64.json 2> $O/r02o_b64.err
$T 200 python bench.py --batch 8 --steps 50 --no-cpu --no-e2e --no-extras > $O/r02o_b8.json 2> $O/r02o_b8.err
python - <<'PY'
import json
for f in ("r02o_b64","r02o_b8"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["value"], d["clocks"]["sm_mhz"], d["roofline"]["per_op_ms"])
    except Exception as e: print(f, "no result", e)
PY
