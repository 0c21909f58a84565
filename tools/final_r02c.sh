#!/bin/bash
# last evidence refresh of round 2 (final tree): tests, smoke, default line, 8-per-GPU record, training record
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/r02g_pytest.log
tail -2 $O/r02g_pytest.log
$T 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
$T 500 python bench.py --steps 20 --warmup 3 --profile-ops > $O/r02g_bench_default.json 2> $O/r02g_bench_default.err
tail -c 200 $O/r02g_bench_default.json; echo
$T 200 python bench.py --impl reference --steps 10 --warmup 2 > $O/r02g_bench_reference.json 2> /dev/null
$T 300 python bench.py --batch 8 --steps 50 --no-cpu --no-e2e --no-extras --profile-ops > $O/r02g_bench_b8.json 2> $O/r02g_bench_b8.err
$T 300 python bench.py --workload train --steps 20 --warmup 5 --profile-ops > $O/r02g_bench_train.json 2> $O/r02g_bench_train.err
python - <<'PY'
import json
for f in ("r02g_bench_default","r02g_bench_b8","r02g_bench_train"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["value"], d["clocks"])
PY
