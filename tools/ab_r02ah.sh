#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests/test_train_gpu.py tests/test_entrypoints_gpu.py tests/test_pretrain_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
$T 300 python bench.py --workload train --steps 30 --warmup 6 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['value'], d.get('loss'))"
done
