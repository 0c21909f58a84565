#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 5 300 $R --master-port 29531 tools/ddp_check.py 2>&1 | tail -5
timeout -k 5 300 $R --master-port 29532 bench.py --gpus 2 --workload train --steps 20 --warmup 5 > $O/r02ac_train_2gpu.json 2> $O/r02ac_train_2gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02ac_train_2gpu.json")); print(round(d["ms_per_step"],3), d["value"], d.get("loss"), {k:d[k] for k in d if "allreduce" in k})
PY
