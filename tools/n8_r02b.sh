#!/bin/bash
# 8 GPUs: all-reduce bucket size of the training step
cd $GRAFT_REPO_ROOT
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
p=29551
for mb in 32 128 1024; do
WSR_BUCKET_MB=$mb timeout -k 5 200 $R --master-port $p bench.py --gpus 8 --workload train --steps 20 --warmup 5 > $O/r02_train8_mb$mb.json 2> $O/r02_train8_mb$mb.err
p=$((p+1))
done
python - <<'PY'
import json
for mb in (32,128,1024):
    try:
        d=json.load(open("gpurun_out/r02_train8_mb%d.json"%mb)); print(mb, round(d["ms_per_step"],3), d["value"], {k:d[k] for k in d if "allreduce" in k and "bytes" not in k})
    except Exception as e: print(mb, "no result", e)
PY
