#!/bin/bash
# 8 GPUs of one box: the default bench line (batch 64 in total, 8 per GPU; sub-records weak + train) and the reference arm under torchrun
cd $GRAFT_REPO_ROOT
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -k 5 600 $R --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_8gpu.json 2> $O/r02_bench_8gpu.err
tail -c 1500 $O/r02_bench_8gpu.json; echo; tail -3 $O/r02_bench_8gpu.err
timeout -k 5 200 $R --master-port 29542 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/r02_ref_8gpu.json 2>/dev/null; head -c 200 $O/r02_ref_8gpu.json; echo
