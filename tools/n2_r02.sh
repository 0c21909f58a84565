#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r02q_bench_2gpu.json 2> $O/r02q_bench_2gpu.err
tail -c 1500 $O/r02q_bench_2gpu.json; tail -5 $O/r02q_bench_2gpu.err
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r02q_ref_2gpu.json 2>/dev/null; head -c 300 $O/r02q_ref_2gpu.json
