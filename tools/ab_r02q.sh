#!/bin/bash
# CTA-pair (cta_group::2) classic-mode convolution: correctness, micro-benchmarks with / without, whole-step A/B
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 120 python tools/prof_conv.py 8 > $O/r02q_conv_b8_pair.txt 2>&1; echo "prof_conv pair rc=$?"; tail -12 $O/r02q_conv_b8_pair.txt
$T 400 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -15 > $O/r02q_ktests.log
tail -3 $O/r02q_ktests.log
if ! grep -q " passed" $O/r02q_ktests.log || grep -q "failed\|error" $O/r02q_ktests.log; then echo "KERNEL TESTS FAILED - stopping"; exit 1; fi
WSR_PAIR=0 $T 120 python tools/prof_conv.py 8 > $O/r02q_conv_b8_nopair.txt 2>&1; tail -12 $O/r02q_conv_b8_nopair.txt
$T 120 python tools/prof_conv.py 64 > $O/r02q_conv_b64_pair.txt 2>&1; tail -12 $O/r02q_conv_b64_pair.txt
WSR_PAIR=0 $T 120 python tools/prof_conv.py 64 > $O/r02q_conv_b64_nopair.txt 2>&1; tail -12 $O/r02q_conv_b64_nopair.txt
$T 400 python -m pytest tests/test_parity_gpu.py tests/test_parity_bench_shapes_gpu.py -x -q -m gpu 2>&1 | tail -4
for pr in 1 0; do
WSR_PAIR=$pr $T 240 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --no-e2e --profile-ops > $O/r02q_b64_pair$pr.json 2> $O/r02q_b64_pair$pr.err
WSR_PAIR=$pr $T 240 python bench.py --batch 8 --steps 50 --no-cpu --no-extras --no-e2e --profile-ops > $O/r02q_b8_pair$pr.json 2> $O/r02q_b8_pair$pr.err
done
python - <<'PY'
import json
for f in ("r02q_b64_pair1","r02q_b64_pair0","r02q_b8_pair1","r02q_b8_pair0"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], d["roofline"]["per_op_ms"])
    except Exception as e: print(f, "no result", e)
PY
