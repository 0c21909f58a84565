#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_groupnorm or vertical" 2>&1 | tail -15 > $O/r02n_ktests.log
tail -3 $O/r02n_ktests.log
$T 300 python -m pytest tests/test_parity_bench_shapes_gpu.py -x -q -m gpu -k config1 2>&1 | tail -40 > $O/r02n_cfg1.log
grep -n "Error\|error\|assert" $O/r02n_cfg1.log | head -10
if ! grep -q " passed" $O/r02n_ktests.log || grep -q "failed\|error" $O/r02n_ktests.log; then echo "KERNEL TESTS FAILED - stopping"; exit 1; fi
for shape in "64 64 64 128 256" "64 192 64 128 256" "64 128 128 64 128" "64 384 128 64 128"; do $T 120 python tools/prof_fuse.py $shape 5; done
WSR_TC_DBG=3 $T 120 python tools/prof_fuse.py 64 64 64 128 256 5
WSR_FUSE_GN=1 $T 400 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -4
WSR_FUSE_GN=1 $T 240 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --no-e2e --profile-ops > $O/r02n_b64_fuse.json 2> $O/r02n_b64_fuse.err
WSR_FUSE_GN=1 $T 240 python bench.py --batch 8 --steps 50 --no-cpu --no-extras --no-e2e > $O/r02n_b8_fuse.json 2> $O/r02n_b8_fuse.err
python - <<'PY'
import json
for f in ("r02n_b64_fuse","r02n_b8_fuse"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], d["roofline"]["per_op_ms"])
    except Exception as e: print(f, "no result", e)
PY
