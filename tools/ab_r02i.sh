#!/bin/bash
# A/B of the round-2 variants on one box; every step under its own timeout so that a hanging kernel cannot hold the box
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_groupnorm or final_conv or vertical" 2>&1 | tail -6 > $O/r02i_ktests.log
tail -2 $O/r02i_ktests.log
if ! grep -q " passed" $O/r02i_ktests.log || grep -q "failed\|error" $O/r02i_ktests.log; then echo "KERNEL TESTS FAILED - stopping"; exit 1; fi
WSR_FUSE_GN=1 $T 400 python -m pytest tests/test_parity_gpu.py tests/test_parity_bench_shapes_gpu.py -x -q -m gpu 2>&1 | tail -6 > $O/r02i_parity_fused.log
tail -2 $O/r02i_parity_fused.log
run() { # name, env..., args
  local name=$1; shift
  env "$@" $T 240 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --no-e2e $BARGS > $O/$name.json 2> $O/$name.err || echo "$name FAILED"
}
BARGS="--profile-ops" run r02i_b64_fuse WSR_FUSE_GN=1
BARGS="--profile-ops" run r02i_b64 WSR_FUSE_GN=0
BARGS="" run r02i_b64_noside WSR_NO_SIDE_STREAM=1
BARGS="--batch 8 --steps 50" run r02i_b8_fuse WSR_FUSE_GN=1
BARGS="--batch 8 --steps 50" run r02i_b8 WSR_FUSE_GN=0
BARGS="--batch 8 --steps 50" run r02i_b8_noside WSR_NO_SIDE_STREAM=1
python - <<'PY'
import json
for f in ("r02i_b64_fuse","r02i_b64","r02i_b64_noside","r02i_b8_fuse","r02i_b8","r02i_b8_noside"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], d["roofline"]["per_op_ms"])
    except Exception as e:
        print(f, "no result", e)
PY
