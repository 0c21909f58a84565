#!/bin/bash
# cluster split-K (partials through distributed shared memory): correctness, sweep of (BN, S), whole-step A/B
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "split_k or cta_pair" 2>&1 | tail -15 > $O/r02t_ktests.log
tail -3 $O/r02t_ktests.log
if ! grep -q " passed" $O/r02t_ktests.log || grep -q "failed\|error" $O/r02t_ktests.log; then echo "KERNEL TESTS FAILED - stopping"; cat $O/r02t_ktests.log; exit 1; fi
echo "== default B=8"; $T 120 python tools/prof_conv.py 8
echo "== nosplit B=8"; WSR_SPLITK=0 $T 120 python tools/prof_conv.py 8
for f in 256,8 256,4 256,2 128,4 128,2 64,2; do echo "== force $f B=8"; WSR_SPLITK_FORCE=$f $T 120 python tools/prof_conv.py 8; done
echo "== default B=16"; $T 120 python tools/prof_conv.py 16
echo "== nosplit B=16"; WSR_SPLITK=0 $T 120 python tools/prof_conv.py 16
$T 400 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -4
for sk in 1 0; do
for b in 8 16 64; do
WSR_SPLITK=$sk $T 240 python bench.py --batch $b --steps 30 --no-cpu --no-extras --no-e2e > $O/r02t_b${b}_sk$sk.json 2> $O/r02t_b${b}_sk$sk.err
done
done
python - <<'PY'
import json
for b in (8,16,64):
  for sk in (1,0):
    f="r02t_b%d_sk%d"%(b,sk)
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "no result", e)
PY
