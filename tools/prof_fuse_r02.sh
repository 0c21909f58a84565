#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
for shape in "64 64 64 128 256" "64 192 64 128 256" "64 128 128 64 128"; do
  for d in 0 1 2 3; do WSR_TC_DBG=$d $T 120 python tools/prof_fuse.py $shape 5; done
done > $O/r02j_fuse_decomp.txt 2>&1
cat $O/r02j_fuse_decomp.txt
$T 120 python tools/prof_fuse.py 16 64 64 128 256 1 > $O/r02j_plain.log 2>&1 && \
$T 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 4 --launch-count 1 -f -o $O/r02j_fuse python tools/prof_fuse.py 16 64 64 128 256 1 > $O/r02j_ncu.log 2>&1
tail -3 $O/r02j_ncu.log
