#!/bin/bash
# one full ncu capture each of the 768->256 @32x64 convolution (B=64) as CTA pairs and as single CTAs
cd $GRAFT_REPO_ROOT
O=gpurun_out
export WSR_PROF_SHAPES=9 WSR_PROF_REPS=2
for pr in 1 0; do
WSR_PAIR=$pr timeout -k 5 120 python tools/prof_conv.py 64 || exit 1
WSR_PAIR=$pr timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o $O/r02q_conv_pair$pr python tools/prof_conv.py 64 > $O/r02q_ncu_pair$pr.log 2>&1
tail -3 $O/r02q_ncu_pair$pr.log
done
ls -la $O/*.ncu-rep
