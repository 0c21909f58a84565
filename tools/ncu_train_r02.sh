#!/bin/bash
# ncu launch list (durations) of one training step at batch 4 (after the warm-up steps)
cd $GRAFT_REPO_ROOT
O=gpurun_out
CMD="python bench.py --workload train --steps 3 --warmup 3"
timeout -k 5 200 $CMD > $O/r02_train_plain.log 2>&1 || exit 1
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1100 --csv --log-file $O/r02_ncu_train_launches.csv $CMD > $O/r02_ncu_train.log 2>&1
tail -2 $O/r02_ncu_train.log | cut -c1-300
grep -c '^"' $O/r02_ncu_train_launches.csv
