#!/bin/bash
# Round-2 evidence run on one B200 (final tree): tests, smoke, default bench line, reference arm, per-GPU-batch 8 record, training record,
# ncu launch list (durations + DRAM bytes) and ncu tensor-pipe / SFU utilisation per launch of one B = 64 reverse step.
# Every step under its own timeout.
cd $GRAFT_REPO_ROOT
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/r02f_pytest.log
tail -2 $O/r02f_pytest.log
$T 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
$T 500 python bench.py --steps 20 --warmup 3 --profile-ops > $O/r02f_bench_default.json 2> $O/r02f_bench_default.err
tail -c 300 $O/r02f_bench_default.json; echo
$T 200 python bench.py --impl reference --steps 10 --warmup 2 > $O/r02f_bench_reference.json 2> /dev/null
$T 300 python bench.py --batch 8 --steps 50 --no-cpu --no-e2e --no-extras --profile-ops > $O/r02f_bench_b8.json 2> $O/r02f_bench_b8.err
$T 300 python bench.py --workload train --steps 20 --warmup 5 --profile-ops > $O/r02f_bench_train.json 2> $O/r02f_bench_train.err
$T 300 python bench.py --steps 300 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02f_bench_sustained.json 2> /dev/null
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras"
$T 200 $BCMD > $O/r02f_plain.log 2>&1 && \
$T 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 700 -c 400 --csv --log-file $O/r02f_ncu_launches.csv $BCMD > $O/r02f_ncu1.log 2>&1
tail -2 $O/r02f_ncu1.log
$T 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 700 -c 200 --csv --log-file $O/r02f_ncu_pipes.csv $BCMD > $O/r02f_ncu2.log 2>&1
tail -2 $O/r02f_ncu2.log
