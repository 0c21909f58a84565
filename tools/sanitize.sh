#!/bin/bash
# compute-sanitizer passes over the hand-written tcgen05 / TMA / mbarrier kernels (small shapes: the kernel test file
# and one smoke()).  Usage (on the GPU box):  bash tools/sanitize.sh [tag]  -> gpurun_out/sanitize_<tag>_<tool>.log
# Each tool runs under its own timeout; the summary line of every log is what gets copied into profiles/.
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
TESTS=${SAN_TESTS:-"tests/test_kernels_gpu.py"}
for TOOL in ${SAN_TOOLS:-memcheck synccheck racecheck}; do
  LOG=$OUT/sanitize_${TAG}_${TOOL}.log
  echo "== $TOOL: pytest $TESTS" > $LOG
  timeout ${SAN_TIMEOUT:-900} $CS --tool $TOOL --print-limit 30 --launch-timeout 0 \
      python -m pytest $TESTS -x -q -m gpu -p no:cacheprovider >> $LOG 2>&1
  echo "== rc=$? (pytest under $TOOL)" >> $LOG
  echo "== $TOOL: smoke()" >> $LOG
  timeout ${SAN_TIMEOUT:-900} $CS --tool $TOOL --print-limit 30 --launch-timeout 0 \
      python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1
  echo "== rc=$? (smoke under $TOOL)" >> $LOG
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $LOG | tail -12
done
