#!/bin/bash
# Runs every GPU test in its own process under a hard timeout, so one hung kernel cannot hide the rest.
# usage: tools/run_gpu_tests_isolated.sh [per-test-seconds] [extra pytest selection args]
T=${1:-90}
shift
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests_isolated.log
: > $LOG
python -m pytest tests --collect-only -q -m gpu "$@" 2>/dev/null | grep '::' > /tmp/ids.txt
while read -r id; do
  s=$(date +%s)
  timeout -s KILL $T python -m pytest "$id" -x -q -p no:cacheprovider > /tmp/one.log 2>&1
  rc=$?
  e=$(date +%s)
  echo "$id rc=$rc $((e - s))s" | tee -a $LOG
  if [ $rc -ne 0 ]; then tail -25 /tmp/one.log >> $LOG; fi
done < /tmp/ids.txt
echo "passed: $(grep -c 'rc=0 ' $LOG) of $(wc -l < /tmp/ids.txt)" | tee -a $LOG
