#!/bin/bash
# tools/micro_sweep.sh -- resident step time of the headline workload (1024 images) against MARS_MICRO_BATCH / MARS_MICRO_OPS
# (layer-group micro-batching, runtime.cu enqueue_run).  usage: bash tools/micro_sweep.sh > profiles/rNN_micro_batch_sweep.txt
echo "MARS_MICRO_BATCH x MARS_MICRO_OPS sweep, batch 1024, tools/one_step.py (3 steps; ms/step of the graph-replayed steps)"
for cfg in "0 -" "32 -" "64 -" "128 -" "256 -" "512 -" "64 22" "128 22" "64 0" "128 0"; do
  set -- $cfg
  if [ "$2" = "-" ]; then unset MARS_MICRO_OPS; else export MARS_MICRO_OPS=$2; fi
  echo -n "micro_batch=$1 micro_ops=${2}: "
  MARS_MICRO_BATCH=$1 python tools/one_step.py 1024 4 2>&1 | tail -1
done
