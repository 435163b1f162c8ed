#!/bin/bash
# usage: run_gpu_round_n.sh <gpus> <logfile> <timeout> <command...>   — gpurun --gpus N with retries while the pod is busy
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $log; then sleep 60; continue; fi
  break
done
echo done >> $log
