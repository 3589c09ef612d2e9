#!/bin/bash
# Runs groups of GPU tests, each group in its own process under a timeout, so a hung or faulting kernel
# cannot take the whole run down. Usage: scripts/gpu_check.sh "<pytest args>" "<pytest args>" ...
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
i=0
for t in "$@"; do
  i=$((i+1))
  echo "=== [$i] $t" | tee -a gpurun_out/summary.txt
  timeout 900 python -m pytest $t -q -m gpu --no-header -p no:cacheprovider > "gpurun_out/group$i.log" 2>&1
  rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/group$i.log)" | tee -a gpurun_out/summary.txt
  grep -E "^(FAILED|ERROR)|Error|error:|assert" "gpurun_out/group$i.log" | head -n 30
done
