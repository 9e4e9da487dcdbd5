#!/bin/bash
# First-contact probe on a B200 box: each test file in its own process so a trapped kernel cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/probe_summary.txt
  tail -5 gpurun_out/$name.log
done
