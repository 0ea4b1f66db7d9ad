#!/bin/bash
# lean2k_kernel on the GPU box: parity first, then A/B against lean2g on the 10 GB bench shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plan_parity.py -x -q -k "few_groups or plan_parity or blanks or hand_over" > gpurun_out/l2k_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/l2k_tests.log
for m in 7 8; do
  echo "== CQG_L2K_MINB=$m"
  CQG_L2K_MINB=$m timeout 300 python tools/run_plan.py group_name 1e10 4 2>&1 | tail -2
done
