#!/bin/bash
# quick look at the scalar lean kernel on a generated table (CQG_LEAN2=0: the first lean kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "LEAN2=$1 $2"; CQG_LEAN2=$1 timeout 300 python tools/run_plan.py $2 ${3:-2e9} 4 2>&1 | tail -3; }
run 1 count_age_gt_40 1e10
run 1 count_height_gt_1_5 1e10
