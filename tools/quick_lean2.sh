#!/bin/bash
# A/B of the scalar lean kernels on a generated table
# CQG_LEAN2: 0 round-1 kernel, 1 lean2 (dp4a masks), 3 lean2 with IMAD.HI masks; CQG_L2_VAR: geometry variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "LEAN2=$1 VAR=$2 $3"; CQG_LEAN2=$1 CQG_L2_VAR=$2 timeout 300 python tools/run_plan.py $3 ${4:-2e9} 4 2>&1 | tail -2; }
run 1 1 count_age_gt_40 1e10
run 1 6 count_age_gt_40 1e10
run 1 2 count_age_gt_40 1e10
run 1 6 count_age_gt_40 1e10
run 1 1 count_age_gt_40 1e10
run 1 1 count_height_gt_1_5 1e10
