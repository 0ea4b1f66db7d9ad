#!/bin/bash
# where a high-cardinality GROUP BY query (BASELINE configs[2] shape, 1.8 M groups) spends its time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CQG_TIMING=1 timeout 600 python tools/run_plan.py group_high_card 1e10 4 > gpurun_out/hc_timing.log 2>&1
echo rc=$?
tail -60 gpurun_out/hc_timing.log
