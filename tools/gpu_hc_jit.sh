#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CQG_JIT_CACHE=/tmp/jitc_$$ CQG_JIT_VERBOSE=1
timeout 600 python tools/run_plan.py group_high_card 2e9 3 > gpurun_out/hc_jit.log 2>&1
echo rc=$?
tail -40 gpurun_out/hc_jit.log
