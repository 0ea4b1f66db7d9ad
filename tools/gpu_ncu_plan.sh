#!/bin/bash
# one ncu --set full capture of a named kernel while tools/run_plan.py runs a plan on a 2 GB table
# usage: tools/gpu_ncu_plan.sh <plan> <kernel regex> <out name>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PLAN=$1; KREG=$2; OUT=$3
timeout 300 python tools/run_plan.py $PLAN 2e9 3 > gpurun_out/${OUT}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${OUT}_plain.log; exit 1; }
cat gpurun_out/${OUT}_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREG -s 1 -c 1 -f -o gpurun_out/$OUT python tools/run_plan.py $PLAN 2e9 3 > gpurun_out/${OUT}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${OUT}_ncu.log
