#!/bin/bash
# Run on the GPU box (gpurun): plain bench, then the ncu launch list of the same command, then one
# full capture of the headline kernel. Outputs land in gpurun_out/; summaries are made by
# tools/summarise_profiles.py on the CPU box and committed under profiles/.
set -u
mkdir -p gpurun_out
R=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-extra"
$CMD > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
CQ_BENCH_BYTES=2e9 $CMD > /dev/null 2>&1 &&
CQ_BENCH_BYTES=2e9 ncu --set full --clock-control none --import-source on -k regex:lean2_kernel -s 2 -c 1 -o gpurun_out/${R}_lean_full $CMD > gpurun_out/${R}_ncu_full.log 2>&1
echo "full capture rc=$?"
