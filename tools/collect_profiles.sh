#!/bin/bash
# Run on the GPU box (gpurun): plain bench, then the ncu launch list of the same command, then one full capture per
# headline kernel (tools/run_plan.py on a 2 GB table). Outputs land in gpurun_out/; summaries are made by
# tools/summarise_profiles.py on the CPU box and committed under profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r02}
python bench.py > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/${R}_bench_plain.err; exit 1; }
echo "plain bench ok"
CMD="python bench.py --steps 2 --warmup 3 --no-extra"
$CMD > gpurun_out/${R}_bench_short.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for spec in "group_name lean2k lean2k" "count_age_gt_40 lean2_kernel lean2" "group_high_card leanhc leanhc"; do
  set -- $spec
  timeout 300 python tools/run_plan.py $1 2e9 3 > gpurun_out/${R}_$3_plain.log 2>&1 || { echo "$1 plain failed"; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/${R}_$3 python tools/run_plan.py $1 2e9 3 > gpurun_out/${R}_$3_ncu.log 2>&1
  echo "$3 capture rc=$?"
done
