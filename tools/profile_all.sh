#!/bin/bash
# Round profile refresh, run on the GPU box through gpurun:  bash tools/profile_all.sh <tag>
# For every workload: a plain bench run (must exit 0), then the ncu launch list of the same command, then one
# `--set full` capture of the hot kernels.  Everything lands in gpurun_out/ and is summarised into profiles/ by
# tools/ncu_summary.py in the build container.
tag=${1:-r01}
out=gpurun_out
for w in tron ttt4 blokus; do
  steps=40; [ $w = blokus ] && steps=24
  python bench.py --workload $w --steps $steps --warmup 3 --no-cpu > $out/plain_${w}_$tag.json 2> $out/plain_${w}_$tag.err || { echo "plain $w failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_${w}_$tag.csv \
      python bench.py --workload $w --steps $steps --warmup 3 --no-cpu > $out/ncu_launches_${w}_$tag.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:tron_step -s 60 -c 1 -o $out/prof_tron_step_$tag -f \
    python bench.py --workload tron --steps 40 --warmup 3 --no-cpu > $out/ncu_full_tron_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttt_rollout -s 40 -c 1 -o $out/prof_ttt_rollout_$tag -f \
    python bench.py --workload ttt4 --steps 40 --warmup 3 --no-cpu > $out/ncu_full_ttt_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blokus_step -s 20 -c 1 -o $out/prof_blokus_step_$tag -f \
    python bench.py --workload blokus --steps 24 --warmup 3 --no-cpu > $out/ncu_full_blokus_step_$tag.log 2>&1
ls -la $out/*_$tag* | awk '{print $5, $9}'
