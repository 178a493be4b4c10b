#!/bin/bash
# usage (on the GPU box, after `python bench.py` has exited 0 on its own): tools/profile_round.sh <tag>
#   1. launch list of the bench command (gpu__time_duration per launch)  -> gpurun_out/launches_<tag>.csv
#   2. one ncu --set full capture of the timed zs_rollout launch          -> gpurun_out/prof_<tag>.ncu-rep
tag=$1
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 20 > gpurun_out/ncu_launches_$tag.log 2>&1
# zs_sim_kernel launches of the bench: the constructor's reset (0), the warm-up rollout (1), the timed rollout (2)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:zs_sim_kernel -s 2 -c 1 -o gpurun_out/prof_$tag -f \
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 20 > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
