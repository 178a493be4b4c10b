#!/bin/bash
# usage (GPU box): tools/profile_others.sh <tag>  -> text summaries (tools/ncu_summary.sh) of one ncu --set full capture each:
# config 3 and config 4 fused rollouts, the encode kernel and the reset kernel at 262,144 bridge envs
tag=$1
cd "$(dirname "$0")/.."
tools/profile_config.sh c3$tag c3_city_evac 32768 20 > /dev/null
tools/ncu_summary.sh gpurun_out/prof_c3$tag.ncu-rep 30 > gpurun_out/summary_c3_$tag.txt 2>&1; rm -f gpurun_out/prof_c3$tag.ncu-rep
tools/profile_config.sh c4$tag c4_maze_safehouse 32768 20 > /dev/null
tools/ncu_summary.sh gpurun_out/prof_c4$tag.ncu-rep 30 > gpurun_out/summary_c4_$tag.txt 2>&1; rm -f gpurun_out/prof_c4$tag.ncu-rep
# zs_sim_kernel launches of probe_kernels.run: reset (0), 30-step rollout (1), 3 + 10 encodes (2..14), 3 + 10 resets (15..27)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:zs_sim_kernel -s 3 -c 1 -o gpurun_out/prof_enc$tag -f python tools/probe_kernels.py c1_bridge_ext 262144 > /dev/null 2>&1
tools/ncu_summary.sh gpurun_out/prof_enc$tag.ncu-rep 30 > gpurun_out/summary_encode_$tag.txt 2>&1; rm -f gpurun_out/prof_enc$tag.ncu-rep
timeout 300 ncu --set full --import-source on --clock-control none -k regex:zs_sim_kernel -s 16 -c 1 -o gpurun_out/prof_rst$tag -f python tools/probe_kernels.py c1_bridge_ext 262144 > /dev/null 2>&1
tools/ncu_summary.sh gpurun_out/prof_rst$tag.ncu-rep 30 > gpurun_out/summary_reset_$tag.txt 2>&1; rm -f gpurun_out/prof_rst$tag.ncu-rep
head -3 gpurun_out/summary_*_$tag.txt
