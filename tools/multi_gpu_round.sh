#!/bin/bash
# usage (on an 8-GPU box): tools/multi_gpu_round.sh <tag>  -> gpurun_out/bench_<tag>_8gpu.json, gpurun_out/configs_<tag>_multi_gpu.jsonl
tag=$1
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --e2e-steps 100 > gpurun_out/bench_${tag}_8gpu.json 2> gpurun_out/bench_${tag}_8gpu.err
out=gpurun_out/configs_${tag}_multi_gpu.jsonl
: > $out
$TR --nproc-per-node 8 --master-port 29512 tools/bench_configs.py c4_maze_safehouse 1048576 20 2>/dev/null | grep '^{' >> $out
$TR --nproc-per-node 8 --master-port 29513 tools/bench_configs.py c1_bridge_ext 8388608 50 tape 2>/dev/null | grep '^{' >> $out
$TR --nproc-per-node 8 --master-port 29514 tools/bench_configs.py c5_bridge_channels 8388608 30 tape 2>/dev/null | grep '^{' >> $out
$TR --nproc-per-node 2 --master-port 29515 tools/bench_configs.py c3_city_evac 65536 40 2>/dev/null | grep '^{' >> $out
tail -c 400 gpurun_out/bench_${tag}_8gpu.json; cut -c1-200 $out
