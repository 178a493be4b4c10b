#!/bin/bash
# usage (GPU box): tools/profile_config.sh <tag> <config name in tests/parity_util.CONFIGS> <envs> <steps>
#   one ncu --set full capture of the timed fused rollout of that configuration -> gpurun_out/prof_<tag>.ncu-rep
tag=$1; name=$2; N=$3; K=$4
cd "$(dirname "$0")/.."
cat > /tmp/_run_cfg.py <<PY
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import probe_large as pl
pl.run("$name", $N, $K)
PY
python /tmp/_run_cfg.py || exit 1
# zs_sim_kernel launches of probe_large.run: the constructor's reset (0), the 2-step warm-up (1), the timed rollout (2)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:zs_sim_kernel -s 2 -c 1 -o gpurun_out/prof_$tag -f \
  python /tmp/_run_cfg.py > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
