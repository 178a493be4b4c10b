#!/bin/bash
# usage (GPU box): tools/sweep_final.sh  -> K=1 / 20 / 512 launch times of configs[1] under the tuning switches, final build
cd "$(dirname "$0")/.."
export PROBE_KS=1,20,512
run() { echo "== $*"; env "$@" python tools/probe_launch_cost.py c1_bridge_ext 4096 2>&1 | grep -E "^K=|zs_step"; }
run ZS_NONE=1
run ZS_WARPS_PER_CTA=1
run ZS_WARPS_PER_CTA=4
run ZS_OCC=6
run ZS_OCC=7
run ZS_TMA_PAIR=1
run ZS_NO_TMA=1
run ZS_SHORT_STEPS=4
run ZS_SHORT_STEPS=16
run ZS_SMEM_SKEW=0
