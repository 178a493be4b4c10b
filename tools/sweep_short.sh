#!/bin/bash
# usage (GPU box): tools/sweep_short.sh -> launch time vs K with one env per warp (short shape) and two (rollout shape)
cd "$(dirname "$0")/.."
export PROBE_KS=1,2,4,6,8,10,12,16,20
run() { echo "== $*"; env "$@" python tools/probe_launch_cost.py c1_bridge_ext 4096 2>&1 | grep -E "^K=|zs_step"; }
run ZS_SHORT_STEPS=1000
run ZS_SHORT_STEPS=0
