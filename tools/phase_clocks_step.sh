#!/bin/bash
# usage (GPU box): tools/phase_clocks_step.sh  -> per-phase cycles of warp 0 in SINGLE-STEP launches (development build)
cd "$(dirname "$0")/.."
cp libzombsole_b200/csrc/libzs_b200.so /tmp/libzs_b200.keep
(cd libzombsole_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden --fmad=false -cudart static -DZS_PHASE_CLOCKS -o libzs_b200.so zs_b200.cu 2>/dev/null)
cat > /tmp/_pcs.py <<PY
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine
cfg, m = pu.build(pu.CONFIGS["c1_bridge_ext"], 4096, 0, auto_reset=True, max_episode_steps=1000)
eng = ZsEngine(cfg, m)
obs = eng.new_obs(); rew, term, trunc = eng.new_outputs()
acts = torch.zeros((4096, 1), dtype=torch.int32, device=eng.device)
for s in range(60):
    eng.fill_synthetic_actions(s, acts)
    eng.step(acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
torch.cuda.synchronize()
PY
timeout 120 python /tmp/_pcs.py 2>&1 | grep "phase cycles" | tail -4
cp /tmp/libzs_b200.keep libzombsole_b200/csrc/libzs_b200.so
