"""Launch-cost probe (not the bench): time of a K-step zs_rollout launch as a function of K, launches issued back to
back until >= 40 ms are inside one CUDA-event pair.  The intercept of the line is the per-launch fixed cost (prologue,
epilogue, ramp), the slope the steady-state cost of a fused step.

    python tools/probe_launch_cost.py [config] [N]
"""
import os
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c1_bridge_ext"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    cfg, m = pu.build(pu.CONFIGS[name], N, 0, auto_reset=True, max_episode_steps=1000)
    eng = ZsEngine(cfg, m)
    obs_bytes = eng.obs_elems * 4 * N
    ring = max(2, -(-2 * 126 * (1 << 20) // obs_bytes))
    obs = eng.new_obs(ring)
    KMAX = 512
    rew, term, trunc = eng.new_outputs(KMAX)
    acts = torch.zeros((KMAX, N, eng.A), dtype=torch.int32, device=eng.device)
    for s in range(KMAX):
        eng.fill_synthetic_actions(s, acts[s])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    eng.rollout(200, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)  # desynchronise the episodes
    torch.cuda.synchronize()
    rows = []
    ks = [int(k) for k in os.environ.get("PROBE_KS", "1,2,3,4,5,8,12,16,20,32,64,128,512").split(",")]
    for K in ks:
        est = 0.03 + 0.006 * K  # ms per launch, rough
        R = max(3, int(40.0 / est))
        for _ in range(3):
            eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(R):
            eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / R
        # one launch alone on an idle GPU
        ev[0].record()
        eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
        ev[1].record()
        torch.cuda.synchronize()
        alone = ev[0].elapsed_time(ev[1])
        rows.append((K, ms, alone))
        print("K=%4d  back-to-back %9.2f us/launch  %7.2f us/step  %.3e env-steps/s   alone %9.2f us" % (
            K, ms * 1e3, ms * 1e3 / K, N * K / ms * 1e3, alone * 1e3), flush=True)
    # the same for single zs_step launches
    o1 = obs[0]
    for _ in range(20):
        eng.step(acts[0], abi.ACTIONS_DISCRETE, o1, rew[0], term[0], trunc[0])
    torch.cuda.synchronize()
    ev[0].record()
    for s in range(2000):
        eng.step(acts[s % KMAX], abi.ACTIONS_DISCRETE, obs[s % ring], rew[0], term[0], trunc[0])
    ev[1].record()
    torch.cuda.synchronize()
    print("zs_step back-to-back: %.2f us/launch" % (ev[0].elapsed_time(ev[1]) / 2000 * 1e3))
    if len(rows) >= 9:
        (k0, m0, _), (k1, m1, _) = rows[8], rows[-1]
        slope = (m1 - m0) / (k1 - k0)
        print("slope %.2f us/step, intercept at K=20: %.2f us" % (slope * 1e3, (m0 - slope * k0) * 1e3))
    eng.close()


if __name__ == "__main__":
    main()
