"""Is the single-step rate bound by the host or by the GPU?  (not the bench)
Back-to-back zs_step launches of 4,096 envs issued three ways: through ZsEngine.step (argument checks every call), as raw
ctypes calls with the pointers computed once, and replayed from a CUDA graph that captured 200 of them (no host in the loop:
the GPU's own time per single-step launch).

    python tools/probe_step_rate.py [config] [N]
"""
import sys
import time
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
    ring = max(2, -(-2 * 126 * (1 << 20) // (eng.obs_elems * 4 * N)))
    obs = eng.new_obs(ring)
    rew, term, trunc = eng.new_outputs(1)
    acts = torch.zeros((512, N, eng.A), dtype=torch.int32, device=eng.device)
    for s in range(512):
        eng.fill_synthetic_actions(s, acts[s])
    eng.rollout(200, 0, acts, abi.ACTIONS_DISCRETE, obs, rew.expand(1, N) if rew.dim() == 1 else rew, term, trunc) if False else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 4000

    def timed(fn, label):
        for s in range(50):
            fn(s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev[0].record()
        for s in range(n):
            fn(s)
        t_issue = time.perf_counter() - t0
        ev[1].record()
        torch.cuda.synchronize()
        print("%-46s %6.2f us per launch on the device clock, %6.2f us of host time to issue one" % (
            label, ev[0].elapsed_time(ev[1]) / n * 1e3, t_issue / n * 1e6), flush=True)

    timed(lambda s: eng.step(acts[s % 512], abi.ACTIONS_DISCRETE, obs[s % ring], rew[0], term[0], trunc[0]), "ZsEngine.step")
    L, h = eng.L, eng.h
    stream = torch.cuda.current_stream(eng.device).cuda_stream
    ap = [acts[s].data_ptr() for s in range(512)]
    op = [obs[s].data_ptr() for s in range(ring)]
    rp, tp, up = rew.data_ptr(), term.data_ptr(), trunc.data_ptr()
    zs_step = L.zs_step
    timed(lambda s: zs_step(h, ap[s % 512], abi.ACTIONS_DISCRETE, op[s % ring], rp, tp, up, None, None, stream), "raw ctypes zs_step, pointers computed once")
    # the GPU's own time: 200 launches captured into a graph, replayed
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=eng.device)
    with torch.cuda.stream(side):
        for s in range(10):
            zs_step(h, ap[s], abi.ACTIONS_DISCRETE, op[s % ring], rp, tp, up, None, None, side.cuda_stream)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for s in range(200):
                zs_step(h, ap[s % 512], abi.ACTIONS_DISCRETE, op[s % ring], rp, tp, up, None, None, side.cuda_stream)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(20):
        g.replay()
    ev[1].record()
    torch.cuda.synchronize()
    print("%-46s %6.2f us per launch on the device clock" % ("CUDA graph of 200 zs_step launches, replayed", ev[0].elapsed_time(ev[1]) / 4000 * 1e3))
    eng.close()


if __name__ == "__main__":
    main()
