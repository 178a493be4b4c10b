#!/usr/bin/env python
"""bench.py — env-steps/s of the batched zombsole hot path (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port, all host threads)
    (N > 1: launched by torch.distributed.run, one rank per GPU)

Workload (config.workload): BASELINE.json configs[1] — 4,096 batched bridge/extermination envs per
GPU, 10 zombies, agent (rifle) + 2 terminator bots, world/simple observation (1,12,111) int32,
uniformly random discrete actions, same-step auto-reset.  A "step" is one transition of the whole
batch (4,096 env-steps per GPU); envs shard over GPUs by global env index with no collective on
the step path (weak scaling).

  value      device-timed (CUDA events, max over ranks) throughput of K steps with the action tape
             [K, N] already resident in HBM, run as ONE fused launch (zs_rollout); observations go to
             a ring of obs buffers larger than L2 so every step's stores miss L2.
  per_step   the same K steps as K zs_step launches (one launch per step).
  e2e        the same metric through the public API ZombsoleVectorEnv.step() with HOST buffers:
             every step copies its actions from pinned host memory and reads observation, reward
             and flags back to pinned host memory inside the timed region.
  roofline   HBM roofline of the fused step kernel: algorithmic bytes (SURVEY.md 8d, 6,431 B per
             env-step for this config) / measured duration vs MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline  the C oracle (a port of the reference's Python path) on the box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
B_ALG = 6431  # algorithmic bytes per env-step, bridge / 13 mobile things / world-simple obs
WORKLOAD = ("BASELINE configs[1]: 4096 batched bridge/extermination envs per GPU, 10 zombies, agent(rifle) + 2 "
            "terminators, world/simple obs (1,12,111) int32, uniform random discrete actions, same-step auto-reset")
ENV_KW = dict(rules_name="extermination", player_names=["terminator", "terminator"], map_name="bridge", agent_id=0,
              initial_zombies=10, minimum_zombies=0, observation_scope="world",
              observation_position_encoding="simple", agent_weapon="rifle")
FALLBACK_HBM_GBS = 6650.0


def ncu_traffic_per_env_step():
    """DRAM bytes per env-step of the step kernel from the committed ncu --set full capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_env_step"]), d["capture"]
    except Exception:
        return None, None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_env(n_envs, seed=0, base=0):
    from libzombsole_b200 import abi
    from oracle import oracle as orc
    cfg = abi.make_config(ENV_KW["rules_name"], ENV_KW["player_names"], [ENV_KW["agent_id"]], ENV_KW["agent_weapon"],
                          ENV_KW["initial_zombies"], ENV_KW["minimum_zombies"], abi.OBS_WORLD, abi.OBS_SIMPLE, 0, False,
                          n_envs, seed=seed, env_index_base=base, max_episode_steps=1000, auto_reset=True)
    return orc.OracleEnv(cfg, ENV_KW["map_name"])


def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1; ignore it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(budget_s=12.0):
    """The oracle port on all host threads over a bounded sample of the same workload."""
    from oracle import oracle as orc
    threads = orc.set_threads(host_threads())
    env = oracle_env(ENVS_PER_GPU)
    env.rollout_synthetic(2, 0)  # warm-up
    t0 = time.perf_counter()
    env.rollout_synthetic(4, 2)
    per_step = (time.perf_counter() - t0) / 4
    steps = max(8, min(2000, int(budget_s / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    env.rollout_synthetic(steps, 6)
    dt = time.perf_counter() - t0
    multi = ENVS_PER_GPU * steps / dt
    env.close()
    orc.set_threads(1)
    env1 = oracle_env(256)
    env1.rollout_synthetic(2, 0)
    s1 = max(8, int(0.25 * steps))
    t0 = time.perf_counter()
    env1.rollout_synthetic(s1, 2)
    single = 256 * s1 / (time.perf_counter() - t0)
    env1.close()
    orc.set_threads(threads)
    return {"value": multi, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "single_core_value": single,
            "sample": "%d envs x %d steps of the same workload (%.1f s), C oracle with OpenMP over envs; "
                      "single_core_value: 256 envs x %d steps on 1 thread" % (ENVS_PER_GPU, steps, dt, s1)}


def run_reference(args, rank, world):
    """--impl reference: the reference's path on the host cores (the oracle port; the reference itself is
    Python and does not travel to the GPU box).  Each step is one transition of a bounded batch."""
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = orc.set_threads(host_threads())
    n = ENVS_PER_GPU
    env = oracle_env(n)
    env.rollout_synthetic(1, 0)
    t0 = time.perf_counter()
    env.rollout_synthetic(2, 1)
    per_step = (time.perf_counter() - t0) / 2
    total = args.steps + args.warmup
    if per_step * total > 150.0:  # keep the whole run within a few minutes
        n = max(64, int(n * 150.0 / (per_step * total)) // 64 * 64)
        env.close()
        env = oracle_env(n)
    env.rollout_synthetic(args.warmup, 0)
    t0 = time.perf_counter()
    env.rollout_synthetic(args.steps, args.warmup)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    env.close()
    sample = "%d envs per step x %d steps on %d host threads (C oracle port of the Python reference)" % (n, args.steps, threads)
    print(json.dumps({
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_step": n},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from libzombsole_b200 import abi
    from libzombsole_b200.gym_env import ZombsoleVectorEnv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N, K, W = ENVS_PER_GPU, args.steps, args.warmup
    env = ZombsoleVectorEnv(num_envs=N, device=dev, seed=args.seed, env_index_base=rank * N, max_episode_steps=1000,
                            auto_reset=True, **ENV_KW)
    eng = env.engine
    obs_bytes = eng.obs_elems * 4 * N
    ring = max(2, -(-2 * 126 * (1 << 20) // obs_bytes))  # obs ring >= 2 x L2 (126 MB)
    obs_ring = eng.new_obs(ring)
    reward, term, trunc = eng.new_outputs(K)
    tape = torch.empty((W + K, N, 1), dtype=torch.int32, device=dev)
    for s in range(W + K):
        eng.fill_synthetic_actions(s, tape[s])
    torch.cuda.synchronize(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sampler = ClockSampler(local_rank)

    # ---------------- fused rollout: `value`
    eng.rollout(W, 0, tape[:W], abi.ACTIONS_DISCRETE, obs_ring, None, None, None)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    barrier()
    ev[0].record()
    eng.rollout(K, W, tape[W:], abi.ACTIONS_DISCRETE, obs_ring, reward, term, trunc)
    ev[1].record()
    barrier()
    launches = eng.launch_count() - launches0
    fused_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))

    # ---------------- one launch per step
    for s in range(W):
        eng.step(tape[s], abi.ACTIONS_DISCRETE, obs_ring[s % ring], reward[0], term[0], trunc[0])
    barrier()
    ev[0].record()
    for s in range(K):
        eng.step(tape[W + s], abi.ACTIONS_DISCRETE, obs_ring[s % ring], reward[s], term[s], trunc[s])
    ev[1].record()
    barrier()
    per_step_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))

    # ---------------- end to end through the public API with host buffers
    Ke = min(K, args.e2e_steps)
    h_actions = torch.empty((Ke + W, N), dtype=torch.int32).pin_memory()
    h_actions.copy_(tape[:Ke + W, :, 0])
    h_obs = torch.empty((N,) + eng.obs_shape, dtype=torch.int32).pin_memory()
    h_rew = torch.empty(N, dtype=torch.float64).pin_memory()
    h_term = torch.empty(N, dtype=torch.bool).pin_memory()
    h_trunc = torch.empty(N, dtype=torch.bool).pin_memory()

    def e2e_step(s):
        o, r, te, tr, _ = env.step(h_actions[s])        # H2D of this step's actions inside env.step
        h_obs.copy_(o, non_blocking=True)
        h_rew.copy_(r, non_blocking=True)
        h_term.copy_(te, non_blocking=True)
        h_trunc.copy_(tr, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()    # the caller owns the results before the next step

    for s in range(W):
        e2e_step(s)
    barrier()
    ev[0].record()
    for s in range(Ke):
        e2e_step(W + s)
    ev[1].record()
    barrier()
    e2e_copy_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))

    # the same through the env's host-output mode: step() takes the pinned host action tensor and returns pinned host
    # tensors that the kernel wrote directly (zero-copy over PCIe, overlapped with the transition) and synchronised on
    env_h = ZombsoleVectorEnv(num_envs=N, device=dev, seed=args.seed, env_index_base=rank * N, max_episode_steps=1000,
                              auto_reset=True, host_outputs=True, **ENV_KW)
    sink = 0
    for s in range(W):
        env_h.step(h_actions[s])
    barrier()
    t0 = time.perf_counter()
    ev[0].record()
    for s in range(Ke):
        o, r, te, tr, _ = env_h.step(h_actions[W + s])   # returns after the stream is idle: the host owns the results
        sink += int(o[0, 0, 0, 0]) + int(te[0])           # the host reads the step's result
    ev[1].record()
    barrier()
    e2e_host_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_host_ms = max_over_ranks(max(ev[0].elapsed_time(ev[1]), e2e_host_wall_ms))
    env_h.close()
    e2e_ms = min(e2e_copy_ms, e2e_host_ms)
    clocks = sampler.stop() if rank == 0 else None

    stats = eng.episode_stats()
    if world > 1:  # the only collective: episode statistics, off the step path
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    stats = stats.cpu().tolist()

    if rank == 0:
        peak, peak_src = hbm_peak()
        total_envs = N * world
        value = total_envs * K / (fused_ms * 1e-3)
        achieved = value * B_ALG / world / 1e9
        traffic_per, traffic_src = ncu_traffic_per_env_step()
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": fused_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": N, "total_envs": total_envs, "seed": args.seed,
                       "launch": "one fused zs_rollout launch for the K timed steps",
                       "l2": "observations are written to a ring of %d buffers (%.0f MB > 126 MB L2); the 3.3 MB "
                             "world state is L2-resident by the nature of a 4096-env batch" % (ring, ring * obs_bytes / 1e6)},
            "per_step": {"value": total_envs * K / (per_step_ms * 1e-3), "unit": "env-steps/s",
                         "ms_per_step": per_step_ms / K, "launches": K},
            "e2e": {"value": total_envs * Ke / (e2e_ms * 1e-3), "unit": "env-steps/s", "steps": Ke,
                    "h2d_bytes_per_step": N * 4, "d2h_bytes_per_step": obs_bytes + N * 8 + 2 * N,
                    "api": ("ZombsoleVectorEnv(host_outputs=True).step(pinned host actions) -> pinned host obs/reward/flags "
                            "written by the kernel over PCIe (zero-copy), stream synchronised before step() returns"
                            if e2e_host_ms <= e2e_copy_ms else
                            "ZombsoleVectorEnv.step(pinned host actions) + obs/reward/flags copied to pinned host"),
                    "copy_variant": {"value": total_envs * Ke / (e2e_copy_ms * 1e-3),
                                     "api": "ZombsoleVectorEnv.step(pinned host actions) + obs/reward/flags copied to pinned host"},
                    "host_outputs_variant": {"value": total_envs * Ke / (e2e_host_ms * 1e-3),
                                             "api": "ZombsoleVectorEnv(host_outputs=True).step(pinned host actions)"}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if traffic_per is None else traffic_per * N * K,
                         "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "zs_sim_kernel<MODE_STEP>",
                         "algorithmic_bytes_per_env_step": B_ALG,
                         "launch_ms": fused_ms, "env_steps_per_launch": N * K},
            "gpu_launches": launches,
            "clocks": clocks,
            "episodes": {"finished": stats[0], "won": stats[1], "mean_length": stats[2] / max(1, stats[0])},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION and above: stdout carries ONE JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
