"""BASELINE configs 3, 4 and 5 at their stated GPU counts (the driver's bench.py line is configs[1]):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_configs.py <config> <total_envs> <steps> [tape]
config = c3_city_evac | c4_maze_safehouse | c1_bridge_ext | c5_bridge_channels (tests/parity_util.CONFIGS).
Envs shard by contiguous global index range, one batch per GPU, no collective on the step path; one fused
zs_rollout launch per rank, timed with CUDA events behind a barrier, max over ranks; uniform random discrete actions,
generated on the device inside the step kernel, or with `tape` read from an action tensor filled beforehand (as bench.py).  Prints one JSON line with env-steps/s over all GPUs and the fraction of the HBM roofline
(SURVEY §8d algorithmic bytes per env-step, MEASURED_PEAKS.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.distributed import shard_envs, all_reduce_stats
from libzombsole_b200.engine import ZsEngine


def main():
    name, total, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    tape = len(sys.argv) > 4 and sys.argv[4] == "tape"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    base, N = shard_envs(total, rank, world)
    cfg, m = pu.build(pu.CONFIGS[name], N, 0, env_index_base=base, auto_reset=True, max_episode_steps=1000)
    eng = ZsEngine(cfg, m)
    obs = eng.new_obs()
    eng.rollout(3, 0, None, abi.ACTIONS_DISCRETE, obs, None, None, None)
    eng.episode_stats(reset=True)
    acts = None
    if tape:
        acts = torch.zeros((K, N, eng.A), dtype=torch.int32, device=eng.device)
        for s in range(K):
            eng.fill_synthetic_actions(3 + s, acts[s])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    eng.rollout(K, 3, acts, abi.ACTIONS_DISCRETE, obs, None, None, None)
    ev[1].record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    stats = all_reduce_stats(eng.episode_stats())
    if rank == 0:
        M, S, cells, A = eng.M, eng.S, eng.cells, eng.A
        b_alg = (12 * M + 2 * S + cells // 4 + 24) + (12 * M + 24) + 4 * A + eng.obs_elems * 4 + 8 * A + 2
        peak = 6538.3
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f).get("hbm_gbs", peak))
        except Exception:
            pass
        rate = total * K / (ms.item() * 1e-3)
        print(json.dumps({"config": name, "total_envs": total, "n_gpus": world, "envs_per_gpu": N, "steps": K,
                          "ms_per_step": ms.item() / K, "env_steps_per_sec": rate, "algorithmic_bytes_per_env_step": b_alg,
                          "roofline_frac": rate * b_alg / (world * peak * 1e9),
                          "episodes": {"finished": int(stats[0]), "won": int(stats[1]), "steps": int(stats[2])},
                          "actions": "uniform random discrete (Philox), " + ("action tape resident in HBM" if tape else "generated inside the step kernel"), "launch": "one fused zs_rollout per rank"}), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
