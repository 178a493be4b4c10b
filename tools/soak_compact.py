"""Soak of the host-facing step (not part of the test suite): host_outputs="compact" against a plain env, step by step over
many episodes, in the expansion modes of zs_step_host (ZS_HOST_DIFF=0 restore ahead + write, =1 difference of the records,
unset: chosen by timing, switching while it runs).

    python tools/soak_compact.py [steps] [envs]
"""
import os
import sys
sys.path.insert(0, ".")
import numpy as np
import torch

KW = dict(rules_name="extermination", player_names=["terminator", "terminator"], map_name="bridge", agent_id=0,
          initial_zombies=10, minimum_zombies=0, observation_scope="world", agent_weapon="rifle")


def run(diff, enc, N, T):
    if diff < 0:
        os.environ.pop("ZS_HOST_DIFF", None)  # the handle times both modes and switches between them
    else:
        os.environ["ZS_HOST_DIFF"] = str(diff)
    from libzombsole_b200.gym_env import ZombsoleVectorEnv
    plain = ZombsoleVectorEnv(num_envs=N, seed=21, max_episode_steps=45, observation_position_encoding=enc, **KW)
    comp = ZombsoleVectorEnv(num_envs=N, seed=21, max_episode_steps=45, observation_position_encoding=enc, host_outputs="compact", **KW)
    plain.reset(); comp.reset()
    rs = np.random.RandomState(4)
    bad = 0
    for t in range(T):
        a = torch.from_numpy(rs.randint(0, 6, size=N).astype(np.int32))
        o, r, te, tr, _ = plain.step(a.cuda())
        co, cr, cte, ctr, _ = comp.step(a)
        if not (torch.equal(o.cpu(), co) and torch.equal(r.cpu().view(torch.int64), cr.view(torch.int64))
                and torch.equal(te.cpu(), cte) and torch.equal(tr.cpu(), ctr)):
            bad += 1
            if bad < 4:
                d = (o.cpu() != co).reshape(N, -1).any(1).nonzero().flatten().tolist()
                print("  step %d differs in envs %s" % (t, d[:8]))
    print("diff=%d %-8s N=%d: %d steps, %d differing steps, %d rows fetched in full" % (diff, enc, N, T, bad, comp.compact_overflows), flush=True)
    plain.close(); comp.close()
    return bad


if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    total = 0
    for diff in (-1, 1, 0):
        for enc in ("simple", "channels"):
            total += run(diff, enc, N, T)
    print("soak_compact: %d differing steps" % total)
    sys.exit(1 if total else 0)
