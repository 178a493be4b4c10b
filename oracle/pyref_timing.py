"""TEST / BENCH INFRASTRUCTURE — times the UNMODIFIED Python reference's own step on the host cores.

This is the "reference's Python CPU step timed on the box's own host cores" that BASELINE.json's north_star asks to be
reported next to the GPU numbers.  The reference package is read from ``baseline/_ref`` (an install of
/root/reference made by ``__graft_entry__.build()`` in the build container; git-ignored, shipped to the GPU box) or, when
that is absent, from ``$ZOMBSOLE_REFERENCE`` / ``/root/reference``.  ``gymnasium`` / ``termcolor`` are not installed in
this image: the stand-ins under ``oracle/shims`` (no arithmetic on the path) are used when the real ones are missing.

Worker (one process, one core):
    python oracle/pyref_timing.py <mode> <seconds> <seed>
  mode "env":   ZombsoleGymEnvDiscreteAction(config 1).step(uniform random action), auto-reset on terminated|truncated,
                resets included (zombsole/gym_env.py:99-164, 327-379)
  mode "world": the same game without observation encoding and reward tracking: Agent.set_action + World.step() + the
                rules' end test, re-initialised when the game ends (zombsole/core.py:72-78, players/agent.py:22-25,
                rules/extermination.py:17-26) — the encoder is 83 % of the reference's step time
prints one JSON line {"steps": n, "seconds": s, "resets": r}.

``measure(seconds, procs)`` starts the workers as plain subprocesses (nothing here imports torch or CUDA).
"""
import json
import os
import random
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIMS = os.path.join(HERE, "shims")
ENV_ARGS = ("extermination", ["terminator", "terminator"], "bridge", 0)  # BASELINE.json configs[0]
ENV_KW = dict(initial_zombies=10, minimum_zombies=0)


def reference_root():
    """Directory that holds the reference's ``zombsole`` package, or None."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("ZOMBSOLE_REFERENCE"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "zombsole", "gym_env.py")):
            return cand
    return None


def _import_reference():
    root = reference_root()
    if root is None:
        raise RuntimeError("the Python reference is not available (baseline/_ref missing)")
    for name in ("gymnasium", "termcolor"):
        try:
            __import__(name)
        except ImportError:
            if SHIMS not in sys.path:
                sys.path.insert(0, SHIMS)
    if root not in sys.path:
        sys.path.insert(0, root)
    import zombsole.gym_env as ge
    return ge


def worker(mode, seconds, seed):
    ge = _import_reference()
    random.seed(seed)
    env = ge.ZombsoleGymEnvDiscreteAction(*ENV_ARGS, **ENV_KW)
    env.reset()
    n_actions = len(env.game_actions)
    steps = resets = 0
    if mode == "env":
        def one():
            _, _, terminated, truncated, _ = env.step(random.randrange(n_actions))
            if terminated or truncated:
                env.reset()
                return 1
            return 0
    else:
        inner = env.env  # the ZombsoleGymEnv behind the discrete wrapper
        game = inner.game

        def one():
            game.agents[0].set_action(env.game_actions[random.randrange(n_actions)])
            game.world.step()
            if game.rules.game_ended() or not game.rules.agents_alive():
                game.__initialize_world__()
                return 1
            return 0
    t_end = time.perf_counter() + min(1.0, 0.25 * seconds)  # warm-up
    while time.perf_counter() < t_end:
        one()
    t0 = time.perf_counter()
    t_end = t0 + seconds
    while True:
        for _ in range(32):
            resets += one()
        steps += 32
        now = time.perf_counter()
        if now >= t_end:
            break
    print(json.dumps({"steps": steps, "seconds": now - t0, "resets": resets}), flush=True)


def _run(mode, seconds, procs):
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), mode, str(seconds), str(1000 + i)],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT)
          for i in range(procs)]
    total, secs, ok = 0.0, [], 0
    for p in ps:
        out, err = p.communicate(timeout=seconds * 4 + 120)
        try:
            d = json.loads(out.strip().splitlines()[-1])
        except Exception:
            continue
        total += d["steps"] / d["seconds"]
        secs.append(d["seconds"])
        ok += 1
    return {"value": total, "processes": ok, "seconds_each": (sum(secs) / len(secs)) if secs else 0.0}


def measure(seconds=3.0, procs=None):
    """-> dict with the three BASELINE.md §4 figures, or {"unavailable": why}."""
    if reference_root() is None:
        return {"unavailable": "baseline/_ref (an install of the Python reference) is not on this box"}
    try:
        procs = procs or max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        procs = procs or (os.cpu_count() or 1)
    one = _run("env", seconds, 1)
    if one["processes"] == 0:
        return {"unavailable": "the reference worker failed to start"}
    many = _run("env", seconds, procs)
    world = _run("world", seconds, 1)
    return {
        "unit": "env-steps/s", "config": "BASELINE configs[0]: ZombsoleGymEnvDiscreteAction, bridge, extermination, 10 zombies, "
                                         "agent + 2 terminators, uniform random discrete actions, resets included",
        "one_process_one_core": one["value"], "all_cores_summed": many["value"], "cores": many["processes"],
        "world_step_only_one_core": world["value"], "seconds_each": seconds,
        "source": os.path.relpath(reference_root(), ROOT) if reference_root().startswith(ROOT) else reference_root(),
        "python": sys.version.split()[0],
    }


if __name__ == "__main__":
    worker(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]))
