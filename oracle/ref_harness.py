"""TEST INFRASTRUCTURE — drives the UNMODIFIED Python reference under injected draws.

Only tests/ and tests/golden/make_golden.py import this.  It needs the reference
tree (``/root/reference`` in the build container; absent on the GPU box), so
everything that must run on the GPU box uses the committed fixtures under
tests/golden/ instead.

What it does
------------
* puts ``oracle/shims`` on sys.path when ``gymnasium``/``termcolor`` are missing
  (they do no arithmetic on the path, see oracle/shims/README.md);
* replaces ``random.shuffle/randint/choice`` by the bound methods of a
  ``random.Random`` subclass whose ``_randbelow(n)`` returns
  ``(Philox4x32-10(seed; env, episode, t_word, k) * n) >> 32`` — the draw contract
  of libzombsole_b200/philox.py.  The reference calls ``random.<fn>`` as module
  attributes (zombsole/core.py:54,76,180,198; things.py:62,103,116;
  weapons.py:43) so nothing in the reference is touched;
* steps ``ZombsoleGymEnv`` / ``MultiagentZombsoleEnv`` (zombsole/gym_env.py:99-164,
  zombsole/gym/multiagent_env.py:111-184) with an action tape and dumps, after
  every step and every reset, the slot-indexed world state, the observation,
  the float64 reward bits, the flags and the number of draws consumed.

Slots: bots in ``player_names`` order, then agents in ``agent_ids`` order, then
zombie slots (``max(initial_zombies, minimum_zombies)`` of them).  Zombies get
slots in spawn order; zombies spawned later by the minimum-zombie flow take the
lowest free zombie slots in dict order.
"""
import os
import random
import struct
import sys

import numpy as np

from libzombsole_b200 import philox

REFERENCE_ROOT = os.environ.get("ZOMBSOLE_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

ACTION_NAMES = [None, "move", "attack_closest", "attack", "heal", "heal_closest"]
ACT_ABSENT = 6  # multi-agent only: the agent's key is missing from the action dict
WEAPON_CODES = {"ZombieClaws": 1, "Knife": 10, "Axe": 11, "Gun": 12, "Rifle": 13, "Shotgun": 14}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "zombsole"))


def import_reference():
    """Import the reference package (with shims if needed); returns the module dict."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for name in ("gymnasium", "termcolor"):
        try:
            __import__(name)
        except ImportError:
            if _SHIMS not in sys.path:
                sys.path.insert(0, _SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import zombsole.gym_env as gym_env
    import zombsole.gym.multiagent_env as multiagent_env
    import zombsole.things as things
    import zombsole.game as game
    return {"gym_env": gym_env, "multiagent_env": multiagent_env, "things": things, "game": game}


class InjectedRandom(random.Random):
    """random.Random whose every _randbelow comes from the Philox draw contract."""

    def __init__(self, seed):
        super().__init__(0)
        self.zs_seed = int(seed)
        self.cell = (0, 0, 0)
        self.k = 0
        self.log = []

    def begin(self, env_index, episode, t_word):
        self.cell = (int(env_index), int(episode), int(t_word))
        self.k = 0

    def _randbelow(self, n):
        u = philox.world_draw(self.zs_seed, self.cell[0], self.cell[1], self.cell[2], self.k)
        self.k += 1
        return philox.randbelow(u, n)


class injected_draws(object):
    """Context manager: rebind random.shuffle/randint/choice to an InjectedRandom."""

    def __init__(self, seed):
        self.rng = InjectedRandom(seed)

    def __enter__(self):
        self._saved = (random.shuffle, random.randint, random.choice)
        random.shuffle, random.randint, random.choice = self.rng.shuffle, self.rng.randint, self.rng.choice
        return self.rng

    def __exit__(self, *exc):
        random.shuffle, random.randint, random.choice = self._saved
        return False


def f64_bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def action_dict(a, multi):
    """[type, dx, dy] -> the reference's action dict."""
    name = ACTION_NAMES[int(a[0])]
    d = {"action_type": name}
    if name in ("move", "attack") or (name == "heal" and (a[1] or a[2])) or multi:
        d["parameter"] = [int(a[1]), int(a[2])]
    return d


class RefRunner(object):
    """One reference env (global index ``env_index``) under injected draws."""

    def __init__(self, cfg, env_index, rng):
        self.mods = import_reference()
        self.cfg = cfg
        self.env_index = env_index
        self.rng = rng
        self.multi = cfg["kind"] == "multi"
        self.episode = 0
        self.steps_in_episode = 0
        rng.begin(env_index, 0, 0)
        if self.multi:
            self.env = self.mods["multiagent_env"].MultiagentZombsoleEnv(
                cfg["rules_name"], list(cfg["player_names"]), cfg["map_name"], list(cfg["agent_ids"]),
                initial_zombies=cfg["initial_zombies"], minimum_zombies=cfg["minimum_zombies"],
                observation_surroundings_width=cfg["surroundings_width"],
                agent_weapons=cfg["agent_weapons"], debug=False)
        else:
            self.env = self.mods["gym_env"].ZombsoleGymEnv(
                cfg["rules_name"], list(cfg["player_names"]), cfg["map_name"], cfg["agent_ids"][0],
                initial_zombies=cfg["initial_zombies"], minimum_zombies=cfg["minimum_zombies"],
                observation_scope=cfg["observation_scope"],
                observation_position_encoding=cfg["observation_position_encoding"],
                agent_weapon=cfg["agent_weapons"] if isinstance(cfg["agent_weapons"], str) else cfg["agent_weapons"][0],
                debug=False)
        self.init_draws = rng.k
        game = self.env.game
        self.n_bots = len(game.players)
        self.n_agents = len(game.agents)
        self.n_zslots = max(cfg["initial_zombies"], cfg["minimum_zombies"])
        self.n_slots = self.n_bots + self.n_agents + self.n_zslots
        self.statics = [t for t in game.map.things if not t.is_decoration]
        self.width, self.height = game.map.size
        self._assign_zombie_slots(fresh=True)

    # ---- slot bookkeeping -------------------------------------------------
    def _assign_zombie_slots(self, fresh):
        Zombie = self.mods["things"].Zombie
        world = self.env.game.world
        if fresh:
            self.zslot = {}
            self.zobj = [None] * self.n_zslots
        in_world = [t for t in world.things.values() if isinstance(t, Zombie)]
        live_ids = set(id(t) for t in in_world)
        for s, z in enumerate(self.zobj):
            if z is not None and id(z) not in live_ids:
                self.zobj[s] = None
                del self.zslot[id(z)]
        for z in in_world:
            if id(z) not in self.zslot:
                s = self.zobj.index(None)
                self.zobj[s] = z
                self.zslot[id(z)] = s

    def _slot_things(self):
        game = self.env.game
        return list(game.players) + list(game.agents) + list(self.zobj)

    # ---- state dump -------------------------------------------------------
    def dump_state(self):
        game = self.env.game
        world = game.world
        DeadBody = self.mods["things"].DeadBody
        M = self.n_slots
        xs = np.full(M, -1, np.int16)
        ys = np.full(M, -1, np.int16)
        life = np.zeros(M, np.int16)
        inw = np.zeros(M, np.uint8)
        weapon = np.zeros(M, np.uint8)
        slot_of = {}
        for s, t in enumerate(self._slot_things()):
            if t is None:
                continue
            slot_of[id(t)] = s
            xs[s], ys[s] = t.position
            life[s] = t.life
            inw[s] = 1 if world.things.get(tuple(t.position)) is t else 0
            weapon[s] = WEAPON_CODES[t.weapon.name]
        order = np.full(M, -1, np.int16)
        n = 0
        for t in world.things.values():
            if id(t) in slot_of:
                order[n] = slot_of[id(t)]
                n += 1
        S = len(self.statics)
        slife = np.zeros(S, np.int16)
        spres = np.zeros(S, np.uint8)
        for i, t in enumerate(self.statics):
            slife[i] = t.life
            spres[i] = 1 if world.things.get(t.position) is t else 0
        dead = np.zeros(self.width * self.height, np.uint8)
        for pos, d in world.decoration.items():
            if isinstance(d, DeadBody):
                dead[pos[1] * self.width + pos[0]] = 1
        return {
            "x": xs, "y": ys, "life": life, "in_world": inw, "weapon": weapon, "order": order,
            "static_life": slife, "static_present": spres, "dead_body": np.packbits(dead, bitorder="little"),
            "counters": np.array([world.t, world.deaths, world.zombie_deaths], np.int32),
        }

    def _obs_array(self, obs):
        if not self.multi:
            return np.asarray(obs, dtype=np.int32)
        ids = list(self.cfg["agent_ids"])
        w = self.cfg["surroundings_width"]
        out = np.zeros((len(ids), 3, w, w), np.int32)
        mask = np.zeros(len(ids), np.uint8)
        for i, aid in enumerate(ids):
            if aid in obs:
                out[i] = np.asarray(obs[aid], dtype=np.int64).astype(np.int32)
                mask[i] = 1
        return out, mask

    # ---- ops --------------------------------------------------------------
    def initial(self):
        rec = self.dump_state()
        obs = self.env.get_observation()
        if self.multi:
            rec["obs"], rec["obs_mask"] = self._obs_array(obs)
        else:
            rec["obs"] = self._obs_array(obs)
        rec["draws"] = np.int32(self.init_draws)
        return rec

    def reset(self):
        self.episode += 1
        self.steps_in_episode = 0
        self.rng.begin(self.env_index, self.episode, 0)
        obs, _ = self.env.reset()
        self._assign_zombie_slots(fresh=True)
        rec = self.dump_state()
        if self.multi:
            rec["obs"], rec["obs_mask"] = self._obs_array(obs)
        else:
            rec["obs"] = self._obs_array(obs)
        rec["draws"] = np.int32(self.rng.k)
        return rec

    def step(self, actions, max_episode_steps=0):
        """actions: int array [A, 3] (type, dx, dy)."""
        world = self.env.game.world
        self.rng.begin(self.env_index, self.episode, world.t + 2)
        ids = list(self.cfg["agent_ids"])
        if self.multi:
            alive_before = list(self.env.agents)
            # type 6 (ZS_ACT_ABSENT) in a multi-agent tape means "key missing" (the reference then
            # heals self, zombsole/gym/multiagent_env.py:129-131)
            act = {aid: action_dict(actions[i], True) for i, aid in enumerate(ids) if int(actions[i][0]) != ACT_ABSENT}
            obs, rew, done, trunc, _ = self.env.step(act)
        else:
            obs, rew, done, trunc, _ = self.env.step(action_dict(actions[0], False))
        self.steps_in_episode += 1
        self._assign_zombie_slots(fresh=False)
        rec = self.dump_state()
        rec["draws"] = np.int32(self.rng.k)
        if self.multi:
            rec["obs"], rec["obs_mask"] = self._obs_array(obs)
            rb = np.zeros(len(ids), np.uint64)
            for i, aid in enumerate(ids):
                if aid in rew:
                    rb[i] = f64_bits(rew[aid])
            rec["reward_bits"] = rb
            rec["alive_before"] = np.array([1 if aid in alive_before else 0 for aid in ids], np.uint8)
            terminated = bool(done and all(done.values())) if done else False
            truncated = bool(trunc and all(trunc.values())) if trunc else False
            # done/truncated dicts carry one shared flag (multiagent_env.py:165-166); with no agent
            # alive before the step the dicts are empty, so re-derive the flag from the rules.
            if not done:
                ended = self.env.game.rules.game_ended()
                terminated = bool(ended)
                truncated = bool((not ended) and (not self.env.game.rules.agents_alive()))
        else:
            rec["obs"] = self._obs_array(obs)
            rec["reward_bits"] = np.array([f64_bits(rew)], np.uint64)
            terminated, truncated = bool(done), bool(trunc)
        if max_episode_steps and self.steps_in_episode >= max_episode_steps:
            truncated = True  # gymnasium TimeLimit (registered with max_episode_steps=1000, gym_env.py:385)
        rec["terminated"] = np.uint8(terminated)
        rec["truncated"] = np.uint8(truncated)
        return rec


STATE_KEYS = ["x", "y", "life", "in_world", "weapon", "order", "static_life", "static_present",
              "dead_body", "counters", "obs", "draws"]


def run_trace(cfg, env_index, seed, actions, max_episode_steps=0):
    """Run one env over an action tape [T, A, 3] with same-tick reset on done|truncated.

    Returns dict of stacked arrays: init_*, step_* [T, ...], reset_* [T, ...] (valid where
    did_reset[t] == 1).
    """
    out = {}
    with injected_draws(seed) as rng:
        runner = RefRunner(cfg, env_index, rng)
        init = runner.initial()
        steps, resets, did_reset = [], [], []
        for t in range(actions.shape[0]):
            rec = runner.step(actions[t], max_episode_steps)
            steps.append(rec)
            if rec["terminated"] or rec["truncated"]:
                resets.append(runner.reset())
                did_reset.append(1)
            else:
                resets.append(None)
                did_reset.append(0)
    for k, v in init.items():
        out["init_" + k] = np.asarray(v)
    for k in steps[0].keys():
        out["step_" + k] = np.stack([np.asarray(r[k]) for r in steps])
    template = next((r for r in resets if r is not None), None)
    if template is not None:
        for k in template.keys():
            zero = np.zeros_like(np.asarray(template[k]))
            out["reset_" + k] = np.stack([np.asarray(r[k]) if r is not None else zero for r in resets])
    out["did_reset"] = np.array(did_reset, np.uint8)
    out["actions"] = np.asarray(actions, np.int32)
    return out
