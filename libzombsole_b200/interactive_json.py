"""JSON over stdio served from ONE batch of worlds on the GPU.

The wire format is the reference's (zombsole/interactive_json.py): one JSON request per input line, one JSON response
per output line —

  requests   {"tag": "GameConfigUpdate", "parameters": {rules_name, map_name, players, agent_ids, ...}}
             {"tag": "GameStatus"} | {"tag": "StartGame"} | {"tag": "GameAction", "parameters": <action>} | {"tag": "Exit"}
  responses  {"tag": "GameState", "parameters": {status, active, config_required, last_observation}}
             {"tag": "GameObservation", "parameters": {observation, reward, done, truncated, info}}
             {"tag": "Error", "parameters": "<message>"}

— and a client that never says which world it means talks to world 0 and reads the same bytes the reference server
prints under the same draws (tests/golden/json_session_*.json are transcripts of the reference's own server).

What is different is what sits behind it.  The reference holds one Python env per process.  This server holds ONE
vector env (``ZombsoleVectorEnv`` / ``MultiagentZombsoleVectorEnv``, ``--num-envs`` worlds resident in HBM) and every
request may carry ``"env": i`` to address world i of the batch: ``StartGame`` re-initialises that world alone (masked
reset), ``GameAction`` advances that world alone (masked step, zs_step_masked) and ``GameStatus`` reports it.
``"env": "all"`` applies ``StartGame`` / ``GameAction`` to the whole batch in one launch — ``parameters`` is then a list
with one action per world — and answers with a ``GameObservations`` list.  A ``GameConfigUpdate`` rebuilds the batch.

The requests are handled by a dispatch table (tag -> handler, needs-parameters flag), not by request classes.

Quirks kept because clients see them: ``GameConfig`` defaults to ten zombies to start with AND to maintain
(interactive_json.py:91-93); the status of a running game is ``null`` and a waiting one is spelled "wating for game"
(:233-239).  Where the reference server dies (a request without "tag", a GameAction before any game, its own render()
call without a renderer, gym_env.py:207) this one answers with an ``Error`` line, or carries on, instead.

Usage:
    python -m libzombsole_b200.interactive_json [-r none] [-m] [--num-envs N] [--seed S] [--env-index I] [--device cuda]
"""
import argparse
import json
import sys

import numpy as np
import torch

from . import abi
from .gym.multiagent_env import MultiagentZombsoleVectorEnv
from .gym_env import ZombsoleVectorEnv, encode_action

TAGS_TEXT = ('GameRequest "tag" must be "GameConfigUpdate", "GameAction", "GameStatus", "StartGame", or "Exit"')


class GameConfig(object):
    """The parameters of a GameConfigUpdate.  Built with ``GameConfig(**parameters)`` so that a missing or unknown key
    is reported in Python's own words, as the reference's constructor call does (interactive_json.py:91-108)."""
    __slots__ = ("rules_name", "map_name", "players", "agent_ids", "initial_zombies", "minimum_zombies",
                 "observation_scope", "observation_position_encoding")

    def __init__(self, rules_name, map_name, players, agent_ids, initial_zombies=10, minimum_zombies=10,
                 observation_scope="world", observation_position_encoding="simple"):
        for name, value in zip(self.__slots__, (rules_name, map_name, players, agent_ids, initial_zombies, minimum_zombies,
                                                observation_scope, observation_position_encoding)):
            setattr(self, name, value)

    @classmethod
    def from_dict(cls, d):
        return cls(**d)


def _line(tag, parameters):
    return json.dumps({"tag": tag, "parameters": parameters})


class BatchedJsonServer(object):
    """``num_envs`` worlds behind one vector-env handle, driven by request lines.  ``env_kwargs`` (seed, env_index_base,
    device) go to the vector env; ``instream`` / ``outstream`` default to stdin / stdout."""

    def __init__(self, render_mode, use_multiagent_env, instream=None, outstream=None, num_envs=1, **env_kwargs):
        if render_mode is not None:
            raise NotImplementedError("rendering is outside the batched simulator's scope (-r none only)")
        self.multi = bool(use_multiagent_env)
        self.num_envs = int(num_envs)
        self.instream = instream if instream is not None else sys.stdin
        self.outstream = outstream if outstream is not None else sys.stdout
        self.env_kwargs = env_kwargs
        self.config = None
        self.vec = None                     # the one handle
        self.active = True
        self.last = [None] * self.num_envs  # per world: the last GameObservation parameters
        self.alive = None                   # multi-agent: per world, the agent ids that still get entries
        #: tag -> (handler, the request must carry "parameters")
        self.dispatch = {
            "GameConfigUpdate": (self._on_config, True),
            "GameStatus": (self._on_status, False),
            "StartGame": (self._on_start, False),
            "GameAction": (self._on_action, True),
            "Exit": (self._on_exit, False),
        }

    # ------------------------------------------------------------------ the loop
    @property
    def gym_env(self):
        return self.vec

    def run(self):
        self._emit(self._state_line(0))
        while self.active:
            message = self.instream.readline()
            if message == "":
                raise EOFError("EOF when reading a line")  # what input() raises in the reference
            self._emit(self.handle(message.rstrip("\n")))

    def _emit(self, text):
        self.outstream.write(text + "\n")
        self.outstream.flush()

    def handle(self, message):
        """One request line -> one response line."""
        try:
            request = json.loads(message)
            if not isinstance(request, dict) or "tag" not in request:
                raise ValueError(TAGS_TEXT)
            tag = request["tag"]
            if tag in ("GameConfigUpdate", "GameAction") and "parameters" not in request:
                raise ValueError(f"A GameRequest with tag {tag} must have key \"parameters\"")
            if tag not in self.dispatch:
                raise ValueError(TAGS_TEXT)
            handler, _ = self.dispatch[tag]
            which = self._which(request.get("env", 0))
            return handler(request.get("parameters"), which)
        except Exception as ex:  # the reference answers every failed decode with an Error line (interactive_json.py:260-262)
            return _line("Error", str(ex))

    def _which(self, env):
        if env == "all":
            return None
        if isinstance(env, bool) or not isinstance(env, int) or not 0 <= env < self.num_envs:
            raise ValueError(f"\"env\" must be \"all\" or a world index below {self.num_envs}")
        return env

    # ------------------------------------------------------------------ responses
    def _status(self, env):
        if not self.active:
            return "exiting"
        if self.last[env] is None:
            return "wating for game"
        return None  # a game in progress reports null, as the reference does

    def _state_line(self, env):
        return _line("GameState", {"status": self._status(env), "active": self.active, "config_required": self.config is None,
                                   "last_observation": self.last[env]})

    # ------------------------------------------------------------------ handlers
    def _on_config(self, parameters, which):
        cfg = GameConfig.from_dict(parameters)
        if self.vec is not None:
            self.vec.close()
            self.vec = None
        common = dict(num_envs=self.num_envs, max_episode_steps=None, auto_reset=False, **self.env_kwargs)
        if self.multi:
            scope = cfg.observation_scope
            width = int(scope[len("surroundings:"):]) if scope.startswith("surroundings:") else 21
            self.vec = MultiagentZombsoleVectorEnv(cfg.rules_name, cfg.players, cfg.map_name, cfg.agent_ids,
                                                   initial_zombies=cfg.initial_zombies, minimum_zombies=cfg.minimum_zombies,
                                                   observation_surroundings_width=width, **common)
        else:
            self.vec = ZombsoleVectorEnv(cfg.rules_name, cfg.players, cfg.map_name, cfg.agent_ids[0],
                                         initial_zombies=cfg.initial_zombies, minimum_zombies=cfg.minimum_zombies,
                                         observation_scope=cfg.observation_scope,
                                         observation_position_encoding=cfg.observation_position_encoding, **common)
        self.config = cfg
        self.last = [None] * self.num_envs
        self.alive = [list(cfg.agent_ids) for _ in range(self.num_envs)] if self.multi else None
        return self._state_line(0 if which is None else which)

    def _on_status(self, parameters, which):
        return self._state_line(0 if which is None else which)

    def _on_exit(self, parameters, which):
        self.active = False
        return self._state_line(0 if which is None else which)

    def _need_env(self):
        if self.vec is None:
            raise ValueError("no game configured: send a GameConfigUpdate first")

    def _mask(self, which):
        m = np.ones(self.num_envs, np.uint8) if which is None else np.zeros(self.num_envs, np.uint8)
        if which is not None:
            m[which] = 1
        return torch.from_numpy(m)

    def _on_start(self, parameters, which):
        self._need_env()
        envs = range(self.num_envs) if which is None else [which]
        self.vec.engine.reset(self._mask(which), self.vec.obs)
        obs = self.vec.obs.cpu().numpy()
        ids = self.config.agent_ids
        for e in envs:
            if self.multi:
                self.alive[e] = list(ids)
                self.last[e] = {"observation": {a: obs[e, i].tolist() for i, a in enumerate(ids)},
                                "reward": {a: 0 for a in ids}, "done": {a: False for a in ids},
                                "truncated": {a: False for a in ids}, "info": {}}
            else:
                self.last[e] = {"observation": obs[e].tolist(), "reward": 0, "done": False, "truncated": False, "info": None}
        return self._observation_line(which)

    def _observation_line(self, which):
        if which is None:
            return _line("GameObservations", list(self.last))
        return _line("GameObservation", self.last[which])

    def _encode_actions(self, parameters, which):
        """-> int32 [N, A, 3] (type, dx, dy) rows; the worlds that are not addressed get no-ops (they are masked out)."""
        N, ids = self.num_envs, self.config.agent_ids
        A = len(ids) if self.multi else 1
        rows = np.zeros((N, A, 3), np.int32)
        if which is None:
            if not isinstance(parameters, list) or len(parameters) != N:
                raise ValueError(f"\"env\": \"all\" needs a list of {N} actions as \"parameters\"")
            todo = list(enumerate(parameters))
        else:
            todo = [(which, parameters)]
        for e, action in todo:
            if self.multi:
                for i, a in enumerate(ids):
                    # a missing key is ZS_ACT_ABSENT: the reference makes that agent heal itself (multiagent_env.py:129-131)
                    rows[e, i] = encode_action(dict(action[a], parameter=action[a].get("parameter", [0, 0]))) if a in action \
                        else (abi.ACT_ABSENT, 0, 0)
            else:
                rows[e, 0] = encode_action(action)
        return rows

    def _on_action(self, parameters, which):
        self._need_env()
        envs = list(range(self.num_envs)) if which is None else [which]
        if any(self.last[e] is None for e in envs):
            raise ValueError("no game in progress: send StartGame first")
        vec, eng = self.vec, self.vec.engine
        rows = torch.from_numpy(self._encode_actions(parameters, which)).to(vec.device)
        amask = vec._mask if self.multi else None
        eng.step_masked(self._mask(which), rows, abi.ACTIONS_FULL, vec.obs, vec.reward, vec._term, vec._trunc, amask)
        obs, reward = vec.obs.cpu().numpy(), vec.reward.cpu().numpy()
        term, trunc = vec._term.cpu().numpy(), vec._trunc.cpu().numpy()
        if self.multi:
            ids = self.config.agent_ids
            P = eng.P
            life = eng.fields["life"][:, P:P + len(ids)].cpu().numpy()
        for e in envs:
            done, truncated = bool(term[e]), bool(trunc[e])
            if self.multi:
                before = self.alive[e]  # the reference's dicts are keyed by the agents alive before the step
                self.last[e] = {"observation": {a: obs[e, i].tolist() for i, a in enumerate(ids) if a in before},
                                "reward": {a: float(reward[e, i]) for i, a in enumerate(ids) if a in before},
                                "done": {a: done for a in before}, "truncated": {a: truncated for a in before}, "info": {}}
                self.alive[e] = [a for i, a in enumerate(ids) if life[e, i] > 0]  # multiagent_env.py:169
            else:
                self.last[e] = {"observation": obs[e].tolist(), "reward": float(reward[e]), "done": done,
                                "truncated": truncated, "info": {}}
        return self._observation_line(which)


#: the reference's name for its server object (zombsole/interactive_json.py:195)
GymEnvManager = BatchedJsonServer


def play_interactive_json(argv=None):
    """zombsole-stdio-json (interactive_json.py:340-359); docopt replaced by argparse, same options plus the batch's."""
    ap = argparse.ArgumentParser(description="Play Zombsole interactively using JSON over stdio (B200 simulator)")
    ap.add_argument("-r", dest="renderer", default="none", help="opencv or none [default: none]")
    ap.add_argument("-m", "--multi-agent", action="store_true", help="Play Multi-Agent Zombsole")
    ap.add_argument("--num-envs", type=int, default=1, help="worlds in the batch; requests address one with \"env\": i")
    ap.add_argument("--seed", type=int, default=0, help="seed of the draw contract")
    ap.add_argument("--env-index", type=int, default=0, help="global env index of world 0 (a word of the draw counter)")
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args(argv)
    if args.renderer not in ["opencv", "none"]:
        print("When using interactive JSON mode, renderer_id must be one of \"opencv\" or \"none\".  Exiting...", file=sys.stderr)
        sys.exit(1)
    render_mode = "human" if args.renderer == "opencv" else None
    BatchedJsonServer(render_mode, args.multi_agent, num_envs=args.num_envs, seed=args.seed, env_index_base=args.env_index,
                      device=args.device).run()


if __name__ == "__main__":
    play_interactive_json()
