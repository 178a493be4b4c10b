"""Multi-agent environments: the reference's dict-of-agents API over the batched B200 simulator.

Mirrors zombsole/gym/multiagent_env.py:
  * ``MultiagentZombsoleEnv`` / ``MultiagentZombsoleEnvDiscreteAction``: the reference's constructor
    arguments and return types (dicts keyed by agent id; observation ``(3, w, w)`` per agent); one world.
  * ``MultiagentZombsoleVectorEnv``: the same arguments plus ``num_envs``; ``step(actions [N, A])`` ->
    ``(obs [N, A, 3, w, w] int32, reward [N, A] float64, terminated [N] bool, truncated [N] bool,
    {"agent_mask": [N, A] bool})`` as CUDA tensors.  ``agent_mask`` marks the agents that were alive
    before the step, i.e. the keys the reference's per-agent dicts would carry
    (multiagent_env.py:88-97,156-166); rewards of the others are 0 and their windows are centred on
    the position where they died.
"""
import numpy as np
import torch

from .. import abi
from ..engine import ZsEngine
from ..gym_env import encode_action
from ..spaces import Box, Dict, Discrete, Text
from ..things import Game


class MultiagentZombsoleVectorEnv(object):
    metadata = {"render.modes": ["human"]}
    reward_range = (-float("inf"), float("inf"))
    game_actions = [
        {"action_type": "move", "parameter": [0, 1]},
        {"action_type": "move", "parameter": [-1, 0]},
        {"action_type": "move", "parameter": [0, -1]},
        {"action_type": "move", "parameter": [1, 0]},
        {"action_type": "attack_closest"},
        {"action_type": "heal"},
        {"action_type": "heal_closest"},
    ]

    def __init__(self, rules_name, player_names, map_name, agent_ids, initial_zombies=0, minimum_zombies=0,
                 render_mode=None, observation_surroundings_width=21, observation_position_encoding_style="channels",
                 agent_weapons="rifle", debug=False, *, num_envs=1, device="cuda", seed=0, env_index_base=0,
                 max_episode_steps=None, auto_reset=True, host_outputs=False):
        if render_mode is not None:
            if render_mode not in self.metadata["render.modes"]:
                raise ValueError("render_mode={} is not supported".format(render_mode))
            raise NotImplementedError("rendering is outside the batched simulator's scope (render_mode=None only)")
        self.render_mode = None
        width = observation_surroundings_width
        if (width % 2 == 0) or (width <= 1):  # observation.py:206-207
            raise ValueError("surroundings width must be an odd number greater than 1")
        lpes = observation_position_encoding_style.lower()
        if lpes not in ["simple", "channels"]:
            raise ValueError(f"{lpes} must be \"simple\" or \"channels\"")
        if lpes != "channels":
            # the reference's get_observation needs get_observation_at_position, which only the channels
            # handler has (multiagent_env.py:92 vs observation.py:145-156): "simple" raises AttributeError there
            raise AttributeError("'SurroundingsSimpleObservation' object has no attribute 'get_observation_at_position'")
        self.surroundings_width = width
        self.position_encoding_style = observation_position_encoding_style
        self.possible_agents = list(agent_ids)
        self.cfg = abi.make_config(rules_name, list(player_names), list(agent_ids), agent_weapons, initial_zombies,
                                   minimum_zombies, abi.OBS_SURROUNDINGS, abi.OBS_CHANNELS, width, True, num_envs,
                                   seed=seed, env_index_base=env_index_base, max_episode_steps=max_episode_steps,
                                   auto_reset=auto_reset)
        self.engine = ZsEngine(self.cfg, map_name, device=device)
        self.num_envs = num_envs
        self.num_agents = len(agent_ids)
        self.device = self.engine.device
        self.debug = debug
        self._ctor = (rules_name, list(player_names), list(agent_ids), initial_zombies, minimum_zombies)
        self.action_spaces = {aid: Discrete(len(self.game_actions)) for aid in self.possible_agents}
        self.observation_spaces = {aid: Box(low=0, high=128, shape=(3, width, width), dtype=np.int32)
                                   for aid in self.possible_agents}
        # host_outputs: the kernel writes observations, rewards, flags and the agent mask straight into pinned host memory and
        # reads the actions from there; step() synchronises the stream once and also leaves the agents' lives after the step
        # in `agent_life_host` (what a host loop needs to know who is still playing, multiagent_env.py:169)
        self.host_outputs = bool(host_outputs)
        if self.host_outputs:
            self.obs, self.reward, self._term, self._trunc = self.engine.new_host_outputs()
            self._mask = torch.ones((num_envs, self.num_agents), dtype=torch.uint8).pin_memory()
            self._actions = torch.zeros((num_envs, self.num_agents, 3), dtype=torch.int32).pin_memory()
            self.agent_life_host = torch.zeros((num_envs, self.num_agents), dtype=torch.int16).pin_memory()
        else:
            self.obs = self.engine.new_obs()
            self.reward, self._term, self._trunc = self.engine.new_outputs()
            self._mask = torch.ones((num_envs, self.num_agents), dtype=torch.uint8, device=self.device)
            self._actions = torch.zeros((num_envs, self.num_agents, 3), dtype=torch.int32, device=self.device)
        self._h2d_done, self._h2d_pending = None, False

    def game(self, env=0):
        rules_name, player_names, agent_ids, iz, mz = self._ctor
        return Game(self.engine, env, rules_name, player_names, agent_ids, iz, mz)

    def get_observation(self):
        self.engine.encode_obs(self.obs)
        if self.host_outputs:
            torch.cuda.current_stream(self.device).synchronize()
        return self.obs

    def _stage_actions(self, actions):
        N, A = self.num_envs, self.num_agents
        if isinstance(actions, (list, tuple)) and len(actions) and isinstance(actions[0], dict):
            # one dict per env, keyed by agent id; a missing key is ZS_ACT_ABSENT (heal self, multiagent_env.py:129-131)
            rows = np.zeros((N, A, 3), np.int32)
            for n, d in enumerate(actions):
                for i, aid in enumerate(self.possible_agents):
                    rows[n, i] = encode_action(dict(d[aid], parameter=d[aid].get("parameter", [0, 0]))) \
                        if aid in d else (abi.ACT_ABSENT, 0, 0)
            self._actions.copy_(torch.from_numpy(rows), non_blocking=not self.host_outputs)
            return self._actions, abi.ACTIONS_FULL
        t = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
        if self.host_outputs and t.dtype == torch.int32 and t.device.type == "cpu" and t.is_pinned():
            pass  # the kernel reads a pinned host action tensor in place
        elif t.dtype != torch.int32 or t.device != self.device:
            pinned_src = t.device.type == "cpu" and t.is_pinned()
            t = t.to(device=self.device, dtype=torch.int32, non_blocking=True)
            if pinned_src:  # the copy is in flight: step() waits for it before the caller may refill the buffer
                self._h2d_done = self._h2d_done or torch.cuda.Event()
                self._h2d_done.record(torch.cuda.current_stream(self.device))
                self._h2d_pending = True
        t = t.contiguous()
        if t.numel() == N * A:
            return t.view(N, A), abi.ACTIONS_DISCRETE
        if t.numel() == 3 * N * A:
            return t.view(N, A, 3), abi.ACTIONS_FULL
        raise ValueError("actions must hold %d discrete ids or (type, dx, dy) rows" % (N * A))

    def step(self, actions):
        """One transition of every world (multiagent_env.py:111-171); discrete id -1 = key missing."""
        a, fmt = self._stage_actions(actions)
        self.engine.step(a, fmt, self.obs, self.reward, self._term, self._trunc, self._mask)
        if self.host_outputs:  # the host owns the results (and any pinned action buffer) when step() returns
            P = self.cfg.n_bots
            self.agent_life_host.copy_(self.engine.fields["life"][:, P:P + self.num_agents], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            self._h2d_pending = False
        if self._h2d_pending:  # a pinned host action buffer is the caller's again when step() returns
            self._h2d_done.synchronize()
            self._h2d_pending = False
        return (self.obs, self.reward, self._term.view(torch.bool), self._trunc.view(torch.bool),
                {"agent_mask": self._mask.view(torch.bool)})  # (0/1 bytes: views, no kernels)

    def reset(self, seed=None, options=None, mask=None):
        self.engine.reset(mask, self.obs)
        if self.host_outputs:
            torch.cuda.current_stream(self.device).synchronize()
        return self.obs, {}

    def rollout(self, n_steps, actions=None, first_step_index=0, obs=None, reward=None, terminated=None, truncated=None):
        fmt = abi.ACTIONS_DISCRETE
        if actions is not None and actions.dim() == 4:
            fmt = abi.ACTIONS_FULL
        self.engine.rollout(n_steps, first_step_index, actions, fmt, self.obs if obs is None else obs, reward,
                            terminated, truncated)
        return self.obs if obs is None else obs

    def close(self):
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False


class MultiagentZombsoleEnv(object):
    """Drop-in for the reference's MultiagentZombsoleEnv (multiagent_env.py:15-216): one world, dicts."""

    metadata = {"render.modes": ["human"]}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, rules_name, player_names, map_name, agent_ids, initial_zombies=0, minimum_zombies=0,
                 render_mode=None, observation_surroundings_width=21, observation_position_encoding_style="channels",
                 agent_weapons="rifle", debug=False, *, device="cuda", seed=0, env_index_base=0):
        self.vec = MultiagentZombsoleVectorEnv(
            rules_name, player_names, map_name, agent_ids, initial_zombies, minimum_zombies, render_mode,
            observation_surroundings_width, observation_position_encoding_style, agent_weapons, debug,
            num_envs=1, device=device, seed=seed, env_index_base=env_index_base, max_episode_steps=None,
            auto_reset=False, host_outputs=True)
        self.position_encoding_style = observation_position_encoding_style
        self.surroundings_width = observation_surroundings_width
        self.agents = list(agent_ids)
        self.possible_agents = list(agent_ids)
        self.action_spaces = {aid: Dict({"action_type": Text(15),
                                         "parameter": Box(low=-10, high=10, shape=(2,), dtype=np.int32)})
                              for aid in self.possible_agents}
        self.observation_spaces = dict(self.vec.observation_spaces)
        self.render_mode = render_mode
        self.game = self.vec.game(0)

    def get_observation(self):
        obs = self.vec.get_observation()[0].numpy().copy()  # (host memory; the caller gets arrays of its own)
        return {aid: obs[i] for i, aid in enumerate(self.possible_agents) if aid in self.agents}

    def step(self, action):
        # one launch whose outputs land in pinned host memory, one synchronisation (the vector env's host_outputs mode)
        obs, reward, term, trunc, info = self.vec.step([action])
        obs, reward = obs[0].numpy().copy(), reward[0].numpy()
        doneflag, truncatedflag = bool(term[0]), bool(trunc[0])
        before = self.agents
        observations = {aid: obs[i] for i, aid in enumerate(self.possible_agents) if aid in before}
        rewards = {aid: float(reward[i]) for i, aid in enumerate(self.possible_agents) if aid in before}
        done = {aid: doneflag for aid in before}
        truncated = {aid: truncatedflag for aid in before}
        life = self.vec.agent_life_host[0].numpy()
        self.agents = [aid for i, aid in enumerate(self.possible_agents) if life[i] > 0]  # multiagent_env.py:169
        return observations, rewards, done, truncated, {}

    def reset(self, seed=None, options=None):
        self.agents = list(self.possible_agents)
        self.vec.reset()
        return self.get_observation(), {}

    def render(self):
        raise ValueError("mode={} is not supported".format(self.render_mode))

    def close(self):
        self.vec.close()

    def __str__(self):
        return "<{} instance>".format(type(self).__name__)

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False


class MultiagentZombsoleEnvDiscreteAction(object):
    """Drop-in for the reference's discrete wrapper (multiagent_env.py:258-319): Discrete(7) per agent."""

    game_actions = MultiagentZombsoleVectorEnv.game_actions

    def __init__(self, rules_name, player_names, map_name, agent_ids, initial_zombies=0, minimum_zombies=0,
                 render_mode=None, observation_surroundings_width=21, agent_weapons="rifle", debug=False,
                 **device_kwargs):
        self.env = MultiagentZombsoleEnv(rules_name, player_names, map_name, agent_ids,
                                         initial_zombies=initial_zombies, minimum_zombies=minimum_zombies,
                                         render_mode=render_mode,
                                         observation_surroundings_width=observation_surroundings_width,
                                         agent_weapons=agent_weapons, debug=debug, **device_kwargs)
        self.action_spaces = {aid: Discrete(len(self.game_actions)) for aid in self.env.possible_agents}
        self.observation_spaces = self.env.observation_spaces
        self.reward_range = self.env.reward_range
        self.metadata = self.env.metadata
        self.render_mode = self.env.render_mode

    def step(self, actions):
        return self.env.step(self.actions(actions))

    def reset(self, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def actions(self, actions):
        return {agent_id: self.game_actions[action] for agent_id, action in actions.items()}

    def reverse_actions(self, actions):
        return {agent_id: self.game_actions.index(action) for agent_id, action in actions.items()}

    def close(self):
        self.env.close()
