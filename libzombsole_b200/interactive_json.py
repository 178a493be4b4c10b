"""JSON over stdio: the reference's only wire format (zombsole/interactive_json.py:1-359), served from the
B200 simulator's drop-in environments.

Same protocol, line by line: one JSON request per input line, one JSON response per output line.

  requests  {"tag": "GameConfigUpdate", "parameters": {rules_name, map_name, players, agent_ids, ...}}
            {"tag": "GameStatus"} | {"tag": "StartGame"} | {"tag": "GameAction", "parameters": <action>} | {"tag": "Exit"}
  responses {"tag": "GameState", "parameters": {status, active, config_required, last_observation}}
            {"tag": "GameObservation", "parameters": {observation, reward, done, truncated, info}}
            {"tag": "Error", "parameters": "<message>"}

Class names, defaults (``minimum_zombies=10`` in GameConfig, interactive_json.py:90-93), message texts and quirks
(the status of a running game is ``null`` because ``_env_status`` falls off its last branch, :233-239; "wating for
game" is spelled that way) are the reference's, so a client written against ``zombsole-stdio-json`` reads the
same bytes.  The environments are ``libzombsole_b200.gym_env.ZombsoleGymEnv`` /
``libzombsole_b200.gym.multiagent_env.MultiagentZombsoleEnv``; randomness follows the counter-based draw
contract (philox.py), selected with ``--seed`` / ``--env-index``.

Usage:
    python -m libzombsole_b200.interactive_json [-r none] [-m] [--seed S] [--env-index I] [--device cuda]
"""
import argparse
import json
import sys
from abc import ABC, abstractmethod
from json import JSONEncoder

from .gym_env import ZombsoleGymEnv
from .gym.multiagent_env import MultiagentZombsoleEnv


class GameResponse(ABC):
    def to_dict(self):
        return {"tag": self.get_tag(), "parameters": self.get_parameters()}

    @abstractmethod
    def get_tag(self):
        pass

    @abstractmethod
    def get_parameters(self):
        pass


class GameStateEncoder(JSONEncoder):
    def default(self, o):
        try:
            d = o.to_dict()
        except TypeError:
            pass
        else:
            return d
        return super().default(o)


class GameStateResponse(GameResponse):
    """interactive_json.py:53-69"""

    def __init__(self, status, active, config_required, last_observation=None):
        self.status = status
        self.active = active
        self.config_required = config_required
        self.last_observation = last_observation

    def get_tag(self):
        return "GameState"

    def get_parameters(self):
        return {"status": self.status, "active": self.active, "config_required": self.config_required,
                "last_observation": self.last_observation}


class GameObservationResponse(GameResponse):
    """interactive_json.py:71-79"""

    def __init__(self, last_observation=None):
        self.last_observation = last_observation

    def get_tag(self):
        return "GameObservation"

    def get_parameters(self):
        return self.last_observation


class ErrorResponse(GameResponse):
    """interactive_json.py:81-89"""

    def __init__(self, message):
        self.message = message

    def get_tag(self):
        return "Error"

    def get_parameters(self):
        return self.message


class GameConfig(object):
    """interactive_json.py:91-108 (the defaults are the reference's: ten zombies to start with AND to maintain)"""

    def __init__(self, rules_name, map_name, players, agent_ids, initial_zombies=10, minimum_zombies=10,
                 observation_scope="world", observation_position_encoding="simple"):
        self.rules_name = rules_name
        self.map_name = map_name
        self.players = players
        self.agent_ids = agent_ids
        self.initial_zombies = initial_zombies
        self.minimum_zombies = minimum_zombies
        self.observation_scope = observation_scope
        self.observation_position_encoding = observation_position_encoding

    @classmethod
    def from_dict(cls, d):
        return cls(**d)


class GameRequest(ABC):
    """interactive_json.py:131-159"""

    @staticmethod
    def decode_hook(jsonobj):
        if "tag" in jsonobj:
            if (jsonobj["tag"] in ["GameConfigUpdate", "GameAction"]) and ("parameters" not in jsonobj):
                raise ValueError(f"A GameRequest with tag {jsonobj['tag']} must have key \"parameters\"")
            if jsonobj["tag"] == "GameConfigUpdate":
                return GameConfigUpdateRequest.from_dict(jsonobj["parameters"])
            elif jsonobj["tag"] == "GameStatus":
                return GameStatusRequest()
            elif jsonobj["tag"] == "Exit":
                return ExitRequest()
            elif jsonobj["tag"] == "StartGame":
                return StartGameRequest()
            elif jsonobj["tag"] == "GameAction":
                return GameActionRequest(jsonobj["parameters"])
            else:
                raise ValueError("GameRequest \"tag\" must be \"GameConfigUpdate\", \"GameAction\", \"GameStatus\", \"StartGame\", or \"Exit\"")
        else:  # simply pass the object through (used where objects are passed as parameters)
            return jsonobj

    @abstractmethod
    def update_game_manager(self, game_manager):
        pass


class GameConfigUpdateRequest(GameRequest):
    def __init__(self, game_config):
        self.game_config = game_config

    @classmethod
    def from_dict(cls, game_config_obj):
        return cls(GameConfig.from_dict(game_config_obj))

    def update_game_manager(self, game_manager):
        game_manager.set_game_config(self.game_config)


class GameStatusRequest(GameRequest):
    def update_game_manager(self, game_manager):
        game_manager.get_game_status()


class ExitRequest(object):
    def update_game_manager(self, game_manager):
        game_manager.exit()


class StartGameRequest(object):
    def update_game_manager(self, game_manager):
        game_manager.start_game()


class GameActionRequest(object):
    def __init__(self, action):
        self.action = action

    def update_game_manager(self, game_manager):
        game_manager.step_with_agent_action(self.action)


class GymEnvManager(object):
    """interactive_json.py:205-338 over the drop-in environments.  ``env_kwargs`` (seed, env_index_base, device) go to
    the environment constructors; ``instream`` / ``outstream`` default to stdin / stdout."""

    def __init__(self, render_mode, use_multiagent_env, instream=None, outstream=None, **env_kwargs):
        self.game_config = None
        self.gym_env = None
        self.keep_going = True
        self.last_observation = None
        self.response_encoder = GameStateEncoder(indent=None)
        self.render_mode = render_mode
        self.use_multiagent_env = use_multiagent_env
        self.instream = instream if instream is not None else sys.stdin
        self.outstream = outstream if outstream is not None else sys.stdout
        self.env_kwargs = env_kwargs

    def _initialize_gym(self):
        if self.game_config is not None:
            if self.gym_env is not None:
                self.gym_env.close()
            c = self.game_config
            if self.use_multiagent_env:
                scope = c.observation_scope
                swidth = int(scope[len("surroundings:"):]) if scope.startswith("surroundings:") else 21
                self.gym_env = MultiagentZombsoleEnv(
                    c.rules_name, c.players, c.map_name, c.agent_ids, initial_zombies=c.initial_zombies,
                    minimum_zombies=c.minimum_zombies, observation_surroundings_width=swidth,
                    render_mode=self.render_mode, debug=False, **self.env_kwargs)
            else:  # single agent
                self.gym_env = ZombsoleGymEnv(
                    c.rules_name, c.players, c.map_name, c.agent_ids[0], initial_zombies=c.initial_zombies,
                    minimum_zombies=c.minimum_zombies, observation_scope=c.observation_scope,
                    observation_position_encoding=c.observation_position_encoding, render_mode=self.render_mode,
                    debug=False, **self.env_kwargs)
            self.last_observation = None

    def _env_status(self):
        if not self.keep_going:
            return "exiting"
        elif self.last_observation is None:
            return "wating for game"
        else:
            return None  # the reference's last branch has no `return`: a game in progress reports null

    def _get_game_state(self):
        return GameStateResponse(self._env_status(), self.keep_going, self.game_config is None, self.last_observation)

    def _response_to_stdout(self, response):
        self.outstream.write(self.response_encoder.encode(response.to_dict()) + "\n")
        self.outstream.flush()

    def run(self):
        self._response_to_stdout(self._get_game_state())
        while self.keep_going:
            message = self.instream.readline()
            if message == "":
                raise EOFError("EOF when reading a line")  # what input() raises in the reference
            message = message.rstrip("\n")
            try:
                obj = json.loads(message, object_hook=GameRequest.decode_hook)
            except Exception as ex:
                self._response_to_stdout(ErrorResponse(str(ex)))
            else:
                obj.update_game_manager(self)

    # ---- the management interface (interactive_json.py:110-129)
    def set_game_config(self, game_config):
        self.game_config = game_config
        self._initialize_gym()
        self._response_to_stdout(self._get_game_state())

    def get_game_status(self):
        self._response_to_stdout(self._get_game_state())

    def _observation_json_ready(self, observation):
        if self.use_multiagent_env:
            return {agent_id: observation[agent_id].tolist() for agent_id in observation}
        return observation.tolist()

    def _initial_values(self):
        if self.use_multiagent_env:
            agent_ids = self.gym_env.possible_agents
            return ({a: 0 for a in agent_ids}, {a: False for a in agent_ids}, {a: False for a in agent_ids}, {})
        return 0, False, False, None

    def start_game(self):
        origobs, _ = self.gym_env.reset()
        reward, done, truncated, info = self._initial_values()
        self.last_observation = {"observation": self._observation_json_ready(origobs), "reward": reward, "done": done,
                                 "truncated": truncated, "info": info}
        self._response_to_stdout(GameObservationResponse(self.last_observation))

    def step_with_agent_action(self, action):
        observation, reward, done, truncated, info = self.gym_env.step(action)
        self.last_observation = {"observation": self._observation_json_ready(observation), "reward": reward, "done": done,
                                 "truncated": truncated, "info": info}
        # The reference calls gym_env.render() here; without a renderer (-r none) that call dies on an undefined name
        # (gym_env.py:207) and takes the reference server down at its first GameAction.  Rendering is outside the
        # batched simulator's scope, so the call is skipped and the protocol carries on.
        self._response_to_stdout(GameObservationResponse(self.last_observation))

    def exit(self):
        self.keep_going = False
        self._response_to_stdout(self._get_game_state())


def play_interactive_json(argv=None):
    """zombsole-stdio-json (interactive_json.py:340-359); docopt replaced by argparse, same options."""
    ap = argparse.ArgumentParser(description="Play Zombsole interactively using JSON over stdio (B200 simulator)")
    ap.add_argument("-r", dest="renderer", default="none", help="opencv or none [default: none]")
    ap.add_argument("-m", "--multi-agent", action="store_true", help="Play Multi-Agent Zombsole")
    ap.add_argument("--seed", type=int, default=0, help="seed of the draw contract")
    ap.add_argument("--env-index", type=int, default=0, help="global env index (a word of the draw counter)")
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args(argv)
    if args.renderer not in ["opencv", "none"]:
        print("When using interactive JSON mode, renderer_id must be one of \"opencv\" or \"none\".  Exiting...", file=sys.stderr)
        sys.exit(1)
    render_mode = "human" if args.renderer == "opencv" else None
    GymEnvManager(render_mode, args.multi_agent, seed=args.seed, env_index_base=args.env_index, device=args.device).run()


if __name__ == "__main__":
    play_interactive_json()
