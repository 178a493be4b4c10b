"""The reference's own tests (tests/test_game.py, tests/test_multiagent_env.py, tests/test_gym_env.py of
jvstinian/libzombsole) run against the drop-in classes: same constructor arguments, same object
protocol (env.game.world.things, agent.position, thing.life), same assertions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_single(**kw):
    from libzombsole_b200.gym_env import ZombsoleGymEnv
    args = dict(rules_name="extermination", player_names=[], map_name="boxed", agent_id=0, initial_zombies=1,
                minimum_zombies=0, render_mode=None, observation_scope="world",
                observation_position_encoding="simple", debug=True)
    args.update(kw)
    return ZombsoleGymEnv(args.pop("rules_name"), args.pop("player_names"), args.pop("map_name"), args.pop("agent_id"), **args)


# ---- tests/test_game.py -------------------------------------------------------------------------
def test_game_targeted_attack():
    from libzombsole_b200.things import Zombie
    gym_env = make_single()
    zombies = [thing for thing in gym_env.game.world.things.values() if isinstance(thing, Zombie)]
    assert len(zombies) > 0
    zombie = zombies[0]
    initial_zombie_life = zombie.life
    zombiepos = zombie.position
    agentpos = gym_env.game.agents[0].position
    relativepos = (zombiepos[0] - agentpos[0], zombiepos[1] - agentpos[1])
    gym_env.step({"action_type": "attack", "parameter": relativepos})
    assert zombie.life < initial_zombie_life


def test_game_targeted_heal():
    gym_env = make_single(player_names=["terminator"])
    gym_env.game.players[0].life = 25
    playerpos = gym_env.game.players[0].position
    agentpos = gym_env.game.agents[0].position
    relativepos = (playerpos[0] - agentpos[0], playerpos[1] - agentpos[1])
    gym_env.step({"action_type": "heal", "parameter": relativepos})
    assert gym_env.game.players[0].life > 25


def test_game_heal_closest():
    gym_env = make_single(player_names=["terminator"])
    gym_env.game.players[0].life = 25
    gym_env.step({"action_type": "heal_closest", "parameter": [0, 0]})
    assert gym_env.game.players[0].life > 25


def test_game_heal_self():
    gym_env = make_single()
    gym_env.game.agents[0].life = 25
    gym_env.step({"action_type": "heal", "parameter": [0, 0]})
    assert gym_env.game.agents[0].life > 25


def test_discrete_game_closest_attack():
    from libzombsole_b200.gym_env import ZombsoleGymEnvDiscreteAction
    from libzombsole_b200.spaces import Discrete
    from libzombsole_b200.things import Zombie
    gym_env = ZombsoleGymEnvDiscreteAction("extermination", [], "boxed", 0, initial_zombies=1, minimum_zombies=0,
                                           render_mode=None, observation_scope="world",
                                           observation_position_encoding="simple", debug=True)
    assert isinstance(gym_env.action_space, (Discrete,))
    zombies = [thing for thing in gym_env.game.world.things.values() if isinstance(thing, Zombie)]
    assert len(zombies) > 0
    zombie = zombies[0]
    initial_zombie_life = zombie.life
    action_id = gym_env.reverse_action({"action_type": "attack_closest"})
    assert action_id == 4
    gym_env.step(action_id)
    assert zombie.life < initial_zombie_life
    gym_env.reset()


# ---- tests/test_gym_env.py (shape tests; check_env needs real gymnasium) ----------------------------
@pytest.mark.parametrize("scope,position_encoding", [("world", "simple"), ("world", "channels")])
def test_observations_world(scope, position_encoding):
    gym_env = make_single(player_names=["terminator"], map_name="bridge", agent_id="0", observation_scope=scope,
                          observation_position_encoding=position_encoding, debug=False)
    observation = gym_env.get_observation()
    map_size = gym_env.game.world.size
    channels = 3 if position_encoding == "channels" else 1
    assert observation.shape == (channels, map_size[1], map_size[0])
    assert observation.dtype == np.int32
    assert gym_env.observation_space.shape == observation.shape


@pytest.mark.parametrize("scope,position_encoding", [("surroundings:11", "simple"), ("surroundings:11", "channels")])
def test_observations_surroundings(scope, position_encoding):
    gym_env = make_single(player_names=["terminator"], map_name="bridge", agent_id="0", observation_scope=scope,
                          observation_position_encoding=position_encoding, debug=False)
    observation = gym_env.get_observation()
    channels = 3 if position_encoding == "channels" else 1
    assert observation.shape == (channels, 11, 11)


def test_registered_ids_and_time_limit():
    from libzombsole_b200 import gym_env as ge
    assert set(ge.REGISTERED) == {"jvstinian/Zombsole-v0", "jvstinian/Zombsole-SurroundingsView-v0"}
    env = ge.make_vector("jvstinian/Zombsole-SurroundingsView-v0", num_envs=8)
    assert env.obs.shape == (8, 1, 21, 21)
    assert env.cfg.max_episode_steps == 1000
    obs, reward, term, trunc, info = env.step(np.full(8, 4))
    assert obs.shape == (8, 1, 21, 21) and reward.shape == (8,) and term.dtype.is_floating_point is False
    env.close()


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match="is not a valid rule name"):
        make_single(rules_name="nope")
    with pytest.raises(ValueError, match="is not a valid player weapon name"):
        make_single(agent_weapon="bazooka")
    with pytest.raises(ValueError, match="is not a valid observation scope"):
        make_single(observation_scope="galaxy")
    with pytest.raises(ValueError, match="odd number greater than 1"):
        make_single(observation_scope="surroundings:4")
    with pytest.raises(ValueError, match="render_mode=x is not supported"):
        make_single(render_mode="x")


# ---- tests/test_multiagent_env.py ------------------------------------------------------------------
def multi(agent_ids, map_name="boxed", initial_zombies=1, player_names=(), discrete=False):
    from libzombsole_b200.gym.multiagent_env import MultiagentZombsoleEnv, MultiagentZombsoleEnvDiscreteAction
    cls = MultiagentZombsoleEnvDiscreteAction if discrete else MultiagentZombsoleEnv
    return cls("extermination", list(player_names), map_name, agent_ids, initial_zombies=initial_zombies,
               minimum_zombies=0, render_mode=None, observation_surroundings_width=21, debug=True)


def test_multiagent_env_shape():
    env = multi([0], player_names=["terminator"])
    observation = env.get_observation()
    map_size = env.game.world.size
    expected = (3, max(map_size[1], 21), max(map_size[0], 21))
    assert len(observation) == 1
    for spobs in observation.values():
        assert spobs.shape == expected


def test_multiagent_1pgame():
    env1p = multi(["0"])
    stepcount = 0
    while True:
        _, _, done, truncated, _ = env1p.step({"0": {"action_type": "attack_closest", "parameter": [0, 0]}})
        if all(done.values()) or all(truncated.values()) or (stepcount >= 10):
            break
        stepcount += 1
    assert stepcount < 10


def test_multiagent_targeted_heal():
    env2p = multi(["0", "1"])
    env2p.game.agents[1].life = 25
    agent1pos = env2p.game.agents[1].position
    agent0pos = env2p.game.agents[0].position
    relativepos = (agent1pos[0] - agent0pos[0], agent1pos[1] - agent0pos[1])
    _ = env2p.step({"0": {"action_type": "heal", "parameter": relativepos}})
    assert env2p.game.agents[1].agent_id == "1"
    assert env2p.game.agents[1].life > 25


def test_multiagent_large_game():
    env32p = multi(list(map(str, range(0, 32))), map_name="fort", initial_zombies=100)
    stepcount = 0
    while True:
        _, _, done, truncated, _ = env32p.step({str(idx): {"action_type": "attack_closest", "parameter": [0, 0]}
                                                for idx in range(0, 32)})
        if all(done.values()) or all(truncated.values()) or (stepcount >= 200):
            break
        stepcount += 1
    assert True


def test_multiagent_discrete_action_game():
    env4p = multi([str(i) for i in range(0, 4)], map_name="fort", initial_zombies=100, discrete=True)
    stepcount = 0
    agent_ids = env4p.env.possible_agents
    while True:
        obs, _, done, truncated, _ = env4p.step({agent_id: env4p.action_spaces[agent_id].sample() for agent_id in agent_ids})
        if all(done.values()) or all(truncated.values()) or (stepcount >= 200):
            break
        stepcount += 1
    assert True


def test_multiagent_simple_encoding_raises_like_reference():
    from libzombsole_b200.gym.multiagent_env import MultiagentZombsoleEnv
    with pytest.raises(AttributeError):
        MultiagentZombsoleEnv("extermination", [], "boxed", ["0"], initial_zombies=1,
                              observation_position_encoding_style="simple")
