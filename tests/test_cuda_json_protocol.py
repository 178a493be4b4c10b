"""The JSON-over-stdio server (libzombsole_b200/interactive_json.py) against transcripts of the reference's own
``GymEnvManager.run()`` (zombsole/interactive_json.py:205-338) under the same draws: every response line must be
byte-identical — protocol tags, error texts, status quirks, observations, rewards and flags."""
import io
import json
import os

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["single", "multi"])
def test_json_session_matches_reference_transcript(name):
    from libzombsole_b200.interactive_json import GymEnvManager
    with open(os.path.join(GOLDEN, "json_session_%s.json" % name)) as f:
        g = json.load(f)
    out = io.StringIO()
    mgr = GymEnvManager(None, g["multi"], instream=io.StringIO("\n".join(g["requests"]) + "\n"), outstream=out,
                        seed=g["seed"], env_index_base=g["env_index"])
    mgr.run()
    got = out.getvalue().splitlines()
    assert len(got) == len(g["responses"])
    for i, (a, b) in enumerate(zip(got, g["responses"])):
        assert a == b, "response %d differs (request %r)" % (i, (["<start>"] + g["requests"])[i][:80])
    if mgr.gym_env is not None:
        mgr.gym_env.close()


def _requests(multi, world, n_actions):
    """A scripted client: StartGame, a few actions (different per world), a status."""
    env = {} if world is None else {"env": world}
    w = 0 if world is None else world
    lines = [json.dumps(dict({"tag": "StartGame"}, **env))]
    single = [{"action_type": "move", "parameter": [0, 1]}, {"action_type": "attack_closest"}, {"action_type": "heal"},
              {"action_type": "move", "parameter": [-1, 0]}, {"action_type": "attack", "parameter": [1, 0]}]
    for t in range(n_actions):
        if multi:
            a = {"0": single[(t + w) % 5], "1": single[(2 * t + w) % 5]} if (t + w) % 4 else {"1": {"action_type": "heal_closest"}}
            a = {k: dict(v, parameter=v.get("parameter", [0, 0])) for k, v in a.items()}
        else:
            a = single[(t + w) % 5]
        lines.append(json.dumps(dict({"tag": "GameAction", "parameters": a}, **env)))
    lines.append(json.dumps(dict({"tag": "GameStatus"}, **env)))
    return lines


@pytest.mark.parametrize("multi", [False, True])
def test_worlds_of_one_batch_answer_like_separate_servers(multi):
    """World i of a batched server (requests with "env": i, interleaved with the other worlds' requests) answers exactly
    like a one-world server whose world has the same global env index: a masked reset / masked step touches one world."""
    from libzombsole_b200.interactive_json import BatchedJsonServer
    N, base, seed, T = 5, 40, 77, 14
    params = ({"rules_name": "extermination", "map_name": "boxed", "players": [], "agent_ids": ["0", "1"],
               "initial_zombies": 3, "minimum_zombies": 0, "observation_scope": "surroundings:5"} if multi else
              {"rules_name": "extermination", "map_name": "bridge", "players": ["terminator"], "agent_ids": [0],
               "initial_zombies": 6, "minimum_zombies": 0})
    config = json.dumps({"tag": "GameConfigUpdate", "parameters": params})
    per_world = [_requests(multi, w, T - w) for w in range(N)]
    # interleave the worlds' scripts round-robin
    script, cursor = [config], [0] * N
    while any(cursor[w] < len(per_world[w]) for w in range(N)):
        for w in range(N):
            if cursor[w] < len(per_world[w]):
                script.append(per_world[w][cursor[w]])
                cursor[w] += 1
    script.append('{"tag": "Exit"}')
    out = io.StringIO()
    srv = BatchedJsonServer(None, multi, instream=io.StringIO("\n".join(script) + "\n"), outstream=out, num_envs=N,
                            seed=seed, env_index_base=base)
    srv.run()
    got = out.getvalue().splitlines()[2:-1]  # (initial state, the config's answer, ..., exit)
    assert len(got) == sum(len(p) for p in per_world)
    answers = [[] for _ in range(N)]
    cursor, k = [0] * N, 0
    while k < len(got):
        for w in range(N):
            if cursor[w] < len(per_world[w]):
                answers[w].append(got[k])
                cursor[w] += 1
                k += 1
    srv.gym_env.close()
    for w in range(N):
        solo_out = io.StringIO()
        solo_script = [config] + _requests(multi, None, T - w) + ['{"tag": "Exit"}']
        # the solo client's action tape must be world w's: _requests varies the actions with the world index
        solo_script = [config] + [json.dumps({k: v for k, v in json.loads(l).items() if k != "env"}) for l in per_world[w]] + ['{"tag": "Exit"}']
        solo = BatchedJsonServer(None, multi, instream=io.StringIO("\n".join(solo_script) + "\n"), outstream=solo_out,
                                 num_envs=1, seed=seed, env_index_base=base + w)
        solo.run()
        want = solo_out.getvalue().splitlines()[2:-1]
        solo.gym_env.close()
        assert answers[w] == want, "world %d of the batch differs from a one-world server" % w


def test_all_worlds_in_one_request():
    from libzombsole_b200.interactive_json import BatchedJsonServer
    N = 4
    lines = [json.dumps({"tag": "GameConfigUpdate", "parameters": {"rules_name": "extermination", "map_name": "bridge",
                                                                   "players": [], "agent_ids": [0], "minimum_zombies": 0}}),
             json.dumps({"tag": "StartGame", "env": "all"}),
             json.dumps({"tag": "GameAction", "env": "all", "parameters": [{"action_type": "attack_closest"}] * N}),
             json.dumps({"tag": "GameAction", "env": 9, "parameters": {"action_type": "heal"}}),
             json.dumps({"tag": "GameAction", "env": "all", "parameters": [{"action_type": "heal"}]}),
             '{"tag": "Exit"}']
    out = io.StringIO()
    srv = BatchedJsonServer(None, False, instream=io.StringIO("\n".join(lines) + "\n"), outstream=out, num_envs=N)
    srv.run()
    got = [json.loads(l) for l in out.getvalue().splitlines()]
    srv.gym_env.close()
    assert got[2]["tag"] == "GameObservations" and len(got[2]["parameters"]) == N
    assert got[3]["tag"] == "GameObservations" and all(p["info"] == {} for p in got[3]["parameters"])
    assert got[4]["tag"] == "Error" and got[5]["tag"] == "Error"
    assert got[6]["parameters"]["status"] == "exiting"
