"""Host-side part of the JSON-over-stdio protocol (no GPU): request decoding and the responses that do not need
an environment, against the reference transcripts (tests/golden/json_session_single.json)."""
import io
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_requests_without_an_environment_answer_like_the_reference():
    from libzombsole_b200.interactive_json import GymEnvManager
    with open(os.path.join(GOLDEN, "json_session_single.json")) as f:
        g = json.load(f)
    # the first four requests (status, malformed JSON, unknown tag, action without parameters) come before any config
    reqs = g["requests"][:4] + ['{"tag": "Exit"}']
    out = io.StringIO()
    GymEnvManager(None, False, instream=io.StringIO("\n".join(reqs) + "\n"), outstream=out).run()
    got = out.getvalue().splitlines()
    assert got[:5] == g["responses"][:5]
    last = json.loads(got[5])
    assert last == {"tag": "GameState", "parameters": {"status": "exiting", "active": False, "config_required": True,
                                                        "last_observation": None}}


def test_game_config_defaults_are_the_reference_ones():
    from libzombsole_b200.interactive_json import GameConfig
    c = GameConfig.from_dict({"rules_name": "extermination", "map_name": "bridge", "players": [], "agent_ids": [0]})
    assert (c.initial_zombies, c.minimum_zombies, c.observation_scope, c.observation_position_encoding) == (10, 10, "world", "simple")
