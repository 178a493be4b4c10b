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
