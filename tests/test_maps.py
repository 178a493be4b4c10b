"""Map compiler vs the reference's Map.from_file (zombsole/game.py:44-97).

Known answers of the reference's own test (tests/test_map.py:6-10) plus a full comparison of
every stock map against the reference parser when the reference tree is present."""
import os

import pytest

from libzombsole_b200.maps import Map, stock_map_names, LABEL_WALL

REF_MAPS = "/root/reference/zombsole/maps"


@pytest.mark.parametrize("map_name,exp_map_size,exp_walls_count,exp_objs_count", [
    ("bridge", (111, 12), 182, 28),
    ("boxed", (15, 8), 14, 0),
    ("fort", (73, 21), 210, 0),
])
def test_map_read(map_name, exp_map_size, exp_walls_count, exp_objs_count):
    lmap = Map.from_map_name(map_name)
    assert lmap.size == exp_map_size
    assert sum(1 for s in lmap.statics if s[2] == LABEL_WALL) == exp_walls_count
    assert len(lmap.objectives) == exp_objs_count


def test_stock_maps_complete():
    assert stock_map_names() == sorted([
        "arduino", "boxed", "bridge", "city_for_evacuation", "city_for_safehouse", "easy_exit", "easy_exit_v2",
        "fort", "hallway", "maze_for_safehouse", "to_the_closet", "village_for_evacuation", "village_for_safehouse"])


def test_survey_map_stats():
    # sizes the BASELINE configs are quoted on
    b = Map.from_map_name("bridge")
    assert (len(b.statics), len(b.player_spawns), len(b.zombie_spawns)) == (198, 20, 53)
    m = Map.from_map_name("maze_for_safehouse")
    assert m.size == (72, 39) and len(m.statics) == 1501 and len(m.objectives) == 18 and len(m.zombie_spawns) == 0
    c = Map.from_map_name("city_for_evacuation")
    assert c.size == (94, 28) and len(c.statics) == 709 and len(c.player_spawns) == 30


def test_ragged_and_blank_lines(tmp_path):
    p = tmp_path / "m"
    p.write_text(u"w  p\n\n  ☒z   \nOo\n", encoding="utf-8")
    m = Map.from_file(str(p))
    assert m.size == (7, 4)  # blanks count for the width, the empty row keeps its index
    assert m.statics == [(0, 0, 4), (2, 2, 1)]
    assert m.player_spawns == [(3, 0)] and m.zombie_spawns == [(3, 2)] and m.objectives == [(0, 3), (1, 3)]


@pytest.mark.skipif(not os.path.isdir(REF_MAPS), reason="reference tree not present")
@pytest.mark.parametrize("name", stock_map_names())
def test_stock_map_equals_reference_parser(name):
    from oracle import ref_harness
    mods = ref_harness.import_reference()
    ref = mods["game"].Map.from_file(os.path.join(REF_MAPS, name))
    got = Map.from_map_name(name)
    assert got.size == ref.size
    ref_statics = [(t.position[0], t.position[1], 1 if t.name == "box" else 4) for t in ref.things if not t.is_decoration]
    assert got.statics == ref_statics
    assert got.player_spawns == ref.player_spawns
    assert got.zombie_spawns == ref.zombie_spawns
    assert got.objectives == ref.objectives
    # and parsing the reference's own UTF-8 file with our parser gives the same thing
    own = Map.from_file(os.path.join(REF_MAPS, name))
    assert (own.size, own.statics, own.player_spawns, own.zombie_spawns, own.objectives) == \
        (got.size, got.statics, got.player_spawns, got.zombie_spawns, got.objectives)
