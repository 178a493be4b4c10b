"""Map loading: the reference's text map format compiled into static tables.

Follows ``Map.from_file`` (zombsole/game.py:44-97) exactly:

* the file is split on ``\\n``; empty lines are skipped but keep their row index;
* position = (column, row); ``width = 1 + max column index of ANY character``
  (blanks included), ``height = 1 + index of the last non-empty row``;
* ``U+2612``/``b``/``B`` -> Box, ``U+2593``/``w``/``W`` -> Wall, ``p``/``P`` player spawn,
  ``z``/``Z`` zombie spawn, ``o``/``O`` objective (position AND an ObjectiveLocation
  decoration); everything else is empty ground;
* ``Map.things`` (here: ``statics``) is in row-major file order.

The 13 stock maps the reference ships (zombsole/maps/*) are input data of the env
API (``map_name="bridge"``); they are stored run-length encoded in
``maps/stock_maps.json`` (tools/pack_maps.py) and expanded back to map text here.
"""
import json
import os
import re

import numpy as np

BOX_ICON = u"☒"
WALL_ICON = u"▓"
LABEL_BOX = 1
LABEL_WALL = 4

_STOCK = None
_STOCK_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps", "stock_maps.json")


def stock_map_names():
    return sorted(_stock().keys())


def _stock():
    global _STOCK
    if _STOCK is None:
        with open(_STOCK_PATH) as f:
            _STOCK = json.load(f)
    return _STOCK


def stock_map_text(map_name):
    """Expand a packed stock map back into map-file text (ASCII spellings)."""
    rows = _stock().get(map_name)
    if rows is None:
        raise FileNotFoundError("no stock map named %r (have: %s)" % (map_name, ", ".join(stock_map_names())))
    out = []
    for row in rows:
        out.append("".join((" " if ch == "." else ch) * int(n) for n, ch in re.findall(r"(\d+)(\D)", row)))
    return "\n".join(out)


class Map(object):
    """A parsed map (same attribute names as the reference's ``Map``, game.py:34-42).

    size            (width, height)
    statics         list of (x, y, label) for boxes/walls in file order
    player_spawns   list of (x, y)
    zombie_spawns   list of (x, y)
    objectives      list of (x, y)
    """

    def __init__(self, size, statics, player_spawns, zombie_spawns, objectives, name=None):
        self.size = size
        self.statics = statics
        self.player_spawns = player_spawns
        self.zombie_spawns = zombie_spawns
        self.objectives = objectives
        self.name = name

    @classmethod
    def from_text(cls, text, name=None):
        statics, player_spawns, zombie_spawns, objectives = [], [], [], []
        max_row = 0
        max_col = 0
        for row_index, line in enumerate(text.split("\n")):
            if not line:
                continue
            max_row = row_index
            for col_index, char in enumerate(line):
                max_col = max(col_index, max_col)
                position = (col_index, row_index)
                if char in (BOX_ICON, "b", "B"):
                    statics.append(position + (LABEL_BOX,))
                elif char in (WALL_ICON, "w", "W"):
                    statics.append(position + (LABEL_WALL,))
                elif char.lower() == "p":
                    player_spawns.append(position)
                elif char.lower() == "z":
                    zombie_spawns.append(position)
                elif char.lower() == "o":
                    objectives.append(position)
        return cls((max_col + 1, max_row + 1), statics, player_spawns, zombie_spawns, objectives, name=name)

    @classmethod
    def from_file(cls, file_path):
        with open(file_path, encoding="utf-8") as map_file:
            return cls.from_text(map_file.read(), name=os.path.basename(file_path))

    @classmethod
    def from_map_name(cls, map_name):
        """A stock map by name, or a path to a map file (the reference resolves names inside its
        own maps/ directory, gym_env.py:54-56; a path is accepted here as well)."""
        if map_name in _stock():
            return cls.from_text(stock_map_text(map_name), name=map_name)
        if os.path.isfile(map_name):
            return cls.from_file(map_name)
        raise FileNotFoundError("no stock map or map file named %r" % (map_name,))

    # ---- tables handed to the C ABI (ZsMap) --------------------------------
    def tables(self):
        def xy(lst):
            return np.ascontiguousarray(np.array([p[:2] for p in lst], dtype=np.int16).reshape(-1, 2))
        return {
            "static_xy": xy(self.statics),
            "static_label": np.ascontiguousarray(np.array([s[2] for s in self.statics], dtype=np.uint8)),
            "player_spawn_xy": xy(self.player_spawns),
            "zombie_spawn_xy": xy(self.zombie_spawns),
            "objective_xy": xy(self.objectives),
        }
