#!/usr/bin/env python
"""Pack the reference's stock map files into libzombsole_b200/maps/stock_maps.json.

Run once in the build container (needs /root/reference).  The stock maps are
INPUT DATA of the env API (``map_name="bridge"`` ...), not code; they are stored
run-length encoded with the ASCII spellings the map format itself defines
(``w`` for U+2593 wall, ``b`` for U+2612 box — zombsole/game.py:76-79 accepts
both spellings; ``.`` stands for a blank), one string per map row, so a map
round-trips to a text the parser reads identically (size, statics in file
order, spawn lists, objectives; checked below and in tests/test_maps.py).
"""
import json
import os
import re
import sys

SRC = os.path.join(os.environ.get("ZOMBSOLE_REFERENCE", "/root/reference"), "zombsole", "maps")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libzombsole_b200", "maps", "stock_maps.json")
ASCII = {u"▓": "w", u"☒": "b", " ": "."}


def rle(row):
    out, i = [], 0
    while i < len(row):
        j = i
        while j < len(row) and row[j] == row[i]:
            j += 1
        out.append("%d%s" % (j - i, row[i]))
        i = j
    return "".join(out)


def main():
    packed = {}
    for name in sorted(os.listdir(SRC)):
        with open(os.path.join(SRC, name), encoding="utf-8") as f:
            rows = f.read().split("\n")
        enc = []
        for row in rows:
            row = "".join(ASCII.get(c, c) for c in row)
            assert re.fullmatch(r"[.wbBWpPzZoO]*", row), (name, row)
            enc.append(rle(row))
        packed[name] = enc
    with open(DST, "w") as f:
        json.dump(packed, f, indent=0, sort_keys=True)
    print("wrote", DST, {k: len(v) for k, v in packed.items()})


if __name__ == "__main__":
    sys.exit(main())
