"""The C-ABI library loads without a GPU and exports every symbol include/zs_b200.h declares;
zs_layout (pure host arithmetic) agrees with the documented state layout."""
import ctypes as C
import os
import re

import pytest

import parity_util as pu
from libzombsole_b200 import abi, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zs_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _native.lib()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(L, name), name
        assert name in abi.PROTOTYPES, "abi.py has no prototype for %s" % name
    assert L.zs_abi_version() == abi.ZS_ABI_VERSION


def test_struct_sizes_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "zs_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(ZsMap), '
                   'sizeof(ZsConfig), sizeof(ZsLayout));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    m, c, l = (int(v) for v in subprocess.check_output([str(exe)]).split())
    assert (m, c, l) == (C.sizeof(abi.ZsMap), C.sizeof(abi.ZsConfig), C.sizeof(abi.ZsLayout))


@pytest.mark.parametrize("name", ["c1_bridge_ext", "c3_city_evac", "c4_maze_safehouse", "multi_fort_32p"])
def test_layout(name):
    L = _native.lib()
    cfg, m = pu.build(pu.CONFIGS[name], 64, 0)
    lay = abi.ZsLayout()
    ma = abi.MapArg(m)
    assert L.zs_layout(C.byref(cfg), C.byref(ma.struct), C.byref(lay)) == 0
    assert lay.n_slots == cfg.n_bots + cfg.n_agents + max(cfg.initial_zombies, cfg.minimum_zombies)
    assert lay.slot_pitch % 16 == 0 and lay.slot_pitch >= lay.n_slots
    assert lay.cells == m.size[0] * m.size[1]
    for f in range(abi.F_COUNT):
        assert lay.offset[f] % 256 == 0 and lay.row_bytes[f] % 16 == 0
        if f:
            assert lay.offset[f] >= lay.offset[f - 1] + 64 * lay.row_bytes[f - 1]
    assert lay.state_bytes >= lay.offset[abi.F_COUNT - 1] + 64 * lay.row_bytes[abi.F_COUNT - 1]
    per_obs = lay.obs_channels * lay.obs_height * lay.obs_width
    assert lay.obs_elems_per_env == per_obs * (cfg.n_agents if cfg.obs_per_agent else 1)


def test_layout_rejects_bad_config():
    L = _native.lib()
    cfg, m = pu.build(pu.CONFIGS["c1_bridge_ext"], 4, 0)
    ma = abi.MapArg(m)
    lay = abi.ZsLayout()
    cfg.rules = 9
    assert L.zs_layout(C.byref(cfg), C.byref(ma.struct), C.byref(lay)) != 0
    assert b"rules" in L.zs_last_error()
    cfg, _ = pu.build(pu.CONFIGS["c1_bridge_ext"], 4, 0)
    cfg.rules = abi.RULES["safehouse"]
    boxed = abi.MapArg(abi.resolve_map("boxed"))
    assert L.zs_layout(C.byref(cfg), C.byref(boxed.struct), C.byref(lay)) != 0
    assert b"objectives" in L.zs_last_error()  # safehouse.py:29-30


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from libzombsole_b200.engine import ZsEngine
    cfg, m = pu.build(pu.CONFIGS["c1_bridge_ext"], 4, 0)
    with pytest.raises(_native.ZsError):
        ZsEngine(cfg, m)
