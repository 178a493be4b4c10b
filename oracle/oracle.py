"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/libzs_oracle.so (the CPU oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from libzombsole_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzs_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "zs_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "zs_b200.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libzs_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.zso_create.restype = C.c_void_p
        L.zso_create.argtypes = [C.POINTER(abi.ZsConfig), C.POINTER(abi.ZsMap)]
        L.zso_destroy.argtypes = [C.c_void_p]
        L.zso_obs_elems.restype = C.c_int64
        L.zso_obs_elems.argtypes = [C.c_void_p]
        L.zso_n_slots.restype = C.c_int32
        L.zso_n_slots.argtypes = [C.c_void_p]
        L.zso_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.zso_encode_obs.argtypes = [C.c_void_p, C.c_void_p]
        L.zso_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 6
        L.zso_synthetic_actions.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.zso_rollout_synthetic.argtypes = [C.c_void_p, C.c_int32, C.c_int64] + [C.c_void_p] * 4
        L.zso_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.zso_export.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
        L.zso_set_life.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.zso_set_static_life.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.zso_set_threads.restype = C.c_int32
        L.zso_set_threads.argtypes = [C.c_int32]
        L.zso_philox4x32_10.argtypes = [C.c_void_p] * 3
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


class OracleEnv(object):
    """N environments stepped by the C oracle; same config/map structs as the CUDA library."""

    def __init__(self, cfg, map_):
        self.L = lib()
        self.cfg = cfg
        self.map_arg = abi.MapArg(abi.resolve_map(map_))
        self.h = self.L.zso_create(C.byref(cfg), C.byref(self.map_arg.struct))
        if not self.h:
            raise RuntimeError("zso_create failed")
        self.N = cfg.num_envs
        self.A = cfg.n_agents
        self.M = self.L.zso_n_slots(self.h)
        self.S = len(self.map_arg.map.statics)
        self.cells = self.map_arg.map.size[0] * self.map_arg.map.size[1]
        self.obs_elems = self.L.zso_obs_elems(self.h)
        self.R = self.A if cfg.obs_per_agent else 1

    def close(self):
        if self.h:
            self.L.zso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask=None):
        obs = np.zeros((self.N, self.obs_elems), np.int32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.zso_reset(self.h, _p(m), _p(obs))
        return obs

    def encode_obs(self):
        obs = np.zeros((self.N, self.obs_elems), np.int32)
        self.L.zso_encode_obs(self.h, _p(obs))
        return obs

    def step(self, actions, fmt):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.N, self.obs_elems), np.int32)
        reward = np.zeros((self.N, self.R), np.float64)
        term = np.zeros(self.N, np.uint8)
        trunc = np.zeros(self.N, np.uint8)
        mask = np.zeros((self.N, self.A), np.uint8)
        draws = np.zeros(self.N, np.int32)
        self.L.zso_step(self.h, _p(actions), fmt, _p(obs), _p(reward), _p(term), _p(trunc), _p(mask), _p(draws))
        return obs, reward, term, trunc, mask, draws

    def synthetic_actions(self, step_index):
        a = np.zeros((self.N, self.A), np.int32)
        self.L.zso_synthetic_actions(self.h, int(step_index), _p(a))
        return a

    def rollout_synthetic(self, n_steps, first_step_index=0, want_obs=True):
        obs = np.zeros((self.N, self.obs_elems), np.int32) if want_obs else None
        reward = np.zeros((n_steps, self.N, self.R), np.float64)
        term = np.zeros((n_steps, self.N), np.uint8)
        trunc = np.zeros((n_steps, self.N), np.uint8)
        self.L.zso_rollout_synthetic(self.h, n_steps, int(first_step_index), _p(obs), _p(reward), _p(term), _p(trunc))
        return obs, reward, term, trunc

    def stats(self, reset=False):
        out = np.zeros(4, np.int64)
        self.L.zso_stats(self.h, _p(out), 1 if reset else 0)
        return out

    def export(self, env):
        M, S = self.M, self.S
        r = {"x": np.zeros(M, np.int16), "y": np.zeros(M, np.int16), "life": np.zeros(M, np.int16),
             "in_world": np.zeros(M, np.uint8), "weapon": np.zeros(M, np.uint8), "order": np.zeros(M, np.int16),
             "static_life": np.zeros(S, np.int16), "static_present": np.zeros(S, np.uint8)}
        dead = np.zeros(self.cells, np.uint8)
        counters = np.zeros(8, np.int32)
        self.L.zso_export(self.h, env, _p(r["x"]), _p(r["y"]), _p(r["life"]), _p(r["in_world"]), _p(r["weapon"]),
                          _p(r["order"]), _p(r["static_life"]), _p(r["static_present"]), _p(dead), _p(counters))
        r["dead_body"] = np.packbits(dead, bitorder="little")
        r["counters"] = counters[:3].copy()
        r["episode"] = int(counters[3])
        r["episode_steps"] = int(counters[4])
        r["reset_draws"] = int(counters[5])
        r["step_draws"] = int(counters[6])
        return r

    def set_life(self, env, slot, life):
        self.L.zso_set_life(self.h, env, slot, life)

    def set_static_life(self, env, index, life):
        self.L.zso_set_static_life(self.h, env, index, life)


def set_threads(n):
    return lib().zso_set_threads(int(n))
