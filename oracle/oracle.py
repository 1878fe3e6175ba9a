"""ctypes binding of oracle/_build/liboracle.so (flatland_oracle.c) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "flatland_oracle.c")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        L.fo_create.restype = P
        L.fo_create.argtypes = [C.c_int] * 4 + [P] * 7
        L.fo_free.argtypes = [P]
        L.fo_reset.argtypes = [P]
        L.fo_step.restype = C.c_int
        L.fo_step.argtypes = [P] * 5
        L.fo_obs.restype = C.c_int
        L.fo_obs.argtypes = [P] * 9
        L.fo_motion_check.argtypes = [C.c_int, P, P, P]
        L.fo_get_state.argtypes = [P] * 12
        for f in ("fo_elapsed", "fo_done_all", "fo_num_targets"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [P]
        L.fo_target_slot.restype = C.c_int
        L.fo_target_slot.argtypes = [P, C.c_int]
        L.fo_get_dist_u16.argtypes = [P, C.c_int, P]
        L.fo_bench_run.restype = C.c_longlong
        L.fo_bench_run.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_uint32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """One environment of the C oracle, built from a generated world (dict of numpy arrays with
    the keys of tests/golden/*.npz: H W N T grid init_pos init_dir target speed earliest latest)."""

    def __init__(self, world):
        L = lib()
        self.H, self.W, self.N, self.T = (int(world[k]) for k in ("H", "W", "N", "T"))
        self._keep = [np.ascontiguousarray(world["grid"], np.uint16),
                      np.ascontiguousarray(world["init_pos"], np.int16),
                      np.ascontiguousarray(world["init_dir"], np.uint8),
                      np.ascontiguousarray(world["target"], np.int16),
                      np.ascontiguousarray(world["speed"], np.float64),
                      np.ascontiguousarray(world["earliest"], np.int32),
                      np.ascontiguousarray(world["latest"], np.int32)]
        self.h = L.fo_create(self.H, self.W, self.N, self.T, *[_p(a) for a in self._keep])

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.fo_free(self.h)
            self.h = None

    def reset(self):
        lib().fo_reset(self.h)

    def step(self, actions, sched_row):
        actions = np.ascontiguousarray(actions, np.uint8)
        sched_row = np.ascontiguousarray(sched_row, np.uint8)
        rewards = np.zeros(self.N, np.int32)
        dones = np.zeros(self.N + 1, np.uint8)
        rc = lib().fo_step(self.h, _p(actions), _p(sched_row), _p(rewards), _p(dones))
        if rc != 0:
            raise Exception("Episode is done, cannot call step()")
        return rewards, dones

    def obs(self):
        n = self.N
        o = dict(attr=np.zeros((n, 83), np.float32), forest=np.zeros((n, 31, 12), np.float32),
                 adjacency=np.zeros((n, 30, 3), np.int32), node_order=np.zeros((n, 31), np.int32),
                 edge_order=np.zeros((n, 30), np.int32), valid_actions=np.zeros((n, 5), np.uint8),
                 dist_target=np.zeros(n, np.float32), deadlocked=np.zeros(n, np.uint8))
        rc = lib().fo_obs(self.h, *[_p(o[k]) for k in ("attr", "forest", "adjacency", "node_order",
                                                        "edge_order", "valid_actions", "dist_target",
                                                        "deadlocked")])
        if rc != 0:
            raise ValueError("WRONG CELL TYPE detected in tree-search")
        return o

    def state(self):
        n = self.N
        s = dict(pos=np.zeros((n, 2), np.int16), dir=np.zeros(n, np.uint8), state=np.zeros(n, np.uint8),
                 ctr=np.zeros(n, np.uint8), mal=np.zeros(n, np.uint8), nmal=np.zeros(n, np.uint16),
                 saved=np.zeros(n, np.uint8), arrival=np.zeros(n, np.int32),
                 old_pos=np.zeros((n, 2), np.int16), old_dir=np.zeros(n, np.int8),
                 sig_mal=np.zeros(n, np.uint8))
        lib().fo_get_state(self.h, *[_p(s[k]) for k in ("pos", "dir", "state", "ctr", "mal", "nmal", "saved",
                                                         "arrival", "old_pos", "old_dir", "sig_mal")])
        return s

    @property
    def elapsed(self):
        return lib().fo_elapsed(self.h)

    @property
    def done_all(self):
        return bool(lib().fo_done_all(self.h))

    def target_slots(self):
        return np.array([lib().fo_target_slot(self.h, i) for i in range(self.N)], np.int16)

    def dist_u16(self):
        ns = lib().fo_num_targets(self.h)
        out = np.zeros((ns, self.H, self.W, 4), np.uint16)
        for s in range(ns):
            lib().fo_get_dist_u16(self.h, s, _p(out[s]))
        return out


def motion_check(cur, nxt):
    cur = np.ascontiguousarray(cur, np.int16)
    nxt = np.ascontiguousarray(nxt, np.int16)
    n = cur.shape[0]
    out = np.zeros(n, np.uint8)
    lib().fo_motion_check(n, _p(cur), _p(nxt), _p(out))
    return out


def bench_run(envs, n_steps, n_threads, seed=1):
    arr = (C.c_void_p * len(envs))(*[e.h for e in envs])
    return lib().fo_bench_run(arr, len(envs), n_steps, n_threads, seed)
