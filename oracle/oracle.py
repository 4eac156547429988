"""ctypes binding of the CPU oracle (oracle/quoridor_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (alphazero_quoridor_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "quoridor_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64p, i32p, f64p, i64p = (C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                  C.POINTER(C.c_longlong))
        vp = C.c_void_p
        L.oq_sizeof_game.restype = C.c_int
        L.oq_reset.argtypes = [vp]
        L.oq_set_position.argtypes = [vp, C.c_uint64, C.c_uint64] + [C.c_int] * 5
        L.oq_get_position.argtypes = [vp, u64p, u64p, i32p]
        L.oq_actions.argtypes = [vp, i32p]
        L.oq_actions.restype = C.c_int
        L.oq_step.argtypes = [vp, C.c_int]
        L.oq_step.restype = C.c_int
        L.oq_has_a_winner.argtypes = [vp, i32p]
        L.oq_has_a_winner.restype = C.c_int
        L.oq_state.argtypes = [vp, f64p]
        L.oq_state.restype = C.c_int
        L.oq_valid_pawn_actions.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, i32p]
        L.oq_valid_pawn_actions.restype = C.c_int
        L.oq_legal_mask.argtypes = [vp, u64p]
        L.oq_state_key.argtypes = [vp]
        L.oq_state_key.restype = C.c_uint64
        L.oq_get_index_error.restype = C.c_int
        L.oq_mcts_new.argtypes = [C.c_int, C.c_double, C.c_int]
        L.oq_mcts_new.restype = vp
        L.oq_mcts_free.argtypes = [vp]
        L.oq_mcts_set_fix_terminal_sign.argtypes = [vp, C.c_int]
        L.oq_mcts_set_seed.argtypes = [vp, C.c_uint64]
        L.oq_mcts_set_rollout_counter.argtypes = [vp, C.c_uint64]
        L.oq_mcts_env_steps.argtypes = [vp]
        L.oq_mcts_env_steps.restype = C.c_longlong
        L.oq_mcts_run.argtypes = [vp, vp, i32p, i32p, f64p]
        L.oq_mcts_run.restype = C.c_int
        L.oq_mcts_root_stats.argtypes = [vp, i32p, f64p]
        L.oq_visits_to_probs.argtypes = [i32p, C.c_int, C.c_double, f64p]
        L.oq_mcts_update_with_move.argtypes = [vp, C.c_int]
        L.oq_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oq_sample_action.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32]
        L.oq_sample_action.restype = C.c_int
        L.oq_rollout.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int, i32p]
        L.oq_rollout.restype = C.c_int
        L.oq_bench_random_games.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, i64p, i32p]
        L.oq_bench_random_games.restype = C.c_longlong
        L.oq_bench_sweeps.argtypes = [u64p, u64p, i32p, C.c_int, C.c_int, u64p]
        L.oq_bench_sweeps.restype = C.c_longlong
        L.oq_bench_pure_mcts.argtypes = [u64p, u64p, i32p, C.c_int, C.c_int, C.c_double, C.c_uint64,
                                         C.c_int, i32p, i64p]
        L.oq_bench_pure_mcts.restype = C.c_longlong
        L.oq_bench_stub_mcts.argtypes = [u64p, u64p, i32p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                         i32p]
        L.oq_bench_stub_mcts.restype = C.c_longlong
        L.oq_max_threads.restype = C.c_int
        L.oq_set_literal_rollouts.argtypes = [C.c_int]
        L.oq_random_game.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int]
        L.oq_random_game.restype = C.c_int
        _lib = L
    return _lib


def _p(arr, ty):
    return arr.ctypes.data_as(C.POINTER(ty))


class OracleGame:
    """Mirror of the reference `Quoridor` surface (quoridor.py:5-610) over the C oracle."""

    def __init__(self):
        self._L = lib()
        self._buf = C.create_string_buffer(self._L.oq_sizeof_game())
        self._g = C.cast(self._buf, C.c_void_p)
        self.reset()

    def reset(self):
        self._L.oq_reset(self._g)

    def copy(self):
        o = OracleGame.__new__(OracleGame)
        o._L = self._L
        o._buf = C.create_string_buffer(self._buf.raw, len(self._buf))
        o._g = C.cast(o._buf, C.c_void_p)
        return o

    def set_position(self, H=0, V=0, p1=4, p2=76, w1=10, w2=10, cur=1):
        self._L.oq_set_position(self._g, H, V, p1, p2, w1, w2, cur)
        return self

    def position(self):
        H, V = C.c_uint64(), C.c_uint64()
        m = (C.c_int * 5)()
        self._L.oq_get_position(self._g, C.byref(H), C.byref(V), m)
        return dict(H=H.value, V=V.value, p1=m[0], p2=m[1], w1=m[2], w2=m[3], cur=m[4])

    def actions(self):
        out = (C.c_int * 140)()
        n = self._L.oq_actions(self._g, out)
        return list(out[:n])

    def legal_mask(self):
        m = (C.c_uint64 * 3)()
        self._L.oq_legal_mask(self._g, m)
        return [m[0], m[1], m[2]]

    def step(self, a):
        return bool(self._L.oq_step(self._g, int(a)))

    def has_a_winner(self):
        w = C.c_int()
        over = self._L.oq_has_a_winner(self._g, C.byref(w))
        return bool(over), (w.value if over else None)

    def state(self):
        out = np.zeros((26, 9, 9), dtype=np.float64)
        rc = self._L.oq_state(self._g, _p(out, C.c_double))
        if rc != 0:
            raise IndexError("reference undefined: pawn off the board")
        return out

    def random_game(self, seed, game_id, cap=3000):
        """Play self to the end with uniform picks from the full legal list (product's pick rule). -> plies."""
        return self._L.oq_random_game(self._g, seed, game_id, cap)

    def state_key(self):
        return self._L.oq_state_key(self._g)

    def sample_action(self, seed, rid, ply):
        return self._L.oq_sample_action(self._g, seed, rid, ply)

    def rollout(self, seed, rid, limit=1000):
        """Mutates self (like pure_mcts.py:86-108 mutates the deep copy). Returns (value, plies)."""
        plies = C.c_int()
        v = self._L.oq_rollout(self._g, seed, rid, limit, C.byref(plies))
        return v, plies.value


def valid_pawn_actions(H, V, loc, opp, player):
    walls = np.zeros(64, dtype=np.int8)
    for i in range(64):
        if (H >> i) & 1:
            walls[i] = 1
        elif (V >> i) & 1:
            walls[i] = -1
    out = (C.c_int * 12)()
    n = lib().oq_valid_pawn_actions(walls.tobytes(), loc, opp, player, out)
    return list(out[:n])


class OracleMCTS:
    """mcts.MCTS / pure_mcts.MCTS (stub_kind: 1 = S1, 2 = S2, 0 = pure MCTS with rollouts)."""

    def __init__(self, stub_kind, c_puct=5, n_playout=100, fix_terminal_sign=False, seed=0, rollout_counter=0):
        self._L = lib()
        self._t = C.c_void_p(self._L.oq_mcts_new(stub_kind, float(c_puct), n_playout))
        self._L.oq_mcts_set_fix_terminal_sign(self._t, int(fix_terminal_sign))
        self._L.oq_mcts_set_seed(self._t, seed)
        self._L.oq_mcts_set_rollout_counter(self._t, rollout_counter)

    def __del__(self):
        try:
            self._L.oq_mcts_free(self._t)
        except Exception:
            pass

    def run(self, game):
        """n_playout playouts from `game`; returns (acts, visits, q) in child insertion order."""
        acts, visits = (C.c_int * 140)(), (C.c_int * 140)()
        qs = (C.c_double * 140)()
        n = self._L.oq_mcts_run(self._t, game._g, acts, visits, qs)
        return list(acts[:n]), list(visits[:n]), list(qs[:n])

    def root_stats(self):
        n, q = C.c_int(), C.c_double()
        self._L.oq_mcts_root_stats(self._t, C.byref(n), C.byref(q))
        return n.value, q.value

    def update_with_move(self, move):
        self._L.oq_mcts_update_with_move(self._t, int(move))

    def env_steps(self):
        return self._L.oq_mcts_env_steps(self._t)


def visits_to_probs(visits, temp):
    v = np.asarray(visits, dtype=np.int32)
    out = np.zeros(len(v), dtype=np.float64)
    lib().oq_visits_to_probs(_p(v, C.c_int), len(v), float(temp), _p(out, C.c_double))
    return out


def philox(seed, rid, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().oq_philox(seed, rid, c2, c3, out)
    return list(out)


# ------------------------------------------------------------------ batched drivers (bench / tests)
def _pos_arrays(H, V, meta5):
    H = np.ascontiguousarray(H, dtype=np.uint64)
    V = np.ascontiguousarray(V, dtype=np.uint64)
    meta5 = np.ascontiguousarray(meta5, dtype=np.int32).reshape(-1, 5)
    assert len(H) == len(V) == len(meta5)
    return H, V, meta5


def bench_random_games(n_games, seed=0, cap=3000, threads=0):
    plies = np.zeros(n_games, dtype=np.int64)
    winner = np.zeros(n_games, dtype=np.int32)
    total = lib().oq_bench_random_games(n_games, seed, cap, threads, _p(plies, C.c_longlong), _p(winner, C.c_int))
    return total, plies, winner


def sweeps(H, V, meta5, threads=0):
    H, V, meta5 = _pos_arrays(H, V, meta5)
    mask = np.zeros((len(H), 3), dtype=np.uint64)
    total = lib().oq_bench_sweeps(_p(H, C.c_uint64), _p(V, C.c_uint64), _p(meta5, C.c_int), len(H), threads,
                                  _p(mask, C.c_uint64))
    return total, mask


def pure_mcts_moves(H, V, meta5, n_playout, c_puct=5.0, seed=0, threads=0, literal_rollouts=False):
    """literal_rollouts=True: every rollout ply runs the full actions() sweep, as pure_mcts.py:7-10 does."""
    H, V, meta5 = _pos_arrays(H, V, meta5)
    lib().oq_set_literal_rollouts(int(literal_rollouts))
    moves = np.zeros(len(H), dtype=np.int32)
    playouts = C.c_longlong()
    steps = lib().oq_bench_pure_mcts(_p(H, C.c_uint64), _p(V, C.c_uint64), _p(meta5, C.c_int), len(H), n_playout,
                                     float(c_puct), seed, threads, _p(moves, C.c_int), C.byref(playouts))
    return steps, playouts.value, moves


def stub_mcts_visits(H, V, meta5, n_playout, c_puct=5.0, stub_kind=2, threads=0):
    H, V, meta5 = _pos_arrays(H, V, meta5)
    visits = np.zeros((len(H), 140), dtype=np.int32)
    sims = lib().oq_bench_stub_mcts(_p(H, C.c_uint64), _p(V, C.c_uint64), _p(meta5, C.c_int), len(H), n_playout,
                                    float(c_puct), stub_kind, threads, _p(visits, C.c_int))
    return sims, visits


def max_threads():
    return lib().oq_max_threads()
