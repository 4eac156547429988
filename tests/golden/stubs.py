"""Deterministic policy-value stubs used to drive MCTS parity.

The same two stubs exist three times, bit-for-bit identical by construction:
  * here, in Python, wrapped around the *reference* `Quoridor` object (golden generation),
  * in `oracle/quoridor_oracle.c` (`stub_eval`),
  * on the device in `alphazero_quoridor_b200/csrc/qz_mcts.cu` (`qz_stub_eval` kernel).

S1 "uniform": prior = float32(1/len(legal)), value = 0.0       -- maximises ties (ordering test)
S2 "hash":    prior = float32 in (0, 1/64], value in [-1, 1)   -- minimises ties (arithmetic test)
S3 "hash/8":  S2's priors, value = S2's value / 8               -- flatter trees, more near-ties

All S2 quantities are dyadic rationals that are exact in float32, so no rounding mode,
summation order or libm difference can creep in.
"""
import numpy as np

M64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15


def splitmix64(x):
    x = (x + GOLDEN) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def state_key(H, V, p1, p2, w1, w2, cur):
    """64-bit key of a position; p1/p2 are taken modulo 256 (two's complement int8)."""
    meta = (p1 & 0xFF) | ((p2 & 0xFF) << 8) | (w1 << 16) | (w2 << 24) | (cur << 32)
    return splitmix64(H ^ splitmix64(V ^ splitmix64(meta)))


def s2_prior(key, a):
    h = splitmix64((key + a * GOLDEN) & M64)
    return np.float32(((h >> 40) + 1)) * np.float32(2.0 ** -30)  # (0, 2^-6]


def s2_value(key):
    h = splitmix64(key ^ 0xABCDEF)
    return float(h >> 40) / float(1 << 23) - 1.0


def masks_of(game):
    """(H, V) u64 masks of a reference Quoridor object."""
    H = V = 0
    for ix, w in enumerate(game._intersections):
        if w == 1:
            H |= 1 << ix
        elif w == -1:
            V |= 1 << ix
    return H, V


def key_of(game):
    H, V = masks_of(game)
    return state_key(H, V, int(game._positions[1]), int(game._positions[2]),
                     int(game._player1_walls_remaining), int(game._player2_walls_remaining),
                     int(game.current_player))


def make_stub(kind, log=None):
    """Return a `policy_value_fn(game)` honouring policy_value_net.py:145-164's contract."""

    def s1(game):
        try:
            acts = game.actions()
        except IndexError:          # off-board terminal leaf: reference would crash
            return iter(()), 0.0
        n = max(len(acts), 1)
        p = np.float32(1.0) / np.float32(n)
        return zip(acts, [p] * len(acts)), 0.0

    def s2(game):
        try:
            acts = game.actions()
        except IndexError:
            return iter(()), 0.0
        key = key_of(game)
        return zip(acts, [s2_prior(key, a) for a in acts]), s2_value(key)

    def s3(game):
        probs, v = s2(game)
        return probs, v / 8.0

    return {"S1": s1, "S2": s2, "S3": s3}[kind]
