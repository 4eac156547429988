"""Bit-exactness of the product's rules header (alphazero_quoridor_b200/csrc/qz_rules.cuh) on the HOST.

qz_rules.cuh is `__host__ __device__`; tests/host_harness compiles the very same source with g++ so the
bitboard rules can be diffed against the oracle and the golden fixtures in a container with no GPU.
This harness is test-only: the product never loads it (the GPU parity tests in test_env_gpu.py go
through the C-ABI of the CUDA library instead).
"""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "qz_host.cpp")
HDR = os.path.join(os.path.dirname(HERE), "alphazero_quoridor_b200", "csrc", "qz_rules.cuh")
SO = os.path.join(HERE, "host_harness", "_build", "libqzhost.so")


@pytest.fixture(scope="module")
def qh():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    csrc = os.path.dirname(HDR)
    deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", SO, SRC])
    L = C.CDLL(SO)
    u64p = C.POINTER(C.c_uint64)
    L.qh_pack_meta.restype = C.c_uint64
    L.qh_pack_meta.argtypes = [C.c_int] * 5 + [C.c_uint, C.c_uint]
    L.qh_apply.argtypes = [u64p, C.c_int]
    L.qh_legal_mask.argtypes = [u64p, u64p]
    L.qh_pawn_moves.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
    L.qh_pawn_moves.restype = C.c_uint
    for fn in (L.qh_pawn_moves_ctx, L.qh_pawn_moves_info):
        fn.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
        fn.restype = C.c_uint
    L.qh_dirs.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p]
    L.qh_dirs_incremental.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_char_p]
    L.qh_spread8.argtypes = [C.c_uint64, C.POINTER(C.c_uint32)]
    L.qh_encode.argtypes = [u64p, C.POINTER(C.c_float)]
    L.qh_action_rank.argtypes = [u64p, C.c_int]
    L.qh_nth_bit64.argtypes = [C.c_uint64, C.c_int]
    return L


def mk_state(qh, H, V, p1, p2, w1, w2, cur, flags=0, ply=0):
    s = (C.c_uint64 * 3)(H, V, qh.qh_pack_meta(p1, p2, w1, w2, cur, flags, ply))
    return s


def unpack_meta(m):
    def i8(x):
        return x - 256 if x > 127 else x
    return dict(p1=i8(m & 0xFF), p2=i8((m >> 8) & 0xFF), w1=(m >> 16) & 0xFF, w2=(m >> 24) & 0xFF,
                cur=(m >> 32) & 0xFF, flags=(m >> 40) & 0xFF, ply=(m >> 48) & 0xFFFF)


def mask_to_ordered(mask3):
    """140-bit mask -> the reference's actions() ordering (quoridor.py:157,420-430)."""
    bits = mask3[0] | (mask3[1] << 64) | (mask3[2] << 128)
    out = [a for a in range(12) if (bits >> a) & 1]
    for ix in range(64):
        if (bits >> (12 + ix)) & 1:
            out.append(12 + ix)
        if (bits >> (76 + ix)) & 1:
            out.append(76 + ix)
    return out


def legal_list(qh, s):
    m = (C.c_uint64 * 3)()
    qh.qh_legal_mask(s, m)
    return mask_to_ordered([m[0], m[1], m[2]])


def rand_walls(rng, n):
    H = V = 0
    tries = 0
    while bin(H | V).count("1") < n and tries < 500:
        tries += 1
        ix = rng.randrange(64)
        r, c = divmod(ix, 8)
        if (H | V) >> ix & 1:
            continue
        if rng.random() < 0.5:
            if (c > 0 and H >> (ix - 1) & 1) or (c < 7 and H >> (ix + 1) & 1):
                continue
            H |= 1 << ix
        else:
            if (r > 0 and V >> (ix - 8) & 1) or (r < 7 and V >> (ix + 8) & 1):
                continue
            V |= 1 << ix
    return H, V


def test_delta_and_nth_bit(qh):
    assert [qh.qh_delta(a) for a in range(12)] == [9, -9, 1, -1, 18, -18, 2, -2, 10, 8, -8, -10]
    rng = random.Random(1)
    for _ in range(300):
        m = rng.getrandbits(64) | 1
        bits = [i for i in range(64) if m >> i & 1]
        k = rng.randrange(len(bits))
        assert qh.qh_nth_bit64(m, k) == bits[k]


def test_spread8(qh):
    rng = random.Random(2)
    for _ in range(2000):
        x = rng.getrandbits(64) if rng.random() < 0.7 else (1 << rng.randrange(64))
        w = (C.c_uint32 * 3)()
        qh.qh_spread8(x, w)
        got = w[0] | (w[1] << 32) | (w[2] << 64)
        want = 0
        for i in range(64):
            if x >> i & 1:
                want |= 1 << ((i // 8) * 9 + (i % 8))
        assert got == want


def test_direction_masks_match_reference_plain_moves(qh):
    """qz_dirs == bits N,S,E,W of _valid_pawn_actions with no adjacent opponent, for all 81 tiles."""
    rng = random.Random(3)
    for it in range(400):
        H, V = rand_walls(rng, rng.randrange(0, 30))
        buf = C.create_string_buffer(81)
        qh.qh_dirs(H, V, buf)
        for t in range(81):
            # player chosen so the goal-row specials (quoridor.py:295,297) stay out of the way
            player = 2 if t >= 72 else 1
            want = O.valid_pawn_actions(H, V, t, -100, player)
            want_bits = sum(1 << a for a in want if a < 4)
            assert buf.raw[t] == want_bits, (hex(H), hex(V), t)
        # incremental placement of one more wall == full rebuild
        ix = rng.randrange(64)
        if (H | V) >> ix & 1:
            continue
        vert = rng.random() < 0.5
        inc = C.create_string_buffer(81)
        qh.qh_dirs_incremental(H, V, ix, int(vert), inc)
        full = C.create_string_buffer(81)
        qh.qh_dirs(H | (0 if vert else 1 << ix), V | (1 << ix if vert else 0), full)
        assert inc.raw == full.raw


def test_corner_masks_match_scalar_corners(qh):
    """The branch-free per-corner masks (QzPawnCtx) reproduce quoridor.py:356-418 on every tile."""
    qh.qh_ctx_corners.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p, C.c_char_p]
    qh.qh_dirs_ctx.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p]
    rng = random.Random(12)
    for it in range(1500):
        H, V = rand_walls(rng, rng.randrange(0, 34))
        a, b = C.create_string_buffer(81), C.create_string_buffer(81)
        qh.qh_ctx_corners(H, V, a, b)
        assert a.raw == b.raw, (hex(H), hex(V))
        d1, d2 = C.create_string_buffer(81), C.create_string_buffer(81)
        qh.qh_dirs_ctx(H, V, d1)
        qh.qh_dirs(H, V, d2)
        assert d1.raw == d2.raw


def test_pawn_moves_golden(qh, pawn_cases):
    for H, V, loc, opp, player, want in pawn_cases:
        got = qh.qh_pawn_moves(H, V, loc, opp, player)
        assert [a for a in range(12) if got >> a & 1] == want
        got = qh.qh_pawn_moves_ctx(H, V, loc, opp, player)
        assert [a for a in range(12) if got >> a & 1] == want
        if 0 <= opp <= 80:
            got = qh.qh_pawn_moves_info(H, V, loc, opp, player)
            assert [a for a in range(12) if got >> a & 1] == want


def test_pawn_moves_vs_oracle_exhaustive_adjacency(qh):
    rng = random.Random(4)
    for it in range(300):
        H, V = rand_walls(rng, rng.randrange(0, 26))
        for L in range(81):
            for d in (9, -9, 1, -1):
                Op = L + d
                if not 0 <= Op <= 80:
                    continue
                for player in (1, 2):
                    want = O.valid_pawn_actions(H, V, L, Op, player)
                    got = qh.qh_pawn_moves(H, V, L, Op, player)
                    assert [a for a in range(12) if got >> a & 1] == want
                    got = qh.qh_pawn_moves_ctx(H, V, L, Op, player)
                    assert [a for a in range(12) if got >> a & 1] == want
                    got = qh.qh_pawn_moves_info(H, V, L, Op, player)
                    assert [a for a in range(12) if got >> a & 1] == want


def test_replay_traces(qh, traces):
    """Golden traces: ordered legal list, step, state planes -- ply by ply (quoridor.py:58-186)."""
    import hashlib
    s = (C.c_uint64 * 3)()
    planes = np.zeros((26, 9, 9), dtype=np.float32)
    for tr in traces:
        qh.qh_initial(s)
        for ply, rec in enumerate(tr["plies"]):
            m = unpack_meta(s[2])
            assert (s[0], s[1]) == (rec["H"], rec["V"])
            assert (m["p1"], m["p2"], m["w1"], m["w2"], m["cur"]) == (rec["p1"], rec["p2"], rec["w1"], rec["w2"], rec["cur"])
            assert m["ply"] == ply and not (m["flags"] & 1)
            assert legal_list(qh, s) == rec["actions"], (tr["policy"], tr["seed"], ply)
            if rec["state"] is not None:
                qh.qh_encode(s, planes.ctypes.data_as(C.POINTER(C.c_float)))
                assert hashlib.sha256(planes.astype(np.uint8).tobytes()).hexdigest() == rec["state"]
            if rec["action"] is None:
                break
            qh.qh_apply(s, rec["action"])
        fin = tr["final"]
        m = unpack_meta(s[2])
        assert (m["p1"], m["p2"], m["cur"]) == (fin["p1"], fin["p2"], fin["cur"])
        assert bool(m["flags"] & 1) == fin["done"] and ((m["flags"] >> 1) & 3) == fin["winner"]


def test_kat(qh, kat):
    for rec in kat["named"] + kat["synthetic"]:
        if rec["actions"] is None:
            continue
        s = mk_state(qh, rec["H"], rec["V"], rec["p1"], rec["p2"], rec["w1"], rec["w2"], rec["cur"])
        assert legal_list(qh, s) == rec["actions"], rec["name"]


def _random_position(rng):
    H, V = rand_walls(rng, rng.randrange(0, 21))
    while True:
        p1 = rng.randrange(0, 72)
        p2 = p1 + rng.choice([9, -9, 1, -1]) if rng.random() < 0.35 else rng.randrange(9, 81)
        if p1 != p2 and 9 <= p2 <= 80:
            break
    return H, V, p1, p2, rng.randrange(1, 11), rng.randrange(1, 11), rng.choice([1, 2])


def test_legal_mask_vs_oracle_random_positions(qh):
    """4000 synthetic positions incl. adjacent pawns, sealed regions and dense walls."""
    rng = random.Random(5)
    n_blocked = 0
    for it in range(4000):
        H, V, p1, p2, w1, w2, cur = _random_position(rng)
        s = mk_state(qh, H, V, p1, p2, w1, w2, cur)
        g = O.OracleGame().set_position(H, V, p1, p2, w1, w2, cur)
        want = g.actions()
        got = legal_list(qh, s)
        assert got == want, (hex(H), hex(V), p1, p2, w1, w2, cur)
        n_blocked += (bin(~(H | V) & (2 ** 64 - 1)).count("1") * 2) - sum(1 for a in want if a >= 12)
    assert n_blocked > 10000      # the path check really rejected candidates


def test_random_games_vs_oracle(qh):
    """Seeded random play through both engines: list, step, planes and rank compared every ply."""
    rng = random.Random(6)
    planes = np.zeros((26, 9, 9), dtype=np.float32)
    for game in range(60):
        s = (C.c_uint64 * 3)()
        qh.qh_initial(s)
        g = O.OracleGame()
        for ply in range(400):
            want = g.actions()
            m3 = (C.c_uint64 * 3)()
            qh.qh_legal_mask(s, m3)
            assert mask_to_ordered([m3[0], m3[1], m3[2]]) == want
            for i, a in enumerate(want):
                assert qh.qh_action_rank(m3, a) == i
            qh.qh_encode(s, planes.ctypes.data_as(C.POINTER(C.c_float)))
            assert np.array_equal(planes.astype(np.float64), g.state())
            if not want:
                break
            wall = [a for a in want if a >= 12]
            a = rng.choice(wall) if (wall and rng.random() < 0.3) else rng.choice(want)
            done = g.step(a)
            qh.qh_apply(s, a)
            m = unpack_meta(s[2])
            pos = g.position()
            assert (s[0], s[1], m["p1"], m["p2"], m["w1"], m["w2"], m["cur"]) == (
                pos["H"], pos["V"], pos["p1"], pos["p2"], pos["w1"], pos["w2"], pos["cur"])
            assert bool(m["flags"] & 1) == done
            if done:
                assert ((m["flags"] >> 1) & 3) == g.has_a_winner()[1]
                break


def test_philox_and_rollouts_match_oracle(qh):
    """The sampled-legality rollout (qz_sample.cuh) == oracle rollout built on the literal rules."""
    qh.qh_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    qh.qh_sample_action.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.c_uint32]
    qh.qh_rollout.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_int)]
    rng = random.Random(8)
    for _ in range(200):
        seed, rid, c2, c3 = rng.getrandbits(64), rng.getrandbits(64), rng.getrandbits(32), rng.getrandbits(32)
        out = (C.c_uint32 * 4)()
        qh.qh_philox(seed, rid, c2, c3, out)
        assert list(out) == O.philox(seed, rid, c2, c3)
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors): zero and all-ones inputs
    out = (C.c_uint32 * 4)()
    qh.qh_philox(0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    qh.qh_philox(M64, M64, 0xFFFFFFFF, 0xFFFFFFFF, out)
    assert [hex(x) for x in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    # single draws on walled positions
    for it in range(600):
        H, V, p1, p2, w1, w2, cur = _random_position(rng)
        s = mk_state(qh, H, V, p1, p2, w1, w2, cur)
        g = O.OracleGame().set_position(H, V, p1, p2, w1, w2, cur)
        seed, rid, ply = rng.getrandbits(64), rng.getrandbits(40), rng.randrange(0, 900)
        assert qh.qh_sample_action(s, seed, rid, ply) == g.sample_action(seed, rid, ply)
    # the table-driven and the capped variants used by the wall-phase kernel agree with the plain sampler
    qh.qh_sample_action_known.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.c_uint32]
    qh.qh_sample_action_capped.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
    n_capped = 0
    for it in range(1500):
        H, V = rand_walls(rng, rng.randrange(12, 21))
        _, _, p1, p2, w1, w2, cur = _random_position(rng)
        s = mk_state(qh, H, V, p1, p2, w1, w2, cur)
        seed, rid, ply = rng.getrandbits(64), rng.getrandbits(40), rng.randrange(0, 900)
        want = qh.qh_sample_action(s, seed, rid, ply)
        assert qh.qh_sample_action_known(s, seed, rid, ply) == want
        capped = qh.qh_sample_action_capped(s, seed, rid, ply, 2)
        assert capped == want or capped == -2
        n_capped += capped == -2
    assert n_capped > 10
    # whole rollouts from the start and from midgames
    for it in range(150):
        if it < 60:
            s = (C.c_uint64 * 3)()
            qh.qh_initial(s)
            g = O.OracleGame()
        else:
            H, V, p1, p2, w1, w2, cur = _random_position(rng)
            s = mk_state(qh, H, V, p1, p2, w1, w2, cur)
            g = O.OracleGame().set_position(H, V, p1, p2, w1, w2, cur)
        seed, rid = rng.getrandbits(64), it
        plies = C.c_int()
        v = qh.qh_rollout(s, seed, rid, 1000, C.byref(plies))
        want_v, want_plies = g.rollout(seed, rid, 1000)
        assert (v, plies.value) == (want_v, want_plies)
        pos = g.position()
        m = unpack_meta(s[2])
        assert (s[0], s[1], m["p1"], m["p2"]) == (pos["H"], pos["V"], pos["p1"], pos["p2"])


M64 = (1 << 64) - 1


def test_tile_table_is_the_transpose_of_the_masks(qh):
    """qz_tile_table (four tiles per multiply) == the eight mask bits of every tile; tiles 81..83 stay empty;
    qz_ctx_store / qz_ctx_load keep every mask."""
    qh.qh_tile_table.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p, C.c_char_p]
    qh.qh_ctx_roundtrip.argtypes = [C.c_uint64, C.c_uint64]
    rng = random.Random(23)
    for it in range(400):
        H, V = rand_walls(rng, rng.randrange(0, 30))
        tab, direct = C.create_string_buffer(84), C.create_string_buffer(81)
        qh.qh_tile_table(H, V, tab, direct)
        assert tab.raw[:81] == direct.raw and tab.raw[81:] == b"\0\0\0"
        if it % 8 == 0:
            assert qh.qh_ctx_roundtrip(H, V) == 1


def test_unchecked_wall_on_occupied_intersection_replaces_it(qh):
    """quoridor.py:246-257 assigns the cell, so with safe=False a wall placed on an occupied intersection REPLACES
    the wall that stood there (the oracle's intersection array does the same)."""
    for first, second in ((12 + 27, 76 + 27), (76 + 9, 12 + 9)):
        g = O.OracleGame()
        s = mk_state(qh, 0, 0, 4, 76, 10, 10, 1)
        for a in (first, second):
            g.step(a)
            qh.qh_apply(s, a)
        pos = g.position()
        assert (s[0], s[1]) == (pos["H"], pos["V"])
        assert bin(s[0] | s[1]).count("1") == 1
        assert legal_list(qh, s) == g.actions()
