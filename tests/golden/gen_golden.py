#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py [--jobs 8]

Imports quoridor.py / mcts.py from /root/reference (read-only) and records, per SURVEY.md 4/8c:
  * replay_traces.json.gz  -- seeded games under three action policies; per ply: ordered legal
                              list, positions, wall masks, walls left, mover, done/winner and a
                              sha256 of the (26,9,9) state tensor  (quoridor.py:58-186)
  * kat.json               -- the known-answer positions of SURVEY.md 4 + random synthetic
                              positions (full ordered legal list + state hash)
  * pawn_cases.json.gz     -- _valid_pawn_actions on random (walls, tile, opponent, player)
                              (quoridor.py:272-353)
  * mcts_golden.json       -- mcts.MCTS(stub, c_puct, n).get_move_probs visit vectors under the
                              deterministic stubs S1/S2, including tree-reuse sequences
                              (mcts.py:103-151)
Nothing here is imported by the product; tests read only the emitted files.
"""
import argparse
import contextlib
import copy
import gzip
import hashlib
import io
import json
import multiprocessing as mp
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from stubs import make_stub, masks_of  # noqa: E402


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def state_hash(game):
    """sha256 over the uint8 cast of state(); values are 0/1 only. None if the reference crashes."""
    try:
        s = game.state()
    except IndexError:
        return None
    assert s.shape == (26, 9, 9) and s.dtype == np.float64
    assert np.all((s == 0) | (s == 1))
    return hashlib.sha256(s.astype(np.uint8).tobytes()).hexdigest()


def snapshot(game, with_actions=True):
    H, V = masks_of(game)
    rec = {
        "p1": int(game._positions[1]), "p2": int(game._positions[2]),
        "H": H, "V": V,
        "w1": int(game._player1_walls_remaining), "w2": int(game._player2_walls_remaining),
        "cur": int(game.current_player),
    }
    over, winner = game.has_a_winner()
    rec["done"] = bool(over)
    rec["winner"] = 0 if winner is None else int(winner)
    if with_actions:
        if over:
            rec["actions"] = None       # reference undefined / unused on terminal states
            rec["state"] = None
        else:
            rec["actions"] = [int(a) for a in game.actions()]
            rec["state"] = state_hash(game)
    return rec


# ----------------------------------------------------------------------------- replay traces
def play_trace(args):
    policy, seed, cap = args
    from quoridor import Quoridor
    rng = random.Random(seed)
    with quiet():
        g = Quoridor()
        plies = []
        done = False
        while not done and len(plies) < cap:
            rec = snapshot(g)
            acts = rec["actions"]
            if not acts:                      # stalemate: reference would crash downstream
                rec["action"] = None
                plies.append(rec)
                break
            pawn = [a for a in acts if a < 12]
            wall = [a for a in acts if a >= 12]
            if policy == "uniform":
                a = acts[rng.randrange(len(acts))]
            elif policy == "wallmix":
                pool = wall if (wall and (not pawn or rng.random() < 0.5)) else pawn
                a = pool[rng.randrange(len(pool))]
            elif policy == "forward":
                # pawn-heavy and goal-directed so that pawns meet and jump
                if wall and rng.random() < 0.15:
                    a = wall[rng.randrange(len(wall))]
                else:
                    fwd = [0, 4, 8, 9] if rec["cur"] == 1 else [1, 5, 10, 11]
                    good = [x for x in pawn if x in fwd]
                    pool = good if (good and rng.random() < 0.75) else pawn
                    a = pool[rng.randrange(len(pool))]
            else:
                raise ValueError(policy)
            rec["action"] = int(a)
            plies.append(rec)
            done = g.step(a)
        final = snapshot(g, with_actions=False)
    return {"policy": policy, "seed": seed, "plies": plies, "final": final}


# ----------------------------------------------------------------------------- KATs
def build_position(p1, p2, H=(), V=(), w1=10, w2=10, cur=1):
    from quoridor import Quoridor
    g = Quoridor()
    g._positions = {1: p1, 2: p2}
    for ix in H:
        g._intersections[ix] = 1
    for ix in V:
        g._intersections[ix] = -1
    g._player1_walls_remaining = w1
    g._player2_walls_remaining = w2
    g.current_player = cur
    g.last_player = 3 - cur
    return g


def kat_cases():
    """SURVEY.md 4 table, as (name, kwargs)."""
    return [
        ("start", dict(p1=4, p2=76)),
        ("row0_bug_a", dict(p1=4, p2=76, H=[4], w1=0)),
        ("row0_bug_b", dict(p1=4, p2=76, H=[3], w1=0)),
        ("row0_bug_c", dict(p1=4, p2=76, V=[4], w1=0)),
        ("row0_bug_d", dict(p1=4, p2=76, V=[3], w1=0)),
        ("jump_open", dict(p1=31, p2=40, w1=0)),
        ("jump_through_wall", dict(p1=30, p2=31, H=[27], w1=0)),
        ("offboard_win_p1", dict(p1=67, p2=76, w1=0)),
        ("offboard_win_p2", dict(p1=4, p2=13, w2=0, cur=2)),
        ("walls0_plane", dict(p1=4, p2=76, w1=0)),
        ("wall_overlap_H", dict(p1=4, p2=76, H=[9])),
        ("wall_overlap_V", dict(p1=4, p2=76, V=[9])),
        ("plane_layout", dict(p1=4, p2=76, H=[9], V=[18], w1=9, w2=9)),
        ("stalemate", dict(p1=4, p2=13, V=[3, 4], H=[12], w1=0)),
        ("p2_start_view", dict(p1=13, p2=76, cur=2)),
    ]


def random_walled_position(rng, n_walls, adjacent=False):
    """Synthetic position: n_walls placed by the cheap prechecks only (quoridor.py:432-461)."""
    H, V = set(), set()
    tries = 0
    while len(H) + len(V) < n_walls and tries < 1000:
        tries += 1
        ix = rng.randrange(64)
        r, c = divmod(ix, 8)
        if ix in H or ix in V:
            continue
        if rng.random() < 0.5:
            if (c > 0 and ix - 1 in H) or (c < 7 and ix + 1 in H):
                continue
            H.add(ix)
        else:
            if (r > 0 and ix - 8 in V) or (r < 7 and ix + 8 in V):
                continue
            V.add(ix)
    while True:
        p1 = rng.randrange(0, 72)           # P1 not on its goal row
        if adjacent:
            p2 = p1 + rng.choice([9, -9, 1, -1])
        else:
            p2 = rng.randrange(9, 81)       # P2 not on its goal row
        if p2 != p1 and 9 <= p2 <= 80:
            break
    return p1, p2, sorted(H), sorted(V)


def kat_worker(job):
    kind, seed = job
    rng = random.Random(seed)
    with quiet():
        if kind == "synthetic":
            nw = rng.randrange(4, 21)
            p1, p2, H, V = random_walled_position(rng, nw, adjacent=rng.random() < 0.3)
            used = len(H) + len(V)
            u1 = rng.randrange(max(0, used - 10), min(10, used) + 1)
            w1, w2 = 10 - u1, 10 - (used - u1)
            cur = rng.choice([1, 2])
            if (w1 if cur == 1 else w2) == 0:   # mover keeps >=1 wall so the sweep runs
                if cur == 1:
                    w1 = 1
                else:
                    w2 = 1
            g = build_position(p1, p2, H, V, w1, w2, cur)
        else:
            raise ValueError(kind)
        rec = snapshot(g)
    rec["name"] = "%s_%d" % (kind, seed)
    return rec


def pawn_cases(n, seed):
    """Direct _valid_pawn_actions calls on random inputs, including non-adjacent/adjacent pawns."""
    from quoridor import Quoridor
    rng = random.Random(seed)
    g = Quoridor()
    out = []
    for i in range(n):
        nw = rng.randrange(0, 28)
        p1, p2, H, V = random_walled_position(rng, nw, adjacent=rng.random() < 0.7)
        walls = np.zeros(64)
        walls[H] = 1
        walls[V] = -1
        loc, opp = (p1, p2) if rng.random() < 0.5 else (p2, p1)
        # also exercise edge tiles anywhere on the board (BFS visits them)
        if rng.random() < 0.3:
            loc = rng.randrange(81)
            opp = loc + rng.choice([9, -9, 1, -1, 2, 18, -18, 10, 30])
            if not (0 <= opp <= 80) or opp == loc:
                opp = (loc + 40) % 81
        player = rng.choice([1, 2])
        res = g._valid_pawn_actions(walls, loc, opp, player)
        Hm = sum(1 << ix for ix in H)
        Vm = sum(1 << ix for ix in V)
        out.append([Hm, Vm, loc, opp, player, [int(a) for a in res]])
    return out


# ----------------------------------------------------------------------------- MCTS goldens
def mcts_worker(job):
    import mcts as ref_mcts
    name, pos, stub_kind, c_puct, n_playout, n_moves, temp = job
    with quiet():
        g = build_position(**pos)
        tree = ref_mcts.MCTS(make_stub(stub_kind), c_puct, n_playout)
        moves = []
        for m in range(n_moves):
            before = snapshot(g, with_actions=False)
            acts, probs = tree.get_move_probs(g, temp)
            root = tree._root
            visits = [int(root._children[a]._n_visits) for a in acts]
            qs = [float(root._children[a]._Q) for a in acts]
            # deterministic move: first max of visits (pure_mcts.py:115 rule)
            move = int(acts[int(np.argmax(visits))])
            moves.append({
                "pos": before, "acts": [int(a) for a in acts], "visits": visits, "q": qs,
                "probs": [float(p) for p in probs], "root_visits": int(root._n_visits),
                "root_q": float(root._Q), "move": move,
            })
            if m + 1 < n_moves:
                tree.update_with_move(move)
                done = g.step(move)
                if done:
                    break
    return {"name": name, "stub": stub_kind, "c_puct": c_puct, "n_playout": n_playout,
            "temp": temp, "moves": moves}


def mcts_jobs():
    start = dict(p1=4, p2=76)
    late_a = dict(p1=64, p2=40, w1=0, w2=0)                       # SURVEY KAT "terminal sign"
    late_b = dict(p1=58, p2=22, H=[20, 22, 45, 51], V=[9, 30, 41, 60], w1=0, w2=0, cur=2)
    late_c = dict(p1=40, p2=49, H=[36, 38, 12], V=[27, 44, 5], w1=0, w2=0)
    late_d = dict(p1=31, p2=30, H=[10, 42, 53], V=[19, 29, 50, 62], w1=0, w2=0, cur=2)
    mid_a = dict(p1=22, p2=58, H=[9, 11, 34, 52], V=[20, 27, 46], w1=6, w2=7)
    mid_b = dict(p1=40, p2=41, H=[27, 37], V=[33, 30, 3], w1=1, w2=4, cur=2)
    one_wall = dict(p1=49, p2=31, H=[41, 25], V=[36, 12], w1=1, w2=0)
    jobs = [
        ("terminal_sign", late_a, "S1", 5, 200, 1, 1.0),
        ("tie_break", start, "S1", 5, 12, 1, 1.0),
        ("start_S1", start, "S1", 5, 140, 1, 1.0),
        ("start_S2", start, "S2", 5, 140, 1, 1.0),
        ("mid_a_S2", mid_a, "S2", 5, 120, 1, 1.0),
        ("mid_b_S2", mid_b, "S2", 5, 150, 1, 1.0),
        ("mid_a_S1", mid_a, "S1", 5, 100, 1, 1.0),
        ("one_wall_S2", one_wall, "S2", 5, 400, 1, 1.0),
        ("late_a_S2_800", late_a, "S2", 5, 800, 1, 1.0),
        ("late_b_S1_800", late_b, "S1", 5, 800, 1, 1.0),
        ("late_b_S2_800", late_b, "S2", 5, 800, 1, 1.0),
        ("late_c_S2_800", late_c, "S2", 5, 800, 1, 1.0),
        ("late_d_S2_cpuct1", late_d, "S2", 1, 600, 1, 1.0),
        ("late_d_S1_cpuct20", late_d, "S1", 20, 600, 1, 1.0),
        # tree reuse (mcts.py:146-151) over several plies
        ("reuse_late_a_S2", late_a, "S2", 5, 300, 8, 1.0),
        ("reuse_late_c_S1", late_c, "S1", 5, 300, 8, 1.0),
        ("reuse_late_b_S2", late_b, "S2", 5, 200, 12, 1.0),
        ("reuse_mid_b_S2", mid_b, "S2", 5, 60, 4, 1.0),
        ("reuse_start_S2", start, "S2", 5, 40, 3, 1.0),
        ("temp_small", late_c, "S2", 5, 200, 1, 1e-3),
        ("late_a_S3_800", late_a, "S3", 5, 800, 1, 1.0),
        ("late_b_S3_800", late_b, "S3", 5, 800, 1, 1.0),
        ("late_c_S3_800", late_c, "S3", 5, 800, 1, 1.0),
        ("late_d_S3_800", late_d, "S3", 5, 800, 1, 1.0),
        ("mid_a_S3", mid_a, "S3", 5, 150, 1, 1.0),
        ("start_S3", start, "S3", 5, 150, 1, 1.0),
        ("one_wall_S3", one_wall, "S3", 5, 400, 1, 1.0),
        ("reuse_late_d_S3", late_d, "S3", 5, 400, 10, 1.0),
        ("reuse_mid_a_S3", mid_a, "S3", 5, 50, 4, 1.0),
    ]
    return jobs


def dump(name, obj, gz=False):
    path = os.path.join(HERE, name)
    data = json.dumps(obj, separators=(",", ":")).encode()
    if gz:
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)
    print("wrote %s (%d bytes raw)" % (path, len(data)), file=sys.stderr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--only", default="")
    ap.add_argument("--games", type=int, default=36, help="games per policy")
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    pool = mp.Pool(args.jobs)

    if not only or "mcts" in only:
        mcts_async = pool.map_async(mcts_worker, mcts_jobs(), chunksize=1)

    if not only or "pawn" in only:
        dump("pawn_cases.json.gz", pawn_cases(6000, 11), gz=True)

    if not only or "kat" in only:
        named = []
        with quiet():
            for name, kw in kat_cases():
                g = build_position(**kw)
                rec = snapshot(g)
                rec["name"] = name
                named.append(rec)
        synth = pool.map(kat_worker, [("synthetic", s) for s in range(240)], chunksize=4)
        dump("kat.json", {"named": named, "synthetic": synth})

    if not only or "traces" in only:
        jobs = []
        for i in range(args.games):
            jobs += [("uniform", 1000 + i, 700), ("wallmix", 2000 + i, 700), ("forward", 3000 + i, 700)]
        traces = pool.map(play_trace, jobs, chunksize=1)
        dump("replay_traces.json.gz", traces, gz=True)

    if not only or "mcts" in only:
        dump("mcts_golden.json", mcts_async.get())


if __name__ == "__main__":
    main()
