"""Pins the CPU oracle (oracle/quoridor_oracle.c) to the reference.

Every fixture under tests/golden/ was produced by running the UNMODIFIED Python reference
(tests/golden/gen_golden.py); the reference itself ships no tests.  CPU only.
"""
import hashlib

import numpy as np
import pytest

from conftest import n_mcts_golden
from oracle import oracle as O


def _state_hash(s):
    assert s.shape == (26, 9, 9) and s.dtype == np.float64
    return hashlib.sha256(s.astype(np.uint8).tobytes()).hexdigest()


def _check_snapshot(g, rec):
    pos = g.position()
    for k in ("p1", "p2", "H", "V", "w1", "w2", "cur"):
        assert pos[k] == rec[k], (k, pos, rec)
    over, winner = g.has_a_winner()
    assert over == rec["done"]
    assert (winner or 0) == rec["winner"]


def test_replay_traces(traces):
    """quoridor.py:138-186 -- ordered legal lists, next states and state tensors, ply by ply."""
    n_plies = 0
    for tr in traces:
        g = O.OracleGame()
        for rec in tr["plies"]:
            _check_snapshot(g, rec)
            assert g.actions() == rec["actions"], (tr["policy"], tr["seed"])
            if rec["state"] is not None:
                assert _state_hash(g.state()) == rec["state"]
            if rec["action"] is None:
                break
            g.step(rec["action"])
            n_plies += 1
        _check_snapshot(g, tr["final"])
    assert n_plies > 20000


def test_kat_positions(kat):
    """SURVEY.md 4 known answers + 240 synthetic walled positions (full ordered legal list)."""
    for rec in kat["named"] + kat["synthetic"]:
        g = O.OracleGame().set_position(rec["H"], rec["V"], rec["p1"], rec["p2"], rec["w1"], rec["w2"], rec["cur"])
        if rec["actions"] is None:
            continue
        assert g.actions() == rec["actions"], rec["name"]
        if rec["state"] is not None:
            assert _state_hash(g.state()) == rec["state"], rec["name"]


def test_survey_kat_values(kat):
    """Spot-check the literal numbers quoted in SURVEY.md 4 (guards the fixture itself)."""
    by = {r["name"]: r for r in kat["named"]}
    assert len(by["start"]["actions"]) == 131 and by["start"]["actions"][:9] == [0, 2, 3, 12, 76, 13, 77, 14, 78]
    assert by["row0_bug_a"]["actions"] == [0, 2, 3]
    assert by["row0_bug_b"]["actions"] == [2, 3]
    assert by["row0_bug_c"]["actions"] == [0, 2, 3]
    assert by["row0_bug_d"]["actions"] == [0]
    assert by["jump_open"]["actions"] == [1, 2, 3, 4, 8, 9]
    assert by["jump_through_wall"]["actions"] == [1, 3, 6, 8, 10]
    assert by["offboard_win_p1"]["actions"] == [1, 2, 3, 4, 8, 9]
    assert by["offboard_win_p2"]["actions"] == [0, 2, 3, 5, 10, 11]
    assert by["stalemate"]["actions"] == []
    start = set(by["start"]["actions"])
    assert start - set(by["wall_overlap_H"]["actions"]) == {20, 21, 22, 85}
    assert start - set(by["wall_overlap_V"]["actions"]) == {21, 77, 85, 93}


def test_pawn_cases(pawn_cases):
    """quoridor.py:272-353 on 6000 random (walls, tile, opponent, player)."""
    for H, V, loc, opp, player, want in pawn_cases:
        assert O.valid_pawn_actions(H, V, loc, opp, player) == want


def test_offboard_terminal():
    g = O.OracleGame().set_position(0, 0, 67, 76, 0, 10, 1)
    assert g.step(4) is True
    assert g.position()["p1"] == 85 and g.position()["cur"] == 1      # mover not rotated on a win
    assert g.has_a_winner() == (True, 1)
    with pytest.raises(IndexError):
        g.state()


@pytest.mark.parametrize("idx", range(n_mcts_golden()))
def test_mcts_golden(mcts_golden, idx):
    """mcts.py:103-151 under the deterministic stubs: visit vectors, Q and probabilities."""
    case = mcts_golden[idx]
    kind = {"S1": 1, "S2": 2, "S3": 3}[case["stub"]]
    tree = O.OracleMCTS(kind, case["c_puct"], case["n_playout"])
    first = case["moves"][0]["pos"]
    g = O.OracleGame().set_position(first["H"], first["V"], first["p1"], first["p2"], first["w1"], first["w2"],
                                    first["cur"])
    for i, mv in enumerate(case["moves"]):
        _check_snapshot(g, mv["pos"])
        acts, visits, qs = tree.run(g)
        assert acts == mv["acts"], case["name"]
        assert visits == mv["visits"], case["name"]
        assert qs == mv["q"], case["name"]              # float64, bit-exact
        n, q = tree.root_stats()
        assert n == mv["root_visits"] and q == mv["root_q"]
        probs = O.visits_to_probs(visits, case["temp"])
        np.testing.assert_allclose(probs, np.array(mv["probs"]), rtol=1e-12, atol=1e-300)
        if i + 1 < len(case["moves"]):
            tree.update_with_move(mv["move"])
            g.step(mv["move"])


def test_rollout_sampler_is_uniform_over_actions():
    """oq_sample_action picks only members of actions() and covers them roughly uniformly."""
    g = O.OracleGame().set_position(H=(1 << 9) | (1 << 34), V=(1 << 20) | (1 << 27) | (1 << 46), p1=22, p2=58,
                                    w1=6, w2=7, cur=1)
    legal = g.actions()
    counts = dict.fromkeys(legal, 0)
    n = 20000
    for rid in range(n):
        a = g.sample_action(1234, rid, 0)
        assert a in counts
        counts[a] += 1
    exp = n / len(legal)
    chi2 = sum((c - exp) ** 2 / exp for c in counts.values())
    assert chi2 < 2.0 * len(legal)          # dof = len-1 ~ 100; 2x is > 6 sigma


def test_rollout_terminates():
    g = O.OracleGame()
    v, plies = g.rollout(7, 0)
    assert v in (-1, 0, 1) and 0 < plies <= 999
