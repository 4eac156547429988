"""Parity AT THE BENCHMARKED SETTINGS (bench.py): the searches the bench times use K > 1 leaves per game per wave
(virtual loss) and finish stuck rollouts at the end of the search, which the exact K = 1 tests do not cover.  Here those
very settings are compared with the K = 1 in-wave engine -- the engine tests/test_mcts_gpu.py pins to the oracle and to the
live reference's visit counts exactly (mcts.py:103-144) -- on 256 positions, with stated bounds on the total-variation
distance of the root visit distributions and on how often the most visited move is the same.

Bounds (measured values in DESIGN.md section 2; printed by every run):
  pure MCTS priors (uniform, stub S1), K = 64, 1000 playouts          mean TV <= 0.002, max TV <= 0.02, same move >= 0.99
  pure MCTS with rollouts, K = 64, 1000 playouts, defer_until_drain   mean TV <= 1.15 x the seed-to-seed TV of two K = 1
                                                                      searches + 0.01 (rollouts are random: two exact
                                                                      searches with different streams differ that much)
  AlphaZero settings, stubs S2 / S3, K = 4, n_playout 100 and 800     see AZ_BOUNDS (bench.py itself runs the AlphaZero
                                                                      sections at K = 1, which is exact)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_POS = 256
# stub -> n_playout -> (mean TV, max TV, min same-best-move rate) for K = 4 against K = 1
# measured on B200 (round 2): S3 0.077 / 0.902 at 100 playouts, 0.048 / 0.965 at 800; S2 0.074 / 0.797 and 0.075 / 0.867.
# (The maximum over the 256 positions is not bounded: under S2's uncorrelated random values single positions flip to a
# different line entirely, max TV 0.99.)  K = 1 -- bench.py's default for the AlphaZero sections -- is exact.
AZ_BOUNDS = {"S3": {100: (0.10, 0.85), 800: (0.07, 0.92)},
             "S2": {100: (0.10, 0.74), 800: (0.10, 0.82)}}


def _positions():
    from alphazero_quoridor_b200.synthetic import midgame_positions
    return midgame_positions(N_POS, seed=21, min_plies=0, max_plies=30)


def _dist(v):
    v = v.double()
    return v / v.sum(1, keepdim=True).clamp(min=1)


def _tv(a, b):
    return 0.5 * (a - b).abs().sum(1)


def _stub_search(st, stub, npl, K):
    from alphazero_quoridor_b200 import tree
    eng = tree.BatchedMCTS(N_POS, tree.StubEvaluator(stub), c_puct=5, n_playout=npl, leaves_per_game=K, reuse_tree=False,
                           allow_large_k=True)
    eng.reset(st)
    eng.search()
    v, _, rn = eng.root_stats(temp=1.0)
    assert (rn == npl).all() and eng.overflow_count() == 0
    return _dist(v)


def _rollout_search(st, seed, K, drain):
    from alphazero_quoridor_b200 import tree
    eng = tree.BatchedMCTS(N_POS, tree.RolloutEvaluator(seed=seed), c_puct=5, n_playout=1000, leaves_per_game=K,
                           reuse_tree=False, defer_until_drain=drain)
    eng.reset(st)
    eng.search()
    v, _, rn = eng.root_stats(temp=1.0)
    assert (rn == 1000).all() and eng.overflow_count() == 0
    return _dist(v)


def test_pure_bench_settings_uniform_priors():
    """bench.py's K = 64 / 1000 playouts under pure MCTS's own priors (uniform, value 0: deterministic)."""
    st = _positions()
    a, b = _stub_search(st, "S1", 1000, 1), _stub_search(st, "S1", 1000, 64)
    tv = _tv(a, b)
    same = (a.argmax(1) == b.argmax(1)).double().mean().item()
    print("S1 K=64 vs K=1 @1000: TV mean %.5f max %.5f same-best %.3f" % (tv.mean().item(), tv.max().item(), same))
    assert tv.mean().item() <= 0.002 and tv.max().item() <= 0.02 and same >= 0.99


def test_pure_bench_settings_with_rollouts():
    """bench.py's exact search (K = 64, 1000 rollouts, stuck rollouts finished at the end of the search) against the
    exact in-wave K = 1 search with the same seed, measured against the noise floor of the rollouts themselves."""
    st = _positions()
    exact_a = _rollout_search(st, seed=11, K=1, drain=False)
    exact_b = _rollout_search(st, seed=12, K=1, drain=False)
    bench = _rollout_search(st, seed=11, K=64, drain=True)
    floor, tv = _tv(exact_a, exact_b), _tv(exact_a, bench)
    same_floor = (exact_a.argmax(1) == exact_b.argmax(1)).double().mean().item()
    same = (exact_a.argmax(1) == bench.argmax(1)).double().mean().item()
    print("rollouts @1000: TV(K=64+drain vs K=1) mean %.4f max %.4f | noise floor TV(K=1 seed a vs b) mean %.4f max %.4f | "
          "same best move %.3f (floor %.3f)" % (tv.mean().item(), tv.max().item(), floor.mean().item(), floor.max().item(),
                                                same, same_floor))
    assert tv.mean().item() <= 1.15 * floor.mean().item() + 0.01
    assert same >= same_floor - 0.08


@pytest.mark.parametrize("stub", ["S3", "S2"])
@pytest.mark.parametrize("npl", [100, 800])
def test_az_bench_settings(stub, npl):
    """The batched AlphaZero setting (K = 4 leaves per game per wave; n_playout 100 = BASELINE configs[2], 800 =
    configs[3]) under the deterministic non-uniform-prior stubs."""
    st = _positions()
    a, b = _stub_search(st, stub, npl, 1), _stub_search(st, stub, npl, 4)
    tv = _tv(a, b)
    same = (a.argmax(1) == b.argmax(1)).double().mean().item()
    print("%s K=4 vs K=1 @%d: TV mean %.4f p90 %.4f max %.4f same-best %.3f"
          % (stub, npl, tv.mean().item(), tv.quantile(0.9).item(), tv.max().item(), same))
    bm, bs = AZ_BOUNDS[stub][npl]
    assert tv.mean().item() <= bm and same >= bs


def test_k_is_capped_for_non_uniform_priors():
    """S2 at K = 64 drifts to TV 0.19 (round-1 measurement), outside any useful tolerance: the engine refuses K above
    MAX_K_NONUNIFORM for evaluators with non-uniform priors unless explicitly overridden."""
    from alphazero_quoridor_b200 import tree
    with pytest.raises(ValueError):
        tree.BatchedMCTS(4, tree.StubEvaluator("S2"), n_playout=16, leaves_per_game=64)
    tree.BatchedMCTS(4, tree.RolloutEvaluator(seed=1), n_playout=16, leaves_per_game=64)          # uniform priors: fine
    tree.BatchedMCTS(4, tree.StubEvaluator("S2"), n_playout=16, leaves_per_game=tree.MAX_K_NONUNIFORM)


def test_k_sweep_table():
    """The table DESIGN.md section 2 quotes (formerly tools/tv_vs_k.py): TV against K = 1 at 1000 playouts."""
    st = _positions()
    for stub in ("S1", "S2", "S3"):
        ref = _stub_search(st, stub, 1000, 1)
        for K in (4, 8, 32, 64):
            p = _stub_search(st, stub, 1000, K)
            tv = _tv(ref, p)
            print("TVTABLE %s K=%d: mean %.5f max %.5f same-best %.3f"
                  % (stub, K, tv.mean().item(), tv.max().item(), (p.argmax(1) == ref.argmax(1)).double().mean().item()))
