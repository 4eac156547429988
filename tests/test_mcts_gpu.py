"""GPU parity of the MCTS kernels (through the C ABI) against the golden fixtures (live reference) and the
CPU oracle.  With one leaf per game per wave the search is the reference's sequential algorithm and root
visit counts / Q must match EXACTLY under a deterministic evaluator (mcts.py:103-151).  With K > 1 leaves per
wave (virtual loss) the stated tolerance is a total-variation bound on the root visit distribution."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import n_mcts_golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

KIND = {"S1": 1, "S2": 2, "S3": 3}


@pytest.fixture(scope="module")
def qz():
    import alphazero_quoridor_b200.mcts as m
    import alphazero_quoridor_b200.pure_mcts as pm
    import alphazero_quoridor_b200.quoridor as q
    import alphazero_quoridor_b200.tree as t

    class NS:
        pass
    ns = NS()
    ns.mcts, ns.pure, ns.q, ns.tree = m, pm, q, t
    return ns


def _facade(qz, pos):
    g = qz.q.Quoridor()
    g._positions = {1: pos["p1"], 2: pos["p2"]}
    for ix in range(64):
        g._intersections[ix] = 1 if pos["H"] >> ix & 1 else (-1 if pos["V"] >> ix & 1 else 0)
    g._player1_walls_remaining, g._player2_walls_remaining = pos["w1"], pos["w2"]
    g.current_player = pos["cur"]
    return g


@pytest.mark.parametrize("idx", range(n_mcts_golden()))
def test_mcts_golden_reference_api(qz, mcts_golden, idx):
    """mcts.MCTS(policy, c_puct, n).get_move_probs / update_with_move vs the recorded reference runs."""
    case = mcts_golden[idx]
    tree = qz.mcts.MCTS(qz.mcts.DeviceStub(case["stub"]), case["c_puct"], case["n_playout"])
    g = _facade(qz, case["moves"][0]["pos"])
    for i, mv in enumerate(case["moves"]):
        acts, probs = tree.get_move_probs(g, case["temp"])
        a2, visits, qs, rn, rq = tree.root_children_stats()
        assert list(acts) == mv["acts"] == a2, case["name"]
        assert visits == mv["visits"], case["name"]
        assert qs == mv["q"], case["name"]                       # float64, bit-exact
        assert (rn, rq) == (mv["root_visits"], mv["root_q"])
        np.testing.assert_allclose(probs, np.array(mv["probs"]), rtol=1e-12, atol=1e-300)
        assert tree._engine.overflow_count() == 0               # an overflowing arena would silently cost exactness
        if i + 1 < len(case["moves"]):
            tree.update_with_move(mv["move"])
            g.step(mv["move"])


def _positions(n, seed, min_plies, max_plies):
    from alphazero_quoridor_b200.synthetic import midgame_positions
    return midgame_positions(n, seed=seed, min_plies=min_plies, max_plies=max_plies)


def _host(qz, states):
    hs = qz.q.BatchedQuoridor(states.shape[0], states=states).host_states()
    H = np.array([d["H"] for d in hs], dtype=np.uint64)
    V = np.array([d["V"] for d in hs], dtype=np.uint64)
    meta5 = np.array([[d["p1"], d["p2"], d["w1"], d["w2"], d["cur"]] for d in hs], dtype=np.int32)
    return hs, H, V, meta5


@pytest.mark.parametrize("kind,n_playout,plies", [("S2", 48, (6, 18)), ("S3", 64, (14, 30)), ("S1", 40, (22, 60)),
                                                   ("S3", 400, (40, 90))])
def test_batched_stub_mcts_vs_oracle(qz, kind, n_playout, plies):
    """Hundreds of independent trees searched side by side == the oracle's sequential search, exactly."""
    n = 192
    states = _positions(n, seed=21 + n_playout, min_plies=plies[0], max_plies=plies[1])
    eng = qz.tree.BatchedMCTS(n, qz.tree.StubEvaluator(kind), c_puct=5, n_playout=n_playout, leaves_per_game=1,
                              reuse_tree=False)
    eng.reset(states)
    eng.search()
    visits, probs, rootn = eng.root_stats(temp=1.0)
    _, H, V, meta5 = _host(qz, states)
    _, want = O.stub_mcts_visits(H, V, meta5, n_playout, c_puct=5.0, stub_kind=KIND[kind])
    assert np.array_equal(visits.cpu().numpy(), want)
    assert (rootn.cpu().numpy() == n_playout).all()
    assert eng.overflow_count() == 0
    p = probs.cpu().numpy()
    np.testing.assert_allclose(p.sum(1), 1.0, rtol=1e-12)


def test_tree_reuse_and_moves_vs_oracle(qz):
    """choose (first-max) + advance (re-root compaction) over several plies == oracle update_with_move."""
    n, n_playout, n_moves = 96, 80, 6
    states = _positions(n, seed=5, min_plies=24, max_plies=70)
    eng = qz.tree.BatchedMCTS(n, qz.tree.StubEvaluator("S3"), c_puct=5, n_playout=n_playout, leaves_per_game=1,
                              reuse_tree=True)
    eng.reset(states)
    hs, _, _, _ = _host(qz, states)
    oracles = [O.OracleMCTS(3, 5, n_playout) for _ in range(n)]
    games = [O.OracleGame().set_position(d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"]) for d in hs]
    alive = [True] * n
    for mv in range(n_moves):
        eng.search()
        visits, _, _ = eng.root_stats(temp=1.0)
        visits = visits.cpu().numpy()
        moves = eng.choose(mode=0).cpu().numpy()
        for i in range(n):
            if not alive[i]:
                continue
            acts, v, _ = oracles[i].run(games[i])
            want = np.zeros(140, dtype=np.int32)
            want[acts] = v
            assert np.array_equal(visits[i], want), (mv, i)
            best = acts[int(np.argmax(v))] if acts else -1
            assert moves[i] == best
            if best < 0:                       # stalemate: the reference would crash here
                alive[i] = False
                continue
            oracles[i].update_with_move(best)
            if games[i].step(best):
                alive[i] = False
        eng.advance(torch.from_numpy(moves))
        # finished games keep being searched on the device (terminal root); only live ones are compared
    assert sum(alive) > n // 2
    assert eng.overflow_count() == 0


def test_fix_terminal_sign_switch(qz):
    """fix_terminal_sign=True flips the reference's inverted terminal value; default keeps it."""
    row = qz.q.pack_state(0, 0, 64, 40, 0, 0, 1)          # P1 one step from winning, no walls
    st = torch.tensor([row], dtype=torch.int64)
    res = {}
    for fix in (False, True):
        eng = qz.tree.BatchedMCTS(1, qz.tree.StubEvaluator("S1"), c_puct=5, n_playout=200, fix_terminal_sign=fix,
                                  reuse_tree=False)
        eng.reset(st)
        eng.search()
        v, _, _, q = eng.root_stats(temp=1.0, want_q=True)
        res[fix] = (v[0, :4].cpu().tolist(), q[0, 0].item())
        o = O.OracleMCTS(1, 5, 200, fix_terminal_sign=fix)
        acts, ov, oq = o.run(O.OracleGame().set_position(0, 0, 64, 40, 0, 0, 1))
        assert res[fix][0] == ov and res[fix][1] == oq[0]
    assert res[False] == ([14, 86, 52, 47], -1.0)            # SURVEY.md 4 KAT "terminal sign"
    assert res[True][1] == 1.0 and res[True][0][0] == max(res[True][0]) > 100   # the winning move dominates once fixed


def test_virtual_loss_tolerance(qz):
    """K = 8 leaves per wave vs the exact K = 1 search: mean total-variation distance of root visit
    distributions <= 0.15 at 256 playouts (stated tolerance for the batched mode; measured 0.128)."""
    n, n_playout = 128, 256
    states = _positions(n, seed=77, min_plies=10, max_plies=40)
    out = {}
    for K in (1, 8):
        eng = qz.tree.BatchedMCTS(n, qz.tree.StubEvaluator("S3"), c_puct=5, n_playout=n_playout, leaves_per_game=K,
                                  reuse_tree=False)
        eng.reset(states)
        eng.search()
        v, _, rn = eng.root_stats(temp=1.0)
        assert (rn.cpu().numpy() == n_playout).all()
        out[K] = v.double().cpu().numpy()
        assert (out[K].sum(1) == n_playout - 1).all()            # first playout only expands the root
    tv = 0.5 * np.abs(out[1] / out[1].sum(1, keepdims=True) - out[8] / out[8].sum(1, keepdims=True)).sum(1)
    print("virtual-loss TV: mean %.4f max %.4f" % (tv.mean(), tv.max()))
    assert tv.mean() <= 0.15
    same_best = (out[1].argmax(1) == out[8].argmax(1)).mean()
    print("same best move: %.3f" % same_best)
    assert same_best >= 0.75


def test_pure_mcts_vs_oracle(qz):
    """Pure MCTS (uniform priors + rollouts, pure_mcts.py:66-115): same Philox rollout streams on both sides
    => identical visit counts and chosen moves at K = 1."""
    n, n_playout, seed = 48, 60, 99
    states = _positions(n, seed=31, min_plies=8, max_plies=50)
    ev = qz.tree.RolloutEvaluator(seed=seed)
    eng = qz.tree.BatchedMCTS(n, ev, c_puct=5, n_playout=n_playout, leaves_per_game=1, reuse_tree=False)
    eng.game_id.copy_(torch.arange(n, dtype=torch.int64) + 1000)
    eng.reset(states)
    eng.search()
    visits, _, _ = eng.root_stats(temp=1.0)
    visits = visits.cpu().numpy()
    moves = eng.choose(mode=0).cpu().numpy()
    hs, _, _, _ = _host(qz, states)
    for i, d in enumerate(hs):
        g = O.OracleGame().set_position(d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"])
        o = O.OracleMCTS(0, 5, n_playout, seed=seed, rollout_counter=(i + 1000) << 24)
        acts, v, _ = o.run(g)
        want = np.zeros(140, dtype=np.int32)
        want[acts] = v
        assert np.array_equal(visits[i], want), i
        assert moves[i] == acts[int(np.argmax(v))]


def test_python_callback_policy(qz, mcts_golden):
    """An arbitrary Python policy_value_fn (the reference's contract) drives the same kernels."""
    from stubs import make_stub
    case = [c for c in mcts_golden if c["name"] == "terminal_sign"][0]
    tree = qz.mcts.MCTS(make_stub("S1"), case["c_puct"], case["n_playout"])
    g = _facade(qz, case["moves"][0]["pos"])
    acts, probs = tree.get_move_probs(g, 1.0)
    _, visits, qs, _, _ = tree.root_children_stats()
    assert list(acts) == case["moves"][0]["acts"] and visits == case["moves"][0]["visits"]
    assert qs == case["moves"][0]["q"]
    case = [c for c in mcts_golden if c["name"] == "late_d_S2_cpuct1"][0]
    tree = qz.mcts.MCTS(make_stub("S2"), case["c_puct"], 150)
    dev = qz.mcts.MCTS(qz.mcts.DeviceStub("S2"), case["c_puct"], 150)
    g = _facade(qz, case["moves"][0]["pos"])
    tree.get_move_probs(g, 1.0)
    dev.get_move_probs(g, 1.0)
    assert tree.root_children_stats() == dev.root_children_stats()


def test_players_and_self_play_loop(qz):
    """MCTSPlayer.choose_action / get_action, pure MCTSPlayer, and Quoridor.start_self_play (quoridor.py:573-610)."""
    np.random.seed(0)
    g = qz.q.Quoridor()
    player = qz.mcts.MCTSPlayer(qz.mcts.DeviceStub("S3"), c_puct=5, n_playout=24, is_selfplay=1)
    move, probs = player.choose_action(g, temp=1.0, return_prob=1)
    assert move in g.actions() and probs.shape == (140,) and abs(probs.sum() - 1.0) < 1e-9
    assert player.get_action(g, temp=1.0) in g.actions()
    player.reset_player()
    # a short self-play game from a late position is too long to reach with random S3; run the real loop
    # with a pawn-only endgame by monkey-patching reset
    def late_reset(self=g):
        qz.q.Quoridor.reset(self)
        self._positions = {1: 58, 2: 22}
        self._player1_walls_remaining = self._player2_walls_remaining = 0
    g.reset = late_reset
    fast = qz.mcts.MCTSPlayer(qz.mcts.DeviceStub("S3"), c_puct=5, n_playout=30, is_selfplay=1, fix_terminal_sign=True)
    winner, data = g.start_self_play(fast, temp=1.0)
    data = list(data)
    assert winner in (1, 2) and len(data) >= 2
    s, p, z = data[0]
    assert s.shape == (26, 9, 9) and p.shape == (140,) and z in (-1.0, 1.0)
    zs = [d[2] for d in data]
    assert zs[-1] == 1.0                                         # the last mover won
    pure = qz.pure.MCTSPlayer(c_puct=5, n_playout=40)
    h = qz.q.Quoridor()
    mv = pure.choose_action(h)
    assert mv in h.actions()


def test_sharding_invariance(qz):
    """Results are keyed by the GLOBAL game index: 64 games searched as one batch == two shards of 32
    (what two ranks would own), moves and visit counts identical -- no collective needed."""
    from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
    from alphazero_quoridor_b200.shard import shard_range

    def run(lo, hi):
        sp = BatchedSelfPlay(hi - lo, qz.tree.RolloutEvaluator(seed=5), c_puct=5, n_playout=40, leaves_per_game=4,
                             pure=True, seed=5, game_id_base=lo)
        out = []
        for _ in range(3):
            mv = sp.step()
            v, _, _ = sp.mcts.root_stats(temp=1.0)
            out.append((mv.cpu(), sp.mcts.root_state.cpu().clone()))
        return out
    whole = run(0, 64)
    parts = [run(*shard_range(64, r, 2)) for r in range(2)]
    for t in range(3):
        assert torch.equal(whole[t][0], torch.cat([parts[0][t][0], parts[1][t][0]]))
        assert torch.equal(whole[t][1], torch.cat([parts[0][t][1], parts[1][t][1]]))


def test_unused_leaf_slots_are_idle(qz):
    """A wave that collects fewer than K leaves per game (the first wave of a search: one) must not evaluate the
    other slots: they hold a finished position, cost no rollout plies and leave the tree untouched."""
    n, K = 96, 8
    start = qz.q.BatchedQuoridor(n).states
    played = []
    for k_leaves in (1, K):
        ev = qz.tree.RolloutEvaluator(seed=5)
        eng = qz.tree.BatchedMCTS(n, ev, c_puct=5, n_playout=2 * K, leaves_per_game=k_leaves, reuse_tree=False)
        eng.reset(start)
        eng.playout_wave(1)
        torch.cuda.synchronize()
        played.append(ev.plies_played())
        if k_leaves == K:
            flags = eng.leaf_flags.view(n, K).cpu().numpy()
            meta = eng.leaf_state.view(n, K, 3)[:, :, 2].cpu().numpy()
            assert (flags[:, 0] & qz.tree.LEAF_INACTIVE == 0).all() and (flags[:, 1:] & qz.tree.LEAF_INACTIVE != 0).all()
            assert (((meta[:, 1:] >> 40) & 1) == 1).all()                 # idle slots: finished position
            assert int(eng.arena.visits.view(n, -1)[:, 0].sum().item()) == n      # one backup per game
    assert played[0] == played[1] and 0 < played[0] <= n * 999            # same rollouts as the K = 1 engine


def test_deferred_stuck_rollouts(qz):
    """defer_depth=3: the stuck rollouts of a wave finish on a side stream and are backed up three waves later;
    defer_until_drain: the stuck rollouts of every wave are finished together when the search ends.
    Nothing may be lost: every playout is counted once, no virtual loss is left behind, the run is deterministic
    and the chosen moves agree with the in-wave engine to the virtual-loss tolerance."""
    from alphazero_quoridor_b200.synthetic import midgame_positions
    n, n_playout, K = 256, 96, 8
    # late positions with walls in hand: many rollouts get ejected to the stuck kernel
    pos = midgame_positions(40000, seed=5, min_plies=26, max_plies=60)
    meta = pos[:, 2]
    sel = ((((meta >> 16) & 0xFF) + ((meta >> 24) & 0xFF)) > 0) & (((meta >> 40) & 1) == 0)
    states = torch.cat([pos[sel][:n // 2], midgame_positions(n, seed=6, min_plies=4, max_plies=30)], 0)[:n].contiguous()
    outs = []
    for defer, at_drain in ((0, False), (3, False), (3, False), (0, True), (0, True)):
        eng = qz.tree.BatchedMCTS(n, qz.tree.RolloutEvaluator(seed=11), c_puct=5, n_playout=n_playout,
                                  leaves_per_game=K, reuse_tree=False, defer_depth=defer, defer_until_drain=at_drain)
        eng.game_id.copy_(torch.arange(n, dtype=torch.int64) + 1000)
        eng.reset(states)
        eng.search()
        visits, _, rootn = eng.root_stats(temp=1.0)
        assert (rootn.cpu().numpy() == n_playout).all()
        tot = visits.sum(1).cpu().numpy()
        # n_playout - 1 child visits (the first playout only expands the root); 0 for a stalemated root
        assert np.isin(tot, (0, n_playout - 1)).all(), (defer, np.unique(tot))
        assert (tot == 0).mean() < 0.05
        for g in (0, 1, n // 2, n - 1):                          # no in-flight marks left anywhere in the tree
            todo = [int(eng.arena.root[g].item())]
            seen = 0
            while todo and seen < 400:
                kids, _ = eng.node_children(g, todo.pop())
                seen += 1
                assert all(k["inflight"] == 0 for k in kids)
                todo.extend(k["slot"] for k in kids if k["slot"] >= 0 and k["visits"] > 1)
        outs.append((visits.cpu().numpy().astype(np.float64), eng.choose(mode=0).cpu().numpy()))
    assert np.array_equal(outs[1][0], outs[2][0]) and np.array_equal(outs[1][1], outs[2][1])     # deterministic
    assert np.array_equal(outs[3][0], outs[4][0]) and np.array_equal(outs[3][1], outs[4][1])     # deterministic
    a = outs[0][0]
    ok = a.sum(1) > 0
    for name, b in (("3 waves late", outs[1][0]), ("at the end of the search", outs[3][0])):
        tv = 0.5 * np.abs(a[ok] / a[ok].sum(1, keepdims=True) - b[ok] / b[ok].sum(1, keepdims=True)).sum(1)
        print("deferred (%s) vs in-wave TV: mean %.4f" % (name, tv.mean()))
        assert tv.mean() <= 0.15


def test_move_sampling_distributions(qz):
    """qz_mcts_choose: mode 1 samples the visit-softmax (mcts.py:185), mode 2 samples
    0.75*probs + 0.25*Dirichlet(0.3) (mcts.py:181) whose mean is 0.75*probs + 0.25/n.  The reference draws from the
    global numpy RNG, so parity is distributional: 16384 identical trees keyed by different game ids."""
    n = 16384
    row = qz.q.pack_state(0, 0, 58, 22, 0, 0, 2)              # pawn-only position with a handful of moves
    states = torch.tensor([row], dtype=torch.int64).expand(n, 3).contiguous()
    eng = qz.tree.BatchedMCTS(n, qz.tree.StubEvaluator("S3"), c_puct=5, n_playout=64, leaves_per_game=1, reuse_tree=False)
    eng.reset(states)
    eng.search()
    visits, probs, _ = eng.root_stats(temp=1.0)
    p = probs[0].cpu().numpy()
    assert torch.equal(visits, visits[:1].expand_as(visits))             # identical trees
    legal = np.nonzero(p)[0]
    nc = len(legal)
    assert nc >= 3
    for mode, expect in ((1, p), (2, np.where(p > 0, 0.75 * p + 0.25 / nc, 0.0))):
        moves = eng.choose(mode=mode, temp=1.0, seed=17).cpu().numpy()
        assert np.isin(moves, legal).all()
        counts = np.bincount(moves, minlength=140).astype(np.float64)
        exp = expect * n
        chi2 = ((counts[legal] - exp[legal]) ** 2 / exp[legal]).sum()
        assert chi2 < 40.0, (mode, chi2, counts[legal], exp[legal])      # dof <= 11: far beyond 6 sigma
        again = eng.choose(mode=mode, temp=1.0, seed=17).cpu().numpy()
        assert np.array_equal(moves, again)                               # counter-based: reproducible
    # temp -> 0 concentrates on the most visited move (mcts.py:185 with temp=1e-3)
    greedy = eng.choose(mode=1, temp=1e-3, seed=3).cpu().numpy()
    assert (greedy == int(np.argmax(visits[0].cpu().numpy()))).mean() > 0.99


def test_reference_private_surface(qz, mcts_golden):
    """`MCTS._root` (TreeNode view: _children/_n_visits/_Q/_P/is_leaf/get_value), `_playout`, and
    pure `_evaluate_rollout` -- the members tests and tools of the reference reach into (mcts.py:12-127)."""
    case = [c for c in mcts_golden if c["name"] == "late_b_S2_800"][0]
    t = qz.mcts.MCTS(qz.mcts.DeviceStub("S2"), case["c_puct"], case["n_playout"])
    g = _facade(qz, case["moves"][0]["pos"])
    assert t._root.is_leaf() and t._root.is_root() and t._root._n_visits == 0
    t._playout(g)
    assert not t._root.is_leaf() and t._root._n_visits == 1
    t2 = qz.mcts.MCTS(qz.mcts.DeviceStub("S2"), case["c_puct"], case["n_playout"])
    t2.get_move_probs(g, 1.0)
    root = t2._root
    kids = root._children
    mv = case["moves"][0]
    assert list(kids.keys()) == mv["acts"]
    assert [k._n_visits for k in kids.values()] == mv["visits"]
    assert [k._Q for k in kids.values()] == mv["q"]
    assert root._n_visits == mv["root_visits"] and root._Q == mv["root_q"]
    best = max(kids.items(), key=lambda kv: kv[1].get_value(case["c_puct"]))      # TreeNode.select, mcts.py:42
    assert best[0] in mv["acts"] and all(0 < k._P <= 2 ** -6 for k in kids.values())
    # pure MCTS
    p = qz.pure.MCTS(c_puct=5, n_playout=20, seed=4)
    h = qz.q.Quoridor()
    vals = [p._evaluate_rollout(h) for _ in range(8)]
    assert set(vals) <= {-1, 0, 1} and h._positions == {1: 4, 2: 76}
    p._playout(h)
    assert abs(p._root._children[0]._P - 1.0 / 131) < 1e-15 and p._root._n_visits == 1


def test_full_size_configs_properties(qz):
    """BASELINE configs[1] and [2] at their full sizes, through properties that need no oracle plus a sampled
    oracle check: every playout is counted exactly once, probabilities are a distribution over legal moves, chosen
    moves are legal, nothing overflows, and the run is reproducible."""
    from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
    # configs[2]: 8192 games, n_playout = 100, c_puct = 5 (deterministic stub as the evaluator, one leaf per wave)
    n, npl = 8192, 100
    states = _positions(n, seed=99, min_plies=0, max_plies=50)
    eng = qz.tree.BatchedMCTS(n, qz.tree.StubEvaluator("S3"), c_puct=5, n_playout=npl, leaves_per_game=1, reuse_tree=False)
    eng.reset(states)
    eng.search()
    visits, probs, rootn = eng.root_stats(temp=1.0)
    assert (rootn == npl).all() and eng.overflow_count() == 0
    tot = visits.sum(1)
    assert ((tot == npl - 1) | (tot == 0)).all()
    env = qz.q.BatchedQuoridor(n, states=states)
    mask = env.legal_mask()
    bits = torch.stack([(mask[:, a >> 6] >> (a & 63)) & 1 for a in range(140)], 1).bool()
    assert not (visits[~bits] != 0).any()                                     # visits only on legal actions
    np.testing.assert_allclose(probs.sum(1)[tot > 0].cpu().numpy(), 1.0, rtol=1e-12)
    moves = eng.choose(mode=0)
    ok = torch.gather(bits, 1, moves.clamp(min=0).long().unsqueeze(1)).squeeze(1) | (moves < 0)
    assert ok.all()
    idx = torch.arange(0, n, 128)
    _, H, V, meta5 = _host(qz, states[idx.to(states.device)].clone())
    _, want = O.stub_mcts_visits(H, V, meta5, npl, c_puct=5.0, stub_kind=3)
    assert np.array_equal(visits[idx.to(visits.device)].cpu().numpy(), want)     # sampled exact check
    # configs[1]: 4096 games, 1000 rollouts per move, 64 leaves per wave, stuck rollouts deferred
    outs = []
    for _ in range(2):
        sp = BatchedSelfPlay(4096, qz.tree.RolloutEvaluator(seed=7), c_puct=5, n_playout=1000, leaves_per_game=64,
                             pure=True, seed=7, defer_depth=4)
        sp.mcts.search()
        v, _, rn = sp.mcts.root_stats(temp=1.0)
        assert (rn == 1000).all() and (v.sum(1) == 999).all() and sp.mcts.overflow_count() == 0
        mv = sp.mcts.choose(mode=0)
        outs.append((v.clone(), mv.clone()))
        assert (v[0].nonzero().flatten().cpu().tolist() == sorted(qz.q.Quoridor().actions()))     # all 131 root moves tried
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_lazy_expansion_equals_eager(qz):
    """Lazy expansion (a node's legal set computed when a playout comes back to it, qz_mcts_extend) against expansion at
    the first visit (the reference's order, pure_mcts.py:75-79): the same trees, visit for visit, at K = 1 and at K = 8
    (no deferral), with real rollouts on the same streams."""
    n = 96
    states = _positions(n, seed=41, min_plies=0, max_plies=40)
    for K, npl in ((1, 70), (8, 200)):
        out = []
        for lazy in (True, False):
            eng = qz.tree.BatchedMCTS(n, qz.tree.RolloutEvaluator(seed=5), c_puct=5, n_playout=npl, leaves_per_game=K,
                                      reuse_tree=False, lazy_expand=lazy)
            assert eng.lazy_expand == lazy
            eng.reset(states)
            eng.search()
            v, _, rn, q = eng.root_stats(temp=1.0, want_q=True)
            assert (rn == npl).all() and eng.overflow_count() == 0
            out.append((v.cpu(), q.cpu(), eng.choose(mode=0).cpu()))
        assert torch.equal(out[0][0], out[1][0]), K
        assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])


def test_node_view_lists_unvisited_children(qz):
    """`node_children` / the TreeNode view show every legal action in actions() order -- children that never got a slot
    included (0 visits, Q 0, their prior), as the reference's freshly expanded children (mcts.py:27-35)."""
    st = _positions(4, seed=3, min_plies=6, max_plies=12)
    eng = qz.tree.BatchedMCTS(4, qz.tree.StubEvaluator("S2"), c_puct=5, n_playout=30, leaves_per_game=1, reuse_tree=True)
    eng.reset(st)
    eng.search()
    legal = qz.q.BatchedQuoridor(4, states=st.clone()).legal_lists()
    visits, _, _ = eng.root_stats(temp=1.0)
    for g in range(4):
        kids, me = eng.node_children(g, int(eng.arena.root[g].item()))
        assert [k["action"] for k in kids] == legal[g] and me["visits"] == 30
        assert sum(k["visits"] for k in kids) == 29
        assert sum(1 for k in kids if k["slot"] >= 0) <= 29 < len(kids)          # most children never got a slot
        for k in kids:
            assert k["visits"] == int(visits[g, k["action"]].item())
            assert 0.0 < k["prior"] <= 2.0 ** -6 + 1e-12                          # S2 priors, present for unvisited ones too
            if k["slot"] < 0:
                assert k["visits"] == 0 and k["q"] == 0.0


def test_arena_overflow_is_counted_not_fatal(qz):
    """An arena that is too small costs exactness, never memory safety: playouts whose slot or block does not fit are
    evaluated where they stand and counted (qz_tree sizing note in include/qzb200.h)."""
    n = 32
    st = _positions(n, seed=9, min_plies=0, max_plies=20)
    for ev, kw in ((qz.tree.StubEvaluator("S3"), {}), (qz.tree.RolloutEvaluator(seed=2), dict(defer_until_drain=True))):
        eng = qz.tree.BatchedMCTS(n, ev, c_puct=5, n_playout=300, leaves_per_game=4, reuse_tree=False, node_cap=400, **kw)
        eng.reset(st)
        eng.search()
        v, _, rn = eng.root_stats(temp=1.0)
        assert eng.overflow_count() > 0
        assert (rn == 300).all() and (v.sum(1) <= 299).all() and (v >= 0).all()
        mv = eng.choose(mode=0)
        assert ((mv >= -1) & (mv < 140)).all()
        eng.check_device()
