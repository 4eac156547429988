"""GPU parity of the env kernels (through the C ABI) against the golden fixtures and the CPU oracle.
Bit-exact: legal-action lists (ordered), next states, state tensors, outcomes (quoridor.py:58-202)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def qz():
    from alphazero_quoridor_b200 import quoridor
    return quoridor


def _meta5(hs):
    return np.array([[d["p1"], d["p2"], d["w1"], d["w2"], d["cur"]] for d in hs], dtype=np.int32)


def _u64(a):
    return np.asarray(a).astype(np.int64).view(np.uint64)


def test_reset_matches_reference_start(qz, kat):
    env = qz.BatchedQuoridor(5)
    hs = env.host_states()
    assert all((d["p1"], d["p2"], d["w1"], d["w2"], d["cur"], d["H"], d["V"], d["flags"], d["ply"]) ==
               (4, 76, 10, 10, 1, 0, 0, 0, 0) for d in hs)
    start = [r for r in kat["named"] if r["name"] == "start"][0]
    assert env.legal_lists()[0] == start["actions"]


def test_golden_traces_batched(qz, traces):
    """All golden games replayed side by side: ordered legal lists, planes and transitions every ply."""
    n = len(traces)
    env = qz.BatchedQuoridor(n)
    maxlen = max(len(t["plies"]) for t in traces)
    checked = 0
    for ply in range(maxlen):
        lists = env.legal_lists()
        planes = env.encode(dtype=torch.float32).cpu().numpy()
        hs = env.host_states()
        acts = np.full(n, -1, dtype=np.int32)
        for i, tr in enumerate(traces):
            if ply >= len(tr["plies"]):
                continue
            rec = tr["plies"][ply]
            d = hs[i]
            assert (d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"]) == (
                rec["H"], rec["V"], rec["p1"], rec["p2"], rec["w1"], rec["w2"], rec["cur"])
            assert lists[i] == rec["actions"], (tr["policy"], tr["seed"], ply)
            if rec["state"] is not None:
                assert hashlib.sha256(planes[i].astype(np.uint8).tobytes()).hexdigest() == rec["state"]
            if rec["action"] is not None:
                acts[i] = rec["action"]
            checked += 1
        env.step(torch.from_numpy(acts))
    hs = env.host_states()
    for i, tr in enumerate(traces):
        fin = tr["final"]
        assert (hs[i]["p1"], hs[i]["p2"], hs[i]["cur"], hs[i]["done"], hs[i]["winner"]) == (
            fin["p1"], fin["p2"], fin["cur"], fin["done"], fin["winner"])
    assert checked > 20000


def test_kat_positions(qz, kat):
    recs = [r for r in kat["named"] + kat["synthetic"] if r["actions"] is not None]
    rows = [qz.pack_state(r["H"], r["V"], r["p1"], r["p2"], r["w1"], r["w2"], r["cur"]) for r in recs]
    env = qz.BatchedQuoridor(len(rows), states=torch.tensor(rows, dtype=torch.int64))
    for rec, got in zip(recs, env.legal_lists()):
        assert got == rec["actions"], rec["name"]
    planes = env.encode().cpu().numpy()
    for rec, p in zip(recs, planes):
        if rec["state"] is not None:
            assert hashlib.sha256(p.astype(np.uint8).tobytes()).hexdigest() == rec["state"], rec["name"]


def test_masks_vs_oracle_synthetic(qz):
    """32768 reachable midgame positions (10-20 random plies): 140-bit masks identical to the oracle's."""
    from alphazero_quoridor_b200.synthetic import midgame_positions
    states = midgame_positions(32768, seed=7)
    env = qz.BatchedQuoridor(states.shape[0], states=states)
    got = _u64(env.legal_mask().cpu().numpy())
    hs = env.host_states()
    H = np.array([d["H"] for d in hs], dtype=np.uint64)
    V = np.array([d["V"] for d in hs], dtype=np.uint64)
    _, want = O.sweeps(H, V, _meta5(hs))
    assert np.array_equal(got, want)
    nwalls = np.array([bin(d["H"] | d["V"]).count("1") for d in hs])
    assert nwalls.min() >= 5 and nwalls.max() >= 18


def test_rollouts_vs_oracle(qz):
    """Kernel rollouts == oracle rollouts (same Philox streams): value, plies and final position."""
    from alphazero_quoridor_b200.rollout import rollout
    from alphazero_quoridor_b200.synthetic import midgame_positions
    seed = 0xC0FFEE
    start = qz.BatchedQuoridor(1).states
    mids = midgame_positions(512, seed=11, min_plies=4, max_plies=24)
    cases = [(start, 1024, 0), (mids, 2, 5000)]
    for states, per, base in cases:
        res, plies, final = rollout(states, per_state=per, seed=seed, rid_base=base, limit=1000, return_final=True)
        res, plies = res.cpu().numpy(), plies.cpu().numpy()
        fin = qz.BatchedQuoridor(final.shape[0], states=final).host_states()
        src = qz.BatchedQuoridor(states.shape[0], states=states).host_states()
        for r in range(0, res.shape[0], 3):       # every third rollout keeps the CPU side short
            d = src[r // per]
            g = O.OracleGame().set_position(d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"])
            v, k = g.rollout(seed, base + r, 1000)
            pos = g.position()
            f = fin[r]
            assert (int(res[r]), int(plies[r])) == (v, k), r
            assert (f["H"], f["V"], f["p1"], f["p2"], f["w1"], f["w2"], f["cur"]) == (
                pos["H"], pos["V"], pos["p1"], pos["p2"], pos["w1"], pos["w2"], pos["cur"])
    assert set(np.unique(res)) <= {-1, 0, 1}


@pytest.mark.parametrize("limit", [1, 2, 40, 129, 130, 258, 400, 5000])
def test_rollout_limits_vs_oracle(qz, limit):
    """pure_mcts.py:86-108 with other `limit`s: the pawn phase runs in 128-ply slices (one kernel pass each), so limits
    around the slice boundaries, below one slice, and large enough to stretch the slices are checked ply for ply."""
    from alphazero_quoridor_b200.rollout import rollout
    from alphazero_quoridor_b200.synthetic import midgame_positions
    from alphazero_quoridor_b200 import _lib
    assert 1 <= _lib.load().qz_rollout_pawn_passes(limit) <= 28
    late = midgame_positions(96, seed=21, min_plies=18, max_plies=44)         # some start inside the pawn phase
    states = torch.cat([qz.BatchedQuoridor(32).states, late], 0).contiguous()
    seed, base = 77 + limit, 31000
    res, plies, final = rollout(states, per_state=2, seed=seed, rid_base=base, limit=limit, return_final=True)
    res, plies = res.cpu().numpy(), plies.cpu().numpy()
    fin = qz.BatchedQuoridor(final.shape[0], states=final).host_states()
    src = qz.BatchedQuoridor(states.shape[0], states=states).host_states()
    assert plies.max() <= max(limit - 1, 0)
    for r in range(res.shape[0]):
        d = src[r // 2]
        g = O.OracleGame().set_position(d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"])
        v, k = g.rollout(seed, base + r, limit)
        pos, f = g.position(), fin[r]
        assert (int(res[r]), int(plies[r])) == (v, k), (limit, r)
        assert (f["H"], f["V"], f["p1"], f["p2"], f["w1"], f["w2"], f["cur"]) == (
            pos["H"], pos["V"], pos["p1"], pos["p2"], pos["w1"], pos["w2"], pos["cur"]), (limit, r)


def test_rollout_is_launch_shape_invariant(qz):
    """Outcome depends only on (state, seed, rid): explicit rids / state_index give the same answers."""
    from alphazero_quoridor_b200.rollout import rollout
    start = qz.BatchedQuoridor(3).states
    a, pa, _ = rollout(start, per_state=500, seed=3, rid_base=100)
    perm = torch.randperm(1500, generator=torch.Generator().manual_seed(0))
    rids = (perm + 100).to(torch.int64)
    sidx = torch.zeros(1500, dtype=torch.int32)
    b, pb, _ = rollout(start, seed=3, rids=rids, state_index=sidx)
    assert torch.equal(a.cpu()[perm], b.cpu()) and torch.equal(pa.cpu()[perm], pb.cpu())


def test_encode_dtypes_and_layouts(qz):
    from alphazero_quoridor_b200.synthetic import midgame_positions
    states = midgame_positions(1003, seed=5)[:1003]          # not a multiple of the 8 games per block
    n = states.shape[0]
    env = qz.BatchedQuoridor(n, states=states)
    ref = env.encode(dtype=torch.float32)
    assert ref.shape == (n, 26, 9, 9)
    assert torch.equal(ref.sum(dim=(2, 3))[:, 3:5], torch.ones(n, 2, device=ref.device))
    # against the rules header's plane definition evaluated on the host oracle for a few games
    for i in (0, 1, 7, 8, n - 1):
        d = env.host_states()[i]
        g = O.OracleGame().set_position(d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"])
        assert np.array_equal(ref[i].cpu().numpy().astype(np.float64), g.state())
    for m in (1, 3, 9):                                      # tiny batches: partial blocks only
        sub = qz.BatchedQuoridor(m, states=states[:m].clone())
        assert torch.equal(sub.encode(dtype=torch.bfloat16).float(), ref[:m])
        assert torch.equal(sub.encode(dtype=torch.float32), ref[:m])
    for dt in (torch.bfloat16, torch.float16):
        assert torch.equal(env.encode(dtype=dt).float(), ref)
    for dt in (torch.float32, torch.bfloat16):
        for cs in (26, 32):
            cl = env.encode(dtype=dt, channels_last=True, c_stride=cs)
            assert cl.shape == (n, cs, 9, 9) and cl.is_contiguous(memory_format=torch.channels_last)
            assert torch.equal(cl[:, :26].float(), ref)
            assert cl[:, 26:].abs().sum().item() == 0


def test_step_safe_mode_and_finished_games(qz):
    env = qz.BatchedQuoridor(4)
    mask = env.legal_mask()
    acts = torch.tensor([0, 1, 12, -1], dtype=torch.int32)        # S (=1) is illegal at the start
    done = env.step(acts, legal_mask=mask)
    hs = env.host_states()
    assert [d["p1"] for d in hs] == [13, 4, 4, 4]
    assert [bool(d["flags"] & qz.FLAG_ILLEGAL) for d in hs] == [False, True, False, False]
    assert [d["cur"] for d in hs] == [2, 1, 2, 1] and done.cpu().tolist() == [0, 0, 0, 0]
    # winning move: off-board jump (SURVEY.md 4 "off-board win P1"); mover is not rotated; later steps are no-ops
    rows = [qz.pack_state(0, 0, 67, 76, 0, 10, 1)]
    env = qz.BatchedQuoridor(1, states=torch.tensor(rows, dtype=torch.int64))
    assert env.step(torch.tensor([4], dtype=torch.int32)).cpu().tolist() == [1]
    d = env.host_states()[0]
    assert (d["p1"], d["cur"], d["done"], d["winner"]) == (85, 1, True, 1)
    env.step(torch.tensor([1], dtype=torch.int32))
    assert env.host_states()[0]["p1"] == 85
    assert env.legal_lists()[0] == []


def test_facade_matches_reference_api(qz, kat):
    """The drop-in `Quoridor` class: same calls and return types as quoridor.py."""
    g = qz.Quoridor()
    assert g.action_space == 140 and g.players == [1, 2] and g.get_current_player() == 1
    acts = g.actions()
    assert len(acts) == 131 and acts[:9] == [0, 2, 3, 12, 76, 13, 77, 14, 78]
    s = g.state()
    assert s.shape == (26, 9, 9) and s.dtype == np.float64
    assert s.sum(axis=(1, 2)).tolist() == [64, 0, 0, 1, 1] + [0] * 9 + [81] + [0] * 9 + [81] + [0]
    assert g.step(0) is False and g.current_player == 2 and g.last_player == 1
    assert g._positions == {1: 13, 2: 76} and g.valid_actions == acts
    # attributes stay authoritative on the host, as callers of the reference assign to them directly
    for rec in kat["named"]:
        if rec["actions"] is None:
            continue
        h = qz.Quoridor()
        h._positions = {1: rec["p1"], 2: rec["p2"]}
        for ix in range(64):
            h._intersections[ix] = 1 if rec["H"] >> ix & 1 else (-1 if rec["V"] >> ix & 1 else 0)
        h._player1_walls_remaining, h._player2_walls_remaining = rec["w1"], rec["w2"]
        h.current_player = rec["cur"]
        assert h.actions() == rec["actions"], rec["name"]
    safe = qz.Quoridor(safe=True)
    with pytest.raises(ValueError):
        safe.step(1)
    # off-board winner: state() raises like the reference (quoridor.py:69 IndexError)
    w = qz.Quoridor()
    w._positions = {1: 67, 2: 76}
    assert w.step(4) is True and w.has_a_winner() == (True, 1) and w.current_player == 1
    with pytest.raises(IndexError):
        w.state()


def test_full_size_properties(qz):
    """BASELINE config 5 scale (2^20 positions): properties that need no oracle, plus a sampled oracle check."""
    from alphazero_quoridor_b200.synthetic import midgame_positions
    n = 1 << 20
    states = midgame_positions(n, seed=7)
    env = qz.BatchedQuoridor(n, states=states)
    m1 = env.legal_mask()
    m2 = env.legal_mask()
    assert torch.equal(m1, m2)                                   # deterministic
    # legal walls are a subset of the precheck candidates: never on an occupied intersection
    H, V = states[:, 0], states[:, 1]
    hl = ((m1[:, 0] >> 12) & ((1 << 52) - 1)) | (m1[:, 1] << 52)
    vl = ((m1[:, 1] >> 12) & ((1 << 52) - 1)) | (m1[:, 2] << 52)
    occ = H | V
    assert not ((hl & occ) != 0).any() and not ((vl & occ) != 0).any()
    # monotonicity: after the mover places a legal wall, the opponent's legal walls are a subset of before
    first_h = hl & (-hl)                                          # lowest legal H wall (0 if none)
    has = first_h != 0
    ix = torch.log2(first_h[has].double().abs()).round().long()
    ix = torch.where(first_h[has] < 0, torch.full_like(ix, 63), ix)
    acts = torch.full((n,), -1, dtype=torch.int32, device=states.device)
    acts[has] = (12 + ix).int()
    env.step(acts, legal_mask=m1)
    assert not ((env.states[:, 2] >> 45) & 1).any()              # nothing flagged illegal
    m3 = env.legal_mask()
    hl3 = ((m3[:, 0] >> 12) & ((1 << 52) - 1)) | (m3[:, 1] << 52)
    vl3 = ((m3[:, 1] >> 12) & ((1 << 52) - 1)) | (m3[:, 2] << 52)
    opp_has_walls = (m3[:, 0] >> 12 != 0) | (m3[:, 1] != 0) | (m3[:, 2] != 0)
    sel = has & opp_has_walls
    assert not ((hl3[sel] & ~hl[sel]) != 0).any() and not ((vl3[sel] & ~vl[sel]) != 0).any()
    # sampled oracle check at this size
    idx = torch.arange(0, n, 257, device=states.device)
    sub = qz.BatchedQuoridor(idx.numel(), states=env.states[idx].clone())
    hs = sub.host_states()
    _, want = O.sweeps(np.array([d["H"] for d in hs], dtype=np.uint64), np.array([d["V"] for d in hs], dtype=np.uint64),
                       _meta5(hs))
    assert np.array_equal(_u64(m3[idx].cpu().numpy()), want)


def test_stuck_rollouts_vs_oracle(qz):
    """Late positions where a player still owns walls but (nearly) none can be placed legally: these rollouts
    leave the per-lane kernel and run in the warp-per-rollout kernel (qz_rollout_stuck_kernel: 32 draw attempts
    decoded at once, only the wall attempts before the first pawn attempt path-checked).  Same Philox streams => value, plies and final position equal the oracle's."""
    from alphazero_quoridor_b200.rollout import rollout
    from alphazero_quoridor_b200.synthetic import midgame_positions
    pos = midgame_positions(60000, seed=123, min_plies=28, max_plies=70)
    meta = pos[:, 2]
    walls = ((meta >> 16) & 0xFF) + ((meta >> 24) & 0xFF)
    live = ((meta >> 40) & 1) == 0
    sel = (walls > 0) & live
    stuck = pos[sel][:160].contiguous()
    assert stuck.shape[0] >= 60, "need late positions with walls in hand"
    seed = 4242
    res, plies, final = rollout(stuck, per_state=1, seed=seed, rid_base=9000, limit=1000, return_final=True)
    res, plies = res.cpu().numpy(), plies.cpu().numpy()
    fin = qz.BatchedQuoridor(final.shape[0], states=final).host_states()
    src = qz.BatchedQuoridor(stuck.shape[0], states=stuck).host_states()
    n_long = 0
    for r, d in enumerate(src):
        g = O.OracleGame().set_position(d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"])
        v, k = g.rollout(seed, 9000 + r, 1000)
        pos_o = g.position()
        f = fin[r]
        assert (int(res[r]), int(plies[r])) == (v, k), r
        assert (f["H"], f["V"], f["p1"], f["p2"], f["w1"], f["w2"], f["cur"]) == (
            pos_o["H"], pos_o["V"], pos_o["p1"], pos_o["p2"], pos_o["w1"], pos_o["w2"], pos_o["cur"])
        n_long += (f["w1"] + f["w2"]) > 0          # still holding walls at the end: stuck all the way
    assert n_long >= 10


def test_full_mask_random_games_vs_oracle(qz):
    """BASELINE config 1, literally: every ply = full legal mask + uniform pick + step.  384 whole games played by
    the three kernels == the oracle's games (literal actions()/step with the same Philox picks): length, winner
    and final position."""
    n, seed = 384, 31337
    env = qz.BatchedQuoridor(n)
    gid = torch.arange(1000, 1000 + n, dtype=torch.int64, device=env.device)
    plies = env.random_play(seed=seed, max_plies=3000, game_id=gid).cpu().numpy()
    hs = env.host_states()
    for i in range(0, n, 2):
        g = O.OracleGame()
        k = g.random_game(seed, 1000 + i, 3000)
        pos = g.position()
        d = hs[i]
        assert k == plies[i], i
        assert (d["H"], d["V"], d["p1"], d["p2"], d["w1"], d["w2"], d["cur"]) == (
            pos["H"], pos["V"], pos["p1"], pos["p2"], pos["w1"], pos["w2"], pos["cur"])
        assert d["done"] == g.has_a_winner()[0] and d["winner"] == (g.has_a_winner()[1] or 0)
    assert 200 < plies.mean() < 500 and np.mean([d["done"] for d in hs]) > 0.97      # a few games hit the ply cap
    # the three-launch-per-ply loop plays the very same games
    env2 = qz.BatchedQuoridor(n)
    plies2 = env2.random_play(seed=seed, max_plies=3000, game_id=gid, fused=False).cpu().numpy()
    assert np.array_equal(plies, plies2) and torch.equal(env.states[:, :2], env2.states[:, :2])
    assert torch.equal(env.states[:, 2] & 0xFFFFFFFFFF, env2.states[:, 2] & 0xFFFFFFFFFF)


def test_facade_replays_golden_traces(qz, traces):
    """The drop-in single-game class, call for call as the reference is used: actions() / state() / step() on
    whole recorded games (one per action policy)."""
    picked = [next(t for t in traces if t["policy"] == pol) for pol in ("uniform", "wallmix", "forward")]
    replayed = 0
    for tr in picked:
        g = qz.Quoridor()
        for rec in tr["plies"][:260]:
            replayed += 1
            assert g.actions() == rec["actions"]
            assert g.get_current_player() == rec["cur"]
            assert (g._positions[1], g._positions[2]) == (rec["p1"], rec["p2"])
            assert (g._player1_walls_remaining, g._player2_walls_remaining) == (rec["w1"], rec["w2"])
            if rec["state"] is not None and rec is tr["plies"][0] or rec["action"] is None:
                s = g.state()
                assert hashlib.sha256(s.astype(np.uint8).tobytes()).hexdigest() == rec["state"]
            if rec["action"] is None:
                break
            done = g.step(rec["action"])
            assert g.valid_actions == rec["actions"]                       # step() refreshes it (quoridor.py:165)
        if len(tr["plies"]) <= 260:
            fin = tr["final"]
            assert done == fin["done"] and g.has_a_winner() == (fin["done"], fin["winner"] or None)
    assert replayed > 300
