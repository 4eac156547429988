"""GPU: batched self-play recording (quoridor.py:573-610 semantics) and the TrainPipeline mirror (train.py)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_selfplay_records_reference_samples():
    """Samples are (state before the move, move probabilities over 140 actions, z) with z = +1 on the winner's
    plies and -1 on the loser's (quoridor.py:596-602); states re-encode to the reference's state()."""
    from alphazero_quoridor_b200.quoridor import BatchedQuoridor, unpack_meta
    from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
    from alphazero_quoridor_b200.tree import StubEvaluator
    sp = BatchedSelfPlay(96, StubEvaluator("S3"), c_puct=5, n_playout=24, leaves_per_game=2, temp=1.0, pure=False,
                         seed=3, max_plies=400, record=True, fix_terminal_sign=True)
    for _ in range(400):
        sp.step()
        if len(sp.sink) >= 8:
            break
    assert len(sp.sink) >= 8, "no game finished"
    for st, pr, z in sp.sink[:8]:
        T = st.shape[0]
        assert pr.shape == (T, 140) and z.shape == (T,)
        np.testing.assert_allclose(pr.sum(1).cpu().numpy(), 1.0, atol=1e-5)
        metas = [unpack_meta(int(m)) for m in st[:, 2].cpu().numpy()]
        assert [m["ply"] for m in metas] == list(range(T))                     # one sample per ply, in order
        assert metas[0]["p1"] == 4 and metas[0]["p2"] == 76 and metas[0]["cur"] == 1
        movers = np.array([m["cur"] for m in metas])
        zz = z.cpu().numpy()
        winner = movers[-1] if zz[-1] == 1.0 else 3 - movers[-1]
        assert set(np.unique(zz)) <= {-1.0, 1.0}
        assert np.array_equal(zz, np.where(movers == winner, 1.0, -1.0))
        # the recorded 24-byte states re-encode to exactly the reference's planes
        planes = BatchedQuoridor(T, states=st.clone()).encode(dtype=torch.float32).cpu().numpy()
        for t in (0, T // 2, T - 1):
            m = metas[t]
            H, V = int(st[t, 0].item()) & (2**64 - 1), int(st[t, 1].item()) & (2**64 - 1)
            g = O.OracleGame().set_position(H, V, m["p1"], m["p2"], m["w1"], m["w2"], m["cur"])
            assert np.array_equal(planes[t].astype(np.float64), g.state())
            legal = g.actions()
            assert pr[t].cpu().numpy()[[a for a in range(140) if a not in legal]].sum() == 0   # mass only on legal moves


def test_train_pipeline_collects_and_updates():
    from alphazero_quoridor_b200.train import TrainPipeline
    torch.manual_seed(0)
    tp = TrainPipeline(n_parallel_games=64, leaves_per_game=4, fix_terminal_sign=True, max_plies=120)
    assert (tp.learn_rate, tp.temp, tp.n_playout, tp.c_puct, tp.buffer_size, tp.batch_size, tp.epochs, tp.kl_targ,
            tp.check_freq, tp.game_batch_num, tp.pure_mcts_playout_num) == (2e-3, 1.0, 400, 5, 10000, 128, 5, 0.02, 50,
                                                                            1500, 1000)            # train.py:17-31
    tp.n_playout = 12
    n = tp.collect_selfplay_data(4)
    assert n >= 4 and len(tp.data_buffer) > tp.batch_size
    st, pr, z = tp.data_buffer[0]
    assert st.shape == (3,) and pr.shape == (140,) and z in (-1.0, 0.0, 1.0)
    loss, entropy = tp.policy_update()
    assert np.isfinite(loss) and np.isfinite(entropy) and 0.1 <= tp.lr_multiplier <= 10


def test_mirror_augmentation():
    """Left-right mirror of (state, probs): planes flip left-right, wall planes flip inside their 8x8 block, action
    probabilities follow the action permutation; mirroring twice is the identity."""
    from alphazero_quoridor_b200.quoridor import BatchedQuoridor
    from alphazero_quoridor_b200.synthetic import midgame_positions
    from alphazero_quoridor_b200.train import MIRROR_ACTION, mirror_samples
    st = midgame_positions(512, seed=2, min_plies=0, max_plies=60)
    probs = torch.rand(512, 140, device=st.device)
    m_st, m_pr = mirror_samples(st, probs)
    back_st, back_pr = mirror_samples(m_st, m_pr)
    assert torch.equal(back_st, st) and torch.equal(back_pr, probs)
    assert sorted(MIRROR_ACTION.tolist()) == list(range(140))
    a = BatchedQuoridor(512, states=st.clone()).encode()
    b = BatchedQuoridor(512, states=m_st.clone()).encode()
    assert torch.equal(b[:, 3:], a[:, 3:].flip(3))                              # pawn / wall-count / turn planes
    assert torch.equal(b[:, :3, :8, :8], a[:, :3, :8, :8].flip(3))              # wall planes: 8x8 block mirrored
    # away from row 0 (where the reference's corner aliasing breaks the symmetry) the mirrored position has exactly
    # the mirrored legal moves: clear the row-0 intersections and lift both pawns to rows >= 2
    H, V, meta = st[:, 0] & ~0xFF, st[:, 1] & ~0xFF, st[:, 2]
    p1 = 18 + (meta & 0xFF) % 54
    p2 = 18 + ((meta >> 8) & 0xFF) % 63
    p2 = torch.where(p2 == p1, 18 + (p2 - 18 + 1) % 63, p2)
    up = torch.stack([H, V, (meta & ~0xFFFF) | p1 | (p2 << 8)], 1).contiguous()
    m_up, _ = mirror_samples(up, probs)
    la = BatchedQuoridor(512, states=up.clone()).legal_lists()
    lb = BatchedQuoridor(512, states=m_up.clone()).legal_lists()
    perm = MIRROR_ACTION.tolist()
    row0 = set(range(12, 20)) | set(range(76, 84))          # candidate walls ON row 0 meet the aliasing themselves
    for i in range(512):
        assert sorted(perm[x] for x in la[i] if x not in row0) == sorted(x for x in lb[i] if x not in row0), i


def test_checkpoint_resume_and_arena(tmp_path):
    from alphazero_quoridor_b200.train import TrainPipeline
    torch.manual_seed(0)
    tp = TrainPipeline(n_parallel_games=64, leaves_per_game=4, fix_terminal_sign=True, max_plies=120)
    tp.n_playout = 12
    tp.collect_selfplay_data(2)
    tp.policy_update()
    path = str(tmp_path / "resume.pt")
    tp.save_checkpoint(path)
    tp2 = TrainPipeline(n_parallel_games=64, leaves_per_game=4, fix_terminal_sign=True, max_plies=120)
    tp2.load_checkpoint(path)
    for (k, a), (_, b) in zip(tp.policy_value_net.get_policy_param().items(), tp2.policy_value_net.get_policy_param().items()):
        assert torch.equal(a, b), k
    assert len(tp2.data_buffer) == len(tp.data_buffer) and tp2.lr_multiplier == tp.lr_multiplier
    assert torch.equal(tp2.data_buffer[3][0], tp.data_buffer[3][0]) and tp2.data_buffer[3][2] == tp.data_buffer[3][2]
    # arena: a random-init net against pure MCTS with a few rollouts -- just has to run and return a ratio
    tp2.pure_mcts_playout_num = 24
    ratio = tp2.policy_evaluate(n_games=16, n_playout=8, max_plies=80)
    assert 0.0 <= ratio <= 1.0
