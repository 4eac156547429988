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
