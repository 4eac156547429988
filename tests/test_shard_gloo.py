"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphazero_quoridor_b200 import shard


def test_shard_ranges_partition_the_games():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            for g in (0, n // 3, n - 1):
                if 0 <= g < n:
                    r = shard.owner_of(g, n, ws)
                    assert spans[r][0] <= g < spans[r][1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        assert shard.world() == (rank, ws, rank)
        lo, hi = shard.shard_range(n_total, rank, ws)
        # each rank "computes" a value that depends only on the GLOBAL game index (as the Philox streams do)
        local = (torch.arange(lo, hi, dtype=torch.int64) * 2654435761 % 1000003).unsqueeze(1)
        allrows = shard.gather_rows(local, n_total)
        want = (torch.arange(n_total, dtype=torch.int64) * 2654435761 % 1000003).unsqueeze(1)
        ok_rows = bool(torch.equal(allrows, want))
        ms, work = shard.reduce_stats(10.0 + rank, [float(hi - lo), 1.0])
        q.put((rank, ok_rows, ms, work))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_reduction_and_gather():
    ws, n_total = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_total, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(ws)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, ok_rows, ms, work in res:
        assert ok_rows
        assert ms == 11.0                      # max over ranks
        assert work == [float(n_total), 2.0]   # sum over ranks


def _train_worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        import numpy as np
        from alphazero_quoridor_b200.policy_value_net import PolicyValueNet
        from alphazero_quoridor_b200.train import allreduce_gradients
        torch.manual_seed(0)
        torch.set_num_threads(1)
        net = PolicyValueNet(use_gpu=False)
        rng = np.random.RandomState(5)
        s = rng.randint(0, 2, size=(16, 26, 9, 9)).astype(np.float64)
        p = rng.dirichlet(np.ones(140), size=16)
        z = rng.choice([-1.0, 1.0], size=16)
        lo, hi = rank * 8, rank * 8 + 8                       # each rank trains on its own half of the batch
        params = list(net.policy_value_net.parameters())
        net.train_step(s[lo:hi], p[lo:hi], z[lo:hi], 2e-3, grad_hook=lambda: allreduce_gradients(params))
        flat = torch.cat([x.detach().reshape(-1) for x in params])
        q.put((rank, flat.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_trainer_gradient_allreduce_two_ranks():
    """The trainer's only collective (gradient averaging): two ranks on different half-batches end up with
    identical weights (conv/linear weights; BatchNorm batch statistics are per rank, as with DDP)."""
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=150) for _ in range(ws))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    import numpy as np
    assert np.array_equal(res[0], res[1])
    assert np.isfinite(res[0]).all()


def _pipeline_worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from alphazero_quoridor_b200.train import TrainPipeline
        torch.set_num_threads(1)
        torch.manual_seed(100 + rank)                         # DIFFERENT initial weights per rank: the constructor must broadcast

        def fake_encode(rows):                                # CPU stand-in for the encode kernel (host logic under test)
            bits = (rows[:, :2, None] >> torch.arange(41)) & 1                     # [B,2,41] -> 82 bits
            x = bits.reshape(rows.shape[0], -1)[:, :81].float().reshape(-1, 1, 9, 9)
            return x.expand(-1, 26, -1, -1).contiguous()
        tp = TrainPipeline(use_gpu=False, encode_fn=fake_encode, seed=3)
        tp.batch_size, tp.epochs = 16, 3
        start = torch.cat([p.detach().reshape(-1) for p in tp.policy_value_net.policy_value_net.parameters()]).clone()
        g = torch.Generator().manual_seed(50 + rank)          # different data on every rank

        def feed(m):
            st = torch.randint(0, 2 ** 62, (m, 3), generator=g, dtype=torch.int64)
            pr = torch.rand((m, 140), generator=g)
            tp.data_buffer.extend(st, pr / pr.sum(1, keepdim=True), torch.randint(0, 2, (m,), generator=g).float() * 2 - 1)
        feed(40 if rank == 0 else 10)                         # rank 1 is NOT ready: nobody may update (or hang)
        ready_first = tp.ready_to_update()
        feed(30)
        ready_second = tp.ready_to_update()
        stats = []
        for _ in range(3):
            tp.policy_update()
            stats.append((tp.last_stats["epochs"], tp.last_stats["kl"], tp.lr_multiplier))
        flat = torch.cat([p.detach().reshape(-1) for p in tp.policy_value_net.policy_value_net.parameters()])
        q.put((rank, ready_first, ready_second, stats, start.numpy(), flat.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_pipeline_collectives_are_rank_consistent():
    """TrainPipeline on two ranks with different buffers, minibatches and initial seeds: the update gate, the KL early
    stop and the learning-rate multiplier are decided collectively (same number of all-reduces on every rank -- a
    rank-local decision here deadlocks), and the ranks start and end with identical weights."""
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in (q.get(timeout=280) for _ in range(ws))}
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    import numpy as np
    assert res[0][0] is False and res[1][0] is False and res[0][1] is True and res[1][1] is True
    assert res[0][2] == res[1][2]                              # epochs run, mean KL, lr multiplier: identical
    assert np.array_equal(res[0][3], res[1][3])                # broadcast at construction
    assert np.array_equal(res[0][4], res[1][4]) and np.isfinite(res[0][4]).all()
    assert not np.array_equal(res[0][3], res[0][4])


def test_replay_ring_matches_deque():
    """ReplayBuffer == collections.deque(maxlen) of train.py:24 (order, eviction), sampling without replacement."""
    from collections import deque
    from alphazero_quoridor_b200.train import ReplayBuffer
    rb, dq = ReplayBuffer(50, "cpu"), deque(maxlen=50)
    g = torch.Generator().manual_seed(0)
    k = 0
    for m in (7, 30, 20, 1, 49, 120, 3):
        st = torch.arange(k, k + m, dtype=torch.int64)[:, None].expand(-1, 3).contiguous()
        rb.extend(st, torch.full((m, 140), 1.0 / 140), torch.ones(m))
        dq.extend(range(k, k + m))
        k += m
        assert len(rb) == len(dq) and rb.ordered()[0][:, 0].tolist() == list(dq)
        if len(rb) >= 10:
            s, _, _ = rb.sample(10, generator=g)
            vals = s[:, 0].tolist()
            assert len(set(vals)) == 10 and set(vals) <= set(dq)
    assert rb[0][0][0].item() == dq[0] and rb[0][2] == 1.0


def test_mirror_samples_is_an_involution_on_cpu():
    """train.mirror_samples (left-right symmetry of states and the 140 move probabilities) is plain tensor arithmetic:
    applying it twice is the identity, the action permutation is a bijection that maps wall (r, c) to (r, 7 - c), and
    z is untouched.  (The GPU suite checks it against the kernels' planes and legal lists.)"""
    from alphazero_quoridor_b200.train import MIRROR_ACTION, mirror_samples
    g = torch.Generator().manual_seed(1)
    n = 64
    H = torch.randint(0, 2 ** 62, (n,), generator=g, dtype=torch.int64)
    V = torch.randint(0, 2 ** 62, (n,), generator=g, dtype=torch.int64) & ~H
    p1 = torch.randint(0, 81, (n,), generator=g)
    p2 = torch.randint(0, 81, (n,), generator=g)
    meta = p1 | (p2 << 8) | (torch.randint(0, 11, (n,), generator=g) << 16) | (torch.randint(0, 11, (n,), generator=g) << 24) \
        | (torch.randint(1, 3, (n,), generator=g) << 32) | (torch.randint(0, 200, (n,), generator=g) << 48)
    st = torch.stack([H, V, meta], 1)
    pr = torch.rand((n, 140), generator=g)
    m_st, m_pr = mirror_samples(st, pr)
    b_st, b_pr = mirror_samples(m_st, m_pr)
    assert torch.equal(b_st, st) and torch.equal(b_pr, pr)
    assert sorted(MIRROR_ACTION.tolist()) == list(range(140))
    assert MIRROR_ACTION[12 + 8 * 3 + 1].item() == 12 + 8 * 3 + 6 and MIRROR_ACTION[76 + 8 * 5 + 0].item() == 76 + 8 * 5 + 7
    assert MIRROR_ACTION[2].item() == 3 and MIRROR_ACTION[8].item() == 9           # E <-> W, NE <-> NW
    assert torch.equal((m_st[:, 2] >> 16), (st[:, 2] >> 16))                        # walls left, mover, flags, ply untouched
    assert torch.equal((m_st[:, 2] & 0xFF) % 9, 8 - (st[:, 2] & 0xFF) % 9)          # pawn columns mirrored
    assert torch.equal((m_st[:, 2] & 0xFF) // 9, (st[:, 2] & 0xFF) // 9)            # rows kept
