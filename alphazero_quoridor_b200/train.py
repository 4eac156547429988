"""Mirror of the reference's train.py (`TrainPipeline`, train.py:12-116) over the batched engine.

Same hyper-parameter attributes and the same three methods (`collect_selfplay_data`, `policy_update`, `run`).
Self-play is not one game at a time: `collect_selfplay_data(n)` advances `n_parallel_games` concurrent games
(selfplay.BatchedSelfPlay, MCTS with the bf16 net on the device) until at least n of them have finished and
appends their (state, mcts_probs, z) samples to the replay buffer.  A sample keeps the 24-byte game state, not
the 26x9x9 float64 tensor of quoridor.py:589; `policy_update` re-encodes the minibatch with the encode kernel.

SURVEY.md 8(f) lists the trainer as "next"; it is provided so that the self-play path has its real caller.  With
torch.distributed initialised (one process per GPU, NCCL) gradients are averaged with an all-reduce per step --
the only collective in the system -- and every rank collects its own shard of games.
"""
from __future__ import print_function

import numpy as np
import torch
import torch.distributed as dist

from .policy_value_net import PolicyValueNet
from .quoridor import BatchedQuoridor
from .selfplay import BatchedSelfPlay
from .tree import NetEvaluator


def allreduce_gradients(parameters):
    """Average the gradients over the ranks (NCCL over NVLink on GPUs, gloo in the CPU tests): the only collective
    in the system.  453,041 parameters = 1.8 MB in fp32, latency-bound; one flattened all-reduce per step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


# ---- left-right mirror symmetry ------------------------------------------------------------------
# tile (r,c) -> (r,8-c); intersection (r,c) -> (r,7-c); pawn actions E<->W, EE<->WW, NE<->NW, SE<->SW.  The commented-out
# get_equi_data of train.py:40-53 is Gomoku-shaped and wrong for Quoridor; this is the symmetry the board really has.
# NOTE the reference's row-0 corner aliasing (quoridor.py:388,:392) is itself mirror-asymmetric, so a mirrored position
# is only approximately equivalent under the reference's rules (exactly equivalent away from row 0).
_MIRROR_PAWN = [0, 1, 3, 2, 4, 5, 7, 6, 9, 8, 11, 10]
MIRROR_ACTION = torch.tensor(_MIRROR_PAWN + [12 + (ix // 8) * 8 + (7 - ix % 8) for ix in range(64)]
                             + [76 + (ix // 8) * 8 + (7 - ix % 8) for ix in range(64)], dtype=torch.int64)


def _mirror_mask64(m):
    """Reverse the 8 columns inside every row of an 8x8 bit mask held in int64 tensors (byte-wise bit reversal)."""
    m = ((m >> 1) & 0x5555555555555555) | ((m & 0x5555555555555555) << 1)
    m = ((m >> 2) & 0x3333333333333333) | ((m & 0x3333333333333333) << 2)
    m = ((m >> 4) & 0x0F0F0F0F0F0F0F0F) | ((m & 0x0F0F0F0F0F0F0F0F) << 4)
    return m


def mirror_samples(states, probs):
    """Left-right mirrored copies of self-play samples: states int64 [m,3] (qz_state rows with on-board pawns), probs
    [m,140].  z is unchanged by the symmetry."""
    H, V, meta = states[:, 0], states[:, 1], states[:, 2]
    p1, p2 = meta & 0xFF, (meta >> 8) & 0xFF
    p1m = (p1 // 9) * 9 + (8 - p1 % 9)
    p2m = (p2 // 9) * 9 + (8 - p2 % 9)
    meta_m = (meta & ~0xFFFF) | p1m | (p2m << 8)
    out = torch.stack([_mirror_mask64(H), _mirror_mask64(V), meta_m], 1)
    return out, probs[:, MIRROR_ACTION.to(probs.device)]


def play_match(engine_a, engine_b, max_plies=300):
    """Arena: two `BatchedMCTS` engines (same number of games, fresh trees every move) play n games against each other
    side by side on the device; A moves first in even games, second in odd ones; every move is the most visited one
    (mcts.py:185 at temp -> 0 / pure_mcts.py:115).  Returns wins / ties and A's score ratio (ties count half), as the
    reference's commented-out policy_evaluate would (train.py:30-31,108)."""
    n, dev = engine_a.n, engine_a.device
    assert engine_b.n == n
    env = BatchedQuoridor(n, device=dev)
    a_color = (torch.arange(n, device=dev) % 2) + 1                         # 1: A moves first, 2: second
    plies = 0
    for plies in range(max_plies):
        meta = env.states[:, 2]
        if bool((((meta >> 40) & 1) == 1).all()):
            break
        engine_a.reset(env.states)
        engine_a.search()
        engine_b.reset(env.states)
        engine_b.search()
        mover = (meta >> 32) & 0xFF
        moves = torch.where(mover == a_color, engine_a.choose(mode=0), engine_b.choose(mode=0))
        env.step(moves)                                                       # finished games / stalemated movers stay put
    meta = env.states[:, 2]
    winner = torch.where(((meta >> 40) & 1) == 1, (meta >> 41) & 3, torch.zeros_like(meta))
    wins_a = int((winner == a_color).sum().item())
    ties = int((winner == 0).sum().item())
    return {"games": n, "wins_a": wins_a, "wins_b": n - wins_a - ties, "ties": ties, "plies": plies + 1,
            "win_ratio_a": (wins_a + 0.5 * ties) / n}


class ReplayBuffer(object):
    """`deque(maxlen=buffer_size)` of train.py:24 as a ring of device tensors: the 24-byte game state (re-encoded by
    the encode kernel when a minibatch is drawn, instead of the float64 [26,9,9] array of quoridor.py:589), the 140
    move probabilities and z.  Appends and sampling never leave the device and never loop over games."""

    def __init__(self, maxlen, device):
        self.maxlen = int(maxlen)
        self.device = torch.device(device)
        self.states = torch.zeros((self.maxlen, 3), dtype=torch.int64, device=self.device)
        self.probs = torch.zeros((self.maxlen, 140), dtype=torch.float32, device=self.device)
        self.z = torch.zeros((self.maxlen,), dtype=torch.float32, device=self.device)
        self.size = 0           # valid entries
        self.head = 0           # next slot to write (the oldest entry once the ring is full)

    def __len__(self):
        return self.size

    def extend(self, states, probs, z):
        """Append m samples (oldest entries are overwritten, like deque(maxlen))."""
        m = states.shape[0]
        if m == 0:
            return
        if m > self.maxlen:
            states, probs, z = states[-self.maxlen:], probs[-self.maxlen:], z[-self.maxlen:]
            m = self.maxlen
        pos = (self.head + torch.arange(m, device=self.device)) % self.maxlen
        self.states[pos] = states.to(self.device)
        self.probs[pos] = probs.to(self.device, torch.float32)
        self.z[pos] = z.to(self.device, torch.float32)
        self.head = (self.head + m) % self.maxlen
        self.size = min(self.maxlen, self.size + m)

    def ordered(self):
        """(states, probs, z) oldest first."""
        if self.size < self.maxlen:
            sl = slice(0, self.size)
            return self.states[sl], self.probs[sl], self.z[sl]
        idx = (self.head + torch.arange(self.maxlen, device=self.device)) % self.maxlen
        return self.states[idx], self.probs[idx], self.z[idx]

    def sample(self, batch_size, generator=None):
        """`random.sample(buffer, batch_size)` (train.py:67): without replacement."""
        assert batch_size <= self.size
        perm = torch.randperm(self.size, generator=generator, device=self.device)[:batch_size]
        if self.size == self.maxlen:
            perm = (self.head + perm) % self.maxlen
        return self.states[perm], self.probs[perm], self.z[perm]

    def clear(self):
        self.size = self.head = 0

    def __getitem__(self, i):
        st, pr, z = self.ordered()
        return st[i], pr[i], float(z[i].item())


def _dist_on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_module(module, src=0):
    """Every rank starts from rank `src`'s parameters AND buffers (BatchNorm running statistics)."""
    if not _dist_on():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


class TrainPipeline(object):
    def __init__(self, init_model=None, n_parallel_games=256, leaves_per_game=4, device=None, seed=0,
                 fix_terminal_sign=False, max_plies=600, use_gpu=True, encode_fn=None):
        # train.py:17-31
        self.learn_rate = 2e-3
        self.lr_multiplier = 1.0
        self.temp = 1.0
        self.n_playout = 400
        self.c_puct = 5
        self.buffer_size = 10000
        self.batch_size = 128
        self.play_batch_size = 1
        self.epochs = 5
        self.kl_targ = 0.02
        self.check_freq = 50
        self.game_batch_num = 1500
        self.best_win_ratio = 0.0
        self.pure_mcts_playout_num = 1000
        self.policy_value_net = PolicyValueNet(model_file=init_model, device=device, use_gpu=use_gpu)
        self.data_buffer = ReplayBuffer(self.buffer_size, self.policy_value_net.device)
        self.n_parallel_games = n_parallel_games
        self.leaves_per_game = leaves_per_game
        self.seed = seed
        self.fix_terminal_sign = fix_terminal_sign
        self.max_plies = max_plies
        self._selfplay = None
        self._encode_fn = encode_fn              # tests on CPU ranks (gloo) supply their own; the product uses the kernel
        self.episode_len = 0
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._gen = torch.Generator(device=self.policy_value_net.device)
        self._gen.manual_seed(int(seed) * 7919 + self.rank)
        broadcast_module(self.policy_value_net.policy_value_net)      # ranks must not start from different inits

    def _engine(self):
        if self._selfplay is None or self._selfplay.mcts.n_playout != self.n_playout:
            self._selfplay = BatchedSelfPlay(
                self.n_parallel_games, NetEvaluator(self.policy_value_net), c_puct=self.c_puct,
                n_playout=self.n_playout, leaves_per_game=self.leaves_per_game, temp=self.temp, pure=False,
                seed=self.seed, game_id_base=self.rank * self.n_parallel_games, max_plies=self.max_plies,
                record=True, fix_terminal_sign=self.fix_terminal_sign, device=self.policy_value_net.device)
        return self._selfplay

    def collect_selfplay_data(self, n_games=1, max_steps=100000):
        """train.py:55-63: play until >= n_games more games have finished; extend the replay buffer.  The samples
        go from the self-play record tensors to the replay ring on the device (one append per flush)."""
        sp = self._engine()
        steps = 0
        start = int(sp.finished_games.item())
        while int(sp.finished_games.item()) - start < n_games and steps < max_steps:
            sp.step()
            steps += 1
            for st, pr, z, lens in sp.flushed:
                self.data_buffer.extend(st, pr, z)
                self.episode_len = int(lens[-1].item()) if lens.numel() else self.episode_len
            sp.flushed = []
        finished = int(sp.finished_games.item()) - start
        sp.check_overflow()
        return finished

    def _encode(self, state_rows):
        """qz_state rows int64 [B,3] -> float32 planes [B,26,9,9] (quoridor.py:58-131) by the encode kernel."""
        if self._encode_fn is not None:
            return self._encode_fn(state_rows)
        rows = state_rows.to(self.policy_value_net.device).contiguous()
        return BatchedQuoridor(rows.shape[0], states=rows, device=self.policy_value_net.device).encode(dtype=torch.float32)

    def _sync_gradients(self):
        allreduce_gradients(list(self.policy_value_net.policy_value_net.parameters()))

    def _all_ranks(self, flag):
        """True iff `flag` holds on EVERY rank (one MIN all-reduce): decisions that gate a collective must be collective."""
        if not _dist_on():
            return bool(flag)
        t = torch.tensor([1.0 if flag else 0.0], device=self.policy_value_net.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def _rank_mean(self, x):
        if not _dist_on():
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.policy_value_net.device)
        dist.all_reduce(t)
        return float(t.item()) / dist.get_world_size()

    def ready_to_update(self):
        """train.py:100 `len(self.data_buffer) > self.batch_size`, on every rank."""
        return self._all_ranks(len(self.data_buffer) > self.batch_size)

    def policy_update(self):
        """train.py:65-92: <= 5 epochs on one minibatch, KL early stop, KL-adaptive learning-rate multiplier.
        Multi-rank: every rank draws its own minibatch, gradients are averaged (one all-reduce per epoch), and the KL
        that steers the early stop and the learning-rate multiplier is the mean over the ranks, so every rank makes the
        same number of collective calls and keeps the same multiplier and weights."""
        st, pr, z = self.data_buffer.sample(self.batch_size, generator=self._gen)
        state_batch = self._encode(st)
        net = self.policy_value_net
        old_probs, old_v = net.policy_value(state_batch)
        winner_batch = z.cpu().numpy()
        loss = entropy = kl = new_v = None
        for i in range(self.epochs):
            loss, entropy = net.train_step(state_batch, pr, z, self.learn_rate * self.lr_multiplier,
                                           grad_hook=self._sync_gradients)
            new_probs, new_v = net.policy_value(state_batch)
            kl = np.mean(np.sum(old_probs * (np.log(old_probs + 1e-10) - np.log(new_probs + 1e-10)), axis=1))
            kl = self._rank_mean(kl)
            if kl > self.kl_targ * 4:
                break
        if kl > self.kl_targ * 2 and self.lr_multiplier > 0.1:
            self.lr_multiplier /= 1.5
        elif kl < self.kl_targ / 2 and self.lr_multiplier < 10:
            self.lr_multiplier *= 1.5
        var = np.var(winner_batch)
        self.last_stats = dict(kl=float(kl), lr_multiplier=self.lr_multiplier, loss=loss, entropy=entropy, epochs=i + 1,
                               explained_var_old=float(1 - np.var(winner_batch - old_v.flatten()) / var) if var > 0 else 0.0,
                               explained_var_new=float(1 - np.var(winner_batch - new_v.flatten()) / var) if var > 0 else 0.0)
        return loss, entropy

    # ---- checkpoint / resume ----
    def save_checkpoint(self, path, model_name=None):
        """Everything needed to resume: the net (also written in the reference's own format, ckpt/<name>.pth, when
        `model_name` is given: policy_value_net.py:198-200), plus what the reference does not save -- optimizer state,
        learning-rate multiplier and the replay buffer."""
        if model_name:
            self.policy_value_net.save_model(model_name)
        st, pr, z = self.data_buffer.ordered()
        torch.save({"net": self.policy_value_net.get_policy_param(),
                    "optimizer": self.policy_value_net.optimizer.state_dict(),
                    "lr_multiplier": self.lr_multiplier,
                    "buffer_states": st.cpu(), "buffer_probs": pr.cpu(), "buffer_z": z.cpu()}, path)

    def load_checkpoint(self, path):
        ck = torch.load(path, map_location=self.policy_value_net.device)
        self.policy_value_net.policy_value_net.load_state_dict(ck["net"])
        self.policy_value_net.optimizer.load_state_dict(ck["optimizer"])
        self.policy_value_net._infer = None
        self.lr_multiplier = ck["lr_multiplier"]
        self.data_buffer.clear()
        self.data_buffer.extend(ck["buffer_states"], ck["buffer_probs"], ck["buffer_z"])
        broadcast_module(self.policy_value_net.policy_value_net)

    # ---- evaluation arena (the commented-out policy_evaluate of train.py:30-31,108) ----
    def policy_evaluate(self, n_games=64, n_playout=None, max_plies=300, seed=0):
        """Win ratio of the current net's MCTS player against pure MCTS (`pure_mcts_playout_num` rollouts per move),
        all games played side by side on the device; the net plays first in even games, second in odd ones.  Both
        sides search with the corrected terminal sign (a player that avoids winning moves cannot be rated)."""
        from .tree import BatchedMCTS, RolloutEvaluator
        dev = self.policy_value_net.device
        npl = self.n_playout if n_playout is None else n_playout
        az = BatchedMCTS(n_games, NetEvaluator(self.policy_value_net), c_puct=self.c_puct, n_playout=npl,
                         leaves_per_game=self.leaves_per_game, reuse_tree=False, fix_terminal_sign=True, device=dev)
        pure = BatchedMCTS(n_games, RolloutEvaluator(seed=seed), c_puct=5, n_playout=self.pure_mcts_playout_num,
                           leaves_per_game=16, reuse_tree=False, fix_terminal_sign=True, device=dev)
        res = play_match(az, pure, max_plies=max_plies)
        self.last_match = res
        return res["win_ratio_a"]

    def run(self):
        """train.py:94-111"""
        try:
            for i in range(self.game_batch_num):
                self.collect_selfplay_data(self.play_batch_size)
                if self.ready_to_update():              # collective decision: no rank may skip the all-reduces
                    loss, entropy = self.policy_update()
                    if self.rank == 0:
                        with open('loss.txt', 'a') as f:
                            f.writelines(str(loss) + '\n')
                if (i + 1) % self.check_freq == 0 and self.rank == 0:
                    self.policy_value_net.save_model('current_policy')
        except KeyboardInterrupt:
            print('\n\rquit')


if __name__ == '__main__':
    TrainPipeline().run()
