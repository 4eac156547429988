"""Batched self-play: thousands of concurrent `Quoridor.start_self_play` loops (quoridor.py:573-610) driven by
`tree.BatchedMCTS`, i.e. the batched form of `TrainPipeline.collect_selfplay_data` (train.py:55-63).

One `step()` = one ply for every game: search (n_playout playouts per game), pick a move
(mcts.py:177-187 / pure_mcts.py:115), optionally record (state, move probabilities, mover), play the move and
re-root (mcts.py:146-151); games that end are scored, handed to the replay sink and restarted from `reset()`.

Games are identified by a GLOBAL game index (`game_id_base + i`), which keys every random stream, so a set of
games gives the same results however it is sharded over GPUs (no collective on this path).
"""
import torch

from .quoridor import BatchedQuoridor
from .tree import BatchedMCTS


class BatchedSelfPlay:
    def __init__(self, n_games, evaluator, c_puct=5, n_playout=400, leaves_per_game=8, temp=1.0, pure=False,
                 seed=0, game_id_base=0, max_plies=600, record=False, fix_terminal_sign=False, device=None,
                 node_cap=None, defer_depth=0, defer_until_drain=False):
        self.pure = bool(pure)
        self.temp = float(temp)
        self.seed = int(seed)
        self.max_plies = int(max_plies)
        self.record = bool(record)
        self.mcts = BatchedMCTS(n_games, evaluator, c_puct=c_puct, n_playout=n_playout,
                                leaves_per_game=leaves_per_game, reuse_tree=not self.pure,
                                fix_terminal_sign=fix_terminal_sign, device=device, node_cap=node_cap,
                                defer_depth=defer_depth, defer_until_drain=defer_until_drain)
        self.n = self.mcts.n
        dev = self.device = self.mcts.device
        self.lib = self.mcts.lib
        # games per slot so far; RNG stream id of slot i's current game = games_played << 20 | (game_id_base + i)
        # (BatchedMCTS keys a rollout by (this id, playout counter), so ids only need to be distinct)
        self.game_id_base = int(game_id_base)
        assert 0 <= self.game_id_base and self.game_id_base + self.n <= (1 << 20), "global game index must be < 2^20"
        self.games_started = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self._slot = torch.arange(self.n, dtype=torch.int64, device=dev) + self.game_id_base
        self._start = BatchedQuoridor(1, device=dev).states
        self.mcts.reset(self._start.expand(self.n, 3).contiguous())
        self._set_game_ids()
        self.finished_games = torch.zeros((), dtype=torch.int64, device=dev)
        self.p1_wins = torch.zeros((), dtype=torch.int64, device=dev)
        self.truncated_games = torch.zeros((), dtype=torch.int64, device=dev)
        self.stalemated_games = torch.zeros((), dtype=torch.int64, device=dev)
        self.finished_plies = torch.zeros((), dtype=torch.int64, device=dev)     # plies of the finished games
        self.moves_played = 0
        if self.record:
            T = self.max_plies
            self.rec_state = torch.zeros((T, self.n, 3), dtype=torch.int64, device=dev)
            self.rec_probs = torch.zeros((T, self.n, 140), dtype=torch.float32, device=dev)
            self._tgrid = torch.arange(T, device=dev)
            # one entry per flush: (states int64 [m,3], probs f32 [m,140], z f32 [m], lengths int64 [games]) with the
            # samples of a game contiguous and in ply order; a replay buffer consumes these as they are
            self.flushed = []

    def _set_game_ids(self):
        self.mcts.game_id.copy_((self.games_started << 20) | self._slot)

    @property
    def sink(self):
        """The flushed samples split per finished game: [(states [T,3], probs [T,140], z [T]), ...]."""
        out = []
        for st, pr, z, lens in self.flushed:
            ls = lens.tolist()
            out.extend(zip(st.split(ls), pr.split(ls), z.split(ls)))
        return out

    def wave_plan(self):
        """Leaves per wave of one move's search (mirrors BatchedMCTS.search)."""
        m = self.mcts
        plan, done, first = [], 0, True
        while done < m.n_playout:
            k = 1 if first else min(m.K, m.n_playout - done)
            plan.append(k)
            done += k
            first = False
        return plan

    def finish_move(self):
        """Everything of step() after the search."""
        return self._after_search()

    def step(self):
        self.mcts.search()
        return self._after_search()

    def _after_search(self):
        m = self.mcts
        if self.pure:
            moves = m.choose(mode=0)
        else:
            moves = m.choose(mode=2, temp=self.temp, seed=self.seed)
        ply = ((m.root_state[:, 2] >> 48) & 0xFFFF)
        if self.record:
            _, probs, _ = m.root_stats(temp=self.temp)
            t = ply.clamp(max=self.max_plies - 1)
            idx = torch.arange(self.n, device=self.device)
            self.rec_state[t, idx] = m.root_state
            self.rec_probs[t, idx] = probs.float()
        m.advance(moves, keep_subtree=not self.pure)
        self.moves_played += self.n
        self._finish_games(moves)
        return moves

    def _finish_games(self, moves):
        m = self.mcts
        meta = m.root_state[:, 2]
        done = ((meta >> 40) & 1).bool()
        ply = (meta >> 48) & 0xFFFF
        trunc = (~done) & (ply >= self.max_plies)
        # a root whose mover has no legal action (QZ_FLAG_STALEMATE positions; the reference prints and returns None,
        # mcts.py:195-196, then crashes in step): choose() gave -1 and nothing was played -- the game ends as a tie
        stale = (~done) & (~trunc) & (moves < 0)
        over = done | trunc | stale
        if self.record and not bool(over.any()):      # recording flushes on the host; otherwise stay asynchronous
            return
        winner = torch.where(done, (meta >> 41) & 3, torch.zeros_like(meta))
        self.finished_games += over.sum()
        self.finished_plies += (ply * over).sum()
        self.p1_wins += (over & (winner == 1)).sum()
        self.truncated_games += trunc.sum()
        self.stalemated_games += stale.sum()
        if self.record:
            self._flush(over, winner, ply)
            self.check_overflow()
        sel = over.to(torch.uint8).contiguous()
        fresh = torch.where(over.unsqueeze(1), self._start.expand(self.n, 3), m.root_state).contiguous()
        m.reset(fresh, select=sel)            # fresh root + start position for the finished slots only
        self.games_started += over.to(torch.int64)
        self._set_game_ids()

    def check_overflow(self):
        """Arena overflows so far (an overflowing leaf stays unexpanded for that playout); warns once when non-zero.
        Synchronises -- call it where the host waits anyway."""
        n = self.mcts.overflow_count()
        if n and not getattr(self, "_overflow_warned", False):
            import warnings
            warnings.warn("tree arena overflowed %d times: raise node_cap (now %d slots per game)" % (n, self.mcts.node_cap))
            self._overflow_warned = True
        return n

    def _flush(self, over, winner, ply):
        """quoridor.py:596-610: z = +1 on the winner's plies, -1 on the loser's (0 for a game without a winner).
        One gather for all finished games -- no per-game host loop."""
        idx = over.nonzero().flatten()
        T = ply[idx].clamp(max=self.max_plies)                                   # samples per finished game
        keep = self._tgrid[None, :] < T[:, None]                                 # [games, T_max]
        st = self.rec_state[:, idx].transpose(0, 1)[keep]                        # game-major, ply order
        pr = self.rec_probs[:, idx].transpose(0, 1)[keep]
        w = winner[idx][:, None].expand(-1, self.max_plies)[keep]
        mover = (st[:, 2] >> 32) & 0xFF
        z = torch.where(w == 0, 0.0, torch.where(mover == w, 1.0, -1.0)).to(torch.float32)
        sel = T > 0
        self.flushed.append((st, pr, z, T[sel]))


class StreamedSelfPlay:
    """S independent `BatchedSelfPlay` sub-batches, each on its own CUDA stream, with their MCTS waves issued
    round-robin.  Rollout lengths are heavy-tailed (a wave ends with a few very long rollouts running on a few
    SMs), so interleaving sub-batches lets one sub-batch's tail overlap another's bulk work.  Games keep their
    global indices (`game_id_base + i`), so results are identical to the un-streamed engine."""

    def __init__(self, n_games, make_evaluator, n_streams=4, game_id_base=0, device=None, **kw):
        assert n_games % n_streams == 0
        per = n_games // n_streams
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.subs = []
        for i, st in enumerate(self.streams):
            with torch.cuda.stream(st):
                self.subs.append(BatchedSelfPlay(per, make_evaluator(), game_id_base=game_id_base + i * per,
                                                 device=self.device, **kw))
        self.n = n_games
        torch.cuda.synchronize(self.device)

    @property
    def moves_played(self):
        return sum(s.moves_played for s in self.subs)

    def step(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)
        plans = [s.wave_plan() for s in self.subs]
        for w in range(max(len(p) for p in plans)):
            for sub, st, plan in zip(self.subs, self.streams, plans):
                if w < len(plan):
                    with torch.cuda.stream(st):
                        sub.mcts.playout_wave(plan[w])
        for sub, st in zip(self.subs, self.streams):
            with torch.cuda.stream(st):
                sub.mcts.drain()
                sub.finish_move()
        for st in self.streams:
            cur.wait_stream(st)
