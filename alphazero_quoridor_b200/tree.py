"""Batched PUCT MCTS on the GPU: the engine behind `mcts.MCTS` / `pure_mcts.MCTS` (mcts.py, pure_mcts.py).

`BatchedMCTS` owns the flat tree arrays (torch CUDA tensors) of n concurrent games and drives the kernels of
libqzb200.so, one wave = select -> legal masks -> evaluate -> expand + backup (stored priors), or
select -> sweep of the revisited leaves -> extend -> rollouts -> backup (uniform priors, lazy expansion).  A child gets
its slot in the arrays the first time the descent picks it (csrc/qz_mcts.cu).  Evaluators:

* `StubEvaluator(kind)`     deterministic parity stubs S1/S2/S3 (tests/golden/stubs.py), on the device
* `RolloutEvaluator(seed)`  pure MCTS: uniform priors + random rollout (pure_mcts.py:13-16,86-108)
* `NetEvaluator(net)`       the 5-block ResNet in PyTorch bf16 (policy_value_net.py), leaves encoded straight
                            into its input tensor by the encode kernel

With `leaves_per_game=1` the search is the reference's sequential algorithm, one playout per game per wave,
and visit counts match the reference exactly under a deterministic evaluator; `leaves_per_game=K>1` collects K
leaves per game per wave with virtual loss (a documented deviation that trades exactness for batch size).
"""
import ctypes as C

import torch

from . import _lib
from .rollout import rollout as _rollout

LEAF_TERMINAL, LEAF_DEPTH_OVERFLOW, LEAF_ARENA_OVERFLOW, LEAF_DUPLICATE, LEAF_INACTIVE, LEAF_PENDING = 1, 2, 4, 8, 16, 32
LEAF_NEEDS_MASK = 64
MAX_CHILDREN = 140
SLOTS_PER_PLAYOUT = 12                 # default arena slots per playout kept (see BatchedMCTS.__init__)
MAX_K_NONUNIFORM = 8                   # leaves per game per wave allowed under non-uniform priors (see BatchedMCTS)
PLAYOUT_BITS = 24                      # rollout stream id = game_id << 24 | playout counter (mod 2^24)
PLAYOUT_MASK = (1 << PLAYOUT_BITS) - 1


class QzTree(C.Structure):
    """struct qz_tree of include/qzb200.h."""
    _fields_ = [("n_games", C.c_int64), ("node_cap", C.c_int32), ("max_depth", C.c_int32),
                ("leaves_per_game", C.c_int32), ("pool_cap", C.c_int32),
                ("prior", C.c_void_p), ("visits", C.c_void_p), ("q", C.c_void_p), ("child_base", C.c_void_p),
                ("node_meta", C.c_void_p), ("root", C.c_void_p), ("n_nodes", C.c_void_p),
                ("root_state", C.c_void_p), ("leaf_node", C.c_void_p), ("leaf_state", C.c_void_p),
                ("path", C.c_void_p), ("path_len", C.c_void_p), ("leaf_flags", C.c_void_p),
                ("prior_pool", C.c_void_p), ("n_pool", C.c_void_p)]


class _Arena:
    """One set of node arrays for n games (a second one is the re-root compaction target)."""

    def __init__(self, n, cap, dev, pool_cap=0):
        tot = n * cap
        # stored (non-uniform) priors: one float per legal action of every expanded node (children get their slot when
        # first visited); pool_cap floats per game
        self.prior_pool = torch.empty(n * pool_cap, dtype=torch.float32, device=dev) if pool_cap else None
        self.n_pool = torch.zeros(n, dtype=torch.int32, device=dev) if pool_cap else None
        self.prior = torch.empty(tot, dtype=torch.float32, device=dev)
        self.visits = torch.empty(tot, dtype=torch.int32, device=dev)
        self.q = torch.empty(tot, dtype=torch.float64, device=dev)
        self.child_base = torch.empty(tot, dtype=torch.int32, device=dev)
        self.node_meta = torch.empty(tot, dtype=torch.int32, device=dev)
        self.root = torch.zeros(n, dtype=torch.int32, device=dev)
        self.n_nodes = torch.zeros(n, dtype=torch.int32, device=dev)
        self.root_state = torch.zeros((n, 3), dtype=torch.int64, device=dev)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in vars(self).values() if torch.is_tensor(t))


# ------------------------------------------------------------------------------------------------ evaluators
class _LeafSet:
    """The per-wave leaf arrays of struct qz_tree.  A search that defers stuck rollouts keeps several sets in
    flight (the select of wave w+1 must not overwrite the paths wave w's deferred backups still need)."""

    def __init__(self, m, max_depth, dev):
        self.leaf_node = torch.empty(m, dtype=torch.int32, device=dev)
        self.leaf_state = torch.zeros((m, 3), dtype=torch.int64, device=dev)
        self.path = torch.empty(m * max_depth, dtype=torch.int32, device=dev)
        self.path_len = torch.empty(m, dtype=torch.int32, device=dev)
        self.leaf_flags = torch.zeros(m, dtype=torch.uint8, device=dev)
        self.leaf_mask = torch.zeros((m, 3), dtype=torch.int64, device=dev)
        self.owed = None            # evaluator state of a deferred wave (None = nothing owed)


class StubEvaluator:
    """Parity stubs evaluated by the qz_stub_eval kernel.  kind: 1 = S1 uniform, 2 = S2 hash, 3 = S3 hash/8."""

    def __init__(self, kind):
        self.kind = {"S1": 1, "S2": 2, "S3": 3}.get(kind, kind)
        self._buf = None

    def evaluate(self, mcts, lset, leaf_rids):
        m = lset.leaf_state.shape[0]
        if self._buf is None or self._buf[0].shape[0] != m:
            self._buf = (torch.empty((m, 140), dtype=torch.float32, device=lset.leaf_state.device),
                         torch.empty((m,), dtype=torch.float64, device=lset.leaf_state.device))
        priors, values = self._buf
        _lib.check(mcts.lib.qz_stub_eval(_lib.ptr(lset.leaf_state), _lib.ptr(lset.leaf_mask), self.kind, _lib.ptr(priors),
                                         _lib.ptr(values), m, mcts._stream()), "qz_stub_eval")
        return dict(priors=priors, value_f64=values)


class RolloutEvaluator:
    """pure_mcts.policy_value_fn + _evaluate_rollout: uniform priors, value = random playout (<= limit-1 plies).

    With `defer_stuck` (used by BatchedMCTS(defer_depth >= 2)) the few stuck rollouts of a wave are finished on a
    side stream while later waves run; their leaves are expanded at once and backed up when the result is in."""

    uniform_prior = True
    can_defer = True

    def __init__(self, seed=0, limit=1000):
        self.seed, self.limit = seed, limit
        self._per_set = {}

    def _buffers(self, lset):
        b = self._per_set.get(id(lset))
        if b is None:
            from .rollout import workspace_words
            m, dev = lset.leaf_state.shape[0], lset.leaf_state.device
            b = dict(ws=torch.zeros((workspace_words(m),), dtype=torch.int64, device=dev),
                     result=torch.empty((m,), dtype=torch.int8, device=dev))
            self._per_set[id(lset)] = b
        return b

    def plies_played(self):
        """Total rollout plies so far (sums the cumulative counters of the workspaces; synchronises)."""
        return sum(int(b["ws"][1].item()) for b in self._per_set.values())

    def zero_plies(self):
        for b in self._per_set.values():
            b["ws"][1] = 0

    def evaluate(self, mcts, lset, leaf_rids, defer=False):
        b = self._buffers(lset)
        _rollout(lset.leaf_state, seed=self.seed, rids=leaf_rids, state_index=mcts._leaf_iota, limit=self.limit,
                 return_plies=False, workspace=b["ws"], result=b["result"], defer_stuck=defer)
        if defer:
            lset.owed = dict(rids=leaf_rids)
        return dict(priors=None, value_i8=b["result"])

    def finish(self, mcts, lset):
        """The deferred pass of `evaluate(..., defer=True)` for this leaf set (call on the side stream)."""
        b = self._buffers(lset)
        _rollout(lset.leaf_state, seed=self.seed, rids=lset.owed["rids"], state_index=mcts._leaf_iota,
                 limit=self.limit, return_plies=False, workspace=b["ws"], result=b["result"], finish=True)
        return b["result"]


class NetEvaluator:
    """Leaf evaluation by the policy-value net (policy_value_net.py:145-164 contract, batched).
    `net` is an alphazero_quoridor_b200.policy_value_net.PolicyValueNet."""

    wants_k = True        # evaluate() takes the number of leaves per game this wave really collected

    def __init__(self, net):
        self.net = net
        self._full = None

    def evaluate(self, mcts, lset, leaf_rids, k_leaves=None):
        m = lset.leaf_state.shape[0]
        if k_leaves is None or k_leaves >= mcts.K:
            probs, value = self.net.evaluate_states(lset.leaf_state)
            return dict(priors=probs, value_f32=value)
        # a wave with unused leaf slots (the first wave of a search collects one leaf per game): only the used slots go
        # through the net; expansion never reads the rows of the others
        idx = mcts.active_slots(k_leaves)
        probs_c, value_c = self.net.evaluate_states(lset.leaf_state.index_select(0, idx))
        if self._full is None or self._full[0].shape[0] != m:
            self._full = (torch.zeros((m, probs_c.shape[1]), dtype=probs_c.dtype, device=probs_c.device),
                          torch.zeros((m,), dtype=value_c.dtype, device=value_c.device))
        self._full[0].index_copy_(0, idx, probs_c)
        self._full[1].index_copy_(0, idx, value_c)
        return dict(priors=self._full[0], value_f32=self._full[1])


# ------------------------------------------------------------------------------------------------ the search
class BatchedMCTS:
    def __init__(self, n_games, evaluator, c_puct=5, n_playout=100, leaves_per_game=1, node_cap=None,
                 max_depth=128, fix_terminal_sign=False, reuse_tree=True, device=None, defer_depth=0,
                 defer_until_drain=False, allow_large_k=False, reuse_factor=4, lazy_expand=True):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = int(n_games)
        self.K = int(leaves_per_game)
        self.c_puct = float(c_puct)
        self.n_playout = int(n_playout)
        self.evaluator = evaluator
        self.uniform_prior = bool(getattr(evaluator, "uniform_prior", False))
        # Virtual loss is only a mild perturbation while the leaves of a wave spread over many near-equal children.
        # Under uniform priors (pure MCTS) K = 64 moves the root visit distribution by a total-variation distance of
        # ~1e-4; under sharply non-uniform priors K = 64 reaches 0.19.  K is therefore capped for
        # evaluators with non-uniform priors unless the caller insists (tests/test_bench_parity.py holds the bounds).
        if not self.uniform_prior and self.K > MAX_K_NONUNIFORM and not allow_large_k:
            raise ValueError("leaves_per_game=%d with a non-uniform-prior evaluator exceeds the parity-tested cap of %d "
                             "(pass allow_large_k=True to override)" % (self.K, MAX_K_NONUNIFORM))
        self.fix_terminal_sign = bool(fix_terminal_sign)
        self.max_depth = int(max_depth)
        # uniform priors (pure MCTS): expand a node when a playout comes BACK to it, not when it is first reached
        self.lazy_expand = self.uniform_prior and bool(lazy_expand)
        # playouts whose nodes one arena may have to hold: one search, or -- with tree reuse -- this move's playouts plus
        # the subtree kept from the moves before.  The reference keeps that subtree without bound (mcts.py:146-151); in
        # forced late-game lines a search keeps most of itself (the recorded reference runs reach 3.5 x n_playout kept
        # visits), so the default holds (reuse_factor + 1) searches; the overflow counter tells when that was too little.
        kept = self.n_playout * ((reuse_factor + 1) if reuse_tree else 1) + self.K
        if node_cap is None:
            # lazy children: a playout adds at most one child slot (<= 4 amortised with block doubling and the copies
            # it leaves behind) and one block (3 header + 4 child slots); the root's block holds all <= 140 children
            node_cap = 160 + kept * SLOTS_PER_PLAYOUT
        self.node_cap = int(node_cap)
        # prior pool: room for two searches of full-width nodes (140 floats each); the deep kept trees live in late
        # positions whose nodes have a handful of legal actions, i.e. a handful of floats
        self.pool_cap = 0 if self.uniform_prior else (self.n_playout * (2 if reuse_tree else 1) + self.K + 1) * MAX_CHILDREN
        dev = self.device
        self.arenas = [_Arena(self.n, self.node_cap, dev, self.pool_cap)]
        if reuse_tree:
            self.arenas.append(_Arena(self.n, self.node_cap, dev, self.pool_cap))
        self.cur = 0
        m = self.n * self.K
        # deferred evaluation (stuck rollouts finished on a side stream while later waves run): wave w's owed
        # backups are applied just before wave w + defer_depth reuses its leaf set
        # defer_until_drain: the deferred passes of ALL waves of a search are launched together when the search
        # drains (before statistics are read / a move is chosen): thousands of stuck rollouts at once fill the SMs,
        # where one wave's few hundred are a latency-bound trickle that crowds the next waves' kernels
        # (profiles/README.md r1m); needs one leaf set per wave of a search.
        self.defer_depth = int(defer_depth) if getattr(evaluator, "can_defer", False) else 0
        self.defer_until_drain = bool(defer_until_drain) and getattr(evaluator, "can_defer", False)
        if self.defer_until_drain:
            self.defer_depth = max(self.defer_depth, 2 + (self.n_playout + self.K - 1) // self.K)
        self.sets = [_LeafSet(m, self.max_depth, dev) for _ in range(max(1, self.defer_depth))]
        self.cur_set = 0
        self.wave_index = 0
        self._active_slots = {}
        # one side stream per leaf set: the stuck passes of consecutive waves are latency-bound (a few hundred
        # long-running blocks each) and must overlap each other, not only the main stream
        self.side_streams = [torch.cuda.Stream(device=dev) for _ in self.sets] if self.defer_depth >= 2 else []
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self._leaf_iota = torch.arange(m, dtype=torch.int32, device=dev)
        self._k_of_leaf = (torch.arange(m, dtype=torch.int64, device=dev) % self.K)
        # RNG stream id of game g (rollouts, move sampling); callers set it to a GLOBAL game index (< 2^40) so that
        # results do not depend on how games are sharded over GPUs.  The Philox stream of a rollout is
        # (game_id << 24) | (playout counter mod 2^24): unique per (game, playout) whatever ids the caller uses.
        self.game_id = torch.arange(self.n, dtype=torch.int64, device=dev)
        self.playouts_done = 0       # playouts since the last reset/advance (per game)
        self.total_playouts = 0
        self.count_tree_steps = False    # bench: accumulate the env steps replayed by the descents
        self.tree_steps = torch.zeros((), dtype=torch.int64, device=dev)
        self._count_sets = []
        self._structs = [[self._make_struct(a, ls) for ls in self.sets] for a in self.arenas]

    # ---- plumbing ----
    def active_slots(self, k):
        """Leaf-slot indices g*K + j (j < k) of a wave that collects k < K leaves per game (int64, cached)."""
        idx = self._active_slots.get(k)
        if idx is None:
            g = torch.arange(self.n, dtype=torch.int64, device=self.device)[:, None] * self.K
            idx = (g + torch.arange(k, dtype=torch.int64, device=self.device)[None, :]).reshape(-1)
            self._active_slots[k] = idx
        return idx

    def _stream(self):
        return _lib.stream_ptr(self.device)

    def _make_struct(self, a, ls):
        t = QzTree()
        t.n_games, t.node_cap, t.max_depth, t.leaves_per_game, t.pool_cap = (self.n, self.node_cap, self.max_depth, self.K,
                                                                              self.pool_cap)
        for name in ("prior", "visits", "q", "child_base", "node_meta", "root", "n_nodes", "root_state"):
            setattr(t, name, getattr(a, name).data_ptr())
        t.prior_pool = a.prior_pool.data_ptr() if a.prior_pool is not None else None
        t.n_pool = a.n_pool.data_ptr() if a.n_pool is not None else None
        for name in ("leaf_node", "leaf_state", "path", "path_len", "leaf_flags"):
            setattr(t, name, getattr(ls, name).data_ptr())
        return t

    @property
    def tree(self):
        return self._structs[self.cur][self.cur_set]

    # the leaf arrays of the most recent wave (read by callback evaluators and tests)
    @property
    def leaf_state(self):
        return self.sets[self.cur_set].leaf_state

    @property
    def leaf_mask(self):
        return self.sets[self.cur_set].leaf_mask

    @property
    def leaf_flags(self):
        return self.sets[self.cur_set].leaf_flags

    @property
    def path_len(self):
        return self.sets[self.cur_set].path_len

    @property
    def leaf_node(self):
        return self.sets[self.cur_set].leaf_node

    @property
    def arena(self):
        return self.arenas[self.cur]

    @property
    def root_state(self):
        return self.arena.root_state

    def nbytes(self):
        return sum(a.nbytes() for a in self.arenas)

    # ---- reference operations ----
    def reset(self, root_states, select=None):
        """Fresh trees (mcts.py:97 / update_with_move(-1)) rooted at `root_states` (int64 [n,3])."""
        root_states = root_states.to(self.device).contiguous()
        assert root_states.shape == (self.n, 3)
        self.drain()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_mcts_init(C.byref(self.tree), _lib.ptr(root_states), _lib.ptr(select),
                                             self._stream()), "qz_mcts_init")
        self.playouts_done = 0

    def _launch_finish(self, idx):
        """Start the deferred pass of wave (self.sets[idx]) on that set's side stream."""
        ls = self.sets[idx]
        cur = torch.cuda.current_stream(self.device)
        side = self.side_streams[idx]
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ls.owed["result"] = self.evaluator.finish(self, ls)
            ls.owed["done"] = torch.cuda.Event()
            ls.owed["done"].record()

    def _settle(self, idx):
        """Apply the backups wave (self.sets[idx]) still owes: wait for its deferred rollouts, then back them up."""
        ls = self.sets[idx]
        if ls.owed is None:
            return
        if "done" not in ls.owed:
            self._launch_finish(idx)
        torch.cuda.current_stream(self.device).wait_event(ls.owed["done"])
        _lib.check(self.lib.qz_mcts_backup_pending(C.byref(self._structs[self.cur][idx]), _lib.ptr(ls.owed["result"]),
                                                   int(self.fix_terminal_sign), self._stream()), "qz_mcts_backup_pending")
        ls.owed = None

    def check_device(self):
        """Wait for this engine's stream and raise QzError if any kernel faulted since the last check (launch return
        codes cannot see asynchronous faults; see qz_stream_check in include/qzb200.h)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_stream_check(self._stream()), "qz_stream_check")

    def counters(self):
        """Structured counters of this engine (synchronises): what a per-wave log line needs."""
        out = {"games": self.n, "leaves_per_game": self.K, "waves": self.wave_index, "playouts_per_game": self.total_playouts,
               "arena_overflows": self.overflow_count(), "nodes_in_use_mean": float(self.arena.n_nodes.double().mean().item()),
               "node_cap": self.node_cap, "launches": _lib.LAUNCHES}
        if hasattr(self.evaluator, "plies_played"):
            out["rollout_plies"] = self.evaluator.plies_played()
        if self.count_tree_steps:
            out["tree_steps"] = int(self.tree_steps.item())
        return out

    def drain(self, check=False):
        """Settle every deferred wave (before reading statistics, choosing a move or re-rooting)."""
        if check:
            self.check_device()
        if self.defer_depth >= 2:
            with torch.cuda.device(self.device):
                order = [(self.cur_set + 1 + i) % len(self.sets) for i in range(len(self.sets))]     # oldest first
                for idx in order:                      # start every pass that is still to run, then collect them
                    if self.sets[idx].owed is not None and "done" not in self.sets[idx].owed:
                        self._launch_finish(idx)
                for idx in order:
                    self._settle(idx)
        if self._count_sets:
            # descent steps of every wave since the last drain in ONE reduction (no per-wave torch kernels)
            lens = torch.stack([ls.path_len for ls in self._count_sets])
            self.tree_steps += (lens.clamp(min=1) - 1).sum()
            self._count_sets = []

    def playout_wave(self, k_leaves=None):
        """One wave: k_leaves playouts per game (mcts.py:103-127)."""
        k = self.K if k_leaves is None else int(k_leaves)
        defer = self.defer_depth >= 2
        with torch.cuda.device(self.device):
            st = self._stream()
            if defer:
                self.cur_set = self.wave_index % len(self.sets)
                self._settle(self.cur_set)          # the wave that used this leaf set defer_depth waves ago
            ls = self.sets[self.cur_set]
            lazy = self.lazy_expand
            _lib.check(self.lib.qz_mcts_select(C.byref(self.tree), self.c_puct, int(self.uniform_prior), k, int(lazy),
                                               _lib.ptr(self.overflow), st), "qz_mcts_select")
            m = self.n * self.K
            if lazy:
                # uniform priors: a leaf's legal actions are only computed when a later playout comes back to it -- swept
                # for the flagged leaves in parallel (a warp per leaf), then the block is built and the descent continued
                _lib.check(self.lib.qz_env_legal_mask_flagged(_lib.ptr(ls.leaf_state), _lib.ptr(ls.leaf_flags),
                                                              LEAF_NEEDS_MASK, _lib.ptr(ls.leaf_mask), m,
                                                              _lib.ptr(ls.leaf_node), self.K, LEAF_DUPLICATE, st),
                           "qz_env_legal_mask_flagged")
                _lib.check(self.lib.qz_mcts_extend(C.byref(self.tree), self.c_puct, _lib.ptr(ls.leaf_mask), 0,
                                                   _lib.ptr(self.overflow), st), "qz_mcts_extend")
            else:
                _lib.check(self.lib.qz_env_legal_mask(_lib.ptr(ls.leaf_state), _lib.ptr(ls.leaf_mask), m, st),
                           "qz_env_legal_mask")
            if self.count_tree_steps:
                if self.defer_until_drain:
                    self._count_sets.append(ls)     # one leaf set per wave: summed once, when the search drains
                else:
                    self.tree_steps += (ls.path_len.clamp(min=1) - 1).sum()
            # rollout / RNG stream of leaf (g,k): (game id, playout counter) in separate bit fields, so two games can
            # never share a stream (pure_mcts.py:86-108 draws fresh randomness for every rollout)
            rids = None
            if getattr(self.evaluator, "uniform_prior", False):
                rids = ((self.game_id << PLAYOUT_BITS).repeat_interleave(self.K)
                        | ((self.total_playouts + self._k_of_leaf) & PLAYOUT_MASK))
            if defer:
                ev = self.evaluator.evaluate(self, ls, rids, defer=True)
            elif getattr(self.evaluator, "wants_k", False):
                ev = self.evaluator.evaluate(self, ls, rids, k_leaves=k)
            else:
                ev = self.evaluator.evaluate(self, ls, rids)
            _lib.check(self.lib.qz_mcts_expand_backup(
                C.byref(self.tree), None if lazy else _lib.ptr(ls.leaf_mask), _lib.ptr(ev.get("priors")),
                _lib.ptr(ev.get("value_f32")), _lib.ptr(ev.get("value_f64")), _lib.ptr(ev.get("value_i8")),
                self.c_puct, int(self.fix_terminal_sign), _lib.ptr(self.overflow), st), "qz_mcts_expand_backup")
            if defer and not self.defer_until_drain:
                # finish the stuck rollouts of this wave on the side stream while the next waves run
                self._launch_finish(self.cur_set)
        self.playouts_done += k
        self.total_playouts += k
        self.wave_index += 1

    def search(self, n_playout=None):
        """get_move_probs' loop (mcts.py:135-139): n_playout playouts for every game."""
        total = self.n_playout if n_playout is None else int(n_playout)
        done = 0
        while done < total:
            # the first playout after a reset/advance may find an unexpanded root: K descents would all
            # stop there, so that wave collects a single leaf per game
            k = 1 if self.playouts_done == 0 else min(self.K, total - done)
            self.playout_wave(k)
            done += k
        self.drain()
        if self.lazy_expand:
            # a root that was evaluated but never revisited (n_playout = 1) gets its block: its children are the
            # reference's freshly expanded ones
            with torch.cuda.device(self.device):
                _lib.check(self.lib.qz_mcts_extend(C.byref(self.tree), self.c_puct, None, 1, _lib.ptr(self.overflow),
                                                   self._stream()), "qz_mcts_extend")

    def root_stats(self, temp=1e-3, want_q=False):
        """(visits int32 [n,140], probs float64 [n,140], root_visits int32 [n][, q float64 [n,140]])
        -- mcts.py:141-144 scattered by action id."""
        dev = self.device
        self.drain()
        visits = torch.empty((self.n, 140), dtype=torch.int32, device=dev)
        probs = torch.empty((self.n, 140), dtype=torch.float64, device=dev)
        rootn = torch.empty((self.n,), dtype=torch.int32, device=dev)
        q = torch.empty((self.n, 140), dtype=torch.float64, device=dev) if want_q else None
        with torch.cuda.device(dev):
            _lib.check(self.lib.qz_mcts_root_stats(C.byref(self.tree), float(temp), _lib.ptr(visits), _lib.ptr(q),
                                                   _lib.ptr(probs), _lib.ptr(rootn), self._stream()),
                       "qz_mcts_root_stats")
        return (visits, probs, rootn, q) if want_q else (visits, probs, rootn)

    def choose(self, mode=0, temp=1e-3, seed=0, noise_eps=0.25, dir_alpha=0.3):
        """Moves int32 [n] (mcts.py:177-187 / pure_mcts.py:115).  mode 0 first-max visits, 1 sample from probs,
        2 sample from (1-eps)*probs + eps*Dirichlet(alpha)."""
        self.drain()
        moves = torch.empty((self.n,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_mcts_choose(C.byref(self.tree), int(mode), float(temp), float(noise_eps),
                                               float(dir_alpha), int(seed) & ((1 << 64) - 1), _lib.ptr(self.game_id),
                                               _lib.ptr(moves), self._stream()), "qz_mcts_choose")
        return moves

    def advance(self, moves, keep_subtree=True):
        """update_with_move (mcts.py:146-151) for every game and step the root states by `moves`.
        keep_subtree=False (or a negative move) discards the tree (pure_mcts.py:142)."""
        moves = moves.to(device=self.device, dtype=torch.int32).contiguous()
        self.drain()
        with torch.cuda.device(self.device):
            if keep_subtree and len(self.arenas) == 2:
                src, dst = self._structs[self.cur][self.cur_set], self._structs[1 - self.cur][self.cur_set]
                _lib.check(self.lib.qz_mcts_reroot(C.byref(src), C.byref(dst), _lib.ptr(moves), 1, _lib.ptr(self.overflow),
                                                   self._stream()), "qz_mcts_reroot")
                self.cur = 1 - self.cur
            else:
                st = self._stream()
                _lib.check(self.lib.qz_env_step(_lib.ptr(self.arena.root_state), _lib.ptr(moves), None, None, self.n, st),
                           "qz_env_step")
                _lib.check(self.lib.qz_mcts_init(C.byref(self.tree), None, None, st), "qz_mcts_init")
        self.playouts_done = 0

    def overflow_count(self):
        return int(self.overflow.item())

    def node_children(self, game, node):
        """`node._children` of slot `node` of game `game` (mcts.py:19-25) in actions() order, children that never got a
        slot included: list of dicts(action, slot, visits, q, prior, inflight), plus the node's own (visits, q, prior,
        resolved slot).  One tiny kernel + one copy: for inspection and the reference-style TreeNode view."""
        self.drain()
        out_i = torch.zeros(4 + 4 * 140, dtype=torch.int32, device=self.device)
        out_d = torch.zeros(2 + 2 * 140, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_mcts_node_children(C.byref(self.tree), int(game), int(node), _lib.ptr(out_i),
                                                      _lib.ptr(out_d), self._stream()), "qz_mcts_node_children")
        oi, od = out_i.cpu().numpy(), out_d.cpu().numpy()
        kids = [dict(action=int(oi[4 + 4 * r]), slot=int(oi[5 + 4 * r]), visits=int(oi[6 + 4 * r]), inflight=int(oi[7 + 4 * r]),
                     q=float(od[2 + 2 * r]), prior=float(od[3 + 2 * r])) for r in range(int(oi[0]))]
        return kids, dict(visits=int(oi[1]), slot=int(oi[2]), q=float(od[0]), prior=float(od[1]))
