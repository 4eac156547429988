"""Drop-in for the reference's mcts.py: `softmax`, `MCTS`, `MCTSPlayer` -- same constructors, methods, return
types and defaults (mcts.py:6-199) -- running the tree on the GPU through `tree.BatchedMCTS` (batch of one).

`policy_value_fn(game)` keeps the reference's contract (policy_value_net.py:145-164): it receives a
`Quoridor`-like object positioned at the leaf and returns `(iterable[(action, prob)], value)`.
Three kinds of callables are accepted:
  * any Python callable  -> called once per leaf on a host `Quoridor` view of the leaf (API-compatible, slow);
  * `DeviceStub(kind)`   -> the parity stubs S1/S2/S3 evaluated by a kernel (no host round trip per leaf);
  * `PolicyValueNet.policy_value_fn` of this package -> the net's batched device path.
For throughput use `tree.BatchedMCTS` / `selfplay.BatchedSelfPlay` with thousands of games instead.
"""
import numpy as np
import torch

from . import tree as _tree
from .quoridor import Quoridor, mask_to_actions, unpack_meta

M64 = (1 << 64) - 1


def softmax(x):
    """mcts.py:6-9"""
    probs = np.exp(x - np.max(x))
    probs /= np.sum(probs)
    return probs


class DeviceStub:
    """A policy_value_fn that the engine evaluates on the device (tests/golden/stubs.py S1/S2/S3).
    Calling it from Python raises: it only exists to select the kernel."""

    def __init__(self, kind):
        self.kind = kind

    def __call__(self, game):
        raise RuntimeError("DeviceStub is evaluated by the qz_stub_eval kernel, not called from Python")


def _game_from_state(row):
    """Host `Quoridor` view of a qz_state row (for Python policy callbacks)."""
    H, V, m = [int(x) & M64 for x in row]
    d = unpack_meta(m)
    g = Quoridor()
    for ix in range(64):
        g._intersections[ix] = 1 if (H >> ix) & 1 else (-1 if (V >> ix) & 1 else 0)
    g._positions = {1: d["p1"], 2: d["p2"]}
    g._player1_walls_remaining, g._player2_walls_remaining = d["w1"], d["w2"]
    g.current_player = d["cur"]
    g.last_player = 3 - d["cur"]
    g._ply = d["ply"]
    return g


class _CallbackEvaluator:
    """Evaluates leaves by calling the user's Python policy_value_fn (mcts.py:117)."""

    def __init__(self, fn):
        self.fn = fn

    def evaluate(self, mcts, lset, leaf_rids):
        leaf_states = lset.leaf_state
        m = leaf_states.shape[0]
        rows = leaf_states.cpu().numpy()
        flags = lset.leaf_flags.cpu().numpy()
        priors = np.zeros((m, 140), dtype=np.float32)
        values = np.zeros((m,), dtype=np.float64)
        for i in range(m):
            if flags[i] & (_tree.LEAF_INACTIVE | _tree.LEAF_TERMINAL):
                continue        # reference calls the policy on terminal leaves too but discards the result
            act_probs, v = self.fn(_game_from_state(rows[i]))
            for a, p in act_probs:
                priors[i, int(a)] = np.float32(p)
            values[i] = float(v)
        dev = leaf_states.device
        return dict(priors=torch.from_numpy(priors).to(dev), value_f64=torch.from_numpy(values).to(dev))


def _make_evaluator(policy_value_fn):
    if isinstance(policy_value_fn, DeviceStub):
        return _tree.StubEvaluator(policy_value_fn.kind)
    owner = getattr(policy_value_fn, "__self__", None)
    if owner is not None and hasattr(owner, "evaluate_states") and getattr(policy_value_fn, "__name__", "") == "policy_value_fn":
        return _tree.NetEvaluator(owner)
    return _CallbackEvaluator(policy_value_fn)


class TreeNode(object):
    """Read-only view of one node of the device tree with the reference's TreeNode surface (mcts.py:12-80):
    `_n_visits`, `_Q`, `_P`, `_u`, `_children` (dict action -> TreeNode, in insertion order), `_parent`,
    `is_leaf()`, `is_root()`, `get_value(c_puct)`.  The statistics live in the flat device arrays of
    tree.BatchedMCTS; this object copies what it is asked for (qz_mcts_node_children).  The device tree gives a child
    its slot when it is first visited; a child without one is the reference's freshly expanded child (0 visits, Q 0, its
    prior, no children) and is shown as such.  Mutation (expand/update/select) happens in the kernels, so those methods
    are not offered here."""

    def __init__(self, engine, game_index, node, parent=None, uniform_prior=False, info=None):
        self._engine, self._g, self._node, self._parent = engine, game_index, int(node), parent
        self._uniform = uniform_prior
        self._info = info                  # dict(visits, q, prior) for a child read with its parent; None = read on demand
        self._kids = None

    def _read(self):
        if self._kids is None and self._node >= 0:
            kids, me = self._engine.node_children(self._g, self._node)
            self._kids = kids
            if self._info is None:
                self._info = me
        return self._info or dict(visits=0, q=0.0, prior=0.0)

    @property
    def _n_visits(self):
        return int((self._info or self._read())["visits"])

    @property
    def _Q(self):
        return float((self._info or self._read())["q"])

    @property
    def _P(self):
        if self._uniform and self._parent is not None:
            return 1.0 / len(self._parent._children)
        return float(np.float32((self._info or self._read())["prior"]))

    @property
    def _children(self):
        if self._node < 0:
            return {}
        self._read()
        return {k["action"]: TreeNode(self._engine, self._g, k["slot"], self, self._uniform,
                                      info=dict(visits=k["visits"], q=k["q"], prior=k["prior"])) for k in self._kids}

    @property
    def _u(self):
        return self.get_value(self._engine.c_puct) - self._Q if self._parent is not None else 0

    def get_value(self, c_puct):
        """mcts.py:64-70 with numpy's promotion (float32 prior * weak scalar rounds to float32 first)."""
        if self._uniform:
            cp = c_puct * self._P
        else:
            cp = float(np.float32(c_puct) * np.float32((self._info or self._read())["prior"]))
        return self._Q + cp * np.sqrt(self._parent._n_visits) / (1 + self._n_visits)

    def is_leaf(self):
        return len(self._children) == 0

    def is_root(self):
        return self._parent is None


class MCTS(object):
    """mcts.py:83-154"""

    def __init__(self, policy_value_fn, c_puct=5, n_playout=1800, leaves_per_game=1, fix_terminal_sign=False,
                 device=None):
        self._policy = policy_value_fn
        self._c_puct = c_puct
        self._n_playout = n_playout
        self._engine = _tree.BatchedMCTS(1, _make_evaluator(policy_value_fn), c_puct=c_puct, n_playout=n_playout,
                                         leaves_per_game=leaves_per_game, fix_terminal_sign=fix_terminal_sign,
                                         reuse_tree=True, device=device)
        self._fresh = True
        self._engine.reset(torch.tensor([Quoridor().packed()], dtype=torch.int64))

    def _set_root_state(self, game):
        row = torch.tensor([game.packed()], dtype=torch.int64, device=self._engine.device)
        self._engine.root_state.copy_(row)

    @property
    def _root(self):
        """The root as a reference-style TreeNode view (mcts.py:97)."""
        self._engine.drain()
        return TreeNode(self._engine, 0, int(self._engine.arena.root[0].item()),
                        uniform_prior=self._engine.uniform_prior)

    def _playout(self, game):
        """mcts.py:103-127: ONE playout from `game` (the reference mutates the game copy it is given; here the
        descent replays the moves on the device and `game` is left untouched)."""
        self._set_root_state(game)
        self._engine.playout_wave(1)

    def get_move_probs(self, game, temp=1e-3):
        """mcts.py:129-144: n_playout playouts from `game`, then (acts, probs) over the root's children in
        insertion (= actions()) order."""
        self._set_root_state(game)
        self._engine.search(self._n_playout)
        visits140, _, _ = self._engine.root_stats(temp=max(float(temp), 1e-12))
        acts, visits = self._root_children(visits140)
        act_probs = softmax(1.0 / temp * np.log(np.array(visits) + 1e-10))
        return tuple(acts), act_probs

    def _root_children(self, visits140=None):
        """(acts in child order, visits) of the root."""
        kids, _ = self._engine.node_children(0, int(self._engine.arena.root[0].item()))
        return [k["action"] for k in kids], [k["visits"] for k in kids]

    def root_children_stats(self):
        """(acts, visits, Q) of the root's children plus (root visits, root Q) -- for tests / inspection."""
        kids, me = self._engine.node_children(0, int(self._engine.arena.root[0].item()))
        return ([k["action"] for k in kids], [k["visits"] for k in kids], [k["q"] for k in kids], me["visits"], me["q"])

    def update_with_move(self, last_move):
        """mcts.py:146-151"""
        mv = torch.tensor([int(last_move)], dtype=torch.int32)
        # the root state is re-read from the caller's game at the next get_move_probs, as in the reference
        self._engine.advance(mv, keep_subtree=True)

    def __str__(self):
        return "MCTS"


class MCTSPlayer(object):
    """mcts.py:157-199"""

    def __init__(self, policy_value_function, c_puct=5, n_playout=2000, is_selfplay=0, **engine_kwargs):
        self.mcts = MCTS(policy_value_function, c_puct, n_playout, **engine_kwargs)
        self._is_selfplay = is_selfplay

    def set_player_ind(self, p):
        self.player = p

    def reset_player(self):
        self.mcts.update_with_move(-1)

    def choose_action(self, game, temp=1e-3, return_prob=0):
        sensible_moves = game.actions()
        move_probs = np.zeros(140)
        if len(sensible_moves) > 0:
            acts, probs = self.mcts.get_move_probs(game, temp)
            move_probs[list(acts)] = probs
            if self._is_selfplay:
                # mcts.py:181: Dirichlet noise on the MOVE distribution (not on the root priors)
                move = np.random.choice(acts, p=0.75 * probs + 0.25 * np.random.dirichlet(0.3 * np.ones(len(probs))))
                self.mcts.update_with_move(move)
            else:
                move = np.random.choice(acts, p=probs)
                self.mcts.update_with_move(-1)
            if return_prob:
                return move, move_probs
            return move
        # mcts.py:195-196: the reference prints a warning and returns None
        return None

    # BASELINE.json's north_star calls this `get_action`; the reference method is choose_action (mcts.py:172)
    get_action = choose_action

    def __str__(self):
        return "MCTS {}".format(self.player)
