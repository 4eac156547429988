"""Drop-in for the reference's pure_mcts.py (pure_mcts.py:1-148): uniform priors, leaf value from a random
rollout of at most 999 plies, move = first max of root visits, tree discarded after every move.

The tree kernels and the rollout kernel are the same ones `tree.BatchedMCTS` uses for thousands of games;
these classes are the single-game mirror of the reference API.
"""
import numpy as np
import torch

from . import tree as _tree
from .mcts import TreeNode as _TreeNodeView
from .quoridor import Quoridor
from .rollout import rollout as _rollout


class TreeNode(_TreeNodeView):
    """pure_mcts.py:19-56 -- read-only view, priors are the uniform 1/len(children) of pure_mcts.py:13-16."""


def rollout_policy_fn(game):
    """pure_mcts.py:7-10 (kept for API parity; the engine samples uniformly on the device instead)."""
    acts = game.actions()
    action_probs = np.random.rand(len(acts))
    return zip(acts, action_probs)


def policy_value_fn(game):
    """pure_mcts.py:13-16"""
    acts = game.actions()
    action_probs = np.ones(len(acts)) / len(acts)
    return zip(acts, action_probs), 0


class MCTS(object):
    """pure_mcts.py:59-125"""

    def __init__(self, policy_value_fn=policy_value_fn, c_puct=5, n_playout=10000, leaves_per_game=1, seed=0,
                 fix_terminal_sign=False, device=None):
        self._policy = policy_value_fn
        self._c_puct = c_puct
        self._n_playout = n_playout
        self._evaluator = _tree.RolloutEvaluator(seed=seed, limit=1000)
        self._engine = _tree.BatchedMCTS(1, self._evaluator, c_puct=c_puct, n_playout=n_playout,
                                         leaves_per_game=leaves_per_game, fix_terminal_sign=fix_terminal_sign,
                                         reuse_tree=False, device=device, lazy_expand=False)   # `_children` right after a playout
        self._engine.reset(torch.tensor([Quoridor().packed()], dtype=torch.int64))

    @property
    def _root(self):
        self._engine.drain()
        return TreeNode(self._engine, 0, int(self._engine.arena.root[0].item()), None, True)

    def _playout(self, game):
        """pure_mcts.py:66-83: one playout (descent, expansion, random rollout, backup) from `game`."""
        row = torch.tensor([game.packed()], dtype=torch.int64, device=self._engine.device)
        self._engine.root_state.copy_(row)
        self._engine.playout_wave(1)

    def _evaluate_rollout(self, game, limit=1000):
        """pure_mcts.py:86-108: one uniformly random playout of at most limit-1 plies from `game`; +1 / -1 from
        the point of view of the side to move in `game`, 0 if nobody won.  (`game` itself is not advanced.)"""
        row = torch.tensor([game.packed()], dtype=torch.int64, device=self._engine.device)
        self._rollouts_done = getattr(self, "_rollouts_done", 0) + 1
        res, _, _ = _rollout(row, per_state=1, seed=self._evaluator.seed, rid_base=(1 << 40) + self._rollouts_done,
                             limit=limit, return_plies=False)
        return int(res.item())

    def get_move(self, game):
        """pure_mcts.py:110-115"""
        row = torch.tensor([game.packed()], dtype=torch.int64, device=self._engine.device)
        self._engine.root_state.copy_(row)
        self._engine.search(self._n_playout)
        return int(self._engine.choose(mode=0).item())

    def root_visits(self):
        visits, _, _ = self._engine.root_stats(temp=1.0)
        return visits[0].cpu().numpy()

    def update_with_move(self, last_move):
        """pure_mcts.py:117-122 -- only ever called with -1 by MCTSPlayer (:142)."""
        self._engine.advance(torch.tensor([int(last_move)], dtype=torch.int32), keep_subtree=False)

    def __str__(self):
        return "MCTS"


class MCTSPlayer(object):
    """pure_mcts.py:128-148"""

    def __init__(self, c_puct=5, n_playout=50, **engine_kwargs):
        self.mcts = MCTS(policy_value_fn, c_puct, n_playout, **engine_kwargs)

    def set_player_ind(self, p):
        self.player = p

    def reset_player(self):
        self.mcts.update_with_move(-1)

    def choose_action(self, game):
        sensible_moves = game.actions()
        if len(sensible_moves) > 0:
            move = self.mcts.get_move(game)
            self.mcts.update_with_move(-1)
            return move
        return None

    get_action = choose_action

    def __str__(self):
        return "MCTS {}".format(self.player)
