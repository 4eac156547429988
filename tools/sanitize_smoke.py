#!/usr/bin/env python3
"""Small run through every kernel for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_quoridor_b200 import tree  # noqa: E402
from alphazero_quoridor_b200.quoridor import BatchedQuoridor  # noqa: E402
from alphazero_quoridor_b200.rollout import rollout  # noqa: E402
from alphazero_quoridor_b200.selfplay import BatchedSelfPlay  # noqa: E402
from alphazero_quoridor_b200.synthetic import midgame_positions  # noqa: E402

n = 203
st = midgame_positions(n, seed=3, min_plies=0, max_plies=70)
env = BatchedQuoridor(n, states=st.clone())
m = env.legal_mask()
env.sample_actions(m, seed=1)
env.step(env.sample_actions(m, seed=1), legal_mask=m)
for dt in (torch.float32, torch.bfloat16):
    env.encode(dtype=dt)
    env.encode(dtype=dt, channels_last=True, c_stride=32)
    env.encode(dtype=dt, channels_last=True, c_stride=26)
BatchedQuoridor(37).random_play(seed=2, max_plies=400)
BatchedQuoridor(37).random_play(seed=2, max_plies=40, fused=False)
rollout(st, per_state=3, seed=5, return_final=True)
late = midgame_positions(4000, seed=5, min_plies=28, max_plies=60)
meta = late[:, 2]
sel = ((((meta >> 16) & 0xFF) + ((meta >> 24) & 0xFF)) > 0) & (((meta >> 40) & 1) == 0)
rollout(late[sel][:64].contiguous(), per_state=2, seed=6)                      # stuck rollouts
for ev, K, reuse, defer in ((tree.StubEvaluator("S3"), 1, True, 0), (tree.StubEvaluator("S2"), 4, True, 0),
                            (tree.RolloutEvaluator(seed=1), 4, False, 3)):
    eng = tree.BatchedMCTS(n, ev, c_puct=5, n_playout=24, leaves_per_game=K, reuse_tree=reuse, defer_depth=defer)
    eng.reset(st)
    eng.search()
    eng.root_stats(temp=1.0, want_q=True)
    for mode in (0, 1, 2):
        mv = eng.choose(mode=mode, temp=1.0, seed=3)
    eng.advance(mv, keep_subtree=reuse)
    eng.search(8)
# round 2: lazy child slots with block growth + forwarding, lazy expansion (flagged sweep + extend), end-of-search stuck
# passes, a one-playout search (root block built at the end), the node view, re-rooting with the prior pool
eng = tree.BatchedMCTS(n, tree.RolloutEvaluator(seed=2), c_puct=5, n_playout=160, leaves_per_game=8, reuse_tree=False,
                       defer_until_drain=True)
eng.reset(st)
eng.search()
eng.choose(mode=0)
eng.node_children(0, 0)
one = tree.BatchedMCTS(n, tree.RolloutEvaluator(seed=3), c_puct=5, n_playout=1, leaves_per_game=1, reuse_tree=False)
one.reset(st)
one.search()
one.root_stats(temp=1.0)
deep = tree.BatchedMCTS(16, tree.StubEvaluator("S2"), c_puct=5, n_playout=300, leaves_per_game=1, reuse_tree=True)
deep.reset(late[sel][:16].contiguous())
for _ in range(3):
    deep.search()
    deep.advance(deep.choose(mode=0), keep_subtree=True)
deep.check_device()
sp = BatchedSelfPlay(64, tree.StubEvaluator("S3"), n_playout=16, leaves_per_game=2, record=True, fix_terminal_sign=True,
                     max_plies=60)
for _ in range(12):
    sp.step()
torch.cuda.synchronize()
print("sanitize smoke done")
