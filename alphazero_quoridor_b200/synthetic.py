"""Synthetic positions for benchmarks and tests (BASELINE.json configs 1/5).

Positions are produced ON THE GPU by truncated uniform-random legal play from `reset()` with the rollout
kernel, so they are reachable, legal and reproducible from (seed, index) alone.  Uniform play places a wall
on ~97% of the plies while walls remain, so k plies give ~k walls.
"""
import torch

from .quoridor import BatchedQuoridor
from .rollout import rollout


def midgame_positions(n, seed=7, min_plies=10, max_plies=20, device=None, rid_base=0):
    """n live positions with `min_plies..max_plies` random plies played (uniformly many of each length).
    Returns int64 [n,3] qz_state rows; finished games (rare) are replaced by the start position."""
    env = BatchedQuoridor(1, device=device)
    start = env.states
    out = []
    lengths = list(range(min_plies, max_plies + 1))
    per = (n + len(lengths) - 1) // len(lengths)
    made = 0
    for i, k in enumerate(lengths):
        cnt = min(per, n - made)
        if cnt <= 0:
            break
        _, _, final = rollout(start, per_state=cnt, seed=seed, rid_base=rid_base + made, limit=k + 1,
                              return_plies=False, return_final=True)
        out.append(final)
        made += cnt
    states = torch.cat(out, 0)
    done = ((states[:, 2] >> 40) & 1).bool()
    if done.any():
        states[done] = start[0]
    # clear the ply counter / flags so the positions look like fresh inputs
    return states.contiguous()
