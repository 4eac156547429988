#!/usr/bin/env python3
"""Write oracle/bench_positions.npz: the root positions bench.py's headline workload holds at every self-play ply.

Run on the GPU box (gpurun) with bench.py's default seed / games / playouts / leaves:

    python tools/gen_bench_positions.py [--plies 30] [--keep 64]

The engine is seeded and deterministic, so ply p of this run IS ply p of `python bench.py` (rank 0): the CPU legs of
bench.py (`cpu_baseline`, `--impl reference`) then search exactly the positions the GPU arm searches at the same ply.
states[p, g] = qz_state row (int64 x 3) of game g BEFORE its move at ply p.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plies", type=int, default=30)
    ap.add_argument("--keep", type=int, default=64)
    a = ap.parse_args()
    args = bench.parse([])
    from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
    from alphazero_quoridor_b200.tree import RolloutEvaluator
    sp = BatchedSelfPlay(args.games, RolloutEvaluator(seed=args.seed, limit=1000), c_puct=args.c_puct,
                         n_playout=args.playouts, leaves_per_game=args.leaves, pure=True, seed=args.seed, game_id_base=0,
                         defer_depth=max(args.defer, 0), defer_until_drain=args.defer < 0)
    rows = []
    for p in range(a.plies):
        rows.append(sp.mcts.root_state[:a.keep].cpu().numpy().copy())
        sp.step()
    torch.cuda.synchronize()
    out = os.path.join(ROOT, "oracle", "bench_positions.npz")
    np.savez_compressed(out, states=np.stack(rows), seed=args.seed, games=args.games, playouts=args.playouts,
                        leaves=args.leaves)
    st = np.stack(rows)
    walls = ((st[:, :, 2] >> 16) & 0xFF) + ((st[:, :, 2] >> 24) & 0xFF)
    print("wrote %s: %s; walls left at ply 0/10/20/last: %s" % (out, st.shape, [float(walls[i].mean()) for i in (0, min(10, a.plies - 1), min(20, a.plies - 1), -1)]))


if __name__ == "__main__":
    main()
