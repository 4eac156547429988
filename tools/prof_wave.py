#!/usr/bin/env python3
"""One self-play ply of bench.py's headline workload (after a few warm-up plies) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full -k regex:... python tools/prof_wave.py`.  Never a source of timings."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warm", type=int, default=8)
    ap.add_argument("--plies", type=int, default=1)
    ap.add_argument("--az", action="store_true", help="AlphaZero-MCTS wave (stub-free: the bf16 net) instead of pure MCTS")
    a = ap.parse_args()
    args = bench.parse([])
    from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
    from alphazero_quoridor_b200.tree import NetEvaluator, RolloutEvaluator
    if a.az:
        from alphazero_quoridor_b200.policy_value_net import PolicyValueNet
        torch.manual_seed(0)
        net = PolicyValueNet(use_gpu=True)
        net.use_cuda_graph = False                       # ncu profiles kernels, not graph replays
        sp = BatchedSelfPlay(args.az_games, NetEvaluator(net), c_puct=args.c_puct, n_playout=100, leaves_per_game=args.az_leaves,
                             temp=1.0, pure=False, seed=args.seed)
    else:
        sp = BatchedSelfPlay(args.games, RolloutEvaluator(seed=args.seed, limit=1000), c_puct=args.c_puct,
                             n_playout=args.playouts, leaves_per_game=args.leaves, pure=True, seed=args.seed,
                             defer_until_drain=True)
    for _ in range(a.warm):
        sp.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(a.plies):
        sp.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done", sp.mcts.counters())


if __name__ == "__main__":
    main()
