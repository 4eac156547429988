#!/usr/bin/env python3
"""Developer micro-benchmarks of the individual kernels (CUDA events, warm-up, L2-sized inputs).
Not the contract benchmark -- see bench.py for that."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_quoridor_b200.quoridor import BatchedQuoridor  # noqa: E402
from alphazero_quoridor_b200.rollout import rollout  # noqa: E402
from alphazero_quoridor_b200.synthetic import midgame_positions  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="sweep,rollout,encode,step")
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--rollouts", type=int, default=1 << 18)
    args = ap.parse_args()
    which = args.which.split(",")
    out = {}
    if "sweep" in which:
        states = midgame_positions(args.n, seed=7)
        env = BatchedQuoridor(args.n, states=states)
        mask = torch.empty((args.n, 3), dtype=torch.int64, device="cuda")
        ms = timed(lambda: env.legal_mask(out=mask))
        out["sweep"] = {"n": args.n, "ms": ms, "sweeps_per_s": args.n / ms * 1e3}
    if "rollout" in which:
        start = BatchedQuoridor(1).states
        res = {}

        def run():
            res["r"] = rollout(start, per_state=args.rollouts, seed=1, limit=1000)
        ms = timed(run, iters=3, warm=1)
        plies = res["r"][1].double()
        out["rollout_from_start"] = {"n": args.rollouts, "ms": ms, "mean_plies": plies.mean().item(),
                                     "max_plies": plies.max().item(),
                                     "env_steps_per_s": plies.sum().item() / ms * 1e3,
                                     "rollouts_per_s": args.rollouts / ms * 1e3,
                                     "p1_win": (res["r"][0] == 1).double().mean().item()}
        late = midgame_positions(4096, seed=3, min_plies=30, max_plies=40)

        def run2():
            res["l"] = rollout(late, per_state=64, seed=2, limit=1000)
        ms = timed(run2, iters=3, warm=1)
        plies = res["l"][1].double()
        out["rollout_pawn_phase"] = {"n": 4096 * 64, "ms": ms, "mean_plies": plies.mean().item(),
                                     "env_steps_per_s": plies.sum().item() / ms * 1e3}
    if "encode" in which:
        n = 1 << 18
        env = BatchedQuoridor(n, states=midgame_positions(n, seed=9))
        for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
            buf = torch.empty((n, 26, 9, 9), dtype=dt, device="cuda")
            ms = timed(lambda: env.encode(out=buf))
            nbytes = buf.numel() * buf.element_size() + n * 24
            out["encode_" + name] = {"n": n, "ms": ms, "GBps": nbytes / ms / 1e6}
        buf = torch.empty((n, 32, 9, 9), dtype=torch.bfloat16, device="cuda", memory_format=torch.channels_last)
        ms = timed(lambda: env.encode(out=buf, channels_last=True, c_stride=32))
        out["encode_bf16_nhwc32"] = {"n": n, "ms": ms, "GBps": (buf.numel() * 2 + n * 24) / ms / 1e6}
    if "step" in which:
        n = 1 << 24
        env = BatchedQuoridor(n)
        acts = torch.zeros(n, dtype=torch.int32, device="cuda")
        done = torch.empty(n, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: env.step(acts, done=done))
        out["step"] = {"n": n, "ms": ms, "GBps": n * (24 + 4 + 24 + 1) / ms / 1e6, "steps_per_s": n / ms * 1e3}
    if "az" in which:
        from alphazero_quoridor_b200.policy_value_net import PolicyValueNet
        from alphazero_quoridor_b200 import tree
        torch.manual_seed(0)
        net = PolicyValueNet(use_gpu=True)
        for n, K, npl in ((8192, 1, 24), (8192, 4, 48), (8192, 8, 96)):
            states = midgame_positions(n, seed=11, min_plies=0, max_plies=40)
            eng = tree.BatchedMCTS(n, tree.NetEvaluator(net), c_puct=5, n_playout=npl, leaves_per_game=K,
                                   reuse_tree=False)
            eng.reset(states)
            eng.search(8)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.search(npl)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            out["az_mcts_K%d" % K] = {"games": n, "playouts": npl, "ms": ms, "sims_per_s": n * npl / ms * 1e3}
        for m in (8192, 65536):
            st = midgame_positions(m, seed=5)
            net.evaluate_states(st)
            ms = timed(lambda: net.evaluate_states(st))
            out["net_forward_%d" % m] = {"ms": ms, "positions_per_s": m / ms * 1e3,
                                         "TFLOPs": m * 62.8e6 / ms / 1e9}
    if "cfg4" in which:
        # BASELINE configs[3], one GPU's shard: 8192 games, n_playout = 800, tree reuse, Dirichlet-mixed moves
        from alphazero_quoridor_b200.policy_value_net import PolicyValueNet
        from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
        from alphazero_quoridor_b200.tree import NetEvaluator
        torch.manual_seed(0)
        net = PolicyValueNet(use_gpu=True)
        sp = BatchedSelfPlay(8192, NetEvaluator(net), c_puct=5, n_playout=800, leaves_per_game=8, temp=1.0, pure=False, seed=1)
        sp.step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            sp.step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        out["cfg4_selfplay_800"] = {"games": 8192, "playouts": 800, "ms_per_move": ms / 2,
                                    "sims_per_s": 2 * 8192 * 800 / ms * 1e3, "tree_GB": sp.mcts.nbytes() / 1e9,
                                    "overflow": sp.mcts.overflow_count(),
                                    "max_nodes_used": int(sp.mcts.arena.n_nodes.max().item()), "node_cap": sp.mcts.node_cap,
                                    "mem_allocated_GB": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
