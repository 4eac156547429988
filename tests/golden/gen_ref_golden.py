#!/usr/bin/env python3
"""Fixtures from the UNMODIFIED reference net, trainer and rollouts (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_ref_golden.py [--jobs 8] [--only net|train|rollout]

Imports /root/reference/{policy_value_net,quoridor,pure_mcts}.py (read-only) and records
  * ref_net.pth           the reference's own `PolicyValueNet(use_gpu=False).save_model(...)` file
                          (policy_value_net.py:198-200) after a few training-mode forwards, so the BatchNorm running
                          statistics are not the trivial (0, 1)
  * ref_net_io.npz        256 positions taken from replay_traces.json.gz (as H, V, p1, p2, w1, w2, cur) with the
                          reference's outputs on their `state()` tensors: eval-mode batch forward (what batched
                          inference must match) and training-mode batch-1 forwards (what the reference's own search sees,
                          policy_value_net.py:117-120,154 -- it never calls .eval())
  * ref_train_step.npz    two consecutive reference `train_step`s (policy_value_net.py:166-192) from ref_net.pth on a
                          fixed minibatch: loss / entropy read out of the reference's own frame (its `loss.data[0]`
                          raises IndexError on current PyTorch AFTER optimizer.step()), per-tensor sums of the updated
                          weights and full copies of the small tensors
  * ref_rollouts.json     value and length of `pure_mcts.MCTS._evaluate_rollout` (pure_mcts.py:86-108) under
                          np.random.seed, from three positions -- the distributions the rollout kernels must follow
Nothing here is imported by the product; tests read only the emitted files.
"""
import argparse
import contextlib
import gzip
import io
import json
import multiprocessing as mp
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

SEED = 20261017


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def set_position(game, pos):
    game._positions = {1: pos["p1"], 2: pos["p2"]}
    for ix in range(64):
        game._intersections[ix] = 1 if (pos["H"] >> ix) & 1 else (-1 if (pos["V"] >> ix) & 1 else 0)
    game._player1_walls_remaining, game._player2_walls_remaining = pos["w1"], pos["w2"]
    game.current_player = pos["cur"]
    game.last_player = 3 - pos["cur"]
    return game


def golden_positions(n):
    traces = json.loads(gzip.open(os.path.join(HERE, "replay_traces.json.gz")).read())
    rng = np.random.RandomState(SEED)
    pool = [p for t in traces for p in t["plies"] if not p["done"] and 0 <= p["p1"] <= 80 and 0 <= p["p2"] <= 80]
    idx = rng.choice(len(pool), size=n, replace=False)
    return [{k: pool[i][k] for k in ("H", "V", "p1", "p2", "w1", "w2", "cur")} for i in idx]


def states_of(positions):
    import quoridor
    g = quoridor.Quoridor()
    out = np.zeros((len(positions), 26, 9, 9), dtype=np.float64)
    for i, p in enumerate(positions):
        out[i] = set_position(g, p).state()
    return out


def gen_net():
    import torch
    import policy_value_net as ref
    torch.manual_seed(SEED)
    torch.set_num_threads(1)
    positions = golden_positions(256)
    x = states_of(positions)
    net = ref.PolicyValueNet(use_gpu=False)
    assert net.policy_value_net.training                       # the reference never leaves training mode
    with torch.no_grad():
        for b in range(8):                                     # move the running statistics away from (0, 1)
            net.policy_value_net(torch.from_numpy(x[32 * b:32 * b + 32]).float())
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.makedirs(os.path.join(tmp, "ckpt"))
    os.chdir(tmp)
    try:
        net.save_model("ref_net")                              # policy_value_net.py:198-200
    finally:
        os.chdir(cwd)
    shutil.copy(os.path.join(tmp, "ckpt", "ref_net.pth"), os.path.join(HERE, "ref_net.pth"))
    sd = {k: v.clone() for k, v in net.policy_value_net.state_dict().items()}
    # training-mode batch-1 forwards: exactly what policy_value_fn computes during the reference's search
    train_p = np.zeros((256, 140), dtype=np.float32)
    train_v = np.zeros((256,), dtype=np.float32)
    with torch.no_grad():
        for i in range(256):
            logp, v = net.policy_value_net(torch.from_numpy(x[i:i + 1]).float())
            train_p[i] = np.exp(logp.numpy()[0])
            train_v[i] = v.numpy()[0, 0]
    net.policy_value_net.load_state_dict(sd)                   # undo the running-statistics updates
    net.policy_value_net.eval()
    eval_p, eval_v = net.policy_value(x)                       # policy_value_net.py:127-143 on the eval-mode module
    net.policy_value_net.train()
    cols = {k: np.array([p[k] for p in positions], dtype=np.uint64 if k in ("H", "V") else np.int32)
            for k in ("H", "V", "p1", "p2", "w1", "w2", "cur")}
    np.savez_compressed(os.path.join(HERE, "ref_net_io.npz"), eval_probs=eval_p.astype(np.float32),
                        eval_value=eval_v.reshape(-1).astype(np.float32), train_probs=train_p, train_value=train_v, **cols)
    dev = np.abs(train_p - eval_p).max(1)
    print("net: 256 positions; eval-vs-train(batch 1) max|dprob| mean %.4f max %.4f; |dvalue| mean %.4f"
          % (dev.mean(), dev.max(), np.abs(train_v - eval_v.reshape(-1)).mean()))


def minibatch(positions):
    rng = np.random.RandomState(SEED + 1)
    probs = rng.dirichlet(np.ones(140) * 0.3, size=len(positions)).astype(np.float32)
    z = rng.choice([-1.0, 1.0], size=len(positions)).astype(np.float32)
    return probs, z


def gen_train():
    import torch
    import policy_value_net as ref
    torch.set_num_threads(1)
    positions = golden_positions(256)[:32]
    x = states_of(positions)
    probs, z = minibatch(positions)
    net = ref.PolicyValueNet(use_gpu=False)
    net.policy_value_net.load_state_dict(torch.load(os.path.join(HERE, "ref_net.pth")))
    out = {}
    for step in range(2):
        try:
            net.train_step(x, probs, z, 2e-3)
            raise SystemExit("reference train_step returned: torch changed, revisit this script")
        except IndexError as e:                                # loss.data[0] on a 0-dim tensor, AFTER optimizer.step()
            tb = e.__traceback__
            while tb.tb_next is not None:
                tb = tb.tb_next
            loc = tb.tb_frame.f_locals
            out["loss%d" % step] = np.float64(loc["loss"].item())
            out["entropy%d" % step] = np.float64(loc["entropy"].item())
            out["value_loss%d" % step] = np.float64(loc["value_loss"].item())
        sd = net.policy_value_net.state_dict()
        out["sums%d" % step] = np.array([sd[k].double().sum().item() for k in sorted(sd)], dtype=np.float64)
        out["abssums%d" % step] = np.array([sd[k].double().abs().sum().item() for k in sorted(sd)], dtype=np.float64)
        for k in ("fc2.weight", "fc2.bias", "fc3.bias", "bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var",
                  "conv3.weight", "res5.bn2.weight", "conv1.weight"):
            out["w%d_%s" % (step, k)] = sd[k].numpy().copy()
    out["keys"] = np.array(sorted(sd))
    np.savez_compressed(os.path.join(HERE, "ref_train_step.npz"), **out)
    print("train: loss %.6f -> %.6f, entropy %.6f -> %.6f" % (out["loss0"], out["loss1"], out["entropy0"], out["entropy1"]))


ROLLOUT_POSITIONS = [
    # name, position, rollouts
    ("start", dict(H=0, V=0, p1=4, p2=76, w1=10, w2=10, cur=1), 384),
    ("mid_few_walls_left", None, 768),            # filled from a golden trace below
    ("pawn_only_p1_ahead", dict(H=0x0000001800000000, V=0x0000000000240000, p1=49, p2=58, w1=0, w2=0, cur=2), 4096),
]


def _rollout_job(args):
    name, pos, seed, n = args
    import pure_mcts
    import quoridor
    np.random.seed(seed)
    tree = pure_mcts.MCTS(pure_mcts.policy_value_fn, 5, 1)
    res = []
    for _ in range(n):
        g = set_position(quoridor.Quoridor(), pos)
        count = [0]
        orig = g.step

        def counted(a, _orig=orig, _c=count):
            _c[0] += 1
            return _orig(a)
        g.step = counted
        with quiet():
            try:
                v = tree._evaluate_rollout(g)
            except ValueError:          # stalemate: actions() == [] and max() of nothing raises (reference undefined;
                v = 2                   # the engine flags the position and scores the rollout 0) -- recorded as 2
        res.append((int(v), count[0]))
    return name, res


def gen_rollouts(jobs):
    traces = json.loads(gzip.open(os.path.join(HERE, "replay_traces.json.gz")).read())
    mid = None
    for t in traces:
        for p in t["plies"]:
            if (not p["done"] and 3 <= p["w1"] + p["w2"] <= 5 and p["w1"] >= 1 and p["w2"] >= 1
                    and 9 <= p["p1"] <= 71 and 9 <= p["p2"] <= 71):
                mid = {k: p[k] for k in ("H", "V", "p1", "p2", "w1", "w2", "cur")}
                break
        if mid:
            break
    assert mid is not None
    plan = []
    for name, pos, n in ROLLOUT_POSITIONS:
        pos = pos or mid
        per = max(1, n // (jobs * 2))
        k = 0
        while k < n:
            m = min(per, n - k)
            plan.append((name, pos, 7000 + len(plan), m))
            k += m
    with mp.Pool(jobs) as pool:
        results = pool.map(_rollout_job, plan, chunksize=1)
    out = {}
    for name, pos, n in ROLLOUT_POSITIONS:
        rs = [r for nm, res in results if nm == name for r in res]
        out[name] = {"position": pos or mid, "values": [r[0] for r in rs], "plies": [r[1] for r in rs]}
        v = np.array(out[name]["values"])
        print("rollouts %-20s n=%d  P(+1)=%.3f P(-1)=%.3f P(0)=%.3f  mean plies %.1f"
              % (name, len(rs), (v == 1).mean(), (v == -1).mean(), (v == 0).mean(), np.mean(out[name]["plies"])))
    with open(os.path.join(HERE, "ref_rollouts.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.only in ("", "net"):
        gen_net()
    if a.only in ("", "train"):
        gen_train()
    if a.only in ("", "rollout"):
        gen_rollouts(a.jobs)
