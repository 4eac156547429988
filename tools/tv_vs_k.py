"""Total-variation distance of root visit distributions, K leaves per wave (virtual loss) against K = 1, at the
bench's search size (1000 playouts) under the deterministic stubs S1 (uniform priors, value 0 -- pure MCTS's
priors) and S2 (pseudo-random priors and values).  DESIGN.md section 2 quotes its output.  Usage: python tools/tv_vs_k.py"""
import sys, os, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_quoridor_b200 import tree
from alphazero_quoridor_b200.synthetic import midgame_positions
n = 256
st = midgame_positions(n, seed=21, min_plies=0, max_plies=30)
out = {}
for stub in ("S1", "S2"):
    ref = None
    for K in (1, 8, 32, 64, 128):
        eng = tree.BatchedMCTS(n, tree.StubEvaluator(stub), c_puct=5, n_playout=1000, leaves_per_game=K, reuse_tree=False)
        eng.reset(st)
        eng.search()
        v, _, _ = eng.root_stats(temp=1.0)
        v = v.double()
        p = v / v.sum(1, keepdim=True).clamp(min=1)
        if K == 1:
            ref = p
        else:
            tv = 0.5 * (p - ref).abs().sum(1)
            same = (p.argmax(1) == ref.argmax(1)).double().mean().item()
            out["%s_K%d" % (stub, K)] = {"tv_mean": tv.mean().item(), "tv_max": tv.max().item(), "same_best": same}
print(json.dumps(out, indent=1))
