"""Debug: why do the e2e replay's final positions differ from the timed run's? (bench.pure_section logic, small sizes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
from alphazero_quoridor_b200.tree import RolloutEvaluator
from alphazero_quoridor_b200.quoridor import unpack_meta
n, steps = 512, 6
ev = RolloutEvaluator(seed=7, limit=1000)
sp = BatchedSelfPlay(n, ev, c_puct=5, n_playout=200, leaves_per_game=16, pure=True, seed=7, defer_until_drain=True)
m = sp.mcts
for _ in range(3):
    sp.step()
snap = dict(root=m.root_state.clone(), started=sp.games_started.clone(), total=m.total_playouts, wave=m.wave_index)
mv_a = [sp.step().clone() for _ in range(steps)]
end_root = m.root_state.clone()
m.reset(snap["root"]); sp.games_started.copy_(snap["started"]); sp._set_game_ids()
m.total_playouts, m.wave_index = snap["total"], snap["wave"]
host = snap["root"].cpu().pin_memory()
mv_b = []
for _ in range(steps):
    m.reset(host.to("cuda", non_blocking=True))
    m.search()
    moves = m.choose(mode=0)
    m.advance(moves, keep_subtree=False)
    host.copy_(m.root_state); torch.cuda.synchronize()
    mv_b.append(moves.clone())
for i, (a, b) in enumerate(zip(mv_a, mv_b)):
    print("step", i, "moves differ in", int((a != b).sum()), "games")
d = (host.cuda() != end_root)
print("rows differing:", int(d.any(1).sum()), "per column:", d.sum(0).tolist())
if d.any():
    g = int(d.any(1).nonzero()[0])
    print(unpack_meta(int(host[g, 2])), unpack_meta(int(end_root[g, 2].item())))
