"""Debug: why do the e2e replay's final positions differ from the timed run's? (bench.pure_section logic, small sizes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
from alphazero_quoridor_b200.tree import RolloutEvaluator
from alphazero_quoridor_b200.quoridor import unpack_meta
import sys as _s
n, steps = (int(_s.argv[1]), int(_s.argv[2])) if len(_s.argv) > 2 else (512, 6)
NPL, KK = (1000, 64) if n >= 4096 else (200, 16)
ev = RolloutEvaluator(seed=7, limit=1000)
sp = BatchedSelfPlay(n, ev, c_puct=5, n_playout=NPL, leaves_per_game=KK, pure=True, seed=7, defer_until_drain=True)
m = sp.mcts
for _ in range(5 if n >= 4096 else 3):
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
    done = ((host[:, 2] >> 40) & 1).bool()
    if bool(done.any()):
        print("finished games at this step:", int(done.sum()))
        host[done] = sp._start.cpu()[0]
        sp.games_started += done.cuda().to(torch.int64)
        sp._set_game_ids()
    mv_b.append(moves.clone())
for i, (a, b) in enumerate(zip(mv_a, mv_b)):
    print("step", i, "moves differ in", int((a != b).sum()), "games")
d = (host.cuda() != end_root)
print("rows differing:", int(d.any(1).sum()), "per column:", d.sum(0).tolist())
if d.any():
    for g in d.any(1).nonzero().flatten().tolist()[:5]:
        print(g, unpack_meta(int(host[g, 2])), unpack_meta(int(end_root[g, 2].item())), "started", int(sp.games_started[g]))
print("overflow", m.overflow_count(), "finished in value run", int(sp.finished_games.item()), "stalemated", int(sp.stalemated_games.item()))
