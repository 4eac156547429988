import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_quoridor_b200 import tree
from alphazero_quoridor_b200.selfplay import StreamedSelfPlay
for streams in (1, 2, 4):
    sp = StreamedSelfPlay(4096, lambda: tree.RolloutEvaluator(seed=1), n_streams=streams, n_playout=1000, c_puct=5.0,
                          leaves_per_game=64, pure=True, seed=1, defer_depth=4)
    for _ in range(3):
        sp.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        sp.step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("streams %d: enqueue %.1f ms/step, total %.1f ms/step" % (streams, (t1 - t0) / 3e-3, (t2 - t0) / 3e-3), flush=True)
    del sp
    torch.cuda.empty_cache()
