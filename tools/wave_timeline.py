"""Timeline of every C-ABI call of one pure-MCTS self-play step (CUDA events on the launching streams):
start / end relative to the step start, per stream.  Usage: python tools/wave_timeline.py STREAMS DEFER"""
import sys, os, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_quoridor_b200 import tree, _lib
from alphazero_quoridor_b200.selfplay import StreamedSelfPlay

streams, defer = int(sys.argv[1]), int(sys.argv[2])


class Proxy:
    def __init__(self, lib, log):
        self._lib, self._log = lib, log

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("qz_") or name in ("qz_last_error_string", "qz_rollout_workspace_bytes", "qz_rollout_pawn_passes"):
            return fn

        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sid = torch.cuda.current_stream().cuda_stream
            e0.record()
            rc = fn(*a)
            e1.record()
            self._log.append((sid, name, e0, e1))
            return rc
        return call


sp = StreamedSelfPlay(4096, lambda: tree.RolloutEvaluator(seed=1), n_streams=streams, n_playout=1000, c_puct=5.0,
                      leaves_per_game=64, pure=True, seed=1, defer_depth=max(defer, 0), defer_until_drain=defer < 0)
for _ in range(3):
    sp.step()
torch.cuda.synchronize()
log = []
prox = Proxy(_lib.load(), log)
for s in sp.subs:
    s.mcts.lib = prox
_lib.load = lambda: prox
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
sp.step()
b.record()
torch.cuda.synchronize()
print("streams %d defer %d: step %.1f ms" % (streams, defer, a.elapsed_time(b)))
sids = {}
rows = []
for sid, name, e0, e1 in log:
    k = sids.setdefault(sid, len(sids))
    rows.append((a.elapsed_time(e0), a.elapsed_time(e1), k, name.replace("qz_", "")))
rows.sort()
lo, hi = float(os.environ.get("TL_LO", "40")), float(os.environ.get("TL_HI", "75"))
for t0, t1, k, name in rows:
    if lo <= t0 <= hi:
        print("%8.2f %8.2f  (%6.2f)  s%-2d %s%s" % (t0, t1, t1 - t0, k, "    " * min(k, 6), name))
