"""Per-call durations with a device synchronize after every call (nothing overlaps): separates what a kernel costs
from what concurrency does to it.  Usage: python tools/serial_spans.py DEFER
(DEFER = waves a stuck pass may lag; 0 = finished in-wave)"""
import sys, os, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_quoridor_b200 import tree, _lib
from alphazero_quoridor_b200.selfplay import StreamedSelfPlay

defer = int(sys.argv[1])
agg = collections.OrderedDict()


class Proxy:
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("qz_") or name in ("qz_last_error_string", "qz_rollout_workspace_bytes", "qz_rollout_pawn_passes"):
            return fn

        def call(*a):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            torch.cuda.synchronize()
            d = agg.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += e0.elapsed_time(e1)
            return rc
        return call


sp = StreamedSelfPlay(4096, lambda: tree.RolloutEvaluator(seed=1), n_streams=1, n_playout=1000, c_puct=5.0,
                      leaves_per_game=64, pure=True, seed=1, defer_depth=max(defer, 0), defer_until_drain=defer < 0)
for _ in range(3):
    sp.step()
torch.cuda.synchronize()
prox = Proxy(_lib.load())
for s in sp.subs:
    s.mcts.lib = prox
_lib.load = lambda: prox
sp.step()
torch.cuda.synchronize()
print("defer %d, serialized:" % defer)
for k, v in agg.items():
    print("   %-28s n=%3d  total %7.2f ms  avg %.3f ms" % (k, v[0], v[1], v[1] / v[0]))
