#!/usr/bin/env python3
"""Contract benchmark (one JSON line on stdout, rank 0).

    python bench.py --gpus N --steps K --warmup W            # ours
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): pure-MCTS self-play, 1000 random rollouts per move, c_puct 5, 4096
concurrent games PER GPU (weak scaling; games are sharded by global game index, no collective on the hot
path).  One "step" = one ply of self-play for every game: 1000 playouts per game (descent, leaf legality sweep,
expansion, a random rollout of <= 999 plies, backup), first-max-visits move, env step, finished games restart.
Metric: env steps per second (tree-descent steps + rollout plies + the moves played), whole job.

The reference arm times the CPU oracle's port of the same path (oracle/quoridor_oracle.c: pure_mcts.py:66-115
on top of the literal quoridor.py rules) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the end-of-search pass runs ~18 stuck-rollout kernels on as many streams: with the default 8 hardware queues they
# would serialise in three rounds (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "env_steps_per_s"
UNIT = "env steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--playouts", type=int, default=1000)
    ap.add_argument("--leaves", type=int, default=64, help="leaves per game per wave (virtual loss)")
    ap.add_argument("--streams", type=int, default=1, help="sub-batches of games on separate CUDA streams")
    ap.add_argument("--defer", type=int, default=-1,
                    help="waves a stuck rollout may lag behind (0 = finish in-wave, -1 = all of them at the end of the search)")
    ap.add_argument("--c-puct", type=float, default=5.0)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the baseline sample")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline micro-section")
    ap.add_argument("--no-az", action="store_true", help="skip the AlphaZero-MCTS (BASELINE configs[2]) side measurement")
    ap.add_argument("--fix-terminal-sign", action="store_true",
                    help="back up a winning move as a win (the reference backs it up as a loss, mcts.py:125, so its "
                         "self-play games hardly ever end); default off = reference behaviour")
    ap.add_argument("--full-games", type=int, default=0, metavar="PLIES",
                    help="after the timed steps keep playing up to PLIES more plies and report finished self-play games/hr")
    return ap.parse_args()


def workload_name(a):
    return "pure_mcts selfplay: %d rollouts/move, %d concurrent games/GPU, c_puct %g, rollout limit 1000" % (
        a.playouts, a.games, a.c_puct)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_pure_mcts_sample(args, target_s, threads=0):
    """The oracle's port of pure_mcts.py timed on the host cores: G start-position games x P playouts with a
    fresh tree each, the reference's algorithm restated literally (full 128-candidate sweep per expansion AND per
    rollout ply, BFS path checks).  Returns the cpu_baseline dict."""
    import numpy as np
    from oracle import oracle as O
    cores = O.max_threads() if threads <= 0 else threads
    H = np.zeros(1, dtype=np.uint64)
    V = np.zeros(1, dtype=np.uint64)
    meta = np.array([[4, 76, 10, 10, 1]], dtype=np.int32)
    # probe: one game per core, 4 playouts
    g0 = max(cores, 1)
    t0 = time.perf_counter()
    steps, playouts, _ = O.pure_mcts_moves(np.repeat(H, g0), np.repeat(V, g0), np.repeat(meta, g0, 0), 4,
                                           c_puct=args.c_puct, seed=args.seed, threads=cores, literal_rollouts=True)
    dt = max(time.perf_counter() - t0, 1e-3)
    per_playout = dt * cores / max(playouts, 1)          # core-seconds per playout
    want_playouts = max(int(target_s * cores / per_playout), cores * 8)
    p = max(8, min(args.playouts, want_playouts // (2 * cores)))
    g = max(cores, min(args.games, want_playouts // p))
    t0 = time.perf_counter()
    steps, playouts, _ = O.pure_mcts_moves(np.repeat(H, g), np.repeat(V, g), np.repeat(meta, g, 0), p,
                                           c_puct=args.c_puct, seed=args.seed + 1, threads=cores, literal_rollouts=True)
    dt = time.perf_counter() - t0
    info = {"value": steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d start-position games x %d playouts (fresh tree each), %.1f s on %d host threads; "
                      "oracle/quoridor_oracle.c: literal quoridor.py rules + pure_mcts.py:66-115, one full actions() "
                      "sweep per rollout ply (the reference's 2 further recomputations per ply, quoridor.py:165 and "
                      "pure_mcts.py:9-10, are not repeated)" % (g, p, dt, cores),
            "playouts_per_s": playouts / dt, "seconds": dt}
    return info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        info = cpu_pure_mcts_sample(args, target_s=max(2.0, min(args.cpu_seconds, 60.0 / max(args.steps + args.warmup, 1))))
        if i >= args.warmup:
            vals.append(info)
        last = info
    steps_total = sum(v["value"] * v["seconds"] for v in vals)
    secs = sum(v["seconds"] for v in vals)
    value = steps_total / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(len(vals), 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "parallelism": "host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from alphazero_quoridor_b200 import _lib
    from alphazero_quoridor_b200.selfplay import StreamedSelfPlay
    from alphazero_quoridor_b200.tree import RolloutEvaluator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    per_stream = args.games // args.streams
    roll_events = []          # (start, end, n_rollouts) of every rollout launch (roofline of the dominant kernel)
    stuck_events = []         # (start, end) of every deferred stuck-rollout pass (side streams)
    evaluators = []

    def make_evaluator():
        ev = RolloutEvaluator(seed=args.seed, limit=1000)
        orig = ev.evaluate

        def timed_eval(mcts, lset, rids, **kw):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = orig(mcts, lset, rids, **kw)
            b.record()
            roll_events.append((a, b, lset.leaf_state.shape[0]))
            return out
        ev.evaluate = timed_eval
        orig_finish = ev.finish

        def timed_finish(mcts, lset):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = orig_finish(mcts, lset)
            b.record()
            stuck_events.append((a, b))
            return out
        ev.finish = timed_finish
        evaluators.append(ev)
        return ev

    sp = StreamedSelfPlay(args.games, make_evaluator, n_streams=args.streams, c_puct=args.c_puct,
                          n_playout=args.playouts, leaves_per_game=args.leaves, pure=True, seed=args.seed,
                          game_id_base=rank * args.games, device=dev, defer_depth=max(args.defer, 0),
                          defer_until_drain=args.defer < 0,
                          fix_terminal_sign=args.fix_terminal_sign)
    engines = [s.mcts for s in sp.subs]
    for m in engines:
        m.count_tree_steps = True

    def rollout_plies():
        return sum(ev.plies_played() for ev in evaluators)

    def tree_steps():
        return sum(int(m.tree_steps.item()) for m in engines)

    def zero_counters():
        for ev in evaluators:
            ev.zero_plies()
        for m in engines:
            m.tree_steps.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sp.step()
    barrier()
    roll_events.clear()
    stuck_events.clear()
    zero_counters()
    moves0 = sp.moves_played
    launches0 = _lib.LAUNCHES
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        sp.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = _lib.LAUNCHES - launches0
    env_steps = rollout_plies() + tree_steps() + (sp.moves_played - moves0)
    playouts = args.games * args.playouts * args.steps
    roll_ms = [a.elapsed_time(b) for a, b, _ in roll_events]
    roll_n = [n for _, _, n in roll_events]
    stuck_ms = [a.elapsed_time(b) for a, b in stuck_events]
    from alphazero_quoridor_b200.shard import reduce_stats
    ms, (env_total, playouts_total) = reduce_stats(ms, [env_steps, playouts], device=dev)   # max time, summed work

    # ---- end to end through the C ABI with HOST buffers: states H2D, search, moves/visits/new states D2H ----
    host_states = torch.empty((args.games, 3), dtype=torch.int64).pin_memory()
    host_moves = torch.empty((args.games,), dtype=torch.int32).pin_memory()
    host_visits = torch.empty((args.games, 140), dtype=torch.int32).pin_memory()
    host_states.copy_(torch.cat([m.root_state for m in engines], 0))
    start_row = sp.subs[0]._start.cpu()[0]
    torch.cuda.synchronize()
    zero_counters()
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    t0.record()
    for _ in range(e2e_steps):
        for st in sp.streams:
            st.wait_stream(cur)
        for i, (sub, st) in enumerate(zip(sp.subs, sp.streams)):
            with torch.cuda.stream(st):
                lo, hi = i * per_stream, (i + 1) * per_stream
                sub.mcts.reset(host_states[lo:hi].to(dev, non_blocking=True))
        plans = [sub.wave_plan() for sub in sp.subs]
        for w in range(max(len(p) for p in plans)):
            for sub, st, plan in zip(sp.subs, sp.streams, plans):
                if w < len(plan):
                    with torch.cuda.stream(st):
                        sub.mcts.playout_wave(plan[w])
        for i, (sub, st) in enumerate(zip(sp.subs, sp.streams)):
            with torch.cuda.stream(st):
                lo, hi = i * per_stream, (i + 1) * per_stream
                m = sub.mcts
                m.drain()
                visits, _, _ = m.root_stats(temp=1.0)
                moves = m.choose(mode=0)
                m.advance(moves, keep_subtree=False)
                host_moves[lo:hi].copy_(moves, non_blocking=True)
                host_visits[lo:hi].copy_(visits, non_blocking=True)
                host_states[lo:hi].copy_(m.root_state, non_blocking=True)
        torch.cuda.synchronize()
        # finished games restart on the host side of the boundary
        done = ((host_states[:, 2] >> 40) & 1).bool()
        if bool(done.any()):
            host_states[done] = start_row
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    e2e_env = rollout_plies() + tree_steps() + args.games * e2e_steps
    e2e_ms, (e2e_env,) = reduce_stats(e2e_ms, [e2e_env], device=dev)
    h2d = args.games * 24
    d2h = args.games * (4 + 140 * 4 + 24)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernels: the rollout launch = qz_rollout_{wall,pawn}_kernel on the main stream plus the
    # deferred qz_rollout_stuck_kernel pass on a side stream.  Algorithmic HBM bytes per rollout: 24 B start state + 4 B
    # state index + 8 B stream id read, 2 x 24 B through the phase hand-over buffer, 1 B result written.
    bytes_per_rollout = 24 + 4 + 8 + 48 + 1
    avg_main = sum(roll_ms) / max(len(roll_ms), 1)
    avg_stuck = sum(stuck_ms) / max(len(stuck_ms), 1)
    avg_ms = avg_main + avg_stuck
    avg_n = sum(roll_n) / max(len(roll_n), 1)
    achieved = bytes_per_rollout * avg_n / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    # whole-step warp-instruction count from the committed launch list of this same command (profiles/r1p_bench_launches.txt,
    # 7 steps of 17 waves x 262,144 leaf slots): the sweep and rollout kernels executed 47.17 G warp instructions per step
    # = 10,585 per leaf slot, select + expand 6.81 G per step.  Since r1q unused leaf slots cost (next to) nothing, so the
    # per-slot figure is applied to the playouts actually made.
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    issue_peak = sm_count * 4 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6           # warp instructions / s
    inst_per_step = 10585.0 * args.games * args.playouts + 6.81e9 * (args.games * args.playouts) / (4096.0 * 1000.0)
    roofline = {"kernel": "qz_rollout_wall_kernel + qz_rollout_pawn_kernel passes (waves) + qz_rollout_stuck_kernel (end of search)",
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of the three kernels in profiles/r1g_rollout_ncu_full.txt
                # (one 262,144-rollout launch: 0.7 + 0.2 + 6.3 MB), scaled to this run's rollouts per launch; it is BELOW
                # the algorithmic 85 B/rollout because the 126 MB L2 absorbs the 48 B/rollout phase hand-over buffer
                "traffic": 7.2e6 * avg_n / 262144.0, "peak_source": peak_src,
                "launches_timed": len(roll_ms), "avg_launch_ms": avg_ms, "avg_main_stream_ms": avg_main,
                "avg_deferred_stuck_pass_ms": avg_stuck, "rollouts_per_launch": avg_n,
                "share_of_step": (sum(roll_ms) + sum(stuck_ms)) / ms if ms > 0 else None,
                "issue_profile": {"source": "profiles/r1p_bench_launches.txt, profiles/r1m_rollout_ncu_full.txt, "
                                            "profiles/r1m_sweep_ncu_full.txt (ncu, B200)",
                                  "whole_step": {"warp_instructions": inst_per_step,
                                                 "achieved_warp_inst_per_s": inst_per_step / (ms / args.steps * 1e-3) if ms > 0 else None,
                                                 "peak_warp_inst_per_s": issue_peak,
                                                 "frac": inst_per_step / (ms / args.steps * 1e-3) / issue_peak if ms > 0 else None,
                                                 "note": "instruction count from the committed launch list (static), time from "
                                                         "this run; the ALU-pipe-bound kernels below reach 0.60-0.76 alone"},
                                  "instruction_share": {"qz_rollout_stuck_kernel": 0.290, "qz_legal_mask_kernel": 0.262,
                                                        "qz_rollout_pawn_kernel": 0.199, "qz_rollout_wall_kernel": 0.123,
                                                        "qz_mcts_select_kernel": 0.089, "qz_mcts_expand_backup_kernel": 0.037},
                                  "smsp_issue_active_pct": {"qz_rollout_wall_kernel": 60.8, "qz_rollout_pawn_kernel (first pass)": 76.0,
                                                            "qz_rollout_stuck_kernel": 43.3, "qz_legal_mask_kernel": 62.8},
                                  "active_lanes_per_instruction": {"qz_rollout_wall_kernel": 18.6, "qz_rollout_pawn_kernel": 23.7,
                                                                   "qz_rollout_stuck_kernel": 18.0, "qz_legal_mask_kernel": 18.7},
                                  "note": "per-kernel numbers are static, from the committed captures"},
                "note": "register-resident by design (24 B of state per game): instruction-issue / latency bound, not HBM "
                        "bound (SURVEY.md 8d), so frac against the HBM peak is ~1e-4 and says nothing -- issue_profile.whole_step "
                        "is the meaningful fraction.  avg_deferred_stuck_pass_ms is the span of one of the ~17 concurrent "
                        "end-of-search stuck passes, so share_of_step sums concurrent streams and exceeds 1.  The HBM-bound "
                        "kernels (step, encode) are under `kernels` with frac ~1."}
    line = {
        "metric": METRIC, "value": env_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "games_per_gpu": args.games, "playouts_per_move": args.playouts,
                   "leaves_per_game_per_wave": args.leaves, "cuda_streams": args.streams,
                   "stuck_rollout_defer_waves": args.defer if args.defer >= 0 else "end of search", "fix_terminal_sign": bool(args.fix_terminal_sign),
                   "parallelism": "games sharded by index x%d, no collective" % world,
                   "l2": "inputs larger than L2: tree arenas %.1f GB/GPU; rollouts are register resident"
                         % (sum(m.nbytes() for m in engines) / 1e9)},
        "mcts_sims_per_s": playouts_total / (ms * 1e-3),
        "moves_per_s": args.games * world * args.steps / (ms * 1e-3),
        "env_steps_per_playout": env_total / playouts_total,
        "clocks": clk,
        "e2e": {"value": e2e_env / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if not args.no_kernels:
        try:
            line["kernels"] = kernel_section(torch, dev, hbm_peak)
        except Exception as ex:  # the headline must survive a failure of the side section
            line["kernels"] = {"error": repr(ex)}
    if args.full_games > 0:
        fin0 = sum(int(sub.finished_games.item()) for sub in sp.subs)
        torch.cuda.synchronize()
        t_full = time.perf_counter()
        plies = 0
        while plies < args.full_games:
            sp.step()
            plies += 1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t_full
        fin = sum(int(sub.finished_games.item()) for sub in sp.subs) - fin0
        trunc = sum(int(sub.truncated_games.item()) for sub in sp.subs)
        line["selfplay_games"] = {"plies_played": plies, "seconds": dt, "games_finished": fin, "games_truncated": trunc,
                                  "games_per_hr": fin / dt * 3600.0, "note": "pure-MCTS self-play, %d playouts/move, "
                                  "games restart when they end; measured on rank 0's shard only" % args.playouts}
    if not args.no_az:
        try:
            line["az_mcts"] = az_section(torch, dev, args)
        except Exception as ex:
            line["az_mcts"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_pure_mcts_sample(args, args.cpu_seconds)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def az_section(torch, dev, args):
    """BASELINE configs[2] on this GPU (side measurement, not the headline): AlphaZero MCTS, n_playout = 100,
    c_puct = 5, random-init 5-block ResNet in bf16, 8192 concurrent games, 4 leaves per game per wave, tree reuse;
    three self-play plies (search + Dirichlet-mixed move + re-root) timed with CUDA events."""
    from alphazero_quoridor_b200.policy_value_net import PolicyValueNet
    from alphazero_quoridor_b200.selfplay import BatchedSelfPlay
    from alphazero_quoridor_b200.tree import NetEvaluator
    torch.manual_seed(0)
    net = PolicyValueNet(use_gpu=True, device=dev)
    n, npl, K = 8192, 100, 4
    sp = BatchedSelfPlay(n, NetEvaluator(net), c_puct=5, n_playout=npl, leaves_per_game=K, temp=1.0, pure=False,
                         seed=args.seed, device=dev)
    sp.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        sp.step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    out = {"workload": "AlphaZero MCTS n_playout=100, c_puct=5, 5-block ResNet bf16 (random init), 8192 games, K=4, tree reuse",
           "sims_per_s": 3 * n * npl / ms * 1e3, "ms_per_move": ms / 3, "moves_per_s": 3 * n / ms * 1e3,
           "net_flops_per_sim": 62.8e6, "net_tflops": 3 * n * npl * 62.8e6 / ms / 1e9,
           "tree_overflow": sp.mcts.overflow_count()}
    del sp, net
    torch.cuda.empty_cache()
    return out


def kernel_section(torch, dev, hbm_peak):
    """Live CUDA-event timings of the other kernels on the path with their algorithmic bytes."""
    from alphazero_quoridor_b200.quoridor import BatchedQuoridor
    from alphazero_quoridor_b200.synthetic import midgame_positions

    def timed(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    out = {}
    n = 1 << 20
    states = midgame_positions(n, seed=7, device=dev)
    env = BatchedQuoridor(n, states=states, device=dev)
    mask = torch.empty((n, 3), dtype=torch.int64, device=dev)
    ms = timed(lambda: env.legal_mask(out=mask))
    out["qz_legal_mask_kernel"] = {"units": n, "ms": ms, "sweeps_per_s": n / ms * 1e3, "bytes_per_unit": 48,
                                   "achieved_GBps": n * 48 / ms / 1e6, "frac_hbm": n * 48 / ms / 1e6 / hbm_peak,
                                   "bound": "issue (2 flood fills per candidate)",
                                   "workload": "BASELINE config 5: 2^20 positions, 10-20 plies of random play"}
    n2 = 1 << 19     # 2.2 GB of bf16 planes: larger than L2
    env2 = BatchedQuoridor(n2, states=states[:n2].clone(), device=dev)
    for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
        buf = torch.empty((n2, 26, 9, 9), dtype=dt, device=dev)
        ms = timed(lambda: env2.encode(out=buf))
        nb = buf.numel() * buf.element_size() + n2 * 24
        out["qz_encode_kernel_" + name] = {"units": n2, "ms": ms, "bytes_per_unit": nb // n2,
                                           "achieved_GBps": nb / ms / 1e6, "frac_hbm": nb / ms / 1e6 / hbm_peak, "bound": "hbm"}
        del buf
    # BASELINE config 1 literally: random games from reset() to terminal, FULL legal set computed on every ply
    ng = 1 << 20
    envg = BatchedQuoridor(ng, device=dev)
    envg.random_play(seed=1, max_plies=3000)
    envg.reset()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    plies = envg.random_play(seed=2, max_plies=3000)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    out["random_play_full_mask"] = {"games": ng, "ms": ms, "env_steps_per_s": float(plies.sum().item()) / ms * 1e3,
                                    "mean_plies": float(plies.float().mean().item()),
                                    "workload": "BASELINE config 1: 2^20 games from reset() to terminal, full 140-action legal "
                                                "set + uniform pick + step on every ply (qz_env_random_play, 2 launches)"}
    del envg
    n3 = 1 << 24     # 403 MB of states: larger than L2
    env3 = BatchedQuoridor(n3, device=dev)
    acts = torch.full((n3,), 2, dtype=torch.int32, device=dev)       # E then (opponent) E ... never terminal
    done = torch.empty(n3, dtype=torch.uint8, device=dev)
    ms = timed(lambda: env3.step(acts, done=done), iters=4, warm=2)
    nb = n3 * (24 + 4 + 24 + 1)
    out["qz_step_kernel"] = {"units": n3, "ms": ms, "bytes_per_unit": 53, "achieved_GBps": nb / ms / 1e6,
                             "frac_hbm": nb / ms / 1e6 / hbm_peak, "bound": "hbm", "steps_per_s": n3 / ms * 1e3}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
