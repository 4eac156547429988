"""CPU-only checks of bench.py's contract pieces that need no GPU: the reference arm's JSON line (same `config` object
as our arm, the keys the driver reads), the position fixture, and the issue-roofline arithmetic over an instruction table
with sampled plies."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_reference_arm_line_and_config():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--playouts", "12", "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    # the same workload object as our arm would print for the same arguments
    a = bench.parse(["--steps", "1", "--warmup", "1", "--playouts", "12"])
    assert line["config"] == bench.config_of(a, 1)
    assert line["metric"] == bench.METRICS["pure"][0]


def test_bench_positions_fixture():
    """oracle/bench_positions.npz: the GPU arm's own positions per self-play ply (tools/gen_bench_positions.py)."""
    H, V, meta5, src = bench.bench_positions(0, 8)
    assert "bench_positions.npz" in src
    assert (H == 0).all() and (V == 0).all() and (meta5 == np.array([4, 76, 10, 10, 1])).all()      # ply 0 = reset()
    H, V, meta5, _ = bench.bench_positions(12, 32)
    assert len(H) == 32 and ((meta5[:, 4] == 1) | (meta5[:, 4] == 2)).all()
    assert ((meta5[:, 0] >= 0) & (meta5[:, 0] <= 80) & (meta5[:, 1] >= 0) & (meta5[:, 1] <= 80)).all()
    walls_placed = np.array([bin(int(h)).count("1") + bin(int(v)).count("1") for h, v in zip(H, V)])
    assert (walls_placed + meta5[:, 2] + meta5[:, 3] == 20).all()                                       # walls are conserved


def test_issue_roofline_interpolates_sampled_plies(tmp_path, monkeypatch):
    table = {"when": "t", "build_hash": bench.build_hash(),
             "args": {"games": 4096, "playouts": 1000, "leaves": 64, "seed": 20261017, "defer": -1},
             "per_ply": {str(p): {"warp_inst": 1e9 * (10 + p), "thread_inst": 1e9 * (10 + p) * 24.0, "launches": 10,
                                  "kernels": {"k": {"warp_inst": 1e9 * (10 + p)}}} for p in (5, 9, 13)}}
    path = tmp_path / "inst_table.json"
    path.write_text(json.dumps(table))
    monkeypatch.setattr(bench, "INST_TABLE", str(path))
    a = bench.parse(["--steps", "9", "--warmup", "5"])                  # plies 5..13: 5, 9, 13 captured, the rest interpolated
    r = bench.issue_roofline(a, 900.0, {"sm_max_mhz": 2000.0}, 100, 4096 * 1000, None)
    want = sum(1e9 * (10 + p) for p in range(5, 14))                    # linear data: interpolation is exact
    assert abs(r["warp_instructions_per_step"] * 9 - want) < 1.0
    assert r["bound"] == "issue" and abs(r["peak"] - 100 * 4 * 2000.0e6 / 1e9) < 1e-9
    assert abs(r["frac"] - want / 0.9 / (100 * 4 * 2000.0e6)) < 1e-12
    assert abs(r["active_lanes_per_instruction"] - 24.0) < 1e-9
    assert r["table"]["stale"] is False and r["table"]["plies_interpolated"] == [6, 7, 8, 10, 11, 12]
    table["build_hash"] = "other"
    path.write_text(json.dumps(table))
    assert bench.issue_roofline(a, 900.0, {}, 100, 4096 * 1000, None)["table"]["stale"] is True
