"""The committed evidence under profiles/ stays readable: bench lines carry the contract keys, the ncu launch list
can be summarised again with tools/summarize_ncu.py and agrees with the committed summary."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(REPO, "profiles")


def test_committed_bench_lines_follow_the_contract():
    for n in (1, 2, 4, 8):
        d = json.load(open(os.path.join(PROFILES, f"r01b_bench_n{n}.json")))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["higher_is_better"] is True
        assert abs(d["value"] - d["config"]["paths_per_step"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    values = [json.load(open(os.path.join(PROFILES, f"r01b_bench_n{n}.json")))["value"] for n in (1, 2, 4, 8)]
    assert values == sorted(values) and values[3] > 7 * values[0]
    ref = json.load(open(os.path.join(PROFILES, "r01b_bench_reference_n1.json")))
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "reference"


def test_round2_bench_lines_and_scaling_curves():
    """Round 2: the contract keys plus the blocks the round-1 review asked for (issue-bound roofline, RMSE against the
    reference, parity audit, traversal counters, strong-scaling leg with an N-independent frame)."""
    lines = {n: json.load(open(os.path.join(PROFILES, f"r02_bench_n{n}.json"))) for n in (1, 2, 4, 8)}
    for n, d in lines.items():
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline",
                  "rmse", "parity", "traversal", "static_4k", "frame_sha"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["gpu_launches"] > 0
        r = d["roofline"]
        assert r["bound"] == "issue" and 0.05 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0.3 < r["issue"]["frac"] < 1.0 and 0.0 < r["hbm"]["dram_frac"] < 0.5
        assert d["rmse"]["ratio"] <= 1.15 and abs(d["rmse"]["luminance_rel_diff"]) <= 0.005
        assert d["parity"]["fast_vs_exact"]["mismatch_rate"] <= 2e-5
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    one = lines[1]
    assert one["gpu_reference"]["list_world"]["whole_frame_launch"]["ms_per_frame"] > 100 * one["ms_per_step"]
    for n in (2, 4, 8):  # weak scaling of the frame loop, strong scaling of the 4K static render, same picture
        assert lines[n]["value"] > 0.94 * n * one["value"], n
        assert one["static_4k"]["ms"] / lines[n]["static_4k"]["ms"] > 0.95 * n, n
        assert lines[n]["static_4k"]["frame_sha"] == one["static_4k"]["frame_sha"]
    c5 = {n: json.load(open(os.path.join(PROFILES, f"r02_c5_n{n}.json")))["ms"] for n in (1, 2, 4, 8)}
    assert c5[1] / c5[8] > 0.95 * 8
    # the round's last build (larger passes for long static renders): same pictures, less time
    last = json.load(open(os.path.join(PROFILES, "r02_bench_n1_final.json")))
    assert last["static_4k"]["frame_sha"] == one["static_4k"]["frame_sha"] and last["static_4k"]["ms"] < one["static_4k"]["ms"]
    assert last["value"] > 0.98 * one["value"] and last["rmse"]["ratio"] <= 1.15
    c5_last = {n: json.load(open(os.path.join(PROFILES, f"r02_c5_n{n}_final.json"))) for n in (1, 2)}
    assert c5_last[1]["ms"] < c5[1] and c5_last[1]["ms"] / c5_last[2]["ms"] > 0.95 * 2
    assert c5_last[1]["mean_luminance"] == c5_last[2]["mean_luminance"]


def test_launch_list_summary_can_be_regenerated(tmp_path):
    out = str(tmp_path / "launches.md")
    subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "summarize_ncu.py"), "launches",
                           os.path.join(PROFILES, "r01b_launches_bench.csv"), out, "python bench.py --steps 3 --warmup 3"])
    new = [l for l in open(out) if l.startswith("| `k_")]
    old = [l for l in open(os.path.join(PROFILES, "r01b_launches_bench.md")) if l.startswith("| `k_")]
    assert new == old and any("k_extend" in l for l in new)
    # round 2: the benchmark step inside bench.py's launch list - the shares bench.py's own stage times must agree with
    out2 = str(tmp_path / "launches2.md")
    subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "summarize_ncu.py"), "launches",
                           os.path.join(PROFILES, "r02_launches_bench.csv"), out2, "bench.py"])
    text = open(out2).read()
    assert "The benchmark step" in text
    share = float(text.split("k_extend ")[-1].split("%")[0])
    stages = json.load(open(os.path.join(PROFILES, "r02_bench_n1.json")))["roofline"]["stage_ms_per_frame"]
    live = 100 * stages["extend"] / sum(stages.values())
    assert abs(share - live) < 5.0, (share, live)
