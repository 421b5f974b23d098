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


def test_launch_list_summary_can_be_regenerated(tmp_path):
    out = str(tmp_path / "launches.md")
    subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "summarize_ncu.py"), "launches",
                           os.path.join(PROFILES, "r01b_launches_bench.csv"), out, "python bench.py --steps 3 --warmup 3"])
    new = [l for l in open(out) if l.startswith("| `k_")]
    old = [l for l in open(os.path.join(PROFILES, "r01b_launches_bench.md")) if l.startswith("| `k_")]
    assert new == old and any("k_extend" in l for l in new)
