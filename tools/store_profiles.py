"""Development aid: copies one measurement set (gpurun_out/*_<tag>.*, made by tools/measure_n1.sh / measure_multi.sh) into
profiles/r02_* and regenerates the ncu summaries.

  python tools/store_profiles.py n1 <tag>          single-GPU set
  python tools/store_profiles.py multi <tag>       bench / c5 lines of 2, 4 and 8 GPUs
"""
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PRO = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")


def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def rows(path):
    return [json.loads(l) for l in open(path) if l.startswith("{")]


def dump(obj, name):
    json.dump(obj, open(os.path.join(PRO, name), "w"), indent=1)


def summarize(*args):
    subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "summarize_ncu.py")] + list(args))


def main():
    kind, tag = sys.argv[1], sys.argv[2]
    if kind == "n1":
        dump(last_json(f"{OUT}/bench_n1_{tag}.json"), "r02_bench_n1.json")
        dump(last_json(f"{OUT}/bench_ref_{tag}.json"), "r02_bench_reference_n1.json")
        cfgs = rows(f"{OUT}/configs_n1_{tag}.json")
        dump(cfgs, "r02_configs_n1.json")
        dump([c for c in cfgs if c["config"].startswith("c5")][0], "r02_c5_n1.json")
        dump(rows(f"{OUT}/audit_{tag}.json"), "r02_audit.json")
        dump(rows(f"{OUT}/kbench_{tag}.json"), "r02_kbench.json")
        shutil.copy(f"{OUT}/buildtiming_{tag}.log", f"{PRO}/r02_buildtiming.log")
        shutil.copy(f"{OUT}/launches_bench_{tag}.csv", f"{PRO}/r02_launches_bench.csv")
        shutil.copy(f"{OUT}/launches_{tag}.csv", f"{PRO}/r02_launches_frames.csv")
        summarize("launches", f"{PRO}/r02_launches_bench.csv", f"{PRO}/r02_launches_bench.md",
                  "RT_GRAPH=0 python bench.py --steps 3 --warmup 3 (first 700 launches; ncu follows the reference-GPU comparator's child process too)")
        summarize("launches", f"{PRO}/r02_launches_frames.csv", f"{PRO}/r02_launches_frames.md",
                  "RT_GRAPH=0 python tools/frame_once.py spheres 11 1920 8 6")
        summarize("full", f"{OUT}/prof_{tag}.ncu-rep", f"{PRO}/r02_kernels_ncu_full.md",
                  "RT_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_extend|k_shade|k_tail -s 21 -c 7 "
                  "python tools/frame_once.py spheres 11 1920 8 6")
    else:
        for n in (2, 4, 8):
            dump(last_json(f"{OUT}/bench_n{n}_{tag}.json"), f"r02_bench_n{n}.json")
            dump(last_json(f"{OUT}/c5_n{n}_{tag}.json"), f"r02_c5_n{n}.json")


if __name__ == "__main__":
    main()
