"""Turns the ncu outputs brought back from the GPU box into the markdown / JSON summaries kept under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv  profiles/rNN_launches_bench.md  "<command>"
  python tools/summarize_ncu.py full     gpurun_out/prof.ncu-rep  profiles/rNN_kernels_ncu_full.md "<command>"
      (a capture of a C2 frame also rewrites profiles/extend_ncu.json - warp instructions, lanes and DRAM bytes per
       segment of the k_extend launches - which bench.py's roofline block reads)
  python tools/summarize_ncu.py full     gpurun_out/prof_c4.ncu-rep profiles/rNN_c4_....md "<command>" "<what was captured>"

`launches`: per-kernel totals and shares of a `--metrics gpu__time_duration.sum` launch list.
`full`:     one column per captured launch of a `--set full` report (read through `ncu -i ... --page raw --csv`).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency: active lanes per instruction (of 32)"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "issued warp instructions per clock per SM (of 4)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active % (FP32 issue)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__cycles_active.avg", "cycles an SM is active, average over the SMs"),
    ("sm__cycles_elapsed.max", "cycles of the launch (an SM active for fewer is idling at the launch's ends)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe (LSU wavefronts) % of peak, over the launch"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slot busy while the SM sub-partition is active (of 1)"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps per scheduler per active cycle"),
    ("smsp__warps_active.avg.per_cycle_active", "resident warps per scheduler"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
]


def short(name):
    """`void k_extend<0, 1>(DScene, ...)` -> `k_extend` (template arguments dropped)."""
    return name.split("(")[0].replace("void ", "").split("<")[0][:70]


def launches(csv_path, out_path, command):
    lines = [l for l in open(csv_path, errors="replace") if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()
    order = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "us")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
        k = short(r["Kernel Name"])
        per.setdefault(k, []).append(us)
        order.append((k, us))
    total = sum(sum(v) for v in per.values())
    out = [f"# ncu launch list of `{command}` (1 x B200)", "",
           "Captured with `ncu --metrics gpu__time_duration.sum --clock-control none --csv` after the same command exited 0 "
           "without ncu. Per-launch times are cold-cache and serialised: compare shares, not absolutes.", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.2f} | {100 * sum(v) / total:.1f}% |")
    # the last complete frame: from the last k_generate to the end
    gen = [i for i, (k, _) in enumerate(order) if k == "k_generate"]
    if gen:
        frame = order[gen[-1]:]
        nxt = [i for i, (k, _) in enumerate(frame) if i > 0 and k == "k_generate"]
        frame = frame[: nxt[0]] if nxt else frame
        out += ["", "Last frame of the run, kernel by kernel (us):", "", "```",
                " ".join(f"{k.replace('k_', '')}:{us:.0f}" for k, us in frame if k.startswith("k_")), "```"]
        render = collections.OrderedDict()
        for k, us in frame:
            if k in ("k_generate", "k_extend", "k_shade", "k_tail", "k_accumulate"):
                render[k] = render.get(k, 0.0) + us
        tot = sum(render.values())
        out += ["", "Shares of the render kernels in that frame: " +
                ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in render.items()) + "."]
    # bench.py's steps: generate, 3 x (extend, shade), tail, present - every occurrence of that launch sequence
    pattern = ["k_generate", "k_extend", "k_shade", "k_extend", "k_shade", "k_extend", "k_shade", "k_tail", "k_present_rgb8"]
    names = [k for k, _ in order]
    steps = [order[i:i + len(pattern)] for i in range(len(order) - len(pattern) + 1) if names[i:i + len(pattern)] == pattern]
    if steps:
        med = [sorted(st[j][1] for st in steps)[len(steps) // 2] for j in range(len(pattern))]
        tot = sum(med)
        agg = collections.OrderedDict()
        for k, us in zip(pattern, med):
            agg[k] = agg.get(k, 0.0) + us
        out += ["", f"The benchmark step (C2 frame + present) occurs {len(steps)} times in the list; median per launch (us):", "", "```",
                " ".join(f"{k.replace('k_', '')}:{us:.0f}" for k, us in zip(pattern, med)), "```", "",
                f"Shares of the step's {tot:.0f} us (serialised, cold caches): " +
                ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in agg.items()) + "."]
    open(out_path, "w").write("\n".join(out) + "\n")
    print("wrote", out_path)


# queue lengths of the three k_extend launches of a C2 frame (bench.py's seed 1000; they move by < 0.1 % with the seed)
C2_EXTEND_SEGMENTS = [2073600, 1727915, 734978]


def full(rep_path, out_path, command, what=None, extend_segments=None):
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kcol = hdr.index("Kernel Name")
    names = []
    count = collections.Counter()
    for d in data:
        k = short(d[kcol]).replace("k_", "")
        names.append(f"{k} #{count[k]}")
        count[k] += 1
    what = what or "one C2 frame: 1920x1080 paths, 1 spp, depth 8, 485-sphere scene"
    out = [f"# `ncu --set full` of the render kernels of {what.split(':')[0]} (1 x B200)", "",
           f"Command: `{command}` (after the same command exited 0 without ncu). {what[0].upper() + what[1:]}; `#k` = the k-th "
           "launch of that kernel in the frame (bounce k for extend / shade).", "",
           "| metric | unit | " + " | ".join(names) + " |", "|---|---|" + "---:|" * len(names)]
    for m, label in FULL_METRICS:
        if m not in hdr:
            continue
        c = hdr.index(m)
        cells = []
        for d in data:
            try:
                cells.append(f"{float(d[c].replace(',', '')):.2f}")
            except ValueError:
                cells.append(d[c])
        out.append(f"| {label} (`{m}`) | {units[c]} | " + " | ".join(cells) + " |")
    open(out_path, "w").write("\n".join(out) + "\n")
    print("wrote", out_path)
    # DRAM traffic of the k_extend launches -> bench.py's roofline.traffic
    r, w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = [float(d[r].replace(",", "")) * scale[units[r]] + float(d[w].replace(",", "")) * scale[units[w]]
           for d in data if short(d[kcol]) == "k_extend"]
    if per and extend_segments is not None:
        # per-segment figures of the k_extend launches -> bench.py's roofline block (profiles/extend_ncu.json)
        ic, lc = hdr.index("smsp__inst_executed.sum"), hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio")
        ext = [d for d in data if short(d[kcol]) == "k_extend"]
        inst = [float(d[ic].replace(",", "")) for d in ext]
        lanes = [float(d[lc].replace(",", "")) for d in ext]
        segs = list(extend_segments)[: len(ext)]
        tpath = os.path.join(REPO, "profiles", "extend_ncu.json")
        json.dump({"launches": "k_extend launches of one C2 frame (the launches bench.py times)",
                   "source": os.path.relpath(out_path, REPO) + " (ncu --set full)",
                   "segments_per_launch": segs, "warp_inst_per_launch": inst, "lanes_per_inst_per_launch": lanes,
                   "dram_bytes_per_launch_each": per,
                   "warp_inst_per_segment": sum(inst) / sum(segs),
                   "lanes_per_inst": sum(i * l for i, l in zip(inst, lanes)) / sum(inst),
                   "dram_bytes_per_launch": sum(per) / len(per), "dram_bytes_per_segment": sum(per) / sum(segs)},
                  open(tpath, "w"), indent=1)
        print("wrote", tpath)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    if mode == "launches":
        launches(src, dst, cmd)
    else:  # full <rep> <out.md> "<command>" ["<what was captured>"]: a C2 capture (no description) also refreshes extend_ncu.json
        what = sys.argv[5] if len(sys.argv) > 5 else None
        full(src, dst, cmd, what, None if what else C2_EXTEND_SEGMENTS)
