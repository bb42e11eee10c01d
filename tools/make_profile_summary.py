#!/usr/bin/env python
"""Turn an ncu capture of the step kernel (+ an optional launch-list CSV) into the committed
summaries under profiles/.

    python tools/make_profile_summary.py gpurun_out/prof.ncu-rep [gpurun_out/launches.csv] --tag r01

Writes profiles/<tag>_step_kernel.json (machine readable, read by bench.py for roofline.traffic),
profiles/<tag>_step_kernel.md (human readable) and profiles/<tag>_launches.csv (the launch list).
Needs only the CUDA toolkit's `ncu` (no GPU).
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import kernel_source_hash  # noqa: E402  (bench.py reports roofline.traffic only while the sources are unchanged)

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__icc_request_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.max",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_static",
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:  # noqa: BLE001
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("launches", nargs="?")
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--name", default="step_kernel", help="file stem: profiles/<tag>_<name>.{json,md}")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--algo-bytes", type=float, default=796.0)
    args = ap.parse_args()
    out_dir = os.path.join(REPO, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units, data = rows[0], rows[1], rows[2:]
    kcol = H.index("Kernel Name")
    launches = []
    for r in data:
        rec = {"kernel": r[kcol]}
        for m in METRICS:
            if m in H:
                rec[m] = num(r[H.index(m)])
                rec[m + "__unit"] = units[H.index(m)]
        stalls = {}
        for i, h in enumerate(H):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                v = num(r[i])
                if v:
                    stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = v
        tot = sum(stalls.values()) or 1.0
        rec["stall_share_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])}
        launches.append(rec)
    first = launches[0]

    def unit_scale(m):
        u = first.get(m + "__unit", "")
        return {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)

    traffic = sum((l["dram__bytes_read.sum"] * unit_scale("dram__bytes_read.sum") +
                   l["dram__bytes_write.sum"] * unit_scale("dram__bytes_write.sum")) for l in launches) / len(launches)
    dur = sum(l["gpu__time_duration.sum"] * unit_scale("gpu__time_duration.sum") for l in launches) / len(launches)
    summary = {
        "tag": args.tag, "source": os.path.basename(args.rep), "kernel": first["kernel"], "launches_profiled": len(launches),
        "kernel_source_hash": kernel_source_hash(),
        "envs_per_launch": args.envs,
        "dram_bytes_per_launch": traffic, "dram_bytes_per_env_step": traffic / args.envs,
        "algorithmic_bytes_per_env_step": args.algo_bytes, "traffic_over_algorithmic": traffic / (args.envs * args.algo_bytes),
        "duration_under_ncu_s": dur,
        "note": "durations under ncu are serialised/cold-cache: compare shares, not absolutes; bench.py times with CUDA events",
        "launch": {k: v for k, v in first.items() if not k.endswith("__unit")},
    }
    # launch-list shares
    if args.launches:
        shutil.copyfile(args.launches, os.path.join(out_dir, "%s_launches.csv" % args.tag))
        lr = list(csv.reader(open(args.launches)))
        hi = [i for i, r in enumerate(lr) if r and r[0] == "ID"][0]
        LH = lr[hi]
        ki, vi = LH.index("Kernel Name"), LH.index("Metric Value")
        agg = defaultdict(lambda: [0, 0.0])
        for r in lr[hi + 1:]:
            if len(r) > vi and num(r[vi]) is not None:
                agg[r[ki]][0] += 1
                agg[r[ki]][1] += num(r[vi])
        tot = sum(v[1] for v in agg.values()) or 1.0
        summary["launch_list"] = [{"kernel": k, "launches": v[0], "total_ns": v[1], "share_pct": round(100 * v[1] / tot, 2)}
                                  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    with open(os.path.join(out_dir, "%s_%s.json" % (args.tag, args.name)), "w") as f:
        json.dump(summary, f, indent=1)
    L = summary["launch"]
    md = ["# %s — ncu summary of `%s`" % (args.tag, first["kernel"]), "",
          "Source: `%s` (`ncu --set full --clock-control none --import-source on`, %d launches in steady state; "
          "bench.py workload: stage_5, %d envs)." % (os.path.basename(args.rep), len(launches), args.envs), "",
          "| metric | value |", "|---|---|"]
    for m in METRICS:
        if m in L and L[m] is not None:
            md.append("| `%s` | %s %s |" % (m, ("%.4g" % L[m]), first.get(m + "__unit", "")))
    md += ["| DRAM traffic per launch (read+write) | %.1f MB = %.0f B per env-step (algorithmic %.0f B, ratio %.2f) |" %
           (traffic / 1e6, traffic / args.envs, args.algo_bytes, summary["traffic_over_algorithmic"]), "",
           "Warp-state samples (share of all samples): " +
           ", ".join("%s %.1f %%" % kv for kv in list(first["stall_share_pct"].items())[:8]), ""]
    if "launch_list" in summary:
        md += ["Launch list of the same command (`--metrics gpu__time_duration.sum`, cold-cache, serialised — shares only):", "",
               "| kernel | launches | share of GPU time |", "|---|---|---|"]
        for e in summary["launch_list"][:8]:
            md.append("| `%s` | %d | %.2f %% |" % (e["kernel"][:70], e["launches"], e["share_pct"]))
        md.append("")
    with open(os.path.join(out_dir, "%s_%s.md" % (args.tag, args.name)), "w") as f:
        f.write("\n".join(md))
    print("\n".join(md))


if __name__ == "__main__":
    main()
