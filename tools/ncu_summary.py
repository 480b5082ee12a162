#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU) into profiles/<name>.md + .json.

    python tools/ncu_summary.py gpurun_out/prof_r1c_cfg2.ncu-rep --points 268435456 --bytes-per-point 8
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--points", type=int, required=True)
    ap.add_argument("--bytes-per-point", type=int, required=True)
    ap.add_argument("--out", default=None)
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    if args.rep.endswith(".csv"):            # `ncu -i x.ncu-rep --page raw --csv` exported on the GPU box
        with open(args.rep) as fh:
            raw = fh.read()
    else:
        raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = rows[2:]
    name = args.out or os.path.splitext(os.path.basename(args.rep))[0]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {"report": os.path.basename(args.rep), "points_per_launch": args.points,
           "algorithmic_bytes_per_launch": args.points * args.bytes_per_point, "launches": []}
    md = ["# ncu summary: %s" % os.path.basename(args.rep), "",
          "`ncu --set full --clock-control none --import-source on` under gpurun; per-launch values "
          "(cold cache, serialised replays -- compare shares, not absolutes). %s" % args.note, ""]
    for r in launches:
        d = dict(zip(hdr, r))
        f = lambda k: float(d[k]) if d.get(k, "") not in ("", "n/a") else float("nan")
        unit = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name"), "metrics": {}}
        for k in KEYS:
            if k in d:
                rec["metrics"][k] = {"value": d[k], "unit": unit.get(k, "")}
        stalls = {k[len(STALLS):].replace("_per_issue_active.ratio", ""): float(d[k]) for k in hdr
                  if k.startswith(STALLS) and k.endswith("per_issue_active.ratio") and d[k] not in ("", "n/a")}
        rec["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = f("dram__bytes_read.sum") * scale.get(unit.get("dram__bytes_read.sum"), 1.0)
        wr = f("dram__bytes_write.sum") * scale.get(unit.get("dram__bytes_write.sum"), 1.0)
        tscale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}
        t = f("gpu__time_duration.sum") * tscale.get(unit.get("gpu__time_duration.sum"), 1e-6)
        wf = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        rec["derived"] = {
            "dram_traffic_bytes": rd + wr,
            "traffic_over_algorithmic": (rd + wr) / out["algorithmic_bytes_per_launch"],
            "duration_s": t,
            "algorithmic_GBps_under_ncu": out["algorithmic_bytes_per_launch"] / t / 1e9,
            "smem_wavefronts_per_warp_point": wf / args.points * 32,
            "smem_bank_conflict_wavefronts_per_warp_point": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / args.points * 32,
            "smem_pipe_busy_frac": wf / 148.0 / f("sm__cycles_active.avg"),
            "instructions_per_point": f("smsp__inst_executed.sum") * 32 / args.points,
        }
        out["launches"].append(rec)
    r0 = out["launches"][0]
    md.append("Kernel: `%s`" % r0["kernel"])
    md.append("")
    md.append("| metric | value |")
    md.append("|---|---|")
    for k, v in r0["metrics"].items():
        md.append("| %s | %s %s |" % (k, v["value"], v["unit"]))
    for k, v in r0["derived"].items():
        md.append("| **%s** | %.6g |" % (k, v))
    md.append("")
    md.append("Top warp stall reasons (warps stalled per issue): " +
              ", ".join("%s %.2f" % kv for kv in r0["stalls_per_issue"].items()))
    os.makedirs(os.path.join(root, "profiles"), exist_ok=True)
    with open(os.path.join(root, "profiles", name + ".json"), "w") as fh:
        json.dump(out, fh, indent=1)
    with open(os.path.join(root, "profiles", name + ".md"), "w") as fh:
        fh.write("\n".join(md) + "\n")
    print("\n".join(md[5:]))


if __name__ == "__main__":
    main()
