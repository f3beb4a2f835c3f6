#!/usr/bin/env python
"""Turns an `ncu --set full` report into the markdown table kept under profiles/ and updates
profiles/ncu_traffic.json (DRAM bytes per launch of the estimate kernel, read by bench.py).

  python tools/ncu_summary.py gpurun_out/prof_all_r1d.ncu-rep profiles/r01d_ncu_summary.md \
      --title "r01d ..." --batch 128
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROWS = [
    ("time (us)", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("regs / thread", "launch__registers_per_thread"),
    ("warp instructions (M)", "smsp__inst_executed.sum"),
    ("DRAM read (MB)", "dram__bytes_read.sum"),
    ("DRAM write (MB)", "dram__bytes_write.sum"),
    ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
    ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("cycles active / elapsed", None),
    ("fp64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("stalled warps per issue: barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("… long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("… short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("… wait (fixed latency)", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("… math_pipe_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("… not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("… mio_throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("… lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * scale.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.check_output(["ncu", "-i", a.report, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, kernels = rows[0], rows[1], rows[2:]

    def get(r, key):
        return (r[hdr.index(key)], units[hdr.index(key)]) if key in hdr else (None, None)

    names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in kernels]
    lines = ["# " + a.title, "",
             "Source: `%s` (`ncu --set full --clock-control none --import-source on`, one launch of "
             "each kernel inside `bench.py`, batch %d x 1280x1024).  Times under ncu are "
             "serialised and cold-cache: compare shares, not absolutes. %s" %
             (os.path.basename(a.report), a.batch, a.note), "",
             "| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    traffic = {}
    for label, key in ROWS:
        cells = []
        for r, nme in zip(kernels, names):
            if key is None:
                act, _ = get(r, "sm__cycles_active.avg")
                ela, _ = get(r, "sm__cycles_elapsed.avg")
                cells.append("%.2f" % (float(act) / float(ela)) if act and ela else "-")
                continue
            v, u = get(r, key)
            if v is None:
                cells.append("-")
            elif "bytes" in key:
                b = to_bytes(v, u)
                cells.append("%.1f" % (b / 1e6))
                traffic.setdefault(nme, 0.0)
                traffic[nme] += b
            elif key == "gpu__time_duration.sum":
                t = float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                cells.append("%.1f" % t)
            elif key == "smsp__inst_executed.sum":
                cells.append("%.1f" % (float(v) / 1e6))
            else:
                cells.append(("%.2f" if key.endswith(".ratio") else "%.1f") % float(v)
                             if "." in v else v)
        lines.append("| " + label + " | " + " | ".join(cells) + " |")
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    tj = os.path.join(os.path.dirname(os.path.abspath(a.out)), "ncu_traffic.json")
    est = [n for n in traffic if n.startswith("estimate")]
    if est:
        j = {"estimate_kernel": {"kernel": est[0], "batch": a.batch,
                                 "dram_bytes_per_launch": traffic[est[0]],
                                 "source": "ncu --set full, %s, dram__bytes_read.sum + "
                                           "dram__bytes_write.sum" % os.path.basename(a.report)},
             "all_kernels_dram_bytes_per_launch": traffic}
        with open(tj, "w") as f:
            json.dump(j, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
