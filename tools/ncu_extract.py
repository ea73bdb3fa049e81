#!/usr/bin/env python3
"""Summarise .ncu-rep captures / launch lists into the small CSVs committed under profiles/.
    python tools/ncu_extract.py metrics OUT.csv "header comment" A.ncu-rep [B.ncu-rep ...]
    python tools/ncu_extract.py launches OUT.csv "header comment" LAUNCHES.csv STEPS"""
import collections, csv, io, re, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def metrics(out, comment, reps):
    cols, names, units = [], [], {}
    for rep in reps:
        h, u, rows = raw(rep)
        ix = {n: i for i, n in enumerate(h)}
        for row in rows:
            names.append(re.sub(r"\(.*", "", row[ix["Kernel Name"]]).replace("void <unnamed>::", ""))
            cols.append({w: row[ix[w]] for w in WANT if w in ix})
            units.update({w: u[ix[w]] for w in WANT if w in ix})
    with open(out, "w") as f:
        f.write(f"# {comment}\n")
        f.write("metric,unit," + ",".join('"%s"' % n for n in names) + "\n")
        for w in WANT:
            if any(w in c for c in cols):
                f.write(f"{w},{units.get(w, '')}," + ",".join(c.get(w, "") for c in cols) + "\n")


def launches(out, comment, src, steps):
    lines = [l for l in open(src) if not l.startswith("==")]
    r = list(csv.reader(lines))
    ix = {n: i for i, n in enumerate(r[0])}
    agg = collections.OrderedDict()
    for row in r[1:]:
        if len(row) < len(r[0]):
            continue
        try:
            v = float(row[ix["Metric Value"]])
        except ValueError:
            continue
        unit = row[ix["Metric Unit"]]
        us = v / 1000 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000)
        key = re.sub(r"\(.*", "", row[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")[:64]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {comment}\n")
        f.write(f"# {sum(a[0] for a in agg.values())} launches over {steps} joint steps, {tot / 1000:.2f} ms serialised "
                "(cold-cache per-launch times: compare SHARES with bench.py's families, not absolutes)\n")
        f.write("kernel,launches,launches_per_step,total_ms,share,avg_us\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f'"{k}",{a[0]},{a[0] / steps:.1f},{a[1] / 1000:.3f},{a[1] / tot:.4f},{a[1] / a[0]:.1f}\n')


if __name__ == "__main__":
    if sys.argv[1] == "metrics":
        metrics(sys.argv[2], sys.argv[3], sys.argv[4:])
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]))
