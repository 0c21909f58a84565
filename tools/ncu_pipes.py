"""Per-kernel tensor-pipe / SFU / issue / DRAM utilisation from an ``ncu --csv`` launch list taken with
--metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,
smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed : one row per (kernel, grid), duration-weighted
averages over ONE reverse step (delimited by stem_assemble launches).

    python tools/ncu_pipes.py profiles/r02_ncu_pipes.csv > profiles/r02_ncu_pipes_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rows = csv.reader(lines)
    hdr = next(rows)
    ix = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r[ix["ID"]]), {"name": re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", r[ix["Kernel Name"]])).replace("wsr::", ""),
                                                   "grid": r[ix["Grid Size"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        m = r[ix["Metric Name"]]
        if m.startswith("gpu__time_duration"):
            v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ix["Metric Unit"]], 1.0)
        d[m] = v
    ids = sorted(launches)
    marks = [i for i in ids if "stem_assemble" in launches[i]["name"]]
    if len(marks) >= 2:
        ids = [i for i in ids if marks[0] <= i < marks[1]]
    T, X, I, D = ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
                  "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed")
    agg = OrderedDict()
    for i in ids:
        d = launches[i]
        t = d.get("gpu__time_duration.sum", 0.0)
        a = agg.setdefault((d["name"], d["grid"]), [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t
        for k, m in enumerate((T, X, I, D)):
            a[2 + k] += t * d.get(m, 0.0)
    total = sum(a[1] for a in agg.values())
    print("# one reverse step, %d launches, %.3f ms of serialised kernel time (ncu, cold caches: compare shares, not absolutes)" % (len(ids), total / 1e6))
    print("# %-58s %-14s %4s %9s %6s %8s %6s %7s %6s" % ("kernel", "grid", "n", "time_us", "share", "tensor%", "sfu%", "issue%", "dram%"))
    for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        t = max(a[1], 1e-9)
        print("%-60s %-14s %4d %9.1f %5.1f%% %8.1f %6.1f %7.1f %6.1f" % (name[:60], grid, a[0], a[1] / 1e3, 100 * a[1] / total, a[2] / t, a[3] / t, a[4] / t, a[5] / t))
    tc = [(k, a) for k, a in agg.items() if k[0].startswith("gemm_tc_kernel")]
    tt = sum(a[1] for _, a in tc)
    if tt > 0:
        print("# gemm_tc_kernel (all convolutions + GEMMs): %.3f ms, duration-weighted tensor-pipe utilisation %.1f %% of peak" % (tt / 1e6, sum(a[2] for _, a in tc) / tt))
    at = [(k, a) for k, a in agg.items() if k[0].startswith("attn")]
    for k, a in at:
        print("# %s %s: tensor %.1f %%, sfu %.1f %%" % (k[0], k[1], a[2] / max(a[1], 1e-9), a[3] / max(a[1], 1e-9)))


if __name__ == "__main__":
    main(sys.argv[1])
