"""Per-kernel summary of an ``ncu --csv`` launch list (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum).

    python tools/ncu_summary.py profiles/r01d_ncu_launches.csv > profiles/r01d_ncu_launch_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("wsr::", "")


def main(path):
    launches = OrderedDict()
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rows = csv.reader(lines)
    hdr = next(rows)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        d = launches.setdefault(r[ix["ID"]], {"name": short(r[ix["Kernel Name"]])})
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m.startswith("gpu__time_duration"):
            val *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        else:
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = val
    agg = OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    print("# kernel                                                        launches        time_ns  share   dram_read_MB  dram_write_MB")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s %6d %14.0f %5.1f%% %12.1f %12.1f" % (name[:62], a[0], a[1], 100 * a[1] / total, a[2] / 1e6, a[3] / 1e6))
    tc = [a for n, a in agg.items() if n.startswith("gemm_tc_kernel")]
    if tc:
        n = sum(a[0] for a in tc)
        print("# gemm_tc_kernel total: %d launches, %.3f ms (%.1f%% of the listed time), DRAM read %.1f MB + write %.1f MB = %.1f MB per launch on average"
              % (n, sum(a[1] for a in tc) / 1e6, 100 * sum(a[1] for a in tc) / total, sum(a[2] for a in tc) / 1e6, sum(a[3] for a in tc) / 1e6,
                 (sum(a[2] for a in tc) + sum(a[3] for a in tc)) / 1e6 / n))


if __name__ == "__main__":
    main(sys.argv[1])
