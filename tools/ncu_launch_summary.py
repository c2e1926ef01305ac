"""Summarises an ncu launch list (the `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv` pass of /opt/skills/guides/B200_PROFILING.md) into the per-kernel table kept under profiles/
and the conv-traffic figure bench.py reports as roofline.traffic.

usage: python tools/ncu_launch_summary.py profiles/r01_launches.csv "header line" [--traffic-json profiles/r01_conv_traffic.json]
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def parse(path):
    launches = OrderedDict()           # id -> dict(name, time_ns, rd, wr)
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(fh) if len(r) >= 15]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    col = {h: i for i, h in enumerate(rows[hdr])}
    for r in rows[hdr + 1:]:
        L = launches.setdefault(r[col["ID"]], dict(name=r[col["Kernel Name"]], time_ns=0.0, rd=0.0, wr=0.0))
        v = float(r[col["Metric Value"]].replace(",", ""))
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            L["time_ns"] = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[col["Metric Unit"]], 1.0)
        elif m == "dram__bytes_read.sum":
            L["rd"] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[col["Metric Unit"]], 1.0)
        elif m == "dram__bytes_write.sum":
            L["wr"] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[col["Metric Unit"]], 1.0)
    return list(launches.values())


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("rd::", "")


def main():
    path, header = sys.argv[1], sys.argv[2]
    launches = parse(path)
    total = sum(L["time_ns"] for L in launches)
    by = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for L in launches:
        b = by[short(L["name"])]
        b[0] += 1
        b[1] += L["time_ns"]
        b[2] += L["rd"]
        b[3] += L["wr"]
    print(f"# {header}")
    print(f"# total {total / 1e6:.3f} ms over {len(launches)} launches (cold-cache, serialised: compare SHARES, not absolutes)\n")
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s} {'dram_rd_MB/launch':>18s} {'dram_wr_MB/launch':>18s}")
    for k, (n, t, rd, wr) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {n:8d} {t / 1e3:10.1f} {100 * t / total:6.1f}% {t / 1e3 / n:8.2f} {rd / n / 1e6:18.2f} {wr / n / 1e6:18.2f}")
    conv = [L for L in launches if "conv_fprop_kernel" in L["name"] or "conv_wgrad_kernel" in L["name"]]
    if conv:
        share = sum(L["time_ns"] for L in conv) / total
        traffic = sum(L["rd"] + L["wr"] for L in conv) / len(conv)
        print(f"\n# tcgen05 conv programs: {len(conv)} launches, {100 * share:.1f} % of the profiled time, "
              f"{traffic / 1e6:.1f} MB DRAM traffic per launch")
        if "--traffic-json" in sys.argv:
            out = sys.argv[sys.argv.index("--traffic-json") + 1]
            with open(out, "w") as fh:
                json.dump({"source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum, cold cache per launch)",
                           "conv_launches": len(conv), "traffic_bytes_per_launch": traffic, "conv_share_of_step_ncu": share}, fh, indent=1)


if __name__ == "__main__":
    main()
