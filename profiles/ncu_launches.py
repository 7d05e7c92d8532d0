"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
one line per launch (device time, DRAM bytes) and the share of every kernel in the whole list.

    python profiles/ncu_launches.py gpurun_out/launches.csv [--only mdb::] > profiles/rNN_launches.md
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def load(path):
    hdr, d = None, OrderedDict()
    for r in csv.reader(open(path, errors="ignore")):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        rec = dict(zip(hdr, r))
        k = int(rec["ID"])
        e = d.setdefault(k, {"name": rec["Kernel Name"], "grid": rec["Grid Size"], "block": rec["Block Size"]})
        e[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    return d


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


def main():
    path = sys.argv[1]
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    d = load(path)
    print("| # | kernel | grid | block | us | DRAM rd MB | DRAM wr MB | GB/s |")
    print("|---|---|---|---|---|---|---|---|")
    tot = defaultdict(lambda: [0, 0.0, 0.0])
    total = 0.0
    for k, e in d.items():
        n = short(e["name"])
        if only and only not in n:
            continue
        us = e.get("gpu__time_duration.sum", 0.0) / 1e3
        rd = e.get("dram__bytes_read.sum", 0.0)
        wr = e.get("dram__bytes_write.sum", 0.0)
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0.0
        print("| %d | %s | %s | %s | %.1f | %.1f | %.1f | %.0f |" % (k, n, e["grid"], e["block"], us, rd / 1e6, wr / 1e6, gbs))
        t = tot[n]
        t[0] += 1; t[1] += us; t[2] += rd + wr
        total += us
    print()
    print("| kernel | launches | total us | share | mean us | mean DRAM MB |")
    print("|---|---|---|---|---|---|")
    for n, (cnt, us, by) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% | %.1f | %.1f |" % (n, cnt, us, 100.0 * us / total if total else 0.0, us / cnt, by / cnt / 1e6))


if __name__ == "__main__":
    main()
