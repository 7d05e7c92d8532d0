"""Summarise an `ncu --set full` report (.ncu-rep) as a markdown table of the metrics the roofline discussion uses.

    python profiles/ncu_report.py gpurun_out/prof.ncu-rep > profiles/rNN_full.md
    python profiles/ncu_report.py gpurun_out/prof.ncu-rep --json profiles/traffic.json   # per-kernel DRAM traffic for bench.py

Needs the `ncu` binary (present in the build container; no GPU required to read a report).
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st requests"),
]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    u = unit.lower()
    return f * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)


def to_us(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(unit.lower(), 1.0)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        e = {"kernel": name}
        for m, label in METRICS:
            if m in idx:
                e[label] = (r[idx[m]], units[idx[m]])
        out.append(e)
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        agg = {}
        for e in out:
            k = e["kernel"].split("<")[0]
            b = to_bytes(*e["DRAM rd"]) + to_bytes(*e["DRAM wr"])
            a = agg.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
            a["launches"] += 1; a["dram_bytes"] += b; a["us"] += to_us(*e["time"])
        res = {k: {"dram_bytes_per_launch": v["dram_bytes"] / v["launches"], "us_per_launch_under_ncu": v["us"] / v["launches"],
                   "launches_captured": v["launches"], "report": rep.split("/")[-1]} for k, v in agg.items()}
        json.dump(res, open(path, "w"), indent=1, sort_keys=True)
        return
    print("| kernel | grid x block | regs | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | SM % | occupancy % | warp inst | st sectors/request |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for e in out:
        us = to_us(*e["time"])
        rd, wr = to_bytes(*e["DRAM rd"]), to_bytes(*e["DRAM wr"])
        sr = float(e["st requests"][0].replace(",", "")) if "st requests" in e else 0.0
        ss = float(e["st sectors"][0].replace(",", "")) if "st sectors" in e else 0.0
        print("| %s | %s x %s | %s | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.3g | %s |" % (
            e["kernel"], e["grid"][0], e["block"][0], e["regs"][0], us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3,
            float(e["DRAM %"][0]), float(e["SM %"][0]), float(e["occupancy %"][0]), float(e["warp inst"][0].replace(",", "")),
            "%.1f" % (ss / sr) if sr else "-"))


if __name__ == "__main__":
    main()
