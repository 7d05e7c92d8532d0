"""DRAM traffic per launch of the hot kernels, measured on THIS build: runs `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`
over profiles/prof_driver.py (the M2 problem, every hot operation twice) in a child process and prints ONE JSON line

    {"<kernel>": <mean dram bytes per launch>, ...}

bench.py calls it after its timed regions (roofline.traffic); nothing timed ever runs under the profiler.

    python profiles/measure_traffic.py [--size 256] [--json profiles/traffic.json]
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = "k_jacobian_fill|k_jacobian_rows|k_fd_coupling|k_coupling_jacobian|k_residual_|k_spmv_|k_cg_update|k_dirichlet_rows"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        log = os.path.join(tmp, "launches.csv")
        cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
               "--kernel-name", "regex:" + KERNELS, "--launch-count", "40", "--csv", "--log-file", log,
               sys.executable, os.path.join(ROOT, "profiles", "prof_driver.py"), str(a.size), "1"]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=220)
        if p.returncode != 0 or not os.path.exists(log):
            print("ncu failed: " + (p.stderr or p.stdout)[-300:], file=sys.stderr)
            return 1
        hdr, per = None, defaultdict(lambda: defaultdict(float))
        for r in csv.reader(open(log, errors="ignore")):
            if r and r[0] == "ID":
                hdr = r
                continue
            if hdr is None or len(r) < len(hdr):
                continue
            rec = dict(zip(hdr, r))
            name = rec["Kernel Name"].replace("void ", "").split("(")[0].split("<")[0].replace("mdb::", "")
            per[(name, rec["ID"])][rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
        agg = defaultdict(list)
        for (name, _), m in per.items():
            agg[name].append(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0))
        out = {k: sum(v) / len(v) for k, v in agg.items()}
        print(json.dumps(out))
        if a.json:
            json.dump({k: {"dram_bytes_per_launch": v, "launches_captured": len(agg[k])} for k, v in out.items()}, open(a.json, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
