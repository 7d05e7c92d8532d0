#!/bin/bash
# round 2, capture f: stream priority A/B for the Jacobian step
mkdir -p gpurun_out
export MDB_BENCH_NO_UNSTRUCTURED=1
run() { timeout 60 env "$@" python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2f.json 2> gpurun_out/r2f.err; python - "$*" <<'P'
import json,sys
try:
    d=json.load(open("gpurun_out/r2f.json"))
    print(sys.argv[1], "| ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["extra"]["jacobian_phases_ms"].items()}, "roof %.3f"%d["roofline"]["frac"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
}
run MDB_AUX_PRIO=1
run MDB_AUX_PRIO=1 MDB_FILL_NO_TMA=1
run MDB_AUX_PRIO=1 MDB_FILL_NO_TMA=1 MDB_JAC_ROWS_FIRST=1
run MDB_JAC_ROWS_FIRST=1
run MDB_AUX_PRIO=1 MDB_JAC_ROWS_FIRST=1
