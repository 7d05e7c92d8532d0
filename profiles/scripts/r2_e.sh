#!/bin/bash
# round 2, capture e: record rows kernel + merged fill + FD coupling v3 — parity subset, serial and overlapped phase times, fill variants
mkdir -p gpurun_out
export MDB_BENCH_NO_UNSTRUCTURED=1
timeout 60 python __graft_entry__.py smoke > gpurun_out/r2e_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 gpurun_out/r2e_smoke.log
[ $rc -ne 0 ] && exit 1
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_mortar.py tests/test_gpu_dn.py -m gpu -x -q --timeout 60 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
run() { timeout 60 env "$@" python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2e.json 2> gpurun_out/r2e.err; python - "$*" <<'P'
import json,sys
try:
    d=json.load(open("gpurun_out/r2e.json"))
    print(sys.argv[1], "| ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["extra"]["jacobian_phases_ms"].items()}, "roof %.3f"%d["roofline"]["frac"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
}
run MDB_JAC_SERIAL=1
run MDB_JAC_SERIAL=1 MDB_FILL_NO_TMA=1
run X=1
run MDB_FILL_NO_TMA=1
run MDB_FILL_NO_TMA=1 MDB_FILL_BLOCKS_PER_SM=12
run MDB_FILL_TMA_BLOCKS_PER_SM=3
