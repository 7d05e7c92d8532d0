#!/bin/bash
# round 2, capture d: FD coupling kernel v2 (plan tables) - parity subset, serial phase times, ncu of rows + fd kernels
mkdir -p gpurun_out
export MDB_BENCH_NO_UNSTRUCTURED=1
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
MDB_JAC_SERIAL=1 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2d.json 2> gpurun_out/r2d.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r2d.json"))
print("serial ms/step %.4f"%d["ms_per_step"], d["extra"]["jacobian_phases_ms"])
P
python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2d2.json 2> gpurun_out/r2d2.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r2d2.json"))
print("overlap ms/step %.4f"%d["ms_per_step"], d["extra"]["jacobian_phases_ms"])
P
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_fd_coupling|k_jacobian_rows_q1" --launch-skip 2 --launch-count 2 -o gpurun_out/r2d_k -f python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2d_ncu1.log 2>&1
