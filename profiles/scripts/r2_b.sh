#!/bin/bash
# round 2, capture b: stand-alone times of the Jacobian step's kernels (serial switch) + ncu launch list
mkdir -p gpurun_out
for v in "X=1" "MDB_FD_STAGED=1" "MDB_FILL_NO_TMA=1"; do
  env MDB_BENCH_NO_UNSTRUCTURED=1 MDB_JAC_SERIAL=1 $v python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2b.json 2> gpurun_out/r2b.err
  python - "$v" <<'P'
import json,sys
d=json.load(open("gpurun_out/r2b.json"))
print("serial", sys.argv[1], "ms/step %.4f"%d["ms_per_step"], d["extra"]["jacobian_phases_ms"])
P
done
MDB_BENCH_NO_UNSTRUCTURED=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2b_ncu.log 2>&1
python profiles/ncu_launches.py gpurun_out/r2b_launches.csv 2>/dev/null | head -40
