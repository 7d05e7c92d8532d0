#!/bin/bash
# round 2, capture a: parity suite + Jacobian step with the staged / register FD coupling kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
MDB_BENCH_NO_UNSTRUCTURED=1 MDB_FD_STAGED=1 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2a_bench_staged.json 2> gpurun_out/r2a_bench_staged.err
MDB_BENCH_NO_UNSTRUCTURED=1 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2a_bench_fast.json 2> gpurun_out/r2a_bench_fast.err
MDB_BENCH_NO_UNSTRUCTURED=1 MDB_FILL_NO_TMA=1 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2a_bench_fast_notma.json 2> gpurun_out/r2a_bench_fast_notma.err
python - <<'P'
import json
for f in ("staged","fast","fast_notma"):
    try:
        d=json.load(open("gpurun_out/r2a_bench_%s.json"%f))
        print(f, "ms/step %.4f"%d["ms_per_step"], d["extra"]["jacobian_phases_ms"], "roof %.3f"%d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
P
