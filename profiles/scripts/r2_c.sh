#!/bin/bash
# round 2, capture c: ncu full set of the register FD coupling kernel + launch list of the Jacobian step's kernels
mkdir -p gpurun_out
export MDB_BENCH_NO_UNSTRUCTURED=1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_fd_coupling --launch-count 1 -o gpurun_out/r2c_fd -f python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2c_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --kernel-name regex:"k_jacobian|k_fd_|k_coupling|k_element|k_dirichlet" -c 24 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2c_ncu2.log 2>&1
python profiles/ncu_launches.py gpurun_out/r2c_launches.csv 2>/dev/null | tail -12
