#!/bin/bash
# ncu captures of the unstructured kernels: launch list of the whole driver + full-set capture of the assembly kernels
cd "$(dirname "$0")/../.." || exit 1
TAG=${1:-um}
python profiles/prof_um.py 1024 2 > gpurun_out/${TAG}_plain.log 2>&1 || { cat gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python profiles/prof_um.py 1024 2 > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_um_jacobian|k_um_residual|k_um_coupling" -s 4 -c 8 -o gpurun_out/${TAG}_full -f python profiles/prof_um.py 1024 2 > gpurun_out/${TAG}_ncu2.log 2>&1
MDB_UM_JAC_CELLS=1 ncu --set full --clock-control none --import-source on -k regex:"k_um_jacobian_cells" -s 1 -c 1 -o gpurun_out/${TAG}_full_cells -f python profiles/prof_um.py 1024 2 > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out/${TAG}_*
