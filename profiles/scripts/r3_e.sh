#!/bin/bash
# round 2, session 3, capture e: full-set ncu of the marching residual kernel and of the Qk streaming fill; launch list of the bench step
cd "$(dirname "$0")/../.." || exit 1
python profiles/prof_res.py 256 3 > gpurun_out/r3e_plain.log 2>&1 || { cat gpurun_out/r3e_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"k_residual_march|k_residual_volume_q1_jobs|k_lambda_boundary|k_coupling_residual" -s 4 -c 4 -o gpurun_out/r3e_full_res -f python profiles/prof_res.py 256 3 > gpurun_out/r3e_ncu1.log 2>&1
python profiles/qk_time.py 2048 > gpurun_out/r3e_qk_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_qk_fill|k_jacobian_rows_qk_list" -s 6 -c 3 -o gpurun_out/r3e_full_qk -f python profiles/qk_time.py 1024 > gpurun_out/r3e_ncu2.log 2>&1
python profiles/prof_driver.py 256 2 > gpurun_out/r3e_drv_plain.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r3e_launches.csv python profiles/prof_driver.py 256 2 > gpurun_out/r3e_ncu3.log 2>&1
ls -la gpurun_out/r3e_*
