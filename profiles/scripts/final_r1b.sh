# final checks and DG captures of the round
mkdir -p gpurun_out
timeout -s KILL 700 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -5 | tee gpurun_out/pytest_full_r1f.log
DG_TIME_NO_SOLVE=1 DG_TIME_REPS=2 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dg_r1f.csv python profiles/dg_time.py 9 > gpurun_out/ncu_dg_r1f.log 2>&1; echo "ncu dg list rc=$?"
DG_TIME_NO_SOLVE=1 DG_TIME_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_dg_jacobian|k_dg_residual|k_dg_face" -s 3 -c 3 -o gpurun_out/prof_dg_r1f -f python profiles/dg_time.py 9 > gpurun_out/ncu_dg_full.log 2>&1; echo "ncu dg full rc=$?"
