mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_dg.py tests/test_gpu_dn.py tests/test_golden.py -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -5
DG_TIME_NO_SOLVE=1 timeout 100 python profiles/dg_time.py 9 | tail -2
DG_TIME_NO_SOLVE=1 DG_TIME_REPS=3 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dg_residual -c 4 --csv --log-file gpurun_out/dg_res.csv python profiles/dg_time.py 9 > /dev/null 2>&1
echo "k_dg_residual ns: $(grep k_dg_residual gpurun_out/dg_res.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
