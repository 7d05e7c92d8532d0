mkdir -p gpurun_out
for w in 8 4 2; do
  MDB_DG_WPB=$w DG_TIME_NO_SOLVE=1 DG_TIME_REPS=3 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dg_jacobian -c 6 --csv --log-file gpurun_out/dg_wpb$w.csv python profiles/dg_time.py 9 > /dev/null 2>&1
  echo "wpb=$w: $(grep k_dg_jacobian gpurun_out/dg_wpb$w.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
