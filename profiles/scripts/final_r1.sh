# final captures of the round: bench line, ncu launch list of the same command, full-set capture of the top kernels, DG launch list
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r1f.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1f_ref.json 2> gpurun_out/bench_r1f_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_r1f.log 2>&1; echo "ncu list rc=$?"
DG_TIME_NO_SOLVE=1 DG_TIME_REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dg_r1f.csv python profiles/dg_time.py 9 > gpurun_out/ncu_dg_r1f.log 2>&1; echo "ncu dg rc=$?"
