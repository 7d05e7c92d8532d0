mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -8 | tee gpurun_out/pytest_full_r1g.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r1g.json; tail -3 gpurun_out/bench_r1g.err
