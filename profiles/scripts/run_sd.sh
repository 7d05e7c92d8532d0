set -x
g++ -std=c++11 -I include examples/stokesdarcy2.cc -o /tmp/stokesdarcy2 -L dune-multidomain_b200 -lmdb200 -Wl,-rpath,$PWD/dune-multidomain_b200
/tmp/stokesdarcy2 examples/stokesdarcy2.ini 0 > gpurun_out/r2j_example_sd.log 2>&1; echo rc=$? >> gpurun_out/r2j_example_sd.log
/tmp/stokesdarcy2 examples/stokesdarcy2.ini 1 >> gpurun_out/r2j_example_sd.log 2>&1; echo rc=$? >> gpurun_out/r2j_example_sd.log
timeout 600 python profiles/sd_time.py 40 80 160 > gpurun_out/r2j_sd_time.txt 2>&1
tail -5 gpurun_out/r2j_example_sd.log; cat gpurun_out/r2j_sd_time.txt | cut -c1-900
