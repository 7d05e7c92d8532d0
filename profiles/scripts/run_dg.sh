python -m pytest tests/test_gpu_dn.py -x -q -m gpu -k driver 2>&1 | grep -v "^E   *+" | tail -30
d=$(mktemp -d)
g++ -std=c++11 -O2 -I include examples/testpoisson-dirichlet-neumann.cc -o $d/tdn -L dune-multidomain_b200 -lmdb200 -Wl,-rpath,$PWD/dune-multidomain_b200
sed 's/refine = 6/refine = 9/' examples/poisson-dirichlet-neumann.ini > $d/r9.ini
( time $d/tdn $d/r9.ini dg2 dg2 ) > gpurun_out/dn_dg2_refine9.txt 2>&1
tail -5 gpurun_out/dn_dg2_refine9.txt
