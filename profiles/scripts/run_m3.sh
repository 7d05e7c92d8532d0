#!/bin/bash
# BASELINE configs[2] at full size through the C++ facade: 4096^2 cells, 200 implicit Euler steps, BiCGSTAB with the multigrid V-cycle
cd "$(dirname "$0")/../.." || exit 1
g++ -O2 -std=c++11 -I include examples/testinstationarypoisson.cc -o /tmp/tip -L dune-multidomain_b200 -lmdb200 -Wl,-rpath,$PWD/dune-multidomain_b200 || exit 1
{ echo "# examples/testinstationarypoisson 12 2.0 0.0002 0.04 2 euler mg   (host vectors: every step uploads u_old and downloads u_new through the facade)"; time /tmp/tip 12 2.0 0.0002 0.04 2 euler mg; } > gpurun_out/r2r_instationary_M3_mg.txt 2>&1
cat gpurun_out/r2r_instationary_M3_mg.txt
