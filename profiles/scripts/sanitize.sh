#!/bin/bash
# compute-sanitizer passes over small parity cases (the SSOR / ILU0 sweeps, the halo flags and the bulk-async copies are
# hand-rolled acquire/release protocols).  memcheck on the assembly + solver + Stokes-Darcy cases, racecheck (shared memory hazards)
# on the staged kernels; synccheck on the mbarrier / bulk-copy kernels.  Output: gpurun_out/sanitize_<tool>.log, summary at the end.
# Run on the GPU box:  bash profiles/scripts/sanitize.sh
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, log tag, pytest selection...
  tool=$1; tag=$2; shift 2
  timeout 900 $SAN --tool $tool --error-exitcode 99 --print-limit 20 --target-processes all \
    python -m pytest -x -q -p no:cacheprovider "$@" > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "$tool $tag rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/sanitize_${tool}_${tag}.log) clean summaries; $(grep -h 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_${tag}.log | sort | uniq -c | tr '\n' ';')  $(tail -1 gpurun_out/sanitize_${tool}_${tag}.log)"
}
run memcheck parity tests/test_gpu_parity.py -k "jacobian_entries or residual_entries or spmv or one_preconditioner_application"
run memcheck precond tests/test_gpu_parity.py -k "sequential_preconditioners_match_oracle"
run memcheck sd tests/test_gpu_sd.py -k "sd_6x4 or golden"
run memcheck um tests/test_gpu_um.py -k "um_6x5 or um_cvcf"
run racecheck parity tests/test_gpu_parity.py -k "jacobian_entries or residual_entries or spmv"
run racecheck um tests/test_gpu_um.py -k "um_6x5"
run synccheck parity tests/test_gpu_parity.py -k "jacobian_entries or residual_entries or spmv"
