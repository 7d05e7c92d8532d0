#!/bin/bash
# round 2, capture h (N GPUs): multi-GPU parity tests, weak-scaling bench with / without the halo overlap, strong-scaling point
N=${1:-2}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 100 > gpurun_out/r2h_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2h_pytest_n$N.log
run() { tag=$1; shift; timeout 150 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --quick --steps 10 $EXTRA > gpurun_out/r2h_${tag}_n$N.json 2> gpurun_out/r2h_${tag}_n$N.err; echo "$tag rc=$?"; python - gpurun_out/r2h_${tag}_n$N.json <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d["extra"]
    print(" ms/step %.4f value %.3e scaling %s"%(d["ms_per_step"], d["value"], d["scaling"]))
    print(" spmv", round(e["spmv"]["ms"],4), "cg it/s", round(e["cg"]["iters_per_s"],1), "bicg it/s", round(e["bicgstab"]["iters_per_s"],1))
    print(" solve_check", {k:v for k,v in e["solve_check"].items() if k!="method"})
    print(" partition_check", e["partition_check"])
except Exception as ex:
    print(" FAILED", ex)
P
}
EXTRA=""
run weak X=1
run weak_noovl MDB_HALO_NO_OVERLAP=1
EXTRA="--strong"
run strong X=1
