#!/bin/bash
# round 2, session 3, capture b (N GPUs): multi-GPU parity tests with the peer-memory scalar all-reduce, weak-scaling bench with the
# peer windows against NCCL scalars (MDB_SCALARS_NCCL=1)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 250 > gpurun_out/r3b_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3b_pytest_n$N.log
run() { tag=$1; shift; timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --quick --steps 10 $EXTRA > gpurun_out/r3b_${tag}_n$N.json 2> gpurun_out/r3b_${tag}_n$N.err; echo "$tag rc=$?"; python - gpurun_out/r3b_${tag}_n$N.json <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d["extra"]
    print(" ms/step %.4f value %.3e scaling %s"%(d["ms_per_step"], d["value"], d["scaling"]))
    print(" spmv", round(e["spmv"]["ms"],4), "cg it/s", round(e["cg"]["iters_per_s"],1), "bicg it/s", round(e["bicgstab"]["iters_per_s"],1))
    print(" solve_check", {k:v for k,v in e["solve_check"].items() if k!="method"})
    print(" tts", e.get("time_to_solution"))
except Exception as ex:
    print(" FAILED", ex)
P
}
EXTRA=""
run ring X=1
run nccl MDB_SCALARS_NCCL=1
