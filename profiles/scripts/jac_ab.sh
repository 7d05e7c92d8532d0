#!/bin/bash
# A/B of the Jacobian step's stream plan: fill kernel variant x priority of the auxiliary stream
cd "$(dirname "$0")/../.." || exit 1
for cfg in ${JAC_AB_CFGS:-"base:" "auxprio:MDB_AUX_PRIO=1" "notma:MDB_FILL_NO_TMA=1" "notma_auxprio:MDB_FILL_NO_TMA=1,MDB_AUX_PRIO=1" "notma_auxprio_fd1:MDB_FILL_NO_TMA=1,MDB_AUX_PRIO=1,MDB_FD_BLOCKS_PER_SM=1" "notma_auxprio_fd2:MDB_FILL_NO_TMA=1,MDB_AUX_PRIO=1,MDB_FD_BLOCKS_PER_SM=2" "notma_auxprio_fd3:MDB_FILL_NO_TMA=1,MDB_AUX_PRIO=1,MDB_FD_BLOCKS_PER_SM=3" "tma_fd2:MDB_FD_BLOCKS_PER_SM=2" "tma_auxprio_fd2:MDB_AUX_PRIO=1,MDB_FD_BLOCKS_PER_SM=2"}; do
  name=${cfg%%:*}; envs=$(echo ${cfg#*:} | tr "," " ")
  env $envs MDB_BENCH_NO_UNSTRUCTURED=1 python bench.py --gpus 1 --steps 20 --warmup 3 --quick --no-cpu-baseline > gpurun_out/jacab_$name.json 2> gpurun_out/jacab_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/jacab_$name.json"))
print("$name", "ms/step %.4f" % d["ms_per_step"], {k: round(v,3) for k,v in d["extra"]["jacobian_phases_ms"].items()})
PY
done
