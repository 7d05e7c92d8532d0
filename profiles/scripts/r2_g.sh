#!/bin/bash
# round 2, capture g (1 GPU): full bench line with the new self-checks, e2e solve and the ncu traffic pass
mkdir -p gpurun_out
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 420 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2g_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r2g_bench.json"))
e=d["extra"]
print("ms/step %.4f value %.3e"%(d["ms_per_step"], d["value"]))
print("roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k not in ("note","step")})
print("step", d["roofline"]["step"]["frac"], d["roofline"]["step"]["frac_of_8TBs"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
for k in ("residual_assembly","spmv","cg","bicgstab","solve_check","e2e_solve"):
    print(k, e.get(k))
print("cpu", d.get("cpu_baseline"))
print("um", {k:v for k,v in e.get("unstructured",{}).items() if k!="cpu_port"})
P
