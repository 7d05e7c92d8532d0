import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    x=d["extra"]
    print(f.split('/')[-1], "jac ms %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in x["jacobian_phases_ms"].items()}, "roof %.3f"%d["roofline"]["frac"],
          "res ms %.3f ker %.3f"%(x["residual_assembly"]["ms"], x["residual_assembly"]["kernel_ms"]), "spmv ms %.3f frac %.3f"%(x["spmv"]["ms"],x["spmv"]["frac"]),
          "cg ms/it %.3f frac %.3f"%(x["cg"]["ms_per_iter"], x["cg"]["frac"]), "bicg it/s %.1f"%x["bicgstab"]["iters_per_s"], "e2e ms %.3f"%d["e2e"]["ms_per_step"])
    for k in ("unstructured", "unstructured_partition", "stokes_darcy", "time_to_solution"):
        if k in x:
            print("  ", k, json.dumps(x[k])[:1500])
