import json,sys
for p in sys.argv[1:]:
    d=json.loads(open(p).read().splitlines()[-1])
    c=d["config"]
    print(p, "N=%d"%d["n_gpus"], d["scaling"], "%.2f G upd/s"%(d["value"]/1e9), "%.1f us/tick"%(d["ms_per_step"]*1e3), "launches", d["gpu_launches"])
    print("   host phases us:", {k: round(v,1) for k,v in (c.get("phase_us_rank0") or {}).items()})
    print("   kernels us/step:", c.get("kernel_us_per_step_rank0"))
    print("   e2e %.2f G"%(d["e2e"]["value"]/1e9), "owned", c.get("owned_per_rank"))
