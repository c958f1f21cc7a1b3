import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], round(d["value"]/1e9,2),"G upd/s", round(d["ms_per_step"]*1e3,1),"us/tick", "launches",d["gpu_launches"])
print("  "+"  ".join("%s=%.1f"%(k["name"],k["avg_us"]) for k in d["kernels"]))
