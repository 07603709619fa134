import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:round(d[k],2) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","gpu_launches")}, {k:round(d["roofline"][k],3) for k in ("achieved","frac","kernel_ms_per_step","clear_ms_per_step","layout_ms_per_step")})
if d.get("e2e"): print("   e2e", {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k!="what"})
