import json, sys
line = [l for l in open(sys.argv[1]) if l.startswith("{")][-1]
d = json.loads(line)
print(sys.argv[1], {k: round(d[k], 2) if isinstance(d[k], float) else d[k] for k in ("value", "ms_per_step", "gpu_launches")},
      {k: round(d["roofline"][k], 3) for k in ("achieved", "frac", "kernel_ms_per_step", "build_kernels_ms_per_step", "clear_ms_per_step", "layout_ms_per_step") if k in d["roofline"]})
if d.get("e2e"):
    print("   e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "what"})
