"""Dev tool: one line per bench JSON file (ms per step, per-kernel ms / roofline fraction)."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], "launches/step %.0f" % (d["gpu_launches"] / d["steps"]),
              {k.replace("_kernel", ""): (round(v["ms_per_step"], 3), round(v["frac"], 3)) for k, v in d["kernels"].items()},
              "sweep f+b %.3f ms frac %.3f" % (d["level_sweep"]["ms_fwd_bwd"], d["level_sweep"]["frac"]))
    except Exception as e:
        print(f, "unreadable:", e)
