import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d.get("kernels",{})
        print(f.split("/")[-1], "tok/s=%.0f ms=%.2f e2e=%.0f launches=%s roof=%.3f side=%.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"] or 0, k.get("rank_r_side_path_ms_per_step",0)), "eager_ref=", (d.get("gpu_eager_baseline") or {}).get("value"), (d.get("gpu_eager_baseline") or {}).get("speedup_of_this_build"))
    except Exception as e:
        print(f, "ERR", e)
