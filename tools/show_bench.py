#!/usr/bin/env python
"""Print the interesting numbers of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench.log"""
import json, sys
for path in sys.argv[1:]:
    line = [l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print("==", path)
    if "stages" in d:
        print("value %.1fM frames/s  ms/step %.3f  n_gpus %d" % (d["value"] / 1e6, d["ms_per_step"], d["n_gpus"]))
        for k in ("e2e", "e2e_packed", "e2e_device_pack"):
            if k in d:
                print("  %-16s %.1fM  %s" % (k, d[k]["value"] / 1e6, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in d[k].items() if kk not in ("api", "value", "unit")}))
        for k, v in d["stages"].items():
            if isinstance(v, dict):
                print("  stage %-24s %.4f ms  frac %.3f %s" % (k, v["ms"], v.get("frac", 0), ("tflops %.0f" % v["achieved_tflops"]) if "achieved_tflops" in v else ("gbs %.0f" % v.get("achieved_gbs", 0))))
        r = d["roofline"]
        print("  roofline %.1f TF frac %.3f (sustained %.3f)" % (r["achieved"], r["frac"], r.get("frac_sustained", 0)))
        if "cpu_baseline" in d:
            print("  cpu", round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
        print("  clocks", d.get("clocks"))
    for name, e in (d.get("extra") or {}).items():
        show = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in e.items() if k in ("ms_per_iter", "estep_ms", "mstep_ms", "finalize_ms", "allreduce_ms", "ms_per_step", "value", "error", "loglik")}
        print("  extra.%s %s" % (name, show), (e.get("roofline") or {}).get("frac"), (e.get("cpu_baseline") or {}).get("value"))
    if "ms_per_iter" in d:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ("ms_per_iter", "estep_ms", "mstep_ms", "finalize_ms", "allreduce_ms", "n_gpus")})
