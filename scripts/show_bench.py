import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fit       %.0f evals/s  frac %.4f  e2e %.0f" % (l["value"], l["roofline"]["frac"], l["e2e"]["value"]))
if "posterior" in l:
    p = l["posterior"]
    print("posterior %.4g points/s  frac %.4f  %.1f ms/step  e2e %.4g" % (p["value"], p["roofline"]["frac"], p["ms_per_step"], p["e2e"]["value"]))
if "cpu_baseline" in l:
    print("cpu", l["cpu_baseline"])
