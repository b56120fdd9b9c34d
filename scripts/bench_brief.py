"""prints the essentials of a bench.py JSON line (the last line of the file that starts with '{')"""
import json
import sys

txt = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if not txt:
    print("no JSON line in", sys.argv[1])
    sys.exit(0)
j = json.loads(txt[-1])
print("N=%s value %.4f s  e2e %.4f s  iterations %s  true residual %.3e  launches %s" % (j["n_gpus"], j["value"], j["e2e"]["value"], j.get("iterations"), j.get("final_true_rel_residual", float("nan")), j.get("gpu_launches")))
r = j["roofline"]
print("roofline %s: %.0f GB/s = %.3f of %.0f (nominal %.3f), traffic %s" % (r["kernel"], r["achieved"], r["frac"], r["peak"], r.get("frac_nominal", float("nan")), r.get("traffic")))
if j.get("mg_setup"):
    print("mg_setup", {k: round(v, 3) for k, v in j["mg_setup"].items() if k != "how"})
for k, v in sorted(j["kernels"].items(), key=lambda kv: -kv[1]["share"]):
    print("   %-20s share %.3f  %9.1f us x%-5d %s" % (k, v["share"], v["ms_per_launch"] * 1e3, v["launches"], None if v["GBps"] is None else round(v["GBps"])))
print("host_side", j.get("host_side"), "clocks", j.get("clocks"))
for key in ("parity", "cpu_baseline"):
    if key in j:
        print(key, json.dumps(j[key])[:700])
for name, o in (j.get("other_workloads") or {}).items():
    print("other", name, json.dumps({k: v for k, v in o.items() if k not in ("kernels", "desc")})[:900])
