"""Print the fields of a bench.py JSON line that matter when reading a run by eye.  usage: python tools/show_bench.py file.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.strip().startswith("{")][-1])
if "unavailable" in d:
    print(d); sys.exit(0)
print("value", round(d["value"], 1), d["unit"], "| ms/step", round(d["ms_per_step"], 2), "| e2e", round((d.get("e2e") or {}).get("value", 0), 1))
print("kernel ms/step", d.get("kernel_ms_per_step"))
for k in ("roofline", "roofline_topk"):
    r = d.get(k)
    if r:
        print(k, r["kernel"][:40], "achieved", round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3),
              "| per search", round(r.get("frac_per_search_incl_select", 0), 3), "| hw_frac", round(r.get("hw_frac", 0), 3))
print("parity", d.get("parity"))
print("cpu", (d.get("cpu_baseline") or {}).get("value"), "| launches", d.get("gpu_launches"), "| clocks", d.get("clocks"))
