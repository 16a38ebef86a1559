"""print the headline fields of a bench.py JSON line: python tools/_pb.py file.json"""
import json
import sys
for f in sys.argv[1:]:
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "gemm_ms", round(d["roofline"]["gemm_ms_per_step"], 3),
          "frac", round(d["roofline"]["frac"], 3), "launches/step", d["gpu_launches"] / d["steps"], "MHz", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w"])
