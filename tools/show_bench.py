#!/usr/bin/env python
"""tools/show_bench.py FILE -- compact view of a bench.py JSON line (headline, legs, alignment, neighbours)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("headline", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"),
      "frac", d["roofline"]["frac"], "clocks", d.get("clocks"))
print("parity", {k: v for k, v in d["parity"].items() if "marginal" in k or k == "gate"})
if "cpu_baseline" in d:
    print("cpu_baseline", d["cpu_baseline"])
for k, v in d.get("legs", {}).items():
    items = [(a, b) for a, b in v.items() if isinstance(b, dict)] if k in ("cfg3", "cfg2_in_flight") else [(k, v)]
    for kk, vv in items:
        keep = ("us_per_launch", "us_per_batch", "ms_per_step", "ms_parse", "ms_parse_plus_gather", "value", "exposed_allreduce_us_per_step", "sentences_per_s")
        print(k, kk, {a: b for a, b in vv.items() if a in keep}, "frac", vv.get("roofline", {}).get("frac"),
              "marg", vv.get("parity", {}).get("marginal_max_abs_vs_f64"))
a = d.get("alignment")
if a:
    print("align ms", a["ms"], "frac", a["roofline"]["frac"], "reduced", a["reduced"]["ms"], "bwd", a["backward"]["ms"],
          {k: v for k, v in a.items() if k not in ("roofline", "reduced", "backward", "parity", "workload")})
print("neighbours", json.dumps(d.get("neighbours"), indent=1))
