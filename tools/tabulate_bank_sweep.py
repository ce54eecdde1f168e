#!/usr/bin/env python
"""Prints tools/sweep_bank_repeat.py's JSON as a table: schedule x stream count, us per
iteration launched / replayed from a CUDA graph."""
import collections
import json
import sys

d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep_bank_repeat.json"))
t = collections.defaultdict(dict)
for p in d["points"]:
    t[(p["schedule"], p["bank_repeat_variant"], p.get("ctas_per_sm"))][p["streams"]] = (p["us_per_iteration"], p["graph_us_per_iteration"])
Ss = sorted({p["streams"] for p in d["points"]})
print("schedule".ljust(28), *[str(s).rjust(14) for s in Ss])
for k, v in t.items():
    print(str(k).ljust(28), *[f"{v[s][0]:6.1f}/{v[s][1]:6.1f}".rjust(14) for s in Ss])
