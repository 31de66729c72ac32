#!/usr/bin/env python3
"""profiles/<name>: DRAM bytes per launch of the proposal kernel from an ncu_summary.py JSON of the bench command
(the full-size launches of 32 proposals), in the form bench.py's `roofline.traffic` reads.

    python scripts/make_traffic.py <ncu_score.json> <out.json> "<source note>"
"""
import json
import sys

recs = [r for r in json.load(open(sys.argv[1])) if "score_proposals" in r["kernel"]]
full = [r for r in recs if r.get("grid", "").replace(" ", "").endswith(",32,1)")] or recs
b = [r["dram_read_bytes"] + r["dram_write_bytes"] for r in full]
json.dump({"dram_bytes_per_launch": sum(b) / len(b), "launches_averaged": len(b), "kernel": full[0]["kernel"].replace("void ", ""),
           "grid": full[0].get("grid"), "source": sys.argv[3]}, open(sys.argv[2], "w"), indent=1)
print(open(sys.argv[2]).read())
