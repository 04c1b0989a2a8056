#!/usr/bin/env python
"""Markdown table of the multi-GPU bench lines (profiles/r2_scale_<tag>_n<N>.json): value, efficiency against the N=1 line of
the same tag measured on the same box (weak: value / (N x value_1); strong: the same formula -- total work is fixed, value is
the whole-job rate), e2e, clocks, gather cost.  usage: scripts/scale_table.py [dir]"""
import glob
import json
import os
import re
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "profiles"
lines = {}
for f in sorted(glob.glob(os.path.join(d, "r2_scale_*_n*.json"))):
    m = re.match(r"r2_scale_(.+)_n(\d+)\.json", os.path.basename(f))
    try:
        txt = [ln for ln in open(f).read().strip().splitlines() if ln.startswith("{")]
        lines[(m.group(1), int(m.group(2)))] = json.loads(txt[-1])
    except Exception as e:  # noqa: BLE001
        print(f"<!-- {f}: {e} -->")
print("| config | N | chains (total / per GPU) | grad-evals/s | vs N=1 | efficiency | e2e | ms/step | SM MHz | gather to rank 0 |")
print("|---|---|---|---|---|---|---|---|---|---|")
for (tag, n), L in sorted(lines.items()):
    base = lines.get((tag, 1))
    sp = L["value"] / base["value"] if base else None
    c = L["config"]
    e2e = (L.get("e2e") or {}).get("value")
    g = L.get("gather_to_rank0")
    print("| %s | %d | %s / %s | %.3g | %s | %s | %s | %.0f | %s | %s |" % (
        tag + " (" + L["scaling"] + ")", n, c.get("chains_total"), c.get("chains_per_gpu"), L["value"],
        "%.2fx" % sp if sp else "-", "%.2f" % (sp / n) if sp else "-", "%.3g" % e2e if e2e else "-", L["ms_per_step"],
        (L.get("clocks") or {}).get("sm_mhz"), "%.0f ms, %.1f GB, %.0f GB/s" % (g["ms"], g["bytes_to_rank0"] / 1e9, g["GB_per_s"]) if g else "-"))
