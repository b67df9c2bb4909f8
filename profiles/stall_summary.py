#!/usr/bin/env python3
"""Condense `ncu --page source --csv` exports (gpurun_out/r2_source_*.csv: per-instruction warp-stall samples, too large to
commit) into profiles/r2_stall_samples.md: stall-reason shares, shares by opcode, the hottest instructions and the executed
local-memory instructions of the last captured launch of each kernel."""
import collections
import csv
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
md = ["# Round 2 - warp-stall samples per instruction (`ncu --set full --import-source on`, source page)", "",
      "Regenerate with `python profiles/stall_summary.py` from `gpurun_out/r2_source_*.csv` (made by `profiles/final_run_ncu.sh` /",
      "`final_run_cfgs.sh`; o_proj-size tensor, last captured launch of each kernel).", ""]
for fn, title in (("r2_source_stats_fast.csv", "stats_fast_kernel<true, false> (last tile-row range of o_proj, 3584 CTAs)"),
                  ("r2_source_greedy_init.csv", "greedy_init_kernel<true> (third range)"),
                  ("r2_source_greedy_par.csv", "greedy_par_kernel<true, false> (all passes, 16-CTA cluster, 128 registers)")):
    f = G / fn
    if not f.exists():
        continue
    rows = list(csv.reader(open(f)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = heads[-1]
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith("stall") and "Not Issued" not in h]
    tot, by_op, inst_op, lines = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), []
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2) if m else "?"
        n = int(r[ix["# Samples"]] or 0)
        ie = int(r[ix["Instructions Executed"]] or 0)
        inst_op[op] += ie
        st = {c[6:]: int(r[ix[c]] or 0) for c in cols}
        for k, v in st.items():
            tot[k] += v
            by_op[op.split(".")[0]][k] += v
        tot["_samples"] += n
        by_op[op.split(".")[0]]["_samples"] += n
        lines.append((n, src[:64], st))
    ns = max(tot["_samples"], 1)
    md += [f"## `{title}`", "", f"{ns} samples.  Stall reasons: " + ", ".join(f"{k} {100 * v / ns:.1f} %" for k, v in tot.most_common(9) if k != "_samples") + ".", "",
           "| opcode | share of samples | main reasons |", "|---|---:|---|"]
    for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["_samples"])[:10]:
        md.append(f"| `{op}` | {100 * c['_samples'] / ns:.1f} % | " + ", ".join(f"{k} {v}" for k, v in c.most_common(4) if k != "_samples") + " |")
    md += ["", "Hottest instructions:", ""]
    for n, src, st in sorted(lines, key=lambda t: -t[0])[:8]:
        md.append(f"* {100 * n / ns:.1f} % `{src}` - " + ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v))
    loc = {op: v for op, v in inst_op.items() if op.startswith(("STL", "LDL"))}
    md += ["", f"Local-memory instructions executed: {sum(loc.values())} of {sum(inst_op.values())} warp instructions ("
           + ", ".join(f"{k} {v}" for k, v in sorted(loc.items())) + ").", ""]
(P / "r2_stall_samples.md").write_text("\n".join(md) + "\n")
print("\n".join(md[:30]))
