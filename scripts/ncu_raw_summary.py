#!/usr/bin/env python
"""Per-kernel table from an `ncu --page raw --csv` export of a `--set full` capture.
usage: ncu_raw_summary.py raw.csv out_prefix "header comment"   -> out_prefix.json (what bench.py reads) + out_prefix.txt"""
import collections
import csv
import json
import re
import sys

raw, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def col(r, name, default=0.0):
    for h in hdr:
        if h.startswith(name):
            try:
                v = float(r[ix[h]].replace(",", ""))
            except ValueError:
                return default
            u = units[ix[h]]
            if u in ("Mbyte", "MByte"):
                v *= 1e6
            elif u in ("Kbyte", "KByte"):
                v *= 1e3
            elif u in ("Gbyte", "GByte"):
                v *= 1e9
            elif u == "ms":
                v *= 1e3
            elif u == "ns":
                v /= 1e3
            elif u in ("second", "s"):
                v *= 1e6
            return v
    return default


groups = collections.OrderedDict()
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    name = re.sub(r"\(int\)", "", name)
    name = name.split("(")[0]
    grid = r[ix["Grid Size"]] if "Grid Size" in ix else ""
    key = (name, grid)
    g = groups.setdefault(key, [])
    g.append(dict(us=col(r, "gpu__time_duration.sum"), rd=col(r, "dram__bytes_read.sum"), wr=col(r, "dram__bytes_write.sum"),
                  dram=col(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                  tensor=col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  issue=col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  inst=col(r, "smsp__inst_executed.sum"), regs=col(r, "launch__registers_per_thread"),
                  smem=col(r, "launch__shared_mem_per_block_dynamic")))
table = collections.OrderedDict()
lines = [f"# {note}", "# one row per distinct kernel / grid, averaged over its launches; cold caches, serialised (ncu replay);",
         "# dram write counts only what left L2 during the kernel",
         f"{'kernel':58s} {'grid':>14s} {'n':>3s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'dram%':>6s} {'tens%':>6s} {'issue%':>6s} {'regs':>4s} {'smemKB':>6s}"]
for (name, grid), g in groups.items():
    n = len(g)
    avg = {k: sum(x[k] for x in g) / n for k in g[0]}
    table.setdefault(name, []).append({"grid": grid, "launches": n, "us": round(avg["us"], 2), "dram_read_bytes": int(avg["rd"]),
                                       "dram_write_bytes": int(avg["wr"]), "dram_pct": round(avg["dram"], 1),
                                       "tensor_pct": round(avg["tensor"], 1), "issue_pct": round(avg["issue"], 1)})
    lines.append(f"{name[:58]:58s} {grid:>14s} {n:3d} {avg['us']:8.1f} {avg['rd'] / 1e6:8.1f} {avg['wr'] / 1e6:8.1f} {avg['dram']:6.1f} "
                 f"{avg['tensor']:6.1f} {avg['issue']:6.1f} {int(avg['regs']):4d} {avg['smem'] / 1024:6.1f}")
json.dump(table, open(out + ".json", "w"), indent=1)
open(out + ".txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
