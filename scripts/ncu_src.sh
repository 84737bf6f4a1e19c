#!/bin/bash
# usage: scripts/ncu_src.sh <kernel-regex> <out-prefix> <cmd...>: one full capture (2nd launch) + key metrics + top SASS stall sites
K=$1; O=gpurun_out/$2; shift 2
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 1 -c 1 -o /tmp/cap -f "$@" > $O.log 2>&1
ncu -i /tmp/cap.ncu-rep --page raw --csv > $O.raw.csv 2>/dev/null
ncu -i /tmp/cap.ncu-rep --page source --csv > $O.src.csv 2>/dev/null
python - "$O.raw.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "l1tex__throughput.avg.pct", "lts__throughput.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread"]
for r in rows[2:]:
    for w in want:
        for h in hdr:
            if h.startswith(w):
                print(f"{h:80s} {r[idx[h]]}"); break
PY
python scripts/ncu_src_top.py $O.src.csv "" 0 45
