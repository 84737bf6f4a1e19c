#!/bin/bash
# usage: scripts/ncu_one.sh <kernel-regex> <out-prefix> <python script + args...>
# One `ncu --set full` capture of the first launches of a kernel; exports the raw page as CSV (the .ncu-rep stays on the box).
K=$1; O=gpurun_out/$2; shift 2
ncu --set full --clock-control none --import-source on -k "regex:$K" -c 4 -o /tmp/cap -f "$@" > $O.log 2>&1
ncu -i /tmp/cap.ncu-rep --page raw --csv > $O.raw.csv 2>/dev/null
ncu -i /tmp/cap.ncu-rep --page details --csv > $O.details.csv 2>/dev/null
python - "$O.raw.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("---")
    for w in want:
        for h in hdr:
            if h.startswith(w):
                print(f"{h:80s} {r[idx[h]]}")
                break
PY
