#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_flow -c 1 -f -o gpurun_out/r02_flow_ncu python scripts/profile_decode.py --new 5 --reps 1 > gpurun_out/r02_flow_ncu.log 2>&1
echo "ncu exit $?"
timeout 120 ncu -i gpurun_out/r02_flow_ncu.ncu-rep --page raw --csv > gpurun_out/r02_flow_ncu_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_flow_ncu_raw.csv')))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
for h,v in zip(hdr,vals):
    if any(k in h for k in ('issue_stalled','gpu__time_duration','dram__bytes','registers','inst_executed.sum','smsp__inst_executed.avg','warps_active','dram_throughput','lts__t_sector_hit','issue_active','icc','inst_cache','idc')):
        print(h,'=',v)
PY
