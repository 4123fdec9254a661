#!/bin/bash
# usage: scripts/ncu_export.sh prefix   -- turns prefix.ncu-rep into prefix.raw.csv + prefix.source.csv.gz + prefix.summary.txt and
# removes the report (gpurun_out/ is capped at 64 MiB; a --set full report with sources is ~9 MB per kernel)
p="$1"
ncu -i "$p.ncu-rep" --page raw --csv > "$p.raw.csv" 2>/dev/null
ncu -i "$p.ncu-rep" --page source --csv 2>/dev/null | gzip -9 > "$p.source.csv.gz"
python scripts/ncu_summary.py "$p" > "$p.summary.txt" 2>&1
rm -f "$p.ncu-rep"
