#!/bin/bash
# One `ncu --set full` capture on the GPU box, reduced there to the two small text files that travel back:
#   gpurun_out/<tag>_raw.csv     raw page (all counters, one row per launch)  -> tools/ncu_summary.py
#   gpurun_out/<tag>_stalls.txt  per-kernel stall digest of the source page   (tools/ncu_stalls.py)
# usage: [NCU_SKIP=n] tools/ncu_capture.sh <tag> <kernel-regex> <launch-count> <command...>   (NCU_SKIP: matching launches to skip first)
# The command is run once without ncu first (it must exit 0); the .ncu-rep stays in /tmp (gpurun_out is capped at 64 MiB).
set -u
tag=$1; regex=$2; count=$3; shift 3
export SAPCU_WS_CAP_GB=${SAPCU_WS_CAP_GB:-3}      # small workspaces: ncu saves / restores device memory around every replay
mkdir -p gpurun_out
"$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "$tag: plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "${NCU_SKIP:-0}" -c "$count" -f -o /tmp/${tag} "$@" > gpurun_out/${tag}_ncu.log 2>&1 || { echo "$tag: ncu failed"; tail -5 gpurun_out/${tag}_ncu.log; exit 1; }
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv > /tmp/${tag}_src.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/${tag}_src.csv "" 14 > gpurun_out/${tag}_stalls.txt 2>&1
rm -f /tmp/${tag}.ncu-rep /tmp/${tag}_src.csv
echo "$tag: ok ($(wc -l < gpurun_out/${tag}_raw.csv) raw rows)"
