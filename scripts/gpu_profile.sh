#!/bin/bash
# ncu launch list (per-kernel device time + DRAM bytes) of a short 1B decode.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_decode.py --new 4 --reps 2 > gpurun_out/profile_${TAG}.log 2>&1
echo "launch list exit $?"; tail -3 gpurun_out/profile_${TAG}.log
