#!/bin/bash
# ncu --set full capture of ONE persistent-decode launch (8 tokens) of the 1B model + a short timing run.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01}
python scripts/profile_decode.py --new 65 --reps 2 2>&1 | tail -2
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 1 -c 1 -o gpurun_out/mega_${TAG} \
    python scripts/profile_decode.py --new 9 --reps 2 > gpurun_out/mega_profile_${TAG}.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/mega_profile_${TAG}.log; ls -la gpurun_out/mega_${TAG}.ncu-rep
