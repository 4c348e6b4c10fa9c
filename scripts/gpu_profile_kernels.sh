#!/bin/bash
# ncu --set full captures of individual kernels.   usage: gpu_profile_kernels.sh TAG REGEX SKIP COUNT
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01}; RX=${2:-gemv_ring}; SKIP=${3:-118}; CNT=${4:-6}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${RX} -s ${SKIP} -c ${CNT} -o gpurun_out/k_${TAG} \
    python scripts/profile_decode.py --new 4 --reps 2 > gpurun_out/k_profile_${TAG}.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/k_${TAG}.ncu-rep
