#!/bin/bash
# ncu --set full captures of individual decode kernels on the per-phase (graph) path.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01}
export SV_MEGA=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_decode_fused -s 30 -c 2 -o gpurun_out/attn_${TAG} \
    python scripts/profile_decode.py --new 4 --reps 2 > gpurun_out/attn_profile_${TAG}.log 2>&1
echo "ncu attn exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemvp_kernel -s 120 -c 5 -o gpurun_out/gemv_${TAG} \
    python scripts/profile_decode.py --new 4 --reps 2 > gpurun_out/gemv_profile_${TAG}.log 2>&1
echo "ncu gemv exit $?"
ls -la gpurun_out/*.ncu-rep
