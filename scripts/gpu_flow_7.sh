#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-1500; }
for v in st4 st3; do
  SV_LIB_PATH=starvector_b200/libstarvector_b200_$v.so TAILN=1 run g_bench_$v python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
done
SV_LIB_PATH=starvector_b200/libstarvector_b200_st3.so TAILN=45 run g_timeline_st3 python scripts/flow_timeline.py --new 8
