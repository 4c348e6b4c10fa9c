#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-1500; }
TAILN=1 run j_bench python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
SV_FLOW=0 TAILN=1 run j_bench_graph python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
TAILN=6 run j_engine python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
TAILN=48 run j_timeline python scripts/flow_timeline.py --new 8 --json gpurun_out/r02_flow_timeline_j.json
