#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-1500; }
TAILN=1 run k_bench_graph python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
SV_FLOW=1 TAILN=1 run k_bench_flow python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
TAILN=6 run k_engine python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
TMO=900 TAILN=6 run k_full1b python -m pytest tests/test_full_1b_gpu.py -q --tb=short -m gpu
