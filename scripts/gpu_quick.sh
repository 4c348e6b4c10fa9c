#!/bin/bash
# Quick GPU iteration: smoke, engine parity (default mode), 1B mode-equivalence, short bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-5} gpurun_out/$name.log | cut -c1-900; }
TAILN=1 run smoke python __graft_entry__.py --smoke
run engine_mega python -m pytest tests/test_engine_gpu.py tests/test_facade_gpu.py -q --tb=short -m gpu -x
run full_1b python -m pytest tests/test_full_1b_gpu.py -q --tb=short -m gpu -x
TAILN=1 run bench_short python bench.py --steps 1 --warmup 1 --max-new-tokens 256 --no-cpu-baseline
