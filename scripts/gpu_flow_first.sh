#!/bin/bash
# First GPU run of the dataflow decode kernel: parity of everything that decodes, then A/B against the graph path.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-900; }
TAILN=30 run flow_timeline python scripts/flow_timeline.py --new 8 --json gpurun_out/r02_flow_timeline.json
TAILN=12 run flow_engine python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu -x
TAILN=12 run flow_full1b python -m pytest tests/test_full_1b_gpu.py -q --tb=short -m gpu -x
TAILN=1 run bench_flow python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline
SV_FLOW=2 TAILN=1 run bench_flow2 python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline
SV_FLOW=0 TAILN=1 run bench_graph python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline
TAILN=30 run flow_timeline_ctx2k python scripts/flow_timeline.py --ctx 2048 --new 8 --json gpurun_out/r02_flow_timeline_ctx2k.json
TAILN=8 run rest python -m pytest tests -q --tb=short -m gpu -x --deselect tests/test_engine_gpu.py --deselect tests/test_full_1b_gpu.py
