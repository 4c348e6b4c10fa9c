#!/bin/bash
# Round-2 re-entry check: parity at the benchmarked shapes, tiled vs row-major weights A/B, beam baseline, in-graph timeline.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/c1_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-3} gpurun_out/c1_$name.log | cut -c1-1500; }
TMO=600 TAILN=8 run tests python -m pytest tests/test_engine_gpu.py tests/test_full_1b_gpu.py -q --tb=short -m gpu -x
TAILN=1 run bench_tiled python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
SV_TILED=0 TAILN=1 run bench_rowmajor python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
TAILN=1 run beam python scripts/beam_bench.py --max-new-tokens 256 --repeats 2
TAILN=2 run timeline python scripts/timeline_decode.py --ctx 2048 --new 24 --json gpurun_out/c1_timeline_ctx2300.json
