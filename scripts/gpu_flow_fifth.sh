#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-1500; }
TAILN=6 run e_engine python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu -x
TAILN=50 run e_timeline python scripts/flow_timeline.py --new 8 --json gpurun_out/r02_flow_timeline_e.json
for la in 0 8 12; do
  SV_FLOW_L2AHEAD=$la TAILN=1 run e_bench_la$la python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
done
SV_FLOW=3 TAILN=1 run e_bench_norealloc python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
TAILN=6 run e_full1b python -m pytest tests/test_full_1b_gpu.py -q --tb=short -m gpu -x -k "modes_agree or matches_cpu or batch8"
