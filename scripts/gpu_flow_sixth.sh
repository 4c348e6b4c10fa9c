#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-1500; }
TAILN=6 run f_engine python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
TAILN=50 run f_timeline python scripts/flow_timeline.py --new 8 --json gpurun_out/r02_flow_timeline_f.json
for la in 0 8; do
  SV_FLOW_L2AHEAD=$la TAILN=1 run f_bench_la$la python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
done
SV_FLOW=3 TAILN=1 run f_bench_norealloc python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
SV_FLOW_L2AHEAD=0 TAILN=50 run f_timeline_la0 python scripts/flow_timeline.py --new 8
