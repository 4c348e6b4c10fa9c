#!/bin/bash
# One GPU call that checks the whole tree: the -m gpu suite, smoke(), and a short bench of both decode modes.
#   scripts/gpurun_retry.sh 2400 'bash scripts/gpu_check.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/check_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/check_$name.log | cut -c1-1200; }
TMO=1500 TAILN=12 run suite python -m pytest tests -q --tb=short -m gpu
TAILN=3 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=1 run bench_graph python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
SV_FLOW=1 TAILN=1 run bench_flow python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline --no-extras
