#!/bin/bash
# Whole-tree check + round-2 measurements in ONE call (a call costs ~4 GPU-minutes before the first command runs).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/c2_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-3} gpurun_out/c2_$name.log | cut -c1-1800; }
TMO=400 TAILN=12 run beam_tests python -m pytest tests/test_beam_gpu.py -q --tb=short -m gpu
TAILN=1 run beam_bench python scripts/beam_bench.py --max-new-tokens 512 --repeats 2
TAILN=2 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 TAILN=15 run suite python -m pytest tests -q --tb=short -m gpu --deselect tests/test_beam_gpu.py
TAILN=1 run bench python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras
echo "=== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 380 --csv --log-file gpurun_out/c2_launches.csv python bench.py --steps 1 --warmup 1 --max-new-tokens 48 --no-cpu-baseline --no-extras > gpurun_out/c2_ncu_bench.log 2>&1; echo "exit $?"
echo "=== ncu beam kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_ -s 8 -c 4 -f -o gpurun_out/c2_beam_ncu python scripts/beam_bench.py --max-new-tokens 16 --repeats 1 > gpurun_out/c2_ncu_beam.log 2>&1; echo "exit $?"
timeout 120 ncu -i gpurun_out/c2_beam_ncu.ncu-rep --page raw --csv > gpurun_out/c2_beam_ncu_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
