#!/bin/bash
# Full GPU check used under gpurun: per-kernel parity, engine parity in each decode mode, smoke, short bench.
# Each stage runs in its own process so a trapped kernel cannot poison the later stages.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
run ops python -m pytest tests/test_ops_gpu.py -q --tb=short -m gpu
run engine_graph_pdl python -m pytest tests/test_engine_gpu.py tests/test_facade_gpu.py -q --tb=short -m gpu
SV_MEGA=2 run engine_mega_setmaxnreg python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
SV_MEGA=1 run engine_mega python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
SV_PDL=0 run engine_graph_nopdl python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
SV_DECODE=legacy run engine_legacy python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
run v2 python -m pytest tests/test_v2_gpu.py -q --tb=short -m gpu
SV_DECODE=fused run v2_fused python -m pytest tests/test_v2_gpu.py -q --tb=short -m gpu
run preprocess python -m pytest tests/test_preprocess_gpu.py -q --tb=short -m gpu
SV_STEP_GRAPH=1 run step_graph python -m pytest tests/test_engine_gpu.py tests/test_beam_gpu.py -q --tb=short -m gpu
run beam python -m pytest tests/test_beam_gpu.py -q --tb=short -m gpu
run widening python -m pytest tests/test_widening_gpu.py -q --tb=short -m gpu
SV_MEGA=1 run widening_mega python -m pytest tests/test_widening_gpu.py -q --tb=short -m gpu -k streaming
run full_1b python -m pytest tests/test_full_1b_gpu.py -q --tb=short -m gpu
run smoke python __graft_entry__.py --smoke
TAILN=2 run bench_short python bench.py --steps 1 --warmup 1 --max-new-tokens 256 --no-cpu-baseline
SV_MEGA=1 TAILN=2 run bench_short_mega python bench.py --steps 1 --warmup 1 --max-new-tokens 256 --no-cpu-baseline
SV_MEGA=2 TAILN=2 run bench_short_mega2 python bench.py --steps 1 --warmup 1 --max-new-tokens 256 --no-cpu-baseline
