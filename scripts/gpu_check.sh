#!/bin/bash
# Full GPU check used under gpurun: per-kernel parity, engine parity (fallback linear, then tcgen05), smoke, short bench.
# Each stage runs in its own process so a trapped kernel cannot poison the later stages.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 12 gpurun_out/$name.log; }
run ops_basic   python -m pytest tests/test_ops_gpu.py -q --tb=short -m gpu -k "layernorm or attention"
run ops_rowgroup python -m pytest tests/test_ops_gpu.py -q --tb=short -m gpu -k "rowgroup and not agrees"
run ops_tcgen05 python -m pytest tests/test_ops_gpu.py -q --tb=short -m gpu -k "tcgen05"
SV_LINEAR_IMPL=rowgroup run engine_rowgroup python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu
run engine_auto python -m pytest tests/test_engine_gpu.py tests/test_facade_gpu.py -q --tb=short -m gpu
run smoke python __graft_entry__.py --smoke
run bench_short python bench.py --steps 1 --warmup 1 --max-new-tokens 256 --no-cpu-baseline
