#!/bin/bash
# First GPU call of the next round: everything that was built after round 1's GPU budget ran out gets measured or
# checked in one go (about 6-8 minutes).  Output lands in gpurun_out/r02_*.
#   python -m starvector_b200.build --variant nwc4      # here, before the call: the .so travels with the snapshot
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-700; }
# 1. correctness of what has never run on a GPU
run widening python -m pytest tests/test_widening_gpu.py tests/test_preprocess_gpu.py -q --tb=short -m gpu
SV_STEP_GRAPH=1 run step_graph python -m pytest tests/test_engine_gpu.py tests/test_beam_gpu.py -q --tb=short -m gpu
SV_MEGA=2 run mega_setmaxnreg python -m pytest tests/test_engine_gpu.py tests/test_full_1b_gpu.py -q --tb=short -m gpu
# 2. where the decode step's time goes inside the replayed graph (CUPTI), short and long context
TAILN=22 run timeline_ctx300 python scripts/timeline_decode.py --ctx 32 --new 24 --json gpurun_out/r02_timeline_ctx300.json
TAILN=22 run timeline_ctx2300 python scripts/timeline_decode.py --ctx 2048 --new 24 --json gpurun_out/r02_timeline_ctx2300.json
# 3. A/B of the persistent kernel variants against the default path (same box, same clocks)
for mode in 0 1 2; do
  SV_MEGA=$mode TAILN=1 run bench_mega$mode python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline
done
NWC4=starvector_b200/libstarvector_b200_nwc4.so
if [ -f $NWC4 ]; then
  SV_LIB_PATH=$NWC4 run nwc4_parity python -m pytest tests/test_engine_gpu.py tests/test_full_1b_gpu.py -q --tb=short -m gpu
  SV_LIB_PATH=$NWC4 TAILN=1 run bench_nwc4 python bench.py --steps 2 --warmup 3 --max-new-tokens 512 --no-cpu-baseline
  SV_LIB_PATH=$NWC4 TAILN=22 run timeline_nwc4 python scripts/timeline_decode.py --ctx 32 --new 24 --json gpurun_out/r02_timeline_nwc4.json
fi
# 4. beam search with and without the step graph; preprocessing after the two layout changes
TAILN=1 run beam_eager python scripts/beam_bench.py --num-beams 2 --max-new-tokens 512
SV_STEP_GRAPH=1 TAILN=1 run beam_graph python scripts/beam_bench.py --num-beams 2 --max-new-tokens 512
TAILN=1 run preprocess python scripts/preprocess_bench.py --iters 20
timeout 120 ncu --set full --clock-control none --import-source on -k regex:resize_ -c 2 -f -o gpurun_out/r02_preprocess \
  python scripts/preprocess_bench.py --iters 1 --no-cpu > gpurun_out/r02_ncu_pre.log 2>&1
timeout 30 ncu -i gpurun_out/r02_preprocess.ncu-rep --page raw --csv > gpurun_out/r02_preprocess_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
