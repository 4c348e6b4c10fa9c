#!/bin/bash
# Final check of the round: the whole -m gpu suite as the driver runs it, with per-test timestamps (a killed run still shows
# where the time went), then the beam bench (device vs host-stepped hypothesis at 1B dims).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== suite"
timeout -s INT 700 python -m pytest tests -v --tb=short -m gpu --durations=25 -p no:cacheprovider 2>&1 | awk '{ print strftime("%H:%M:%S"), $0; fflush() }' > gpurun_out/c3_suite.log
echo "exit ${PIPESTATUS[0]} (suite)"
grep -E "passed|failed|FAILED|ERROR|Interrupt" gpurun_out/c3_suite.log | tail -15 | cut -c1-300
echo "=== beam_bench"
timeout 120 python scripts/beam_bench.py --max-new-tokens 512 --repeats 2 > gpurun_out/c3_beam_bench.log 2>&1; echo "exit $?"; tail -1 gpurun_out/c3_beam_bench.log | cut -c1-900
