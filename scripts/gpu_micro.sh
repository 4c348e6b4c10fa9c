#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 scripts/bin/l2_prefetch_test 2>&1 | tee gpurun_out/r02_l2_prefetch_test.log
timeout 120 scripts/bin/hop_latency 2>&1 | tee gpurun_out/r02_hop_latency.log
timeout 200 scripts/bin/ring_stream 2>&1 | tee gpurun_out/r02_ring_stream2.log
