#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/r02_$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-4} gpurun_out/r02_$name.log | cut -c1-1500; }
TAILN=30 run ring_stream scripts/bin/ring_stream
TAILN=60 run flow_timeline_b python scripts/flow_timeline.py --new 8 --json gpurun_out/r02_flow_timeline_b.json
