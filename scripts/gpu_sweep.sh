#!/bin/bash
# Short decode-speed sweeps (256 new tokens, B=1): prints ms/token for each environment setting.
cd "$(dirname "$0")/.."
one() { label=$1; shift; env "$@" python bench.py --steps 1 --warmup 1 --max-new-tokens 256 --no-cpu-baseline > /tmp/sweep.json 2> /tmp/sweep.err; python -c "import sys,json; d=json.loads(open('/tmp/sweep.json').read()); print('$label', round(d['decode_ms_per_token_step'],4), round(d['value'],1), d['engine'][:50])" 2>/dev/null || { echo "$label FAILED"; tail -3 /tmp/sweep.err; }; }
for spec in "$@"; do one "$spec" $spec; done
