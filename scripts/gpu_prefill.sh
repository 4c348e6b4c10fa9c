#!/bin/bash
# prefill latency (ViT + adapter + decoder prefill) under different settings
cd "$(dirname "$0")/.."
for spec in "$@"; do env $spec python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from starvector_b200.config import dims_1b
from starvector_b200.engine import Engine
from starvector_b200.weights import synthetic_images, synthetic_state_dict
d = dims_1b(max_batch=8, max_len=512)
e = Engine(d, 0); e.load_state_dict(synthetic_state_dict(d, seed=0))
for B in (1, 8):
    img = synthetic_images(d, B).cuda(); pr = torch.tensor([[44, 5678]] * B, dtype=torch.int32).cuda()
    for _ in range(3): e.encode_images(img); e.prefill(pr)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): e.encode_images(img); e.prefill(pr)
    b.record(); torch.cuda.synchronize()
    print(os.environ.get("LABEL", ""), "B=%d prefill %.3f ms/image" % (B, a.elapsed_time(b) / 10 / B), flush=True)
e.close()
PY
done
