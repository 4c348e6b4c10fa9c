"""Device timeline of the dataflow decode kernel (SV_MEGA_DEBUG=1): CTA 0 / thread 0 stamps clock64() when the inputs of a
phase have arrived ("ready") and when its outputs are stored ("done"), for the first token of the first launch.

    SV_MEGA_DEBUG=1 python scripts/flow_timeline.py [--batch 1] [--ctx 0] [--new 8] [--json gpurun_out/flow_timeline.json]

wait = previous phase done -> this phase ready (hop latency + the slowest producer CTA); work = ready -> done.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SV_MEGA_DEBUG", "1")
import torch

from starvector_b200.config import dims_1b
from starvector_b200.engine import Engine, GenerationParams
from starvector_b200.weights import synthetic_images, synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--new", type=int, default=8)
ap.add_argument("--ctx", type=int, default=0, help="teacher-force this many tokens first (long-context timeline)")
ap.add_argument("--json", default="")
a = ap.parse_args()
d = dims_1b(max_batch=a.batch, max_len=min(8192, 300 + a.ctx + a.new + 64))
eng = Engine(d, 0)
print(eng.describe(), flush=True)
eng.load_state_dict(synthetic_state_dict(d, seed=0))
img = synthetic_images(d, a.batch, seed=1).cuda()
prompt = torch.tensor([[44, 5678]] * a.batch, dtype=torch.int32).cuda()
for rep in range(2):
    eng.encode_images(img)
    eng.prefill(prompt)
    ids = torch.full((a.batch,), 17, dtype=torch.int32, device="cuda")
    for _ in range(a.ctx):
        eng.decode_step(ids)
    out = eng.generate(GenerationParams(max_new_tokens=a.new, eos_token_id=None, pad_token_id=49152))
    torch.cuda.synchronize()
    ms, steps = eng.last_decode_timing()
    print(f"rep {rep}: {steps} decode steps in {ms:.3f} ms -> {ms / max(steps, 1) * 1000:.1f} us/step", flush=True)
tl = eng.debug_timeline()
L = d.n_layer
names = ["qkv.ready", "qkv.done", "attn.done", "merge.done", "cproj.ready", "cproj.done", "fc.ready", "fc.done", "fc2.ready", "fc2.done"]
if len(tl) >= 10 * L + 2:
    per = {n: 0 for n in names}
    prev = tl[0]
    for l in range(L):
        for k, n in enumerate(names):
            v = tl[l * 10 + k]
            per[n] += v - prev
            prev = v
    tail = tl[10 * L:]
    total = tl[-1] - tl[0]
    print(f"stamps {len(tl)}  total {total} cycles for the first token of the launch (CTA 0)")
    for n in names:
        print(f"  {n:12s} {per[n] / L:9.0f} cycles/layer  ({per[n] * 100.0 / total:5.1f} % of the token)")
    print("  tail (lm_head.ready, lm_head.done, select.done) deltas:", [tail[i] - (tail[i - 1] if i else tl[10 * L - 1]) for i in range(len(tail))])
    if a.json:
        json.dump({"cycles_per_layer": {n: per[n] / L for n in names}, "total_cycles": total, "stamps": tl, "ctx": a.ctx, "batch": a.batch},
                  open(a.json, "w"))
else:
    print("timeline too short:", len(tl))
eng.close()
