"""Device timeline of the dataflow decode kernel (SV_MEGA_DEBUG=1): CTA 0 / thread 0 stamps clock64() when the inputs of a
phase have arrived ("ready") and when its outputs are stored ("done"), for the first token of the first launch.

    SV_MEGA_DEBUG=1 python scripts/flow_timeline.py [--batch 1] [--ctx 0] [--new 8] [--json gpurun_out/flow_timeline.json]

wait = previous phase done -> this phase ready (hop latency + the slowest producer CTA); work = ready -> done.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SV_MEGA_DEBUG", "1")
# the records are compiled out of the shipped library (they cost registers): python -m starvector_b200.build --variant timeline
os.environ.setdefault("SV_LIB_PATH", os.path.join(ROOT, "starvector_b200", "libstarvector_b200_timeline.so"))
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
    if a.ctx:      # long-context timeline: teacher-forced single-token launches; the records are of the last one
        ids = torch.full((a.batch,), 17, dtype=torch.int32, device="cuda")
        for _ in range(a.ctx):
            eng.decode_step(ids, return_logits=False)
        torch.cuda.synchronize()
        continue
    out = eng.generate(GenerationParams(max_new_tokens=a.new, eos_token_id=None, pad_token_id=49152))
    torch.cuda.synchronize()
    ms, steps = eng.last_decode_timing()
    print(f"rep {rep}: {steps} decode steps in {ms:.3f} ms -> {ms / max(steps, 1) * 1000:.1f} us/step", flush=True)
raw = eng.debug_timeline(8192, raw=True)
KIND = ["qkv", "c_proj", "fc", "fc2", "lm_head"]
SUB = {1: "enter", 2: "x_ready", 3: "ln_done", 4: "w0_landed", 5: "w_last", 6: "done"}
ATT = {40: "att.enter", 41: "att.q_ready", 42: "att.blocks", 43: "att.tree", 44: "att.stored", 46: "merge.done", 47: "select.done"}


def name_of(i):
    if i in ATT:
        return ATT[i]
    if i >= 64:
        k = i - 64
        return f"P.{KIND[k // 2]}.start" if k % 2 == 0 and k // 2 < len(KIND) else f"P.{k}"
    return f"{KIND[i // 8]}.{SUB.get(i % 8, i % 8)}"


cons = [((v >> 48) & 0xffff, v & 0xffffffffffff) for v in raw[:4096] if v]
prod = [((v >> 48) & 0xffff, v & 0xffffffffffff) for v in raw[4096:] if v]
if cons:
    t0 = cons[0][1]
    total = cons[-1][1] - t0
    # per-record-name: mean delta from the previous consumer record, over the layers
    sums, counts, order = {}, {}, []
    prev = t0
    for i, t in cons:
        n = name_of(i)
        if n not in sums:
            sums[n] = 0; counts[n] = 0; order.append(n)
        sums[n] += t - prev; counts[n] += 1
        prev = t
    print(f"consumer records {len(cons)}, producer records {len(prod)}; first token of the launch: {total} cycles on CTA 0")
    print("  record              n   mean delta from previous record (cycles)   share")
    for n in order:
        print(f"  {n:18s} {counts[n]:3d}   {sums[n] / counts[n]:10.0f}   {sums[n] * 100.0 / total:5.1f} %")
    # producer lead: at every consumer 'enter' of a GEMV phase, had the producer already issued that phase's last slab?
    if prod:
        print("  producer (cycles after the consumer's first record): first 12:", [(name_of(i), t - t0) for i, t in prod[:12]])
        print("  consumer first 16:", [(name_of(i), t - t0) for i, t in cons[:16]])
    if a.json:
        json.dump({"consumer": [(name_of(i), t - t0) for i, t in cons], "producer": [(name_of(i), t - t0) for i, t in prod],
                   "mean_delta": {n: sums[n] / counts[n] for n in order}, "total_cycles": total, "ctx": a.ctx, "batch": a.batch}, open(a.json, "w"))
else:
    print("no timeline records (SV_MEGA_DEBUG unset or the dataflow kernel is not in use)")
eng.close()
