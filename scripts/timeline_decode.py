"""In-graph kernel timeline of the decode loop (what ncu cannot show: ncu serialises launches and flushes caches).

Runs a short StarVector-1B generate under torch.profiler (CUPTI activity records carry the hardware start/end
timestamps of every kernel node of the replayed CUDA graph), then reports, for the steady-state decode steps:
per-kernel-name time, the gaps between consecutive kernels (negative = overlap from programmatic dependent launch),
and the share of the step in which no kernel of ours was running.  CUPTI adds a little per-launch overhead, so the
step time printed here is a few percent above bench.py's; use it for the SHAPE of the step, not as a bench number.

    python scripts/timeline_decode.py [--batch 1] [--ctx 1024] [--new 24] [--json gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def short_name(name: str) -> str:
    """`void sv::mega::gemv_ring_kernel<true, 1, false>(...)` -> `gemv_ring_kernel<true, 1, false>`."""
    name = name.split("(")[0]
    if name.startswith("void "):
        name = name[5:]
    depth, last = 0, 0
    for i, ch in enumerate(name):
        depth += ch == "<"
        depth -= ch == ">"
        if ch == ":" and depth == 0:
            last = i + 1
    return name[last:]


def analyse(events, step_end_marker: str = "select", skip_steps: int = 2):
    """events: iterable of dicts with name, ts (us), dur (us) for GPU kernels.  A decode step ends with the kernel whose
    name contains `step_end_marker`.  Returns a summary over the steps after the first `skip_steps`."""
    ks = sorted(({"name": short_name(e["name"]), "ts": float(e["ts"]), "dur": float(e["dur"])} for e in events), key=lambda k: k["ts"])
    steps, cur = [], []
    for k in ks:
        cur.append(k)
        if step_end_marker in k["name"]:
            steps.append(cur)
            cur = []
    sizes = defaultdict(int)
    for s in steps:
        sizes[len(s)] += 1
    if not steps:
        return {"steps": 0}
    common = max(sizes, key=sizes.get)                       # the steady-state step has the most frequent launch count
    steady = [s for s in steps[skip_steps:] if len(s) == common] or [s for s in steps if len(s) == common]
    by_name = defaultdict(lambda: {"n": 0, "dur": 0.0, "gap_before": 0.0, "overlap_before": 0.0})
    span = busy = 0.0
    for s in steady:
        span += s[-1]["ts"] + s[-1]["dur"] - s[0]["ts"]
        cover_end = s[0]["ts"]
        for i, k in enumerate(s):
            d = by_name[k["name"]]
            d["n"] += 1
            d["dur"] += k["dur"]
            if i:
                gap = k["ts"] - (s[i - 1]["ts"] + s[i - 1]["dur"])
                d["gap_before" if gap >= 0 else "overlap_before"] += abs(gap)
            start, end = max(k["ts"], cover_end), k["ts"] + k["dur"]
            if end > start:
                busy += end - start
                cover_end = end
    n = len(steady)
    rows = sorted(({"kernel": k, "launches_per_step": v["n"] / n, "us_per_step": v["dur"] / n, "gap_us_per_step": v["gap_before"] / n,
                    "overlap_us_per_step": v["overlap_before"] / n} for k, v in by_name.items()), key=lambda r: -r["us_per_step"])
    return {"steps": n, "launches_per_step": common, "step_us": span / n, "busy_us": busy / n, "idle_us": (span - busy) / n,
            "sum_kernel_us": sum(r["us_per_step"] for r in rows), "kernels": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--ctx", type=int, default=1024, help="decode this many tokens before the profiled window")
    ap.add_argument("--new", type=int, default=24)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile

    from starvector_b200.config import dims_1b
    from starvector_b200.engine import Engine, GenerationParams
    from starvector_b200.weights import synthetic_images, synthetic_state_dict

    d = dims_1b(max_batch=a.batch, max_len=min(8192, 300 + a.ctx + a.new + 64))
    eng = Engine(d, 0)
    eng.load_state_dict(synthetic_state_dict(d, seed=0))
    img = synthetic_images(d, a.batch, seed=1).cuda()
    prompt = torch.tensor([[44, 5678]] * a.batch, dtype=torch.int32).cuda()
    params = GenerationParams(max_new_tokens=a.ctx + a.new, eos_token_id=None, pad_token_id=49152)
    for _ in range(2):                                        # warm-up: graph capture, clocks
        eng.encode_images(img); eng.prefill(prompt); eng.generate(params)
    torch.cuda.synchronize()
    eng.encode_images(img); eng.prefill(prompt)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.generate(params)
        torch.cuda.synchronize()
    ms, steps = eng.last_decode_timing()
    events = [{"name": e.name, "ts": e.time_range.start, "dur": e.time_range.end - e.time_range.start}
              for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tail = sorted(events, key=lambda e: e["ts"])
    out = analyse(tail, skip_steps=max(2, a.ctx))
    out["engine_us_per_step"] = ms / max(steps, 1) * 1000.0
    out["config"] = {"batch": a.batch, "ctx": a.ctx, "new": a.new, "decode": os.environ.get("SV_DECODE", "default"), "pdl": os.environ.get("SV_PDL", "1")}
    print(f"{out['steps']} steady steps, {out.get('launches_per_step')} launches/step, step {out.get('step_us', 0):.1f} us "
          f"(engine timer, whole run: {out['engine_us_per_step']:.1f} us), busy {out.get('busy_us', 0):.1f} us, idle {out.get('idle_us', 0):.1f} us, "
          f"sum of kernel durations {out.get('sum_kernel_us', 0):.1f} us")
    for r in out.get("kernels", [])[:16]:
        print(f"  {r['kernel'][:60]:60s} x{r['launches_per_step']:5.1f}  {r['us_per_step']:8.1f} us  gap {r['gap_us_per_step']:6.1f}  overlap {r['overlap_us_per_step']:6.1f}")
    if a.json:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)
    eng.close()


if __name__ == "__main__":
    main()
