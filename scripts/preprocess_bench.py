"""Image-preprocessing probe (not the headline bench): n RGBA images of side `--side` -> [n,3,224,224].
GPU arm = `ImageTrainProcessor.batch` on host uint8 arrays (uploads + 2 kernels + sync inside the timed region, CUDA-event
and wall-clock timed); CPU arm = the reference recipe on Pillow/torchvision (oracle.preprocess.reference_transform), one
image after another as `ImageEncoder.process_images` does (reference image_encoder.py:113-117).  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from oracle import preprocess as P
    from starvector_b200.preprocess import ImageTrainProcessor

    imgs = [P.synthetic_image(args.side, args.side, 4, seed=i) for i in range(args.n)]
    pinned = [torch.from_numpy(a).pin_memory().numpy() for a in imgs]
    proc = ImageTrainProcessor(size=224, dtype=torch.bfloat16)
    out = {}
    for name, batch in (("pageable", imgs), ("pinned", pinned)):
        for _ in range(3):
            proc.batch(batch)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.iters):
            proc.batch(batch)
        ev1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / args.iters
        out[name] = {"ms_per_batch_wall": round(wall * 1e3, 3), "ms_per_batch_device": round(ev0.elapsed_time(ev1) / args.iters, 3),
                     "images_per_s": round(args.n / wall, 1)}
    in_bytes = sum(a.nbytes for a in imgs)
    if not args.no_cpu:
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            for a in imgs:
                P.reference_transform(a, 224, P.ALPHA_WHITE)
        cpu = (time.perf_counter() - t0) / reps
        out["cpu_reference"] = {"ms_per_batch": round(cpu * 1e3, 2), "images_per_s": round(args.n / cpu, 1), "threads": 1}
    out["config"] = {"n": args.n, "side": args.side, "channels": 4, "out": "bf16 [n,3,224,224]", "input_MB": round(in_bytes / 1e6, 2),
                     "h2d_GBps_at_pinned_rate": round(in_bytes / (out["pinned"]["ms_per_batch_wall"] * 1e-3) / 1e9, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
