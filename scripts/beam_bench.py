"""Beam-search throughput probe (not the headline bench): StarVector-1B dims, synthetic weights, one image,
`num_beams` beams, fixed number of steps (no EOS, early_stopping="never").  Prints one JSON line with tokens/s of the
returned hypothesis and ms per beam step for the device-resident loop (sv_beam_search) and the host-stepped loop, next to the
one-beam graph-replayed loop on the same engine."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-beams", type=int, default=2)
    ap.add_argument("--max-new-tokens", type=int, default=512)
    ap.add_argument("--repeats", type=int, default=2)
    args = ap.parse_args()
    from starvector_b200.beam_search import beam_search
    from starvector_b200.config import dims_1b
    from starvector_b200.engine import Engine, GenerationParams
    from starvector_b200.weights import synthetic_images, synthetic_state_dict

    n_new, nb = args.max_new_tokens, args.num_beams
    d = dims_1b(max_batch=nb, max_len=257 + 2 + n_new + 32)
    eng = Engine(d, 0)
    eng.load_state_dict(synthetic_state_dict(d, seed=0))
    dev = torch.device("cuda", 0)
    img = synthetic_images(d, 1, seed=1).to(dev)
    prompt = torch.tensor([[44, 78]], dtype=torch.int32, device=dev)

    def beams(impl):
        return lambda: beam_search(eng, img, prompt, num_beams=nb, max_new_tokens=n_new, early_stopping="never",
                                   eos_token_id=None, pad_token_id=49152, repetition_penalty=3.1, length_penalty=-1.0, impl=impl)

    def one_beam():
        eng.encode_images(img)
        eng.prefill(prompt)
        return eng.generate(GenerationParams(max_new_tokens=n_new, eos_token_id=None, pad_token_id=49152))

    out = {}
    for name, fn in (("beam_device", beams("device")), ("beam_host_stepped", beams("host")), ("one_beam_graph", one_beam)):
        fn()
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(args.repeats):
            t0 = time.perf_counter()
            ids = fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[name] = {"tokens": int(ids.shape[1]), "seconds": round(best, 4), "tokens_per_s": round(ids.shape[1] / best, 1),
                     "ms_per_step": round(1e3 * best / ids.shape[1], 3)}
        if name != "beam_host_stepped":
            ms, steps = eng.last_decode_timing()
            out[name]["device_ms_per_step"] = round(ms / max(steps, 1), 4)
    out["same_hypothesis"] = bool(torch.equal(beams("device")().cpu(), beams("host")().cpu()))
    out["config"] = {"model": "StarVector-1B dims, synthetic weights", "num_beams": nb, "max_new_tokens": n_new,
                     "timing": "host wall clock around the call incl. encode+prefill, best of %d" % args.repeats}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
