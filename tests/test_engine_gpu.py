"""End-to-end parity of the CUDA path against the CPU oracle, through the C-ABI (ctypes).

Contract (DESIGN.md "parity"): the engine computes in bf16 storage / fp32 accumulate with the
reference's rounding points.  Float outputs are compared with the fp32 oracle and must be at
least as close as ~2x the reference's own bf16 run; greedy ids must equal the bf16 oracle's
wherever the oracle's top-1/top-2 logit margin exceeds the stated tolerance.
"""
import os

import pytest
import torch

from oracle.pipeline import OracleStarVector
from parity import check_greedy_ids, oracle_greedy
from starvector_b200.config import ModelDims, dims_tiny
from starvector_b200.engine import Engine, GenerationParams
from starvector_b200.weights import synthetic_images, synthetic_state_dict

pytestmark = pytest.mark.gpu

PROMPT = [44, 78]
MARGIN_TOL = 0.05          # logits: oracle top-1/top-2 margin below which an id flip is tolerated


def _err(a, ref):
    d = (a.float().cpu() - ref.float().cpu()).abs()
    return d.max().item(), d.mean().item()


def _as_accurate_as_bf16(engine_out, oracle_bf16, oracle_fp32, slack=2.0, floor=2e-2):
    e_max, e_mean = _err(engine_out, oracle_fp32)
    o_max, o_mean = _err(oracle_bf16, oracle_fp32)
    assert e_max <= slack * o_max + floor, f"max err {e_max:.4f} vs bf16-oracle {o_max:.4f}"
    assert e_mean <= slack * o_mean + floor / 10, f"mean err {e_mean:.5f} vs bf16-oracle {o_mean:.5f}"


@pytest.fixture(scope="module")
def tiny():
    d = dims_tiny()
    sd = synthetic_state_dict(d, seed=0, init="randomized")
    eng = Engine(d, 0)
    eng.load_state_dict(sd)
    pad = d.vocab - 4
    o16 = OracleStarVector(d, sd, dtype=torch.bfloat16, pad_token_id=pad)
    o32 = OracleStarVector(d, sd, dtype=torch.float32, pad_token_id=pad)
    img = synthetic_images(d, 2, seed=1)
    yield d, sd, eng, o16, o32, img
    eng.close()


def test_golden_fixture_vision(golden_dir):
    """Engine vs the fixture written by the reference's own ViT/Adapter modules."""
    for norm in ("layer_norm", "batch_norm"):
        g = torch.load(os.path.join(golden_dir, f"tiny_v1_{norm}.pt"), weights_only=False)
        d = ModelDims(**g["dims"])
        sd = synthetic_state_dict(d, seed=g["seed"], init=g["init"])
        eng = Engine(d, 0)
        eng.load_state_dict(sd)
        emb, vit = eng.encode_images(synthetic_images(d, 2, seed=g["image_seed"]), return_embeds=True, return_vit=True)
        v_max, _ = _err(vit, g["vit_out"])
        a_max, _ = _err(emb, g["adapter_out"])
        assert v_max < 0.08 and a_max < 0.08, (norm, v_max, a_max)
        eng.close()


def test_encode_images(tiny):
    d, sd, eng, o16, o32, img = tiny
    emb, vit = eng.encode_images(img, return_embeds=True, return_vit=True)
    _as_accurate_as_bf16(vit, o16.image_encoder(img), o32.image_encoder(img.float()))
    _as_accurate_as_bf16(emb, o16.image_projection(o16.image_encoder(img)),
                         o32.image_projection(o32.image_encoder(img.float())))


def test_prefill_and_teacher_forced_logits(tiny, golden_dir):
    d, sd, eng, o16, o32, img = tiny
    g = torch.load(os.path.join(golden_dir, "tiny_v1_layer_norm.pt"), weights_only=False)
    forced = g["forced_ids"]                                    # [2, 24]
    eng.encode_images(img)
    logits = [eng.prefill(torch.tensor([PROMPT] * 2), return_logits=True)]
    for j in range(forced.shape[1]):
        logits.append(eng.decode_step(forced[:, j]))
    got = torch.stack(logits, dim=1)                            # [2, 25, V]
    _as_accurate_as_bf16(got, g["tf_logits_bf16"], g["tf_logits_fp32"], slack=2.0, floor=3e-2)


def _greedy_contract(got, o16, img, prompt, stop_ids, n_new, **kw):
    """tests/parity.py: equal ids, or a flip at an oracle margin < MARGIN_TOL followed by a teacher-forced re-sync."""
    ref_new, ref_logits = oracle_greedy(o16, img, prompt, stop_ids, n_new, **kw)
    check_greedy_ids(got, ref_new, ref_logits, MARGIN_TOL, lambda ids: o16.teacher_forced_logits(img, prompt, ids),
                     eos_token_id=o16.eos_token_id, repetition_penalty=kw.get("repetition_penalty", 1.0))
    return ref_new


def test_greedy_ids_match_oracle(tiny):
    d, sd, eng, o16, o32, img = tiny
    n_new = 24
    eng.encode_images(img)
    eng.prefill(torch.tensor([PROMPT] * 2))
    got = eng.generate(GenerationParams(max_new_tokens=n_new, eos_token_id=0, pad_token_id=d.vocab - 4))
    _greedy_contract(got, o16, img, PROMPT, (), n_new)


def test_repetition_penalty_eos_and_row0_stop(tiny):
    """HF loop semantics (App. B): penalty on generated ids, EOS->pad, row-0 '</svg>' stops everyone."""
    d, sd, eng, o16, o32, img = tiny
    n_new = 20
    eng.encode_images(img)
    eng.prefill(torch.tensor([PROMPT] * 2))
    free = eng.generate(GenerationParams(max_new_tokens=n_new, repetition_penalty=3.1, eos_token_id=0,
                                         pad_token_id=d.vocab - 4))
    ref = _greedy_contract(free, o16, img, PROMPT, (), n_new, repetition_penalty=3.1)
    if torch.equal(free.cpu().long(), ref):
        stop = ref[0, 4:7].tolist()
        eng.encode_images(img)
        eng.prefill(torch.tensor([PROMPT] * 2))
        got = eng.generate(GenerationParams(max_new_tokens=n_new, repetition_penalty=3.1, eos_token_id=0,
                                            pad_token_id=d.vocab - 4, stop_ids=stop, poll_interval=1))
        _greedy_contract(got, o16, img, PROMPT, stop, n_new, repetition_penalty=3.1)
        assert got.shape[1] <= 7 + 0 or got.shape[1] < n_new


def test_untied_head_random_walk(tiny):
    """An un-tied random lm_head makes greedy a pseudo-random walk: exercises the margin contract."""
    d, sd, eng, o16, o32, img = tiny
    sd2 = dict(sd)
    g = torch.Generator().manual_seed(3)
    sd2["model.svg_transformer.transformer.lm_head.weight"] = (torch.randn(d.vocab, d.hidden, generator=g) * 0.2).to(torch.bfloat16)
    eng2 = Engine(d, 0)
    eng2.load_state_dict(sd2)
    o = OracleStarVector(d, sd2, dtype=torch.bfloat16, pad_token_id=d.vocab - 4)
    o.llm.lm_head.weight = torch.nn.Parameter(sd2["model.svg_transformer.transformer.lm_head.weight"].clone())
    eng2.encode_images(img)
    eng2.prefill(torch.tensor([PROMPT] * 2))
    got = eng2.generate(GenerationParams(max_new_tokens=32, eos_token_id=0, pad_token_id=d.vocab - 4))
    ref = _greedy_contract(got, o, img, PROMPT, (), 32)
    assert len(set(ref[0].tolist())) > 4, "walk degenerate: test lost its power"
    eng2.close()


def test_host_entry_equals_device_path(tiny):
    d, sd, eng, o16, o32, img = tiny
    p = GenerationParams(max_new_tokens=12, eos_token_id=None, pad_token_id=d.vocab - 4)
    eng.encode_images(img)
    eng.prefill(torch.tensor([PROMPT] * 2))
    a = eng.generate(p).cpu()
    b, n = eng.generate_im2svg_host(img.cpu().pin_memory(), torch.tensor([PROMPT] * 2, dtype=torch.int32), p)
    assert n == 12 and torch.equal(a, b)


def test_generate_is_deterministic_and_graph_replay_safe(tiny):
    d, sd, eng, o16, o32, img = tiny
    p = GenerationParams(max_new_tokens=40, eos_token_id=None, pad_token_id=d.vocab - 4)
    outs = []
    for _ in range(3):
        eng.encode_images(img)
        eng.prefill(torch.tensor([PROMPT] * 2))
        outs.append(eng.generate(p).cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    ms, steps = eng.last_decode_timing()
    assert steps == 39 and ms > 0 and eng.launch_count() > 0


def test_sampling_distribution(tiny):
    """do_sample: distribution-level parity only (different RNG): empirical first-token frequencies
    must match softmax(top-p-filtered logits / T) of the oracle's prefill logits."""
    d, sd, eng, o16, o32, img = tiny
    one = img[:1]
    T, top_p, n = 0.25, 0.9, 500
    eng.encode_images(one)
    logits = eng.prefill(torch.tensor([PROMPT]), return_logits=True)[0].float().cpu() / T
    probs = torch.softmax(logits, -1)
    sp, si = probs.sort(descending=True)
    keep = (sp.cumsum(0) - sp) < top_p
    expect = torch.zeros_like(probs)
    expect[si[keep]] = sp[keep] / sp[keep].sum()
    counts = torch.zeros_like(probs)
    for s in range(n):
        eng.encode_images(one)
        eng.prefill(torch.tensor([PROMPT]))
        t = eng.generate(GenerationParams(max_new_tokens=1, do_sample=True, temperature=T, top_p=top_p,
                                          eos_token_id=None, pad_token_id=d.vocab - 4, seed=1000 + s))
        counts[int(t[0, 0])] += 1
    # bf16 logits tie often; which member of a tie at the nucleus boundary survives is an artefact of HF's
    # sort order, so tokens tied with the smallest kept probability are allowed too
    p_min = sp[keep].min()
    outside = (expect == 0) & (probs < p_min * (1 - 1e-6))
    assert counts[outside].sum() == 0, "sampled a token outside the nucleus"
    tied = (expect == 0) & ~outside
    expect = expect * (1 - counts[tied].sum() / n)
    counts = counts.clone(); counts[tied] = 0; n = int(counts.sum())
    tv = 0.5 * (counts / n - expect).abs().sum().item()
    noise = 0.5 * (2 * expect / (3.14159 * n)).sqrt().sum().item()          # E|p_hat - p| summed over the support
    assert tv < 2.0 * noise + 0.02, f"total variation {tv:.3f} vs sampling noise {noise:.3f}"


def test_errors_are_python_exceptions(tiny):
    d, sd, eng, o16, o32, img = tiny
    with pytest.raises(ValueError):
        eng.encode_images(torch.zeros(1, 3, 10, 10))
    eng.encode_images(img)
    eng.prefill(torch.tensor([PROMPT] * 2))
    with pytest.raises(ValueError):
        eng.generate(GenerationParams(max_new_tokens=10 ** 6, pad_token_id=0))


def test_per_row_stop_mode(tiny):
    """`stop_row0_only=0` (the ABI's per-row mode; the reference's StoppingCriteriaSub only looks at row 0):
    every row ends at ITS OWN stop sequence and is padded afterwards, the call ends when all rows have ended.
    Oracle: the reference path run once per image (batch 1, where "row 0" is that image), re-rectangularised."""
    d, sd, eng, o16, o32, img = tiny
    n_new, pad = 24, d.vocab - 4
    free, _ = oracle_greedy(o16, img, PROMPT, (), n_new)
    stop = free[1, 3:5].tolist()                                   # a sequence row 1 emits early; row 0 may or may not
    rows = []
    for b in range(2):
        r, _ = oracle_greedy(o16, img[b:b + 1], PROMPT, stop, n_new)
        rows.append(r[0])
    width = max(len(r) for r in rows)
    ref = torch.full((2, width), pad, dtype=torch.long)
    for b, r in enumerate(rows):
        ref[b, :len(r)] = r
    def first_stop(seq):                                           # index after the first completion of `stop` in seq
        for e in range(len(stop), len(seq) + 1):
            if seq[e - len(stop):e] == stop:
                return e
        return len(seq)

    n1 = first_stop(free[1].tolist())
    assert len(rows[1]) == n1 <= 5                                 # row 1 stopped where the sequence first completes
    eng.encode_images(img)
    eng.prefill(torch.tensor([PROMPT] * 2))
    got = eng.generate(GenerationParams(max_new_tokens=n_new, eos_token_id=0, pad_token_id=pad, stop_ids=stop, stop_row0_only=False,
                                        poll_interval=1)).cpu().long()
    if got.shape == ref.shape:
        assert torch.equal(got, ref), (got.tolist(), ref.tolist())
    else:                                                          # a tolerated bf16 flip changed row 0's length: the stopped row must still be exact
        assert got[1, :n1].tolist() == rows[1].tolist() and (got[1, n1:] == pad).all()


@pytest.mark.parametrize("twin_below,near", [(True, False), (False, False), (False, True)])
def test_greedy_tie_takes_the_lowest_index(tiny, twin_below, near):
    """HF greedy = argmax over the bf16 logits cast to float, lowest index wins ties.  An un-tied lm_head
    with two IDENTICAL rows gives two exactly equal logits at every step: the engine must emit the smaller index, on the
    first token (prefill logits) and on every decode step (lm_head argmax partials), in both decode modes' shared epilogue."""
    d, sd, eng, o16, o32, img = tiny
    eng.encode_images(img[:1])
    top = int(eng.prefill(torch.tensor([PROMPT]), return_logits=True)[0].float().argmax())
    twin = top - 3 if (twin_below and top >= 3) else top + 5
    assert 0 < twin < d.vocab - 8
    head = sd["model.svg_transformer.transformer.transformer.wte.weight"].clone()
    head[twin] = head[top]
    if near:
        # a NEAR twin above the top row: a few elements differ by one bf16 ulp, so the fp32 dot products differ (either way,
        # depending on the hidden state) while the bf16 logits still round to the same value almost always; HF compares the
        # ROUNDED logits, so the lower index must still win wherever they are equal (an argmax over the fp32 accumulators
        # would follow the larger pre-rounding value)
        row = head[top].clone()
        idx = row.abs().argsort()[8:16]
        bits = row.view(torch.int16)
        bits[idx] = bits[idx] + 1
        head[twin] = bits.view(torch.bfloat16)
    sd2 = dict(sd)
    sd2["model.svg_transformer.transformer.lm_head.weight"] = head
    eng2 = Engine(d, 0)
    eng2.load_state_dict(sd2)
    eng2.encode_images(img[:1])
    lg = eng2.prefill(torch.tensor([PROMPT]), return_logits=True)[0].float()
    if not near:
        assert lg[twin] == lg[top] == lg.max()
    n_new = 24 if near else 6
    got = eng2.generate(GenerationParams(max_new_tokens=n_new, eos_token_id=None, pad_token_id=d.vocab - 4)).cpu()
    if lg[twin] == lg[top] == lg.max():
        assert int(got[0, 0]) == min(top, twin)
    # every later step: wherever the two twins are the maximum, the smaller index must have been chosen
    eng2.encode_images(img[:1])
    eng2.prefill(torch.tensor([PROMPT]))
    for s in range(n_new - 1):
        step_logits = eng2.decode_step(got[:, s])[0].float()
        assert int(got[0, s + 1]) == int(step_logits.argmax()), (s, int(got[0, s + 1]))   # first maximum of the bf16 logits
        if step_logits[top] == step_logits[twin] == step_logits.max():
            assert int(got[0, s + 1]) == min(top, twin), (s, int(got[0, s + 1]))
    eng2.close()
