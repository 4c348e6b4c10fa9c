"""Widening rows of SURVEY.md §8f-4 on the GPU: token streaming (`sv_generate_stream`, the facade's `streamer=` kwarg the
reference's serving worker passes), `num_return_sequences` (`generate_im2svg_grpo`), batches above `max_batch`.  These are
self-consistency tests of host-side control flow around kernels whose parity the other GPU test files establish."""
import os

import pytest
import torch

from oracle import preprocess as PRE
from starvector_b200.config import dims_tiny
from starvector_b200.engine import Engine, GenerationParams
from starvector_b200.modeling import StarVectorForCausalLM
from starvector_b200.preprocess import SiglipImageProcessor
from starvector_b200.weights import synthetic_images, synthetic_state_dict

pytestmark = pytest.mark.gpu
PROMPT = [44, 78]


@pytest.fixture(scope="module")
def engine():
    d = dims_tiny()
    eng = Engine(d, 0)
    eng.load_state_dict(synthetic_state_dict(d, seed=0, init="randomized"))
    yield d, eng, synthetic_images(d, 2, seed=1)
    eng.close()


@pytest.fixture(scope="module")
def model():
    d = dims_tiny()
    sd = synthetic_state_dict(d, seed=0, init="randomized")
    m = StarVectorForCausalLM.from_config(dims=d, state_dict=sd)
    yield d, sd, m
    m.model.engine.close()


def test_streaming_callback_sees_exactly_the_returned_tokens(engine):
    """sv_generate_stream: chunks arrive in order, their concatenation is the rectangle sv_generate returns; a truthy
    return value cancels at the next poll; an exception in the callback cancels and propagates."""
    d, eng, img = engine
    p = GenerationParams(max_new_tokens=45, eos_token_id=None, pad_token_id=d.vocab - 4, poll_interval=8)

    def run(cb=None, params=p):
        eng.encode_images(img)
        eng.prefill(torch.tensor([PROMPT] * 2))
        return eng.generate(params, on_tokens=cb).cpu()

    plain = run()
    chunks = []
    streamed = run(lambda ids, first: chunks.append((first, ids.clone())) and False)
    assert torch.equal(streamed, plain)
    assert [c[0] for c in chunks] == [sum(x[1].shape[1] for x in chunks[:i]) for i in range(len(chunks))]
    assert len(chunks) >= 5 and torch.equal(torch.cat([c[1] for c in chunks], dim=1), plain)
    # with a stop condition armed the same holds (row 0 stop after 7 tokens)
    stop = plain[0, 4:7].tolist()
    ps = GenerationParams(max_new_tokens=45, eos_token_id=None, pad_token_id=d.vocab - 4, poll_interval=4, stop_ids=stop)
    want = run(params=ps)
    chunks.clear()
    got = run(lambda ids, first: chunks.append((first, ids.clone())) and False, params=ps)
    assert torch.equal(got, want) and torch.equal(torch.cat([c[1] for c in chunks], dim=1), want)
    # cancel
    seen = []
    cut = run(lambda ids, first: seen.append(ids.shape[1]) or sum(seen) >= 16)
    assert 16 <= cut.shape[1] < 45 and torch.equal(cut, plain[:, : cut.shape[1]])

    def boom(ids, first):
        raise KeyError("from the callback")

    with pytest.raises(KeyError):
        run(boom)
    assert torch.equal(run(), plain)                           # the engine is usable afterwards


def test_grpo_num_return_sequences(model):
    """starvector_base.py:261-286: G completions per image = rows repeated adjacently, sampled independently."""
    d, sd, m = model
    img = synthetic_images(d, 2, seed=1).cuda()
    P = len(m.model.svg_transformer.tokenizer("<svg")["input_ids"])
    kw = dict(max_length=d.query_length + P + 10)
    two = m.model.generate_im2svg_grpo({"image": img}, use_nucleus_sampling=False, num_return_sequences=2, **kw)
    assert two["outputs"].shape[0] == 4 and len(two["raw_svg"]) == 4
    assert torch.equal(two["outputs"][0], two["outputs"][1]) and torch.equal(two["outputs"][2], two["outputs"][3])
    one = m.model.generate_im2svg_grpo({"image": img}, use_nucleus_sampling=False, num_beams=1, **kw)   # (G = 1 keeps the default 2 beams, :277-280)
    assert torch.equal(one["outputs"], two["outputs"][::2])
    assert two["inputs_embeds"].shape == (2, d.query_length + P, d.hidden)
    s = m.model.generate_im2svg_grpo({"image": img[:1]}, num_return_sequences=4, temperature=2.0, top_p=1.0, **kw)
    assert s["outputs"].shape[0] == 4
    assert len({tuple(r.tolist()) for r in s["outputs"]}) >= 3          # independent samples per copy
    with pytest.raises(ValueError):
        m.model.generate_im2svg_grpo({"image": img}, num_return_sequences=3, **kw)      # 6 rows > max_batch 4


def test_forward_scoring_matches_oracle(model):
    """`StarVectorForCausalLM.forward` (starvector_arch.py:161-184): logits of G completions per image over a shared visual
    prefix (prefilled once, KV rows replicated in `.repeat` order), against the oracle's one full forward."""
    from oracle.pipeline import OracleStarVector

    d, sd, m = model
    img = synthetic_images(d, 2, seed=1).cuda()
    grpo = m.model.generate_im2svg_grpo({"image": img}, use_nucleus_sampling=False, num_beams=1, max_length=d.query_length + 2 + 1)
    vision_embeds = grpo["inputs_embeds"]                                   # [2, Q+P, H], the reference's hand-over to forward()
    G, T, keep = 2, 9, 5
    gen = torch.Generator().manual_seed(11)
    ids = torch.randint(1, d.vocab - 8, (2 * G, T), generator=gen)
    mask = torch.ones(2 * G, vision_embeds.shape[1] + T, dtype=torch.long)
    mask[1, -2:] = 0                                                         # a right-padded completion is accepted
    out = m(vision_embeds, ids, G, mask, keep)
    assert out.loss is None and out.logits.shape == (2 * G, keep, d.vocab)
    full = m.forward(vision_embeds, ids, num_generations=G, attention_mask=None, num_logits_to_keep=0).logits
    assert full.shape == (2 * G, T, d.vocab) and torch.equal(full[:, -keep:], out.logits)
    refs = {}
    for dt in (torch.bfloat16, torch.float32):
        o = OracleStarVector(d, sd, dtype=dt, pad_token_id=d.vocab - 4)
        emb = torch.cat([vision_embeds.cpu().to(dt).repeat(G, 1, 1), o.llm.transformer.wte(ids)], dim=1)
        with torch.no_grad():
            refs[dt] = o.llm(inputs_embeds=emb, use_cache=False).logits[:, -T:].float()
    e = (full.cpu() - refs[torch.float32]).abs()
    o_err = (refs[torch.bfloat16] - refs[torch.float32]).abs()
    assert e.max().item() <= 2.0 * o_err.max().item() + 3e-2 and e.mean().item() <= 2.0 * o_err.mean().item() + 3e-3, \
        (e.max().item(), o_err.max().item(), e.mean().item(), o_err.mean().item())
    with pytest.raises(NotImplementedError):
        left = mask.clone(); left[0, 0] = 0
        m(vision_embeds, ids, G, left, keep)
    with pytest.raises(ValueError):
        m(vision_embeds, ids[:3], G, None, keep)


def test_grpo_shares_the_visual_prefix(model):
    """num_return_sequences = G encodes and prefills every image ONCE (its KV rows are replicated by `sv_expand_batch`), never
    the G-times expanded batch: the ViT and the prefill only ever see B rows, and each of the G greedy completions of an image
    equals the single completion of a plain B-row call (same prefill batch, so bit-identical prefix; decode rows are
    independent of each other)."""
    d, sd, m = model
    img = synthetic_images(d, 2, seed=1).cuda()
    P = len(m.model.svg_transformer.tokenizer("<svg")["input_ids"])
    kw = dict(use_nucleus_sampling=False, max_length=d.query_length + P + 6)
    eng = m.model.engine
    calls = []
    enc, pre = eng.encode_images, eng.prefill

    def rec_encode(px, *a, **k):
        calls.append(("encode", int(px.shape[0])))
        return enc(px, *a, **k)

    def rec_prefill(ids, *a, **k):
        calls.append(("prefill", int(ids.shape[0])))
        return pre(ids, *a, **k)

    eng.encode_images, eng.prefill = rec_encode, rec_prefill
    try:
        shared = m.model.generate_im2svg_grpo({"image": img}, num_return_sequences=2, **kw)["outputs"].cpu()
    finally:
        del eng.encode_images, eng.prefill                      # drop the instance attributes: the class methods are back
    assert ("prefill", 2) in calls and ("encode", 2) in calls and all(n == 2 for _, n in calls), calls
    one = m.model.generate_im2svg_ids({"image": img}, num_beams=1, **kw).cpu()
    assert shared.shape[0] == 4 and shared.shape[1] == one.shape[1], (shared.shape, one.shape)
    assert torch.equal(shared[0::2], one) and torch.equal(shared[1::2], one), (shared.tolist(), one.tolist())


def test_more_images_than_max_batch_runs_in_groups(model):
    """6 images on an engine that holds 4: two groups, merged with the row-0 stop rule of the sharded path."""
    from starvector_b200.parallel import merge_generated

    d, sd, m = model
    img = synthetic_images(d, 6, seed=4).cuda()
    tok = m.model.svg_transformer.tokenizer
    P = len(tok("<svg")["input_ids"])
    kw = dict(use_nucleus_sampling=False, num_beams=1, max_length=d.query_length + P + 12)
    all6 = m.model.generate_im2svg_ids({"image": img}, **kw)
    first = m.model.generate_im2svg_ids({"image": img[:4]}, **kw)[:, P:]
    rest = m.model.generate_im2svg_ids({"image": img[4:]}, stop_ids=(), **kw)[:, P:]
    want = merge_generated([first, rest], tok("</svg>")["input_ids"], tok.pad_token_id)
    assert all6.shape[0] == 6 and torch.equal(all6[:, P:], want)
    assert len(m.model.generate_im2svg({"image": img}, **kw)) == 6


def test_streamer_kwarg_streams_tokens(model):
    """serve/model_worker.py:131-181: generate in a thread with `streamer=`, iterate the streamer for text."""
    from threading import Thread

    from transformers import TextIteratorStreamer

    d, sd, m = model
    tok = m.model.svg_transformer.tokenizer
    P = len(tok("<svg")["input_ids"])
    img = synthetic_images(d, 1, seed=1).cuda()
    kw = dict(use_nucleus_sampling=False, num_beams=1, max_length=d.query_length + P + 20)

    class Collect:
        def __init__(self):
            self.tokens, self.ended = [], False

        def put(self, value):
            self.tokens.append(value.clone())

        def end(self):
            self.ended = True

    c = Collect()
    ids = m.model.generate_im2svg_ids({"image": img}, streamer=c, **kw)
    assert c.ended and all(t.shape == (1,) for t in c.tokens)
    assert torch.equal(torch.stack(c.tokens, dim=1), ids[:, P:].cpu())
    assert torch.equal(ids, m.model.generate_im2svg_ids({"image": img}, **kw))

    streamer = TextIteratorStreamer(tok, skip_prompt=False, skip_special_tokens=True, timeout=60)
    result = {}
    thread = Thread(target=lambda: result.update(text=m.model.generate_im2svg(batch={"image": img}, streamer=streamer, **kw)))
    thread.start()
    pieces = [piece for piece in streamer]
    thread.join()
    assert "".join(pieces) == tok.decode(ids[0, P:].tolist(), skip_special_tokens=True) and len(result["text"]) == 1
    with pytest.raises(ValueError):
        m.model.generate_im2svg({"image": img}, streamer=Collect(), num_beams=2, max_length=kw["max_length"])


def test_siglip_golden_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "preprocess_v1.pt"), weights_only=False)
    proc = SiglipImageProcessor(size=384)
    for case in g["siglip_cases"]:
        a = PRE.synthetic_image(*case["hwc"], seed=case["seed"])
        got = proc(images=a).pixel_values[0].cpu()
        assert PRE.tensor_sha256(got) == case["sha256_f32"] and PRE.tensor_sha256(got.to(torch.bfloat16)) == case["sha256_bf16"]
