"""The reference-facing Python surface (StarVectorForCausalLM.generate_im2svg / .generate)."""
import warnings

import pytest
import torch

from oracle.pipeline import OracleStarVector
from starvector_b200.config import dims_tiny
from starvector_b200.modeling import StarVectorForCausalLM
from starvector_b200.weights import synthetic_images, synthetic_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    d = dims_tiny()
    sd = synthetic_state_dict(d, seed=0, init="randomized")
    m = StarVectorForCausalLM.from_config(dims=d, state_dict=sd)
    yield d, sd, m
    m.model.engine.close()


def test_quickstart_shaped_call(model):
    d, sd, m = model
    m.cuda(); m.eval()
    img = synthetic_images(d, 1, seed=1)
    batch = {"image": img.to(torch.float16).cuda()}
    # literal quickstart.py:19 kwargs: defaults add do_sample=True, top_p=0.9, num_beams=2 -> beam-sample
    out = m.generate_im2svg(batch, max_length=d.query_length + 2 + 16, temperature=1.5, length_penalty=-1,
                            repetition_penalty=3.1)
    assert isinstance(out, list) and len(out) == 1 and out[0].startswith("<svg")


def test_process_images_then_generate_like_quickstart(model):
    """quickstart.py:15-19: `process_images([pil])[0].to(float16).cuda()` -> generate_im2svg; pixels == the reference recipe."""
    from PIL import Image

    from oracle import preprocess as P

    d, sd, m = model
    arr = P.synthetic_image(90, 61, 4, seed=3)
    pix = m.process_images([Image.fromarray(arr, "RGBA")])
    assert isinstance(pix, list) and pix[0].shape == (1, 3, d.image_size, d.image_size) and pix[0].is_cuda
    assert torch.equal(pix[0][0].cpu(), P.reference_transform(arr, d.image_size, P.ALPHA_WHITE))
    image = pix[0].to(torch.float16).cuda()
    out = m.generate_im2svg({"image": image}, use_nucleus_sampling=False, num_beams=1, max_length=d.query_length + 2 + 8)
    assert len(out) == 1 and out[0].startswith("<svg")


def test_greedy_strings_match_oracle(model):
    d, sd, m = model
    img = synthetic_images(d, 2, seed=1)
    tok = m.model.svg_transformer.tokenizer
    prompt = tok("<svg")["input_ids"]
    kw = dict(use_nucleus_sampling=False, num_beams=1, max_length=d.query_length + len(prompt) + 12)
    got = m.model.generate_im2svg({"image": img.cuda()}, **kw)
    o = OracleStarVector(d, sd, dtype=torch.bfloat16, pad_token_id=tok.pad_token_id)
    ref = o.generate_im2svg_ids(img, prompt, tok("</svg>")["input_ids"], **kw)
    assert got == tok.batch_decode(ref, skip_special_tokens=True)


def test_transformer_generate_with_inputs_embeds(model):
    d, sd, m = model
    img = synthetic_images(d, 2, seed=1)
    emb, _ = m.model.engine.encode_images(img, return_embeds=True)
    prompt = torch.tensor([m.model.svg_transformer.tokenizer("<svg")["input_ids"]] * 2)
    inputs_embeds = torch.cat([emb, m.model._get_embeddings(prompt)], dim=1)
    a = m.model.svg_transformer.transformer.generate(
        inputs_embeds=inputs_embeds, attention_mask=torch.ones(inputs_embeds.shape[:2], dtype=torch.long),
        do_sample=False, num_beams=1, max_length=inputs_embeds.shape[1] + 10)
    b = m.model.generate_im2svg_ids({"image": img}, use_nucleus_sampling=False, num_beams=1,
                                    max_length=inputs_embeds.shape[1] + 10)
    assert torch.equal(a.cpu(), b[:, prompt.shape[1]:].cpu())


def test_max_length_shorter_than_prefix_raises(model):
    d, sd, m = model
    with pytest.raises(ValueError, match="max_length"):
        m.generate_im2svg({"image": synthetic_images(d, 1)}, num_beams=1, max_length=d.query_length)   # < Q + P
