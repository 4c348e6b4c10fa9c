"""CPU check of the beam-search bookkeeping (starvector_b200/beam_search.py) against HF generate(num_beams>1).

The bookkeeping only needs four engine calls; here they are served by the CPU oracle's HF decoder (with its own KV
cache), so any divergence from `GenerationMixin._beam_search` is caught without a GPU.  The GPU test
(tests/test_beam_gpu.py) then runs the same function on the real engine.
"""
import pytest
import torch

from oracle.pipeline import OracleStarVector
from starvector_b200.beam_search import beam_search
from starvector_b200.config import dims_tiny
from starvector_b200.weights import synthetic_images, synthetic_state_dict

PROMPT = [44, 78]


class OracleBackedEngine:
    """encode_images / prefill / decode_step / reorder_cache on top of the oracle's CPU modules."""

    def __init__(self, oracle):
        self.o, self.dims = oracle, oracle.dims

    def encode_images(self, image):
        self.embeds = self.o.image_projection(self.o.image_encoder(image.to(self.o.dtype)))

    def prefill(self, prompt_ids, return_logits=False):
        x = torch.cat([self.embeds, self.o.llm.transformer.wte(prompt_ids.long())], dim=1)
        out = self.o.llm(inputs_embeds=x, use_cache=True)
        self.cache = out.past_key_values
        return out.logits[:, -1, :].float()

    def decode_step(self, ids):
        out = self.o.llm(input_ids=ids.long().view(-1, 1), past_key_values=self.cache, use_cache=True)
        self.cache = out.past_key_values
        return out.logits[:, -1, :].float()

    def reorder_cache(self, idx):
        self.cache.reorder_cache(idx.long())


@pytest.fixture(scope="module")
def setup():
    torch.set_num_threads(1)
    d = dims_tiny(max_batch=8)
    sd = dict(synthetic_state_dict(d, seed=0, init="randomized"))
    g = torch.Generator().manual_seed(3)          # un-tied random head: a non-degenerate search space
    sd["model.svg_transformer.transformer.lm_head.weight"] = (torch.randn(d.vocab, d.hidden, generator=g) * 0.2).to(torch.bfloat16)
    o = OracleStarVector(d, sd, dtype=torch.float32, pad_token_id=d.vocab - 4)
    o.llm.lm_head.weight = torch.nn.Parameter(sd["model.svg_transformer.transformer.lm_head.weight"].float())
    return d, o, synthetic_images(d, 2, seed=1).float()


@pytest.mark.parametrize("nb,lp,rp,stop", [(2, 1.0, 1.0, ()), (3, 1.0, 1.0, ()), (2, -1.0, 3.1, ()), (2, 1.0, 1.0, "row0"),
                                           (2, 2.0, 1.0, ())])
def test_beam_search_matches_hf(setup, nb, lp, rp, stop):
    d, o, img = setup
    n_new = 14
    kw = dict(use_nucleus_sampling=False, num_beams=nb, length_penalty=lp, repetition_penalty=rp,
              max_length=d.query_length + len(PROMPT) + n_new)
    stop_ids = ()
    if stop == "row0":
        base = o.generate_im2svg_ids(img, PROMPT, (), **kw)
        stop_ids = tuple(base[0, 2 + 4: 2 + 6].tolist())
    ref = o.generate_im2svg_ids(img, PROMPT, stop_ids, **kw)[:, len(PROMPT):]
    got = beam_search(OracleBackedEngine(o), img, torch.tensor([PROMPT] * 2), num_beams=nb, max_new_tokens=n_new,
                      repetition_penalty=rp, length_penalty=lp, early_stopping=True, eos_token_id=0,
                      pad_token_id=d.vocab - 4, stop_ids=stop_ids)
    assert got.shape == ref.shape and torch.equal(got, ref), (got.tolist(), ref.tolist())


@pytest.mark.parametrize("nb,lp", [(2, 1.0), (3, 1.0), (2, -1.0)])
def test_beam_search_without_early_stopping_matches_hf(setup, nb, lp):
    """v2 passes no `early_stopping` (starvector_v2.py:53-57 returns {}), so HF's default False applies: the loop runs on
    until the best running beam can no longer beat the worst finished one."""
    import warnings

    d, o, img = setup
    n_new = 14
    emb, mask, _ = o.prepare_generation_inputs(img, PROMPT)
    kw = o.generation_kwargs({"inputs_embeds": emb, "attention_mask": mask, "use_nucleus_sampling": False, "num_beams": nb,
                              "length_penalty": lp, "max_length": d.query_length + len(PROMPT) + n_new}, ())
    kw.pop("top_p"); kw.pop("temperature")
    kw["early_stopping"] = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = o.llm.generate(**kw)
    got = beam_search(OracleBackedEngine(o), img, torch.tensor([PROMPT] * 2), num_beams=nb, max_new_tokens=n_new,
                      length_penalty=lp, early_stopping=False, eos_token_id=0, pad_token_id=d.vocab - 4)
    assert got.shape == ref.shape and torch.equal(got, ref), (got.tolist(), ref.tolist())
