"""CPU oracle of ``StarVectorBase.generate_im2svg`` (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates, line for line, reference starvector/model/models/starvector_base.py:
``_prepare_generation_inputs`` (:203-221), ``_get_generation_kwargs`` (:223-241),
``_get_im2svg_specific_kwargs`` (:289-295), ``StoppingCriteriaSub`` (:9-20) and the
``generate`` + concat at :255-256.  Tokenizer calls are replaced by explicit id lists (no
tokenizer files exist offline): ``prompt_ids`` stands for ``tokenizer('<svg')`` and
``stop_ids`` for ``tokenizer('</svg>')``.

The decoder is the installed ``transformers`` ``GPTBigCodeForCausalLM`` — the class the
reference loads by name at llm/starcoder.py:33 — driven through ``GenerationMixin.generate``.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Sequence

import torch

from . import vision

VIS = "model.image_encoder.visual_encoder."
LNV = "model.image_encoder.ln_vision."
ADP = "model.image_projection."
LLM = "model.svg_transformer.transformer."


def _sub(sd: Dict[str, torch.Tensor], prefix: str, dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        if k.startswith(prefix):
            out[k[len(prefix):]] = v.to(dtype) if v.is_floating_point() else v
    return out


def build_hf_decoder(dims, sd: Dict[str, torch.Tensor], dtype: torch.dtype, eos_token_id: int, pad_token_id: int):
    """GPTBigCodeForCausalLM with starcoderbase-1b's structure (llm/starcoder.py:16-34)."""
    from transformers import GPTBigCodeConfig, GPTBigCodeForCausalLM

    cfg = GPTBigCodeConfig(
        vocab_size=dims.vocab, n_positions=dims.n_positions, n_embd=dims.hidden, n_layer=dims.n_layer,
        n_head=dims.n_head, n_inner=dims.n_inner, activation_function="gelu_pytorch_tanh",
        resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=dims.ln_eps,
        scale_attn_weights=True, use_cache=True, attention_softmax_in_fp32=True,
        scale_attention_softmax_in_fp32=True, multi_query=True,
        bos_token_id=eos_token_id, eos_token_id=eos_token_id, pad_token_id=pad_token_id,  # starcoder.py:22-24
    )
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.device("meta"):
            model = GPTBigCodeForCausalLM(cfg)
        model = model.to_empty(device="cpu").to(dtype)
    llm = _sub(sd, LLM, dtype)
    llm.setdefault("lm_head.weight", llm["transformer.wte.weight"])
    missing, unexpected = model.load_state_dict(llm, strict=False, assign=True)
    missing = [m for m in missing if not m.endswith(".attn.bias") and not m.endswith("transformer.bias")]
    if missing or unexpected:
        raise RuntimeError(f"decoder state dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
    model.lm_head.weight = model.transformer.wte.weight          # tied (train/util.py:68-77)
    # non-persistent causal-mask buffer was created on meta; rebuild it
    n = cfg.max_position_embeddings
    model.transformer.register_buffer("bias", torch.tril(torch.ones((n, n), dtype=torch.bool)), persistent=False)
    model.eval()
    model.generation_config.eos_token_id = eos_token_id
    model.generation_config.bos_token_id = eos_token_id
    model.generation_config.pad_token_id = pad_token_id
    return model


def _make_stopping_criteria(stop_ids: Sequence[int]):
    """StoppingCriteriaSub (starvector_base.py:9-20): ROW 0 ONLY, a python bool for the whole batch."""
    from transformers.generation.stopping_criteria import StoppingCriteria, StoppingCriteriaList

    stops = [list(stop_ids)] if len(stop_ids) else []

    class _RowZeroStop(StoppingCriteria):
        def __call__(self, input_ids, scores, **kw):
            for s in stops:
                if input_ids[0][-len(s):].tolist() == s:
                    return True
            return False

    return StoppingCriteriaList([_RowZeroStop()])


class OracleStarVector:
    """Reference-equivalent StarVector-v1 (CLIP ViT + Adapter + GPTBigCode) on CPU."""

    def __init__(self, dims, state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.bfloat16,
                 eos_token_id: Optional[int] = 0, pad_token_id: int = 49152):
        self.dims = dims
        self.dtype = dtype
        self.adapter_norm = {0: "layer_norm", 1: "batch_norm"}[dims.adapter_norm]
        self.vis = _sub(state_dict, VIS, dtype)
        self.lnv = _sub(state_dict, LNV, dtype)
        self.adp = _sub(state_dict, ADP, dtype)
        self.eos_token_id = eos_token_id
        self.pad_token_id = pad_token_id
        self.llm = build_hf_decoder(dims, state_dict, dtype, 0 if eos_token_id is None else eos_token_id, pad_token_id)

    # -- stages --------------------------------------------------------------------------
    @torch.no_grad()
    def image_encoder(self, image: torch.Tensor) -> torch.Tensor:
        return vision.image_encoder_forward(image, self.vis, self.lnv, patch=self.dims.patch_size,
                                            heads=self.dims.vit_heads, layers=self.dims.vit_layers)

    @torch.no_grad()
    def image_projection(self, embedded: torch.Tensor) -> torch.Tensor:
        return vision.adapter_forward(embedded, self.adp, self.adapter_norm)

    @torch.no_grad()
    def prepare_generation_inputs(self, image: torch.Tensor, prompt_ids: Sequence[int]):
        """starvector_base.py:203-221."""
        image = image.to(self.dtype)                                                 # :206
        embedded_image = self.image_encoder(image)                                   # :208
        embedded_image = self.image_projection(embedded_image)                       # :209
        embedded_att = torch.ones(embedded_image.size()[:-1], dtype=torch.long)      # :210
        prompt = torch.tensor([list(prompt_ids)] * image.size(0), dtype=torch.long)  # :213-216
        attention_mask = torch.cat([embedded_att, torch.ones_like(prompt)], dim=1)   # :217
        inputs_embeds = self.llm.transformer.wte(prompt)                             # :218 (v1:16-18)
        inputs_embeds = torch.cat([embedded_image, inputs_embeds], dim=1)            # :219
        return inputs_embeds, attention_mask, prompt

    def generation_kwargs(self, base: dict, stop_ids: Sequence[int]) -> dict:
        """starvector_base.py:223-241 + :289-295."""
        kw = {
            "inputs_embeds": base["inputs_embeds"],
            "attention_mask": base["attention_mask"],
            "do_sample": base.get("use_nucleus_sampling", True),
            "top_p": base.get("top_p", 0.9),
            "temperature": base.get("temperature", 1),
            "num_beams": base.get("num_beams", 2),
            "max_length": base.get("max_length", 30),
            "min_length": base.get("min_length", 1),
            "repetition_penalty": base.get("repetition_penalty", 1.0),
            "length_penalty": base.get("length_penalty", 1.0),
            "use_cache": base.get("use_cache", True),
            "stopping_criteria": _make_stopping_criteria(stop_ids),
        }
        kw.update({"early_stopping": True, "pad_token_id": self.pad_token_id})
        return kw

    # -- the path ------------------------------------------------------------------------
    @torch.no_grad()
    def generate_im2svg_ids(self, image: torch.Tensor, prompt_ids: Sequence[int], stop_ids: Sequence[int] = (),
                            return_logits: bool = False, **kwargs):
        """starvector_base.py:243-256 up to (not including) tokenizer.batch_decode.

        Returns ``LongTensor [B, P + n_new]`` (prompt ids followed by the generated ids); with
        ``return_logits`` also the per-step fp32 logits ``[n_new, B, V]`` HF selected from.
        """
        inputs_embeds, attention_mask, prompt = self.prepare_generation_inputs(image, prompt_ids)
        kw = self.generation_kwargs({**kwargs, "inputs_embeds": inputs_embeds, "attention_mask": attention_mask},
                                    stop_ids)
        if self.eos_token_id is None:
            # throughput configs disable EOS so exactly max_new tokens come out
            kw["eos_token_id"] = None
            self.llm.generation_config.eos_token_id = None
        if not kw["do_sample"]:
            kw.pop("top_p"); kw.pop("temperature")
        if kw["num_beams"] == 1:
            kw.pop("early_stopping", None); kw.pop("length_penalty")
        if return_logits:
            kw.update(output_logits=True, return_dict_in_generate=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = self.llm.generate(**kw)                                           # :255
        seq = out.sequences if return_logits else out
        ids = torch.cat([prompt, seq], dim=1)                                       # :256
        if return_logits:
            return ids, torch.stack([l.float() for l in out.logits], dim=0)
        return ids

    @torch.no_grad()
    def teacher_forced_logits(self, image: torch.Tensor, prompt_ids: Sequence[int],
                              forced: torch.Tensor) -> torch.Tensor:
        """Logits `[B, n+1, V]` (fp32): position j is the distribution for generated token j
        given the visual prefix, the prompt and ``forced[:, :j]`` — one full forward, no cache."""
        inputs_embeds, attention_mask, _ = self.prepare_generation_inputs(image, prompt_ids)
        t0 = inputs_embeds.shape[1]
        emb = torch.cat([inputs_embeds, self.llm.transformer.wte(forced)], dim=1)
        out = self.llm(inputs_embeds=emb, use_cache=False)
        return out.logits[:, t0 - 1:, :].float()


    @torch.no_grad()
    def teacher_forced_logits_at(self, image: torch.Tensor, prompt_ids: Sequence[int], forced: torch.Tensor,
                                 steps: Sequence[int]) -> torch.Tensor:
        """`teacher_forced_logits(...)[:, steps]` without materialising the other positions' logits (long contexts):
        one full forward of the decoder body, lm_head on the selected positions only.  `[B, len(steps), V]` fp32."""
        inputs_embeds, _, _ = self.prepare_generation_inputs(image, prompt_ids)
        t0 = inputs_embeds.shape[1]
        emb = torch.cat([inputs_embeds, self._embed(forced)], dim=1)
        hidden = self._body(inputs_embeds=emb, use_cache=False).last_hidden_state
        idx = torch.tensor([t0 - 1 + int(j) for j in steps])
        return self.llm.lm_head(hidden[:, idx]).float()

    def _embed(self, ids):
        return self.llm.transformer.wte(ids)

    @property
    def _body(self):
        return self.llm.transformer


# ======================================================================================================
# StarVector v2 (8B family): SigLIP vision tower + Adapter + StarCoder2 — reference models/starvector_v2.py,
# image_encoder.py:32-48,108-109 and llm/starcoder2.py:19-32.  Both towers are the installed transformers
# classes the reference loads by name; only the glue is restated.
def build_hf_siglip(dims, sd, dtype):
    """`AutoModel.from_pretrained("google/siglip-...").vision_model` with random-init weights of our state dict."""
    from transformers import SiglipVisionConfig, SiglipVisionModel

    cfg = SiglipVisionConfig(hidden_size=dims.vit_width, intermediate_size=dims.vit_mlp, num_hidden_layers=dims.vit_layers,
                             num_attention_heads=dims.vit_heads, image_size=dims.image_size, patch_size=dims.patch_size,
                             hidden_act="gelu_pytorch_tanh", layer_norm_eps=dims.vit_ln_eps, attention_dropout=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vm = SiglipVisionModel(cfg).vision_model.to(dtype)
    vis = _sub(sd, VIS, dtype)
    missing, unexpected = vm.load_state_dict(vis, strict=False)
    missing = [m for m in missing if not m.startswith("head.")]      # pooling head: output discarded by the reference
    if missing or unexpected:
        raise RuntimeError(f"siglip state dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
    return vm.eval()


def build_hf_starcoder2(dims, sd, dtype, eos_token_id):
    from transformers import Starcoder2Config, Starcoder2ForCausalLM

    cfg = Starcoder2Config(
        vocab_size=dims.vocab, hidden_size=dims.hidden, intermediate_size=dims.n_inner, num_hidden_layers=dims.n_layer,
        num_attention_heads=dims.n_head, num_key_value_heads=dims.n_kv_head, hidden_act="gelu_pytorch_tanh",
        max_position_embeddings=dims.n_positions, norm_epsilon=dims.ln_eps, use_cache=True,
        bos_token_id=eos_token_id, eos_token_id=eos_token_id, sliding_window=dims.sliding_window or None, use_bias=True,
        rope_parameters={"rope_type": "default", "rope_theta": float(dims.rope_theta)},
        residual_dropout=0.0, embedding_dropout=0.0, attention_dropout=0.0,
    )
    llm = _sub(sd, LLM, dtype)
    tied = "lm_head.weight" not in llm or torch.equal(llm["lm_head.weight"], llm["model.embed_tokens.weight"])
    cfg.tie_word_embeddings = tied                       # an explicit, different lm_head (tests) stays un-tied
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = Starcoder2ForCausalLM(cfg).to(dtype)
    llm.setdefault("lm_head.weight", llm["model.embed_tokens.weight"])
    missing, unexpected = model.load_state_dict(llm, strict=False)
    if missing or unexpected:
        raise RuntimeError(f"starcoder2 state dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
    if tied:
        model.lm_head.weight = model.model.embed_tokens.weight
    model.eval()
    model.generation_config.eos_token_id = eos_token_id
    model.generation_config.bos_token_id = eos_token_id
    model.generation_config.pad_token_id = None        # v2 passes no pad id: HF falls back to eos (starvector_v2.py:53-57)
    return model


class OracleStarVectorV2(OracleStarVector):
    """Reference-equivalent StarVector v2 on CPU (models/starvector_v2.py)."""

    def __init__(self, dims, state_dict, dtype=torch.bfloat16, eos_token_id: Optional[int] = 0):
        self.dims = dims
        self.dtype = dtype
        self.adapter_norm = {0: "layer_norm", 1: "batch_norm"}[dims.adapter_norm]
        self.adp = _sub(state_dict, ADP, dtype)
        self.eos_token_id = eos_token_id
        self.pad_token_id = None
        self.vision = build_hf_siglip(dims, state_dict, dtype)
        self.llm = build_hf_starcoder2(dims, state_dict, dtype, 0 if eos_token_id is None else eos_token_id)

    @torch.no_grad()
    def image_encoder(self, image):
        return self.vision(image)["last_hidden_state"]                              # image_encoder.py:108-109

    @torch.no_grad()
    def prepare_generation_inputs(self, image, prompt_ids):
        image = image.to(self.dtype)
        embedded_image = self.image_projection(self.image_encoder(image))
        embedded_att = torch.ones(embedded_image.size()[:-1], dtype=torch.long)
        prompt = torch.tensor([list(prompt_ids)] * image.size(0), dtype=torch.long)
        attention_mask = torch.cat([embedded_att, torch.ones_like(prompt)], dim=1)
        inputs_embeds = self.llm.model.embed_tokens(prompt)                         # starvector_v2.py:45-47
        return torch.cat([embedded_image, inputs_embeds], dim=1), attention_mask, prompt

    def generation_kwargs(self, base, stop_ids):
        kw = super().generation_kwargs(base, stop_ids)
        kw.pop("pad_token_id")                                                     # _get_im2svg_specific_kwargs -> {} (v2:53-57)
        kw.pop("early_stopping")                                                   # v2 passes none: HF default (False) applies
        return kw

    def _embed(self, ids):
        return self.llm.model.embed_tokens(ids)

    @property
    def _body(self):
        return self.llm.model

    @torch.no_grad()
    def teacher_forced_logits(self, image, prompt_ids, forced):
        inputs_embeds, _, _ = self.prepare_generation_inputs(image, prompt_ids)
        t0 = inputs_embeds.shape[1]
        emb = torch.cat([inputs_embeds, self.llm.model.embed_tokens(forced)], dim=1)
        out = self.llm(inputs_embeds=emb, use_cache=False)
        return out.logits[:, t0 - 1:, :].float()
