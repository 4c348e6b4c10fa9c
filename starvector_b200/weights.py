"""Synthetic (seeded) checkpoints under the reference's state-dict names.

There are no hub weights offline, so the oracle and the engine share weights generated
here.  Names follow the reference module tree (reference:
starvector/model/models/starvector_base.py:29-36, image_encoder/clip_model.py:167-179,
adapters/adapter.py:19-28, and transformers' GPTBigCodeForCausalLM), so a real
``from_pretrained`` state dict loads through the same path.

``init="hf_default"`` reproduces the *distributions* of the reference's default
initialisers (normal(0, .02) decoder, zero biases, unit LayerNorms, xavier adapter);
``init="randomized"`` additionally perturbs every bias / LayerNorm affine so parity tests
exercise those terms.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Tuple

import torch

from .config import ModelDims

VIS = "model.image_encoder.visual_encoder."
LNV = "model.image_encoder.ln_vision."
ADP = "model.image_projection."
DEC = "model.svg_transformer.transformer.transformer."
LM_HEAD = "model.svg_transformer.transformer.lm_head.weight"


DEC2 = "model.svg_transformer.transformer.model."


def weight_shapes_v2(d: ModelDims) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """StarVector v2 (8B family) tensors: SiglipVisionTransformer keys (image_encoder.py:32-48; the pooling `head` is
    omitted — the reference discards its output, image_encoder.py:109), Adapter, Starcoder2ForCausalLM keys
    (llm/starcoder2.py:19-32)."""
    W, Q, H = d.vit_width, d.query_length, d.hidden
    kv, hq = d.n_kv_head * d.head_dim, d.n_head * d.head_dim
    yield VIS + "embeddings.patch_embedding.weight", (W, 3, d.patch_size, d.patch_size), "conv"
    yield VIS + "embeddings.patch_embedding.bias", (W,), "bias"
    yield VIS + "embeddings.position_embedding.weight", (Q, W), "dec"
    for i in range(d.vit_layers):
        p = f"{VIS}encoder.layers.{i}."
        yield p + "layer_norm1.weight", (W,), "ln_w"
        yield p + "layer_norm1.bias", (W,), "ln_b"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield p + f"self_attn.{n}.weight", (W, W), "linear"
            yield p + f"self_attn.{n}.bias", (W,), "bias"
        yield p + "layer_norm2.weight", (W,), "ln_w"
        yield p + "layer_norm2.bias", (W,), "ln_b"
        yield p + "mlp.fc1.weight", (d.vit_mlp, W), "linear"
        yield p + "mlp.fc1.bias", (d.vit_mlp,), "bias"
        yield p + "mlp.fc2.weight", (W, d.vit_mlp), "linear"
        yield p + "mlp.fc2.bias", (W,), "bias"
    yield VIS + "post_layernorm.weight", (W,), "ln_w"
    yield VIS + "post_layernorm.bias", (W,), "ln_b"
    yield ADP + "c_fc.weight", (2 * W, W), "xavier"
    yield ADP + "c_fc.bias", (2 * W,), "bias"
    yield ADP + "c_proj.weight", (H, 2 * W), "xavier"
    yield ADP + "c_proj.bias", (H,), "bias"
    if d.adapter_norm == 0:
        yield ADP + "norm.weight", (Q, H), "ln_w"
        yield ADP + "norm.bias", (Q, H), "ln_b"
    else:
        yield ADP + "norm.weight", (Q,), "ln_w"
        yield ADP + "norm.bias", (Q,), "ln_b"
        yield ADP + "norm.running_mean", (Q,), "bn_mean"
        yield ADP + "norm.running_var", (Q,), "bn_var"
    yield DEC2 + "embed_tokens.weight", (d.vocab, H), "dec"
    for i in range(d.n_layer):
        p = f"{DEC2}layers.{i}."
        yield p + "input_layernorm.weight", (H,), "ln_w"
        yield p + "input_layernorm.bias", (H,), "ln_b"
        yield p + "self_attn.q_proj.weight", (hq, H), "dec"
        yield p + "self_attn.q_proj.bias", (hq,), "bias"
        yield p + "self_attn.k_proj.weight", (kv, H), "dec"
        yield p + "self_attn.k_proj.bias", (kv,), "bias"
        yield p + "self_attn.v_proj.weight", (kv, H), "dec"
        yield p + "self_attn.v_proj.bias", (kv,), "bias"
        yield p + "self_attn.o_proj.weight", (H, hq), "dec_proj"
        yield p + "self_attn.o_proj.bias", (H,), "bias"
        yield p + "post_attention_layernorm.weight", (H,), "ln_w"
        yield p + "post_attention_layernorm.bias", (H,), "ln_b"
        yield p + "mlp.c_fc.weight", (d.n_inner, H), "dec"
        yield p + "mlp.c_fc.bias", (d.n_inner,), "bias"
        yield p + "mlp.c_proj.weight", (H, d.n_inner), "dec_proj"
        yield p + "mlp.c_proj.bias", (H,), "bias"
    yield DEC2 + "norm.weight", (H,), "ln_w"
    yield DEC2 + "norm.bias", (H,), "ln_b"


def weight_shapes(d: ModelDims) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (name, shape, kind) for every tensor; kind drives the initialiser."""
    if d.variant == 1:
        yield from weight_shapes_v2(d)
        return
    W, Q, H = d.vit_width, d.query_length, d.hidden
    kv = d.n_kv_head * d.head_dim
    yield VIS + "conv1.weight", (W, 3, d.patch_size, d.patch_size), "conv"
    yield VIS + "class_embedding", (W,), "vit_embed"
    yield VIS + "positional_embedding", (Q, W), "vit_embed"
    yield VIS + "ln_pre.weight", (W,), "ln_w"
    yield VIS + "ln_pre.bias", (W,), "ln_b"
    for i in range(d.vit_layers):
        p = f"{VIS}transformer.resblocks.{i}."
        yield p + "ln_1.weight", (W,), "ln_w"
        yield p + "ln_1.bias", (W,), "ln_b"
        yield p + "attn.in_proj_weight", (3 * W, W), "xavier"
        yield p + "attn.in_proj_bias", (3 * W,), "bias"
        yield p + "attn.out_proj.weight", (W, W), "linear"
        yield p + "attn.out_proj.bias", (W,), "bias"
        yield p + "ln_2.weight", (W,), "ln_w"
        yield p + "ln_2.bias", (W,), "ln_b"
        yield p + "mlp.c_fc.weight", (d.vit_mlp, W), "linear"
        yield p + "mlp.c_fc.bias", (d.vit_mlp,), "bias"
        yield p + "mlp.c_proj.weight", (W, d.vit_mlp), "linear"
        yield p + "mlp.c_proj.bias", (W,), "bias"
    yield LNV + "weight", (W,), "ln_w"
    yield LNV + "bias", (W,), "ln_b"
    yield ADP + "c_fc.weight", (2 * W, W), "xavier"
    yield ADP + "c_fc.bias", (2 * W,), "bias"
    yield ADP + "c_proj.weight", (H, 2 * W), "xavier"
    yield ADP + "c_proj.bias", (H,), "bias"
    if d.adapter_norm == 0:
        yield ADP + "norm.weight", (Q, H), "ln_w"
        yield ADP + "norm.bias", (Q, H), "ln_b"
    else:
        yield ADP + "norm.weight", (Q,), "ln_w"
        yield ADP + "norm.bias", (Q,), "ln_b"
        yield ADP + "norm.running_mean", (Q,), "bn_mean"
        yield ADP + "norm.running_var", (Q,), "bn_var"
    yield DEC + "wte.weight", (d.vocab, H), "dec"
    yield DEC + "wpe.weight", (d.n_positions, H), "dec"
    for i in range(d.n_layer):
        p = f"{DEC}h.{i}."
        yield p + "ln_1.weight", (H,), "ln_w"
        yield p + "ln_1.bias", (H,), "ln_b"
        yield p + "attn.c_attn.weight", (H + 2 * kv, H), "dec"
        yield p + "attn.c_attn.bias", (H + 2 * kv,), "bias"
        yield p + "attn.c_proj.weight", (H, H), "dec_proj"
        yield p + "attn.c_proj.bias", (H,), "bias"
        yield p + "ln_2.weight", (H,), "ln_w"
        yield p + "ln_2.bias", (H,), "ln_b"
        yield p + "mlp.c_fc.weight", (d.n_inner, H), "dec"
        yield p + "mlp.c_fc.bias", (d.n_inner,), "bias"
        yield p + "mlp.c_proj.weight", (H, d.n_inner), "dec_proj"
        yield p + "mlp.c_proj.bias", (H,), "bias"
    yield DEC + "ln_f.weight", (H,), "ln_w"
    yield DEC + "ln_f.bias", (H,), "ln_b"


def synthetic_state_dict(
    d: ModelDims, seed: int = 0, init: str = "hf_default", dtype: torch.dtype = torch.bfloat16,
    logit_gain: float = 1.0, device=None,
) -> Dict[str, torch.Tensor]:
    """Seeded random checkpoint (CPU tensors in `dtype`).  `lm_head.weight` is tied to `wte`.

    `device` (a CUDA device) draws the values on the GPU instead: same distributions, other values than the CPU stream,
    for throughput runs of models no CPU oracle is run against (the 15 GB StarVector-8B replica in seconds, not minutes).

    `logit_gain` > 1 scales `wte` (hence the tied lm_head) to widen greedy top-1/top-2
    margins for the "peaked" parity configuration (SURVEY.md §7 hard parts (d)).
    """
    if init not in ("hf_default", "randomized"):
        raise ValueError(f"unknown init {init!r}")
    g = torch.Generator(device=device if device is not None else "cpu").manual_seed(seed)
    rnd = init == "randomized"
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in weight_shapes(d):
        t = torch.empty(shape, dtype=torch.float32, device=device)
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            b = 1.0 / math.sqrt(fan_in)
            t.uniform_(-b, b, generator=g)
        elif kind == "vit_embed":
            t.normal_(0.0, d.vit_width ** -0.5, generator=g)
        elif kind == "linear":
            b = 1.0 / math.sqrt(shape[1])
            t.uniform_(-b, b, generator=g)
        elif kind == "xavier":
            b = math.sqrt(6.0 / (shape[0] + shape[1]))
            t.uniform_(-b, b, generator=g)
        elif kind == "dec":
            t.normal_(0.0, 0.02, generator=g)
            if (name.endswith("wte.weight") or name.endswith("embed_tokens.weight")) and logit_gain != 1.0:
                t.mul_(logit_gain)
        elif kind == "dec_proj":
            t.normal_(0.0, 0.02 / math.sqrt(2 * d.n_layer), generator=g)
        elif kind == "bias":
            t.normal_(0.0, 0.02, generator=g) if rnd else t.zero_()
        elif kind == "ln_w":
            t.normal_(0.0, 0.1, generator=g).add_(1.0) if rnd else t.fill_(1.0)
        elif kind == "ln_b":
            t.normal_(0.0, 0.1, generator=g) if rnd else t.zero_()
        elif kind == "bn_mean":
            t.normal_(0.0, 0.1, generator=g) if rnd else t.zero_()
        elif kind == "bn_var":
            t.uniform_(0.5, 1.5, generator=g) if rnd else t.fill_(1.0)
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[name] = t.to(dtype)
    sd[LM_HEAD] = sd[(DEC2 + "embed_tokens.weight") if d.variant == 1 else (DEC + "wte.weight")]
    return sd


def synthetic_images(d: ModelDims, batch: int, seed: int = 1, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """`[B,3,S,S]` CLIP-normalised noise, the shape `process_images` produces (image_encoder.py:112-117)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, d.image_size, d.image_size, generator=g).to(dtype)
