"""Restatement of the reference's CLIP ViT tower and Adapter on CPU (torch, any dtype).

Each function cites the reference lines it follows.  Weights come in as a flat dict with
the reference's own parameter names (relative to the module), so the same tensors feed the
reference modules, this oracle and the CUDA engine.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _ln(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """clip_model.py:117-124 — LayerNorm run in the weight dtype, cast back to x's dtype."""
    orig = x.dtype
    return F.layer_norm(x.type(w.dtype), (x.shape[-1],), w, b, eps).type(orig)


def _mha(x: torch.Tensor, p: Dict[str, torch.Tensor], pre: str, heads: int) -> torch.Tensor:
    """nn.MultiheadAttention(x,x,x, need_weights=False) with seq-first input `[L,B,E]`.

    clip_model.py:134,148-150.  Restates torch's packed in-projection → per-head SDPA →
    out-projection sequence (torch/nn/functional.py `multi_head_attention_forward`) so the
    result is bit-identical to the module on CPU.
    """
    L, B, E = x.shape
    hd = E // heads
    proj = F.linear(x, p[pre + "attn.in_proj_weight"], p[pre + "attn.in_proj_bias"])
    proj = proj.unflatten(-1, (3, E)).unsqueeze(0).transpose(0, -2).squeeze(-2).contiguous()
    q, k, v = proj[0], proj[1], proj[2]
    q = q.view(L, B * heads, hd).transpose(0, 1).view(B, heads, L, hd)
    k = k.view(L, B * heads, hd).transpose(0, 1).view(B, heads, L, hd)
    v = v.view(L, B * heads, hd).transpose(0, 1).view(B, heads, L, hd)
    o = F.scaled_dot_product_attention(q, k, v, None, 0.0, False)
    o = o.permute(2, 0, 1, 3).contiguous().view(B * L, E)
    o = F.linear(o, p[pre + "attn.out_proj.weight"], p[pre + "attn.out_proj.bias"])
    return o.view(L, B, E)


def clip_vit_forward(img: torch.Tensor, p: Dict[str, torch.Tensor], *, patch: int, heads: int,
                     layers: int) -> torch.Tensor:
    """VisionTransformer.forward (clip_model.py:181-191): `[B,3,S,S]` → `[B,Q,W]` (no ln_post/proj)."""
    x = F.conv2d(img, p["conv1.weight"], None, stride=patch)                    # :182
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)                  # :183-184
    cls = p["class_embedding"].to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype)
    x = torch.cat([cls, x], dim=1)                                              # :185
    x = x + p["positional_embedding"].to(x.dtype)                               # :186
    x = _ln(x, p["ln_pre.weight"], p["ln_pre.bias"])                            # :187
    x = x.permute(1, 0, 2)                                                      # :188  NLD -> LND
    for i in range(layers):                                                     # :189 (130-155)
        pre = f"transformer.resblocks.{i}."
        x = x + _mha(_ln(x, p[pre + "ln_1.weight"], p[pre + "ln_1.bias"]), p, pre, heads)   # :153
        h = _ln(x, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"])
        h = F.linear(h, p[pre + "mlp.c_fc.weight"], p[pre + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)                                        # QuickGELU :126-128
        h = F.linear(h, p[pre + "mlp.c_proj.weight"], p[pre + "mlp.c_proj.bias"])
        x = x + h                                                               # :154
    return x.permute(1, 0, 2)                                                   # :190


def image_encoder_forward(img: torch.Tensor, vis: Dict[str, torch.Tensor], lnv: Dict[str, torch.Tensor],
                          *, patch: int, heads: int, layers: int) -> torch.Tensor:
    """ImageEncoder.forward, clip branch (image_encoder.py:91-94): ViT then `ln_vision`."""
    e = clip_vit_forward(img, vis, patch=patch, heads=heads, layers=layers)
    return _ln(e, lnv["weight"], lnv["bias"])


def adapter_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], adapter_norm: str) -> torch.Tensor:
    """Adapter.forward in eval mode (adapters/adapter.py:33-39; dropout is the identity)."""
    h = F.linear(x, p["c_fc.weight"], p["c_fc.bias"])                           # :34
    h = h * torch.sigmoid(h)                                                    # Swish :9-10
    z = F.linear(h, p["c_proj.weight"], p["c_proj.bias"])                       # :36
    if adapter_norm == "layer_norm":                                            # :25-26 LayerNorm([Q,H])
        return F.layer_norm(z, tuple(p["norm.weight"].shape), p["norm.weight"], p["norm.bias"], 1e-5)
    if adapter_norm == "batch_norm":                                            # :27-28 BatchNorm1d(Q), eval
        return F.batch_norm(z, p["norm.running_mean"], p["norm.running_var"], p["norm.weight"],
                            p["norm.bias"], False, 0.1, 1e-5)
    raise ValueError(adapter_norm)
