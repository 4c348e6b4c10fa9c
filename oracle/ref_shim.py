"""Import the reference's own in-tree modules (authoring container only).

`/root/reference` does not exist on the GPU box; callers must check `available()` first.
Only the fairscale import (used for grad-checkpointing, clip_model.py:10) needs a shim
(SURVEY.md §8c).  Nothing here copies reference code: it imports it in place.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "starvector"))


def load():
    """Returns (VisionTransformer, LayerNorm, Adapter) classes of the reference."""
    if not available():
        raise RuntimeError("/root/reference is not mounted")
    for n in ("fairscale", "fairscale.nn", "fairscale.nn.checkpoint",
              "fairscale.nn.checkpoint.checkpoint_activations"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["fairscale.nn.checkpoint.checkpoint_activations"].checkpoint_wrapper = lambda m, **k: m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from starvector.model.image_encoder.clip_model import VisionTransformer, LayerNorm
    from starvector.model.adapters.adapter import Adapter
    return VisionTransformer, LayerNorm, Adapter
