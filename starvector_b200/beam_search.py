"""Beam search / beam-sample over the engine (the reference's default is `num_beams=2`,
`do_sample=True`, `early_stopping=True` — starvector_base.py:231-234,292-295).

Two implementations with one contract.  `impl="device"` (the default on a real engine): `sv_beam_search` -- candidate
selection, the bookkeeping below (restated once more in csrc/sv_beam_core.h and pinned to HF by tests/test_beam_core.py) and
the cache permutation all run inside the replayed decode graph; the host only polls a done flag.  `impl="host"`
(`SV_BEAM=host`, or an engine without `beam_search_device`): the loop below, one `sv_decode_step` + `sv_reorder_cache` per
step -- the torch restatement the device code is tested against, and the fallback for vocabularies that do not fit the
candidate kernel's shared memory.

The model forward of every step runs in the CUDA engine (beams are image rows: `sv_decode_step`, and the KV cache
is permuted with `sv_reorder_cache`); the per-step bookkeeping below is a restatement of transformers'
`GenerationMixin._beam_search` (generation/utils.py:2844-3425 in the installed 5.5.0: `_get_top_k_continuations`,
`_get_running_beams_for_next_iteration`, `_update_finished_beams`, `_check_early_stop_heuristic`,
`_beam_search_has_unfinished_sequences`) for the decoder-only / `inputs_embeds` case the reference uses, on a handful of
tiny `[B, num_beams]` tensors with torch ops — host logic, like HF's own.

With `do_sample=False` the result is deterministic and is tested for equality with HF on the oracle; with
`do_sample=True` (beam-sample) candidates are drawn with `torch.multinomial`, so parity is distribution-level only.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch

from .engine import Engine


def _gather_beams(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """[B, nb, ...] gathered along dim 1 with idx [B, k]  (utils.py:2856-2873)."""
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.gather(t, 1, idx.expand(-1, -1, *t.shape[2:]))


def _topk_stable(x: torch.Tensor, k: int):
    """Top-k along the last dim with a DEFINED tie rule: equal values keep index order (lower flat index = lower beam, then
    lower token id, first) -- the rule the device kernels follow (csrc/sv_beam_core.h).  bf16 logits tie often, and
    torch.topk leaves the order of ties to the backend."""
    v, i = torch.sort(x, dim=-1, descending=True, stable=True)
    return v[..., :k], i[..., :k]


def _process_log_probs(log_probs: torch.Tensor, generated: torch.Tensor, repetition_penalty: float, do_sample: bool,
                       temperature: float, top_p: float, min_keep: int) -> torch.Tensor:
    """logits_processor(flat_running_sequences, log_probs): repetition penalty -> temperature -> top-p (App. B.3)."""
    if repetition_penalty != 1.0 and generated.shape[1] > 0:
        score = torch.gather(log_probs, 1, generated)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        log_probs = log_probs.scatter(1, generated, score)
    if do_sample:
        if temperature != 1.0:
            log_probs = log_probs / temperature
        if top_p < 1.0:
            sorted_logits, sorted_idx = torch.sort(log_probs, descending=False)
            cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
            remove = cum <= (1 - top_p)
            remove[..., -min_keep:] = False
            log_probs = log_probs.masked_fill(remove.scatter(1, sorted_idx, remove), float("-inf"))
    return log_probs


@torch.no_grad()
def beam_search(engine: Engine, image: Optional[torch.Tensor], prompt_ids: Optional[torch.Tensor], *, num_beams: int,
                max_new_tokens: int, inputs_embeds: Optional[torch.Tensor] = None, do_sample: bool = False, temperature: float = 1.0, top_p: float = 1.0, repetition_penalty: float = 1.0,
                length_penalty: float = 1.0, early_stopping=True, eos_token_id: Optional[int] = 0, pad_token_id: int = 0,
                stop_ids: Sequence[int] = (), seed: int = 0, impl: Optional[str] = None) -> torch.Tensor:
    """Returns int64 `[B, n_generated]`: the best finished (or running) beam per image, new tokens only."""
    src = inputs_embeds if inputs_embeds is not None else image
    B, nb = src.shape[0], int(num_beams)
    if B * nb > engine.dims.max_batch:
        raise ValueError(f"batch {B} x num_beams {nb} exceeds the engine's max_batch {engine.dims.max_batch}")
    if impl is None:
        impl = os.environ.get("SV_BEAM", "device")
    if impl not in ("device", "host"):
        raise ValueError("impl must be 'device' or 'host'")
    on_device = impl == "device" and hasattr(engine, "beam_search_device")
    # _expand_inputs_for_generation: every image row becomes num_beams adjacent rows
    if inputs_embeds is not None:
        logits = engine.prefill_embeds(inputs_embeds.repeat_interleave(nb, dim=0), return_logits=not on_device)
    else:
        engine.encode_images(image.repeat_interleave(nb, dim=0))
        logits = engine.prefill(prompt_ids.repeat_interleave(nb, dim=0), return_logits=not on_device)
    if on_device:
        fill_dev = (pad_token_id if pad_token_id is not None else eos_token_id) if eos_token_id is not None else -1
        try:
            return engine.beam_search_device(
                B, num_beams=nb, max_new_tokens=max_new_tokens, do_sample=do_sample, temperature=temperature, top_p=top_p,
                repetition_penalty=repetition_penalty, length_penalty=length_penalty, early_stopping=early_stopping,
                eos_token_id=eos_token_id, pad_token_id=fill_dev, stop_ids=stop_ids, seed=seed).long()
        except NotImplementedError:          # vocabulary too large for the candidate kernel: the host-stepped loop below
            if inputs_embeds is not None:
                logits = engine.prefill_embeds(inputs_embeds.repeat_interleave(nb, dim=0), return_logits=True)
            else:
                logits = engine.prefill(prompt_ids.repeat_interleave(nb, dim=0), return_logits=True)
    dev, V = logits.device, engine.dims.vocab
    max_length, cur_len, prompt_len = int(max_new_tokens), 0, 0          # generated-token coordinates (inputs_embeds)
    n_eos = 0 if eos_token_id is None else 1
    K = max(2, 1 + n_eos) * nb                                            # beams_to_keep
    top_mask = torch.zeros(K, dtype=torch.bool, device=dev)
    top_mask[:nb] = True
    fill = (pad_token_id if pad_token_id is not None else eos_token_id) if eos_token_id is not None else -1   # HF: `pad if pad is not None`
    running_sequences = torch.full((B, nb, max_length), fill, dtype=torch.int64, device=dev)
    sequences = running_sequences.clone()
    running_beam_scores = torch.zeros((B, nb), dtype=torch.float32, device=dev)
    running_beam_scores[:, 1:] = -1e9
    beam_scores = torch.full((B, nb), -1e9, dtype=torch.float32, device=dev)
    is_sent_finished = torch.zeros((B, nb), dtype=torch.bool, device=dev)
    unsatisfied = torch.ones((B, 1), dtype=torch.bool, device=dev)       # is_early_stop_heuristic_unsatisfied
    running_beam_indices = torch.full((B, nb, max_length), -1, dtype=torch.int32, device=dev)
    beam_indices = running_beam_indices.clone()
    gen = torch.Generator(device=dev).manual_seed(seed)
    stop = list(stop_ids)
    batch_offset = (torch.arange(B, device=dev) * nb).view(-1, 1)

    while True:
        log_probs = torch.log_softmax(logits.float(), dim=-1)
        log_probs = _process_log_probs(log_probs, running_sequences[:, :, :cur_len].reshape(B * nb, cur_len),
                                       repetition_penalty, do_sample, temperature, top_p, min_keep=max(2, 1 + n_eos))
        log_probs = (log_probs.view(B, nb, V) + running_beam_scores[:, :, None]).reshape(B, nb * V)
        # ---- _get_top_k_continuations
        if do_sample:
            topk_indices = torch.multinomial(torch.softmax(log_probs, dim=-1), num_samples=K, generator=gen)
            topk_log_probs = torch.gather(log_probs, 1, topk_indices)
        else:
            topk_log_probs, topk_indices = _topk_stable(log_probs, K)
        topk_beam = topk_indices // V
        topk_running_beam_indices = _gather_beams(running_beam_indices, topk_beam)
        topk_running_sequences = _gather_beams(running_sequences, topk_beam)
        topk_running_sequences[:, :, cur_len] = topk_indices % V
        topk_running_beam_indices[:, :, cur_len - prompt_len] = (topk_beam + batch_offset).to(torch.int32)
        # ---- stopping criteria on the flattened candidates: MaxLength | EOS | StoppingCriteriaSub (row 0 -> everyone)
        flat = topk_running_sequences[:, :, : cur_len + 1].reshape(B * K, cur_len + 1)
        hits = torch.zeros(B * K, dtype=torch.bool, device=dev)
        if cur_len + 1 >= max_length:
            hits |= True
        if eos_token_id is not None:
            hits |= flat[:, -1] == eos_token_id
        if stop and cur_len + 1 >= len(stop) and flat[0, -len(stop):].tolist() == stop:    # starvector_base.py:15-20
            hits |= True
        hits = hits.view(B, K)
        # ---- _get_running_beams_for_next_iteration
        topk_running_log_probs = topk_log_probs + hits.to(torch.float32) * -1.0e9
        next_idx = _topk_stable(topk_running_log_probs, nb)[1]
        running_sequences = _gather_beams(topk_running_sequences, next_idx)
        running_beam_scores = _gather_beams(topk_running_log_probs, next_idx)
        running_beam_indices = _gather_beams(topk_running_beam_indices, next_idx)
        # ---- _update_finished_beams
        just_finished = hits & top_mask[None, :]
        fin_log_probs = topk_log_probs / ((cur_len + 1 - prompt_len) ** length_penalty)
        full = torch.all(is_sent_finished, dim=-1, keepdim=True) & (early_stopping is True)
        fin_log_probs = fin_log_probs + full.to(torch.float32) * -1.0e9
        fin_log_probs = fin_log_probs + (~unsatisfied).to(torch.float32) * -1.0e9
        fin_log_probs = fin_log_probs + (~just_finished) * -1.0e9
        merged_scores = torch.cat((beam_scores, fin_log_probs), dim=1)
        top_merged = _topk_stable(merged_scores, nb)[1]
        sequences = _gather_beams(torch.cat((sequences, topk_running_sequences), dim=1), top_merged)
        beam_scores = _gather_beams(merged_scores, top_merged)
        beam_indices = _gather_beams(torch.cat((beam_indices, topk_running_beam_indices), dim=1), top_merged)
        is_sent_finished = _gather_beams(torch.cat((is_sent_finished, just_finished), dim=1), top_merged)
        # ---- cache permutation for the next forward, loop bookkeeping
        beam_idx = running_beam_indices[..., cur_len - prompt_len].reshape(-1)
        cur_len += 1
        # _check_early_stop_heuristic
        if early_stopping == "never" and length_penalty > 0.0:
            best_len = max_length - prompt_len
        else:
            best_len = cur_len - prompt_len
        best_running = running_beam_scores[:, :1] / (best_len ** length_penalty)
        worst_finished = torch.where(is_sent_finished, torch.min(beam_scores, dim=1, keepdim=True)[0],
                                     torch.tensor(-1.0e9, device=dev))
        unsatisfied = unsatisfied & torch.any(best_running > worst_finished, dim=-1, keepdim=True)
        # _beam_search_has_unfinished_sequences
        improvement_possible = bool(torch.any(unsatisfied))
        exists_open_beam = not (bool(torch.all(is_sent_finished)) and early_stopping is True)
        valid_continuations = not bool(torch.all(hits))
        if not (improvement_possible and exists_open_beam and valid_continuations):
            break
        engine.reorder_cache(beam_idx)
        logits = engine.decode_step(running_sequences[:, :, cur_len - 1].reshape(-1))

    best = sequences[:, 0, :]                                           # num_return_sequences = 1
    max_generated = int(((beam_indices[:, 0, :] + 1).bool()).sum(dim=1).max())
    return best[:, :max_generated]
