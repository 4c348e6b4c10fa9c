"""CPU check of the DEVICE beam-search bookkeeping (csrc/sv_beam_core.h, the code beam_step_kernel runs) against HF
generate(num_beams > 1): the library's host replays `sv_beam_row_candidates_host` / `sv_beam_step_host` drive a whole
search over the CPU oracle's logits (the oracle's HF decoder is the forward pass, with its own KV cache permuted by the
parent rows the step returns), and the returned hypotheses must equal HF's exactly -- the same contract
tests/test_beam_logic.py holds the torch restatement (starvector_b200/beam_search.py) to.  No GPU: the library only has
to load."""
import ctypes as C

import pytest
import torch

from oracle.pipeline import OracleStarVector
from starvector_b200 import _lib
from starvector_b200.beam_search import beam_search
from starvector_b200.config import dims_tiny
from starvector_b200.weights import synthetic_images, synthetic_state_dict
from test_beam_logic import OracleBackedEngine

PROMPT = [44, 78]


def _fp(t):
    return t.ctypes.data_as(C.POINTER(C.c_float)) if hasattr(t, "ctypes") else C.cast(t.data_ptr(), C.POINTER(C.c_float))


def _ip(t):
    return C.cast(t.data_ptr(), C.POINTER(C.c_int32))


def core_beam_search(eng, image, prompt_ids, *, num_beams, max_new_tokens, do_sample=False, temperature=1.0, top_p=1.0,
                     repetition_penalty=1.0, length_penalty=1.0, early_stopping=True, eos_token_id=0, pad_token_id=0,
                     stop_ids=(), seed=0, trace=None):
    """The loop sv_beam_search runs on the device, on the host: same stages, same state blob."""
    lib = _lib.load()
    B, nb = image.shape[0], num_beams
    R, K, V = B * nb, 2 * nb, eng.dims.vocab
    bp = _lib.BeamParams()
    bp.num_beams, bp.max_new_tokens, bp.do_sample = nb, max_new_tokens, int(do_sample)
    bp.early_stopping = 2 if early_stopping == "never" else int(bool(early_stopping))
    bp.temperature, bp.top_p, bp.repetition_penalty, bp.length_penalty = temperature, top_p, repetition_penalty, length_penalty
    bp.eos_token_id = -1 if eos_token_id is None else eos_token_id
    bp.pad_token_id = pad_token_id
    bp.n_stop_ids = len(stop_ids)
    for i, s in enumerate(stop_ids):
        bp.stop_ids[i] = int(s)
    bp.seed = seed
    assert lib.sv_beam_params_check(C.byref(bp), B) == 0
    state = C.create_string_buffer(lib.sv_beam_state_bytes())
    prefix_len = eng.dims.query_length + prompt_ids.shape[1]
    assert lib.sv_beam_state_init_host(C.byref(bp), B, prefix_len, state) == 0
    stride = max_new_tokens
    run_seq = torch.full((2, R, stride), pad_token_id, dtype=torch.int32)
    fin_seq = torch.full((2, R, stride), pad_token_id, dtype=torch.int32)
    eng.encode_images(image.repeat_interleave(nb, dim=0))
    logits = eng.prefill(prompt_ids.repeat_interleave(nb, dim=0), return_logits=True)
    key = torch.empty(R, K, dtype=torch.float32)
    val = torch.empty(R, K, dtype=torch.float32)
    tok = torch.empty(R, K, dtype=torch.int32)
    nxt = torch.empty(R, dtype=torch.int32)
    src = torch.empty(R, dtype=torch.int32)
    plan = torch.zeros(64, dtype=torch.int32)
    parity, cur = C.c_int32(), C.c_int32()
    fin_len = (C.c_int32 * 8)()
    cache_hi = prefix_len - 1
    step = 0
    while True:
        lg = logits.float().contiguous()
        lib.sv_beam_state_read_host(state, C.byref(parity), C.byref(cur), fin_len, None)
        scores = torch.frombuffer(state, dtype=torch.float32, count=12)[4:12].clone()    # running_scores inside the blob
        for r in range(R):
            seq = run_seq[parity.value, r, : cur.value].contiguous()
            rc = lib.sv_beam_row_candidates_host(C.byref(bp), _fp(lg[r]), V, _ip(seq), cur.value, float(scores[r]), step, r,
                                                 _fp(key[r]), _fp(val[r]), _ip(tok[r]))
            assert rc == 0
        cont = lib.sv_beam_step_host(C.byref(bp), B, V, stride, state, _fp(key), _fp(val), _ip(tok), _ip(run_seq), _ip(fin_seq),
                                     cache_hi, _ip(nxt), _ip(src), _ip(plan))
        assert cont in (0, 1)
        if trace is not None:
            trace.append((nxt.clone(), src.clone(), plan.clone()))
        step += 1
        cache_hi += 1
        if not cont:
            break
        eng.reorder_cache(src)
        logits = eng.decode_step(nxt)
    lib.sv_beam_state_read_host(state, C.byref(parity), C.byref(cur), fin_len, None)
    n_gen = max(fin_len[b * nb] for b in range(B))
    return fin_seq[parity.value, 0::nb, :n_gen].long()


@pytest.fixture(scope="module")
def setup():
    torch.set_num_threads(1)
    d = dims_tiny(max_batch=8)
    sd = dict(synthetic_state_dict(d, seed=0, init="randomized"))
    g = torch.Generator().manual_seed(3)
    sd["model.svg_transformer.transformer.lm_head.weight"] = (torch.randn(d.vocab, d.hidden, generator=g) * 0.2).to(torch.bfloat16)
    o = OracleStarVector(d, sd, dtype=torch.float32, pad_token_id=d.vocab - 4)
    o.llm.lm_head.weight = torch.nn.Parameter(sd["model.svg_transformer.transformer.lm_head.weight"].float())
    return d, o, synthetic_images(d, 2, seed=1).float()


def test_state_blob_layout():
    """The test reads the running scores out of the blob: pin the offsets it relies on (cur_len, done, parity, pad, then
    running_scores[8])."""
    lib = _lib.load()
    bp = _lib.BeamParams()
    bp.num_beams, bp.max_new_tokens, bp.repetition_penalty, bp.temperature = 2, 4, 1.0, 1.0
    state = C.create_string_buffer(lib.sv_beam_state_bytes())
    assert lib.sv_beam_state_init_host(C.byref(bp), 2, 7, state) == 0
    ints = torch.frombuffer(state, dtype=torch.int32, count=4)
    assert ints.tolist()[:3] == [0, 0, 0]
    fl = torch.frombuffer(state, dtype=torch.float32, count=12)[4:]
    assert fl.tolist() == [0.0, -1e9, 0.0, -1e9, 0.0, -1e9, 0.0, -1e9]
    assert lib.sv_beam_params_check(C.byref(bp), 5) != 0          # 5 x 2 rows > 8


@pytest.mark.parametrize("nb,lp,rp,stop", [(2, 1.0, 1.0, ()), (3, 1.0, 1.0, ()), (2, -1.0, 3.1, ()), (2, 1.0, 1.0, "row0"),
                                           (2, 2.0, 1.0, ()), (4, 1.0, 1.3, ())])
def test_device_bookkeeping_matches_hf(setup, nb, lp, rp, stop):
    d, o, img = setup
    n_new = 14
    kw = dict(use_nucleus_sampling=False, num_beams=nb, length_penalty=lp, repetition_penalty=rp,
              max_length=d.query_length + len(PROMPT) + n_new)
    stop_ids = ()
    if stop == "row0":
        base = o.generate_im2svg_ids(img, PROMPT, (), **kw)
        stop_ids = tuple(base[0, 2 + 4: 2 + 6].tolist())
    ref = o.generate_im2svg_ids(img, PROMPT, stop_ids, **kw)[:, len(PROMPT):]
    got = core_beam_search(OracleBackedEngine(o), img, torch.tensor([PROMPT] * 2), num_beams=nb, max_new_tokens=n_new,
                           repetition_penalty=rp, length_penalty=lp, early_stopping=True, eos_token_id=0,
                           pad_token_id=d.vocab - 4, stop_ids=stop_ids)
    assert got.shape == ref.shape and torch.equal(got, ref), (got.tolist(), ref.tolist())


@pytest.mark.parametrize("nb,lp,es", [(2, 1.0, False), (3, 1.0, False), (2, -1.0, False), (2, 1.0, "never")])
def test_device_bookkeeping_early_stopping_modes(setup, nb, lp, es):
    """early_stopping False (what v2 gets, starvector_v2.py:53-57) and "never": against the torch restatement, which
    tests/test_beam_logic.py pins to HF for False."""
    d, o, img = setup
    n_new = 14
    kw = dict(num_beams=nb, max_new_tokens=n_new, length_penalty=lp, early_stopping=es, eos_token_id=0, pad_token_id=d.vocab - 4)
    ref = beam_search(OracleBackedEngine(o), img, torch.tensor([PROMPT] * 2), **kw)
    got = core_beam_search(OracleBackedEngine(o), img, torch.tensor([PROMPT] * 2), **kw)
    assert got.shape == ref.shape and torch.equal(got, ref), (got.tolist(), ref.tolist())


def test_kv_plan_reproduces_the_parent_rows(setup):
    """The KV suffix-copy plan (copy_src / copy_lo / copy_hi + the divergence matrix) must leave every cache row equal to
    its parent's history: replay the plan on a symbolic cache (row -> list of (position, writer-token) entries) and
    compare with the full-row permutation HF's `_reorder_cache` makes."""
    d, o, img = setup
    nb, n_new = 3, 12
    trace = []
    core_beam_search(OracleBackedEngine(o), img, torch.tensor([PROMPT] * 2), num_beams=nb, max_new_tokens=n_new,
                     early_stopping=False, eos_token_id=None, pad_token_id=d.vocab - 4, trace=trace)
    R = 2 * nb
    prefix = d.query_length + len(PROMPT)
    full = [[("p", i) for i in range(prefix)] for _ in range(R)]       # what a whole-row gather keeps
    part = [list(row) for row in full]                                 # what the suffix copies keep
    for step, (nxt, src, plan) in enumerate(trace):
        p = plan.tolist()
        copy_src, copy_lo, copy_hi, cont = p[40:48], p[48:56], p[56], p[57]
        if not cont:
            break
        new_full = [list(full[int(src[r])]) for r in range(R)]
        staged = {r: part[copy_src[r]][copy_lo[r]: copy_hi + 1] for r in range(R) if copy_src[r] >= 0 and copy_lo[r] <= copy_hi}
        for r, seg in staged.items():
            part[r][copy_lo[r]: copy_hi + 1] = seg
        full = new_full
        assert part == full, f"step {step}: suffix copies diverge from the whole-row permutation"
        for r in range(R):                                             # the forward pass appends this step's token
            full[r].append((step, int(nxt[r]), r))
            part[r].append((step, int(nxt[r]), r))


def test_beam_sample_candidates_follow_the_distribution():
    """Beam-sample: the first candidate of a row is a draw from softmax(processed log-probs) (Gumbel top-K = sampling
    without replacement); check the empirical distribution of the first draw and that no filtered token is ever drawn."""
    lib = _lib.load()
    V, nb = 24, 2
    bp = _lib.BeamParams()
    bp.num_beams, bp.max_new_tokens, bp.do_sample = nb, 8, 1
    bp.temperature, bp.top_p, bp.repetition_penalty, bp.length_penalty = 0.8, 0.9, 1.0, 1.0
    bp.eos_token_id, bp.pad_token_id = 0, 0
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(V, generator=g) * 2.0).to(torch.bfloat16).float().contiguous()
    lp = torch.log_softmax(logits, -1) / 0.8
    sl, si = torch.sort(lp, descending=False)
    remove = sl.softmax(-1).cumsum(-1) <= (1 - 0.9)
    remove[-2:] = False
    keep = torch.ones(V, dtype=torch.bool)
    keep[si[remove]] = False
    expect = torch.where(keep, lp, torch.tensor(float("-inf"))).softmax(-1)
    counts = torch.zeros(V)
    key = torch.empty(4, dtype=torch.float32); val = torch.empty(4, dtype=torch.float32); tok = torch.empty(4, dtype=torch.int32)
    empty = torch.zeros(1, dtype=torch.int32)
    n = 4000
    for s in range(n):
        bp.seed = 1000 + s
        assert lib.sv_beam_row_candidates_host(C.byref(bp), _fp(logits), V, _ip(empty), 0, 0.0, 3, 1, _fp(key), _fp(val), _ip(tok)) == 0
        finite = [int(tok[k]) for k in range(4) if val[k] > float("-inf")]
        assert len(set(finite)) == len(finite) >= 2               # draws without replacement; min_tokens_to_keep = 2
        for k in range(4):
            if val[k] > float("-inf"):
                assert bool(keep[tok[k]]) and abs(float(val[k]) - float(lp[tok[k]])) < 1e-5
        counts[tok[0]] += 1
    tv = 0.5 * (counts / n - expect).abs().sum().item()
    assert tv < 0.04, tv


class _QuantizedLogits(OracleBackedEngine):
    """The oracle's logits rounded to multiples of 0.25: candidates tie at almost every step, as bf16 logits do at 49k vocab."""

    def prefill(self, prompt_ids, return_logits=False):
        return torch.round(super().prefill(prompt_ids, return_logits) * 4) / 4

    def decode_step(self, ids):
        return torch.round(super().decode_step(ids) * 4) / 4


@pytest.mark.parametrize("nb,rp", [(2, 1.0), (3, 1.0), (4, 2.0)])
def test_tie_rule_is_the_same_in_both_loops(setup, nb, rp):
    """Equal scores are ordered by lower beam, then lower token id, in the device bookkeeping (sv_beam_core.h, argmax rounds with
    removal + merge) and in the host-stepped loop (stable sort): with heavily tied logits the two must still return the same
    hypotheses.  (HF itself leaves the order of ties to torch.topk, so this is pinned between our two implementations.)"""
    d, o, img = setup
    n_new = 12
    kw = dict(num_beams=nb, max_new_tokens=n_new, repetition_penalty=rp, early_stopping=True, eos_token_id=0, pad_token_id=d.vocab - 4)
    ids = torch.tensor([PROMPT] * 2)
    ref = beam_search(_QuantizedLogits(o), img, ids, **kw)
    # the quantised model really ties: count exact ties among the top candidates of the first step
    q = _QuantizedLogits(o)
    q.encode_images(img)
    first = torch.log_softmax(q.prefill(ids, return_logits=True).float(), -1)
    top = first.topk(8).values
    assert int((top[:, 1:] == top[:, :-1]).sum()) >= 2, "no ties among the top-8 of the first step: the test lost its power"
    got = core_beam_search(_QuantizedLogits(o), img, ids, **kw)
    assert got.shape == ref.shape and torch.equal(got, ref), (got.tolist(), ref.tolist())


class _SeededLogitsEngine:
    """No model: the logits of a row are seeded noise of its token history (bf16-valued, full 1B vocabulary), so both loops see
    identical logits for identical histories -- long searches at the real vocabulary size without a GPU."""

    class _Dims:
        pass

    def __init__(self, vocab, num_beams, scale=2.0):
        import hashlib

        self._sha = hashlib.sha256
        self.dims = self._Dims()
        self.dims.vocab, self.dims.max_batch, self.dims.query_length = vocab, 8, 5
        self.nb, self.scale, self.hist = num_beams, scale, None

    def _logits(self):
        rows = []
        for h in self.hist:
            seed = int.from_bytes(self._sha(repr(h).encode()).digest()[:7], "little")
            g = torch.Generator().manual_seed(seed)
            rows.append((torch.randn(self.dims.vocab, generator=g) * self.scale).to(torch.bfloat16).float())
        return torch.stack(rows)

    def encode_images(self, image):
        pass

    def prefill(self, prompt_ids, return_logits=False):
        self.hist = [(("image", r // self.nb),) for r in range(prompt_ids.shape[0])]
        return self._logits()

    def decode_step(self, ids):
        self.hist = [h + (int(t),) for h, t in zip(self.hist, ids.tolist())]
        return self._logits()

    def reorder_cache(self, idx):
        self.hist = [self.hist[int(i)] for i in idx.tolist()]


def test_long_search_at_the_1b_vocabulary_matches_the_host_loop():
    """256 steps, 49,156-entry vocabulary, repetition_penalty 3.1, length_penalty -1, early_stopping "never" (the beam bench's
    settings): running scores reach -10^3, every step reorders two beams that share a long prefix, the repetition-penalty set
    grows to hundreds of tokens.  The host replays of the device stages must return the torch loop's hypothesis."""
    V, n_new = 49156, 256
    kw = dict(num_beams=2, max_new_tokens=n_new, repetition_penalty=3.1, length_penalty=-1.0, early_stopping="never",
              eos_token_id=None, pad_token_id=49152)
    img, ids = torch.zeros(1, 3, 4, 4), torch.tensor([[1, 2]])
    ref = beam_search(_SeededLogitsEngine(V, 2), img, ids, **kw)
    got = core_beam_search(_SeededLogitsEngine(V, 2), img, ids, **kw)
    assert ref.shape == (1, n_new) and torch.equal(got, ref)
