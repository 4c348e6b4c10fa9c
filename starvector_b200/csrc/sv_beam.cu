// Beam search / beam-sample inside the replayed decode graph (the reference's default generate() mode is
// num_beams=2, starvector_base.py:231-241,289-295).  Three small kernels follow the lm_head of every decode step, so the
// host never sees a logit, a score or a cache permutation:
//   beam_candidates_kernel  one CTA per cache row (= running beam): log-softmax of the bf16 logits row in shared memory,
//                           HF's logits processors on the log-probs (repetition penalty over the beam's own tokens,
//                           temperature, top-p), + the beam's running score, then the row's K best continuations
//                           (beam-sample: the K first draws without replacement, by Gumbel-perturbed top-K);
//   beam_step_kernel        one CTA: merges the rows of each image, runs the bookkeeping of
//                           `GenerationMixin._beam_search` (sv_beam_core.h, shared with the host replay), moves the
//                           token sequences, and writes the next tokens' embeddings for the following decode step;
//   beam_kv_copy_kernel     `_reorder_cache` without moving the cache: a child row only receives the suffix of its
//                           parent's K / V^T rows from the position where the two rows diverged (a per-image divergence
//                           matrix is part of the state) -- typically a handful of tokens, not 2 x the live cache.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/starvector_b200.h"
#include "sv_beam_core.h"
#include "sv_kernels.h"

namespace sv {

using svbeam::Params;
using svbeam::Plan;
using svbeam::State;

constexpr int kBeamThreads = 1024;

namespace {

SV_DEVINL float block_max_f(float v, float* sm) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sm[0];
#pragma unroll
  for (int w = 1; w < kBeamThreads / 32; ++w) r = fmaxf(r, sm[w]);
  return r;
}
SV_DEVINL float block_sum_f(float v, float* sm) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kBeamThreads / 32; ++w) r += sm[w];     // fixed order: deterministic
  return r;
}
// (value desc, index asc) order: is (bv, bi) ahead of (av, ai)?
SV_DEVINL bool ahead(float bv, int bi, float av, int ai) { return bv > av || (bv == av && bi < ai); }

// block-wide first element in (value desc, index asc) order among those strictly BEHIND (lim_v, lim_i)
SV_DEVINL void block_argmax_behind(const float* __restrict__ s, int V, float lim_v, int lim_i, float* smf, int* smi,
                                   float& out_v, int& out_i) {
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += kBeamThreads) {
    const float v = s[i];
    if (ahead(lim_v, lim_i, v, i) && ahead(v, i, bv, bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ahead(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { smf[threadIdx.x >> 5] = bv; smi[threadIdx.x >> 5] = bi; }
  __syncthreads();
  bv = smf[0]; bi = smi[0];
#pragma unroll
  for (int w = 1; w < kBeamThreads / 32; ++w)
    if (ahead(smf[w], smi[w], bv, bi)) { bv = smf[w]; bi = smi[w]; }
  out_v = bv; out_i = bi;
}

}  // namespace

// ---- K1: per-row candidates -------------------------------------------------------------------------------------
// dynamic shared memory: float score[V] | uint32 seen[(V + 31) / 32]
__global__ void __launch_bounds__(kBeamThreads) beam_candidates_kernel(const bf16* __restrict__ logits,
                                                                       const Params* __restrict__ pp, const State* st,
                                                                       const int32_t* __restrict__ run_seq,
                                                                       float* __restrict__ cand_key,
                                                                       float* __restrict__ cand_val,
                                                                       int32_t* __restrict__ cand_tok) {
  if (st->done) return;
  extern __shared__ float sc[];
  __shared__ float smf[kBeamThreads / 32];
  __shared__ int smi[kBeamThreads / 32];
  const int r = blockIdx.x, tid = threadIdx.x;
  const int V = pp->vocab, K = pp->K, R = pp->B * pp->nb, cur = st->cur_len;
  const float rp = pp->rep_penalty, T = pp->temperature, top_p = pp->top_p;
  const bool sample = pp->do_sample != 0;
  uint32_t* seen = reinterpret_cast<uint32_t*>(sc + V);
  const bool use_rp = rp != 1.0f && cur > 0;
  if (use_rp) {                       // tokens this running beam has generated (RepetitionPenaltyLogitsProcessor input)
    for (int i = tid; i < (V + 31) / 32; i += kBeamThreads) seen[i] = 0u;
    __syncthreads();
    const int32_t* seq = run_seq + ((int64_t)st->parity * R + r) * pp->seq_stride;
    for (int i = tid; i < cur; i += kBeamThreads) {
      const int t = seq[i];
      if (t >= 0 && t < V) atomicOr(&seen[t >> 5], 1u << (t & 31));
    }
    __syncthreads();
  }
  const bf16* lr = logits + (int64_t)r * V;
  // log_softmax(logits.float()): x - max - log(sum(exp(x - max)))
  float mx = -INFINITY;
  for (int i = tid; i < V; i += kBeamThreads) { const float x = __bfloat162float(lr[i]); sc[i] = x; mx = fmaxf(mx, x); }
  mx = block_max_f(mx, smf);
  float z = 0.f;
  for (int i = tid; i < V; i += kBeamThreads) z += expf(sc[i] - mx);
  z = block_sum_f(z, smf);
  const float logz = logf(z);
  for (int i = tid; i < V; i += kBeamThreads) {
    const bool sn = use_rp && ((seen[i >> 5] >> (i & 31)) & 1u);
    sc[i] = svbeam::process_logprob((sc[i] - mx) - logz, sn, rp, sample, T);
  }
  __syncthreads();
  if (sample && top_p < 1.0f) {
    // TopPLogitsWarper(min_tokens_to_keep): a token stays iff the mass of strictly more probable tokens is < top_p, or it is
    // one of the min_keep most probable.  Bisection on the probability threshold (as the one-beam sampler does).
    float m2 = -INFINITY;
    for (int i = tid; i < V; i += kBeamThreads) m2 = fmaxf(m2, sc[i]);
    m2 = block_max_f(m2, smf);
    float z2 = 0.f;
    for (int i = tid; i < V; i += kBeamThreads) z2 += expf(sc[i] - m2);
    z2 = block_sum_f(z2, smf);
    const float inv = 1.0f / z2;
    float lo = 0.f, hi = 1.f;
    for (int it = 0; it < 30; ++it) {
      const float mid = 0.5f * (lo + hi);
      float m = 0.f;
      for (int i = tid; i < V; i += kBeamThreads) { const float q = expf(sc[i] - m2) * inv; m += q > mid ? q : 0.f; }
      m = block_sum_f(m, smf);
      if (m < top_p) hi = mid; else lo = mid;
    }
    float kv = INFINITY;                 // the min_keep-th best (value, index): everything not behind it is kept
    int ki = -1;
    for (int j = 0; j < pp->min_keep; ++j) block_argmax_behind(sc, V, kv, ki, smf, smi, kv, ki);
    __syncthreads();
    for (int i = tid; i < V; i += kBeamThreads) {
      const float v = sc[i];
      const bool keep = (expf(v - m2) * inv > lo) || !ahead(kv, ki, v, i);
      if (!keep) sc[i] = -INFINITY;
    }
    __syncthreads();
  }
  // ordering keys: score + running beam score (+ Gumbel noise for beam-sample); K rounds of block argmax with removal
  const float rs = st->running_scores[r];
  for (int i = tid; i < V; i += kBeamThreads) {
    float key = sc[i] + rs;
    if (sample && key > -INFINITY) key += svbeam::gumbel_noise(pp->seed, cur, r, i);
    sc[i] = key;
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    float bv; int bi;
    block_argmax_behind(sc, V, INFINITY, -1, smf, smi, bv, bi);
    if (tid == 0) {
      const int tok = bi == 0x7fffffff ? 0 : bi;
      // the candidate's log-prob, recomputed from the logit (the key may carry noise)
      const bool sn = use_rp && ((seen[tok >> 5] >> (tok & 31)) & 1u);
      const float s = svbeam::process_logprob((__bfloat162float(lr[tok]) - mx) - logz, sn, rp, sample, T);
      cand_key[r * K + k] = bv;
      cand_val[r * K + k] = bv == -INFINITY ? -INFINITY : s + rs;
      cand_tok[r * K + k] = tok;
      sc[tok] = -INFINITY;
    }
    __syncthreads();
  }
}

// ---- K2: bookkeeping, sequence moves, next-token embedding -----------------------------------------------------------
__global__ void __launch_bounds__(kBeamThreads) beam_step_kernel(const Params* __restrict__ pp, State* st, Plan* plan_out,
                                                                 const float* __restrict__ cand_key,
                                                                 const float* __restrict__ cand_val,
                                                                 const int32_t* __restrict__ cand_tok, int32_t* run_seq,
                                                                 int32_t* fin_seq, GenState* gs, int advance,
                                                                 const bf16* __restrict__ wte,
                                                                 const bf16* __restrict__ wpe, bf16* __restrict__ x, int h,
                                                                 int n_positions, int32_t* next_ids) {
  // every thread reads the flag BEFORE thread 0 can rewrite the state below (it may set done in this very step)
  const int was_done = st->done;
  __syncthreads();
  if (was_done) return;
  __shared__ Plan plan;
  __shared__ int s_oldp, s_pos;
  __shared__ int s_finlen[svbeam::kMaxRows];
  const int tid = threadIdx.x;
  const int nb = pp->nb, K = pp->K, R = pp->B * nb, stride = pp->seq_stride;
  const int pad_fill = pp->pad_id;          // (from the device-resident parameters: the captured graph serves every call)
  if (tid == 0) {
    const Params p = *pp;
    float mval[svbeam::kMaxRows * 2];
    int32_t mbeam[svbeam::kMaxRows * 2], mtok[svbeam::kMaxRows * 2];    // B * K = 2 * B * nb <= 16
    for (int b = 0; b < p.B; ++b)
      svbeam::merge_candidates(p, cand_key + b * nb * K, cand_val + b * nb * K, cand_tok + b * nb * K, mval + b * K,
                               mbeam + b * K, mtok + b * K);
    State s = *st;
    s_oldp = s.parity;
    const int cache_hi = advance ? gs->cur_len : gs->cur_len - 1;
    svbeam::beam_step(p, s, mval, mbeam, mtok, run_seq + (int64_t)s.parity * R * stride, cache_hi, plan);
    *st = s;
    *plan_out = plan;
    if (plan.cont && advance) gs->cur_len += 1;
    gs->done = plan.cont ? 0 : 1;
    s_pos = gs->cur_len;
    for (int r = 0; r < R; ++r) s_finlen[r] = s.fin_len[r];
  }
  __syncthreads();
  const int oldp = s_oldp, newp = oldp ^ 1, L = plan.old_len;
  for (int r = 0; r < R; ++r) {
    const int32_t* src = run_seq + ((int64_t)oldp * R + plan.run_parent[r]) * stride;
    int32_t* dst = run_seq + ((int64_t)newp * R + r) * stride;
    for (int i = tid; i <= L; i += kBeamThreads) dst[i] = i < L ? src[i] : plan.run_tok[r];
    int32_t* dstf = fin_seq + ((int64_t)newp * R + r) * stride;
    if (plan.fin_old[r] >= 0) {
      const int32_t* srcf = fin_seq + ((int64_t)oldp * R + plan.fin_old[r]) * stride;
      const int n = s_finlen[r];
      for (int i = tid; i <= L; i += kBeamThreads) dstf[i] = i < n ? srcf[i] : pad_fill;
    } else {
      const int32_t* srcp = run_seq + ((int64_t)oldp * R + plan.fin_parent[r]) * stride;
      for (int i = tid; i <= L; i += kBeamThreads) dstf[i] = i < L ? srcp[i] : plan.fin_tok[r];
    }
  }
  if (!plan.cont) return;
  // next step's input rows: wte[token] + wpe[position] (bf16 add), as select_fused_kernel does for one beam
  int pos = s_pos;
  pos = pos >= n_positions ? n_positions - 1 : pos;
  const int hv = h >> 3, V = pp->vocab;
  for (int i = tid; i < R * hv; i += kBeamThreads) {
    const int b = i / hv, c = (i % hv) * 8;
    int id = plan.run_tok[b];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    float e[8], q[8];
    unpack8(ldg_cached(wte + (int64_t)id * h + c), e);
    if (wpe) {
      unpack8(ldg_cached(wpe + (int64_t)pos * h + c), q);
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] += q[j];
    }
    *reinterpret_cast<uint4*>(x + (int64_t)b * h + c) = pack8(e);
  }
  if (tid < R) next_ids[tid] = plan.run_tok[tid];
}

// ---- K3: KV suffix copies (phase 0: parent rows -> staging, phase 1: staging -> child rows) -------------------------------
// grid (chunks, rows, layers).  K [row][kvh][tcap][D]: one contiguous run per kv head; V^T [row][kvh][D][tcap]: one short
// run per (kv head, dim).
__global__ void __launch_bounds__(256) beam_kv_copy_kernel(bf16* kc, bf16* vc, bf16* kc2, bf16* vc2, int64_t layer_stride,
                                                           int n_kv, int tcap, int D, const Plan* __restrict__ plan,
                                                           int phase) {
  if (!plan->cont) return;
  const int r = blockIdx.y, layer = blockIdx.z;
  const int src_row = plan->copy_src[r];
  if (src_row < 0) return;
  const int lo = plan->copy_lo[r], hi = plan->copy_hi;
  if (lo > hi) return;
  const int n = hi - lo + 1;
  const int64_t row_elems = (int64_t)n_kv * tcap * D;
  const bf16* ks = (phase == 0 ? kc : kc2) + layer * layer_stride + (phase == 0 ? src_row : r) * row_elems;
  const bf16* vs = (phase == 0 ? vc : vc2) + layer * layer_stride + (phase == 0 ? src_row : r) * row_elems;
  bf16* kd = (phase == 0 ? kc2 : kc) + layer * layer_stride + r * row_elems;
  bf16* vd = (phase == 0 ? vc2 : vc) + layer * layer_stride + r * row_elems;
  const int vec_per_key = D / 8;
  const int nk = n_kv * n * vec_per_key;                       // 16-byte vectors of K
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, ts = gridDim.x * blockDim.x;
  for (int i = t0; i < nk; i += ts) {
    const int kvh = i / (n * vec_per_key), rem = i % (n * vec_per_key);
    const int64_t off = ((int64_t)kvh * tcap + lo) * D + (int64_t)rem * 8;
    *reinterpret_cast<uint4*>(kd + off) = *reinterpret_cast<const uint4*>(ks + off);
  }
  const int nv = n_kv * D * n;                                 // 2-byte elements of V^T
  for (int i = t0; i < nv; i += ts) {
    const int line = i / n, t = i % n;                         // line = kvh * D + dim
    const int64_t off = (int64_t)line * tcap + lo + t;
    vd[off] = vs[off];
  }
}

// ---- launchers ----------------------------------------------------------------------------------------------------
size_t beam_candidates_smem(int vocab) { return (size_t)vocab * 4 + (size_t)((vocab + 31) / 32) * 4; }

cudaError_t beam_init(int vocab) {
  const size_t need = beam_candidates_smem(vocab);
  if (need > 220 * 1024) return cudaErrorInvalidValue;          // a logits row must fit the SM's shared memory
  return cudaFuncSetAttribute(beam_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
}

void launch_beam_candidates(const bf16* logits, int vocab, int rows, const Params* p, const State* st, const int32_t* run_seq,
                            float* cand_key, float* cand_val, int32_t* cand_tok, cudaStream_t st_) {
  beam_candidates_kernel<<<rows, kBeamThreads, beam_candidates_smem(vocab), st_>>>(logits, p, st, run_seq, cand_key, cand_val,
                                                                                   cand_tok);
  count_launch();
}

void launch_beam_step(const Params* p, State* st, Plan* plan, const float* cand_key, const float* cand_val,
                      const int32_t* cand_tok, int32_t* run_seq, int32_t* fin_seq, GenState* gs, int advance,
                      const bf16* wte, const bf16* wpe, bf16* x, int h, int n_positions, int32_t* next_ids,
                      cudaStream_t st_) {
  beam_step_kernel<<<1, kBeamThreads, 0, st_>>>(p, st, plan, cand_key, cand_val, cand_tok, run_seq, fin_seq, gs, advance,
                                                wte, wpe, x, h, n_positions, next_ids);
  count_launch();
}

void launch_beam_kv_copy(bf16* kc, bf16* vc, bf16* kc2, bf16* vc2, int64_t layer_stride, int n_layer, int rows, int n_kv,
                         int tcap, int D, const Plan* plan, cudaStream_t st_) {
  for (int phase = 0; phase < 2; ++phase)
    beam_kv_copy_kernel<<<dim3(4, rows, n_layer), 256, 0, st_>>>(kc, vc, kc2, vc2, layer_stride, n_kv, tcap, D, plan, phase);
  count_launch(2);
}

}  // namespace sv

// =====================================================================================================================
// Host replays of the two device stages (no GPU needed): the test hooks behind tests/test_beam_core.py, which runs a whole
// beam search with them on the CPU oracle's logits and compares with HF generate(num_beams > 1).
extern "C" {

static svbeam::Params params_from_abi(const sv_beam_params* bp, int32_t batch, int32_t vocab, int32_t seq_stride) {
  svbeam::Params p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.nb = bp->num_beams; p.K = 2 * bp->num_beams; p.vocab = vocab; p.max_length = bp->max_new_tokens;
  p.eos_id = bp->eos_token_id;
  p.pad_id = bp->pad_token_id;
  p.n_stop = bp->n_stop_ids;
  for (int i = 0; i < bp->n_stop_ids && i < svbeam::kMaxStop; ++i) p.stop_ids[i] = bp->stop_ids[i];
  p.do_sample = bp->do_sample; p.early_stopping = bp->early_stopping;
  p.min_keep = std::max(2, 1 + (bp->eos_token_id >= 0 ? 1 : 0));
  p.seq_stride = seq_stride;
  p.temperature = bp->temperature; p.top_p = bp->top_p; p.rep_penalty = bp->repetition_penalty;
  p.length_penalty = bp->length_penalty; p.seed = bp->seed;
  return p;
}

int sv_beam_params_check(const sv_beam_params* bp, int32_t batch) {
  if (!bp || batch < 1) return SV_ERR_INVALID;
  if (bp->num_beams < 2 || batch * bp->num_beams > svbeam::kMaxRows) return SV_ERR_INVALID;
  if (bp->max_new_tokens < 1 || bp->n_stop_ids < 0 || bp->n_stop_ids > svbeam::kMaxStop) return SV_ERR_INVALID;
  if (bp->early_stopping < 0 || bp->early_stopping > 2) return SV_ERR_INVALID;
  if (bp->do_sample && !(bp->temperature > 0.f)) return SV_ERR_INVALID;
  if (!(bp->repetition_penalty > 0.f)) return SV_ERR_INVALID;
  return SV_OK;
}

int sv_beam_state_bytes(void) { return (int)sizeof(svbeam::State); }

int sv_beam_state_init_host(const sv_beam_params* bp, int32_t batch, int32_t first_cache_pos, void* state) {
  if (sv_beam_params_check(bp, batch) != SV_OK || !state) return SV_ERR_INVALID;
  svbeam::Params p = params_from_abi(bp, batch, 8, bp->max_new_tokens);
  svbeam::init_state(p, *reinterpret_cast<svbeam::State*>(state), first_cache_pos);
  return SV_OK;
}

// One logits row (fp32 values of the bf16 logits) -> its K = 2 * num_beams best continuations, serially and with HF's
// exact top-p (sort + cumulative sum).  seq: the row's generated tokens so far.
int sv_beam_row_candidates_host(const sv_beam_params* bp, const float* logits, int32_t vocab, const int32_t* seq,
                                int32_t seq_len, float running_score, int32_t step, int32_t row, float* cand_key,
                                float* cand_val, int32_t* cand_tok) {
  if (!bp || !logits || vocab < 1 || !cand_key || !cand_val || !cand_tok) return SV_ERR_INVALID;
  const int K = 2 * bp->num_beams;
  const bool sample = bp->do_sample != 0;
  const int min_keep = std::max(2, 1 + (bp->eos_token_id >= 0 ? 1 : 0));
  std::vector<char> seen(vocab, 0);
  for (int i = 0; i < seq_len; ++i) if (seq[i] >= 0 && seq[i] < vocab) seen[seq[i]] = 1;
  float mx = -INFINITY;
  for (int i = 0; i < vocab; ++i) mx = std::max(mx, logits[i]);
  float z = 0.f;
  for (int i = 0; i < vocab; ++i) z += expf(logits[i] - mx);
  const float logz = logf(z);
  std::vector<float> s(vocab);
  for (int i = 0; i < vocab; ++i)
    s[i] = svbeam::process_logprob((logits[i] - mx) - logz, seq_len > 0 && seen[i], bp->repetition_penalty, sample, bp->temperature);
  std::vector<float> val = s;
  if (sample && bp->top_p < 1.0f) {
    std::vector<int> order(vocab);
    for (int i = 0; i < vocab; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return s[a] < s[b]; });    // ascending, as torch.sort
    float m2 = s[order[vocab - 1]], z2 = 0.f;
    for (int i = 0; i < vocab; ++i) z2 += expf(s[i] - m2);
    float cum = 0.f;
    for (int j = 0; j < vocab; ++j) {
      cum += expf(s[order[j]] - m2) / z2;
      if (cum <= 1.0f - bp->top_p && j < vocab - min_keep) val[order[j]] = -INFINITY;
    }
  }
  std::vector<float> key(vocab);
  for (int i = 0; i < vocab; ++i) {
    key[i] = val[i] + running_score;
    if (sample && key[i] > -INFINITY) key[i] += svbeam::gumbel_noise(bp->seed, step, row, i);
  }
  for (int k = 0; k < K; ++k) {
    int best = -1;
    for (int i = 0; i < vocab; ++i)
      if (best < 0 || key[i] > key[best]) best = i;            // ties: the lower index
    cand_key[k] = key[best];
    cand_val[k] = key[best] == -INFINITY ? -INFINITY : val[best] + running_score;
    cand_tok[k] = best;
    key[best] = -INFINITY;
  }
  return SV_OK;
}

// One bookkeeping step on the host: row candidates [batch * num_beams][K] -> merged -> beam_step.  `state` is the opaque
// State blob; run_seq / fin_seq are the double-buffered sequence arrays [2][batch * num_beams][seq_stride].  plan_out (optional)
// receives the Plan as int32 [sizeof(Plan) / 4].  Returns 1 while the search continues, 0 when it is over, < 0 on error.
int sv_beam_step_host(const sv_beam_params* bp, int32_t batch, int32_t vocab, int32_t seq_stride, void* state,
                      const float* cand_key, const float* cand_val, const int32_t* cand_tok, int32_t* run_seq,
                      int32_t* fin_seq, int32_t cache_hi, int32_t* next_tokens, int32_t* src_rows, int32_t* plan_out) {
  if (sv_beam_params_check(bp, batch) != SV_OK || !state || !cand_key || !cand_val || !cand_tok || !run_seq || !fin_seq)
    return SV_ERR_INVALID;
  svbeam::Params p = params_from_abi(bp, batch, vocab, seq_stride);
  svbeam::State& s = *reinterpret_cast<svbeam::State*>(state);
  const int nb = p.nb, K = p.K, R = batch * nb;
  float mval[svbeam::kMaxRows * 2];
  int32_t mbeam[svbeam::kMaxRows * 2], mtok[svbeam::kMaxRows * 2];
  for (int b = 0; b < batch; ++b)
    svbeam::merge_candidates(p, cand_key + b * nb * K, cand_val + b * nb * K, cand_tok + b * nb * K, mval + b * K, mbeam + b * K,
                             mtok + b * K);
  const int oldp = s.parity;
  svbeam::Plan plan;
  memset(&plan, 0, sizeof(plan));
  svbeam::beam_step(p, s, mval, mbeam, mtok, run_seq + (int64_t)oldp * R * seq_stride, cache_hi, plan);
  const int newp = oldp ^ 1, L = plan.old_len;
  const int fill = bp->pad_token_id;
  for (int r = 0; r < R; ++r) {                                // the same moves beam_step_kernel makes
    const int32_t* src = run_seq + ((int64_t)oldp * R + plan.run_parent[r]) * seq_stride;
    int32_t* dst = run_seq + ((int64_t)newp * R + r) * seq_stride;
    for (int i = 0; i <= L; ++i) dst[i] = i < L ? src[i] : plan.run_tok[r];
    int32_t* dstf = fin_seq + ((int64_t)newp * R + r) * seq_stride;
    if (plan.fin_old[r] >= 0) {
      const int32_t* srcf = fin_seq + ((int64_t)oldp * R + plan.fin_old[r]) * seq_stride;
      for (int i = 0; i <= L; ++i) dstf[i] = i < s.fin_len[r] ? srcf[i] : fill;
    } else {
      const int32_t* srcp = run_seq + ((int64_t)oldp * R + plan.fin_parent[r]) * seq_stride;
      for (int i = 0; i <= L; ++i) dstf[i] = i < L ? srcp[i] : plan.fin_tok[r];
    }
    if (next_tokens) next_tokens[r] = plan.run_tok[r];
    if (src_rows) src_rows[r] = plan.run_parent[r];
  }
  if (plan_out) memcpy(plan_out, &plan, sizeof(plan));
  return plan.cont;
}

// Final read-out of a State blob: parity of the live sequence buffers and the lengths of the finished hypotheses.
int sv_beam_state_read_host(const void* state, int32_t* parity, int32_t* cur_len, int32_t* fin_len8, float* beam_scores8) {
  if (!state) return SV_ERR_INVALID;
  const svbeam::State& s = *reinterpret_cast<const svbeam::State*>(state);
  if (parity) *parity = s.parity;
  if (cur_len) *cur_len = s.cur_len;
  for (int r = 0; r < svbeam::kMaxRows; ++r) {
    if (fin_len8) fin_len8[r] = s.fin_len[r];
    if (beam_scores8) beam_scores8[r] = s.beam_scores[r];
  }
  return SV_OK;
}

}  // extern "C"
