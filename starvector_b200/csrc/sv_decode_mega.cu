// Persistent decode kernel: `nsteps` whole tokens (all layers, lm_head, token selection) in ONE
// cooperative launch, one CTA per SM.  This is the B=1..8 greedy hot loop of the BASELINE workloads.
//
// Why: a decode step streams 2.24 GB of weights but is cut into ~146 dependent phases of a few MB
// each; as separate kernels every phase pays launch ramp + drain (measured: 122 launches/step,
// ~25% of the HBM roofline).  Here the phase boundaries are grid barriers that only the compute
// warps take, while a dedicated producer warp keeps streaming weights through them:
//
//   warp 8 (producer, one elected lane): walks the STATIC weight schedule of the whole launch
//       (layer 0 c_attn tiles, c_proj, c_fc, mlp.c_proj, layer 1 ..., lm_head, next token ...) and
//       copies [R rows x 1024 k] weight slabs into a 5-slot shared-memory ring with
//       cp.async.bulk (TMA bulk copy, one instruction per 2 KB row segment, completion on the
//       slot's "full" mbarrier).  It never waits for activations, so ~165 KB per SM (24 MB chip
//       wide) of HBM reads stay in flight across barriers, LayerNorm prologues and attention.
//   warps 0-7 (consumers): wait on "full", read 128-bit MMA fragments from the slot (row pitch
//       = 2 KB + 64 B, bank-conflict free), multiply with register-resident activation fragments
//       (mma.sync m16n8k16, weights = A, the <= 8 image rows = B), release the slot ("empty"
//       mbarrier, one arrive per warp), and at tile end do the deterministic cross-warp split-K
//       reduction + the reference's bf16 epilogue (bias, gelu, residual, KV-cache append, argmax).
//   Attention (split-KV, MQA: the 16 query heads are the MMA M dimension) reads K / V^T with
//       L2-only loads; CTA-level partials are merged by a second short phase.
//
// Work split: N output rows are tiled R <= 16 rows at a time so that every CTA owns the same
// number of rows (2048 -> 147 x 14, 2304 -> 144 x 16, 8192 -> 147 x 4 x 14, 49156 -> 147 x 21 x 16).
// All cross-CTA data (activations, partials, state) is written with plain stores and read with
// ld.global.cg after a release/acquire grid barrier; weights use the async proxy only.
// Every wait (mbarrier, grid barrier) is bounded and traps instead of hanging the GPU.
#include <cstdio>
#include <cstdlib>

#include "sv_kernels.h"
#include "sv_ring.cuh"
#include "sv_select.cuh"

namespace sv {
namespace mega {


// ---- grid barrier among the consumer threads of all CTAs (monotonic counter, wrap-safe compare)
SV_DEVINL void grid_barrier(unsigned int* ctr, unsigned int& target, int ncta) {
  target += (unsigned int)ncta;
  consumer_sync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned int v;
    for (uint32_t it = 0;; ++it) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if ((int)(v - target) >= 0) break;
      if (it > (1u << 24)) __trap();
    }
    __threadfence();
  }
  consumer_sync();
}

enum { EPI_PLAIN = 0, EPI_QKV = 1, EPI_LMHEAD = 2 };

struct Ctx {
  const Args* a;
  uint8_t* smem;
  int cta, ncta, warp, lane, g, t;
  float* red;     // [2][NWC][16][8]
  float* stat;    // [NWC][8]
  // optional (per-phase ring kernels): parameters staged into shared memory BEFORE the programmatic-dependency wait, so that
  // their HBM misses overlap the previous kernel's tail instead of sitting on this kernel's critical path
  uint32_t ln_s = 0;            // shared address of [ln_w row | ln_b row] (K bf16 each), 0 = read them from global
  const float* bias_s = nullptr;   // [tile][16] biases of this CTA's output rows
};

// ---- consumer: one GEMV phase  Y[B,N] = epi( LN?(X)[B,K] . W[N,K]^T )
// LN_BIGK compiles in the LayerNorm path for K > 2048 (v2); v1 kernels are instantiated without it so their register
// allocation is untouched.
template <bool HAS_LN, int EPI, bool LN_BIGK = false>
SV_DEVINL void gemv_phase(const Ctx& cx, Ring& r, const bf16* __restrict__ X, const bf16* __restrict__ bias,
                          const bf16* res, bf16* Y, int N, int K, int act, const bf16* __restrict__ ln_w,
                          const bf16* __restrict__ ln_b, const Layer* L) {
  const Args& a = *cx.a;
  const Plan p = make_plan(N, K, cx.cta, cx.ncta);
  const int warp = cx.warp, g = cx.g, t = cx.t;
  const int cps = p.KS >> 5;                         // 32-wide chunks per slot row
  const int cpws = (cps + NWC - 1) / NWC;            // chunks per warp per slot (<= CPW)
  const bool row_ok = g < a.B;
  const bf16* xp = X + (int64_t)(row_ok ? g : 0) * K + 8 * t;
  const bool big_k = p.nstg > 2;

  // activations for the whole phase live in registers when K <= 2048 (8 fragments per lane)
  uint4 xr[2 * CPW];
#pragma unroll
  for (int i = 0; i < 2 * CPW; ++i) xr[i] = make_uint4(0u, 0u, 0u, 0u);
  if (!big_k && p.ntile > 0) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int j = 0; j < CPW; ++j) {
        const int cl = warp + NWC * j;
        const bool okc = ks < p.nstg && j < cpws && cl < cps;
        if (okc && row_ok) xr[ks * CPW + j] = ldcg16(xp + (ks * cps + cl) * 32);
      }
    }
    if constexpr (HAS_LN) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 2 * CPW; ++i) {
        float f[8];
        unpack8(xr[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[j];
      }
      s = quad_sum(s);
      if (t == 0) cx.stat[warp * 8 + g] = s;
      consumer_sync();
      float mean = 0.f;
#pragma unroll
      for (int w = 0; w < NWC; ++w) mean += cx.stat[w * 8 + g];
      mean /= (float)K;
      float q = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const bool okc = ks < p.nstg && j < cpws && (warp + NWC * j) < cps;
          if (okc) {
            float f[8];
            unpack8(xr[ks * CPW + j], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float dlt = f[e] - mean; q += dlt * dlt; }
          }
        }
      }
      q = quad_sum(q);
      consumer_sync();
      if (t == 0) cx.stat[warp * 8 + g] = q;
      consumer_sync();
      float var = 0.f;
#pragma unroll
      for (int w = 0; w < NWC; ++w) var += cx.stat[w * 8 + g];
      const float rstd = 1.0f / sqrtf(var / (float)K + a.ln_eps);
      // (the weight ring keeps HBM busy on its own, so the LN affine is fetched late to save registers)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cl = warp + NWC * j;
          const bool okc = ks < p.nstg && j < cpws && cl < cps;
          float f[8], wf[8], bfv[8];
          unpack8(xr[ks * CPW + j], f);
          const int ch = okc ? ks * cps + cl : 0;
          if (cx.ln_s) {
            unpack8(lds16(cx.ln_s + (ch * 32 + 8 * t) * 2), wf);
            unpack8(lds16(cx.ln_s + (K + ch * 32 + 8 * t) * 2), bfv);
          } else {
            unpack8(ldg_cached(ln_w + ch * 32 + 8 * t), wf);
            unpack8(ldg_cached(ln_b + ch * 32 + 8 * t), bfv);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = (row_ok && okc) ? (f[e] - mean) * rstd * wf[e] + bfv[e] : 0.f;
          xr[ks * CPW + j] = pack8(f);       // ln output is a bf16 tensor in the reference; 0 on padded chunks
        }
      }
    }
  }

  // LayerNorm with K > 2048 (StarCoder2: H = 4608): row statistics in two streaming passes over x (L2 resident),
  // the normalisation itself happens per slab inside the MMA loop.
  float ln_mean = 0.f, ln_rstd = 1.f;
  if constexpr (HAS_LN && LN_BIGK) {
    if (big_k && p.ntile > 0) {
      float sv = 0.f;
      for (int ks = 0; ks < p.nstg; ++ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cl = warp + NWC * j;
          if (row_ok && j < cpws && cl < cps) {
            float f[8];
            unpack8(ldcg16(xp + (ks * cps + cl) * 32), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) sv += f[e];
          }
        }
      }
      sv = quad_sum(sv);
      if (t == 0) cx.stat[warp * 8 + g] = sv;
      consumer_sync();
#pragma unroll
      for (int w = 0; w < NWC; ++w) ln_mean += cx.stat[w * 8 + g];
      ln_mean /= (float)K;
      float q = 0.f;
      for (int ks = 0; ks < p.nstg; ++ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cl = warp + NWC * j;
          if (row_ok && j < cpws && cl < cps) {
            float f[8];
            unpack8(ldcg16(xp + (ks * cps + cl) * 32), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float dlt = f[e] - ln_mean; q += dlt * dlt; }
          }
        }
      }
      q = quad_sum(q);
      consumer_sync();
      if (t == 0) cx.stat[warp * 8 + g] = q;
      consumer_sync();
      float var = 0.f;
#pragma unroll
      for (int w = 0; w < NWC; ++w) var += cx.stat[w * 8 + g];
      ln_rstd = 1.0f / sqrtf(var / (float)K + a.ln_eps);
    }
  }

  float c[4] = {0.f, 0.f, 0.f, 0.f};
  int pos_now = 0;
  if constexpr (EPI == EPI_QKV) pos_now = __ldcg(&a.state->cur_len);         // read here, not behind the last MMA
  for (int tl = 0; tl < p.ntile; ++tl) {
    const int tile = p.tile0 + tl;
    // the epilogue thread's residual value: requested now, used after the MMAs (an L2 round trip off the tail)
    float res_pre = 0.f;
    {
      const int n_ = threadIdx.x & 15, mm_ = threadIdx.x >> 4, col_ = tile * p.R + n_;
      if (res != nullptr && threadIdx.x < 128 && n_ < p.R && col_ < N && mm_ < a.B) res_pre = __bfloat162float(__ldcg(res + (int64_t)mm_ * N + col_));
    }
    if (!big_k) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        if (ks < p.nstg) {
          mbar_wait(r.full0 + 8u * r.slot, r.phase);
          const uint32_t sb = r.base + r.slot * SLOT_BYTES + g * p.pitch + t * 16;
#pragma unroll
          for (int j = 0; j < CPW; ++j) {
            const int cl = warp + NWC * j;
            if (j < cpws && cl < cps) {
              const uint4 lo = lds16(sb + cl * 64), hi = lds16(sb + 8 * p.pitch + cl * 64);
              const uint4 xv = xr[ks * CPW + j];
              mma_bf16_16816(c, lo.x, hi.x, lo.y, hi.y, xv.x, xv.y);
              mma_bf16_16816(c, lo.z, hi.z, lo.w, hi.w, xv.z, xv.w);
            }
          }
          __syncwarp();
          if (cx.lane == 0) mbar_arrive(r.empty0 + 8u * r.slot);
          r.advance();
        }
      }
    } else {
      // K > 2048: activation fragments are fetched per slab from L2, one slab ahead of their use (with HAS_LN the
      // LayerNorm affine of the same columns rides along and the fragment is normalised after the slab's MMAs).
      constexpr bool LNB = HAS_LN && LN_BIGK;
      uint4 xc[CPW], xn[CPW], wn[LNB ? CPW : 1], bn[LNB ? CPW : 1];
      auto fetch = [&](int ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cl = warp + NWC * j;
          const bool okc = ks < p.nstg && j < cpws && cl < cps;
          const int ch = okc ? ks * cps + cl : 0;
          xn[j] = (row_ok && okc) ? ldcg16(xp + ch * 32) : make_uint4(0u, 0u, 0u, 0u);
          if constexpr (LNB) {
            wn[j] = okc ? ldg_cached(ln_w + ch * 32 + 8 * t) : make_uint4(0u, 0u, 0u, 0u);
            bn[j] = okc ? ldg_cached(ln_b + ch * 32 + 8 * t) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      };
      auto promote = [&](int ks) {             // xn (raw) -> xc (what the MMAs consume)
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          if constexpr (LNB) {
            const bool okc = ks < p.nstg && j < cpws && (warp + NWC * j) < cps;
            float f[8], wf[8], bfv[8];
            unpack8(xn[j], f); unpack8(wn[j], wf); unpack8(bn[j], bfv);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = (row_ok && okc) ? (f[e] - ln_mean) * ln_rstd * wf[e] + bfv[e] : 0.f;
            xc[j] = pack8(f);
          } else {
            xc[j] = xn[j];
          }
        }
      };
      fetch(0);
      promote(0);
      for (int ks = 0; ks < p.nstg; ++ks) {
        fetch(ks + 1);
        mbar_wait(r.full0 + 8u * r.slot, r.phase);
        const uint32_t sb = r.base + r.slot * SLOT_BYTES + g * p.pitch + t * 16;
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cl = warp + NWC * j;
          if (j < cpws && cl < cps) {
            const uint4 lo = lds16(sb + cl * 64), hi = lds16(sb + 8 * p.pitch + cl * 64);
            mma_bf16_16816(c, lo.x, hi.x, lo.y, hi.y, xc[j].x, xc[j].y);
            mma_bf16_16816(c, lo.z, hi.z, lo.w, hi.w, xc[j].z, xc[j].w);
          }
        }
        __syncwarp();
        if (cx.lane == 0) mbar_arrive(r.empty0 + 8u * r.slot);
        r.advance();
        promote(ks + 1);
      }
    }
    // ---- tile finished: deterministic cross-warp split-K reduction + epilogue
    float* rd = cx.red + (tl & 1) * (NWC * 16 * 8);
    rd[(warp * 16 + g) * 8 + 2 * t] = c[0]; rd[(warp * 16 + g) * 8 + 2 * t + 1] = c[1];
    rd[(warp * 16 + g + 8) * 8 + 2 * t] = c[2]; rd[(warp * 16 + g + 8) * 8 + 2 * t + 1] = c[3];
    c[0] = c[1] = c[2] = c[3] = 0.f;
    consumer_sync();
    if (threadIdx.x < 128) {
      const int n = threadIdx.x & 15, mm = threadIdx.x >> 4;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < NWC; ++w) acc += rd[(w * 16 + n) * 8 + mm];
      const int col = tile * p.R + n;
      const bool ok = n < p.R && col < N && mm < a.B;
      float v = 0.f;
      if (ok) {
        const float bv = bias ? (cx.bias_s ? cx.bias_s[tl * 16 + n] : __bfloat162float(bias[col])) : 0.f;
        const float rv = res_pre;
        v = epilogue_elem(acc, bv, act, res != nullptr, rv);
        const bf16 vb = __float2bfloat16_rn(v);
        Y[(int64_t)mm * N + col] = vb;
        if constexpr (EPI == EPI_QKV) {
          const int q_cols = a.n_head * D, j = col - q_cols;
          const int pos = pos_now;
          if (j >= 0 && pos < a.tcap) {
            if (j < a.n_kv * D) {
              const int kvh = j / D, dim = j % D;
              L->kc[(((int64_t)mm * a.n_kv + kvh) * a.tcap + pos) * D + dim] = vb;
            } else {
              const int jj = j - a.n_kv * D, kvh = jj / D, dim = jj % D;
              L->vc[(((int64_t)mm * a.n_kv + kvh) * D + dim) * a.tcap + pos] = vb;
            }
          }
        }
      }
      if constexpr (EPI == EPI_LMHEAD) {
        float bv = ok ? v : -INFINITY;
        int bi = ok ? col : 0x7fffffff;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (n == 0 && mm < a.B) {
          a.amax_val[(int64_t)tile * 8 + mm] = bv;
          a.amax_idx[(int64_t)tile * 8 + mm] = bi;
        }
      }
    }
    // red[] is double-buffered by tile parity: one barrier per tile
  }
}


// ---- attention phase A: CTA-level partials.  item = (image, kv head, key chunk c), strided over CTAs.
SV_DEVINL void attention_partials(const Ctx& cx, const Layer* L) {
  const Args& a = *cx.a;
  const int warp = cx.warp, g = cx.g, t = cx.t;
  const int group = a.n_head / a.n_kv;
  const int nkeys = __ldcg(&a.state->cur_len) + 1;
  const int blocks = (nkeys + 31) / 32;
  const int per = (blocks + a.att_ncta - 1) / a.att_ncta;
  const int nact = (blocks + per - 1) / per;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  float* att = reinterpret_cast<float*>(cx.smem + OFF_ATT);
  const int nitems = a.B * a.n_kv * nact;
  for (int item = cx.cta; item < nitems; item += cx.ncta) {
    const int c = item % nact, bk = item / nact, kvh = bk % a.n_kv, b = bk / a.n_kv;
    const int blk0 = c * per, blk1 = min(blocks, blk0 + per);
    float acc[D / 8][4], mrow[2], lrow[2];
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
    mrow[0] = mrow[1] = -INFINITY; lrow[0] = lrow[1] = 0.f;
    if (blk0 + warp < blk1) {
      const bf16* qrow = a.qkv + (int64_t)b * a.qkv_cols + (int64_t)kvh * group * D;
      uint32_t qa[D / 16][4];
#pragma unroll
      for (int jj = 0; jj < D / 32; ++jj) {
        uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = make_uint4(0u, 0u, 0u, 0u);
        if (g < group) lo = ldcg16(qrow + (int64_t)g * D + 32 * jj + 8 * t);
        if (g + 8 < group) hi = ldcg16(qrow + (int64_t)(g + 8) * D + 32 * jj + 8 * t);
        qa[2 * jj][0] = lo.x; qa[2 * jj][1] = hi.x; qa[2 * jj][2] = lo.y; qa[2 * jj][3] = hi.y;
        qa[2 * jj + 1][0] = lo.z; qa[2 * jj + 1][1] = hi.z; qa[2 * jj + 1][2] = lo.w; qa[2 * jj + 1][3] = hi.w;
      }
      const bf16* kb_ = L->kc + (int64_t)bk * a.tcap * D;
      const bf16* vb_ = L->vc + (int64_t)bk * D * a.tcap;
      for (int blk = blk0 + warp; blk < blk1; blk += NWC)
        attn_block(qa, kb_, vb_, a.tcap, blk * 32, min(nkeys, blk * 32 + 32), scale_log2, acc, mrow, lrow, g, t);
    }
    float lq[2] = {quad_sum(lrow[0]), quad_sum(lrow[1])};
    // tree merge 8 -> 4 -> 2 -> 1 warps through a 4-partial shared buffer
    if (warp >= 4) attn_store_to(att + (warp - 4) * PSZ, acc, mrow, lq, g, t);
    consumer_sync();
    if (warp < 4) attn_merge_from(att + warp * PSZ, acc, mrow, lq, g, t);
    consumer_sync();
    if (warp == 2 || warp == 3) attn_store_to(att + (warp - 2) * PSZ, acc, mrow, lq, g, t);
    consumer_sync();
    if (warp < 2) attn_merge_from(att + warp * PSZ, acc, mrow, lq, g, t);
    consumer_sync();
    if (warp == 1) attn_store_to(att, acc, mrow, lq, g, t);
    consumer_sync();
    if (warp == 0) {
      attn_merge_from(att, acc, mrow, lq, g, t);
      attn_store_to(a.attn_partial + ((int64_t)bk * a.att_ncta + c) * PSZ, acc, mrow, lq, g, t);
    }
    consumer_sync();
  }
}

// ---- attention phase B: merge the CTA partials, one output element per thread (all CTAs)
SV_DEVINL void attention_merge(const Ctx& cx) {
  const Args& a = *cx.a;
  const int group = a.n_head / a.n_kv;
  const int nkeys = __ldcg(&a.state->cur_len) + 1;
  const int blocks = (nkeys + 31) / 32;
  const int per = (blocks + a.att_ncta - 1) / a.att_ncta;
  const int nact = (blocks + per - 1) / per;
  const int total = a.B * a.n_head * D;
  for (int o = cx.cta * NCT + (int)threadIdx.x; o < total; o += cx.ncta * NCT) {
    const int dim = o % D, head = (o / D) % a.n_head, b = o / (D * a.n_head);
    const int kvh = head / group, r = head % group;
    const float* p0 = a.attn_partial + ((int64_t)(b * a.n_kv + kvh) * a.att_ncta) * PSZ;
    float M = -INFINITY;
    for (int c = 0; c < nact; ++c) M = fmaxf(M, __ldcg(p0 + (int64_t)c * PSZ + r));
    float Lsum = 0.f, A = 0.f;
    for (int c = 0; c < nact; ++c) {
      const float m = __ldcg(p0 + (int64_t)c * PSZ + r);
      const float sc = (m == -INFINITY) ? 0.f : exp2f(m - M);
      Lsum += __ldcg(p0 + (int64_t)c * PSZ + 16 + r) * sc;
      A += __ldcg(p0 + (int64_t)c * PSZ + 32 + r * D + dim) * sc;
    }
    a.attn[(int64_t)b * a.n_head * D + head * D + dim] = __float2bfloat16_rn(A / Lsum);
  }
}

// ---- token selection (CTA 0): argmax partials (or penalised full scan) -> HF bookkeeping -> embedding
SV_DEVINL void select_phase(const Ctx& cx, int ntiles) {
  const Args& a = *cx.a;
  if (cx.cta != 0) return;
  if (__ldcg(&a.state->done)) return;
  AmaxPair* sm = reinterpret_cast<AmaxPair*>(cx.red);
  int* s_tok = reinterpret_cast<int*>(cx.smem + OFF_TOK);
  const int tid = threadIdx.x;
  const float rp = a.params->rep_penalty;
  for (int b = 0; b < a.B; ++b) {
    AmaxPair best{-INFINITY, 0x7fffffff};
    if (rp == 1.0f) {
      for (int i = tid; i < ntiles; i += NCT)
        best = amax_better(best, AmaxPair{__ldcg(a.amax_val + (int64_t)i * 8 + b), __ldcg(a.amax_idx + (int64_t)i * 8 + b)});
    } else {
      const bf16* lr = a.logits + (int64_t)b * a.vocab;
      const uint8_t* sr = a.seen + (int64_t)b * a.vocab;
      for (int i = tid; i < a.vocab; i += NCT) {
        float v = __bfloat162float(__ldcg(lr + i));
        if (__ldcg(sr + i)) v = v < 0.f ? v * rp : v / rp;
        best = amax_better(best, AmaxPair{v, i});
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      AmaxPair other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = amax_better(best, other);
    }
    consumer_sync();
    if (cx.lane == 0) sm[cx.warp] = best;
    consumer_sync();
    if (tid == 0) {
      for (int w = 1; w < NWC; ++w) best = amax_better(best, sm[w]);
      s_tok[b] = best.i == 0x7fffffff ? 0 : best.i;
    }
  }
  consumer_sync();
  if (tid == 0) {
    select_apply_tokens(s_tok, a.B, a.vocab, a.state, a.params, a.seen, a.next_ids, a.out_ids, 1);
    __threadfence();
  }
  consumer_sync();
  int pos = a.state->cur_len;        // written by this CTA's thread 0 just above (same-CTA visibility after bar)
  pos = pos >= a.n_positions ? a.n_positions - 1 : pos;
  const int hv = a.H >> 3;
  for (int i = tid; i < a.B * hv; i += NCT) {
    const int b = i / hv, col = (i % hv) * 8;
    int id = s_tok[b];
    id = id < 0 ? 0 : (id >= a.vocab ? a.vocab - 1 : id);
    float e[8], q[8];
    unpack8(ldg_cached(a.wte + (int64_t)id * a.H + col), e);
    unpack8(ldg_cached(a.wpe + (int64_t)pos * a.H + col), q);
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] += q[j];
    *reinterpret_cast<uint4*>(a.x + (int64_t)b * a.H + col) = pack8(e);
  }
}

// REALLOC = true is the warp-specialised register split (opt-in SV_MEGA=2): the CTA is launched with three warpgroups
// (384 threads x 168 registers = the whole register file), warpgroup 2 (the producer warp + three idle warps) gives
// registers back with `setmaxnreg.dec` and the two consumer warpgroups take them with `setmaxnreg.inc`
// (128 x (168-56) = 256 x (224-168)), so the GEMV/attention code is no longer compiled against the 168-register cap that
// nine equal warps impose (one SM sub-partition would have to hold three of them).
constexpr int NTHREADS_REALLOC = NCT + 128;
constexpr int REGS_PRODUCER = 56, REGS_CONSUMER = 224;

template <bool REALLOC>
__global__ void __launch_bounds__(REALLOC ? NTHREADS_REALLOC : NTHREADS, 1) decode_mega_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  Ring ring;
  ring.base = smem_u32(smem);
  ring.full0 = smem_u32(smem + OFF_BAR);
  ring.empty0 = ring.full0 + 8u * STAGES;
  ring.slot = 0; ring.phase = 0; ring.nslots = STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(ring.full0 + 8u * s, 1); mbar_init(ring.empty0 + 8u * s, NWC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int qkv_n = a.qkv_cols;
  if (warp >= NWC) {
    if constexpr (REALLOC) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
      if (warp > NWC) return;          // the other three warps of the producer warpgroup only exist to hand registers over
    }
    // =========================== producer ===========================
    for (int s = 0; s < a.nsteps; ++s) {
      for (int l = 0; l < a.n_layer; ++l) {
        const Layer* L = a.layers + l;
        produce_phase(ring, L->attn_w, qkv_n, a.H, cta, ncta, lane);
        produce_phase(ring, L->proj_w, a.H, a.H, cta, ncta, lane);
        produce_phase(ring, L->fc_w, a.I, a.H, cta, ncta, lane);
        produce_phase(ring, L->fc2_w, a.H, a.I, cta, ncta, lane);
      }
      produce_phase(ring, a.lm_head, a.vocab, a.H, cta, ncta, lane);
    }
    return;   // in-flight bulk copies are all consumed (and thus complete) before the consumers exit
  }
  // =========================== consumers ===========================
  if constexpr (REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_CONSUMER));
  Ctx cx;
  cx.a = &a; cx.smem = smem; cx.cta = cta; cx.ncta = ncta; cx.warp = warp; cx.lane = lane; cx.g = lane >> 2; cx.t = lane & 3;
  cx.red = reinterpret_cast<float*>(smem + OFF_RED);
  cx.stat = reinterpret_cast<float*>(smem + OFF_STAT);
  unsigned int target = 0;
  const int ntiles_lm = make_plan(a.vocab, a.H, 0, ncta).ntiles;
  int dbg_i = 0;
  const bool dbg_on = a.dbg != nullptr && cta == 0 && threadIdx.x == 0;
#define SV_STAMP() do { if (dbg_on && s == 0 && dbg_i < 1000) a.dbg[dbg_i++] = clock64(); } while (0)
#define SV_GRID_BARRIER() do { SV_STAMP(); grid_barrier(a.barrier_ctr, target, ncta); SV_STAMP(); } while (0)
  for (int s = 0; s < a.nsteps; ++s) {
    for (int l = 0; l < a.n_layer; ++l) {
      const Layer* L = a.layers + l;
      gemv_phase<true, EPI_QKV>(cx, ring, a.x, L->attn_b, nullptr, a.qkv, qkv_n, a.H, 0, L->ln1_w, L->ln1_b, L);
      SV_GRID_BARRIER();
      attention_partials(cx, L);
      SV_GRID_BARRIER();
      attention_merge(cx);
      SV_GRID_BARRIER();
      gemv_phase<false, EPI_PLAIN>(cx, ring, a.attn, L->proj_b, a.x, a.x, a.H, a.H, 0, nullptr, nullptr, L);
      SV_GRID_BARRIER();
      gemv_phase<true, EPI_PLAIN>(cx, ring, a.x, L->fc_b, nullptr, a.h, a.I, a.H, 2 /*gelu_tanh*/, L->ln2_w, L->ln2_b, L);
      SV_GRID_BARRIER();
      gemv_phase<false, EPI_PLAIN>(cx, ring, a.h, L->fc2_b, a.x, a.x, a.H, a.I, 0, nullptr, nullptr, L);
      SV_GRID_BARRIER();
    }
    gemv_phase<true, EPI_LMHEAD>(cx, ring, a.x, nullptr, nullptr, a.logits, a.vocab, a.H, 0, a.lnf_w, a.lnf_b, nullptr);
    SV_GRID_BARRIER();
    select_phase(cx, ntiles_lm);
    SV_GRID_BARRIER();
  }
}

// ------------------------------------------------------------------------------------------
// The same weight ring as ONE-PHASE kernels (per-phase CUDA-graph decode path): a producer warp
// streams this GEMV's slabs through shared memory (starting before the PDL dependency wait, weights
// are immutable), 8 consumer warps do LayerNorm prologue / MMA / epilogue.  ~165 KB of HBM reads
// in flight per SM instead of the 64 KB a register-landing GEMV can hold.
struct RingGemvArgs {
  Args a;            // B, ln_eps, n_head, n_kv, tcap, state, amax_* (fields the epilogues read)
  Layer L;           // kc / vc for the QKV epilogue
  const bf16 *X, *W, *bias, *res, *ln_w, *ln_b;
  bf16* Y;
  int N, K, act;
  int nslots;        // ring depth of THIS launch
  const void* next_w;            // weights the NEXT GEMV of the step will stream (immutable): prefetched into L2 here,
  unsigned long long next_bytes; // so HBM keeps streaming across the kernel boundary / the attention kernel
};

// Each CTA asks L2 to fetch its 1/ncta share of the next weight matrix (cp.async.bulk.prefetch.L2, 4 KB pieces,
// one per lane).  The 126 MB L2 holds the current (<= 33.5 MB) and the next (<= 33.5 MB) matrices of a layer.
SV_DEVINL void l2_prefetch_share(const void* base, unsigned long long bytes, int cta, int ncta, int lane) {
  if (base == nullptr || bytes == 0) return;
  const unsigned long long per = ((bytes + ncta - 1) / ncta + 4095ull) & ~4095ull;
  const unsigned long long lo = (unsigned long long)cta * per;
  if (lo >= bytes) return;
  const unsigned long long hi = lo + per < bytes ? lo + per : bytes;
  const char* p = reinterpret_cast<const char*>(base);
  for (unsigned long long off = lo + (unsigned long long)lane * 4096ull; off < hi; off += 32ull * 4096ull) {
    const unsigned long long n = hi - off < 4096ull ? ((hi - off) & ~15ull) : 4096ull;
    if (n >= 16)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"((uint32_t)n) : "memory");
  }
}
constexpr int RING_BIAS_TILES = 8;     // biases staged for up to this many tiles per CTA (mlp.c_fc has 4)
SV_DEVINL constexpr int ring_smem_bytes(int nslots) {
  return nslots * SLOT_BYTES + RED_BYTES + NWC * 8 * 4 + 2 * 8 * 8 + 16 + 2 * 2 * KS_MAX * 2 + RING_BIAS_TILES * 16 * 4 + 256;
}

// (A 2-CTA/SM register budget (96 regs) so that consecutive kernels co-reside under PDL was measured 25% slower.)
template <bool HAS_LN, int EPI, bool LN_BIGK = false>
__global__ void __launch_bounds__(NTHREADS, RING_MINBLOCKS) gemv_ring_kernel(const RingGemvArgs ra) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int off_red = ra.nslots * SLOT_BYTES, off_stat = off_red + RED_BYTES, off_bar = off_stat + NWC * 8 * 4;
  Ring ring;
  ring.base = smem_u32(smem);
  ring.full0 = smem_u32(smem + off_bar);
  ring.empty0 = ring.full0 + 8u * 8;
  ring.slot = 0; ring.phase = 0; ring.nslots = (uint32_t)ra.nslots;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ra.nslots; ++s) { mbar_init(ring.full0 + 8u * s, 1); mbar_init(ring.empty0 + 8u * s, NWC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == NWC) {
    produce_phase(ring, ra.W, ra.N, ra.K, cta, ncta, lane);     // no dependency on the previous kernel
    l2_prefetch_share(ra.next_w, ra.next_bytes, cta, ncta, lane);
    return;
  }
  Ctx cx;
  cx.a = &ra.a; cx.smem = smem; cx.cta = cta; cx.ncta = ncta; cx.warp = warp; cx.lane = lane; cx.g = lane >> 2; cx.t = lane & 3;
  cx.red = reinterpret_cast<float*>(smem + off_red);
  cx.stat = reinterpret_cast<float*>(smem + off_stat);
  // immutable parameters (LayerNorm affine, biases of this CTA's rows) are staged into shared memory before the wait on the
  // previous kernel: their HBM misses (~1 us each, two per LayerNorm kernel, one per epilogue) overlap that kernel's tail
  {
    const int off_par = (off_bar + 2 * 8 * 8 + 15) & ~15;
    if (HAS_LN && !LN_BIGK && ra.K <= 2 * KS_MAX) {
      cx.ln_s = smem_u32(smem + off_par);
      const int nv = ra.K / 8;                                   // 16-byte vectors per row
      for (int i = threadIdx.x; i < 2 * nv; i += NCT)
        *reinterpret_cast<uint4*>(smem + off_par + i * 16) = ldg_cached((i < nv ? ra.ln_w : ra.ln_b) + (i % nv) * 8);
    }
    if (ra.bias != nullptr) {
      float* bs = reinterpret_cast<float*>(smem + off_par + 2 * 2 * KS_MAX * 2);
      const Plan p = make_plan(ra.N, ra.K, cta, ncta);
      if (p.tpc <= RING_BIAS_TILES) {
        for (int i = threadIdx.x; i < p.ntile * 16; i += NCT) {
          const int col = (p.tile0 + i / 16) * p.R + (i % 16);
          bs[i] = (i % 16) < p.R && col < ra.N ? __bfloat162float(ra.bias[col]) : 0.f;
        }
        cx.bias_s = bs;
      }
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  gemv_phase<HAS_LN, EPI, LN_BIGK>(cx, ring, ra.X, ra.bias, ra.res, ra.Y, ra.N, ra.K, ra.act, ra.ln_w, ra.ln_b, &ra.L);
}

}  // namespace mega

// ---- host side
static int g_mega_ncta = 0;
static bool g_mega_realloc_ok = false;
static char g_mega_why[256] = "decode_mega_init not called";
const char* decode_mega_status() { return g_mega_why; }

cudaError_t decode_mega_init() {
  cudaError_t e = cudaFuncSetAttribute(mega::decode_mega_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       mega::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  // the experimental variant must never make engine creation fail: its errors only disable it
  bool realloc_attr_ok = cudaFuncSetAttribute(mega::decode_mega_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              mega::SMEM_BYTES) == cudaSuccess;
  if (!realloc_attr_ok) cudaGetLastError();
  int dev = 0, nsm = 0, per_sm = 0, per_sm_realloc = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mega::decode_mega_kernel<false>, mega::NTHREADS, mega::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  if (!realloc_attr_ok || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_realloc, mega::decode_mega_kernel<true>,
                                                                         mega::NTHREADS_REALLOC, mega::SMEM_BYTES) != cudaSuccess) {
    per_sm_realloc = 0;
    cudaGetLastError();
  }
  g_mega_ncta = (coop && per_sm >= 1) ? nsm : 0;
  g_mega_realloc_ok = coop && per_sm_realloc >= 1;
  snprintf(g_mega_why, sizeof(g_mega_why), "sms=%d coop=%d blocks_per_sm=%d (setmaxnreg variant: %d) smem=%d threads=%d -> ncta=%d",
           nsm, coop, per_sm, per_sm_realloc, mega::SMEM_BYTES, mega::NTHREADS, g_mega_ncta);
  return cudaSuccess;
}
bool decode_mega_realloc_supported() { return g_mega_realloc_ok; }
int decode_mega_ncta() { return g_mega_ncta; }
bool decode_mega_supported(int H, int I, int head_dim, int max_batch) {
  auto okk = [](int K) { return K % 32 == 0 && (K <= mega::KS_MAX ? true : K % mega::KS_MAX == 0); };
  return mega::NWC == 8 && g_mega_ncta > 0 && head_dim == mega::D && okk(H) && okk(I) && H <= 2 * mega::KS_MAX && max_batch <= 8;
}

cudaError_t launch_decode_mega(const MegaLaunch& m, cudaStream_t st) {
  mega::Args a{};
  a.layers = reinterpret_cast<const mega::Layer*>(m.layers_dev);
  a.n_layer = m.n_layer; a.B = m.B; a.H = m.H; a.I = m.I; a.n_head = m.n_head; a.n_kv = m.n_kv; a.qkv_cols = m.qkv_cols;
  a.vocab = m.vocab; a.tcap = m.tcap; a.n_positions = m.n_positions; a.ln_eps = m.ln_eps;
  a.wte = m.wte; a.wpe = m.wpe; a.lnf_w = m.lnf_w; a.lnf_b = m.lnf_b; a.lm_head = m.lm_head;
  a.x = m.x; a.qkv = m.qkv; a.attn = m.attn; a.h = m.h; a.logits = m.logits;
  a.attn_partial = m.attn_partial; a.amax_val = m.amax_val; a.amax_idx = m.amax_idx;
  a.state = m.state; a.params = m.params; a.seen = m.seen; a.next_ids = m.next_ids; a.out_ids = m.out_ids;
  a.barrier_ctr = m.barrier_ctr; a.nsteps = m.nsteps; a.att_ncta = m.att_ncta; a.dbg = m.dbg;
  cudaError_t e = cudaMemsetAsync(m.barrier_ctr, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  void* args[] = {&a};
  if (m.realloc)
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(mega::decode_mega_kernel<true>), dim3(g_mega_ncta),
                                    dim3(mega::NTHREADS_REALLOC), args, mega::SMEM_BYTES, st);
  else
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(mega::decode_mega_kernel<false>), dim3(g_mega_ncta),
                                    dim3(mega::NTHREADS), args, mega::SMEM_BYTES, st);
  count_launch();
  return e;
}


// ---- per-phase ring GEMV launchers (used by the CUDA-graph decode path)
template <bool HAS_LN, int EPI, bool LN_BIGK = false>
static void launch_ring_t(const mega::RingGemvArgs& ra, int ncta, bool pdl, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(mega::NTHREADS); cfg.stream = st;
  cfg.dynamicSmemBytes = mega::ring_smem_bytes(ra.nslots);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, mega::gemv_ring_kernel<HAS_LN, EPI, LN_BIGK>, ra);
  count_launch();
}

cudaError_t gemv_ring_init() {   // set the shared-memory opt-in outside of any stream capture
  cudaError_t e;
#define SV_RING_ATTR(LN, EPI)                                                                                          \
  e = cudaFuncSetAttribute(mega::gemv_ring_kernel<LN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, mega::SMEM_BYTES); \
  if (e != cudaSuccess) return e;
  SV_RING_ATTR(true, mega::EPI_QKV) SV_RING_ATTR(true, mega::EPI_PLAIN) SV_RING_ATTR(true, mega::EPI_LMHEAD)
  SV_RING_ATTR(false, mega::EPI_PLAIN)
#undef SV_RING_ATTR
  e = cudaFuncSetAttribute(mega::gemv_ring_kernel<true, mega::EPI_PLAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mega::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(mega::gemv_ring_kernel<true, mega::EPI_LMHEAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mega::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaSuccess;
}

bool gemv_ring_supported(int K, bool has_ln) { (void)has_ln; return K >= 32 && K % 32 == 0; }

// CTAs of one ring GEMV: one per SM in every build.  With RING_MINBLOCKS = 2 the CTA is small enough for two per SM, and
// the second slot is deliberately left free: it is where the NEXT kernel's CTA (launched early through PDL) becomes
// resident and starts filling its ring while this kernel is still computing.
static int ring_ncta() {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  return nsm;
}

int gemv_ring_ntiles(int N) {
  const int nsm = ring_ncta();
  const int rows_per_cta = (N + nsm - 1) / nsm, tpc = (rows_per_cta + 15) / 16, R = (rows_per_cta + tpc - 1) / tpc;
  return (N + R - 1) / R;
}

void launch_gemv_ring(const RingGemvLaunch& g, cudaStream_t st) {
  mega::RingGemvArgs ra{};
  ra.a.B = g.B; ra.a.ln_eps = g.ln_eps; ra.a.n_head = g.n_head; ra.a.n_kv = g.n_kv; ra.a.tcap = g.tcap; ra.a.state = const_cast<GenState*>(g.state);
  ra.a.amax_val = g.amax_val; ra.a.amax_idx = g.amax_idx;
  ra.L.kc = g.kcache; ra.L.vc = g.vtcache;
  ra.X = g.X; ra.W = g.W; ra.bias = g.bias; ra.res = g.res; ra.ln_w = g.ln_w; ra.ln_b = g.ln_b; ra.Y = g.Y;
  ra.N = g.N; ra.K = g.K; ra.act = g.act;
  ra.next_w = g.next_w; ra.next_bytes = g.next_bytes;
  const int nsm = ring_ncta();
  {   // ring depth: what this CTA will stream, capped so the next kernel's CTA can co-reside (227 KB per SM)
    static int cap = 0;
    if (cap == 0) { const char* c = getenv("SV_RING_SLOTS"); cap = c ? atoi(c) : mega::STAGES; if (cap < 1 || cap > 6) cap = mega::STAGES; }
    const int rows_per_cta = (g.N + nsm - 1) / nsm, tpc = (rows_per_cta + 15) / 16;
    int ks = 32;
    for (int c : {1024, 768, 512, 256, 128, 64}) if (c <= g.K && g.K % c == 0) { ks = c; break; }
    const int need = tpc * (g.K / ks);
    ra.nslots = need < cap ? need : cap;
    if (ra.nslots < 1) ra.nslots = 1;
  }
  const bool ln = g.ln_w != nullptr;
  if (ln && g.K > 2 * mega::KS_MAX) {       // LayerNorm over K > 2048 (v2): separate instantiations
    if (g.epi == mega::EPI_LMHEAD) launch_ring_t<true, mega::EPI_LMHEAD, true>(ra, nsm, g.pdl, st);
    else launch_ring_t<true, mega::EPI_PLAIN, true>(ra, nsm, g.pdl, st);
    return;
  }
  if (ln && g.epi == mega::EPI_QKV) launch_ring_t<true, mega::EPI_QKV>(ra, nsm, g.pdl, st);
  else if (ln && g.epi == mega::EPI_LMHEAD) launch_ring_t<true, mega::EPI_LMHEAD>(ra, nsm, g.pdl, st);
  else if (ln) launch_ring_t<true, mega::EPI_PLAIN>(ra, nsm, g.pdl, st);
  else launch_ring_t<false, mega::EPI_PLAIN>(ra, nsm, g.pdl, st);
}

}  // namespace sv
