// Dataflow decode kernel: `nsteps` whole tokens (all layers, lm_head, greedy token selection) in ONE cooperative launch,
// one CTA per SM, and NO grid barriers and NO fences on the token's critical path.
//
// The decode step of the reference (GPTBigCodeBlock, starvector/model/gpt_bigcode/modeling_gpt_bigcode.py:670-755; loop:
// HF GenerationMixin._sample) is a chain of ~146 small all-to-all dependent phases that streams 2.24 GB
// of weights.  As separate kernels every phase pays a launch boundary (round 1: 122 launches, 1.0 ms/token = 0.34 of the HBM
// roofline); behind grid barriers it pays red.release + polling + membar + re-load (1.4 ms/token).  Here:
//
//   * weights: a producer warp per CTA walks the STATIC weight schedule of the whole launch and keeps a 5-slot shared
//     memory ring full with cp.async.bulk (sv_ring.cuh), never waiting for activations: HBM streams straight through
//     phase boundaries;
//   * activations travel between CTAs as FLAGGED WORDS through L2 (the point of coherence): a 32-bit word holds one bf16
//     value + a 16-bit phase tag (fp32 payloads: 64-bit word, 32-bit tag).  A single aligned 4/8-byte store is
//     single-copy atomic, so the consumer polls the data itself with ld.relaxed.gpu and needs neither a flag, a fence nor a
//     barrier: one phase hop costs one L2 store + one L2 load (~0.3 us) instead of ~2 us.  Tags come from a monotonic
//     phase counter, buffers are cleared when a new sequence starts, so a stale word can never carry the expected tag;
//   * a buffer is only rewritten one full all-to-all phase after its last read (DESIGN.md "flow hazards"), so no
//     double buffering and no back-pressure signalling is needed;
//   * attention: split-KV items of 8 warps x 32 keys spread over CTAs, CTA-local tree merge, partials (m,l,acc) as flagged
//     fp32 words, then a distributed merge (one warp per (image, head, 32 dims)) -- two short hops instead of a
//     cluster barrier; the current token's k/v never round-trip through the cache before they are used: the CTA that owns
//     the last key block takes them from the flagged QKV vector and appends them to the cache itself;
//   * the KV cache of older tokens is read with ld.global.cg one full token after it was written with st.global.cg.
//
// Every wait is bounded and traps instead of hanging the GPU.
#include <cstdio>
#include <cstdlib>

#include "sv_kernels.h"
#include "sv_ring.cuh"
#include "sv_select.cuh"

namespace sv {
namespace flow {

using namespace mega;

constexpr int MAXS = 64;                      // attention key splits per (image, kv head)
constexpr int FLOW_OFF_STAT = OFF_STAT;       // [2][NWC][8] floats: needs 2x the mega layout's room
constexpr int FLOW_OFF_BAR = FLOW_OFF_STAT + 2 * NWC * 8 * 4;
constexpr int FLOW_OFF_TOK = FLOW_OFF_BAR + 2 * STAGES * 8 + 4 * 8;      // + 2 full / 2 empty barriers of the LayerNorm ring
constexpr int LN_MAX_H = 2 * KS_MAX;          // LayerNorm width the parameter ring holds (the flow kernel needs H <= 2048)
constexpr int LN_SLOT_BYTES = 2 * LN_MAX_H * 2;                          // weight row + bias row, bf16
constexpr int FLOW_OFF_STATE = FLOW_OFF_TOK + 64;                        // CTA 0: GenState + GenParamsDev working copies
constexpr int FLOW_OFF_LN = (FLOW_OFF_STATE + (int)sizeof(GenState) + (int)sizeof(GenParamsDev) + 127) & ~127;   // [2][LN_SLOT_BYTES]
constexpr int FLOW_SMEM_BYTES = FLOW_OFF_LN + 2 * LN_SLOT_BYTES + 128;
static_assert(FLOW_SMEM_BYTES <= 232448, "dataflow decode kernel: shared memory over the 227 KB per-CTA limit");

struct FlowArgs {
  const Layer* layers;
  int n_layer, B, H, I, n_head, n_kv, qkv_cols, vocab, tcap, n_positions;
  float ln_eps;
  const bf16 *wte, *wpe, *lnf_w, *lnf_b, *lm_head;
  bf16* x_plain;                 // [B][H] bf16: input of the first step when first_plain; refreshed by every select
  bf16* logits;                  // [B][vocab] bf16 (plain stores; read by the host path / the penalised scan)
  uint32_t *xa, *xb, *qkv, *att, *hb;          // flagged bf16 words: [B][H], [B][H], [B][qkv_cols], [B][H], [B][I]
  unsigned long long *part, *amax;             // flagged fp32 words [B*n_kv][MAXS][PSZ]; argmax partials [ntiles][8]
  GenState* state;
  const GenParamsDev* params;
  uint8_t* seen;
  int32_t *next_ids, *out_ids;
  int nsteps, step0, cur_len0, first_plain, do_select;
  int l2_ahead;                  // weight slabs the producer asks L2 to fetch ahead of the shared-memory ring (0 = off)
  long long* dbg;                // optional: CTA 0 / thread 0 clock64() stamps of the first step
};

// ---- relaxed gpu-scope accesses (always served by L2)
SV_DEVINL uint4 ld_rlx16(const void* p) {
  uint4 r;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
SV_DEVINL uint32_t ld_rlx32(const void* p) {
  uint32_t r;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
SV_DEVINL unsigned long long ld_rlx64(const void* p) {
  unsigned long long r;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}
SV_DEVINL void st_rlx32(void* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
SV_DEVINL void st_rlx64(void* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
SV_DEVINL void st_rlx16B(void* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
SV_DEVINL void spin_guard(uint32_t& it) { if (++it > (1u << 26)) __trap(); }

// phase tags
SV_DEVINL uint32_t tag16(uint32_t gp) { return ((gp % 65535u) + 1u) << 16; }    // in the upper half of a bf16 word
SV_DEVINL unsigned long long tag32(uint32_t gp) { return (unsigned long long)(gp + 1u) << 32; }
SV_DEVINL unsigned long long fword(float v, unsigned long long T) { return T | (unsigned long long)__float_as_uint(v); }

// 8 consecutive bf16 values out of flagged words; nonzero result = at least one word does not carry tag E yet
SV_DEVINL uint32_t ll_get8(const uint32_t* p, uint32_t E, uint4& out) {
  const uint4 a = ld_rlx16(p), b = ld_rlx16(p + 4);
  out.x = __byte_perm(a.x, a.y, 0x5410); out.y = __byte_perm(a.z, a.w, 0x5410);
  out.z = __byte_perm(b.x, b.y, 0x5410); out.w = __byte_perm(b.z, b.w, 0x5410);
  return ((a.x ^ E) | (a.y ^ E) | (a.z ^ E) | (a.w ^ E) | (b.x ^ E) | (b.y ^ E) | (b.z ^ E) | (b.w ^ E)) >> 16;
}
SV_DEVINL void ll_put8(uint32_t* p, uint32_t E, const uint4& v) {
  st_rlx16B(p, make_uint4(E | (v.x & 0xffffu), E | (v.x >> 16), E | (v.y & 0xffffu), E | (v.y >> 16)));
  st_rlx16B(p + 4, make_uint4(E | (v.z & 0xffffu), E | (v.z >> 16), E | (v.w & 0xffffu), E | (v.w >> 16)));
}

struct FCtx {
  const FlowArgs* a;
  uint8_t* smem;
  int cta, ncta, warp, lane, g, t;
  float* red;     // [2][NWC][16][8]
  float* stat;    // [2][NWC][8]
  long long* dbg; // nullptr unless this thread records the timeline
  int dbg_i;
  bool slow_select;   // repetition penalty armed: the select phase scans the full logits row
};
// timeline records (SV_MEGA_DEBUG): [id << 48 | clock64], CTA 0 only; consumer thread 0 fills dbg[0..4096), the producer
// warp's lane 0 dbg[4096..8192).  ids: 8 * kind + {1 enter, 2 x ready, 3 LayerNorm done, 4 first weight slab landed, 5 last
// slab consumed, 6 outputs stored} with kind 0 qkv, 1 c_proj, 2 fc, 3 mlp.c_proj, 4 lm_head; 40.. attention; 64 + 2 * kind
// (+1) = producer starts (has issued) the kind's slabs.
enum { ST_ENTER = 1, ST_XREADY = 2, ST_LN = 3, ST_W0 = 4, ST_WLAST = 5, ST_DONE = 6, ST_ATT_ENTER = 40, ST_ATT_Q = 41, ST_ATT_BLK = 42,
       ST_ATT_TREE = 43, ST_ATT_DONE = 44, ST_MERGE_DONE = 46, ST_SELECT_DONE = 47, ST_PROD = 64 };
constexpr int DBG_HALF = 4096;
SV_DEVINL void stamp_raw(long long* dbg, int& i, int id) {
  if (dbg && i < DBG_HALF) dbg[i++] = (long long)(((unsigned long long)id << 48) | ((unsigned long long)clock64() & 0xffffffffffffull));
}
SV_DEVINL void stamp(FCtx& cx, int id) { stamp_raw(cx.dbg, cx.dbg_i, id); }

enum { EPI_LL = 0, EPI_LMHEAD = 2 };

// LayerNorm parameters ride the producer's schedule too: (weight, bias) rows land in a 2-slot shared-memory mini-ring one or
// two phases before the consumers need them, instead of 16 dependent trips to HBM in the LayerNorm prologue.
struct LnRing {
  uint32_t base, full0, empty0, slot, phase;
  SV_DEVINL void advance() { if (++slot == 2u) { slot = 0; phase ^= 1u; } }
};
SV_DEVINL void produce_ln(LnRing& lr, const bf16* ln_w, const bf16* ln_b, int N, int K, int cta, int ncta, int lane) {
  if (make_plan(N, K, cta, ncta).ntile <= 0) return;          // the consumers skip the phase as well
  if (lane == 0) {
    const uint32_t fb = lr.full0 + 8u * lr.slot, dst = lr.base + lr.slot * LN_SLOT_BYTES;
    mbar_wait(lr.empty0 + 8u * lr.slot, lr.phase ^ 1u);
    mbar_expect_tx(fb, (uint32_t)(4 * K));
    bulk_g2s(dst, ln_w, (uint32_t)(2 * K), fb);
    bulk_g2s(dst + LN_MAX_H * 2, ln_b, (uint32_t)(2 * K), fb);
  }
  lr.advance();
}

// ---- consumer: one GEMV phase  Y[B,N] = epi( LN?(X)[B,K] . W[N,K]^T ) on flagged activations.
// X: flagged [B][K] carrying tag EX.  res (optional): flagged [B][N], tag ER.  EPI_LL: Y flagged [B][N], tag EY.
// EPI_LMHEAD: plain bf16 logits + one flagged argmax partial per (tile, image).
// has_ln / epi are run-time (warp-uniform) switches on purpose: the kernel holds ONE copy of this code for its five call
// patterns (a 256 KB kernel thrashed the instruction cache at every phase change).
SV_DEVINL void gemv_flow(FCtx& cx, Ring& r, LnRing& lr, const bool has_ln, const int epi, const uint32_t* __restrict__ X, uint32_t EX,
                         const bf16* __restrict__ bias, const uint32_t* res, uint32_t ER, uint32_t* Y, uint32_t EY, int N, int K, int act,
                         uint32_t gp, int kind) {
  const FlowArgs& a = *cx.a;
  const Plan p = make_plan(N, K, cx.cta, cx.ncta);
  if (p.ntile <= 0) return;
  stamp(cx, 8 * kind + ST_ENTER);                          // nothing to do here: go and wait where this CTA has work
  const int warp = cx.warp, g = cx.g, t = cx.t;
  const int cps = p.KS >> 5;                         // 32-wide chunks per slot row
  const int cpws = (cps + NWC - 1) / NWC;            // chunks per warp per slot (<= CPW)
  const bool row_ok = g < a.B;
  const uint32_t* xp = X + (int64_t)(row_ok ? g : 0) * K + 8 * t;
  const bool big_k = p.nstg > 2;

  // activations of the whole phase live in registers when K <= 2048 (8 fragments per lane)
  uint4 xr[2 * CPW];
#pragma unroll
  for (int i = 0; i < 2 * CPW; ++i) xr[i] = make_uint4(0u, 0u, 0u, 0u);
  if (!big_k) {
    if (row_ok) {
      uint32_t bad, it = 0;
      do {
        bad = 0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
          for (int j = 0; j < CPW; ++j) {
            const int cl = warp + NWC * j;
            if (ks < p.nstg && j < cpws && cl < cps) bad |= ll_get8(xp + (ks * cps + cl) * 32, EX, xr[ks * CPW + j]);
          }
        }
        if (bad) spin_guard(it);
      } while (bad);
    }
    stamp(cx, 8 * kind + ST_XREADY);
    if (has_ln) {
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
      if (t == 0) cx.stat[NWC * 8 + warp * 8 + g] = q;       // second half of stat[]: no write-after-read barrier needed
      consumer_sync();
      float var = 0.f;
#pragma unroll
      for (int w = 0; w < NWC; ++w) var += cx.stat[NWC * 8 + w * 8 + g];
      const float rstd = 1.0f / sqrtf(var / (float)K + a.ln_eps);
      mbar_wait(lr.full0 + 8u * lr.slot, lr.phase);          // staged by the producer warp, normally long ago
      const uint32_t lnw = lr.base + lr.slot * LN_SLOT_BYTES + 16 * t, lnb = lnw + LN_MAX_H * 2;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cl = warp + NWC * j;
          const bool okc = ks < p.nstg && j < cpws && cl < cps;
          float f[8], wf[8], bfv[8];
          unpack8(xr[ks * CPW + j], f);
          const int ch = okc ? ks * cps + cl : 0;
          unpack8(lds16(lnw + ch * 64), wf);
          unpack8(lds16(lnb + ch * 64), bfv);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = (row_ok && okc) ? (f[e] - mean) * rstd * wf[e] + bfv[e] : 0.f;
          xr[ks * CPW + j] = pack8(f);       // ln output is a bf16 tensor in the reference; 0 on padded chunks
        }
      }
      __syncwarp();
      if (cx.lane == 0) mbar_arrive(lr.empty0 + 8u * lr.slot);
      lr.advance();
      stamp(cx, 8 * kind + ST_LN);
    }
  }

  float c[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tl = 0; tl < p.ntile; ++tl) {
    const int tile = p.tile0 + tl;
    // the epilogue thread's residual word: requested now, checked after the MMAs (it was written a phase ago)
    const int en = threadIdx.x & 15, emm = threadIdx.x >> 4;
    const int ecol = tile * p.R + en;
    const bool eok = threadIdx.x < 128 && en < p.R && ecol < N && emm < a.B;
    uint32_t rword = 0;
    if (res != nullptr && eok) rword = ld_rlx32(res + (int64_t)emm * N + ecol);
    float bias_v = 0.f;                                  // fetched now: an HBM miss here must not sit behind the last MMA
    if (bias != nullptr && eok) bias_v = __bfloat162float(bias[ecol]);
    if (!big_k) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        if (ks < p.nstg) {
          mbar_wait(r.full0 + 8u * r.slot, r.phase);
          if (tl == 0 && ks == 0) stamp(cx, 8 * kind + ST_W0);
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
      // K > 2048: activation fragments are fetched per slab from L2, one slab ahead of their use.  The loads of slab
      // ks + 1 are ISSUED before the MMAs of slab ks and only CHECKED after them, so the L2 round trip hides behind the
      // weight wait + MMAs (checking at once cost one exposed round trip per slab: 8 per mlp.c_proj tile).
      uint4 xc[CPW], xn[CPW], ra[CPW], rb[CPW];
      auto issue = [&](int ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cl = warp + NWC * j;
          if (row_ok && ks < p.nstg && j < cpws && cl < cps) {
            const uint32_t* q = xp + (ks * cps + cl) * 32;
            ra[j] = ld_rlx16(q); rb[j] = ld_rlx16(q + 4);
          }
        }
      };
      auto finish = [&](int ks) {                          // raw words -> fragments; spins (re-loading) until every tag matches
#pragma unroll
        for (int j = 0; j < CPW; ++j) xn[j] = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok && ks < p.nstg) {
          uint32_t it = 0;
          for (;;) {
            uint32_t bad = 0;
#pragma unroll
            for (int j = 0; j < CPW; ++j) {
              const int cl = warp + NWC * j;
              if (j < cpws && cl < cps) {
                bad |= ((ra[j].x ^ EX) | (ra[j].y ^ EX) | (ra[j].z ^ EX) | (ra[j].w ^ EX) | (rb[j].x ^ EX) | (rb[j].y ^ EX) | (rb[j].z ^ EX) |
                        (rb[j].w ^ EX)) >> 16;
                xn[j].x = __byte_perm(ra[j].x, ra[j].y, 0x5410); xn[j].y = __byte_perm(ra[j].z, ra[j].w, 0x5410);
                xn[j].z = __byte_perm(rb[j].x, rb[j].y, 0x5410); xn[j].w = __byte_perm(rb[j].z, rb[j].w, 0x5410);
              }
            }
            if (!bad) break;
            spin_guard(it);
            issue(ks);
          }
        }
      };
      issue(0);
      finish(0);
      if (tl == 0) stamp(cx, 8 * kind + ST_XREADY);
      for (int ks = 0; ks < p.nstg; ++ks) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) xc[j] = xn[j];
        issue(ks + 1);
        mbar_wait(r.full0 + 8u * r.slot, r.phase);
        if (tl == 0 && ks == 0) stamp(cx, 8 * kind + ST_W0);
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
        finish(ks + 1);
      }
    }
    // ---- tile finished: deterministic cross-warp split-K reduction + epilogue
    if (tl == p.ntile - 1) stamp(cx, 8 * kind + ST_WLAST);
    float* rd = cx.red + (tl & 1) * (NWC * 16 * 8);
    rd[(warp * 16 + g) * 8 + 2 * t] = c[0]; rd[(warp * 16 + g) * 8 + 2 * t + 1] = c[1];
    rd[(warp * 16 + g + 8) * 8 + 2 * t] = c[2]; rd[(warp * 16 + g + 8) * 8 + 2 * t + 1] = c[3];
    c[0] = c[1] = c[2] = c[3] = 0.f;
    consumer_sync();
    if (threadIdx.x < 128) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < NWC; ++w) acc += rd[(w * 16 + en) * 8 + emm];
      float v = 0.f;
      if (eok) {
        const float bv = bias_v;
        float rv = 0.f;
        if (res != nullptr) {
          uint32_t it = 0;
          while ((rword ^ ER) >> 16) { spin_guard(it); rword = ld_rlx32(res + (int64_t)emm * N + ecol); }
          rv = __uint_as_float(rword << 16);
        }
        v = epilogue_elem(acc, bv, act, res != nullptr, rv);
        const bf16 vb = __float2bfloat16_rn(v);
        if (epi == EPI_LL) {
          st_rlx32(Y + (int64_t)emm * N + ecol, EY | (uint32_t)__bfloat16_as_ushort(vb));
        } else {
          a.logits[(int64_t)emm * N + ecol] = vb;
        }
      }
      if (epi == EPI_LMHEAD) {
        // penalised selection reads the logits themselves: order them before the partial word that announces the tile
        // (all 16 stores a partial covers come from this half-warp)
        if (cx.slow_select) { __threadfence(); __syncwarp(); }
        // greedy = argmax over the bf16 logits cast to float, lowest index wins ties (HF _sample)
        float bv = eok ? v : -INFINITY;
        int bi = eok ? ecol : 0x7fffffff;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (en == 0 && emm < a.B) {
          // [tag16 | bf16 bits of the value | index]: one atomic 8-byte word
          const unsigned long long w = ((unsigned long long)(tag16(gp) | (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(bv))) << 32) |
                                       (unsigned long long)(uint32_t)bi;
          st_rlx64(a.amax + (int64_t)tile * 8 + emm, w);
        }
      }
    }
    // red[] is double-buffered by tile parity: one barrier per tile
  }
  stamp(cx, 8 * kind + ST_DONE);
}

// ---- attention, hop 1.  item = (image, kv head, key range c) = one CTA.
// Two sub-phases with different work splits, so that no (m, l, acc) partial ever has to be merged across warps:
//   1. keys split over warps: warp w computes S = q.K^T for key blocks w (and w + 8) of the item on the tensor cores
//      (the <= 16 query heads of the kv head are the MMA M dimension), the row maxima go through shared memory, and with the
//      item-wide maximum every warp turns its scores into P = exp2(S - M) once; P is parked in shared memory in exactly the
//      register layout the P.V MMA wants as its A operand (so there is no transpose);
//   2. output dims split over warps: warp w computes out[:, 16w .. 16w+15] = P . V over ALL keys of the item: it owns a
//      disjoint slice of the output, its V^T rows are 16-byte coalesced loads that are issued before the softmax finishes.
// One item covers <= 16 key blocks (512 keys).  Up to 512 keys of context a single item holds the whole row: it
// normalises and writes the attention output directly (no merge hop).  Longer rows are cut into items of <= 8 blocks
// whose (m, l, acc) partials go to merge_flow as flagged fp32 words.
constexpr int ATT_R = 2;                                  // key blocks per warp and item
constexpr int ATT_BLKS = NWC * ATT_R;                     // key blocks per item
constexpr int ATT_P_BYTES = ATT_BLKS * 2 * 32 * 16;       // P fragments: [block][h][lane] x 16 bytes
constexpr int OFF_ATT_M = OFF_ATT + ATT_P_BYTES;          // float [NWC][16] row maxima, then [NWC][16] row sums
static_assert(ATT_P_BYTES + 2 * NWC * 16 * 4 <= ATT_BYTES, "attention scratch must fit the mega layout's tree-merge buffer");

SV_DEVINL void attn_split(int nkeys, int& nact, int& per) {
  const int nblk = (nkeys + 31) / 32;
  if (nblk <= ATT_BLKS) { nact = 1; per = nblk; return; }
  nact = min(MAXS, (nblk + NWC - 1) / NWC);
  per = (nblk + nact - 1) / nact;
  nact = (nblk + per - 1) / per;
}

// S (log2 domain, scaled, masked) of one 32-key block: thread (g, t) gets keys kb + 8t + 2j + e for head rows g (s[j][e])
// and g + 8 (s[j][2 + e]) -- the layout the P.V A operand needs (sv_attention.cu, fragment trick).
SV_DEVINL void qk_block(const uint32_t (&qa)[D / 16][4], const bf16* __restrict__ kbase, int kb, int key_end, float scale_log2,
                        float (&s)[4][4], int g, int t) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    int key = kb + 8 * (g >> 1) + 2 * j + (g & 1);
    key = key < key_end ? key : key_end - 1;
    const bf16* kp = kbase + (int64_t)key * D + 8 * t;
#pragma unroll
    for (int jj = 0; jj < D / 32; ++jj) {
      const uint4 w = ldcg16(kp + 32 * jj);
      mma_bf16_16816(s[j], qa[2 * jj][0], qa[2 * jj][1], qa[2 * jj][2], qa[2 * jj][3], w.x, w.y);
      mma_bf16_16816(s[j], qa[2 * jj + 1][0], qa[2 * jj + 1][1], qa[2 * jj + 1][2], qa[2 * jj + 1][3], w.z, w.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool valid = (kb + 8 * t + 2 * j + e) < key_end;
      s[j][e] = valid ? s[j][e] * scale_log2 : -INFINITY;
      s[j][2 + e] = valid ? s[j][2 + e] * scale_log2 : -INFINITY;
    }
  }
}

SV_DEVINL void attention_flow(FCtx& cx, const Layer* L, int cur_len, uint32_t E, uint32_t gp) {
  const FlowArgs& a = *cx.a;
  const int warp = cx.warp, lane = cx.lane, g = cx.g, t = cx.t;
  const int group = a.n_head / a.n_kv;
  const int nkeys = cur_len + 1;
  const int nblk = (nkeys + 31) / 32;
  int nact, per;
  attn_split(nkeys, nact, per);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  const uint32_t pbuf = smem_u32(cx.smem + OFF_ATT);
  float* mbuf = reinterpret_cast<float*>(cx.smem + OFF_ATT_M);
  float* lbuf = mbuf + NWC * 16;
  const unsigned long long T = tag32(gp);
  const int nitems = a.B * a.n_kv * nact;
  stamp(cx, ST_ATT_ENTER);
  for (int item = cx.cta; item < nitems; item += cx.ncta) {
    const int c = item % nact, bk = item / nact, kvh = bk % a.n_kv, b = bk / a.n_kv;
    const int blk0 = c * per, blk1 = min(nblk, blk0 + per), nb = blk1 - blk0;
    bf16* kb_ = L->kc + (int64_t)bk * a.tcap * D;
    bf16* vb_ = L->vc + (int64_t)bk * D * a.tcap;
    const uint32_t* qkv_row = a.qkv + (int64_t)b * a.qkv_cols;
    // ---- sub-phase 1: scores of this warp's key blocks
    float s[ATT_R][4][4];
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (warp < nb) {
      const uint32_t* qrow = qkv_row + (int64_t)kvh * group * D;
      uint32_t qa[D / 16][4];
      {
        uint4 lo[D / 32], hi[D / 32];
#pragma unroll
        for (int jj = 0; jj < D / 32; ++jj) { lo[jj] = make_uint4(0u, 0u, 0u, 0u); hi[jj] = make_uint4(0u, 0u, 0u, 0u); }
        uint32_t bad, it = 0;
        do {
          bad = 0;
#pragma unroll
          for (int jj = 0; jj < D / 32; ++jj) {
            if (g < group) bad |= ll_get8(qrow + (int64_t)g * D + 32 * jj + 8 * t, E, lo[jj]);
            if (g + 8 < group) bad |= ll_get8(qrow + (int64_t)(g + 8) * D + 32 * jj + 8 * t, E, hi[jj]);
          }
          if (bad) spin_guard(it);
        } while (bad);
#pragma unroll
        for (int jj = 0; jj < D / 32; ++jj) {
          qa[2 * jj][0] = lo[jj].x; qa[2 * jj][1] = hi[jj].x; qa[2 * jj][2] = lo[jj].y; qa[2 * jj][3] = hi[jj].y;
          qa[2 * jj + 1][0] = lo[jj].z; qa[2 * jj + 1][1] = hi[jj].z; qa[2 * jj + 1][2] = lo[jj].w; qa[2 * jj + 1][3] = hi[jj].w;
        }
      }
      stamp(cx, ST_ATT_Q);
      // the warp that reads the newest key block appends the current token's k / v to the cache first
      // (GPTBigCodeAttention.forward: key_value = cat(layer_past, key_value), vendored modeling_gpt_bigcode.py:265-267);
      // the other warps read that V^T column only after the CTA barrier below
      if (blk1 == nblk && ((nb - 1) % NWC) == warp && cur_len < a.tcap) {
        const uint32_t* kll = qkv_row + (int64_t)a.n_head * D + (int64_t)kvh * D + 4 * lane;
        const uint32_t* vll = kll + (int64_t)a.n_kv * D;
        uint4 kw, vw;
        uint32_t bad, it = 0;
        do {
          kw = ld_rlx16(kll); vw = ld_rlx16(vll);
          bad = ((kw.x ^ E) | (kw.y ^ E) | (kw.z ^ E) | (kw.w ^ E) | (vw.x ^ E) | (vw.y ^ E) | (vw.z ^ E) | (vw.w ^ E)) >> 16;
          if (bad) spin_guard(it);
        } while (bad);
        uint2 kp;
        kp.x = __byte_perm(kw.x, kw.y, 0x5410); kp.y = __byte_perm(kw.z, kw.w, 0x5410);
        __stcg(reinterpret_cast<uint2*>(kb_ + (int64_t)cur_len * D + 4 * lane), kp);
        unsigned short* vt = reinterpret_cast<unsigned short*>(vb_) + (int64_t)(4 * lane) * a.tcap + cur_len;
        __stcg(vt, (unsigned short)(vw.x & 0xffffu));
        __stcg(vt + a.tcap, (unsigned short)(vw.y & 0xffffu));
        __stcg(vt + 2 * (int64_t)a.tcap, (unsigned short)(vw.z & 0xffffu));
        __stcg(vt + 3 * (int64_t)a.tcap, (unsigned short)(vw.w & 0xffffu));
        __threadfence_block();
        __syncwarp();
      }
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        const int bi = warp + r * NWC;
        if (bi < nb) {
          const int kb = (blk0 + bi) * 32;
          qk_block(qa, kb_, kb, min(nkeys, kb + 32), scale_log2, s[r], g, t);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mx0 = fmaxf(mx0, fmaxf(s[r][j][0], s[r][j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[r][j][2], s[r][j][3]));
          }
        }
      }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
    }
    if (t == 0) { mbuf[warp * 16 + g] = mx0; mbuf[warp * 16 + g + 8] = mx1; }
    consumer_sync();
    stamp(cx, ST_ATT_BLK);
    // ---- item-wide row maxima, P = exp2(S - M) parked as MMA A fragments, row sums
    float M0 = -INFINITY, M1 = -INFINITY;
#pragma unroll
    for (int w = 0; w < NWC; ++w) { M0 = fmaxf(M0, mbuf[w * 16 + g]); M1 = fmaxf(M1, mbuf[w * 16 + g + 8]); }
    float rs0 = 0.f, rs1 = 0.f;
    if (warp < nb) {
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        const int bi = warp + r * NWC;
        if (bi < nb) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s[r][j][0] = exp2f(s[r][j][0] - M0); s[r][j][1] = exp2f(s[r][j][1] - M0);
            s[r][j][2] = exp2f(s[r][j][2] - M1); s[r][j][3] = exp2f(s[r][j][3] - M1);
            rs0 += s[r][j][0] + s[r][j][1]; rs1 += s[r][j][2] + s[r][j][3];
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 pa;
            pa.x = pack_bf16x2(s[r][2 * h][0], s[r][2 * h][1]);
            pa.y = pack_bf16x2(s[r][2 * h][2], s[r][2 * h][3]);
            pa.z = pack_bf16x2(s[r][2 * h + 1][0], s[r][2 * h + 1][1]);
            pa.w = pack_bf16x2(s[r][2 * h + 1][2], s[r][2 * h + 1][3]);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(pbuf + ((bi * 2 + h) * 32 + lane) * 16), "r"(pa.x), "r"(pa.y),
                         "r"(pa.z), "r"(pa.w) : "memory");
          }
        }
      }
    }
    rs0 = quad_sum(rs0); rs1 = quad_sum(rs1);
    if (t == 0) { lbuf[warp * 16 + g] = rs0; lbuf[warp * 16 + g + 8] = rs1; }
    consumer_sync();
    // ---- sub-phase 2: out[:, 16 * warp + {0..15}] = P . V over every key block of the item
    float acc[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    const bf16* v0 = vb_ + (int64_t)(16 * warp + g) * a.tcap + blk0 * 32 + 8 * t;       // V^T row of n-tile 0; n-tile 1: + 8 rows
#pragma unroll 1
    for (int b0 = 0; b0 < nb; b0 += 4) {
      uint4 vv[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (b0 + i < nb) {
          vv[i][0] = ldcg16(v0 + (b0 + i) * 32);
          vv[i][1] = ldcg16(v0 + 8 * (int64_t)a.tcap + (b0 + i) * 32);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (b0 + i < nb) {
          const uint4 p0 = lds16(pbuf + (((b0 + i) * 2 + 0) * 32 + lane) * 16), p1 = lds16(pbuf + (((b0 + i) * 2 + 1) * 32 + lane) * 16);
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            mma_bf16_16816(acc[n], p0.x, p0.y, p0.z, p0.w, vv[i][n].x, vv[i][n].y);
            mma_bf16_16816(acc[n], p1.x, p1.y, p1.z, p1.w, vv[i][n].z, vv[i][n].w);
          }
        }
      }
    }
    float L0 = 0.f, L1 = 0.f;
#pragma unroll
    for (int w = 0; w < NWC; ++w) { L0 += lbuf[w * 16 + g]; L1 += lbuf[w * 16 + g + 8]; }
    stamp(cx, ST_ATT_TREE);
    if (nact == 1) {
      // the whole row was in this item: normalise and publish the attention output (bf16, like the reference's attn_output)
      uint32_t* orow = a.att + (int64_t)b * a.n_head * D + (int64_t)kvh * group * D;
      const float i0 = 1.0f / L0, i1 = 1.0f / L1;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int d = 16 * warp + 8 * n + 2 * t;
        if (g < group) {
          st_rlx32(orow + g * D + d, E | (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[n][0] * i0)));
          st_rlx32(orow + g * D + d + 1, E | (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[n][1] * i0)));
        }
        if (g + 8 < group) {
          st_rlx32(orow + (g + 8) * D + d, E | (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[n][2] * i1)));
          st_rlx32(orow + (g + 8) * D + d + 1, E | (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[n][3] * i1)));
        }
      }
    } else {
      unsigned long long* ws = a.part + ((int64_t)bk * MAXS + c) * PSZ;
      if (warp == 0 && t == 0) {
        st_rlx64(ws + g, fword(M0, T)); st_rlx64(ws + g + 8, fword(M1, T));
        st_rlx64(ws + 16 + g, fword(L0, T)); st_rlx64(ws + 16 + g + 8, fword(L1, T));
      }
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int d = 16 * warp + 8 * n + 2 * t;
        st_rlx64(ws + 32 + g * D + d, fword(acc[n][0], T)); st_rlx64(ws + 32 + g * D + d + 1, fword(acc[n][1], T));
        st_rlx64(ws + 32 + (g + 8) * D + d, fword(acc[n][2], T)); st_rlx64(ws + 32 + (g + 8) * D + d + 1, fword(acc[n][3], T));
      }
    }
    stamp(cx, ST_ATT_DONE);
    consumer_sync();                                      // the scratch is reused by the CTA's next item
  }
}

// ---- attention, hop 2: distributed merge of the item partials; task = (image, head, 32 output dims) = one warp
SV_DEVINL void merge_flow(FCtx& cx, int cur_len, uint32_t E, uint32_t gp) {
  const FlowArgs& a = *cx.a;
  const int group = a.n_head / a.n_kv;
  int nact, per;
  attn_split(cur_len + 1, nact, per);
  if (nact == 1) return;                               // the single item published the output itself
  const unsigned long long T = tag32(gp);
  const int ntasks = a.B * a.n_head * (D / 32);
  for (int task = cx.cta + cx.ncta * cx.warp; task < ntasks; task += cx.ncta * NWC) {
    const int q4 = task % (D / 32), head = (task / (D / 32)) % a.n_head, b = task / ((D / 32) * a.n_head);
    const int kvh = head / group, rr = head % group, dim = q4 * 32 + cx.lane;
    const unsigned long long* p0 = a.part + ((int64_t)(b * a.n_kv + kvh) * MAXS) * PSZ;
    float M = -INFINITY, Lsum = 0.f, A = 0.f;
    for (int c0 = 0; c0 < nact; c0 += 8) {
      unsigned long long wm[8], wl[8], wa[8];
      uint32_t it = 0;
      bool bad;
      do {
        bad = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c0 + j < nact) {
            const unsigned long long* pc = p0 + (int64_t)(c0 + j) * PSZ;
            wm[j] = ld_rlx64(pc + rr); wl[j] = ld_rlx64(pc + 16 + rr); wa[j] = ld_rlx64(pc + 32 + rr * D + dim);
            bad |= ((wm[j] ^ T) >> 32) != 0 || ((wl[j] ^ T) >> 32) != 0 || ((wa[j] ^ T) >> 32) != 0;
          }
        }
        if (bad) spin_guard(it);
      } while (bad);
      float Mn = M;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c0 + j < nact) Mn = fmaxf(Mn, __uint_as_float((uint32_t)wm[j]));
      const float sc0 = (M == -INFINITY) ? 0.f : exp2f(M - Mn);
      Lsum *= sc0; A *= sc0; M = Mn;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < nact) {
          const float m = __uint_as_float((uint32_t)wm[j]);
          const float sc = (m == -INFINITY) ? 0.f : exp2f(m - M);
          Lsum += __uint_as_float((uint32_t)wl[j]) * sc;
          A += __uint_as_float((uint32_t)wa[j]) * sc;
        }
      }
    }
    const bf16 o = __float2bfloat16_rn(A / Lsum);
    st_rlx32(a.att + (int64_t)b * a.n_head * D + head * D + dim, E | (uint32_t)__bfloat16_as_ushort(o));
  }
}

// ---- next step's input: x = wte[token] + wpe[position] (GPTBigCodeModel.forward), as flagged words (+ a plain copy)
SV_DEVINL void embed_flow(const FCtx& cx, const int* toks, int pos, uint32_t E) {
  const FlowArgs& a = *cx.a;
  pos = pos >= a.n_positions ? a.n_positions - 1 : pos;
  const int hv = a.H >> 3;
  for (int i = threadIdx.x; i < a.B * hv; i += NCT) {
    const int b = i / hv, col = (i % hv) * 8;
    int id = toks[b];
    id = id < 0 ? 0 : (id >= a.vocab ? a.vocab - 1 : id);
    float e[8], q[8];
    unpack8(ldg_cached(a.wte + (int64_t)id * a.H + col), e);
    if (a.wpe) {
      unpack8(ldg_cached(a.wpe + (int64_t)pos * a.H + col), q);
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] += q[j];
    }
    const uint4 v = pack8(e);
    ll_put8(a.xa + (int64_t)b * a.H + col, E, v);
    *reinterpret_cast<uint4*>(a.x_plain + (int64_t)b * a.H + col) = v;
  }
}

// ---- token selection (CTA 0): argmax partials (or penalised scan of the logits) -> HF bookkeeping -> embedding
SV_DEVINL void select_flow(FCtx& cx, int ntiles, uint32_t gp, int next_pos, uint32_t Enext) {
  const FlowArgs& a = *cx.a;
  AmaxPair* sm = reinterpret_cast<AmaxPair*>(cx.red);
  int* s_tok = reinterpret_cast<int*>(cx.smem + FLOW_OFF_TOK);
  GenState* st = reinterpret_cast<GenState*>(cx.smem + FLOW_OFF_STATE);          // working copies (decode_flow_kernel prologue):
  const GenParamsDev* prm = reinterpret_cast<const GenParamsDev*>(st + 1);         // no global round trips in the bookkeeping
  const int tid = threadIdx.x;
  const float rp = prm->rep_penalty;
  const uint32_t T16 = tag16(gp) >> 16;
  const bool done = st->done != 0;                 // only this CTA ever writes the state during the launch
  for (int b = 0; b < a.B; ++b) {
    AmaxPair best{-INFINITY, 0x7fffffff};
    // the partial words double as "this tile's logits are complete" (the penalised path fenced before writing them)
    for (int i0 = tid; i0 < ntiles; i0 += 4 * NCT) {
      unsigned long long w[4];
      uint32_t it = 0;
      bool bad;
      do {
        bad = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j * NCT;
          if (i < ntiles) { w[j] = ld_rlx64(a.amax + (int64_t)i * 8 + b); bad |= (uint32_t)(w[j] >> 48) != T16; }
        }
        if (bad) spin_guard(it);
      } while (bad);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i0 + j * NCT < ntiles) {
          const float v = __uint_as_float(((uint32_t)(w[j] >> 32) & 0xffffu) << 16);
          best = amax_better(best, AmaxPair{v, (int)(uint32_t)w[j]});
        }
      }
    }
    if (rp != 1.0f) {
      best = AmaxPair{-INFINITY, 0x7fffffff};
      consumer_sync();                                // every partial seen by some thread -> all logits are in L2
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
  if (tid == 0 && !done) {
    select_apply_tokens(s_tok, a.B, a.vocab, st, prm, a.seen, a.next_ids, a.out_ids, 1);
    a.state->cur_len = st->cur_len; a.state->step = st->step; a.state->done = st->done;      // for the host / the next launch
    for (int b = 0; b < a.B; ++b) a.state->unfinished[b] = st->unfinished[b];
  }
  consumer_sync();
  embed_flow(cx, s_tok, next_pos, Enext);           // after `done` the other CTAs keep stepping until the launch ends
}

// ---- the static schedule: phase q of a token = layer q / 4, kind q % 4 (0 c_attn, 1 attn.c_proj, 2 mlp.c_fc, 3 mlp.c_proj);
// q == 4 * n_layer is the lm_head
struct PhaseW { const bf16 *W, *ln_w, *ln_b; int N, K, kind; };
SV_DEVINL PhaseW phase_weights(const FlowArgs& a, int q) {
  PhaseW w;
  w.ln_w = nullptr; w.ln_b = nullptr;
  if (q == 4 * a.n_layer) { w.W = a.lm_head; w.N = a.vocab; w.K = a.H; w.kind = 4; w.ln_w = a.lnf_w; w.ln_b = a.lnf_b; return w; }
  const Layer* L = a.layers + (q >> 2);
  w.kind = q & 3;
  switch (w.kind) {
    case 0: w.W = L->attn_w; w.N = a.qkv_cols; w.K = a.H; w.ln_w = L->ln1_w; w.ln_b = L->ln1_b; break;
    case 1: w.W = L->proj_w; w.N = a.H; w.K = a.H; break;
    case 2: w.W = L->fc_w; w.N = a.I; w.K = a.H; w.ln_w = L->ln2_w; w.ln_b = L->ln2_b; break;
    default: w.W = L->fc2_w; w.N = a.H; w.K = a.I; break;
  }
  return w;
}

// Position of one CTA in the weight schedule of the launch: (token, phase, tile, k slab).  The producer warp keeps two
// of them: `cur` feeds the shared-memory ring, `pf` runs `l2_ahead` slabs further and only asks L2 to fetch
// (cp.async.bulk.prefetch.L2), so that HBM keeps streaming while the ring is full and the consumers sit in a
// latency-bound phase (attention, hops): the ring then refills from L2 at L2 speed.
struct WeightWalk {
  int s, q, tl, ks, cta, ncta;
  bool done;
  PhaseW w;
  Plan p;
  SV_DEVINL void load(const FlowArgs& a) {
    for (;;) {
      if (s >= a.nsteps) { done = true; return; }
      w = phase_weights(a, q);
      p = make_plan(w.N, w.K, cta, ncta);
      if (p.ntile > 0) return;
      if (++q > 4 * a.n_layer) { q = 0; ++s; }
    }
  }
  SV_DEVINL void init(const FlowArgs& a, int cta_, int ncta_) { s = 0; q = 0; tl = 0; ks = 0; cta = cta_; ncta = ncta_; done = false; load(a); }
  SV_DEVINL bool at_phase_start() const { return tl == 0 && ks == 0; }
  SV_DEVINL void next(const FlowArgs& a) {
    if (++ks < p.nstg) return;
    ks = 0;
    if (++tl < p.ntile) return;
    tl = 0;
    if (++q > 4 * a.n_layer) { q = 0; ++s; }
    load(a);
  }
  SV_DEVINL const bf16* row_ptr(int lane) const { return w.W + (int64_t)((p.tile0 + tl) * p.R + lane) * w.K + (int64_t)ks * p.KS; }
  SV_DEVINL int rows() const { return min(p.R, w.N - (p.tile0 + tl) * p.R); }
};

// The producer also asks L2 for the K / V^T blocks this CTA's attention items of layer l will read (everything but the
// current token, which arrives as flagged words): issued when the ring starts on the layer's c_attn weights, i.e. a few
// microseconds before the attention phase, so its dependent loads hit L2 instead of HBM.
SV_DEVINL void prefetch_kv_l2(const FlowArgs& a, const Layer* L, int cur_len, int cta, int ncta, int lane) {
  if (cur_len <= 0) return;
  const int nkeys = cur_len + 1, nblk = (nkeys + 31) / 32;
  int nact, per;
  attn_split(nkeys, nact, per);
  const int nitems = a.B * a.n_kv * nact;
  for (int item = cta; item < nitems; item += ncta) {
    const int c = item % nact, bk = item / nact;
    const int key0 = c * per * 32, key1 = min(cur_len, min(nblk, c * per + per) * 32);       // cached keys of the item
    if (key1 <= key0) continue;
    const char* kp = reinterpret_cast<const char*>(L->kc + ((int64_t)bk * a.tcap + key0) * D);
    const int kbytes = (key1 - key0) * D * 2, piece = ((kbytes + 31) / 32 + 15) & ~15;
    if (lane * piece < kbytes)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kp + lane * piece), "r"((uint32_t)min(piece, kbytes - lane * piece)) : "memory");
    const int vbytes = ((key1 - key0) * 2 + 15) & ~15;
#pragma unroll
    for (int r = 0; r < D / 32; ++r) {
      const char* vp = reinterpret_cast<const char*>(L->vc + ((int64_t)bk * D + lane + 32 * r) * a.tcap + key0);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(vp), "r"((uint32_t)vbytes) : "memory");
    }
  }
}

template <bool REALLOC>
__global__ void __launch_bounds__(REALLOC ? NCT + 128 : NTHREADS, 1) decode_flow_kernel(const FlowArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  Ring ring;
  ring.base = smem_u32(smem);
  ring.full0 = smem_u32(smem + FLOW_OFF_BAR);
  ring.empty0 = ring.full0 + 8u * STAGES;
  ring.slot = 0; ring.phase = 0; ring.nslots = STAGES;
  LnRing lnr;
  lnr.base = smem_u32(smem + FLOW_OFF_LN);
  lnr.full0 = ring.empty0 + 8u * STAGES;
  lnr.empty0 = lnr.full0 + 16u;
  lnr.slot = 0; lnr.phase = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(ring.full0 + 8u * s, 1); mbar_init(ring.empty0 + 8u * s, NWC); }
    for (int s = 0; s < 2; ++s) { mbar_init(lnr.full0 + 8u * s, 1); mbar_init(lnr.empty0 + 8u * s, NWC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= NWC) {
    if constexpr (REALLOC) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(56));
      if (warp > NWC) return;
    }
    // =========================== producer: the static weight schedule of the whole launch ===========================
    long long* pdbg = (a.dbg != nullptr && cta == 0 && lane == 0) ? a.dbg + DBG_HALF : nullptr;
    int pi = 0;
    WeightWalk cur, pf;
    cur.init(a, cta, ncta);
    pf.init(a, cta, ncta);
    auto prefetch_slab = [&]() {
      if (lane < pf.rows())
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf.row_ptr(lane)), "r"((uint32_t)(pf.p.KS * 2)) : "memory");
      pf.next(a);
    };
    for (int i = 0; i < a.l2_ahead && !pf.done; ++i) prefetch_slab();
    while (!cur.done) {
      if (a.l2_ahead > 0 && !pf.done) prefetch_slab();
      if (cur.at_phase_start()) {
        if (cur.s > 0) pdbg = nullptr;
        stamp_raw(pdbg, pi, ST_PROD + 2 * cur.w.kind);
        if (cur.w.ln_w != nullptr) produce_ln(lnr, cur.w.ln_w, cur.w.ln_b, cur.w.N, cur.w.K, cta, ncta, lane);
        if (cur.w.kind == 0 && a.l2_ahead > 0) prefetch_kv_l2(a, a.layers + (cur.q >> 2), a.cur_len0 + cur.s, cta, ncta, lane);
      }
      const uint32_t fb = ring.full0 + 8u * ring.slot;
      const int rows = cur.rows();
      if (lane == 0) {
        mbar_wait(ring.empty0 + 8u * ring.slot, ring.phase ^ 1u);
        mbar_expect_tx(fb, (uint32_t)(rows * cur.p.KS * 2));
      }
      __syncwarp();
      if (lane < rows) bulk_g2s(ring.base + ring.slot * SLOT_BYTES + lane * cur.p.pitch, cur.row_ptr(lane), (uint32_t)(cur.p.KS * 2), fb);
      ring.advance();
      cur.next(a);
    }
    return;   // in-flight bulk copies are all consumed (and thus complete) before the consumers exit
  }
  // =========================== consumers ===========================
  if constexpr (REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(224));
  FCtx cx;
  cx.a = &a; cx.smem = smem; cx.cta = cta; cx.ncta = ncta; cx.warp = warp; cx.lane = lane; cx.g = lane >> 2; cx.t = lane & 3;
  cx.red = reinterpret_cast<float*>(smem + OFF_RED);
  cx.stat = reinterpret_cast<float*>(smem + FLOW_OFF_STAT);
  cx.dbg = (a.dbg != nullptr && cta == 0 && threadIdx.x == 0) ? a.dbg : nullptr;
  cx.dbg_i = 0;
  cx.slow_select = a.params->rep_penalty != 1.0f;
  if (cta == 0 && a.do_select) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + FLOW_OFF_STATE);
    const uint32_t* s0 = reinterpret_cast<const uint32_t*>(a.state);
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(a.params);
    constexpr int n0 = (int)sizeof(GenState) / 4, n1 = (int)sizeof(GenParamsDev) / 4;
    for (int i = threadIdx.x; i < n0 + n1; i += NCT) dst[i] = i < n0 ? s0[i] : s1[i - n0];
    consumer_sync();
  }
  const int ntiles_lm = make_plan(a.vocab, a.H, 0, ncta).ntiles;
  const uint32_t nl1 = (uint32_t)a.n_layer + 1u;
  if (a.first_plain && cta == 0) {
    // the step-0 input was written as plain bf16 by the kernel that selected / embedded the previous token
    const uint32_t E0 = tag16((uint32_t)a.step0 * nl1);
    const int hv = a.H >> 3;
    for (int i = threadIdx.x; i < a.B * hv; i += NCT) {
      const int b = i / hv, col = (i % hv) * 8;
      ll_put8(a.xa + (int64_t)b * a.H + col, E0, __ldcg(reinterpret_cast<const uint4*>(a.x_plain + (int64_t)b * a.H + col)));
    }
  }
  for (int s = 0; s < a.nsteps; ++s) {
    const uint32_t gs = (uint32_t)(a.step0 + s);
    const int cur_len = a.cur_len0 + s;
    if (s == 1) cx.dbg = nullptr;
    for (int q = 0; q <= 4 * a.n_layer; ++q) {
      const int l = q >> 2;
      const uint32_t gp = gs * nl1 + (uint32_t)l;          // the lm_head (q = 4 * n_layer) is "layer n_layer"
      const uint32_t E = tag16(gp), En = tag16(gp + 1u);
      const PhaseW w = phase_weights(a, q);
      const Layer* L = a.layers + (l < a.n_layer ? l : 0);
      const uint32_t *X, *res = nullptr;
      uint32_t *Y = nullptr, EY = E;
      const bf16* bias = nullptr;
      int act = 0, epi = EPI_LL;
      switch (w.kind) {
        case 0: X = a.xa; bias = L->attn_b; Y = a.qkv; break;
        case 1: X = a.att; bias = L->proj_b; res = a.xa; Y = a.xb; break;
        case 2: X = a.xb; bias = L->fc_b; Y = a.hb; act = 2 /*gelu_tanh*/; break;
        case 3: X = a.hb; bias = L->fc2_b; res = a.xb; Y = a.xa; EY = En; break;
        default: X = a.xa; epi = EPI_LMHEAD; break;
      }
      gemv_flow(cx, ring, lnr, w.ln_w != nullptr, epi, X, E, bias, res, E, Y, EY, w.N, w.K, act, gp, w.kind);
      if (w.kind == 0) {
        attention_flow(cx, L, cur_len, E, gp);
        merge_flow(cx, cur_len, E, gp);
        stamp(cx, ST_MERGE_DONE);
      }
    }
    const uint32_t gp = gs * nl1 + (uint32_t)a.n_layer;
    if (a.do_select && cta == 0) {
      select_flow(cx, ntiles_lm, gp, cur_len + 1, tag16((gs + 1u) * nl1));
      stamp(cx, ST_SELECT_DONE);
    }
  }
}

}  // namespace flow

// ---- host side
static int g_flow_ncta = 0;
static bool g_flow_realloc_ok = false;
static char g_flow_why[256] = "decode_flow_init not called";
const char* decode_flow_status() { return g_flow_why; }

cudaError_t decode_flow_init() {
  cudaError_t e = cudaFuncSetAttribute(flow::decode_flow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, flow::FLOW_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  bool realloc_attr_ok = cudaFuncSetAttribute(flow::decode_flow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              flow::FLOW_SMEM_BYTES) == cudaSuccess;
  if (!realloc_attr_ok) cudaGetLastError();
  int dev = 0, nsm = 0, per_sm = 0, per_sm_realloc = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flow::decode_flow_kernel<false>, mega::NTHREADS, flow::FLOW_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  if (!realloc_attr_ok || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_realloc, flow::decode_flow_kernel<true>, mega::NCT + 128,
                                                                         flow::FLOW_SMEM_BYTES) != cudaSuccess) {
    per_sm_realloc = 0;
    cudaGetLastError();
  }
  g_flow_ncta = (coop && per_sm >= 1) ? nsm : 0;
  g_flow_realloc_ok = coop && per_sm_realloc >= 1;
  snprintf(g_flow_why, sizeof(g_flow_why), "sms=%d coop=%d blocks_per_sm=%d (setmaxnreg variant: %d) smem=%d threads=%d -> ncta=%d", nsm,
           coop, per_sm, per_sm_realloc, flow::FLOW_SMEM_BYTES, mega::NTHREADS, g_flow_ncta);
  return cudaSuccess;
}
bool decode_flow_realloc_supported() { return g_flow_realloc_ok; }
int decode_flow_ncta() { return g_flow_ncta; }
int decode_flow_max_splits() { return flow::MAXS; }
int decode_flow_partial_floats() { return mega::PSZ; }
bool decode_flow_supported(int H, int I, int head_dim, int max_batch, int window, bool rope) {
  auto okk = [](int K) { return K % 32 == 0 && (K <= mega::KS_MAX ? true : K % mega::KS_MAX == 0); };
  return mega::NWC == 8 && g_flow_ncta > 0 && head_dim == mega::D && okk(H) && okk(I) && H <= 2 * mega::KS_MAX && max_batch <= 8 &&
         window == 0 && !rope;
}

cudaError_t launch_decode_flow(const FlowLaunch& m, cudaStream_t st) {
  flow::FlowArgs a{};
  a.layers = reinterpret_cast<const mega::Layer*>(m.layers_dev);
  a.n_layer = m.n_layer; a.B = m.B; a.H = m.H; a.I = m.I; a.n_head = m.n_head; a.n_kv = m.n_kv; a.qkv_cols = m.qkv_cols;
  a.vocab = m.vocab; a.tcap = m.tcap; a.n_positions = m.n_positions; a.ln_eps = m.ln_eps;
  a.wte = m.wte; a.wpe = m.wpe; a.lnf_w = m.lnf_w; a.lnf_b = m.lnf_b; a.lm_head = m.lm_head;
  a.x_plain = m.x_plain; a.logits = m.logits;
  a.xa = m.xa; a.xb = m.xb; a.qkv = m.qkv; a.att = m.att; a.hb = m.hb; a.part = m.part; a.amax = m.amax;
  a.state = m.state; a.params = m.params; a.seen = m.seen; a.next_ids = m.next_ids; a.out_ids = m.out_ids;
  a.nsteps = m.nsteps; a.step0 = m.step0; a.cur_len0 = m.cur_len0; a.first_plain = m.first_plain; a.do_select = m.do_select;
  a.l2_ahead = m.l2_ahead;
  a.dbg = m.dbg;
  void* args[] = {&a};
  cudaError_t e;
  if (m.realloc)
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(flow::decode_flow_kernel<true>), dim3(g_flow_ncta), dim3(mega::NCT + 128), args,
                                    flow::FLOW_SMEM_BYTES, st);
  else
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(flow::decode_flow_kernel<false>), dim3(g_flow_ncta), dim3(mega::NTHREADS), args,
                                    flow::FLOW_SMEM_BYTES, st);
  count_launch();
  return e;
}

}  // namespace sv
