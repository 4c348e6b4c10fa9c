// Decode-step kernels, register-landing generation (`SV_GEMV=regs`) + the fused token-selection kernel.
//
// A decode step streams every decoder weight once (2.24 GB for StarVector-1B), so it is HBM-bound; per layer the
// step is five kernels (ln_1+c_attn+KV append | attention | c_proj+residual | ln_2+c_fc+gelu | mlp.c_proj+residual),
// then ln_f+lm_head (+ per-tile argmax partials) and one single-CTA kernel that finishes token selection, the HF
// stop/EOS bookkeeping and the next token's embedding.
//
// gemvp_kernel here is the first persistent GEMV: 128-bit weight loads land in REGISTERS (16 per lane in flight,
// 64 KB per SM), rows tiled R <= 16 per CTA so every SM streams the same bytes, LayerNorm fused as a prologue on
// register-resident activation fragments.  The default path has since moved to the shared-memory weight ring
// (sv_decode_mega.cu: gemv_ring_kernel, ~165 KB in flight per SM); this version stays as an A/B reference and for
// shapes the ring does not take.  select_fused_kernel (below) is used by every decode mode.
//
// All kernels are written for Programmatic Dependent Launch: weights (never written at run time) are prefetched
// BEFORE `griddepcontrol.wait`; everything produced by the previous kernel is read after the wait with L2-only
// loads (ld.global.cg), because a co-resident CTA of the previous kernel may have left a stale copy of an
// in-place-updated buffer in this SM's L1.
#include "sv_kernels.h"
#include "sv_select.cuh"

namespace sv {

SV_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SV_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SV_DEVINL uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

template <typename... KArgs, typename... Args>
static void launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                      Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);
  count_launch();
}

// ------------------------------------------------------------------------------------------
enum { EPI_PLAIN = 0, EPI_QKV = 1, EPI_LMHEAD = 2 };

struct GemvArgs {
  const bf16* X; const bf16* W; const bf16* bias; const bf16* res; bf16* Y;
  int B, N, K, act;
  int R, tiles_per_cta, ntiles;        // R <= 16 weight rows per tile; CTA c owns tiles [c*tpc, (c+1)*tpc)
  const bf16* ln_w; const bf16* ln_b; float ln_eps;
  // EPI_QKV: append this token's K row / V^T column to the cache at position state->cur_len
  bf16* kcache; bf16* vtcache; const GenState* state; int q_cols, n_kv, d, tcap;
  // EPI_LMHEAD: per (tile, image row) argmax partials of the bf16-rounded logits
  float* amax_val; int* amax_idx;
};

constexpr int kGemvBatch = 8;   // 32-wide K chunks per warp per load batch (x2 weight rows per lane)

// Persistent weight-streaming GEMV.  grid ~= #SMs, one CTA per SM, NW warps.
//   * the N output features are cut into tiles of R <= 16 rows so that #tiles is a whole number of
//     CTA-loads (host picks R, tiles_per_cta from N and the SM count: 2048 -> 147 x 14 rows,
//     8192 -> 147 x 4 x 14, 49156 -> 147 x 21 x 16): every SM streams the same number of bytes;
//   * warp w owns K chunks w, w+NW, ...; lane (g,t) loads 16 bytes of rows g and g+8 per chunk, so a
//     batch puts 16 x 128-bit loads per lane (64 KB per 8-warp CTA) in flight;
//   * the next batch's loads are issued right after the current batch's MMAs, BEFORE the cross-warp
//     reduction / epilogue of a finished tile, so HBM never idles inside the kernel;
//   * with HAS_LN the activation rows are loaded and layer-normalised ONCE per CTA and stay in
//     registers as ready-made MMA "B" fragments for every tile.
template <int NW, bool HAS_LN, int EPI>
__global__ void __launch_bounds__(NW * 32, 1) gemvp_kernel(const GemvArgs a) {
  __shared__ float red[2][NW][16][8];
  __shared__ float stat[NW][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nchunks = a.K >> 5;
  const int cpw = (nchunks + NW - 1) / NW;                       // chunks per warp
  const int nb = (cpw + kGemvBatch - 1) / kGemvBatch;            // load batches per tile (1 when HAS_LN)
  const int tile0 = blockIdx.x * a.tiles_per_cta;
  const int ntile = min(a.tiles_per_cta, a.ntiles - tile0);
  const int nitems = ntile * nb;
  const bool row_ok = g < a.B;
  const bf16* xp = a.X + (int64_t)(row_ok ? g : 0) * a.K + 8 * t;

  uint4 wlo[kGemvBatch], whi[kGemvBatch], xr[kGemvBatch];
  auto load_item = [&](int item) {
    const int tile = tile0 + item / nb, base = (item % nb) * kGemvBatch;
    const int r_lo = tile * a.R + g, r_hi = r_lo + 8;
    const bool ok_lo = g < a.R && r_lo < a.N, ok_hi = g + 8 < a.R && r_hi < a.N;
    const bf16* p_lo = a.W + (int64_t)r_lo * a.K + 8 * t;
    const bf16* p_hi = a.W + (int64_t)r_hi * a.K + 8 * t;
#pragma unroll
    for (int i = 0; i < kGemvBatch; ++i) {
      const int ch = warp + NW * (base + i);
      const bool okc = (base + i) < cpw && ch < nchunks;
      wlo[i] = (okc && ok_lo) ? ldg_stream(p_lo + ch * 32) : make_uint4(0u, 0u, 0u, 0u);
      whi[i] = (okc && ok_hi) ? ldg_stream(p_hi + ch * 32) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  auto load_x = [&](int item) {           // activations written by the previous kernel: L2-only loads
    const int base = (item % nb) * kGemvBatch;
#pragma unroll
    for (int i = 0; i < kGemvBatch; ++i) {
      const int ch = warp + NW * (base + i);
      xr[i] = (row_ok && (base + i) < cpw && ch < nchunks) ? ldcg16(xp + ch * 32) : make_uint4(0u, 0u, 0u, 0u);
    }
  };

  pdl_launch_dependents();
  if (nitems > 0) load_item(0);           // weights never change: prefetch before the dependency wait
  uint4 lw[HAS_LN ? kGemvBatch : 1], lb[HAS_LN ? kGemvBatch : 1];
  if constexpr (HAS_LN) {
#pragma unroll
    for (int i = 0; i < kGemvBatch; ++i) {
      const int ch = warp + NW * i;
      const bool okc = i < cpw && ch < nchunks;
      lw[i] = okc ? ldg_cached(a.ln_w + ch * 32 + 8 * t) : make_uint4(0u, 0u, 0u, 0u);
      lb[i] = okc ? ldg_cached(a.ln_b + ch * 32 + 8 * t) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  pdl_wait();
  if (nitems <= 0) return;

  if constexpr (HAS_LN) {
    load_x(0);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kGemvBatch; ++i) {
      float f[8];
      unpack8(xr[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
    s = quad_sum(s);
    if (t == 0) stat[warp][g] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) mean += stat[w][g];
    mean /= (float)a.K;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kGemvBatch; ++i) {
      const int ch = warp + NW * i;
      if (i < cpw && ch < nchunks) {
        float f[8];
        unpack8(xr[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float dlt = f[j] - mean; q += dlt * dlt; }
      }
    }
    q = quad_sum(q);
    __syncthreads();
    if (t == 0) stat[warp][g] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) var += stat[w][g];
    const float rstd = 1.0f / sqrtf(var / (float)a.K + a.ln_eps);
#pragma unroll
    for (int i = 0; i < kGemvBatch; ++i) {
      float f[8], wf[8], bfv[8];
      unpack8(xr[i], f); unpack8(lw[i], wf); unpack8(lb[i], bfv);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = row_ok ? (f[j] - mean) * rstd * wf[j] + bfv[j] : 0.f;
      xr[i] = pack8(f);                   // ln output is a bf16 tensor in the reference; 0 for padded chunks
    }
  }

  float c[4] = {0.f, 0.f, 0.f, 0.f};
  for (int item = 0; item < nitems; ++item) {
    if constexpr (!HAS_LN) load_x(item);
#pragma unroll
    for (int i = 0; i < kGemvBatch; ++i) {
      mma_bf16_16816(c, wlo[i].x, whi[i].x, wlo[i].y, whi[i].y, xr[i].x, xr[i].y);
      mma_bf16_16816(c, wlo[i].z, whi[i].z, wlo[i].w, whi[i].w, xr[i].z, xr[i].w);
    }
    if (item + 1 < nitems) load_item(item + 1);
    if ((item + 1) % nb != 0) continue;
    // ---- tile finished: deterministic cross-warp split-K reduction, then the epilogue
    const int tl = item / nb, tile = tile0 + tl;
    float (*rd)[16][8] = red[tl & 1];
    rd[warp][g][2 * t] = c[0]; rd[warp][g][2 * t + 1] = c[1];
    rd[warp][g + 8][2 * t] = c[2]; rd[warp][g + 8][2 * t + 1] = c[3];
    c[0] = c[1] = c[2] = c[3] = 0.f;
    __syncthreads();
    if (threadIdx.x < 128) {
      const int n = threadIdx.x & 15, mm = threadIdx.x >> 4;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) acc += rd[w][n][mm];
      const int col = tile * a.R + n;
      const bool ok = n < a.R && col < a.N && mm < a.B;
      float v = 0.f;
      if (ok) {
        const float bv = a.bias ? __bfloat162float(a.bias[col]) : 0.f;
        float rv = 0.f;
        if (a.res) rv = __bfloat162float(__ldcg(a.res + (int64_t)mm * a.N + col));
        v = epilogue_elem(acc, bv, a.act, a.res != nullptr, rv);
        const bf16 vb = __float2bfloat16_rn(v);
        a.Y[(int64_t)mm * a.N + col] = vb;
        if constexpr (EPI == EPI_QKV) {
          const int j = col - a.q_cols;
          const int pos = a.state->cur_len;
          if (j >= 0 && pos < a.tcap) {
            if (j < a.n_kv * a.d) {
              const int kvh = j / a.d, dim = j % a.d;
              a.kcache[(((int64_t)mm * a.n_kv + kvh) * a.tcap + pos) * a.d + dim] = vb;
            } else {
              const int jj = j - a.n_kv * a.d, kvh = jj / a.d, dim = jj % a.d;
              a.vtcache[(((int64_t)mm * a.n_kv + kvh) * a.d + dim) * a.tcap + pos] = vb;
            }
          }
        }
      }
      if constexpr (EPI == EPI_LMHEAD) {
        // argmax over the tile's <= 16 vocabulary entries of image row mm (lanes n = 0..15 are adjacent)
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
    // red[] is double-buffered by tile parity: one barrier per tile is enough
  }
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms = n > 0 ? n : 148;
  }
  return g_num_sms;
}

// Row tiling so that every SM streams (almost) the same number of weight rows.
static void plan_tiles(GemvArgs& a) {
  const int rows_per_cta = (a.N + num_sms() - 1) / num_sms();
  a.tiles_per_cta = (rows_per_cta + 15) / 16;
  a.R = (rows_per_cta + a.tiles_per_cta - 1) / a.tiles_per_cta;
  a.ntiles = (a.N + a.R - 1) / a.R;
}
int gemv_ntiles(int N) { GemvArgs a{}; a.N = N; plan_tiles(a); return a.ntiles; }

static int pick_nw(int K) { return (K / 32) > 128 ? 16 : 8; }
bool gemv8_supported(int K, bool has_ln) { return K % 32 == 0 && (!has_ln || K / 32 <= kGemvBatch * pick_nw(K)); }

template <int NW>
static void launch_gemvp_nw(GemvArgs& a, bool has_ln, int epi, bool pdl, cudaStream_t st) {
  plan_tiles(a);
  dim3 grid((a.ntiles + a.tiles_per_cta - 1) / a.tiles_per_cta), block(NW * 32);
  if (has_ln) {
    if (epi == EPI_QKV) launch_ex(gemvp_kernel<NW, true, EPI_QKV>, grid, block, 0, st, pdl, a);
    else if (epi == EPI_LMHEAD) launch_ex(gemvp_kernel<NW, true, EPI_LMHEAD>, grid, block, 0, st, pdl, a);
    else launch_ex(gemvp_kernel<NW, true, EPI_PLAIN>, grid, block, 0, st, pdl, a);
  } else {
    launch_ex(gemvp_kernel<NW, false, EPI_PLAIN>, grid, block, 0, st, pdl, a);
  }
}
static void launch_gemvp(GemvArgs& a, bool has_ln, int epi, bool pdl, cudaStream_t st) {
  if (pick_nw(a.K) == 8) launch_gemvp_nw<8>(a, has_ln, epi, pdl, st);
  else launch_gemvp_nw<16>(a, has_ln, epi, pdl, st);
}

void launch_gemv8(const bf16* x, const bf16* w, const bf16* bias, const bf16* res, bf16* y, int B, int N, int K,
                  int act, const bf16* ln_w, const bf16* ln_b, float ln_eps, bool pdl, cudaStream_t st) {
  GemvArgs a{};
  a.X = x; a.W = w; a.bias = bias; a.res = res; a.Y = y; a.B = B; a.N = N; a.K = K; a.act = act;
  a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps;
  launch_gemvp(a, ln_w != nullptr, EPI_PLAIN, pdl, st);
}

void launch_gemv8_qkv(const bf16* x, const bf16* w, const bf16* bias, bf16* y, int B, int N, int K, const bf16* ln_w,
                      const bf16* ln_b, float ln_eps, bf16* kcache, bf16* vtcache, const GenState* state, int q_cols,
                      int n_kv, int d, int tcap, bool pdl, cudaStream_t st) {
  GemvArgs a{};
  a.X = x; a.W = w; a.bias = bias; a.Y = y; a.B = B; a.N = N; a.K = K; a.act = 0;
  a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps;
  a.kcache = kcache; a.vtcache = vtcache; a.state = state; a.q_cols = q_cols; a.n_kv = n_kv; a.d = d; a.tcap = tcap;
  launch_gemvp(a, true, EPI_QKV, pdl, st);
}

void launch_gemv8_lmhead(const bf16* x, const bf16* w, bf16* logits, int B, int N, int K, const bf16* ln_w,
                         const bf16* ln_b, float ln_eps, float* amax_val, int* amax_idx, bool pdl, cudaStream_t st) {
  GemvArgs a{};
  a.X = x; a.W = w; a.Y = logits; a.B = B; a.N = N; a.K = K; a.act = 0;
  a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps; a.amax_val = amax_val; a.amax_idx = amax_idx;
  launch_gemvp(a, true, EPI_LMHEAD, pdl, st);
}

// ------------------------------------------------------------------------------------------
// Token selection + HF stop bookkeeping + next-token embedding in ONE single-CTA kernel.
// Greedy without a repetition penalty reduces the lm_head's per-tile argmax partials; otherwise the
// full bf16 logits row is scanned (penalty changes the order).  Semantics: SURVEY.md App. B.3-6.
__global__ void __launch_bounds__(1024) select_fused_kernel(const bf16* __restrict__ logits, int vocab, int batch,
                                                            const float* __restrict__ amax_val,
                                                            const int* __restrict__ amax_idx, int ntiles,
                                                            GenState* state, const GenParamsDev* __restrict__ p,
                                                            uint8_t* seen, int32_t* next_ids, int32_t* out_ids,
                                                            int advance_len, const bf16* __restrict__ wte,
                                                            const bf16* __restrict__ wpe, bf16* __restrict__ x, int h,
                                                            int n_positions) {
  pdl_launch_dependents();
  pdl_wait();
  if (state->done) return;
  __shared__ AmaxPair sm[32];
  __shared__ int s_tok[8];
  const int tid = threadIdx.x;
  const float rp = p->rep_penalty;
  const bool use_partials = (rp == 1.0f) && amax_val != nullptr;
  for (int b = 0; b < batch; ++b) {
    AmaxPair best{-INFINITY, 0x7fffffff};
    if (use_partials) {
      for (int i = tid; i < ntiles; i += 1024) {
        const float v = __ldcg(amax_val + (int64_t)i * 8 + b);
        const int id = __ldcg(amax_idx + (int64_t)i * 8 + b);
        best = amax_better(best, AmaxPair{v, id});
      }
    } else {
      const bf16* lr = logits + (int64_t)b * vocab;
      const uint8_t* sr = seen + (int64_t)b * vocab;
      for (int i = tid; i < vocab; i += 1024) {
        float v = __bfloat162float(__ldcg(lr + i));
        if (sr[i]) v = v < 0.f ? v * rp : v / rp;
        best = amax_better(best, AmaxPair{v, i});
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      AmaxPair other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = amax_better(best, other);
    }
    __syncthreads();
    if ((tid & 31) == 0) sm[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 32; ++w) best = amax_better(best, sm[w]);
      s_tok[b] = best.i == 0x7fffffff ? 0 : best.i;
    }
  }
  __syncthreads();
  if (tid == 0) select_apply_tokens(s_tok, batch, vocab, state, p, seen, next_ids, out_ids, advance_len);
  __syncthreads();
  // --- next step's input: wte[token] + wpe[position] (bf16 add), GPTBigCodeModel.forward
  int pos = state->cur_len;
  pos = pos >= n_positions ? n_positions - 1 : pos;
  const int hv = h >> 3;
  for (int i = tid; i < batch * hv; i += 1024) {
    const int b = i / hv, c = (i % hv) * 8;
    int id = s_tok[b];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    float e[8], q[8];
    unpack8(ldg_cached(wte + (int64_t)id * h + c), e);
    if (wpe) {                                                  // RoPE models have no learned position table
      unpack8(ldg_cached(wpe + (int64_t)pos * h + c), q);
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] += q[j];
    }
    *reinterpret_cast<uint4*>(x + (int64_t)b * h + c) = pack8(e);
  }
}

void launch_select_fused(const bf16* logits, int vocab, int batch, const float* amax_val, const int* amax_idx,
                         int ntiles, GenState* state, const GenParamsDev* params, uint8_t* seen, int32_t* next_ids,
                         int32_t* out_ids, int advance_len, const bf16* wte, const bf16* wpe, bf16* x, int h,
                         int n_positions, bool pdl, cudaStream_t st) {
  launch_ex(select_fused_kernel, dim3(1), dim3(1024), 0, st, pdl, logits, vocab, batch, amax_val, amax_idx, ntiles,
            state, params, seen, next_ids, out_ids, advance_len, wte, wpe, x, h, n_positions);
}

}  // namespace sv
