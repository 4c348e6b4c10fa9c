// Token selection fused with the next decode step's input (used by the CUDA-graph decode path): one single-CTA kernel finishes
// the greedy argmax from the lm_head kernel's per-tile partials (or scans the penalised logits), applies the HF stop / EOS
// bookkeeping (sv_select.cuh) and writes the next token's embedding.  Programmatic-dependent-launch ready: everything the
// previous kernel produced is read after `griddepcontrol.wait` with L2-only loads.
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
