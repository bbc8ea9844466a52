// feas.cu — K1: the dense pods x templates Filter pass (SchedulablePodGroups for EVERY pending pod,
// core/scaleup/orchestrator/orchestrator.go:603-638 -> RunFiltersOnNode, plugin_runner.go:131).
//
// thread = pod.  NodeResourcesFit (noderesources/fit.go:649-736) is `request_r > free_r` for every
// requested resource; the int64 operands are order-preserving dictionary encoded at load time:
//     rank_req(v)  = 1-based index of v among the sorted distinct positive requests of that resource
//                    (0 when the pod does not request it: such a resource is never checked, fit.go:670-704)
//     rank_free(f) = number of distinct request values <= f
//     request > free  <=>  rank_req > rank_free            (exact, both directions)
// and the ranks of all resources are packed into W 32-bit words with one guard bit per field, so one
// subtraction per word compares every resource at once (a cleared guard bit = a borrow = "insufficient").
// Per 32 templates a lane assembles its pod's row of verdict bits, ANDs the class words of the
// size-independent plugins (pre_ok / post_ok), and a 5-stage shuffle transpose of the warp's 32x32 bit
// block yields the template-major words of the output bit matrix.  Output is flushed in full 32 B
// sectors with a popcount per template for the fit histogram.
#include <climits>

#include "engine.h"

namespace cae {

__global__ void expand_pods_kernel(const int32_t* __restrict__ pend_spec, int p_begin, int Pl, int W,
                                   const uint32_t* __restrict__ spec_w, const int32_t* __restrict__ spec_sc,
                                   const int32_t* __restrict__ spec_dc, uint32_t* __restrict__ pod_w,
                                   int32_t* __restrict__ pod_sc, int32_t* __restrict__ pod_dc) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Pl) return;
  int spec = pend_spec[p_begin + p];
  for (int w = 0; w < W; ++w) pod_w[(size_t)w * Pl + p] = spec_w[(size_t)spec * FEAS_MAX_W + w];
  pod_sc[p] = spec_sc[spec];
  pod_dc[p] = spec_dc[spec];
}

int launch_expand_pods(Engine* e) {
  if (e->Pl == 0) return 0;
  expand_pods_kernel<<<(e->Pl + 255) / 256, 256, 0, e->stream>>>(e->dobj.pend_spec, e->p_begin, e->Pl, e->W, e->d_spec_w,
                                                                   e->d_spec_sc, e->d_spec_dc, e->d_pod_w, e->d_pod_sc, e->d_pod_dc);
  e->stats.kernel_launches++;
  CAE_KERNEL_OK();
  return 0;
}

constexpr int K1_THREADS = 256;
constexpr int K1_TCHUNK = 128;  // templates per CTA
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_PAD = K1_TCHUNK + 4;

struct FeasGuards { uint32_t g[FEAS_MAX_W]; };

// lane i holds row i of a 32x32 bit matrix; afterwards lane j holds column j (bit i = M[i][j])
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t lowmask = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
    x = (lane & s) ? ((x & ~lowmask) | ((y & ~lowmask) >> s)) : ((x & lowmask) | ((y & lowmask) << s));
  }
  return x;
}

template <int W, bool REASONS>
__global__ void __launch_bounds__(K1_THREADS)
feasibility_kernel(int Pl, int Plw, int T, int Tw, int N, int U, FeasGuards guards,
                   const uint32_t* __restrict__ pod_w, const int32_t* __restrict__ pod_sc,
                   const int32_t* __restrict__ pod_dc, const uint32_t* __restrict__ tmpl_w,
                   const int32_t* __restrict__ tmpl_slots,
                   const uint32_t* __restrict__ pre_ok, const uint32_t* __restrict__ post_ok,
                   const uint8_t* __restrict__ pre_code, const uint8_t* __restrict__ post_code,
                   uint32_t* __restrict__ fit_bits, int32_t* __restrict__ fit_count,
                   uint8_t* __restrict__ reasons) {
  __shared__ __align__(16) uint32_t s_tw[W > 0 ? W : 1][K1_TCHUNK];
  __shared__ uint32_t s_out[K1_WARPS][K1_PAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x * K1_THREADS + tid;
  const int t0 = blockIdx.y * K1_TCHUNK;
  const int tn = min(K1_TCHUNK, T - t0);

  for (int i = tid; i < W * K1_TCHUNK; i += K1_THREADS) {
    const int w = i / K1_TCHUNK, j = i % K1_TCHUNK;
    s_tw[w][j] = (j < tn) ? tmpl_w[(size_t)w * T + t0 + j] : 0u;  // guard bits clear: every pod "fails" on padding
  }
  const bool valid = p < Pl;
  uint32_t pw[W > 0 ? W : 1];
#pragma unroll
  for (int w = 0; w < W; ++w) pw[w] = valid ? pod_w[(size_t)w * Pl + p] : 0u;
  const int sc = valid ? pod_sc[p] : 0;
  const int dc = valid ? pod_dc[p] : 0;
  __syncthreads();

#pragma unroll 1
  for (int tw = 0; tw < K1_TCHUNK / 32; ++tw) {
    const int wglob = t0 / 32 + tw;
    uint32_t row = 0;
    if (wglob < Tw) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int tl = tw * 32 + j;
        uint32_t bad = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) bad |= ~(s_tw[w][tl] - pw[w]) & guards.g[w];  // a cleared guard bit = insufficient
        if (bad == 0) row |= 1u << j;
        if (REASONS) {
          const int t = t0 + tl;
          if (valid && t < T) {
            // first failing plugin in Filter order: static plugins, NodeResourcesFit, then PTS / IPA
            uint8_t r = pre_code[(size_t)sc * U + N + t] & 0x0F;
            if (r == 0) r = (bad != 0 || tmpl_slots[t] < 1) ? CAE_R_FIT : post_code[(size_t)dc * T + t];
            reasons[(size_t)t * Pl + p] = r;
          }
        }
      }
      row &= valid ? (pre_ok[(size_t)sc * Tw + wglob] & post_ok[(size_t)dc * Tw + wglob]) : 0u;
    }
    s_out[warp][tw * 32 + lane] = warp_transpose32(row, lane);  // word of template t0 + tw*32 + lane over this warp's pods
  }
  __syncthreads();
  // flush: 8 consecutive words (one 32 B sector) per template row; popcount -> per-template counts
  const int pw0 = blockIdx.x * K1_WARPS;
  for (int i = tid; i < K1_TCHUNK * K1_WARPS; i += K1_THREADS) {
    const int tl = i / K1_WARPS, wv = i % K1_WARPS;
    const int t = t0 + tl;
    const uint32_t word = (t < T) ? s_out[wv][tl] : 0u;
    int c = __popc(word);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    if (t < T) {
      if (pw0 + wv < Plw && fit_bits) fit_bits[(size_t)t * Plw + pw0 + wv] = word;
      if (wv == 0 && c) atomicAdd(&fit_count[t], c);
    }
  }
}

template <int W>
static void launch_feas_w(Engine* e, bool want_reasons) {
  dim3 grid((e->Pl + K1_THREADS - 1) / K1_THREADS, (e->T + K1_TCHUNK - 1) / K1_TCHUNK);
  if (grid.x == 0 || grid.y == 0) return;
  FeasGuards g;
  for (int w = 0; w < FEAS_MAX_W; ++w) g.g[w] = e->feas_guard[w];
  if (want_reasons)
    feasibility_kernel<W, true><<<grid, K1_THREADS, 0, e->stream>>>(
        e->Pl, e->Plw, e->T, e->Tw, e->N, e->U, g, e->d_pod_w, e->d_pod_sc, e->d_pod_dc, e->d_tmpl_w, e->d_tmpl_slots,
        e->d_pre_ok, e->d_post_ok, e->d_pre_code, e->d_post_code, e->d_fit_bits, e->d_fit_count, e->d_reasons);
  else
    feasibility_kernel<W, false><<<grid, K1_THREADS, 0, e->stream>>>(
        e->Pl, e->Plw, e->T, e->Tw, e->N, e->U, g, e->d_pod_w, e->d_pod_sc, e->d_pod_dc, e->d_tmpl_w, e->d_tmpl_slots,
        e->d_pre_ok, e->d_post_ok, e->d_pre_code, e->d_post_code, e->d_fit_bits, e->d_fit_count, e->d_reasons);
  e->stats.kernel_launches++;
}

int launch_feasibility(Engine* e, bool want_reasons) {
  CAE_CUDA(cudaMemsetAsync(e->d_fit_count, 0, sizeof(int32_t) * e->T, e->stream));
  switch (e->W) {
    case 0: launch_feas_w<0>(e, want_reasons); break;
    case 1: launch_feas_w<1>(e, want_reasons); break;
    case 2: launch_feas_w<2>(e, want_reasons); break;
    case 3: launch_feas_w<3>(e, want_reasons); break;
    default: launch_feas_w<4>(e, want_reasons); break;
  }
  CAE_KERNEL_OK();
  return 0;
}

}  // namespace cae
