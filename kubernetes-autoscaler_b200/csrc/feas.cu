// feas.cu — K1: the dense pods x templates Filter pass (SchedulablePodGroups for EVERY pending pod,
// core/scaleup/orchestrator/orchestrator.go:603-638 -> RunFiltersOnNode, plugin_runner.go:131).
//
// thread = pod.  NodeResourcesFit (noderesources/fit.go:649-736) is `request_r > free_r` for every
// requested resource; the int64 operands are order-preserving dictionary encoded at load time:
//     rank_req(v)  = 1-based index of v among the sorted distinct positive requests of that resource
//                    (0 when the pod does not request it: such a resource is never checked, fit.go:670-704)
//     rank_free(f) = number of distinct request values <= f
//     request > free  <=>  rank_req > rank_free            (exact, both directions)
// The template ranks are stored BIT-SLICED: slice b, word tw holds bit b of the ranks of templates
// tw*32 .. tw*32+31.  A lane compares its pod's rank against 32 templates at once with the classic
// MSB-first bit-serial comparator (gt |= eq & r & ~f; eq &= ~(r ^ f)), ~2.5 logic ops per slice per
// 32 evaluations, then ANDs the class words of the size-independent plugins (pre_ok / post_ok).
// A 5-stage shuffle transpose of the warp's 32x32 verdict block yields the template-major words of
// the output bit matrix; they are flushed in full 32 B sectors and pop-counted into the fit histogram.
//
// LUT variant (the default when the tables fit in shared memory): because the ranks are small dictionaries,
// "rank_req <= rank_free" for 32 templates is ONE word of a threshold bitmap indexed by (dim, rank_req):
//     lut[base_a + k][tw] bit j = (k <= rank_free_a(template tw*32+j))
// so a pod's verdict word is the AND of A shared-memory words and its two class words (pre_ok / post_ok,
// read through L1: neighbouring pods share classes) instead of ~3 logic ops per rank bit.  Rows are staged
// with an odd pitch so that lanes reading different rows hit different banks.
#include <algorithm>
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
constexpr int K1_TW = 16;                 // template words (x32 templates) per CTA
constexpr int K1_TCHUNK = K1_TW * 32;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_PAD = K1_TCHUNK + 4;

// Fused exchange of the fit histogram over peer memory (see cae_peer_attach in include/caengine.h)
struct PeerPush {
  int world;                 // 0 = disabled
  int32_t* accum[8];         // every rank's accumulator slot for this step (P2P-mapped)
  int32_t* arrive[8];        // every rank's arrival counter of that slot
  int32_t* done_ctr;         // local: thread blocks finished
};

// Publishes the all-reduced histogram once every rank's contribution has arrived, and clears the slot.
__global__ void peer_wait_kernel(int32_t* __restrict__ accum, volatile int32_t* arrive, int target, int T,
                                 int32_t* __restrict__ fit_count, int32_t* __restrict__ status) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int ok = 1;
    while (*arrive < target) {
      if (clock64() - t0 > 4000000000ll) { ok = 0; break; }  // ~2 s: a peer died; fail instead of hanging the GPU
      __nanosleep(200);
    }
    __threadfence_system();
    s_ok = ok;
    if (!ok) atomicExch(status, 1);
  }
  __syncthreads();
  if (!s_ok) return;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    fit_count[t] = __ldcg(&accum[t]);
    accum[t] = 0;
  }
}

struct FeasLayout {
  uint32_t fstart;           // bit b: slice b starts a field
  int nb;                    // real slices (the rest is zero padding)
  uint8_t sword[32], sshift[32];
};

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


// The LAST thread block to finish owns the complete local histogram: it adds it into every rank's exchange
// buffer over NVLink (system-scope atomics on peer memory), then signals arrival.
__device__ __forceinline__ void peer_push_tail(const PeerPush& pp, int32_t* __restrict__ fit_count, int T, int tid, int nthreads) {
  __shared__ int s_last;
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(pp.done_ctr, 1) == (int)(gridDim.x * gridDim.y) - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int t = tid; t < T; t += nthreads) {
      const int v = __ldcg(&fit_count[t]);
      if (v)
        for (int r = 0; r < pp.world; ++r) atomicAdd_system(pp.accum[r] + t, v);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pp.world) atomicAdd_system(pp.arrive[tid], 1);
    if (tid == 0) *pp.done_ctr = 0;
  }
}

// Same transpose with the per-lane constants hoisted: stage s sends rotl(x, amt_s) and merges under keep_s
// (1 funnel shift + 1 shuffle + 1 LOP3 per stage).
struct TransposeConsts { uint32_t amt[5], keep[5]; };
__device__ __forceinline__ TransposeConsts transpose_consts(int lane) {
  TransposeConsts c;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int s = 16 >> i;
    const uint32_t lowmask = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
    c.amt[i] = (lane & s) ? (uint32_t)s : (uint32_t)(32 - s);   // set lanes send their low part up, clear lanes their high part down
    const uint32_t k = (lane & s) ? ~lowmask : lowmask;
    asm volatile("mov.b32 %0, %1;" : "=r"(c.keep[i]) : "r"(k));  // opaque: keep it in a register, do not re-derive it per use
  }
  return c;
}
__device__ __forceinline__ uint32_t warp_transpose32c(uint32_t x, const TransposeConsts& c) {
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const uint32_t y = __shfl_xor_sync(0xffffffffu, __funnelshift_l(x, x, c.amt[i]), 16 >> i);
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(x) : "r"(x), "r"(y), "r"(c.keep[i]));  // keep ? x : y, bitwise
  }
  return x;
}

struct LutLayout {
  int A_rows;                // resource rows; then SC static-class rows, then DC dynamic-class rows
  int SC, DC;
  int base[CAE_MAX_RES];
  uint32_t mask[CAE_MAX_RES];
  uint8_t word[CAE_MAX_RES], shift[CAE_MAX_RES];
};
constexpr int K1_LPITCH = K1_TW + 1;     // odd row pitch: distinct rows -> distinct banks

template <int A, bool REASONS, int NW>
__global__ void __launch_bounds__(NW * 32)
feasibility_lut_kernel(int Pl, int Plw, int T, int Tw, int N, int U, int W, LutLayout lay,
                       const uint32_t* __restrict__ pod_w, const int32_t* __restrict__ pod_sc,
                       const int32_t* __restrict__ pod_dc, const uint32_t* __restrict__ rlut,
                       const int32_t* __restrict__ tmpl_slots,
                       const uint32_t* __restrict__ pre_ok, const uint32_t* __restrict__ post_ok,
                       const uint8_t* __restrict__ pre_code, const uint8_t* __restrict__ post_code,
                       uint32_t* __restrict__ fit_bits, int32_t* __restrict__ fit_count,
                       uint8_t* __restrict__ reasons, PeerPush pp) {
  extern __shared__ uint32_t k1_smem[];
  const int rows = lay.A_rows;                                 // resource rows only: the class rows stay in global / L1
  uint32_t* s_lut = k1_smem;                                   // [rows][K1_LPITCH]
  constexpr int NT = NW * 32;
  constexpr int PAD = K1_TCHUNK + 32 / NW;   // flush reads (wv, tl) hit 32 distinct banks
  uint32_t* s_out = s_lut + (size_t)rows * K1_LPITCH;          // [NW][PAD]
  int32_t* s_cnt = reinterpret_cast<int32_t*>(s_out + NW * PAD);  // [K1_TCHUNK]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x * NT + tid;
  const int tw0 = blockIdx.y * K1_TW;
  const int t0 = tw0 * 32;

  for (int i = tid; i < rows * K1_TW; i += NT) {
    const int row = i / K1_TW, w = i % K1_TW;
    s_lut[row * K1_LPITCH + w] = (tw0 + w < Tw) ? rlut[(size_t)row * Tw + tw0 + w] : 0u;
  }
  for (int i = tid; i < K1_TCHUNK; i += NT) s_cnt[i] = 0;
  const bool valid = p < Pl;
  int off[A > 0 ? A : 1];
  {
    uint32_t pw[FEAS_MAX_W];
#pragma unroll
    for (int w = 0; w < FEAS_MAX_W; ++w) pw[w] = (valid && w < W) ? pod_w[(size_t)w * Pl + p] : 0u;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const uint32_t word = lay.word[a] == 0 ? pw[0] : lay.word[a] == 1 ? pw[1] : lay.word[a] == 2 ? pw[2] : pw[3];
      off[a] = (lay.base[a] + (int)((word >> lay.shift[a]) & lay.mask[a])) * K1_LPITCH;
    }
  }
  const int sc = valid ? pod_sc[p] : 0;
  const int dc = valid ? pod_dc[p] : 0;
  // class words of this pod for the 16 template words (mostly L1 hits: neighbouring pods share classes)
  uint32_t cls[K1_TW];
#pragma unroll
  for (int tw = 0; tw < K1_TW; ++tw)
    cls[tw] = (valid && tw0 + tw < Tw) ? (__ldg(&pre_ok[(size_t)sc * Tw + tw0 + tw]) & __ldg(&post_ok[(size_t)dc * Tw + tw0 + tw])) : 0u;
  const TransposeConsts tc = transpose_consts(lane);
  __syncthreads();

#pragma unroll
  for (int tw = 0; tw < K1_TW; ++tw) {
    uint32_t fit = 0xffffffffu;
#pragma unroll
    for (int a = 0; a < A; ++a) fit &= s_lut[off[a] + tw];
    const uint32_t row = fit & cls[tw];
    if (REASONS) {
      const int wglob = tw0 + tw;
      if (valid && wglob < Tw) {
        for (int j = 0; j < 32; ++j) {
          const int t = wglob * 32 + j;
          if (t >= T) break;
          uint8_t rs = pre_code[(size_t)sc * U + N + t] & 0x0F;
          if (rs == 0) rs = (!((fit >> j) & 1u) || tmpl_slots[t] < 1) ? CAE_R_FIT : post_code[(size_t)dc * T + t];
          reasons[(size_t)t * Pl + p] = rs;
        }
      }
    }
    const uint32_t col = warp_transpose32c(row, tc);
    s_out[warp * PAD + tw * 32 + lane] = col;
    atomicAdd(&s_cnt[tw * 32 + lane], __popc(col));
  }
  __syncthreads();
  const int pw0 = blockIdx.x * NW;
  for (int i = tid; i < K1_TCHUNK * NW; i += NT) {
    const int tl = i / NW, wv = i % NW;
    const int t = t0 + tl;
    if (t < T && pw0 + wv < Plw && fit_bits) fit_bits[(size_t)t * Plw + pw0 + wv] = s_out[wv * PAD + tl];
  }
  for (int i = tid; i < K1_TCHUNK; i += NT) {
    const int c = s_cnt[i];
    if (c && t0 + i < T) atomicAdd(&fit_count[t0 + i], c);
  }
  if (pp.world) peer_push_tail(pp, fit_count, T, tid, NT);
}

template <int B, bool REASONS>
__global__ void __launch_bounds__(K1_THREADS)
feasibility_kernel(int Pl, int Plw, int T, int Tw, int N, int U, int W, FeasLayout lay,
                   const uint32_t* __restrict__ pod_w, const int32_t* __restrict__ pod_sc,
                   const int32_t* __restrict__ pod_dc, const uint32_t* __restrict__ tslice,
                   const int32_t* __restrict__ tmpl_slots,
                   const uint32_t* __restrict__ pre_ok, const uint32_t* __restrict__ post_ok,
                   const uint8_t* __restrict__ pre_code, const uint8_t* __restrict__ post_code,
                   uint32_t* __restrict__ fit_bits, int32_t* __restrict__ fit_count,
                   uint8_t* __restrict__ reasons, PeerPush pp) {
  __shared__ uint32_t s_sl[B > 0 ? B : 1][K1_TW];
  __shared__ uint32_t s_out[K1_WARPS][K1_PAD];
  __shared__ int32_t s_cnt[K1_TCHUNK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x * K1_THREADS + tid;
  const int tw0 = blockIdx.y * K1_TW;
  const int t0 = tw0 * 32;

  for (int i = tid; i < B * K1_TW; i += K1_THREADS) {
    const int b = i / K1_TW, w = i % K1_TW;
    s_sl[b][w] = (tw0 + w < Tw) ? tslice[(size_t)b * Tw + tw0 + w] : 0u;
  }
  for (int i = tid; i < K1_TCHUNK; i += K1_THREADS) s_cnt[i] = 0;
  const bool valid = p < Pl;
  // the pod's rank bits as all-ones / all-zeros masks, one register per slice
  uint32_t r[B > 0 ? B : 1];
  {
    uint32_t pw[FEAS_MAX_W];
#pragma unroll
    for (int w = 0; w < FEAS_MAX_W; ++w) pw[w] = (valid && w < W) ? pod_w[(size_t)w * Pl + p] : 0u;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const uint32_t word = lay.sword[b] == 0 ? pw[0] : lay.sword[b] == 1 ? pw[1] : lay.sword[b] == 2 ? pw[2] : pw[3];
      r[b] = b < lay.nb ? 0u - ((word >> lay.sshift[b]) & 1u) : 0u;
    }
  }
  const int sc = valid ? pod_sc[p] : 0;
  const int dc = valid ? pod_dc[p] : 0;
  __syncthreads();

#pragma unroll 1
  for (int tw = 0; tw < K1_TW; ++tw) {
    const int wglob = tw0 + tw;
    uint32_t row = 0;
    if (wglob < Tw) {
      // bit-serial "rank_req > rank_free" for 32 templates at once; fields concatenated, MSB first
      uint32_t gt = 0, eq = 0;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const uint32_t f = s_sl[b][tw];
        if ((lay.fstart >> b) & 1u) eq = 0xffffffffu;  // uniform: a new field starts
        gt |= eq & r[b] & ~f;
        eq &= ~(r[b] ^ f);
      }
      row = valid ? (~gt & pre_ok[(size_t)sc * Tw + wglob] & post_ok[(size_t)dc * Tw + wglob]) : 0u;
      if (REASONS) {
        if (valid) {
          for (int j = 0; j < 32; ++j) {
            const int t = wglob * 32 + j;
            if (t >= T) break;
            // first failing plugin in Filter order: static plugins, NodeResourcesFit, then PTS / IPA
            uint8_t rs = pre_code[(size_t)sc * U + N + t] & 0x0F;
            if (rs == 0) rs = (((gt >> j) & 1u) || tmpl_slots[t] < 1) ? CAE_R_FIT : post_code[(size_t)dc * T + t];
            reasons[(size_t)t * Pl + p] = rs;
          }
        }
      }
    }
    const uint32_t col = warp_transpose32(row, lane);  // word of template t0 + tw*32 + lane over this warp's pods
    s_out[warp][tw * 32 + lane] = col;
    if (col) atomicAdd(&s_cnt[tw * 32 + lane], __popc(col));
  }
  __syncthreads();
  // flush: 8 consecutive words (one 32 B sector) per template row
  const int pw0 = blockIdx.x * K1_WARPS;
  for (int i = tid; i < K1_TCHUNK * K1_WARPS; i += K1_THREADS) {
    const int tl = i / K1_WARPS, wv = i % K1_WARPS;
    const int t = t0 + tl;
    if (t < T && pw0 + wv < Plw && fit_bits) fit_bits[(size_t)t * Plw + pw0 + wv] = s_out[wv][tl];
  }
  for (int i = tid; i < K1_TCHUNK; i += K1_THREADS) {
    const int c = s_cnt[i];
    if (c && t0 + i < T) atomicAdd(&fit_count[t0 + i], c);
  }
  if (pp.world) peer_push_tail(pp, fit_count, T, tid, K1_THREADS);
}

static PeerPush peer_push_args(Engine* e) {
  PeerPush pp{};
  if (e->peer_world > 1 && e->T <= Engine::PEER_CAP) {
    const int slot = (int)(e->peer_step & 1);
    pp.world = e->peer_world;
    for (int r = 0; r < e->peer_world; ++r) {
      pp.accum[r] = e->peer_base[r] + (size_t)slot * Engine::PEER_CAP;
      pp.arrive[r] = e->peer_base[r] + (size_t)2 * Engine::PEER_CAP + slot;
    }
    pp.done_ctr = e->d_xbuf + (size_t)2 * Engine::PEER_CAP + 8;
  }
  return pp;
}

static void launch_peer_wait(Engine* e) {
  const int slot = (int)(e->peer_step & 1);
  e->peer_uses[slot] += 1;
  e->peer_step += 1;
  peer_wait_kernel<<<1, 256, 0, e->stream>>>(e->d_xbuf + (size_t)slot * Engine::PEER_CAP,
                                             e->d_xbuf + (size_t)2 * Engine::PEER_CAP + slot,
                                             (int)(e->peer_uses[slot] * e->peer_world), e->T, e->d_fit_count,
                                             e->d_xbuf + (size_t)2 * Engine::PEER_CAP + 9);
  e->stats.kernel_launches++;
}

template <int B>
static void launch_feas_b(Engine* e, bool want_reasons, dim3 grid, const PeerPush& pp) {
  FeasLayout lay;
  lay.fstart = e->feas_fstart;
  lay.nb = e->feas_B;
  for (int b = 0; b < 32; ++b) { lay.sword[b] = e->feas_sword[b]; lay.sshift[b] = e->feas_sshift[b]; }
  if (want_reasons)
    feasibility_kernel<B, true><<<grid, K1_THREADS, 0, e->stream>>>(
        e->Pl, e->Plw, e->T, e->Tw, e->N, e->U, e->W, lay, e->d_pod_w, e->d_pod_sc, e->d_pod_dc, e->d_tslice, e->d_tmpl_slots,
        e->d_pre_ok, e->d_post_ok, e->d_pre_code, e->d_post_code, e->d_fit_bits, e->d_fit_count, e->d_reasons, pp);
  else
    feasibility_kernel<B, false><<<grid, K1_THREADS, 0, e->stream>>>(
        e->Pl, e->Plw, e->T, e->Tw, e->N, e->U, e->W, lay, e->d_pod_w, e->d_pod_sc, e->d_pod_dc, e->d_tslice, e->d_tmpl_slots,
        e->d_pre_ok, e->d_post_ok, e->d_pre_code, e->d_post_code, e->d_fit_bits, e->d_fit_count, e->d_reasons, pp);
}

template <int A, bool REASONS, int NW>
static int launch_feas_lut_arw(Engine* e, const PeerPush& pp, const LutLayout& lay, int rows) {
  dim3 grid((e->Pl + NW * 32 - 1) / (NW * 32), (e->Tw + K1_TW - 1) / K1_TW);
  const size_t smem = sizeof(uint32_t) * ((size_t)std::max(rows, 1) * K1_LPITCH + NW * (K1_TCHUNK + 32 / NW) + K1_TCHUNK);
  auto kern = feasibility_lut_kernel<A, REASONS, NW>;
  if (smem > 48 * 1024) CAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NW * 32, smem, e->stream>>>(
      e->Pl, e->Plw, e->T, e->Tw, e->N, e->U, e->W, lay, e->d_pod_w, e->d_pod_sc, e->d_pod_dc, e->d_rlut, e->d_tmpl_slots,
      e->d_pre_ok, e->d_post_ok, e->d_pre_code, e->d_post_code, e->d_fit_bits, e->d_fit_count, e->d_reasons, pp);
  return 0;
}

template <int A>
static int launch_feas_lut_a(Engine* e, bool want_reasons, const PeerPush& pp, const LutLayout& lay, int rows) {
  if (e->k1_warps == 8)
    return want_reasons ? launch_feas_lut_arw<A, true, 8>(e, pp, lay, rows) : launch_feas_lut_arw<A, false, 8>(e, pp, lay, rows);
  return want_reasons ? launch_feas_lut_arw<A, true, 16>(e, pp, lay, rows) : launch_feas_lut_arw<A, false, 16>(e, pp, lay, rows);
}

constexpr int K1_LUT_MAX_ROWS = 1024;    // 68 KB of threshold rows per thread block at most

int launch_feasibility(Engine* e, bool want_reasons) {
  CAE_CUDA(cudaMemsetAsync(e->d_fit_count, 0, sizeof(int32_t) * e->T, e->stream));
  if (e->Pl == 0 || e->Tw == 0) return 0;
  const PeerPush pp = peer_push_args(e);
  const int rows = e->lut_rows;
  if (!e->force_bitslice && rows <= K1_LUT_MAX_ROWS) {
    LutLayout lay{};
    lay.A_rows = e->lut_rows; lay.SC = e->SC; lay.DC = e->DC;
    for (int a = 0; a < e->A; ++a) {
      lay.base[a] = e->lut_base[a]; lay.mask[a] = e->lut_mask[a]; lay.word[a] = e->lut_word[a]; lay.shift[a] = e->lut_shift[a];
    }
    int rc = 0;
    switch (e->A) {
      case 0: rc = launch_feas_lut_a<0>(e, want_reasons, pp, lay, rows); break;
      case 1: rc = launch_feas_lut_a<1>(e, want_reasons, pp, lay, rows); break;
      case 2: rc = launch_feas_lut_a<2>(e, want_reasons, pp, lay, rows); break;
      case 3: rc = launch_feas_lut_a<3>(e, want_reasons, pp, lay, rows); break;
      case 4: rc = launch_feas_lut_a<4>(e, want_reasons, pp, lay, rows); break;
      case 5: rc = launch_feas_lut_a<5>(e, want_reasons, pp, lay, rows); break;
      case 6: rc = launch_feas_lut_a<6>(e, want_reasons, pp, lay, rows); break;
      case 7: rc = launch_feas_lut_a<7>(e, want_reasons, pp, lay, rows); break;
      default: rc = launch_feas_lut_a<8>(e, want_reasons, pp, lay, rows); break;
    }
    if (rc) return rc;
  } else {
    // slices beyond feas_B are all-zero with r = 0: they change nothing (padding to a multiple of 4)
    dim3 grid((e->Pl + K1_THREADS - 1) / K1_THREADS, (e->Tw + K1_TW - 1) / K1_TW);
    const int Bp = e->feas_B == 0 ? 0 : (e->feas_B + 3) / 4 * 4;
    switch (Bp) {
      case 0: launch_feas_b<0>(e, want_reasons, grid, pp); break;
      case 4: launch_feas_b<4>(e, want_reasons, grid, pp); break;
      case 8: launch_feas_b<8>(e, want_reasons, grid, pp); break;
      case 12: launch_feas_b<12>(e, want_reasons, grid, pp); break;
      case 16: launch_feas_b<16>(e, want_reasons, grid, pp); break;
      case 20: launch_feas_b<20>(e, want_reasons, grid, pp); break;
      case 24: launch_feas_b<24>(e, want_reasons, grid, pp); break;
      case 28: launch_feas_b<28>(e, want_reasons, grid, pp); break;
      default: launch_feas_b<32>(e, want_reasons, grid, pp); break;
    }
  }
  e->stats.kernel_launches++;
  CAE_KERNEL_OK();
  if (pp.world) {
    launch_peer_wait(e);
    CAE_KERNEL_OK();
  }
  return 0;
}

}  // namespace cae
