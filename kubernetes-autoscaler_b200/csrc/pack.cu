// pack.cu — K3: BinpackingNodeEstimator.Estimate (estimator/binpacking_estimator.go:97-247) on the GPU.
//
// One WARP per template (templates are independent simulations; inside one, placement order is
// sequential by construction), persistent with an atomic work counter.  Per-template node state
// lives in a per-warp global slab (L1/L2 resident); lanes stride over the nodes.
//
// Plain groups (identical pods, no topology spread / inter-pod affinity involvement) are placed in
// CLOSED FORM:
//   * tryToScheduleOnExistingNodes (:141-164): SchedulePodOnAnyNodeMatching scans cyclically from
//     lastIndex (predicate/plugin_runner.go:81,123), so identical pods are dealt round-robin over the
//     added nodes with spare capacity k_j: every node gets min(k_j, L), the first `rem` nodes in
//     cyclic order with k_j > L get one more (L = largest lap count with sum min(k_j, L) <= n).
//   * tryToScheduleOnNewNodes (:168-247): only the last added node is tried, so each new node takes
//     min(remaining, k_new) until the limiter denies (:222); an empty last node stops the group (:212);
//     a pod that fits no fresh node still adds one (:227-240).
// Dynamic groups (dyn.cuh) run the reference's per-pod loop, but against INCREMENTAL counters
// instead of a PreFilter rescan per pod: per-domain counts are seeded from the cluster base counts,
// the nodes added so far and the run's placement log, then updated on every placement; nodes that
// failed are stamped and skipped until a counter minimum / presence changes (only then can a failed
// node become feasible again).
//
// FM = true turns the same machinery into HintingSimulator.TrySchedulePods on the CLUSTER snapshot
// (simulator/scheduling/hinting_simulator.go:53-135; filterOutSchedulableByPacking,
// core/podlistprocessor/filter_out_schedulable.go:96-126): one simulation (one warp), no template, the node
// list is the N cluster nodes, pods arrive as runs of consecutive identical pods in the caller's order and
// every pod is placed with the per-pod loop (hint first, then SchedulePodOnAnyNodeMatching from lastIndex),
// with the SimilarPodsScheduling shortcut (similar_pods.go:59-104).
#include <algorithm>
#include <climits>
#include <vector>

#include "engine.h"

namespace cae {

struct PackParams {
  int E, T, N, U, t_begin, t_end, cap, has_dyn, dstride, log_cap, nblk;
  size_t bmax_off;
  const int32_t *order, *order_n;
  const int32_t* perm;     // work order of the templates (NULL = index order)
  const uint8_t* pre_code;
  const int32_t *spec_sc, *spec_dc;
  const int64_t* tmpl_free;  // [A][T]
  const int32_t *tmpl_slots, *max_nodes, *pc_of;
  const unsigned long long* port_conf;
  const int64_t* c_free;  // [A][N]
  const int32_t* c_slots;
  int act_dim[CAE_MAX_RES];
  int32_t *node_count, *pod_count, *sched, *work_counter, *status;
  unsigned char* scratch;
  size_t scratch_per_warp;
  // FM (filter-out-schedulable) only
  int fm_runs, fm_last_index, fm_break;
  const int32_t *fm_run_off, *fm_pods, *fm_hint, *fm_class, *fm_class_ctrl;
  const uint8_t* fm_node_ok;
  int32_t *fm_assigned, *fm_out;          // [P] node or -1; {lastIndex, overflowing controllers, pods scheduled}
  int32_t* fm_ctrl_cnt;                   // [controllers] classes stored per controller (zeroed)
  uint8_t *fm_class_mark, *fm_ctrl_over;  // [classes] known unschedulable, [controllers] overflowing (zeroed)
  int fm_nctrl;
};

__device__ __forceinline__ int wsum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long wsum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int wmax(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int wmin(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// per-warp description of the dynamic group being placed (shared memory, uniform reads)
struct WarpDyn {
  int nq;
  int qid[DYN_MAX_Q], kind[DYN_MAX_Q], k[DYN_MAX_Q], host[DYN_MAX_Q], Dc[DYN_MAX_Q], tslot[DYN_MAX_Q];
  int wown[DYN_MAX_Q], self[DYN_MAX_Q], maxskew[DYN_MAX_Q], mindom[DYN_MAX_Q], elig_new[DYN_MAX_Q], dsw[DYN_MAX_Q];
  int minv[DYN_MAX_Q], nmin[DYN_MAX_Q], ndom[DYN_MAX_Q], tot[DYN_MAX_Q], boff[DYN_MAX_Q];
  int aff_self;
};

constexpr int PACK_WARPS = 4;
constexpr int PACK_HIST = 256;             // capacities up to this use the histogram; above, a binary search
#ifndef PACK_MIN_BLOCKS
#define PACK_MIN_BLOCKS 1
#endif

template <int A, bool FM>
__global__ void __launch_bounds__(PACK_WARPS * 32, PACK_MIN_BLOCKS) pack_kernel(DevObjects o, DynTables d, PackParams p) {
  __shared__ WarpDyn s_wd[PACK_WARPS];
  __shared__ int s_hist[PACK_WARPS][PACK_HIST + 1];   // plain groups: #nodes per capacity value (closed-form lap count)
  const int lane = threadIdx.x & 31;
  WarpDyn& wd = s_wd[threadIdx.x >> 5];
  int* hist = s_hist[threadIdx.x >> 5];
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned char* slab = p.scratch + (size_t)warp_global * p.scratch_per_warp;
  const int N = p.N, NT = p.N + p.T;
  if (FM && warp_global != 0) return;
  const int Neff = (p.has_dyn || FM) ? N : 0;  // cluster nodes carry run state only when a placement can reach them
  const int X = Neff + p.cap;
  int32_t* hdr = reinterpret_cast<int32_t*>(slab);  // stamp / version counters survive across launches
  int64_t* nfree = reinterpret_cast<int64_t*>(slab + 16);                             // [A][X]
  unsigned long long* nports = reinterpret_cast<unsigned long long*>(nfree + (size_t)(A > 0 ? A : 1) * X);  // [X]
  int32_t* nslots = reinterpret_cast<int32_t*>(nports + X);                           // [X]
  int32_t* kbuf = nslots + X;                                                         // [X] capacities of the current pass
  int32_t* stamp = kbuf + X;                                                          // [X] "failed at stamp"
  int32_t* wcnt = stamp + X;                                                          // [DYN_MAX_Q][dstride]
  int32_t* wpres = wcnt + (size_t)DYN_MAX_Q * p.dstride;                              // [DYN_MAX_Q][dstride]
  int32_t* wver = wpres + (size_t)DYN_MAX_Q * p.dstride;                              // [DYN_MAX_Q][dstride] slot version
  int32_t* logbuf = wver + (size_t)DYN_MAX_Q * p.dstride;                             // [log_cap][3]
  int32_t* bslots = logbuf + (size_t)p.log_cap * 3;                                   // [nblk] upper bound of the pod slots in a 32-node block
  int32_t* cmw = bslots + p.nblk;                                                     // [nblk/32+2] candidate-block masks of the current group
  int32_t* kcap = kbuf + Neff;                                                        // capacities of the added nodes for the current group
  int64_t* bmax = reinterpret_cast<int64_t*>(slab + p.bmax_off);                      // [A][nblk] upper bound of the free capacity in a block
  uint8_t* nsched = reinterpret_cast<uint8_t*>(bmax + (size_t)(A > 0 ? A : 1) * p.nblk);  // [X]
  int stamp_ctr = hdr[0] + 1, gver_ctr = hdr[1] + 1;

  for (;;) {
    int t = 0;
    if (lane == 0) {
      const int i = atomicAdd(p.work_counter, 1);
      t = i >= p.t_end - p.t_begin ? p.t_end : (p.perm ? p.perm[i] : p.t_begin + i);
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= p.t_end) break;

    int64_t tfree[A > 0 ? A : 1];
#pragma unroll
    for (int a = 0; a < A; ++a) tfree[a] = FM ? 0 : p.tmpl_free[(size_t)a * p.T + t];
    const int tslots = FM ? 0 : p.tmpl_slots[t];
    const int max_nodes = (!FM && p.max_nodes) ? p.max_nodes[t] : 0;
    const int col_new = N + p.T + t;  // universe column of the sanitized template
    int n_new = 0, nodes_with_pods = 0, pods_total = 0, last_index = FM ? p.fm_last_index : 0, log_n = 0;
    bool new_nodes_available = !FM, cl_init = false, overflow = false, fm_stop = false, fm_moved = false;
    const int n_groups = FM ? p.fm_runs : p.order_n[t];

    // ---- shared helpers -----------------------------------------------------------------------
    auto slot_of = [&](int q, int x) -> int {
      if (x < Neff) return d.dom[(size_t)wd.k[q] * NT + x];
      if (wd.host[q]) return wd.Dc[q] + 1 + (x - Neff);
      return wd.tslot[q];
    };
    auto elig_of = [&](int q, int x) -> bool {
      return x < Neff ? d.elig[(size_t)wd.qid[q] * p.U + x] != 0 : wd.elig_new[q] != 0;
    };
    auto log_append = [&](bool mine, int x, int spec, int cnt) {  // warp-aggregated append
      unsigned m = __ballot_sync(0xffffffffu, mine);
      if (!m) return;
      int rank = __popc(m & ((1u << lane) - 1));
      if (mine) {
        int idx = log_n + rank;
        if (idx < p.log_cap) { logbuf[idx * 3] = x; logbuf[idx * 3 + 1] = spec; logbuf[idx * 3 + 2] = cnt; }
      }
      log_n += __popc(m);
      if (log_n > p.log_cap) overflow = true;
    };

    // Per 32-node block of added nodes: upper bounds of the free capacity / pod slots.  free only shrinks, so a
    // stale bound stays valid; blocks whose bound is below the request are skipped without touching their nodes.
    auto refresh_block = [&](int b) {  // exact maxima of block b (all lanes)
      const int j = b * 32 + lane;
      const bool in = j < n_new;
      const int x = Neff + (in ? j : 0);
      int ms = in ? nslots[x] : INT_MIN;
      ms = wmax(ms);
#pragma unroll
      for (int a = 0; a < A; ++a) {
        long long v = in ? nfree[(size_t)a * X + x] : LLONG_MIN;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) bmax[(size_t)a * p.nblk + b] = v;
      }
      if (lane == 0) bslots[b] = ms;
    };

    for (int gi = 0; gi < n_groups; ++gi) {
      const int g = FM ? 0 : (p.order[(size_t)t * p.E + gi] & ~ORDER_NOT_ON_FRESH);
      const int pb = FM ? p.fm_run_off[gi] : o.group_off[g];
      int n = (FM ? p.fm_run_off[gi + 1] : o.group_off[g + 1]) - pb;
      const int spec = o.pend_spec[FM ? p.fm_pods[pb] : pb];
      int64_t req[A > 0 ? A : 1];
#pragma unroll
      for (int a = 0; a < A; ++a) req[a] = o.ps_req[(size_t)spec * R + p.act_dim[a]];
      const int sc = p.spec_sc[spec];
      const int dc = p.has_dyn ? p.spec_dc[spec] : 0;
      const bool static_new = !FM && (p.pre_code[(size_t)sc * p.U + col_new] & 0x0F) == 0;
      const int plist = o.ps_port_list[spec];
      const bool has_ports = o.port_off[plist + 1] > o.port_off[plist];
      const unsigned long long pconf = has_ports ? p.port_conf[plist] : 0ull;  // port sets this pod collides with
      const unsigned long long pbit = has_ports ? (1ull << p.pc_of[plist]) : 0ull;
      bool feeds = p.has_dyn && !FM && d.group_feeds[g];
      if (FM && p.has_dyn) {   // runs are not groups: log every placement that some counter counts
        bool any = false;
        for (int q = lane; q < d.Q; q += 32) any |= d.wmat[(size_t)q * d.S + spec] != 0;
        feeds = __ballot_sync(0xffffffffu, any) != 0u;
      }
      int placed = 0;

      if (dc == 0 && !FM) {
        // ======================= plain group: closed form =======================================
        if (n_new > 0 && static_new) {
          const int list_len = N + n_new;
          const int s = last_index >= N ? last_index - N : 0;  // first added node in cyclic scan order
          const int nb = (n_new + 31) >> 5;
          // candidate blocks: lanes test 32 block summaries at once
          auto block_may_fit = [&](int b2) -> bool {
            bool ok = bslots[b2] >= 1;
#pragma unroll
            for (int a = 0; a < A; ++a) ok &= !(req[a] > 0 && bmax[(size_t)a * p.nblk + b2] < req[a]);
            return ok;
          };
          long long total = 0;
          int kmax = 0;
#pragma unroll
          for (int i = 0; i < PACK_HIST / 32; ++i) hist[lane * (PACK_HIST / 32) + i + 1] = 0;
          __syncwarp();
          for (int cb = 0; cb < nb; cb += 32) {
            const int b2 = cb + lane;
            unsigned cm = __ballot_sync(0xffffffffu, b2 < nb && block_may_fit(b2));
            if (lane == 0) cmw[cb >> 5] = (int)cm;
            while (cm) {
              const int bb = cb + __ffs(cm) - 1;
              cm &= cm - 1;
              const int j = bb * 32 + lane;
              int kblk = 0;
              if (j < n_new) {
                const int x = Neff + j;
                int k = min(nslots[x], n);
                if (k > 0 && (nports[x] & pconf)) k = 0;
#pragma unroll
                for (int a = 0; a < A; ++a) {
                  if (req[a] > 0 && k > 0) {
                    const int64_t f = nfree[(size_t)a * X + x];
                    if (f < req[a]) k = 0;
                    else if (f < (int64_t)k * req[a]) k = (int)(f / req[a]);
                  }
                }
                if (has_ports) k = min(k, 1);
                kcap[j] = k;
                total += k;
                kmax = max(kmax, k);
                kblk = k;
              }
              {  // histogram of the capacities: lanes with the same value elect one writer (no atomics)
                const bool cnt = kblk > 0 && kblk <= PACK_HIST;
                const unsigned peers = __match_any_sync(0xffffffffu, cnt ? kblk : 0);
                if (cnt && lane == __ffs(peers) - 1) hist[kblk] += __popc(peers);
                __syncwarp();
              }
              // a candidate block that cannot take a single pod had a stale bound: tighten it so that the
              // following groups skip it without touching its nodes
              if (__ballot_sync(0xffffffffu, kblk > 0) == 0u) {
                refresh_block(bb);
                if (lane == 0) cmw[cb >> 5] &= ~(1 << (bb - cb));
              }
            }
          }
          total = wsum_ll(total);
          kmax = wmax(kmax);
          __syncwarp();
          if (total > 0) {
            auto sum_min = [&](int lim) -> long long {  // sum over candidate blocks of min(k_j, lim)
              long long f = 0;
              for (int cb = 0; cb < nb; cb += 32) {
                unsigned cm = (unsigned)cmw[cb >> 5];
                while (cm) {
                  const int bb = cb + __ffs(cm) - 1;
                  cm &= cm - 1;
                  const int j = bb * 32 + lane;
                  if (j < n_new) f += min(kcap[j], lim);
                }
              }
              return wsum_ll(f);
            };
            int L, rem;
            if (total <= n) { L = kmax; rem = 0; }
            else if (kmax <= PACK_HIST) {
              // f(L) = sum_j min(k_j, L) = sum_{l <= L} G(l), G(l) = #{k_j >= l}: two warp scans over the histogram
              constexpr int PB = PACK_HIST / 32;
              int c[PB], G[PB];
              int lane_tot = 0;
#pragma unroll
              for (int i = 0; i < PB; ++i) { c[i] = hist[lane * PB + i + 1]; lane_tot += c[i]; }
              int suf = lane_tot;   // inclusive suffix sum over lanes
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_down_sync(0xffffffffu, suf, off);
                if (lane + off < 32) suf += v;
              }
              int run = suf - lane_tot, gsum = 0;   // counts of the lanes above
#pragma unroll
              for (int i = PB - 1; i >= 0; --i) { run += c[i]; G[i] = run; gsum += run; }
              int pre = gsum;       // inclusive prefix sum over lanes
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, off);
                if (lane >= off) pre += v;
              }
              int f = pre - gsum, cnt = 0, fbest = 0;
#pragma unroll
              for (int i = 0; i < PB; ++i) {
                f += G[i];          // f(lane * PB + i + 1)
                if (f <= n) { ++cnt; fbest = f; }
              }
              L = wsum(cnt);        // f is strictly increasing up to kmax and f(kmax) = total > n
              rem = n - wmax(fbest);
            } else {
              int lo = 0, hi = kmax;  // f(lo) <= n < f(hi), f(L) = sum min(k_j, L)
              while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sum_min(mid) <= n) lo = mid; else hi = mid;
              }
              L = lo;
              rem = (int)(n - sum_min(L));
            }
            // cyclic walk from node s over the candidate blocks: block(s) is visited twice (its tail first, its head last)
            int seen = 0, last_dist = -1, newly = 0, got = 0;
            const int bs = s >> 5;
            for (int step = 0; step <= nb; ++step) {
              int bb = bs + step;
              if (bb >= nb) bb -= nb;
              if (step == nb) bb = bs;
              const unsigned cw = (unsigned)cmw[bb >> 5];
              if (!((cw >> (bb & 31)) & 1u)) continue;
              const int j = bb * 32 + lane;
              bool in = j < n_new;
              if (bb == bs) in = in && (step == 0 ? j >= s : (step == nb && j < s));
              const int x = Neff + (j < n_new ? j : 0);
              const int k = in ? kcap[j] : 0;
              const bool extra_c = in && k > L;
              const unsigned m = __ballot_sync(0xffffffffu, extra_c);
              const int rank = seen + __popc(m & ((1u << lane) - 1));
              const int mj = in ? min(k, L) + ((extra_c && rank < rem) ? 1 : 0) : 0;
              seen += __popc(m);
              if (mj > 0) {
#pragma unroll
                for (int a = 0; a < A; ++a) if (req[a] > 0) nfree[(size_t)a * X + x] -= (int64_t)mj * req[a];
                nslots[x] -= mj;
                nports[x] |= pbit;
                if (!nsched[x]) { nsched[x] = 1; newly++; }
                got += mj;
                // the pod placed last sits at the furthest position served in the final lap
                if (rem > 0 ? (extra_c && rank < rem) : (k >= L)) { int dd = j - s; if (dd < 0) dd += n_new; last_dist = dd; }
              }
              if (feeds) log_append(mj > 0, x, spec, mj);
            }
            got = wsum(got);
            newly = wsum(newly);
            last_dist = wmax(last_dist);
            placed += got;
            nodes_with_pods += newly;
            n -= got;
            if (last_dist >= 0) {
              int jl = s + last_dist;
              if (jl >= n_new) jl -= n_new;
              last_index = (N + jl + 1) % list_len;
            }
            __syncwarp();
          }
        }
        if (n > 0 && new_nodes_available) {
          // after the pass above no added node (the last one included) can take this pod any more
          const bool stop = (n_new > 0) && !nsched[Neff + n_new - 1];  // last node still empty (:212)
          if (!stop) {
            int k_new = 0;
            if (static_new) {
              k_new = min(tslots, n);
#pragma unroll
              for (int a = 0; a < A; ++a) {
                if (req[a] > 0 && k_new > 0) {
                  if (tfree[a] < req[a]) k_new = 0;
                  else if (tfree[a] < (int64_t)k_new * req[a]) k_new = (int)(tfree[a] / req[a]);
                }
              }
              if (has_ports) k_new = min(k_new, 1);  // DaemonSet port conflicts are part of static_new
            }
            long long allowed = max_nodes < 0 ? 0 : (max_nodes == 0 ? (long long)INT_MAX : max((long long)max_nodes - n_new, 0ll));
            if (allowed > p.cap - n_new) { allowed = p.cap - n_new; }
            int add, fill = 0;
            if (k_new <= 0) {
              add = allowed >= 1 ? 1 : 0;  // the node is added, the pod still fails on it (:235-240)
              if (allowed < 1) new_nodes_available = false;
            } else {
              const long long need = ((long long)n + k_new - 1) / k_new;
              if (need > allowed) { add = (int)allowed; new_nodes_available = false; }
              else add = (int)need;
              fill = (int)min((long long)n, (long long)add * k_new);
            }
            for (int base = 0; base < add; base += 32) {
              const int i = base + lane;
              const bool in = i < add;
              const int x = Neff + n_new + (in ? i : 0);
              const int mj = (!in || k_new <= 0) ? 0 : min(k_new, fill - i * k_new);
              if (in) {
#pragma unroll
                for (int a = 0; a < A; ++a) nfree[(size_t)a * X + x] = tfree[a] - (req[a] > 0 ? (int64_t)mj * req[a] : 0);
                nslots[x] = tslots - mj;
                nports[x] = mj > 0 ? pbit : 0ull;
                nsched[x] = mj > 0;
                stamp[x] = 0;
              }
              if (feeds) log_append(in && mj > 0, x, spec, mj);
            }
            if (k_new > 0) { nodes_with_pods += add; placed += fill; n -= fill; }
            const int b_first = n_new >> 5;
            n_new += add;
            __syncwarp();
            if (add > 0) {  // loose but valid bounds for the blocks that received nodes; tightened lazily
              for (int b2 = b_first + lane; b2 <= (n_new - 1) >> 5; b2 += 32) {
                const bool fresh_blk = b2 > b_first || (b_first << 5) == n_new - add;
#pragma unroll
                for (int a = 0; a < A; ++a) {
                  const int64_t old = fresh_blk ? LLONG_MIN : bmax[(size_t)a * p.nblk + b2];
                  bmax[(size_t)a * p.nblk + b2] = tfree[a] > old ? tfree[a] : old;
                }
                const int olds = fresh_blk ? INT_MIN : bslots[b2];
                bslots[b2] = tslots > olds ? tslots : olds;
              }
            }
            __syncwarp();
          }
        }
      } else {
        // ======================= dynamic group: per-pod loop on incremental counters ===============
        const bool host_spread = o.ps_hostname_spread[spec] != 0;
        int cur_stamp = ++stamp_ctr;
        const int gver = ++gver_ctr;      // version of this group's working counters (lazy copy-on-write)
        const int n_new_start = n_new;
        int fb_fail_stamp = -1, fb_mark = 0;
        // ---- describe the group's counters ----
        __syncwarp();
        if (lane == 0) {
          int nq = 0;
          for (int q = p.has_dyn ? d.dc_q_off[dc] : 0; q < (p.has_dyn ? d.dc_q_off[dc + 1] : 0); ++q) {
            if (!d.q_active[q]) continue;
            const int k = d.q_k[q], kind = d.q_kind[q];
            wd.qid[nq] = q; wd.kind[nq] = kind; wd.k[nq] = k; wd.host[nq] = d.is_host[k]; wd.Dc[nq] = d.Dc[k];
            const int td = FM ? -1 : d.dom[(size_t)k * NT + N + t];   // FM: no template, nothing is ever added
            wd.tslot[nq] = td < 0 ? -1 : (td < d.Dc[k] ? td : d.Dc[k]);
            wd.wown[nq] = d.q_wown[q]; wd.self[nq] = d.q_self[q];
            wd.maxskew[nq] = kind == Q_PTS ? o.pts_max_skew[d.q_p0[q]] : 0;
            wd.mindom[nq] = kind == Q_PTS ? o.pts_min_domains[d.q_p0[q]] : 0;
            wd.elig_new[nq] = FM ? 0 : d.elig[(size_t)q * p.U + col_new];
            wd.dsw[nq] = FM ? 0 : d.ds_w[(size_t)q * p.T + t];
            wd.tot[nq] = d.base_tot[q];
            wd.boff[nq] = d.q_base_off[q];
            wd.minv[nq] = d.st_min1[q]; wd.nmin[nq] = d.st_nmin[q]; wd.ndom[nq] = d.st_ndom[q];
            ++nq;
          }
          wd.nq = nq;
          wd.aff_self = p.has_dyn ? d.dc_aff_self[dc] : 0;
        }
        __syncwarp();
        const int nq = wd.nq;
        // Working counters are copy-on-write over the cluster base counts: a slot is valid only when its
        // version equals this group's, otherwise it reads as its default (base count for cluster domains,
        // DaemonSet weight for the fresh hostname domain of an added node, 0 for a template-only value).
        auto def_cnt = [&](int q, int sl) -> int {
          const int Dc = wd.Dc[q];
          return sl < Dc ? d.base_cnt[wd.boff[q] + sl] : (sl == Dc ? 0 : (wd.elig_new[q] ? wd.dsw[q] : 0));
        };
        auto def_pres = [&](int q, int sl) -> int {
          const int Dc = wd.Dc[q];
          return sl < Dc ? d.base_pres[wd.boff[q] + sl] : (sl == Dc ? 0 : (wd.elig_new[q] ? 1 : 0));
        };
        auto rd_cnt = [&](int q, int sl) -> int {
          const size_t o2 = (size_t)q * p.dstride + sl;
          if (wver[o2] == gver) return wcnt[o2];
          return def_cnt(q, sl);
        };
        auto rd_pres = [&](int q, int sl) -> int {
          const size_t o2 = (size_t)q * p.dstride + sl;
          if (wver[o2] == gver) return wpres[o2];
          return def_pres(q, sl);
        };
        auto wr = [&](int q, int sl, int c, int pr) {  // single lane
          const size_t o2 = (size_t)q * p.dstride + sl;
          wcnt[o2] = c; wpres[o2] = pr; wver[o2] = gver;
        };
        auto recompute = [&](int q) {  // min / #domains over the present domains of a spread counter
          const int len = wd.Dc[q] + 1 + (wd.host[q] ? n_new : 0);
          int mn = INT_MAX, nd = 0;
          for (int i = lane; i < len; i += 32) if (rd_pres(q, i) > 0) { mn = min(mn, rd_cnt(q, i)); ++nd; }
          mn = wmin(mn);
          nd = wsum(nd);
          int nm = 0;
          for (int i = lane; i < len; i += 32) if (rd_pres(q, i) > 0 && rd_cnt(q, i) == mn) ++nm;
          nm = wsum(nm);
          __syncwarp();
          if (lane == 0) { wd.minv[q] = mn; wd.nmin[q] = nm; wd.ndom[q] = nd; }
          __syncwarp();
        };
        // ---- seed: nodes added so far (O(1) per counter), then the run's placement log if other groups feed us ----
        bool need_log = false;
        for (int q = 0; q < nq; ++q) {
          bool full = false;
          if (wd.elig_new[q] && n_new > 0) {
            const int dsw = wd.dsw[q];
            if (wd.host[q]) {  // n_new fresh hostname domains, each holding the DaemonSet weight
              if (lane == 0) {
                wd.tot[q] += n_new * dsw;
                if (wd.kind[q] == Q_PTS) {
                  wd.ndom[q] += n_new;
                  if (dsw < wd.minv[q]) { wd.minv[q] = dsw; wd.nmin[q] = n_new; }
                  else if (dsw == wd.minv[q]) wd.nmin[q] += n_new;
                }
              }
            } else if (wd.tslot[q] >= 0) {  // all added nodes share the template's value of this key
              const int sl = wd.tslot[q];
              const int c0 = rd_cnt(q, sl), p0 = rd_pres(q, sl), c1 = c0 + n_new * dsw;
              __syncwarp();
              if (lane == 0) {
                wr(q, sl, c1, p0 + n_new);
                wd.tot[q] += n_new * dsw;
                if (wd.kind[q] == Q_PTS && p0 == 0) {
                  wd.ndom[q] += 1;
                  if (c1 < wd.minv[q]) { wd.minv[q] = c1; wd.nmin[q] = 1; }
                  else if (c1 == wd.minv[q]) wd.nmin[q] += 1;
                }
              }
              if (wd.kind[q] == Q_PTS && p0 > 0 && dsw > 0) full = true;
            }
          }
          __syncwarp();
          if (FM || d.q_nfeed[wd.qid[q]] - (wd.wown[q] > 0 ? 1 : 0) > 0) need_log = true;   // FM: earlier runs of this very spec count too
          if (full) recompute(q);
        }
        __syncwarp();
        if (need_log && log_n > 0) {  // pods other groups (FM: earlier runs) placed before: replay the placement log
          const int nlog = min(log_n, p.log_cap);
          for (int q = 0; q < nq; ++q) {
            const int qid = wd.qid[q];
            // lanes stride over the log.  Pass 1 materialises the touched copy-on-write slots with their defaults
            // (identical values from every lane), pass 2 stamps the version and adds the weights atomically.
            auto entry = [&](int i, int& w, size_t& o2) -> bool {
              const int x = logbuf[i * 3];
              w = d.wmat[(size_t)qid * d.S + logbuf[i * 3 + 1]];
              if (w == 0 || !elig_of(q, x)) return false;
              const int sl = slot_of(q, x);
              if (sl < 0) return false;
              o2 = (size_t)q * p.dstride + sl;
              if (wver[o2] != gver) { wcnt[o2] = def_cnt(q, sl); wpres[o2] = def_pres(q, sl); }
              return true;
            };
            bool touched = false;
            int dt = 0;
            for (int i = lane; i < nlog; i += 32) {
              int w; size_t o2;
              if (entry(i, w, o2)) { touched = true; dt += w * logbuf[i * 3 + 2]; }
            }
            __syncwarp();
            for (int i = lane; i < nlog; i += 32) {
              const int x = logbuf[i * 3];
              const int w = d.wmat[(size_t)qid * d.S + logbuf[i * 3 + 1]];
              if (w == 0 || !elig_of(q, x)) continue;
              const int sl = slot_of(q, x);
              if (sl < 0) continue;
              const size_t o2 = (size_t)q * p.dstride + sl;
              wver[o2] = gver;
              atomicAdd(&wcnt[o2], w * logbuf[i * 3 + 2]);
            }
            __syncwarp();
            dt = wsum(dt);
            touched = __ballot_sync(0xffffffffu, touched) != 0u;
            if (lane == 0) wd.tot[q] += dt;
            __syncwarp();
            if (touched && wd.kind[q] == Q_PTS) recompute(q);
          }
          __syncwarp();
        }

        auto ensure_cluster = [&]() {  // run state of the cluster nodes, needed once a fallback can place onto them
          if (cl_init) return;
          for (int x = lane; x < N; x += 32) {
#pragma unroll
            for (int a = 0; a < A; ++a) nfree[(size_t)a * X + x] = p.c_free[(size_t)a * N + x];
            nslots[x] = p.c_slots[x];
            nports[x] = 0ull;
            nsched[x] = 0;
          }
          cl_init = true;
          __syncwarp();
        };
        // RunFilterPlugins on node x (default plugin order), per lane
        auto eval = [&](int x) -> int {
          const int col = x < Neff ? x : col_new;
          const int code = p.pre_code[(size_t)sc * p.U + col] & 0x0F;
          if (code) return code;
          if (nports[x] & pconf) return CAE_R_NODE_PORTS;
          bool fail = nslots[x] < 1;
#pragma unroll
          for (int a = 0; a < A; ++a) fail |= (req[a] > 0 && req[a] > nfree[(size_t)a * X + x]);
          if (fail) return CAE_R_FIT;
          bool aff_any = false, pods_exist = true;
          long long aff_tot = 0;
          for (int q = 0; q < nq; ++q) {
            const int kind = wd.kind[q];
            const int sl = slot_of(q, x);
            const int c = sl >= 0 ? rd_cnt(q, sl) : 0;
            if (kind == Q_PTS) {  // podtopologyspread/filtering.go:314-359
              if (sl < 0) return CAE_R_PTS_MISSING_LABEL;
              const long long minm = wd.ndom[q] < wd.mindom[q] ? 0 : wd.minv[q];
              if ((long long)c + wd.self[q] - minm > wd.maxskew[q]) return CAE_R_PTS_SKEW;
            } else if (kind == Q_AFF) {  // interpodaffinity/filtering.go:382-408
              aff_any = true;
              if (sl < 0) return CAE_R_IPA_AFFINITY;
              if (c <= 0) pods_exist = false;
              aff_tot += wd.tot[q];
            } else {
              if (aff_any) {
                if (!pods_exist && !(aff_tot == 0 && wd.aff_self)) return CAE_R_IPA_AFFINITY;
                aff_any = false;
              }
              if (sl >= 0 && c > 0) return kind == Q_ANTI ? CAE_R_IPA_ANTI_AFFINITY : CAE_R_IPA_EXISTING_ANTI_AFFINITY;
            }
          }
          if (aff_any && !pods_exist && !(aff_tot == 0 && wd.aff_self)) return CAE_R_IPA_AFFINITY;
          return CAE_R_OK;
        };
        // ForceAddPod on node x (uniform x) + counter upkeep
        auto place = [&](int x) {
          const bool newly = !nsched[x];
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int a = 0; a < A; ++a) if (req[a] > 0) nfree[(size_t)a * X + x] -= req[a];
            nslots[x] -= 1;
            nports[x] |= pbit;
            nsched[x] = 1;
          }
          if (newly) ++nodes_with_pods;
          bool bump = false;
          for (int q = 0; q < nq; ++q) {
            const int w = wd.wown[q];
            if (w == 0 || !elig_of(q, x)) continue;
            const int sl = slot_of(q, x);
            if (sl < 0) continue;
            const int old = rd_cnt(q, sl), pr = rd_pres(q, sl);
            __syncwarp();
            if (lane == 0) { wr(q, sl, old + w, pr); wd.tot[q] += w; }
            if (wd.kind[q] == Q_PTS) {
              if (old == wd.minv[q]) {
                const int nm = wd.nmin[q] - 1;
                __syncwarp();
                if (lane == 0) wd.nmin[q] = nm;
                __syncwarp();
                if (nm == 0) { recompute(q); bump = true; }
              }
            } else if (old == 0) bump = true;  // 0 -> positive can make affinity satisfiable elsewhere
            __syncwarp();
          }
          if (feeds) log_append(lane == 0, x, spec, 1);
          if (bump) cur_stamp = ++stamp_ctr;
          ++placed;
          --n;
          __syncwarp();
        };
        // addNewNodeToSnapshot (:249-265): a sanitized copy of the template joins the list
        auto add_node = [&]() {
          const int j = n_new, x = Neff + j;
          if (lane == 0) {
#pragma unroll
            for (int a = 0; a < A; ++a) nfree[(size_t)a * X + x] = tfree[a];
            nslots[x] = tslots;
            nports[x] = 0ull;
            nsched[x] = 0;
            stamp[x] = 0;
            const int b2 = j >> 5;
            const bool first = (j & 31) == 0;
#pragma unroll
            for (int a = 0; a < A; ++a) {
              const int64_t old = bmax[(size_t)a * p.nblk + b2];
              bmax[(size_t)a * p.nblk + b2] = (first || tfree[a] > old) ? tfree[a] : old;
            }
            bslots[b2] = first ? tslots : max(bslots[b2], tslots);
          }
          n_new = j + 1;
          bool bump = false;
          for (int q = 0; q < nq; ++q) {
            const int en = wd.elig_new[q], dsw = wd.dsw[q];
            const int sl = slot_of(q, x);
            const bool pts = wd.kind[q] == Q_PTS;
            if (wd.host[q]) {  // a brand-new hostname domain (its default already reads dsw / present)
              if (en) {
                const int old_nd = wd.ndom[q];
                __syncwarp();
                if (lane == 0) {
                  wd.tot[q] += dsw;
                  if (pts) {
                    wd.ndom[q] = old_nd + 1;
                    if (dsw < wd.minv[q]) { wd.minv[q] = dsw; wd.nmin[q] = 1; }
                    else if (dsw == wd.minv[q]) wd.nmin[q] += 1;
                  }
                }
                __syncwarp();
                // a node joining can only HELP other nodes by lifting the domain count over minDomains
                // (the global minimum stops being treated as 0, filtering.go:55-68)
                if (pts && old_nd < wd.mindom[q] && wd.ndom[q] >= wd.mindom[q]) bump = true;
                if (!pts && dsw > 0) bump = true;
              }
            } else if (en && sl >= 0) {
              const int c0 = rd_cnt(q, sl), p0 = rd_pres(q, sl);
              const int old_nd = wd.ndom[q], old_mn = wd.minv[q];
              __syncwarp();
              if (lane == 0) {
                wr(q, sl, c0 + dsw, p0 + 1);
                wd.tot[q] += dsw;
                if (pts && p0 == 0) {
                  wd.ndom[q] = old_nd + 1;
                  if (c0 + dsw < wd.minv[q]) { wd.minv[q] = c0 + dsw; wd.nmin[q] = 1; }
                  else if (c0 + dsw == wd.minv[q]) wd.nmin[q] += 1;
                }
              }
              __syncwarp();
              if (pts && p0 > 0 && dsw > 0) recompute(q);
              if (pts && ((old_nd < wd.mindom[q] && wd.ndom[q] >= wd.mindom[q]) || wd.minv[q] > old_mn)) bump = true;
              if (!pts && dsw > 0) bump = true;
            }
            __syncwarp();
          }
          if (bump) cur_stamp = ++stamp_ctr;
          __syncwarp();
        };
        (void)n_new_start;

        if (FM) {
          // ---- HintingSimulator.TrySchedulePods over the cluster nodes, pod by pod ----
          ensure_cluster();
          const int len = N;
          const int run_n = n;
          const int cls = p.fm_class ? p.fm_class[p.fm_pods[pb]] : -1;
          bool blocked = cls >= 0 && p.fm_class_mark[cls] != 0;   // IsSimilarUnschedulable (similar_pods.go:68-84)
          bool run_failed = false;  // a pod of this run fitted nowhere: the identical pods behind it see the same state
          for (int i = 0; i < run_n; ++i) {
            const int pod = p.fm_pods[pb + i];
            int where = -1;
            if (!fm_stop) {
              const int h = p.fm_hint ? p.fm_hint[pod] : -1;   // tryScheduleUsingHints (:80-106); lastIndex untouched
              if (h >= 0 && h < N && (!p.fm_node_ok || p.fm_node_ok[h]) && eval(h) == CAE_R_OK) { place(h); where = h; }
              if (where < 0 && !blocked && !run_failed && len > 0) {
                // SchedulePodOnAnyNodeMatching (:117): whole list, cyclic from lastIndex
                int hit = -1;
                for (int base = 0; base < len && hit < 0; base += 32) {
                  const int pos = base + lane;
                  int idx = last_index + pos;
                  if (idx >= len) idx -= len;
                  bool ok = false;
                  if (pos < len && !o.node_unschedulable[idx] && (!p.fm_node_ok || p.fm_node_ok[idx]) && stamp[idx] != cur_stamp) {
                    ok = eval(idx) == CAE_R_OK;
                    if (!ok) stamp[idx] = cur_stamp;
                  }
                  const unsigned m = __ballot_sync(0xffffffffu, ok);
                  if (m) hit = __shfl_sync(0xffffffffu, idx, __ffs(m) - 1);
                }
                if (hit >= 0) {
                  place(hit);
                  last_index = (hit + 1) % len;
                  fm_moved = true;
                  where = hit;
                } else {
                  run_failed = true;
                  if (cls >= 0) {  // SetUnschedulable (similar_pods.go:87-104): at most 10 infos per controller
                    const int ctrl = p.fm_class_ctrl[cls];
                    const int cnt = p.fm_ctrl_cnt[ctrl];
                    __syncwarp();
                    if (lane == 0) {
                      if (cnt >= 10) p.fm_ctrl_over[ctrl] = 1;
                      else { p.fm_ctrl_cnt[ctrl] = cnt + 1; p.fm_class_mark[cls] = 1; }
                    }
                    if (cnt < 10) blocked = true;
                    __syncwarp();
                  }
                }
              }
              if (where < 0 && p.fm_break) fm_stop = true;   // breakOnFailure (:71-73)
            }
            if (lane == 0) p.fm_assigned[pod] = where;
          }
        } else {
        // ---- tryToScheduleOnExistingNodes: per pod, first passing added node in cyclic order ----
        while (n > 0 && n_new > 0) {
          const int s = last_index >= N ? last_index - N : 0;
          const int nb = (n_new + 31) >> 5, bs = s >> 5;
          int found = -1;
          // cyclic order from node s, block by block; block(s) is visited twice (tail first, head last)
          for (int step = 0; step <= nb && found < 0; ++step) {
            int bb = bs + step;
            if (bb >= nb) bb -= nb;
            if (step == nb) bb = bs;
            bool may = bslots[bb] >= 1;  // block bounds: nothing in this block can take the pod (uniform)
#pragma unroll
            for (int a = 0; a < A; ++a) may &= !(req[a] > 0 && bmax[(size_t)a * p.nblk + bb] < req[a]);
            if (!may) continue;
            const int j = bb * 32 + lane;
            bool in = j < n_new;
            if (bb == bs) in = in && (step == 0 ? j >= s : (step == nb && j < s));
            const int x = Neff + (j < n_new ? j : 0);
            bool ok = false;
            if (in && stamp[x] != cur_stamp) {
              ok = eval(x) == CAE_R_OK;
              if (!ok) stamp[x] = cur_stamp;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m) found = __shfl_sync(0xffffffffu, j, __ffs(m) - 1);
          }
          if (found < 0) break;  // first pod that fits nowhere ends this phase for the group (:158)
          place(Neff + found);
          last_index = (N + found + 1) % (N + n_new);
        }
        // ---- tryToScheduleOnNewNodes ----
        while (n > 0 && new_nodes_available) {
          bool found = false;
          if (n_new > 0) {
            const int xl = Neff + n_new - 1;
            const int r = eval(xl);
            if (r == CAE_R_OK) { place(xl); found = true; }
            else if (host_spread && r == CAE_R_PTS_SKEW) {
              // SchedulePodOnAnyNodeMatching(name != lastNodeName) (:190-205): whole list, cyclic from lastIndex
              ensure_cluster();
              const int len = N + n_new, lastpos = N + n_new - 1;
              int hit = -1;
              if (fb_fail_stamp == cur_stamp) {
                // everything scanned before still fails; only nodes added since can pass
                int best = INT_MAX;
                for (int base = fb_mark; base < n_new - 1; base += 32) {
                  const int j = base + lane;
                  int dist = INT_MAX;
                  if (j < n_new - 1 && eval(Neff + j) == CAE_R_OK) { dist = N + j - last_index; if (dist < 0) dist += len; }
                  best = min(best, wmin(dist));
                }
                if (best != INT_MAX) { hit = last_index + best; if (hit >= len) hit -= len; }
              } else {
                for (int base = 0; base < len && hit < 0; base += 32) {
                  const int pos = base + lane;
                  int idx = last_index + pos;
                  if (idx >= len) idx -= len;
                  bool ok = false;
                  if (pos < len && idx != lastpos) {
                    const int x = idx < N ? idx : Neff + (idx - N);
                    if (!(idx < N && o.node_unschedulable[idx]) && stamp[x] != cur_stamp) {  // plugin_runner.go:92-94
                      ok = eval(x) == CAE_R_OK;
                      if (!ok) stamp[x] = cur_stamp;
                    }
                  }
                  const unsigned m = __ballot_sync(0xffffffffu, ok);
                  if (m) hit = __shfl_sync(0xffffffffu, idx, __ffs(m) - 1);
                }
              }
              if (hit >= 0) {
                place(hit < N ? hit : Neff + (hit - N));
                last_index = (hit + 1) % len;
                found = true;
              } else { fb_fail_stamp = cur_stamp; fb_mark = n_new; }
            }
          }
          if (!found) {
            if (n_new > 0 && !nsched[Neff + n_new - 1]) break;  // last node still empty (:212)
            const bool permit = !(max_nodes < 0 || (max_nodes > 0 && n_new >= max_nodes)) && n_new < p.cap;
            if (!permit) { new_nodes_available = false; break; }  // (:222)
            const int stamp_before = cur_stamp;
            add_node();
            if (fb_fail_stamp == stamp_before && cur_stamp != stamp_before) fb_fail_stamp = -1;
            if (eval(Neff + n_new - 1) != CAE_R_OK) break;  // (:238-240)
            place(Neff + n_new - 1);
          }
        }
        }  // !FM
      }
      pods_total += placed;
      if (!FM && lane == 0 && p.sched) p.sched[(size_t)t * p.E + g] = placed;
    }
    if (FM) {
      int over = 0;
      for (int c = lane; c < p.fm_nctrl; c += 32) over += p.fm_ctrl_over[c] != 0;
      over = wsum(over);
      if (lane == 0) {
        hdr[0] = stamp_ctr; hdr[1] = gver_ctr;
        p.fm_out[0] = last_index; p.fm_out[1] = over; p.fm_out[2] = pods_total; p.fm_out[3] = fm_moved ? 1 : 0;
        if (overflow && p.status) atomicExch(p.status, 1);
      }
    } else if (lane == 0) {
      hdr[0] = stamp_ctr; hdr[1] = gver_ctr;
      p.node_count[t] = nodes_with_pods;
      p.pod_count[t] = pods_total;
      if (overflow && p.status) atomicExch(p.status, 1);
    }
    __syncwarp();
  }
}

template <int A, bool FM>
static void launch_pack_a(int blocks, cudaStream_t st, const DevObjects& o, const DynTables& d, const PackParams& p) {
  pack_kernel<A, FM><<<blocks, PACK_WARPS * 32, 0, st>>>(o, d, p);
}

template <bool FM>
static void launch_pack_any(int A, int blocks, cudaStream_t st, const DevObjects& o, const DynTables& d, const PackParams& p) {
  switch (A) {
    case 0: launch_pack_a<0, FM>(blocks, st, o, d, p); break;
    case 1: launch_pack_a<1, FM>(blocks, st, o, d, p); break;
    case 2: launch_pack_a<2, FM>(blocks, st, o, d, p); break;
    case 3: launch_pack_a<3, FM>(blocks, st, o, d, p); break;
    case 4: launch_pack_a<4, FM>(blocks, st, o, d, p); break;
    case 5: launch_pack_a<5, FM>(blocks, st, o, d, p); break;
    case 6: launch_pack_a<6, FM>(blocks, st, o, d, p); break;
    case 7: launch_pack_a<7, FM>(blocks, st, o, d, p); break;
    default: launch_pack_a<8, FM>(blocks, st, o, d, p); break;
  }
}

// slab layout shared by the estimator and the filter pass: returns bytes per warp, fills the layout fields of p
static size_t pack_layout(const Engine* e, PackParams& p, int cap, bool cluster_state, int log_cap) {
  p.cap = cap;
  const int Neff = cluster_state ? e->N : 0;
  const size_t X = (size_t)Neff + cap;
  int dmax = 1;
  for (int k = 0; k < e->dyn.K; ++k) dmax = std::max(dmax, e->dyn.Dc[k] + 1 + (e->dyn.is_host[k] ? cap : 0));
  p.dstride = p.has_dyn ? dmax : 1;
  p.log_cap = log_cap;
  const int A1 = std::max(e->A, 1);
  p.nblk = (cap + 31) / 32 + 1;
  size_t per_warp = 16 + X * ((size_t)A1 * 8 + 8 + 4 + 4 + 4) + (size_t)3 * DYN_MAX_Q * p.dstride * 4 + (size_t)p.log_cap * 12 + (size_t)(p.nblk + p.nblk / 32 + 2) * 4;
  per_warp = (per_warp + 7) & ~(size_t)7;
  p.bmax_off = per_warp;
  per_warp += (size_t)A1 * p.nblk * 8 + X;
  per_warp = (per_warp + 255) & ~(size_t)255;
  return per_warp;
}

static void pack_common(const Engine* e, PackParams& p) {
  p.E = e->E; p.T = e->T; p.N = e->N; p.U = e->U;
  p.has_dyn = e->has_dynamic ? 1 : 0;
  for (int a = 0; a < CAE_MAX_RES; ++a) p.act_dim[a] = e->act_dim[a];
  p.order = e->d_order; p.order_n = e->d_order_n; p.pre_code = e->d_pre_code; p.spec_sc = e->d_spec_sc; p.spec_dc = e->d_spec_dc;
  p.tmpl_free = e->d_tmpl_free; p.tmpl_slots = e->d_tmpl_slots; p.max_nodes = e->d_max_nodes;
  p.pc_of = e->d_pc_of; p.port_conf = e->d_port_conf; p.c_free = e->d_c_free; p.c_slots = e->d_c_slots;
  p.node_count = e->d_counts2; p.pod_count = e->d_counts2 + e->T; p.sched = e->d_sched;
  p.work_counter = e->d_work_counter; p.status = e->d_work_counter + 1;
}

int launch_pack(Engine* e) {
  int nt = e->t_end - e->t_begin;
  if (nt <= 0) return 0;
  PackParams p{};
  pack_common(e, p);
  p.t_begin = e->t_begin; p.t_end = e->t_end;
  // node capacity of a slab: the largest limiter cap, or (unlimited) one node per pod + 1
  // (every added node but possibly one holds >= 1 pod)
  const int cap = std::max(1, std::min(e->P + 1, e->pack_cap));
  const size_t X = (size_t)(p.has_dyn ? e->N : 0) + cap;
  const int log_cap = p.has_dyn ? (int)std::min<size_t>(4 * X + 1024, (size_t)1 << 24) : 1;
  const size_t per_warp = pack_layout(e, p, cap, p.has_dyn != 0, log_cap);
  const int A1 = std::max(e->A, 1);
  int warps = std::min(nt, e->sm_count * e->pack_warps_per_sm);
  const size_t budget = (size_t)24 << 30;  // keep the slabs within 24 GiB of the 180 GB HBM
  if (per_warp * warps > budget) warps = (int)std::max<size_t>(1, budget / per_warp);
  int blocks = (warps + PACK_WARPS - 1) / PACK_WARPS;
  warps = blocks * PACK_WARPS;
  size_t need = per_warp * warps;
  const size_t sig = per_warp * 1000003u + X * 10007u + (size_t)p.dstride * 101u + (size_t)p.log_cap * 7u + (size_t)A1;
  if (need > e->pack_scratch_bytes) {
    if (e->d_pack_scratch) cudaFree(e->d_pack_scratch);
    e->d_pack_scratch = nullptr;
    e->pack_scratch_bytes = 0;
    CAE_CUDA(cudaMalloc(&e->d_pack_scratch, need));
    e->pack_scratch_bytes = need;
    e->pack_layout_sig = 0;
  }
  if (sig != e->pack_layout_sig) {  // stamps / versions are only meaningful within one slab layout
    CAE_CUDA(cudaMemsetAsync(e->d_pack_scratch, 0, need, e->stream));
    e->pack_layout_sig = sig;
  }
  p.scratch = static_cast<unsigned char*>(e->d_pack_scratch);
  p.scratch_per_warp = per_warp;
  CAE_CUDA(cudaMemsetAsync(e->d_work_counter, 0, sizeof(int32_t) * 2, e->stream));
  if (e->pack_lpt) {   // longest processing time first: fewer idle SMs while the last templates finish
    std::vector<long long> cost(nt);
    CAE_CUDA(cudaMemcpyAsync(cost.data(), e->d_tmpl_cost + e->t_begin, sizeof(long long) * nt, cudaMemcpyDeviceToHost, e->stream));
    CAE_CUDA(cudaStreamSynchronize(e->stream));
    std::vector<int32_t> perm(nt);
    for (int i = 0; i < nt; ++i) perm[i] = e->t_begin + i;
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return cost[a - e->t_begin] > cost[b - e->t_begin]; });
    CAE_CUDA(cudaMemcpyAsync(e->d_perm, perm.data(), sizeof(int32_t) * nt, cudaMemcpyHostToDevice, e->stream));
    CAE_CUDA(cudaStreamSynchronize(e->stream));   // perm lives on this stack frame
    p.perm = e->d_perm;
  }
  launch_pack_any<false>(e->A, blocks, e->stream, e->dobj, e->dyn, p);
  e->stats.kernel_launches++;
  CAE_KERNEL_OK();
  return 0;
}

// HintingSimulator.TrySchedulePods on the cluster snapshot (one warp).  `in` = device blob laid out by
// cae_filter_schedulable (api.cu); scratch (class marks, controller counters) is zeroed here.
int launch_filter(Engine* e, const FilterLaunch& f) {
  PackParams p{};
  pack_common(e, p);
  p.t_begin = 0; p.t_end = 1;
  p.fm_runs = f.runs; p.fm_last_index = f.last_index; p.fm_break = f.break_on_failure;
  p.fm_run_off = f.run_off; p.fm_pods = f.pods; p.fm_hint = f.hint; p.fm_class = f.cls; p.fm_class_ctrl = f.class_ctrl;
  p.fm_node_ok = f.node_ok; p.fm_assigned = f.assigned; p.fm_out = f.out;
  p.fm_ctrl_cnt = f.ctrl_cnt; p.fm_class_mark = f.class_mark; p.fm_ctrl_over = f.ctrl_over; p.fm_nctrl = f.nctrl;
  const int log_cap = p.has_dyn ? f.n_pods + 1024 : 1;   // one entry per placement at most
  const size_t per_warp = pack_layout(e, p, 1, true, log_cap);
  const size_t need = per_warp * PACK_WARPS;
  const size_t sig = per_warp * 1000003u + (size_t)e->N * 10007u + (size_t)p.dstride * 101u + (size_t)p.log_cap * 7u + (size_t)std::max(e->A, 1);
  if (need > e->fm_scratch_bytes) {
    if (e->d_fm_scratch) cudaFree(e->d_fm_scratch);
    e->d_fm_scratch = nullptr;
    e->fm_scratch_bytes = 0;
    CAE_CUDA(cudaMalloc(&e->d_fm_scratch, need));
    e->fm_scratch_bytes = need;
    e->fm_layout_sig = 0;
  }
  if (sig != e->fm_layout_sig) {
    CAE_CUDA(cudaMemsetAsync(e->d_fm_scratch, 0, need, e->stream));
    e->fm_layout_sig = sig;
  }
  p.scratch = static_cast<unsigned char*>(e->d_fm_scratch);
  p.scratch_per_warp = per_warp;
  CAE_CUDA(cudaMemsetAsync(e->d_work_counter, 0, sizeof(int32_t) * 2, e->stream));
  launch_pack_any<true>(e->A, 1, e->stream, e->dobj, e->dyn, p);
  e->stats.kernel_launches++;
  CAE_KERNEL_OK();
  return 0;
}

}  // namespace cae
