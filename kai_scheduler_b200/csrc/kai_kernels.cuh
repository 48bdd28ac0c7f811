// kai_kernels.cuh — sm_100a kernels of the scheduling-cycle engine.
//
//   k_node_totals    Σ node Allocatable over ready nodes          (proportion.go:252-288)
//   k_queue_usage    per-queue Allocated / Request scatter-add    (proportion.go:347-401)
//   k_fair_share     hierarchical fair-share division per level   (resource_division.go:26-357)
//   k_action         persistent cooperative kernel running a whole Action (allocate) on device:
//                    node tiles resident in shared memory, one fit+score+argmax sweep per
//                    allocateTask, one all-to-all slot exchange per sweep, replicated sequencer.
//
// All arithmetic that feeds a decision is IEEE binary64 with explicit round-to-nearest
// intrinsics (no FMA contraction; the file is also compiled with -fmad=false) in the
// reference's operation order.
#pragma once
#include <cfloat>
#include <cstdint>

#include "kai_device.cuh"
#include "kai_seq.cuh"

namespace kai {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_relaxed_b128(void *p, unsigned long long lo, unsigned long long hi) {
  asm volatile("{ .reg .b128 q; mov.b128 q, {%1, %2}; st.relaxed.gpu.global.b128 [%0], q; }" ::"l"(p), "l"(lo),
               "l"(hi)
               : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_b128(void *p, unsigned long long lo, unsigned long long hi) {
  asm volatile("{ .reg .b128 q; mov.b128 q, {%1, %2}; st.relaxed.sys.global.b128 [%0], q; }" ::"l"(p), "l"(lo),
               "l"(hi)
               : "memory");
}
__device__ __forceinline__ void ld_relaxed_sys_b128(const void *p, unsigned long long &lo, unsigned long long &hi) {
  asm volatile("{ .reg .b128 q; ld.relaxed.sys.global.b128 q, [%2]; mov.b128 {%0, %1}, q; }"
               : "=l"(lo), "=l"(hi)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void ld_relaxed_b128(const void *p, unsigned long long &lo, unsigned long long &hi) {
  asm volatile("{ .reg .b128 q; ld.relaxed.gpu.global.b128 q, [%2]; mov.b128 {%0, %1}, q; }"
               : "=l"(lo), "=l"(hi)
               : "l"(p)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// K_totals: proportion.setTotalResources (proportion.go:252-288)
// ---------------------------------------------------------------------------------------------
__global__ void k_node_totals(DevSnap s) {
  double acc[QR] = {0, 0, 0};
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < s.N; n += gridDim.x * blockDim.x) {
    if (!(s.nflags[n] & KAI_NODE_READY)) continue;
    for (int r = 0; r < QR; r++) {
      double v = s.alloc[(size_t)r * s.N + n];
      if (s.foreign) v = __dsub_rn(v, s.foreign[(size_t)r * s.N + n]);
      acc[r] = __dadd_rn(acc[r], v);
    }
  }
  __shared__ double sh[QR][32];
  for (int r = 0; r < QR; r++) {
    double v = acc[r];
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[r][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < QR) {
    double v = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; w++) v = __dadd_rn(v, sh[threadIdx.x][w]);
    atomicAdd(&s.total[threadIdx.x], v);
  }
}

// ---------------------------------------------------------------------------------------------
// K_usage: proportion.updateQueuesCurrentResourceUsage (proportion.go:347-401)
// ---------------------------------------------------------------------------------------------
__global__ void k_queue_usage(DevSnap s) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < s.T; t += gridDim.x * blockDim.x) {
    int st = s.t_status[t];
    bool allocated = (st & kAllocatedStatuses) != 0;
    if (!allocated && st != KAI_POD_PENDING) continue;
    int j = s.t_job[t];
    bool preemptible = (s.j_flags[j] & KAI_JOB_PREEMPTIBLE) != 0;
    for (int q = s.j_queue[j]; q >= 0; q = s.q_parent[q])
      for (int r = 0; r < QR; r++) {
        double v = s.t_req[(size_t)t * s.R + r];
        if (v == 0) continue;
        atomicAdd(&s.q_request[(size_t)r * s.Q + q], v);
        if (allocated) {
          atomicAdd(&s.q_alloc[(size_t)r * s.Q + q], v);
          if (!preemptible) atomicAdd(&s.q_alloc_np[(size_t)r * s.Q + q], v);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// K_fairshare: resource_division.SetResourcesShare for every sibling group of one tree level.
// One thread per (group, resource); groups of a level are independent, resources are independent.
// scratch: w [3][Q], rr [3][Q] (NaN = absent)
// ---------------------------------------------------------------------------------------------
struct FsGroup {
  const int *members;
  int n;
};
__device__ __forceinline__ FsGroup fs_group(const DevSnap &s, int parent) {
  FsGroup g;
  if (parent < 0) {
    g.members = s.top_queues;
    g.n = s.n_top;
  } else {
    g.members = s.q_children + s.q_child_begin[parent];
    g.n = s.q_child_begin[parent + 1] - s.q_child_begin[parent];
  }
  return g;
}

__device__ void fs_set_resource_share(const DevSnap &s, FsGroup g, int r, double total, double k, double *w,
                                      double *rr) {
  const size_t ro = (size_t)r * s.Q;
  const double *deserved = s.q_deserved + ro, *limit = s.q_limit + ro, *oqw = s.q_oqw + ro;
  const double *request = s.q_request + ro;
  const double *usage = s.q_usage ? s.q_usage + ro : nullptr;
  double *fair = s.q_fair + ro;
  w += ro;
  rr += ro;
  auto satisfied = [&](int q) {  // resource_division.go:283-290
    if (request[q] <= fair[q]) return true;
    if (limit[q] != KAI_UNLIMITED && limit[q] <= fair[q]) return true;
    return false;
  };
  auto remaining_requested = [&](int q) {  // :317-325
    double requested = requestable_share(limit[q], request[q]);
    if (requested < fair[q]) return 0.0;
    return __dsub_rn(requested, fair[q]);
  };
  // :92-109 setDeservedResource
  double remaining = total;
  for (int i = 0; i < g.n; i++) {
    int q = g.members[i];
    double d = deserved[q];
    if (d == KAI_UNLIMITED) d = total;
    double amount = fmin(d, requestable_share(limit[q], request[q]));
    fair[q] = __dadd_rn(fair[q], amount);
    remaining = __dsub_rn(remaining, amount);
    rr[q] = __longlong_as_double(0x7ff8000000000000LL);
  }
  if (!(remaining > 0)) return;
  // :111-144 divideOverQuotaResource — priorities in descending order (:146-162)
  for (int pass = 0; pass < 2; pass++) {
    long long cur_prio = 0x7fffffffffLL;
    for (;;) {
      long long p = -0x7fffffffffLL;
      bool any = false;
      for (int i = 0; i < g.n; i++) {
        long long qp = s.q_priority[g.members[i]];
        if (qp < cur_prio && (!any || qp > p)) {
          p = qp;
          any = true;
        }
      }
      if (!any) break;
      cur_prio = p;
      if (pass == 0) {
        // :164-222 divideUpToFairShare over the queues of this priority
        for (;;) {
          bool another = false;
          double round_amount = remaining;
          double total_weights = 0;  // :307-315
          for (int i = 0; i < g.n; i++) {
            int q = g.members[i];
            if (s.q_priority[q] != p) continue;
            if (remaining_requested(q) > 0) total_weights = __dadd_rn(total_weights, oqw[q]);
          }
          double wsum = 0.0;
          if (total_weights != 0) {  // :224-251 calcShareWeights
            for (int i = 0; i < g.n; i++) {
              int q = g.members[i];
              if (s.q_priority[q] != p) continue;
              w[q] = 0.0;
              if (satisfied(q)) continue;
              double n_weight = __ddiv_rn(oqw[q], total_weights);
              double n_usage = usage ? usage[q] : 0.0;
              double t = __dsub_rn(n_weight, n_usage);
              double t2 = __dmul_rn(k, t);
              double sw = fmax(0.0, __dadd_rn(n_weight, t2));
              w[q] = sw;
              wsum = __dadd_rn(wsum, sw);
            }
          }
          if (wsum == 0) break;
          for (int i = 0; i < g.n; i++) {
            int q = g.members[i];
            if (s.q_priority[q] != p) continue;
            if (remaining == 0) break;
            if (satisfied(q)) continue;
            double requested = remaining_requested(q);
            if (oqw[q] == 0) continue;
            double nqw = __ddiv_rn(w[q], wsum);
            double fs = __dmul_rn(round_amount, nqw);
            double give = 0;  // :264-281 getResourceToGiveInCurrentRound
            if (requested <= fs) {
              give = requested;
              rr[q] = __longlong_as_double(0x7ff8000000000000LL);
            } else {
              double rf = floor(fs);
              if (rf > 0) give = rf;
              double left = __dsub_rn(fs, give);
              if (left > 0) rr[q] = left;
            }
            if (give == 0) continue;
            fair[q] = __dadd_rn(fair[q], give);
            remaining = __dsub_rn(remaining, give);
            another = another || requested < fs;
          }
          if (!another || remaining == 0) break;
        }
      } else {
        if (remaining <= 0) break;
        // :253-262 divideRemainingResource, order :335-357 (remaining desc, creation asc, uid asc)
        for (;;) {
          if (remaining == 0) break;
          int best = -1;
          for (int i = 0; i < g.n; i++) {
            int q = g.members[i];
            if (s.q_priority[q] != p) continue;
            double a = rr[q];
            if (a != a) continue;
            if (best < 0) {
              best = q;
              continue;
            }
            double b = rr[best];
            bool better;
            if (a > b)
              better = true;
            else if (a < b)
              better = false;
            else if (s.q_creation[q] != s.q_creation[best])
              better = s.q_creation[q] < s.q_creation[best];
            else
              better = s.q_uid_rank[q] < s.q_uid_rank[best];
            if (better) best = q;
          }
          if (best < 0) break;
          rr[best] = __longlong_as_double(0x7ff8000000000000LL);
          double give = fmin(1.0, remaining);
          fair[best] = __dadd_rn(fair[best], give);
          remaining = __dsub_rn(remaining, give);
        }
      }
    }
  }
}

// single CTA; levels separated by __syncthreads (proportion.go:410-423 setFairShareForQueues)
__global__ void k_fair_share(DevSnap s, double k_value, double *w, double *rr) {
  for (int lvl = 0; lvl < s.n_levels; lvl++) {
    int g0 = s.level_group_begin[lvl], g1 = s.level_group_begin[lvl + 1];
    int n_items = (g1 - g0) * QR;
    for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
      int parent = s.level_groups[g0 + it / QR];
      int r = it % QR;
      double total = parent < 0 ? s.total[r] : s.q_fair[(size_t)r * s.Q + parent];
      fs_set_resource_share(s, fs_group(s, parent), r, total, k_value, w, rr);
    }
    __syncthreads();
  }
  // GetAllocatableShare per queue (resource_share.go:51-61): static for the rest of the cycle
  for (int i = threadIdx.x; i < QR * s.Q; i += blockDim.x)
    s.q_allocatable[i] = allocatable_share(s.q_deserved[i], s.q_fair[i], s.q_limit[i]);
}

}  // namespace kai
