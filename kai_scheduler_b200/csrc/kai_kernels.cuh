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

namespace kai {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_relaxed_b128(void *p, unsigned long long lo, unsigned long long hi) {
  asm volatile("{ .reg .b128 q; mov.b128 q, {%1, %2}; st.relaxed.gpu.global.b128 [%0], q; }" ::"l"(p), "l"(lo),
               "l"(hi)
               : "memory");
}
__device__ __forceinline__ void ld_relaxed_b128(const void *p, unsigned long long &lo, unsigned long long &hi) {
  asm volatile("{ .reg .b128 q; ld.relaxed.gpu.global.b128 q, [%2]; mov.b128 {%0, %1}, q; }"
               : "=l"(lo), "=l"(hi)
               : "l"(p)
               : "memory");
}

__device__ __forceinline__ double requestable_share(double max_allowed, double request) {
  if (max_allowed == KAI_UNLIMITED) return request;
  return fmin(max_allowed, request);
}
// resource_share.go:51-61
__device__ __forceinline__ double allocatable_share(double deserved, double fair, double max_allowed) {
  if (deserved == KAI_UNLIMITED) return max_allowed;
  double a = fmax(deserved, fair);
  if (max_allowed != KAI_UNLIMITED) a = fmin(max_allowed, a);
  return a;
}
// resource_quantities.go:81-97
__device__ __forceinline__ int compare_quantities(double q, double o) {
  if (q == KAI_UNLIMITED) return o == KAI_UNLIMITED ? 0 : 1;
  if (o == KAI_UNLIMITED) return -1;
  if (q > o) return 1;
  if (q < o) return -1;
  return 0;
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
}

// ---------------------------------------------------------------------------------------------
// K_action: the persistent action kernel
// ---------------------------------------------------------------------------------------------
struct Track {  // global min/max of NonAllocated(res) over nodes with Allocatable(res) != 0 (pack.go:66-86)
  double mn, mx;
  int cnt_mn, cnt_mx;
  int dirty;
};

struct Decision {
  double req[KAI_MAX_RES];
  double mn, mx;
  int task, res, strategy, gpu_task, pipeline_only, nominated, pred_class, best_effort;
};

struct Winner {
  double score;
  uint32_t rank;
  uint32_t flags;
  int node;
};
// winner flag bits (per resource: gpu bits 0-2, cpu bits 3-5) + bit 6 fits idle
enum { WF_B_EQ_MX = 1, WF_A_EQ_MN = 2, WF_A_LT_MN = 4, WF_FITS_IDLE = 64, WF_HAS_RES0 = 128 };

struct Ctl {  // broadcast block, written by thread 0
  int job, n_items, job_ok, item_ok, need_minmax, stop;
  unsigned int seq;
  Decision dec;
  Winner win;
  Track trk[2];  // 0 gpu, 1 cpu
};

struct Tile {  // shared-memory node tile of this CTA
  double *I, *L;        // [R][npc]
  double *Agpu, *Acpu;  // [npc]
  double *gpu_count;    // [npc]
  int *rank;            // [npc]
  uint32_t *flags;      // [npc]
  int npc, base, count, R;
};

struct Seq {  // replicated sequencer state (thread 0 of every CTA)
  const DevSnap *s;
  const kai_config *cfg;
  Replica rp;
  Tile *tile;
  Ctl *ctl;
  int n_ops, ops_cap;
  int root_len;
  int n_items;
  bool is_cta0;
  long long pods_placed, pods_evicted, decisions, nodes_scanned, n_visits, minmax_exchanges;
  kai_job_visit *visits;
  int visits_cap;
  int error;
  long long t_pop, t_prep, t_scan, t_xchg, t_apply, t_finish, t_init;  // clock64 phase totals (thread 0)
};

__device__ __forceinline__ double &q_alloc(Seq &q, int r, int qi) { return q.rp.q_alloc[(size_t)r * q.s->Q + qi]; }
__device__ __forceinline__ double &q_alloc_np(Seq &q, int r, int qi) {
  return q.rp.q_alloc_np[(size_t)r * q.s->Q + qi];
}

// ---- PodInfo helpers ----
__device__ __forceinline__ bool should_allocate(const Seq &q, int t, bool real) {  // pod_info.go:518-521
  int st = q.rp.t_status[t];
  return st == KAI_POD_PENDING || (!real && st == KAI_POD_RELEASING && q.rp.t_virtual[t]);
}

// ---- node tile mutation by the owning CTA (node_info.go:457-551) ----
__device__ void node_add_task(Seq &q, int t) {
  const DevSnap &s = *q.s;
  int n = q.rp.t_node[t];
  int st = q.rp.t_status[t];
  q.rp.t_node_status[t] = st;
  Tile &tl = *q.tile;
  int ln = n - tl.base;
  if (ln < 0 || ln >= tl.count) return;
  for (int r = 0; r < s.R; r++) {
    double v = s.t_req[(size_t)t * s.R + r];
    double &I = tl.I[r * tl.npc + ln], &L = tl.L[r * tl.npc + ln];
    if (st == KAI_POD_RELEASING) {
      L = __dadd_rn(L, v);
      I = __dsub_rn(I, v);
    } else if (st == KAI_POD_PIPELINED) {
      L = __dsub_rn(L, v);
    } else {
      I = __dsub_rn(I, v);
    }
  }
}
__device__ void node_remove_task(Seq &q, int t, int n) {
  const DevSnap &s = *q.s;
  int st = q.rp.t_node_status[t];
  Tile &tl = *q.tile;
  int ln = n - tl.base;
  if (ln < 0 || ln >= tl.count) return;
  for (int r = 0; r < s.R; r++) {
    double v = s.t_req[(size_t)t * s.R + r];
    double &I = tl.I[r * tl.npc + ln], &L = tl.L[r * tl.npc + ln];
    if (st == KAI_POD_RELEASING) {
      L = __dsub_rn(L, v);
      I = __dadd_rn(I, v);
    } else if (st == KAI_POD_PIPELINED) {
      L = __dadd_rn(L, v);
    } else {
      I = __dadd_rn(I, v);
    }
  }
}

// ---- PodGroupInfo.UpdateTaskStatus (job_info.go:253-264) + podset counters ----
__device__ void set_status(Seq &q, int t, int status) {
  int old = q.rp.t_status[t];
  int ps = q.s->t_podset[t];
  if (old & kActiveAllocated) q.rp.ps_active_alloc[ps]--;
  if (status & kActiveAllocated) q.rp.ps_active_alloc[ps]++;
  q.rp.t_status[t] = status;
  q.rp.j_req_valid[q.s->t_job[t]] = 0;
}

// ---- proportion event handlers (proportion.go:443-489) ----
__device__ void queue_allocate(Seq &q, int t, bool add) {
  const DevSnap &s = *q.s;
  int j = s.t_job[t];
  bool preemptible = (s.j_flags[j] & KAI_JOB_PREEMPTIBLE) != 0;
  for (int qi = s.j_queue[j]; qi >= 0; qi = s.q_parent[qi])
    for (int r = 0; r < QR; r++) {
      double v = s.t_req[(size_t)t * s.R + r];
      double &a = q_alloc(q, r, qi);
      a = add ? __dadd_rn(a, v) : __dsub_rn(a, v);
      if (!preemptible) {
        double &b = q_alloc_np(q, r, qi);
        b = add ? __dadd_rn(b, v) : __dsub_rn(b, v);
      }
    }
}

// ---- Statement (framework/statement.go) ----
__device__ void push_op(Seq &q, const Op &op) {
  if (q.n_ops >= q.ops_cap) {
    q.error = 1;
    return;
  }
  q.rp.ops[q.n_ops++] = op;
}
__device__ void stmt_allocate(Seq &q, int t, int n) {  // :297-358
  Op op;
  op.kind = OP_ALLOCATE;
  op.task = t;
  op.prev_status = q.rp.t_status[t];
  op.prev_node = q.rp.t_node[t];
  op.next_node = n;
  op.prev_virtual = q.rp.t_virtual[t];
  op.undo_index = -1;
  op.pad = 0;
  set_status(q, t, KAI_POD_ALLOCATED);
  q.rp.t_node[t] = n;
  node_add_task(q, t);
  queue_allocate(q, t, true);
  push_op(q, op);
  q.rp.t_virtual[t] = 1;
}
__device__ void unallocate(Seq &q, int t, int prev_virtual) {  // :392-427
  set_status(q, t, KAI_POD_PENDING);
  node_remove_task(q, t, q.rp.t_node[t]);
  q.rp.t_node[t] = -1;
  q.rp.t_virtual[t] = (unsigned char)prev_virtual;
  queue_allocate(q, t, false);
}
__device__ void stmt_pipeline(Seq &q, int t, int n) {  // :197-295 (task not yet on the node)
  Op op;
  op.kind = OP_PIPELINE;
  op.task = t;
  op.prev_status = q.rp.t_status[t];
  op.prev_node = q.rp.t_node[t];
  op.next_node = n;
  op.prev_virtual = q.rp.t_virtual[t];
  op.undo_index = -1;
  op.pad = 0;
  set_status(q, t, KAI_POD_PIPELINED);
  q.rp.t_node[t] = n;
  node_add_task(q, t);
  queue_allocate(q, t, true);
  push_op(q, op);
  q.rp.t_virtual[t] = 1;
}
__device__ void unpipeline(Seq &q, const Op &op) {  // :432-476
  int t = op.task;
  set_status(q, t, op.prev_status);
  int host = q.rp.t_node[t];
  q.rp.t_node[t] = op.prev_node;
  q.rp.t_virtual[t] = (unsigned char)op.prev_virtual;
  node_remove_task(q, t, host);
  queue_allocate(q, t, false);
}
__device__ void undo_op(Seq &q, int i) {  // :597-643 (allocate-action subset: no undo chains survive)
  Op op = q.rp.ops[i];
  if (op.kind == OP_ALLOCATE)
    unallocate(q, op.task, op.prev_virtual);
  else if (op.kind == OP_PIPELINE)
    unpipeline(q, op);
  q.ctl->trk[0].dirty = q.ctl->trk[1].dirty = 1;
}
__device__ void stmt_rollback(Seq &q, int cp) {  // :48-61
  for (int i = q.n_ops - 1; i >= cp; i--) undo_op(q, i);
  q.n_ops = cp;
}
__device__ void stmt_convert_all_allocated_to_pipelined(Seq &q, int job) {  // :483-520
  int n0 = q.n_ops;
  for (int i = 0; i < n0; i++) {
    Op op = q.rp.ops[i];
    if (op.kind != OP_ALLOCATE || q.s->t_job[op.task] != job) continue;
    int node = q.rp.t_node[op.task];
    unallocate(q, op.task, 1);
    stmt_pipeline(q, op.task, node);
  }
  int k = 0;
  for (int i = 0; i < q.n_ops; i++) {
    Op op = q.rp.ops[i];
    if (op.kind == OP_ALLOCATE && q.s->t_job[op.task] == job) continue;
    q.rp.ops[k++] = op;
  }
  q.n_ops = k;
  q.ctl->trk[0].dirty = q.ctl->trk[1].dirty = 1;
}
__device__ void stmt_commit(Seq &q) {  // :536-571
  for (int i = 0; i < q.n_ops; i++) {
    Op op = q.rp.ops[i];
    if (op.kind == OP_ALLOCATE) {
      // BindPod -> updatePodOnSession(Binding) (session.go:111-125): active-allocated -> active-allocated
      q.rp.t_status[op.task] = KAI_POD_BINDING;
      q.rp.t_node_status[op.task] = KAI_POD_BINDING;
      q.rp.j_req_valid[q.s->t_job[op.task]] = 0;
      q.pods_placed++;
    } else if (op.kind == OP_PIPELINE) {
      q.pods_placed++;
    } else if (op.kind == OP_EVICT) {
      q.pods_evicted++;
    }
  }
  q.n_ops = 0;
}

// ---- podset / task selection (api/podgroup_info/allocation_info.go) ----
__device__ bool podset_less(const Seq &q, int a, int b) {  // subgroup_order.go:31-62, name order = index order
  int ln = q.rp.ps_active_alloc[a], rn = q.rp.ps_active_alloc[b];
  int lm = q.s->ps_min[a], rm = q.s->ps_min[b];
  bool lsat = ln >= lm, rsat = rn >= rm;
  if (!lsat && !rsat) return a < b;
  if (!lsat) return true;
  if (!rsat) return false;
  double lr = __ddiv_rn((double)ln, (double)lm);
  double rr = __ddiv_rn((double)rn, (double)rm);
  if (lr < rr) return true;
  if (rr < lr) return false;
  return a < b;
}
__device__ int sorted_podsets(const Seq &q, int job, int *out) {
  int b = q.s->j_ps_begin[job], e = q.s->j_ps_begin[job + 1];
  int n = 0;
  for (int ps = b; ps < e; ps++) {  // insertion sort with the PodSetOrderFn total order
    int i = n++;
    while (i > 0 && podset_less(q, ps, out[i - 1])) {
      out[i] = out[i - 1];
      i--;
    }
    out[i] = ps;
  }
  return n;
}
// :27-54 GetTasksToAllocate; result into q.rp.tta, returns count.  If sum != null also accumulates
// the request of the selected tasks (GetTasksToAllocateInitResource :87-113).
__device__ int tasks_to_allocate(Seq &q, int job, bool real, double *sum) {
  const DevSnap &s = *q.s;
  int *order = q.rp.ps_order;
  int nps = sorted_podsets(q, job, order);
  int unsat = 0;
  for (int k = 0; k < nps; k++)
    if (q.rp.ps_active_alloc[order[k]] < s.ps_min[order[k]]) unsat++;
  int max_sets = unsat > 0 ? unsat : 1;
  int n_sets = 0, n = 0;
  if (sum) sum[0] = sum[1] = sum[2] = 0.0;
  for (int k = 0; k < nps && n_sets < max_sets; k++) {
    int ps = order[k];
    int tb = s.ps_task_begin[ps], te = s.ps_task_begin[ps + 1];
    int n_alloc = q.rp.ps_active_alloc[ps];
    int max_tasks = n_alloc >= s.ps_min[ps] ? 1 : s.ps_min[ps] - n_alloc;  // :144-153
    int taken = 0;
    for (int i = tb; i < te && taken < max_tasks; i++) {
      int t = s.ps_sorted_tasks[i];
      if (!should_allocate(q, t, real)) continue;
      if (sum)
        for (int r = 0; r < QR; r++) sum[r] = __dadd_rn(sum[r], s.t_req[(size_t)t * s.R + r]);
      else
        q.rp.tta[n] = t;
      n++;
      taken++;
    }
    if (taken > 0) n_sets++;
  }
  return n;
}
__device__ const double *job_init_resource(Seq &q, int job) {
  double *c = q.rp.j_req + (size_t)job * QR;
  if (!q.rp.j_req_valid[job]) {
    tasks_to_allocate(q, job, false, c);
    q.rp.j_req_valid[job] = 1;
  }
  return c;
}
__device__ bool has_tasks_to_allocate(const Seq &q, int job) {  // :18-25 (isRealAllocation = true)
  const DevSnap &s = *q.s;
  for (int ps = s.j_ps_begin[job]; ps < s.j_ps_begin[job + 1]; ps++)
    for (int i = s.ps_task_begin[ps]; i < s.ps_task_begin[ps + 1]; i++)
      if (q.rp.t_status[s.ps_sorted_tasks[i]] == KAI_POD_PENDING) return true;
  return false;
}

// ---- capacity policy (plugins/proportion/capacity_policy) ----
__device__ bool over_capacity(Seq &q, int job, const double *req) {
  const DevSnap &s = *q.s;
  for (int qi = s.j_queue[job]; qi >= 0; qi = s.q_parent[qi])
    for (int r = 0; r < QR; r++) {
      double lim = s.q_limit[(size_t)r * s.Q + qi];
      if (lim == KAI_UNLIMITED) continue;
      if (req[r] == 0) continue;
      if (lim < __dadd_rn(q_alloc(q, r, qi), req[r])) return true;
    }
  if (s.j_flags[job] & KAI_JOB_PREEMPTIBLE) return false;
  for (int qi = s.j_queue[job]; qi >= 0; qi = s.q_parent[qi])
    for (int r = 0; r < QR; r++) {
      double d = s.q_deserved[(size_t)r * s.Q + qi];
      if (d == KAI_UNLIMITED) continue;
      if (req[r] == 0) continue;
      if (d < __dadd_rn(q_alloc_np(q, r, qi), req[r])) return true;
    }
  return false;
}

// ---- queue ordering (plugins/proportion/queue_order/queue_order.go:19-73) ----
struct QView {
  double fair[QR], alloc[QR], deserved[QR], limit[QR];
  int priority;
  long long creation;
};
__device__ void load_qview(Seq &q, int qi, QView &v) {
  const DevSnap &s = *q.s;
  for (int r = 0; r < QR; r++) {
    size_t o = (size_t)r * s.Q + qi;
    v.fair[r] = s.q_fair[o];
    v.alloc[r] = q.rp.q_alloc[o];
    v.deserved[r] = s.q_deserved[o];
    v.limit[r] = s.q_limit[o];
  }
  v.priority = s.q_priority[qi];
  v.creation = s.q_creation[qi];
}
__device__ double dominant_share(const QView &v, const double *alloc, const double *total) {  // queue_resource_share.go:142-166
  double dom = 0.0;
  for (int r = 0; r < QR; r++) {
    double allocatable = allocatable_share(v.deserved[r], v.fair[r], v.limit[r]);
    if (allocatable == KAI_UNLIMITED) allocatable = total[r];
    double value = allocatable == 0 ? __dmul_rn(alloc[r], 1000.0) : __ddiv_rn(alloc[r], allocatable);
    dom = fmax(dom, value);
  }
  return dom;
}
__device__ int queue_order_result(const QView &l, const QView &r, const double *lreq, const double *rreq,
                                  const double *total) {
  bool lo = true, ro = true;  // :87-100
  for (int i = 0; i < QR; i++) {
    if (l.fair[i] >= l.alloc[i]) lo = false;
    if (r.fair[i] >= r.alloc[i]) ro = false;
  }
  if (!lo && ro) return -1;
  if (lo && !ro) return 1;
  double lw[QR], rw[QR];  // :102-128
  bool ls = true, rs = true;
  for (int i = 0; i < QR; i++) {
    lw[i] = __dadd_rn(l.alloc[i], lreq[i]);
    rw[i] = __dadd_rn(r.alloc[i], rreq[i]);
    if (compare_quantities(lw[i], l.deserved[i]) > 0) ls = false;
    if (compare_quantities(rw[i], r.deserved[i]) > 0) rs = false;
  }
  if (ls && !rs) return -1;
  if (rs && !ls) return 1;
  if (l.priority > r.priority) return -1;  // :75-85
  if (l.priority < r.priority) return 1;
  bool lv = false, rv = false;  // :130-180
  double la[QR], ra[QR];
  for (int i = 0; i < QR; i++) {
    la[i] = allocatable_share(l.deserved[i], l.fair[i], l.limit[i]);
    ra[i] = allocatable_share(r.deserved[i], r.fair[i], r.limit[i]);
    if (la[i] == 0 && lw[i] > 0) lv = true;
    if (ra[i] == 0 && rw[i] > 0) rv = true;
  }
  if (lv && !rv) return 1;
  if (!lv && rv) return -1;
  double lsh = dominant_share(l, lw, total), rsh = dominant_share(r, rw, total);  // :182-201
  if (lsh < rsh) return -1;
  if (lsh > rsh) return 1;
  lsh = dominant_share(l, l.alloc, total);  // :203-219
  rsh = dominant_share(r, r.alloc, total);
  if (lsh < rsh) return -1;
  if (lsh > rsh) return 1;
  bool l_le_r = true, r_le_l = true;  // :221-233
  for (int i = 0; i < QR; i++) {
    if (compare_quantities(la[i], ra[i]) > 0) l_le_r = false;
    if (compare_quantities(ra[i], la[i]) > 0) r_le_l = false;
  }
  if (!r_le_l && l_le_r) return -1;
  if (!l_le_r && r_le_l) return 1;
  if (l.creation < r.creation) return -1;  // :235-240
  return 1;
}

// ---- job-order tree (actions/utils/job_order_by_queue.go), one node per queue ----
__device__ __forceinline__ bool qn_is_leaf(const Seq &q, int qi) { return q.s->q_nchildren[qi] == 0; }
__device__ __forceinline__ int *qn_items(Seq &q, int qi) {
  return qn_is_leaf(q, qi) ? q.rp.leaf_heap + q.s->q_job_begin[qi] : q.rp.child_heap + q.s->q_child_begin[qi];
}
__device__ __forceinline__ int &qn_len(Seq &q, int qi) {
  return qn_is_leaf(q, qi) ? q.rp.leaf_len[qi] : q.rp.child_len[qi];
}
// plugins/elastic/elastic.go:50-63: 0 below, 1 exactly at, 2 above minAvailable
__device__ int elastic_class(const Seq &q, int job) {
  bool exactly = true;
  for (int ps = q.s->j_ps_begin[job]; ps < q.s->j_ps_begin[job + 1]; ps++) {
    int n = q.rp.ps_active_alloc[ps], m = q.s->ps_min[ps];
    if (n < m) return 0;
    if (n > m) exactly = false;
  }
  return exactly ? 1 : 2;
}
// session_plugins.go:227-242 JobOrderFn = priority (priority.go:41-54), elastic (elastic.go:25-48), creation, UID
__device__ bool job_less(const Seq &q, int l, int r) {
  int lp = q.s->j_priority[l], rp = q.s->j_priority[r];
  if (lp > rp) return true;
  if (lp < rp) return false;
  int lc = elastic_class(q, l), rc = elastic_class(q, r);
  if (lc != rc) return lc < rc;
  return q.s->j_order_rank[l] < q.s->j_order_rank[r];
}
__device__ int best_job(Seq &q, int qi) {  // :283-292 getBestJobFromNode
  while (!qn_is_leaf(q, qi)) qi = q.rp.child_heap[q.s->q_child_begin[qi]];
  return q.rp.leaf_heap[q.s->q_job_begin[qi]];
}
__device__ bool node_less(Seq &q, int l, int r) {  // :256-278 buildNodeOrderFn (pending order)
  if (qn_len(q, l) == 0) return true;
  if (qn_len(q, r) == 0) return false;
  const double *lreq = job_init_resource(q, best_job(q, l));
  double lr[QR] = {lreq[0], lreq[1], lreq[2]};
  const double *rreq = job_init_resource(q, best_job(q, r));
  QView lv, rv;
  load_qview(q, l, lv);
  load_qview(q, r, rv);
  return queue_order_result(lv, rv, lr, rreq, q.s->total) < 0;
}
// container/heap up/down with the two comparators
template <bool kJobs>
__device__ __forceinline__ bool heap_less(Seq &q, int a, int b) {
  if (kJobs) return job_less(q, a, b);
  return node_less(q, a, b);
}
template <bool kJobs>
__device__ void heap_up(Seq &q, int *items, int j) {
  for (;;) {
    int i = (j - 1) / 2;
    if (i == j || !heap_less<kJobs>(q, items[j], items[i])) break;
    int t = items[i];
    items[i] = items[j];
    items[j] = t;
    j = i;
  }
}
template <bool kJobs>
__device__ bool heap_down(Seq &q, int *items, int i0, int n) {
  int i = i0;
  for (;;) {
    int j1 = 2 * i + 1;
    if (j1 >= n || j1 < 0) break;
    int j = j1;
    int j2 = j1 + 1;
    if (j2 < n && heap_less<kJobs>(q, items[j2], items[j1])) j = j2;
    if (!heap_less<kJobs>(q, items[j], items[i])) break;
    int t = items[i];
    items[i] = items[j];
    items[j] = t;
    i = j;
  }
  return i > i0;
}
template <bool kJobs>
__device__ void heap_push(Seq &q, int *items, int &len, int x) {
  items[len++] = x;
  heap_up<kJobs>(q, items, len - 1);
}
template <bool kJobs>
__device__ int heap_pop(Seq &q, int *items, int &len) {
  int n = len - 1;
  int t = items[0];
  items[0] = items[n];
  items[n] = t;
  heap_down<kJobs>(q, items, 0, n);
  len = n;
  return items[n];
}
__device__ void mark_ancestors(Seq &q, int qi) {  // :246-250
  for (int c = qi; c >= 0; c = q.s->q_parent[c]) q.rp.qn_flags[c] |= QN_REORDER;
}
__device__ void ensure_chain(Seq &q, int child) {  // :135-175
  for (;;) {
    int p = q.s->q_parent[child];
    if (p < 0) {
      if (!(q.rp.qn_flags[child] & QN_LINKED)) {
        heap_push<false>(q, q.rp.root_heap, q.root_len, child);
        q.rp.qn_flags[child] |= QN_LINKED;
      }
      return;
    }
    bool is_new = !(q.rp.qn_flags[p] & QN_EXISTS);
    if (is_new) {
      q.rp.qn_flags[p] = QN_EXISTS;
      q.rp.child_len[p] = 0;
    }
    if (!(q.rp.qn_flags[child] & QN_LINKED)) {
      heap_push<false>(q, q.rp.child_heap + q.s->q_child_begin[p], q.rp.child_len[p], child);
      q.rp.qn_flags[child] |= QN_LINKED;
    }
    if (!is_new) return;
    child = p;
  }
}
__device__ void push_job(Seq &q, int job) {  // :90-119
  int qi = q.s->j_queue[job];
  if (!qn_is_leaf(q, qi)) return;
  bool needs_linking = !(q.rp.qn_flags[qi] & QN_EXISTS);
  if (needs_linking) {
    q.rp.qn_flags[qi] = QN_EXISTS;
    q.rp.leaf_len[qi] = 0;
  }
  heap_push<true>(q, q.rp.leaf_heap + q.s->q_job_begin[qi], q.rp.leaf_len[qi], job);
  if (needs_linking) ensure_chain(q, qi);
  mark_ancestors(q, qi);
}
__device__ int get_next_node(Seq &q, int *items, int &len) {  // :193-215
  for (;;) {
    if (len == 0) return -1;
    int ni = items[0];
    if (q.rp.qn_flags[ni] & QN_REORDER) {
      if (!heap_down<false>(q, items, 0, len)) heap_up<false>(q, items, 0);  // heap.Fix(0)
      q.rp.qn_flags[ni] &= ~QN_REORDER;
      continue;
    }
    if (qn_len(q, ni) == 0) return -1;
    return ni;
  }
}
__device__ void handle_pop(Seq &q, int qi) {  // :219-243
  for (;;) {
    if (qn_len(q, qi) == 0) {
      int p = q.s->q_parent[qi];
      if (p >= 0)
        heap_pop<false>(q, q.rp.child_heap + q.s->q_child_begin[p], q.rp.child_len[p]);
      else
        heap_pop<false>(q, q.rp.root_heap, q.root_len);
      q.rp.qn_flags[qi] = 0;
      if (p < 0) return;
      qi = p;
      continue;
    }
    mark_ancestors(q, qi);
    return;
  }
}
__device__ int pop_next_job(Seq &q) {  // :61-88
  if (q.root_len == 0) return -1;
  int ni = get_next_node(q, q.rp.root_heap, q.root_len);
  while (ni >= 0 && !qn_is_leaf(q, ni)) ni = get_next_node(q, q.rp.child_heap + q.s->q_child_begin[ni], q.rp.child_len[ni]);
  if (ni < 0) return -1;
  int job = heap_pop<true>(q, q.rp.leaf_heap + q.s->q_job_begin[ni], q.rp.leaf_len[ni]);
  handle_pop(q, ni);
  return job;
}

// ---- min/max trackers ----
// update after a scan placement that lowered NonAllocated(res) of the winning node from b to a
__device__ void track_decrease(Track &t, uint32_t f, double a) {
  if (t.dirty) return;
  if (f & WF_B_EQ_MX) {
    if (--t.cnt_mx == 0) {
      t.dirty = 1;
      return;
    }
  }
  if (f & WF_A_LT_MN) {
    t.mn = a;
    t.cnt_mn = 1;
  } else if (f & WF_A_EQ_MN) {
    t.cnt_mn++;
  }
}

// =============================================================================================
// cooperative pieces (all threads of the CTA)
// =============================================================================================
struct Cand {
  double score;
  uint32_t rank;
  int ln;
};
__device__ __forceinline__ bool better(double sa, uint32_t ra, double sb, uint32_t rb) {
  if (ra == kNoRank) return false;
  if (rb == kNoRank) return true;
  return sa > sb || (sa == sb && ra < rb);
}

// pack.go:45-64
__device__ __forceinline__ double binpack_score(double mn, double mx, double cur, double overall) {
  if (overall == 0) return 0.0;
  if (mx == 0) return 0.0;
  if (mn == mx) return 9.0;
  double t1 = __dsub_rn(cur, mn);
  double t2 = __dsub_rn(mx, mn);
  double t3 = __ddiv_rn(t1, t2);
  double t4 = __dsub_rn(1.0, t3);
  return __dmul_rn(9.0, t4);
}

// The sweep: FittingNode (session.go:201-232) + NodeOrderFn sum (session_plugins.go:427-437) per node of the
// tile, then argmax on (score desc, name rank asc) (session.go:466-485).
__device__ Cand scan_tile(const Tile &tl, const Decision &d, const DevSnap &s, Cand *sh_warp) {
  Cand best;
  best.score = -1.0;
  best.rank = kNoRank;
  best.ln = -1;
  const uint32_t *mask = d.pred_class >= 0 ? s.pred_mask + (size_t)d.pred_class * s.mask_words : nullptr;
  for (int ln = threadIdx.x; ln < tl.count; ln += blockDim.x) {
    bool fit_ri = true, fit_i = true;
#pragma unroll 4
    for (int r = 0; r < tl.R; r++) {
      double I = tl.I[r * tl.npc + ln];
      double avail = __dadd_rn(I, tl.L[r * tl.npc + ln]);
      double rq = d.req[r];
      if (r >= 3) {
        if (rq != 0 && rq > avail) fit_ri = false;
        if (rq != 0 && rq > I) fit_i = false;
      } else {
        if (rq > avail) fit_ri = false;
        if (rq > I) fit_i = false;
      }
    }
    if (!fit_ri) continue;
    int n = tl.base + ln;
    if (mask && !((mask[n >> 5] >> (n & 31)) & 1u)) continue;
    double score = 0.0;
    score = __dadd_rn(score, (d.best_effort || fit_i) ? 100.0 : 0.0);                       // nodeavailability
    score = __dadd_rn(score, 0.0);                                                          // gpusharingorder
    bool cpu_only_node = !(tl.flags[ln] & KAI_NODE_NOT_CPU_ONLY) && tl.Agpu[ln] <= 0;
    score = __dadd_rn(score, (!d.gpu_task && cpu_only_node) ? 10.0 : 0.0);                  // resourcetype
    score = __dadd_rn(score, (d.nominated == n) ? 1000000.0 : 0.0);                         // nominatednode
    double cur = __dadd_rn(tl.I[d.res * tl.npc + ln], tl.L[d.res * tl.npc + ln]);
    double overall = d.res == KAI_RES_GPU ? tl.Agpu[ln] : tl.Acpu[ln];
    double place;
    if (d.strategy == KAI_PLACEMENT_BINPACK) {
      place = binpack_score(d.mn, d.mx, cur, overall);
    } else {  // spread.go:16-36
      double cnt = d.res == KAI_RES_GPU ? (double)(long long)tl.gpu_count[ln] : overall;
      place = cnt == 0 ? 0.0 : __ddiv_rn(cur, cnt);
    }
    score = __dadd_rn(score, place);
    uint32_t rk = (uint32_t)tl.rank[ln];
    if (better(score, rk, best.score, best.rank)) {
      best.score = score;
      best.rank = rk;
      best.ln = ln;
    }
  }
  // warp argmax
  for (int o = 16; o > 0; o >>= 1) {
    double os = __shfl_down_sync(0xffffffffu, best.score, o);
    uint32_t orank = __shfl_down_sync(0xffffffffu, best.rank, o);
    int oln = __shfl_down_sync(0xffffffffu, best.ln, o);
    if (better(os, orank, best.score, best.rank)) {
      best.score = os;
      best.rank = orank;
      best.ln = oln;
    }
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh_warp[warp] = best;
  __syncthreads();
  if (warp == 0) {
    int nw = blockDim.x >> 5;
    Cand c;
    if (lane < nw)
      c = sh_warp[lane];
    else {
      c.score = -1.0;
      c.rank = kNoRank;
      c.ln = -1;
    }
    for (int o = 16; o > 0; o >>= 1) {
      double os = __shfl_down_sync(0xffffffffu, c.score, o);
      uint32_t orank = __shfl_down_sync(0xffffffffu, c.rank, o);
      int oln = __shfl_down_sync(0xffffffffu, c.ln, o);
      if (better(os, orank, c.score, c.rank)) {
        c.score = os;
        c.rank = orank;
        c.ln = oln;
      }
    }
    best = c;
  }
  return best;  // valid on thread 0
}

// slot layout per CTA and parity: 4 x u64 = A{score bits, [seq:24][flags:8][rank:32]}, B{cur_a gpu, cur_a cpu} tagged
// through a second pair C{seq, 0}.  6 x u64 rounded to 8.
constexpr int kSlotWords = 8;

// all-to-all exchange of the per-CTA candidates; executed by warp 0; result broadcast through ctl.win
__device__ void exchange_candidates(const ActionParams &p, Ctl &ctl, const Tile &tl, const Decision &d, Cand local,
                                    unsigned int seq) {
  const int lane = threadIdx.x & 31;
  unsigned long long *buf = p.xbuf + (size_t)(seq & 1) * kMaxGrid * kSlotWords;
  const unsigned int tag = seq & 0xffffffu;
  if (lane == 0) {
    uint32_t flags = 0;
    double a_gpu = 0, a_cpu = 0;
    if (local.rank != kNoRank) {
      int ln = local.ln;
      bool fit_i = true;
      for (int r = 0; r < tl.R; r++) {
        double I = tl.I[r * tl.npc + ln];
        double rq = d.req[r];
        if (r >= 3 ? (rq != 0 && rq > I) : (rq > I)) fit_i = false;
      }
      bool to_idle = !d.pipeline_only && (d.best_effort || fit_i);
      if (to_idle) flags |= WF_FITS_IDLE;
      for (int k = 0; k < 2; k++) {
        int res = k == 0 ? KAI_RES_GPU : KAI_RES_CPU;
        double overall = k == 0 ? tl.Agpu[ln] : tl.Acpu[ln];
        if (overall == 0 || d.req[res] == 0) continue;
        double I = tl.I[res * tl.npc + ln], L = tl.L[res * tl.npc + ln];
        double b = __dadd_rn(I, L);
        double a = to_idle ? __dadd_rn(__dsub_rn(I, d.req[res]), L) : __dadd_rn(I, __dsub_rn(L, d.req[res]));
        const Track &t = ctl.trk[k];
        uint32_t f = 0;
        if (b == t.mx) f |= WF_B_EQ_MX;
        if (a < t.mn)
          f |= WF_A_LT_MN;
        else if (a == t.mn)
          f |= WF_A_EQ_MN;
        flags |= f << (3 * k);
        if (k == 0)
          a_gpu = a;
        else
          a_cpu = a;
      }
    }
    unsigned long long *slot = buf + (size_t)blockIdx.x * kSlotWords;
    // payload first (only fetched when a WF_A_LT_MN bit is set), then the candidate word pair
    st_relaxed_b128(slot + 2, (unsigned long long)__double_as_longlong(a_gpu), (unsigned long long)tag);
    st_relaxed_b128(slot + 4, (unsigned long long)__double_as_longlong(a_cpu), (unsigned long long)tag);
    unsigned long long hi = ((unsigned long long)tag << 40) | ((unsigned long long)(flags & 0xffu) << 32) |
                            (unsigned long long)local.rank;
    st_relaxed_b128(slot, (unsigned long long)__double_as_longlong(local.score), hi);
  }
  // gather: lane l polls slots l, l+32, ...
  double bs = -1.0;
  uint32_t brank = kNoRank, bflags = 0;
  int bslot = -1;
  for (int c = lane; c < p.grid; c += 32) {
    const unsigned long long *slot = buf + (size_t)c * kSlotWords;
    unsigned long long lo, hi;
    do {
      ld_relaxed_b128(slot, lo, hi);
    } while ((unsigned int)(hi >> 40) != tag);
    double sc = __longlong_as_double((long long)lo);
    uint32_t rk = (uint32_t)(hi & 0xffffffffu);
    if (better(sc, rk, bs, brank)) {
      bs = sc;
      brank = rk;
      bflags = (uint32_t)((hi >> 32) & 0xffu);
      bslot = c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    double os = __shfl_xor_sync(0xffffffffu, bs, o);
    uint32_t orank = __shfl_xor_sync(0xffffffffu, brank, o);
    uint32_t ofl = __shfl_xor_sync(0xffffffffu, bflags, o);
    int osl = __shfl_xor_sync(0xffffffffu, bslot, o);
    if (better(os, orank, bs, brank)) {
      bs = os;
      brank = orank;
      bflags = ofl;
      bslot = osl;
    }
  }
  if (lane == 0) {
    ctl.win.score = bs;
    ctl.win.rank = brank;
    ctl.win.flags = bflags;
    ctl.win.node = brank == kNoRank ? -1 : p.s.rank_to_node[brank];
    if (brank != kNoRank) {
      // tracker maintenance; fetch cur_a only when a new global minimum was established
      for (int k = 0; k < 2; k++) {
        uint32_t f = (bflags >> (3 * k)) & 7u;
        double a = 0;
        if (f & WF_A_LT_MN) {
          const unsigned long long *slot = buf + (size_t)bslot * kSlotWords + 2 + 2 * k;
          unsigned long long lo, hi;
          do {
            ld_relaxed_b128(slot, lo, hi);
          } while ((unsigned int)hi != tag);
          a = __longlong_as_double((long long)lo);
        }
        if (f) track_decrease(ctl.trk[k], f, a);
      }
    }
  }
}

// min/max exchange (rare): every CTA publishes local (mn, mx, cnt_mn, cnt_mx) for gpu and cpu
__device__ void exchange_minmax(const ActionParams &p, Ctl &ctl, const Tile &tl, unsigned int seq, double *sh_d,
                                int *sh_i) {
  // block-level reduction by thread 0 over per-thread partials (tiles are small; rare path)
  double mn[2] = {DBL_MAX, DBL_MAX}, mx[2] = {0, 0};
  for (int ln = threadIdx.x; ln < tl.count; ln += blockDim.x) {
    for (int k = 0; k < 2; k++) {
      int res = k == 0 ? KAI_RES_GPU : KAI_RES_CPU;
      double overall = k == 0 ? tl.Agpu[ln] : tl.Acpu[ln];
      if (overall == 0) continue;
      double cur = __dadd_rn(tl.I[res * tl.npc + ln], tl.L[res * tl.npc + ln]);
      if (cur < mn[k]) mn[k] = cur;
      if (cur > mx[k]) mx[k] = cur;
    }
  }
  for (int k = 0; k < 2; k++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
    sh_d[warp * 4 + 0] = mn[0];
    sh_d[warp * 4 + 1] = mx[0];
    sh_d[warp * 4 + 2] = mn[1];
    sh_d[warp * 4 + 3] = mx[1];
  }
  __syncthreads();
  for (int w = 0; w < nw; w++) {
    mn[0] = fmin(mn[0], sh_d[w * 4 + 0]);
    mx[0] = fmax(mx[0], sh_d[w * 4 + 1]);
    mn[1] = fmin(mn[1], sh_d[w * 4 + 2]);
    mx[1] = fmax(mx[1], sh_d[w * 4 + 3]);
  }
  // counts at the local extremes
  int c[4] = {0, 0, 0, 0};
  for (int ln = threadIdx.x; ln < tl.count; ln += blockDim.x) {
    for (int k = 0; k < 2; k++) {
      int res = k == 0 ? KAI_RES_GPU : KAI_RES_CPU;
      double overall = k == 0 ? tl.Agpu[ln] : tl.Acpu[ln];
      if (overall == 0) continue;
      double cur = __dadd_rn(tl.I[res * tl.npc + ln], tl.L[res * tl.npc + ln]);
      if (cur == mn[k]) c[2 * k]++;
      if (cur == mx[k]) c[2 * k + 1]++;
    }
  }
  for (int i = 0; i < 4; i++)
    for (int o = 16; o > 0; o >>= 1) c[i] += __shfl_xor_sync(0xffffffffu, c[i], o);
  __syncthreads();
  if (lane == 0)
    for (int i = 0; i < 4; i++) sh_i[warp * 4 + i] = c[i];
  __syncthreads();
  if (warp != 0) return;
  unsigned long long *buf = p.mmbuf + (size_t)(seq & 1) * kMaxGrid * 8;
  const unsigned long long tag = seq;
  if (lane == 0) {
    int tot[4] = {0, 0, 0, 0};
    for (int w = 0; w < nw; w++)
      for (int i = 0; i < 4; i++) tot[i] += sh_i[w * 4 + i];
    unsigned long long *slot = buf + (size_t)blockIdx.x * 8;
    // four tagged 128-bit words: {value, [tag:32][count:32]}
    st_relaxed_b128(slot + 0, (unsigned long long)__double_as_longlong(mn[0]), (tag << 32) | (unsigned int)tot[0]);
    st_relaxed_b128(slot + 2, (unsigned long long)__double_as_longlong(mx[0]), (tag << 32) | (unsigned int)tot[1]);
    st_relaxed_b128(slot + 4, (unsigned long long)__double_as_longlong(mn[1]), (tag << 32) | (unsigned int)tot[2]);
    st_relaxed_b128(slot + 6, (unsigned long long)__double_as_longlong(mx[1]), (tag << 32) | (unsigned int)tot[3]);
  }
  // every lane walks all slots redundantly in the same order (deterministic combine), lane 0 keeps the result
  if (lane == 0) {
    double gmn[2] = {DBL_MAX, DBL_MAX}, gmx[2] = {0, 0};
    long long cmn[2] = {0, 0}, cmx[2] = {0, 0};
    for (int cta = 0; cta < p.grid; cta++) {
      const unsigned long long *slot = buf + (size_t)cta * 8;
      for (int k = 0; k < 2; k++) {
        unsigned long long lo, hi;
        do {
          ld_relaxed_b128(slot + 4 * k, lo, hi);
        } while ((hi >> 32) != (tag & 0xffffffffu));
        double v = __longlong_as_double((long long)lo);
        int cnt = (int)(hi & 0xffffffffu);
        if (cnt > 0) {
          if (v < gmn[k]) {
            gmn[k] = v;
            cmn[k] = cnt;
          } else if (v == gmn[k])
            cmn[k] += cnt;
        }
        do {
          ld_relaxed_b128(slot + 4 * k + 2, lo, hi);
        } while ((hi >> 32) != (tag & 0xffffffffu));
        v = __longlong_as_double((long long)lo);
        cnt = (int)(hi & 0xffffffffu);
        if (cnt > 0) {
          if (v > gmx[k]) {
            gmx[k] = v;
            cmx[k] = cnt;
          } else if (v == gmx[k])
            cmx[k] += cnt;
        }
      }
    }
    for (int k = 0; k < 2; k++) {
      ctl.trk[k].mn = gmn[k];
      ctl.trk[k].mx = gmx[k];
      ctl.trk[k].cnt_mn = (int)cmn[k];
      ctl.trk[k].cnt_mx = (int)cmx[k];
      ctl.trk[k].dirty = 0;
    }
  }
}

// =============================================================================================
// sequencer steps (thread 0)
// =============================================================================================
// InitializeWithJobs (input_jobs.go:21-68) in canonical order: leaf queues ascending, jobs of a queue in
// JobOrderFn order (the Go map order is unspecified).  Leaf heaps start as sorted arrays.
__device__ void seq_init_job_order(Seq &q) {
  const DevSnap &s = *q.s;
  for (int qi = 0; qi < s.Q; qi++) {
    if (s.q_nchildren[qi] != 0) continue;
    if (q.rp.leaf_len[qi] == 0) continue;  // filled by the parallel init below
    q.rp.qn_flags[qi] = QN_EXISTS;
    ensure_chain(q, qi);
    mark_ancestors(q, qi);
  }
}

// builds ctl.dec for item k of the current job; returns false when the task cannot be placed at all
__device__ bool seq_prepare_task(Seq &q, int t) {
  const DevSnap &s = *q.s;
  Ctl &c = *q.ctl;
  Decision &d = c.dec;
  const double *rq = s.t_req + (size_t)t * s.R;
  bool gpu_task = rq[KAI_RES_GPU] > 0;
  // predicates.go:196-200 -> capacity_policy.go:51-61 with node_info.go:734-744
  double creq[QR] = {rq[KAI_RES_CPU], rq[KAI_RES_MEM], gpu_task ? 1.0 : 0.0};
  if (over_capacity(q, s.t_job[t], creq)) return false;
  for (int r = 0; r < KAI_MAX_RES; r++) d.req[r] = r < s.R ? rq[r] : 0.0;
  d.task = t;
  d.gpu_task = gpu_task;
  d.res = gpu_task ? KAI_RES_GPU : KAI_RES_CPU;
  d.strategy = gpu_task ? q.cfg->gpu_placement : q.cfg->cpu_placement;
  d.pipeline_only = 0;
  d.nominated = s.t_nominated ? s.t_nominated[t] : -1;
  d.pred_class = s.t_pred_class ? s.t_pred_class[t] : -1;
  bool empty = !(rq[KAI_RES_GPU] > 0.01) && !(rq[KAI_RES_CPU] >= 10) && !(rq[KAI_RES_MEM] >= 10.0 * 1024 * 1024);
  for (int r = 3; r < s.R; r++)
    if (rq[r] >= 10) empty = false;
  d.best_effort = empty;
  c.need_minmax = (d.strategy == KAI_PLACEMENT_BINPACK) && c.trk[gpu_task ? 0 : 1].dirty;
  return true;
}

__device__ void seq_apply_winner(Seq &q, int t) {
  Ctl &c = *q.ctl;
  q.decisions++;
  q.nodes_scanned += q.s->N;
  if (c.win.node < 0) {
    c.item_ok = 0;
    return;
  }
  if (c.win.flags & WF_FITS_IDLE)
    stmt_allocate(q, t, c.win.node);
  else
    stmt_pipeline(q, t, c.win.node);
  c.item_ok = 1;
}

// job_info.go:443-464 ShouldPipelineJob
__device__ bool should_pipeline_job(const Seq &q, int job) {
  const DevSnap &s = *q.s;
  for (int ps = s.j_ps_begin[job]; ps < s.j_ps_begin[job + 1]; ps++) {
    bool has_pipelined = false;
    int active = 0;
    for (int i = s.ps_task_begin[ps]; i < s.ps_task_begin[ps + 1]; i++) {
      int st = q.rp.t_status[s.ps_sorted_tasks[i]];
      if (st == KAI_POD_PIPELINED)
        has_pipelined = true;
      else if (st & kActiveAllocated)
        active++;
    }
    if (has_pipelined && active < s.ps_min[ps]) return true;
  }
  return false;
}

__device__ void record_visit(Seq &q, int job, int outcome) {
  if (q.is_cta0 && q.n_visits < q.visits_cap) {
    q.visits[q.n_visits].job = job;
    q.visits[q.n_visits].outcome = outcome;
  }
  q.n_visits++;
}

// =============================================================================================
// the kernel
// =============================================================================================
__global__ void __launch_bounds__(kThreads, 1) k_action(const __grid_constant__ ActionParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ Ctl ctl;
  __shared__ Tile tile;
  __shared__ Seq seq;
  __shared__ Cand sh_warp[kThreads / 32];
  __shared__ double sh_d[(kThreads / 32) * 4];
  __shared__ int sh_i[(kThreads / 32) * 4];
  const DevSnap &s = p.s;
  const int tid = threadIdx.x;

  // ---- carve the node tile ----
  if (tid == 0) {
    int npc = p.nodes_per_cta;
    unsigned char *ptr = smem;
    tile.npc = npc;
    tile.R = s.R;
    tile.base = p.node_base + blockIdx.x * npc;
    int end = min(p.node_base + p.node_count, tile.base + npc);
    tile.count = max(0, end - tile.base);
    tile.I = (double *)ptr;
    ptr += sizeof(double) * s.R * npc;
    tile.L = (double *)ptr;
    ptr += sizeof(double) * s.R * npc;
    tile.Agpu = (double *)ptr;
    ptr += sizeof(double) * npc;
    tile.Acpu = (double *)ptr;
    ptr += sizeof(double) * npc;
    tile.gpu_count = (double *)ptr;
    ptr += sizeof(double) * npc;
    tile.rank = (int *)ptr;
    ptr += sizeof(int) * npc;
    tile.flags = (uint32_t *)ptr;
    // replica pointers
    unsigned char *a = p.replica_arena + (size_t)blockIdx.x * p.replica_bytes;
    auto take = [&](size_t bytes) {
      unsigned char *r = a;
      a += (bytes + 15) & ~(size_t)15;
      return r;
    };
    Replica &rp = seq.rp;
    rp.q_alloc = (double *)take(sizeof(double) * QR * s.Q);
    rp.q_alloc_np = (double *)take(sizeof(double) * QR * s.Q);
    rp.t_status = (int *)take(sizeof(int) * s.T);
    rp.t_node = (int *)take(sizeof(int) * s.T);
    rp.t_node_status = (int *)take(sizeof(int) * s.T);
    rp.t_virtual = (unsigned char *)take(s.T);
    rp.ps_active_alloc = (int *)take(sizeof(int) * s.S);
    rp.j_req = (double *)take(sizeof(double) * QR * s.J);
    rp.j_req_valid = (unsigned char *)take(s.J);
    rp.leaf_heap = (int *)take(sizeof(int) * s.J);
    rp.leaf_len = (int *)take(sizeof(int) * s.Q);
    rp.child_heap = (int *)take(sizeof(int) * s.Q);
    rp.child_len = (int *)take(sizeof(int) * s.Q);
    rp.root_heap = (int *)take(sizeof(int) * (s.n_top + 1));
    rp.qn_flags = (unsigned char *)take(s.Q);
    rp.ops = (Op *)take(sizeof(Op) * p.ops_cap);
    rp.tta = (int *)take(sizeof(int) * (s.max_job_tasks + 1));
    rp.ps_order = (int *)take(sizeof(int) * (s.max_job_podsets + 1));
    seq.s = &p.s;
    seq.cfg = &p.cfg;
    seq.tile = &tile;
    seq.ctl = &ctl;
    seq.n_ops = 0;
    seq.ops_cap = p.ops_cap;
    seq.root_len = 0;
    seq.is_cta0 = blockIdx.x == 0;
    seq.pods_placed = seq.pods_evicted = seq.decisions = seq.nodes_scanned = seq.n_visits = 0;
    seq.minmax_exchanges = 0;
    seq.visits = p.visits;
    seq.visits_cap = p.visits_cap;
    seq.error = 0;
    seq.t_pop = seq.t_prep = seq.t_scan = seq.t_xchg = seq.t_apply = seq.t_finish = seq.t_init = 0;
    ctl.trk[0].dirty = ctl.trk[1].dirty = 1;
    ctl.seq = p.seq0;
    ctl.stop = 0;
  }
  __syncthreads();
  long long tk0 = clock64();

  // ---- load the tile (coalesced per resource row) and the replica state ----
  for (int ln = tid; ln < tile.count; ln += blockDim.x) {
    int n = tile.base + ln;
    for (int r = 0; r < s.R; r++) {
      tile.I[r * tile.npc + ln] = s.idle[(size_t)r * s.N + n];
      tile.L[r * tile.npc + ln] = s.rel[(size_t)r * s.N + n];
    }
    tile.Agpu[ln] = s.alloc[(size_t)KAI_RES_GPU * s.N + n];
    tile.Acpu[ln] = s.alloc[(size_t)KAI_RES_CPU * s.N + n];
    tile.gpu_count[ln] = s.gpu_count[n];
    tile.rank[ln] = s.name_rank[n];
    tile.flags[ln] = s.nflags[n];
  }
  {
    Replica &rp = seq.rp;
    for (int i = tid; i < QR * s.Q; i += blockDim.x) {
      rp.q_alloc[i] = s.q_alloc[i];
      rp.q_alloc_np[i] = s.q_alloc_np[i];
    }
    for (int i = tid; i < s.T; i += blockDim.x) {
      rp.t_status[i] = s.t_status[i];
      rp.t_node[i] = s.t_node[i];
      rp.t_node_status[i] = s.t_node_status[i];
      rp.t_virtual[i] = s.t_virtual[i];
    }
    for (int i = tid; i < s.S; i += blockDim.x) {
      int c = 0;
      for (int k = s.ps_task_begin[i]; k < s.ps_task_begin[i + 1]; k++)
        if (s.t_status[s.ps_sorted_tasks[k]] & kActiveAllocated) c++;
      rp.ps_active_alloc[i] = c;
    }
    for (int i = tid; i < s.J; i += blockDim.x) rp.j_req_valid[i] = 0;
    for (int i = tid; i < s.Q; i += blockDim.x) {
      rp.leaf_len[i] = 0;
      rp.child_len[i] = 0;
      rp.qn_flags[i] = 0;
    }
  }
  __syncthreads();
  // leaf heaps: one thread per leaf queue filters its jobs (input_jobs.go:24-45) and keeps them in
  // JobOrderFn order (host order = priority desc, (creation, uid); elastic class fixed up by insertion)
  for (int qi = tid; qi < s.Q; qi += blockDim.x) {
    if (s.q_nchildren[qi] != 0) continue;
    int b = s.q_job_begin[qi], e = s.q_job_begin[qi + 1];
    int *heap = seq.rp.leaf_heap + b;
    int n = 0;
    for (int k = b; k < e; k++) {
      int job = s.q_jobs_sorted[k];
      // FilterUnready: podset.go:114-120; FilterNonPending
      bool ready = true;
      int pending = 0;
      for (int ps = s.j_ps_begin[job]; ps < s.j_ps_begin[job + 1]; ps++) {
        int alive = 0, gated = 0;
        for (int i = s.ps_task_begin[ps]; i < s.ps_task_begin[ps + 1]; i++) {
          int st = s.t_status[s.ps_sorted_tasks[i]];
          if (st & kAlive) alive++;
          if (st & KAI_POD_GATED) gated++;
          if (st == KAI_POD_PENDING) pending++;
        }
        if (alive - gated < s.ps_min[ps]) ready = false;
      }
      if (!ready || pending == 0) continue;
      int i = n++;
      while (i > 0 && job_less(seq, job, heap[i - 1])) {
        heap[i] = heap[i - 1];
        i--;
      }
      heap[i] = job;
    }
    seq.rp.leaf_len[qi] = n;
  }
  __syncthreads();
  if (tid == 0) seq_init_job_order(seq);
  __syncthreads();
  if (tid == 0) seq.t_init = clock64() - tk0;

  // ---- allocate action main loop (actions/allocate/allocate.go:46-111) ----
  for (;;) {
    if (tid == 0) {
      long long tk = clock64();
      int job = pop_next_job(seq);
      ctl.job = job;
      ctl.n_items = 0;
      ctl.job_ok = 0;
      if (job >= 0) {
        seq.n_ops = 0;
        // common/allocate.go:20-36 AllocateJob
        int n = tasks_to_allocate(seq, job, true, nullptr);
        double req[QR] = {0, 0, 0};
        for (int k = 0; k < n; k++)
          for (int r = 0; r < QR; r++) req[r] = __dadd_rn(req[r], s.t_req[(size_t)seq.rp.tta[k] * s.R + r]);
        if (!over_capacity(seq, job, req)) {
          // tasks_to_allocate already emits tasks grouped in PodSetOrderFn order, which is the order
          // allocateSubGroupSetOnNodes/allocatePodSet visit them in (common/allocate.go:62-119)
          ctl.n_items = n;
          ctl.job_ok = 1;
        }
      }
      if (seq.error) ctl.stop = 1;
      seq.t_pop += clock64() - tk;
    }
    __syncthreads();
    if (ctl.job < 0 || ctl.stop) break;
    bool job_success = ctl.job_ok != 0;
    if (job_success) {
      const int n_items = ctl.n_items;
      for (int k = 0; k < n_items; k++) {
        if (tid == 0) {
          long long tk = clock64();
          int t = seq.rp.tta[k];
          ctl.need_minmax = 0;
          ctl.item_ok = seq_prepare_task(seq, t) ? 1 : 0;
          if (ctl.need_minmax) seq.minmax_exchanges++;
          seq.t_prep += clock64() - tk;
        }
        __syncthreads();
        if (!ctl.item_ok) {
          job_success = false;
          break;
        }
        if (ctl.need_minmax) {
          unsigned int sq = ctl.seq;
          exchange_minmax(p, ctl, tile, sq, sh_d, sh_i);
          __syncthreads();
          if (tid == 0) ctl.seq = sq + 1;
          __syncthreads();
        }
        if (tid == 0) {
          int tk = ctl.dec.gpu_task ? 0 : 1;
          ctl.dec.mn = ctl.trk[tk].mn;
          ctl.dec.mx = ctl.trk[tk].mx;
        }
        __syncthreads();
        long long tk1 = clock64();
        Cand local = scan_tile(tile, ctl.dec, s, sh_warp);
        if (tid < 32) {
          long long tk2 = clock64();
          local.score = __shfl_sync(0xffffffffu, local.score, 0);
          local.rank = __shfl_sync(0xffffffffu, local.rank, 0);
          local.ln = __shfl_sync(0xffffffffu, local.ln, 0);
          unsigned int sq = ctl.seq;
          exchange_candidates(p, ctl, tile, ctl.dec, local, sq);
          if (tid == 0) {
            long long tk3 = clock64();
            ctl.seq = sq + 1;
            seq_apply_winner(seq, ctl.dec.task);
            long long tk4 = clock64();
            seq.t_scan += tk2 - tk1;
            seq.t_xchg += tk3 - tk2;
            seq.t_apply += tk4 - tk3;
          }
        }
        __syncthreads();
        if (!ctl.item_ok) {
          job_success = false;
          break;
        }
      }
    }
    if (tid == 0) {
      long long tk = clock64();
      int job = ctl.job;
      if (job_success) {
        if (should_pipeline_job(seq, job)) stmt_convert_all_allocated_to_pipelined(seq, job);
        stmt_commit(seq);
        record_visit(seq, job, 1);
        if (has_tasks_to_allocate(seq, job)) push_job(seq, job);
      } else {
        stmt_rollback(seq, 0);  // Discard (statement.go:522-534)
        record_visit(seq, job, 0);
      }
      if (seq.error) ctl.stop = 1;
      seq.t_finish += clock64() - tk;
    }
    __syncthreads();
    if (ctl.stop) break;
  }

  // ---- write back: tiles by their owners, session state and counters by CTA 0 ----
  __syncthreads();
  for (int ln = tid; ln < tile.count; ln += blockDim.x) {
    int n = tile.base + ln;
    for (int r = 0; r < s.R; r++) {
      s.idle[(size_t)r * s.N + n] = tile.I[r * tile.npc + ln];
      s.rel[(size_t)r * s.N + n] = tile.L[r * tile.npc + ln];
    }
  }
  if (blockIdx.x == 0) {
    Replica &rp = seq.rp;
    for (int i = tid; i < QR * s.Q; i += blockDim.x) {
      s.q_alloc[i] = rp.q_alloc[i];
      s.q_alloc_np[i] = rp.q_alloc_np[i];
    }
    for (int i = tid; i < s.T; i += blockDim.x) {
      s.t_status[i] = rp.t_status[i];
      s.t_node[i] = rp.t_node[i];
      s.t_node_status[i] = rp.t_node_status[i];
      s.t_virtual[i] = rp.t_virtual[i];
    }
    if (tid == 0) {
      p.counters[0] = seq.n_visits;
      p.counters[1] = seq.decisions;
      p.counters[2] = seq.nodes_scanned;
      p.counters[3] = seq.pods_placed;
      p.counters[4] = seq.pods_evicted;
      p.counters[5] = seq.minmax_exchanges;
      p.counters[6] = seq.error;
      p.counters[7] = ctl.seq;
      p.counters[8] = seq.t_init;
      p.counters[9] = seq.t_pop;
      p.counters[10] = seq.t_prep;
      p.counters[11] = seq.t_scan;
      p.counters[12] = seq.t_xchg;
      p.counters[13] = seq.t_apply;
      p.counters[14] = seq.t_finish;
    }
  }
}

}  // namespace kai
