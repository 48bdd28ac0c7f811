// kai_action.cuh — the persistent action kernel (allocate) and its two prepare kernels.
//
//   k_prep_jobs    per job: podset status counters, readiness, JobOrderFn sort key, cached
//                  GetTasksToAllocateInitResource                                        (grid-parallel)
//   k_prep_queues  per leaf queue: eligible jobs in JobOrderFn order                     (thread per queue)
//   k_action       whole Action on device, one CTA per SM, cooperative launch:
//                    CTA 0      the sequencer (warp 0): job-order tree, capacity checks, statement; its
//                               per-queue state lives in shared memory, per-task/job state is the session
//                               state in HBM/L2 (single copy)
//                    CTA 1..G-1 scanners: each keeps its slice of the node tables in shared memory for the
//                               whole action and answers every decision record with one 16-byte candidate
//                  sequencer -> scanners: a 16-word decision record (tagged 128-bit relaxed stores) plus a
//                  list of node deltas since the previous record; scanners -> sequencer: one slot per CTA.
//
// Exactness notes
//   * FittingNode + NodeOrderFn + sortNodesByScore (framework/session.go:201-264,466-485) are evaluated as
//     "argmax over fitting nodes of (score desc, name-rank asc)" with the reference's f64 operation order.
//   * binpack min/max (pack.go:66-86) are tracked incrementally with counts of nodes at the extremes; any
//     event that could change an extreme without being observable marks the tracker dirty and forces a
//     min/max exchange before the next sweep that needs it.
//   * same-node batching: after a sweep picked node n for a pod, the owner of n also reports for how many
//     further pods with the SAME request/flags node n provably stays the argmax (its score does not drop
//     below the winning score, it still fits in the same mode, min/max stay put).  Those pods are placed
//     without a sweep.  DESIGN.md §5 has the argument.
#pragma once
#include <cfloat>
#include <cstdint>

#include "kai_device.cuh"
#include "kai_kernels.cuh"

namespace kai {

constexpr unsigned long long kKeyNone = ~0ull;
constexpr uint32_t kRankNone = 0xFFFFFFu;  // 24-bit rank field
constexpr int kMaxRepeat = 10;             // 6 flag bits per repeat in one 64-bit word

__device__ __forceinline__ unsigned long long make_job_key(int priority, int cls, int order_rank) {
  unsigned long long pinv = (unsigned long long)(unsigned int)(0x40000000 - priority) & 0x7fffffffull;
  return (pinv << 33) | ((unsigned long long)cls << 31) | (unsigned long long)(order_rank & 0x7fffffff);
}

// ---------------------------------------------------------------------------------------------
// prepare kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_prep_jobs(DevSnap s, int filter_non_pending, int filter_unready) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < s.J; j += gridDim.x * blockDim.x) {
    bool ready = true, below = false, exactly = true;
    int pending = 0;
    for (int ps = s.j_ps_begin[j]; ps < s.j_ps_begin[j + 1]; ps++) {
      int act = 0, pend = 0, pipe = 0, alive = 0, gated = 0;
      for (int i = s.ps_task_begin[ps]; i < s.ps_task_begin[ps + 1]; i++) {
        int st = s.t_status[i];  // tasks of a podset are contiguous
        if (st & kActiveAllocated) act++;
        if (st == KAI_POD_PENDING) pend++;
        if (st == KAI_POD_PIPELINED) pipe++;
        if (st & kAlive) alive++;
        if (st & KAI_POD_GATED) gated++;
      }
      s.ps_cnt0[ps] = act;
      s.ps_cnt0[s.S + ps] = pend;
      s.ps_cnt0[2 * s.S + ps] = pipe;
      int m = s.ps_min[ps];
      if (alive - gated < m) ready = false;  // podset.go:114-120
      pending += pend;
      if (act < m) below = true;  // elastic.go:50-63
      if (act > m) exactly = false;
    }
    int cls = below ? 0 : (exactly ? 1 : 2);
    int q = s.j_queue[j];
    bool eligible = (!filter_unready || ready) && (!filter_non_pending || pending > 0) && q >= 0 &&
                    s.q_nchildren[q] == 0;  // input_jobs.go:24-63
    s.j_key0[j] = eligible ? make_job_key(s.j_priority[j], cls, s.j_order_rank[j]) : kKeyNone;
    // GetTasksToAllocateInitResource(job, isRealAllocation=false) (allocation_info.go:87-113) for the common
    // single-podset job; other jobs are evaluated lazily by the sequencer.  Tasks of a podset are stored in
    // TaskOrderFn order (the host renumbers them), so "the first k that should allocate" is a linear scan.
    s.j_req_valid[j] = 0;
    JobRec rec;
    rec.req0[0] = rec.req0[1] = rec.req0[2] = 0;
    rec.n_tta = -1;
    rec.ps0 = s.j_ps_begin[j];
    rec.n_podsets = s.j_ps_begin[j + 1] - s.j_ps_begin[j];
    rec.tb = rec.n_podsets > 0 ? s.ps_task_begin[rec.ps0] : 0;
    rec.cnt[0] = rec.cnt[1] = rec.cnt[2] = 0;
    rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
    if (rec.n_podsets == 1) {
      int ps = rec.ps0;
      int act = s.ps_cnt0[ps], m = s.ps_min[ps];
      rec.cnt[0] = act;
      rec.cnt[1] = s.ps_cnt0[s.S + ps];
      rec.cnt[2] = s.ps_cnt0[2 * s.S + ps];
      int max_tasks = act >= m ? 1 : m - act;
      double sum[QR] = {0, 0, 0};
      int taken = 0;
      bool prefix = true;  // the selected tasks are exactly the first `taken` tasks and all are Pending
      for (int t = s.ps_task_begin[ps]; t < s.ps_task_begin[ps + 1] && taken < max_tasks; t++) {
        int st = s.t_status[t];
        if (!(st == KAI_POD_PENDING || (st == KAI_POD_RELEASING && s.t_virtual[t]))) {
          prefix = false;
          continue;
        }
        if (st != KAI_POD_PENDING) prefix = false;
        for (int r = 0; r < QR; r++) sum[r] = __dadd_rn(sum[r], s.t_req[(size_t)t * s.R + r]);
        taken++;
      }
      for (int r = 0; r < QR; r++) {
        s.j_req[(size_t)j * QR + r] = sum[r];
        rec.req0[r] = sum[r];
      }
      s.j_req_valid[j] = 1;
      if (prefix && act == 0 && rec.cnt[2] == 0) rec.n_tta = taken;
    }
    s.jrec[j] = rec;
  }
}

__global__ void k_prep_queues(DevSnap s) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < s.Q; q += gridDim.x * blockDim.x) {
    int b = s.q_job_begin[q], e = s.q_job_begin[q + 1];
    int n = 0;
    int *out = s.leaf_sorted + b;
    for (int k = b; k < e; k++) {
      int job = s.q_jobs_sorted[k];
      unsigned long long key = s.j_key0[job];
      if (key == kKeyNone) continue;
      int i = n++;
      while (i > 0 && key < s.j_key0[out[i - 1]]) {
        out[i] = out[i - 1];
        i--;
      }
      out[i] = job;
    }
    s.leaf_count[q] = n;
  }
}

// ---------------------------------------------------------------------------------------------
// action kernel state
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

// tracker event bits per resource (gpu bits 0-2, cpu bits 3-5)
enum { WF_B_EQ_MX = 1, WF_A_EQ_MN = 2, WF_A_LT_MN = 4 };
enum { SLOT_TO_IDLE = 64, SLOT_HAS_REPEAT = 128 };

struct Winner {
  double score;
  uint32_t rank;
  uint32_t flags;
  int node;
};

struct Batch {  // same-node batching state
  int valid, node, to_idle, left, idx;
  unsigned long long fl;  // 6 tracker-event bits per repeat
};

enum { DK_SCAN = 1, DK_MINMAX = 2, DK_FLUSH = 3, DK_DONE = 4 };
enum { DB_GPU_TASK = 1, DB_BEST_EFFORT = 2, DB_PIPELINE_ONLY = 4, DB_BATCHING = 8, DB_DIRTY0 = 16, DB_DIRTY1 = 32 };

struct Ctl {  // sequencer control block (shared memory of CTA 0), written by lane 0
  int job, n_items, job_ok, item_ok, need_minmax, use_batch, stop;
  unsigned int seq;  // sequence number of the next decision record
  int n_delta;       // node deltas queued for the next record
  Decision dec;
  Winner win;
  Track trk[2];  // 0 gpu, 1 cpu
  Batch batch;
  unsigned long long dw[kDecWords];
  // context of the job being allocated: its podset counters stay here and are written back at the end
  int ctx_job, ctx_ps, ctx_fresh, ctx_queue, ctx_preempt, ctx_base;
  int ctx_cnt[3];
};

struct Tile {  // shared-memory node tile of this CTA
  double *I, *L;        // [R][npc]
  double *Agpu, *Acpu;  // [npc]
  double *gpu_count;    // [npc]
  int *rank;            // [npc]
  uint32_t *flags;      // [npc]
  int npc, base, count, R;
};

struct Seq {  // sequencer state (lane 0 of warp 0 of CTA 0)
  const DevSnap *s;
  const kai_config *cfg;
  const ActionParams *p;
  Replica rp;
  Tile *tile;
  Ctl *ctl;
  int n_ops, ops_cap;
  int root_len;
  int batching;
  bool is_cta0;
  long long pods_placed, pods_evicted, sweeps, nodes_scanned, n_visits, minmax_exchanges, batched;
  kai_job_visit *visits;
  int visits_cap;
  int error;
  long long t_pop, t_prep, t_scan, t_xchg, t_apply, t_finish, t_init;  // clock64 phase totals (thread 0)
  long long t_key, n_key, t_tta, t_heap;
};

__device__ __forceinline__ double &q_alloc(Seq &q, int r, int qi) { return q.rp.q_alloc[(size_t)r * q.s->Q + qi]; }
__device__ __forceinline__ double &q_alloc_np(Seq &q, int r, int qi) {
  return q.rp.q_alloc_np[(size_t)r * q.s->Q + qi];
}
__device__ __forceinline__ void invalidate_chain(Seq &q, int qi) {
  for (int c = qi; c >= 0; c = __ldg(&q.s->q_parent[c])) q.rp.qkey[c].valid = 0;
}

__device__ __forceinline__ bool job_touched(const Seq &q, int j) { return (q.rp.touched[j >> 5] >> (j & 31)) & 1u; }
__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }
// podset status counters: those of the job being allocated live in the control block
__device__ __forceinline__ int ps_get(const Seq &q, int ps, int which) {
  if (ps == q.ctl->ctx_ps) return q.ctl->ctx_cnt[which];
  return q.rp.ps_active_alloc[(size_t)which * q.s->S + ps];
}
__device__ __forceinline__ void ps_add(Seq &q, int ps, int which, int d) {
  if (ps == q.ctl->ctx_ps)
    q.ctl->ctx_cnt[which] += d;
  else
    q.rp.ps_active_alloc[(size_t)which * q.s->S + ps] += d;
}

// ---- PodInfo helpers ----
__device__ __forceinline__ bool should_allocate(const Seq &q, int t, bool real) {  // pod_info.go:518-521
  int st = q.rp.t_status[t];
  return st == KAI_POD_PENDING || (!real && st == KAI_POD_RELEASING && q.rp.t_virtual[t]);
}

// ---- node mutations (node_info.go:457-551) are queued as deltas for the scanner that owns the node ----
enum { ND_ADD = 0, ND_ADD_PIPELINED = 1, ND_ADD_RELEASING = 2, ND_REM = 3, ND_REM_PIPELINED = 4, ND_REM_RELEASING = 5 };
__device__ void seq_flush_deltas(Seq &q);  // single-threaded FLUSH exchange when the delta list is full
__device__ void emit_delta(Seq &q, int node, int code, int t) {
  Ctl &c = *q.ctl;
  if (c.n_delta >= kMaxDelta) seq_flush_deltas(q);
  unsigned long long data = (unsigned long long)(unsigned int)(node | (code << 28)) | ((unsigned long long)(unsigned int)t << 32);
  st_relaxed_b128(q.p->delta + ((size_t)(c.seq & 1) * kMaxDelta + c.n_delta) * 2, data, (unsigned long long)c.seq);
  c.n_delta++;
}
__device__ void node_add_task(Seq &q, int t, int n, int st) {  // n = task node, st = task status (just set)
  q.rp.t_node_status[t] = st;
  emit_delta(q, n, st == KAI_POD_RELEASING ? ND_ADD_RELEASING : (st == KAI_POD_PIPELINED ? ND_ADD_PIPELINED : ND_ADD), t);
}
__device__ void node_remove_task(Seq &q, int t, int n) {
  int st = q.rp.t_node_status[t];
  emit_delta(q, n, st == KAI_POD_RELEASING ? ND_REM_RELEASING : (st == KAI_POD_PIPELINED ? ND_REM_PIPELINED : ND_REM), t);
}
// applied by the owning scanner to its tile row (lane r handles resource r)
__device__ __forceinline__ void apply_delta_row(double &I, double &L, int code, double v) {
  switch (code) {
    case ND_ADD: I = __dsub_rn(I, v); break;
    case ND_ADD_PIPELINED: L = __dsub_rn(L, v); break;
    case ND_ADD_RELEASING: L = __dadd_rn(L, v); I = __dsub_rn(I, v); break;
    case ND_REM: I = __dadd_rn(I, v); break;
    case ND_REM_PIPELINED: L = __dadd_rn(L, v); break;
    case ND_REM_RELEASING: L = __dsub_rn(L, v); I = __dadd_rn(I, v); break;
  }
}

// ---- PodGroupInfo.UpdateTaskStatus (job_info.go:253-264) + podset counters ----
// job / old may be passed when the caller already knows them (saves dependent L2 loads)
__device__ void set_status(Seq &q, int t, int status, int job = -1, int old = -1) {
  Ctl &c = *q.ctl;
  if (old < 0) old = q.rp.t_status[t];
  int j = job >= 0 ? job : __ldg(&q.s->t_job[t]);
  int ps = (j == c.ctx_job && c.ctx_ps >= 0) ? c.ctx_ps : __ldg(&q.s->t_podset[t]);
  if (old & kActiveAllocated) ps_add(q, ps, 0, -1);
  if (status & kActiveAllocated) ps_add(q, ps, 0, +1);
  if (old == KAI_POD_PENDING) ps_add(q, ps, 1, -1);
  if (status == KAI_POD_PENDING) ps_add(q, ps, 1, +1);
  if (old == KAI_POD_PIPELINED) ps_add(q, ps, 2, -1);
  if (status == KAI_POD_PIPELINED) ps_add(q, ps, 2, +1);
  q.rp.t_status[t] = status;
  q.rp.j_req_valid[j] = 0;
  q.rp.touched[j >> 5] |= 1u << (j & 31);
  int qi = j == c.ctx_job ? c.ctx_queue : __ldg(&q.s->j_queue[j]);
  invalidate_chain(q, qi);  // the job may be the best pending job of its queue chain
}

// ---- proportion event handlers (proportion.go:443-489) ----
__device__ void queue_allocate(Seq &q, int t, bool add, int job = -1) {
  const DevSnap &s = *q.s;
  Ctl &c = *q.ctl;
  int j = job >= 0 ? job : __ldg(&s.t_job[t]);
  bool preemptible;
  int qi;
  if (j == c.ctx_job) {
    preemptible = c.ctx_preempt != 0;
    qi = c.ctx_queue;
  } else {
    preemptible = (__ldg(&s.j_flags[j]) & KAI_JOB_PREEMPTIBLE) != 0;
    qi = __ldg(&s.j_queue[j]);
  }
  double v[QR];
  if (t == c.dec.task)
    for (int r = 0; r < QR; r++) v[r] = c.dec.req[r];
  else
    for (int r = 0; r < QR; r++) v[r] = __ldg(&s.t_req[(size_t)t * s.R + r]);
  for (; qi >= 0; qi = __ldg(&s.q_parent[qi])) {
    for (int r = 0; r < QR; r++) {
      double &a = q_alloc(q, r, qi);
      a = add ? __dadd_rn(a, v[r]) : __dsub_rn(a, v[r]);
      if (!preemptible) {
        double &b = q_alloc_np(q, r, qi);
        b = add ? __dadd_rn(b, v[r]) : __dsub_rn(b, v[r]);
      }
    }
    q.rp.qkey[qi].valid = 0;
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
// `fresh`: the task belongs to the context job and is known to be Pending / unplaced / not virtual
__device__ void stmt_place(Seq &q, int t, int n, int kind, bool fresh) {  // :297-358 Allocate, :197-295 Pipeline
  Op op;
  op.kind = kind;
  op.task = t;
  if (fresh) {
    op.prev_status = KAI_POD_PENDING;
    op.prev_node = -1;
    op.prev_virtual = 0;
  } else {
    op.prev_status = q.rp.t_status[t];
    op.prev_node = q.rp.t_node[t];
    op.prev_virtual = q.rp.t_virtual[t];
  }
  op.next_node = n;
  op.undo_index = -1;
  op.pad = 0;
  int job = fresh ? q.ctl->ctx_job : -1;
  int st = kind == OP_ALLOCATE ? KAI_POD_ALLOCATED : KAI_POD_PIPELINED;
  set_status(q, t, st, job, op.prev_status);
  q.rp.t_node[t] = n;
  node_add_task(q, t, n, st);
  queue_allocate(q, t, true, job);
  push_op(q, op);
  q.rp.t_virtual[t] = 1;
}
__device__ void stmt_allocate(Seq &q, int t, int n, bool fresh = false) { stmt_place(q, t, n, OP_ALLOCATE, fresh); }
__device__ void stmt_pipeline(Seq &q, int t, int n, bool fresh = false) { stmt_place(q, t, n, OP_PIPELINE, fresh); }
__device__ void unallocate(Seq &q, int t, int prev_virtual) {  // :392-427
  set_status(q, t, KAI_POD_PENDING);
  node_remove_task(q, t, q.rp.t_node[t]);
  q.rp.t_node[t] = -1;
  q.rp.t_virtual[t] = (unsigned char)prev_virtual;
  queue_allocate(q, t, false);
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
__device__ void node_state_disturbed(Seq &q) {  // a node changed outside a sweep: trackers and batch are stale
  q.ctl->trk[0].dirty = q.ctl->trk[1].dirty = 1;
  q.ctl->batch.valid = 0;
}
__device__ void undo_op(Seq &q, int i) {  // :597-643 (allocate-action subset: no undo chains survive)
  Op op = q.rp.ops[i];
  if (op.kind == OP_ALLOCATE)
    unallocate(q, op.task, op.prev_virtual);
  else if (op.kind == OP_PIPELINE)
    unpipeline(q, op);
  node_state_disturbed(q);
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
  node_state_disturbed(q);
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
  int ln = ps_get(q, a, 0), rn = ps_get(q, b, 0);
  int lm = __ldg(&q.s->ps_min[a]), rm = __ldg(&q.s->ps_min[b]);
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
  int b = __ldg(&q.s->j_ps_begin[job]), e = __ldg(&q.s->j_ps_begin[job + 1]);
  if (e - b == 1) {
    out[0] = b;
    return 1;
  }
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
// :27-54 GetTasksToAllocate; result into q.rp.tta, returns count.  If sum != null only accumulates the
// request of the selected tasks (GetTasksToAllocateInitResource :87-113).
__device__ int tasks_to_allocate(Seq &q, int job, bool real, double *sum) {
  const DevSnap &s = *q.s;
  int *order = q.rp.ps_order;
  int nps = sorted_podsets(q, job, order);
  int unsat = 0;
  for (int k = 0; k < nps; k++)
    if (ps_get(q, order[k], 0) < __ldg(&s.ps_min[order[k]])) unsat++;
  int max_sets = unsat > 0 ? unsat : 1;
  int n_sets = 0, n = 0;
  if (sum) sum[0] = sum[1] = sum[2] = 0.0;
  for (int k = 0; k < nps && n_sets < max_sets; k++) {
    int ps = order[k];
    int tb = __ldg(&s.ps_task_begin[ps]), te = __ldg(&s.ps_task_begin[ps + 1]);
    int n_alloc = ps_get(q, ps, 0);
    int m = __ldg(&s.ps_min[ps]);
    int max_tasks = n_alloc >= m ? 1 : m - n_alloc;  // :144-153
    int taken = 0;
    for (int i = tb; i < te && taken < max_tasks; i++) {
      int t = i;  // tasks of a podset are stored in TaskOrderFn order
      if (!should_allocate(q, t, real)) continue;
      if (sum)
        for (int r = 0; r < QR; r++) sum[r] = __dadd_rn(sum[r], __ldg(&s.t_req[(size_t)t * s.R + r]));
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
  if (!job_touched(q, job)) {
    const JobRec *rec = q.s->jrec + job;
    if (rec->n_podsets == 1) return rec->req0;
  }
  double *c = q.rp.j_req + (size_t)job * QR;
  if (!q.rp.j_req_valid[job]) {
    tasks_to_allocate(q, job, false, c);
    q.rp.j_req_valid[job] = 1;
  }
  return c;
}
__device__ bool has_tasks_to_allocate(const Seq &q, int job) {  // :18-25 (isRealAllocation = true)
  for (int ps = __ldg(&q.s->j_ps_begin[job]); ps < __ldg(&q.s->j_ps_begin[job + 1]); ps++)
    if (ps_get(q, ps, 1) > 0) return true;
  return false;
}
// job_info.go:443-464 ShouldPipelineJob
__device__ bool should_pipeline_job(const Seq &q, int job) {
  for (int ps = __ldg(&q.s->j_ps_begin[job]); ps < __ldg(&q.s->j_ps_begin[job + 1]); ps++) {
    int pipe = ps_get(q, ps, 2);
    if (pipe > 0 && ps_get(q, ps, 0) - pipe < __ldg(&q.s->ps_min[ps])) return true;
  }
  return false;
}

// ---- capacity policy (plugins/proportion/capacity_policy) ----
__device__ bool over_capacity(Seq &q, int job, const double *req) {
  const DevSnap &s = *q.s;
  bool preemptible = (__ldg(&s.j_flags[job]) & KAI_JOB_PREEMPTIBLE) != 0;
  for (int qi = __ldg(&s.j_queue[job]); qi >= 0; qi = __ldg(&s.q_parent[qi]))
    for (int r = 0; r < QR; r++) {
      if (req[r] == 0) continue;
      double lim = __ldg(&s.q_limit[(size_t)r * s.Q + qi]);
      if (lim != KAI_UNLIMITED && lim < __dadd_rn(q_alloc(q, r, qi), req[r])) return true;
    }
  if (preemptible) return false;
  for (int qi = __ldg(&s.j_queue[job]); qi >= 0; qi = __ldg(&s.q_parent[qi]))
    for (int r = 0; r < QR; r++) {
      if (req[r] == 0) continue;
      double d = __ldg(&s.q_deserved[(size_t)r * s.Q + qi]);
      if (d != KAI_UNLIMITED && d < __dadd_rn(q_alloc_np(q, r, qi), req[r])) return true;
    }
  return false;
}

// ---- job-order tree (actions/utils/job_order_by_queue.go), one node per queue ----
__device__ __forceinline__ bool qn_is_leaf(const Seq &q, int qi) { return __ldg(&q.s->q_nchildren[qi]) == 0; }
__device__ __forceinline__ int leaf_len(const Seq &q, int qi) {
  return (q.rp.leaf_end[qi] - q.rp.leaf_head[qi]) + q.rp.ovl_len[qi];
}
__device__ __forceinline__ int qn_len(const Seq &q, int qi) {
  return qn_is_leaf(q, qi) ? leaf_len(q, qi) : q.rp.child_len[qi];
}
// the leaf priority queue: sorted run [head, end) + overflow heap for re-pushed jobs.  JobOrderFn
// (session_plugins.go:227-242: priority, elastic, creation, UID) is a strict total order on the packed key,
// so any exact priority queue pops in the same order as container/heap.
__device__ int leaf_peek(const Seq &q, int qi) {
  int h = q.rp.leaf_head[qi], e = q.rp.leaf_end[qi];
  int a = h < e ? q.rp.leaf_heap[h] : -1;
  int b = q.rp.ovl_len[qi] > 0 ? q.rp.leaf_heap[__ldg(&q.s->q_job_begin[qi])] : -1;
  if (a < 0) return b;
  if (b < 0) return a;
  return q.rp.j_key[b] < q.rp.j_key[a] ? b : a;
}
__device__ int leaf_pop(Seq &q, int qi) {
  int h = q.rp.leaf_head[qi], e = q.rp.leaf_end[qi];
  int base = __ldg(&q.s->q_job_begin[qi]);
  int a = h < e ? q.rp.leaf_heap[h] : -1;
  int n = q.rp.ovl_len[qi];
  int b = n > 0 ? q.rp.leaf_heap[base] : -1;
  bool from_ovl = a < 0 || (b >= 0 && q.rp.j_key[b] < q.rp.j_key[a]);
  if (!from_ovl) {
    q.rp.leaf_head[qi] = h + 1;
    return a;
  }
  // binary-heap pop on the overflow area
  int *it = q.rp.leaf_heap + base;
  n--;
  it[0] = it[n];
  int i = 0;
  for (;;) {
    int j1 = 2 * i + 1;
    if (j1 >= n) break;
    int j = j1;
    if (j1 + 1 < n && q.rp.j_key[it[j1 + 1]] < q.rp.j_key[it[j1]]) j = j1 + 1;
    if (!(q.rp.j_key[it[j]] < q.rp.j_key[it[i]])) break;
    int t = it[i];
    it[i] = it[j];
    it[j] = t;
    i = j;
  }
  q.rp.ovl_len[qi] = n;
  return b;
}
__device__ int elastic_class(const Seq &q, int job) {  // plugins/elastic/elastic.go:50-63
  bool exactly = true;
  for (int ps = __ldg(&q.s->j_ps_begin[job]); ps < __ldg(&q.s->j_ps_begin[job + 1]); ps++) {
    int n = ps_get(q, ps, 0), m = __ldg(&q.s->ps_min[ps]);
    if (n < m) return 0;
    if (n > m) exactly = false;
  }
  return exactly ? 1 : 2;
}
__device__ void leaf_push(Seq &q, int qi, int job) {
  q.rp.j_key[job] = make_job_key(__ldg(&q.s->j_priority[job]), elastic_class(q, job), __ldg(&q.s->j_order_rank[job]));
  int base = __ldg(&q.s->q_job_begin[qi]);
  int n = q.rp.ovl_len[qi];
  if (base + n >= q.rp.leaf_head[qi] && q.rp.leaf_head[qi] < q.rp.leaf_end[qi]) {
    q.error = 2;  // cannot happen while pushes follow pops
    return;
  }
  int *it = q.rp.leaf_heap + base;
  it[n] = job;
  int j = n;
  for (;;) {
    int i = (j - 1) / 2;
    if (i == j || !(q.rp.j_key[it[j]] < q.rp.j_key[it[i]])) break;
    int t = it[i];
    it[i] = it[j];
    it[j] = t;
    j = i;
  }
  q.rp.ovl_len[qi] = n + 1;
}
__device__ int best_job(Seq &q, int qi) {  // :283-292 getBestJobFromNode
  while (!qn_is_leaf(q, qi)) qi = q.rp.child_heap[__ldg(&q.s->q_child_begin[qi])];
  return leaf_peek(q, qi);
}

// queue_order.go:19-73 on cached per-node keys.  A key is recomputed when the queue's Allocated or its
// best pending job changed since it was last used (invalidate_chain / queue_allocate).
__device__ const QKey &queue_key(Seq &q, int qi) {
  QKey &k = q.rp.qkey[qi];
  if (k.valid) return k;
  long long tkk = clock64();
  q.n_key++;
  const DevSnap &s = *q.s;
  const double *req = job_init_resource(q, best_job(q, qi));
  bool over = true, starved = true, viol = false;
  double dj = 0.0, dr = 0.0;
  for (int r = 0; r < QR; r++) {
    size_t o = (size_t)r * s.Q + qi;
    double alloc = q.rp.q_alloc[o];
    double with_job = __dadd_rn(alloc, req[r]);
    if (__ldg(&s.q_fair[o]) >= alloc) over = false;                                   // :87-100
    if (compare_quantities(with_job, __ldg(&s.q_deserved[o])) > 0) starved = false;  // :102-128
    double la = __ldg(&s.q_allocatable[o]);
    if (la == 0 && with_job > 0) viol = true;  // :130-180
    double denom = la == KAI_UNLIMITED ? s.total[r] : la;  // queue_resource_share.go:142-166
    double vj = denom == 0 ? __dmul_rn(with_job, 1000.0) : __ddiv_rn(with_job, denom);
    double vr = denom == 0 ? __dmul_rn(alloc, 1000.0) : __ddiv_rn(alloc, denom);
    dj = fmax(dj, vj);
    dr = fmax(dr, vr);
  }
  k.over = over;
  k.starved = starved;
  k.viol = viol;
  k.drf_job = dj;
  k.drf = dr;
  k.priority = __ldg(&s.q_priority[qi]);
  k.valid = 1;
  q.t_key += clock64() - tkk;
  return k;
}
__device__ bool node_less(Seq &q, int l, int r) {  // :256-278 buildNodeOrderFn (pending order)
  if (qn_len(q, l) == 0) return true;
  if (qn_len(q, r) == 0) return false;
  const QKey kl = queue_key(q, l);
  const QKey kr = queue_key(q, r);
  if (!kl.over && kr.over) return true;
  if (kl.over && !kr.over) return false;
  if (kl.starved && !kr.starved) return true;
  if (kr.starved && !kl.starved) return false;
  if (kl.priority > kr.priority) return true;
  if (kl.priority < kr.priority) return false;
  if (kl.viol && !kr.viol) return false;
  if (!kl.viol && kr.viol) return true;
  if (kl.drf_job < kr.drf_job) return true;
  if (kl.drf_job > kr.drf_job) return false;
  if (kl.drf < kr.drf) return true;
  if (kl.drf > kr.drf) return false;
  const DevSnap &s = *q.s;
  bool l_le_r = true, r_le_l = true;  // :221-233
  for (int i = 0; i < QR; i++) {
    double la = __ldg(&s.q_allocatable[(size_t)i * s.Q + l]), ra = __ldg(&s.q_allocatable[(size_t)i * s.Q + r]);
    if (compare_quantities(la, ra) > 0) l_le_r = false;
    if (compare_quantities(ra, la) > 0) r_le_l = false;
  }
  if (!r_le_l && l_le_r) return true;
  if (!l_le_r && r_le_l) return false;
  return __ldg(&s.q_creation[l]) < __ldg(&s.q_creation[r]);  // :235-240
}
// container/heap over queue nodes
__device__ void qheap_up(Seq &q, int *items, int j) {
  for (;;) {
    int i = (j - 1) / 2;
    if (i == j || !node_less(q, items[j], items[i])) break;
    int t = items[i];
    items[i] = items[j];
    items[j] = t;
    j = i;
  }
}
__device__ bool qheap_down(Seq &q, int *items, int i0, int n) {
  int i = i0;
  for (;;) {
    int j1 = 2 * i + 1;
    if (j1 >= n || j1 < 0) break;
    int j = j1;
    int j2 = j1 + 1;
    if (j2 < n && node_less(q, items[j2], items[j1])) j = j2;
    if (!node_less(q, items[j], items[i])) break;
    int t = items[i];
    items[i] = items[j];
    items[j] = t;
    i = j;
  }
  return i > i0;
}
__device__ void qheap_push(Seq &q, int *items, int &len, int x) {
  items[len++] = x;
  qheap_up(q, items, len - 1);
}
__device__ int qheap_pop(Seq &q, int *items, int &len) {
  int n = len - 1;
  int t = items[0];
  items[0] = items[n];
  items[n] = t;
  qheap_down(q, items, 0, n);
  len = n;
  return items[n];
}
__device__ void mark_ancestors(Seq &q, int qi) {  // :246-250 (+ key invalidation: best job / heap tops changed)
  for (int c = qi; c >= 0; c = __ldg(&q.s->q_parent[c])) {
    q.rp.qn_flags[c] |= QN_REORDER;
    q.rp.qkey[c].valid = 0;
  }
}
__device__ void ensure_chain(Seq &q, int child) {  // :135-175
  for (;;) {
    int p = __ldg(&q.s->q_parent[child]);
    if (p < 0) {
      if (!(q.rp.qn_flags[child] & QN_LINKED)) {
        qheap_push(q, q.rp.root_heap, q.root_len, child);
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
      qheap_push(q, q.rp.child_heap + __ldg(&q.s->q_child_begin[p]), q.rp.child_len[p], child);
      q.rp.qn_flags[child] |= QN_LINKED;
      invalidate_chain(q, p);
    }
    if (!is_new) return;
    child = p;
  }
}
__device__ void push_job(Seq &q, int job) {  // :90-119
  int qi = __ldg(&q.s->j_queue[job]);
  if (!qn_is_leaf(q, qi)) return;
  bool needs_linking = !(q.rp.qn_flags[qi] & QN_EXISTS);
  if (needs_linking) q.rp.qn_flags[qi] = QN_EXISTS;
  leaf_push(q, qi, job);
  invalidate_chain(q, qi);
  if (needs_linking) ensure_chain(q, qi);
  mark_ancestors(q, qi);
}
// owner = queue whose children heap `items` is (or -1 for the root heap)
__device__ int get_next_node(Seq &q, int *items, int &len, int owner) {  // :193-215
  for (;;) {
    if (len == 0) return -1;
    int ni = items[0];
    if (q.rp.qn_flags[ni] & QN_REORDER) {
      if (!qheap_down(q, items, 0, len)) qheap_up(q, items, 0);  // heap.Fix(0)
      q.rp.qn_flags[ni] &= ~QN_REORDER;
      if (owner >= 0) invalidate_chain(q, owner);
      continue;
    }
    if (qn_len(q, ni) == 0) return -1;
    return ni;
  }
}
__device__ void handle_pop(Seq &q, int qi) {  // :219-243
  for (;;) {
    if (qn_len(q, qi) == 0) {
      int p = __ldg(&q.s->q_parent[qi]);
      if (p >= 0) {
        qheap_pop(q, q.rp.child_heap + __ldg(&q.s->q_child_begin[p]), q.rp.child_len[p]);
        invalidate_chain(q, p);
      } else {
        qheap_pop(q, q.rp.root_heap, q.root_len);
      }
      q.rp.qn_flags[qi] = 0;
      q.rp.qkey[qi].valid = 0;
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
  int ni = get_next_node(q, q.rp.root_heap, q.root_len, -1);
  while (ni >= 0 && !qn_is_leaf(q, ni))
    ni = get_next_node(q, q.rp.child_heap + __ldg(&q.s->q_child_begin[ni]), q.rp.child_len[ni], ni);
  if (ni < 0) return -1;
  int job = leaf_pop(q, ni);
  {  // warm L1 for the next pops of this queue
    int h = q.rp.leaf_head[ni], e = q.rp.leaf_end[ni];
    if (h < e) {
      const JobRec *r1 = q.s->jrec + q.rp.leaf_heap[h];
      prefetch_l1(r1);
      if (h + 1 < e) prefetch_l1(q.s->jrec + q.rp.leaf_heap[h + 1]);
    }
  }
  invalidate_chain(q, ni);
  handle_pop(q, ni);
  return job;
}

// ---- min/max trackers ----
// update after a placement that lowered NonAllocated(res) of a node from b to a (a < b)
__device__ __forceinline__ void track_decrease(Track &t, uint32_t f, double a) {
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
__device__ __forceinline__ uint32_t track_flags(const Track &t, double b, double a) {
  uint32_t f = 0;
  if (b == t.mx) f |= WF_B_EQ_MX;
  if (a < t.mn)
    f |= WF_A_LT_MN;
  else if (a == t.mn)
    f |= WF_A_EQ_MN;
  return f;
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
  if (ra == kRankNone) return false;
  if (rb == kRankNone) return true;
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

// FittingNode (session.go:201-232) + NodeOrderFn sum (session_plugins.go:427-437) of one node row given as
// Idle/Releasing vectors.  Returns false if the node does not fit; fit_i = fits on Idle alone.
__device__ __forceinline__ bool node_key(const Decision &d, int R, const double *I, const double *L, int stride,
                                         double a_gpu, double a_cpu, double gpu_count, uint32_t nflags, int n,
                                         double &score, bool &fit_i) {
  bool fit_ri = true;
  fit_i = true;
  for (int r = 0; r < R; r++) {
    double i = I[r * stride];
    double avail = __dadd_rn(i, L[r * stride]);
    double rq = d.req[r];
    if (r >= 3) {
      if (rq != 0 && rq > avail) fit_ri = false;
      if (rq != 0 && rq > i) fit_i = false;
    } else {
      if (rq > avail) fit_ri = false;
      if (rq > i) fit_i = false;
    }
  }
  if (!fit_ri) return false;
  score = 0.0;
  score = __dadd_rn(score, (d.best_effort || fit_i) ? 100.0 : 0.0);  // nodeavailability.go:29-40
  score = __dadd_rn(score, 0.0);                                     // gpusharingorder (whole GPUs)
  bool cpu_only_node = !(nflags & KAI_NODE_NOT_CPU_ONLY) && a_gpu <= 0;
  score = __dadd_rn(score, (!d.gpu_task && cpu_only_node) ? 10.0 : 0.0);  // resourcetype.go:29-41
  score = __dadd_rn(score, (d.nominated == n) ? 1000000.0 : 0.0);        // nominatednode.go:29-41
  double cur = __dadd_rn(I[d.res * stride], L[d.res * stride]);
  double overall = d.res == KAI_RES_GPU ? a_gpu : a_cpu;
  double place;
  if (d.strategy == KAI_PLACEMENT_BINPACK) {
    place = binpack_score(d.mn, d.mx, cur, overall);
  } else {  // spread.go:16-36
    double cnt = d.res == KAI_RES_GPU ? (double)(long long)gpu_count : overall;
    place = cnt == 0 ? 0.0 : __ddiv_rn(cur, cnt);
  }
  score = __dadd_rn(score, place);
  return true;
}

// The sweep over this CTA's tile followed by the block argmax on (score desc, name rank asc).
__device__ Cand scan_tile(const Tile &tl, const Decision &d, const DevSnap &s, Cand *sh_warp) {
  Cand best;
  best.score = -1.0;
  best.rank = kRankNone;
  best.ln = -1;
  const uint32_t *mask = d.pred_class >= 0 ? s.pred_mask + (size_t)d.pred_class * s.mask_words : nullptr;
  for (int ln = threadIdx.x; ln < tl.count; ln += blockDim.x) {
    int n = tl.base + ln;
    if (mask && !((__ldg(&mask[n >> 5]) >> (n & 31)) & 1u)) continue;
    double score;
    bool fit_i;
    if (!node_key(d, tl.R, tl.I + ln, tl.L + ln, tl.npc, tl.Agpu[ln], tl.Acpu[ln], tl.gpu_count[ln], tl.flags[ln], n,
                  score, fit_i))
      continue;
    uint32_t rk = (uint32_t)tl.rank[ln];
    if (better(score, rk, best.score, best.rank)) {
      best.score = score;
      best.rank = rk;
      best.ln = ln;
    }
  }
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
      c.rank = kRankNone;
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

// slot layout per CTA and parity (8 x u64):
//   A {score bits, [tag:24][flags:8][repeat:8][rank:24]}
//   B {cur_a gpu bits, tag}   C {cur_a cpu bits, tag}   D {repeat tracker-event bits, tag}
constexpr int kSlotWords = 8;

// candidate of this CTA -> slot words, including the same-node repeat analysis (lane 0 of warp 0)
__device__ void publish_candidate(const Track *trk, const Tile &tl, const Decision &d, Cand local,
                                  unsigned long long *slot, unsigned int tag, int batching) {
  uint32_t flags = 0, repeat = 0;
  double a_gpu = 0, a_cpu = 0;
  unsigned long long rep_flags = 0;
  if (local.rank != kRankNone) {
    const int ln = local.ln, R = tl.R, n = tl.base + ln;
    double I[KAI_MAX_RES], L[KAI_MAX_RES];
    for (int r = 0; r < R; r++) {
      I[r] = tl.I[r * tl.npc + ln];
      L[r] = tl.L[r * tl.npc + ln];
    }
    const double ag = tl.Agpu[ln], ac = tl.Acpu[ln], gc = tl.gpu_count[ln];
    const uint32_t nf = tl.flags[ln];
    bool fit_i = true;
    for (int r = 0; r < R; r++) {
      double rq = d.req[r];
      if (r >= 3 ? (rq != 0 && rq > I[r]) : (rq > I[r])) fit_i = false;
    }
    const bool to_idle = !d.pipeline_only && (d.best_effort || fit_i);  // common/allocate.go:165-174
    if (to_idle) flags |= SLOT_TO_IDLE;
    Track sim[2] = {trk[0], trk[1]};
    bool stop = false;
    // placement 0 is the swept one; placements 1..kMaxRepeat are candidate repeats on the same node
    for (int rep = 0; rep <= kMaxRepeat; rep++) {
      if (rep > 0) {
        if (!batching || stop) break;
        double sc;
        bool fi;
        if (!node_key(d, R, I, L, 1, ag, ac, gc, nf, n, sc, fi)) break;
        bool ti = !d.pipeline_only && (d.best_effort || fi);
        if (ti != to_idle) break;
        if (!(sc >= local.score)) break;  // node n must stay the argmax
      }
      uint32_t f6 = 0;
      double a2[2] = {0, 0};
      for (int k = 0; k < 2; k++) {
        int res = k == 0 ? KAI_RES_GPU : KAI_RES_CPU;
        double overall = k == 0 ? ag : ac;
        if (overall == 0 || d.req[res] == 0) continue;
        double b = __dadd_rn(I[res], L[res]);
        double a = to_idle ? __dadd_rn(__dsub_rn(I[res], d.req[res]), L[res]) : __dadd_rn(I[res], __dsub_rn(L[res], d.req[res]));
        uint32_t f = sim[k].dirty ? 0u : track_flags(sim[k], b, a);
        f6 |= f << (3 * k);
        a2[k] = a;
      }
      if (rep > 0 && (f6 & ((WF_A_LT_MN) | (WF_A_LT_MN << 3)))) break;  // a new minimum needs its value: sweep instead
      // does this placement leave (mn, mx) of the scored resource intact for the following repeats?
      for (int k = 0; k < 2; k++) {
        uint32_t f = (f6 >> (3 * k)) & 7u;
        bool scored = (k == 0) == (d.res == KAI_RES_GPU);
        if (f) {
          Track before = sim[k];
          track_decrease(sim[k], f, a2[k]);
          if (scored && d.strategy == KAI_PLACEMENT_BINPACK &&
              (sim[k].dirty || sim[k].mn != before.mn || sim[k].mx != before.mx))
            stop = true;
        }
      }
      if (rep == 0) {
        flags |= f6 & 0x3fu;
        a_gpu = a2[0];
        a_cpu = a2[1];
      } else {
        rep_flags |= (unsigned long long)(f6 & 0x3fu) << (6 * (rep - 1));
        repeat = rep;
      }
      // apply the placement to the simulated node row (node_info.go:457-493)
      for (int r = 0; r < R; r++) {
        if (to_idle)
          I[r] = __dsub_rn(I[r], d.req[r]);
        else
          L[r] = __dsub_rn(L[r], d.req[r]);
      }
    }
    if (repeat) flags |= SLOT_HAS_REPEAT;
  }
  st_relaxed_b128(slot + 2, (unsigned long long)__double_as_longlong(a_gpu), (unsigned long long)tag);
  st_relaxed_b128(slot + 4, (unsigned long long)__double_as_longlong(a_cpu), (unsigned long long)tag);
  if (repeat) st_relaxed_b128(slot + 6, rep_flags, (unsigned long long)tag);
  unsigned long long hi = ((unsigned long long)tag << 40) | ((unsigned long long)(flags & 0xffu) << 32) |
                          ((unsigned long long)(repeat & 0xffu) << 24) | (unsigned long long)(local.rank & 0xffffffu);
  st_relaxed_b128(slot, (unsigned long long)__double_as_longlong(local.score), hi);
}


// =============================================================================================
// sequencer <-> scanner protocol
// =============================================================================================
// Watchdog for the spin waits: a wait that does not complete within ~2^22 polls records (code, seq, who)
// in counters[24..27], raises the abort flag and lets every waiter fall through so that the kernel ends
// and the host reports KAI_ERR_CUDA instead of hanging the GPU.
struct Spin {
  unsigned int n = 0;
  __device__ __forceinline__ bool expired(const ActionParams &p, int code, unsigned int seq, int who) {
    if ((++n & 0x3ffu) != 0) return false;
    volatile long long *c = p.counters;
    if (c[24] != 0) return true;
    if (n >= (1u << 22)) {
      if (atomicCAS((unsigned long long *)&p.counters[24], 0ull, (unsigned long long)code) == 0ull) {
        c[25] = seq;
        c[26] = who;
        c[27] = blockIdx.x;
      }
      return true;
    }
    return false;
  }
};
// decision record words (each stored as {data, tag}):
//   0  kind | res<<8 | strategy<<16 | bits<<24 | n_delta<<32      1  nominated | pred_class<<32
//   2..9 req[0..7]      10,11 gpu tracker mn,mx      12,13 cpu tracker mn,mx
//   14 gpu cnt_mn | cnt_mx<<32      15 cpu cnt_mn | cnt_mx<<32
__device__ void build_decision_words(Ctl &c, int kind, int batching) {
  const Decision &d = c.dec;
  unsigned long long bits = (d.gpu_task ? DB_GPU_TASK : 0) | (d.best_effort ? DB_BEST_EFFORT : 0) |
                            (d.pipeline_only ? DB_PIPELINE_ONLY : 0) | (batching ? DB_BATCHING : 0) |
                            (c.trk[0].dirty ? DB_DIRTY0 : 0) | (c.trk[1].dirty ? DB_DIRTY1 : 0);
  c.dw[0] = (unsigned long long)kind | ((unsigned long long)d.res << 8) | ((unsigned long long)d.strategy << 16) |
            (bits << 24) | ((unsigned long long)c.n_delta << 32);
  c.dw[1] = (unsigned long long)(unsigned int)d.nominated | ((unsigned long long)(unsigned int)d.pred_class << 32);
  for (int r = 0; r < KAI_MAX_RES; r++) c.dw[2 + r] = (unsigned long long)__double_as_longlong(d.req[r]);
  for (int k = 0; k < 2; k++) {
    c.dw[10 + 2 * k] = (unsigned long long)__double_as_longlong(c.trk[k].mn);
    c.dw[11 + 2 * k] = (unsigned long long)__double_as_longlong(c.trk[k].mx);
    c.dw[14 + k] = (unsigned long long)(unsigned int)c.trk[k].cnt_mn | ((unsigned long long)(unsigned int)c.trk[k].cnt_mx << 32);
  }
}

// Publish the next decision record (warp 0 of CTA 0; lane 0 has prepared ctl.dec / trackers / deltas).
__device__ void seq_publish(const ActionParams &p, Ctl &ctl, int kind) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) {
    build_decision_words(ctl, kind, p.batching);  // deltas are self-validating tagged words: no fence
  }
  __syncwarp();
  unsigned long long *rec = p.dbuf + (size_t)(ctl.seq & 1) * kDecWords * 2;
  if (lane < kDecWords) st_relaxed_b128(rec + 2 * lane, ctl.dw[lane], (unsigned long long)ctl.seq);
  __syncwarp();
}

// Gather the candidate slots of all scanners (warp 0 of CTA 0) and let lane 0 digest the winner.
__device__ void seq_gather_candidates(const ActionParams &p, Ctl &ctl) {
  const int lane = threadIdx.x & 31;
  const unsigned int seq = ctl.seq;
  unsigned long long *buf = p.xbuf + (size_t)(seq & 1) * kMaxGrid * kSlotWords;
  const unsigned int tag = seq & 0xffffffu;
  double bs = -1.0;
  uint32_t brank = kRankNone, bmeta = 0;
  int bslot = -1;
  for (int c = lane; c < p.grid - 1; c += 32) {
    const unsigned long long *slot = buf + (size_t)c * kSlotWords;
    unsigned long long lo, hi;
    {
      Spin spin;
      do {
        ld_relaxed_b128(slot, lo, hi);
      } while (((unsigned int)(hi >> 40) != tag) && !spin.expired(p, 1, (unsigned int)seq, (int)(c)));
    }
    double sc = __longlong_as_double((long long)lo);
    uint32_t rk = (uint32_t)(hi & 0xffffffu);
    if (better(sc, rk, bs, brank)) {
      bs = sc;
      brank = rk;
      bmeta = (uint32_t)((hi >> 24) & 0xffffu);  // [flags:8][repeat:8]
      bslot = c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    double os = __shfl_xor_sync(0xffffffffu, bs, o);
    uint32_t orank = __shfl_xor_sync(0xffffffffu, brank, o);
    uint32_t om = __shfl_xor_sync(0xffffffffu, bmeta, o);
    int osl = __shfl_xor_sync(0xffffffffu, bslot, o);
    if (better(os, orank, bs, brank)) {
      bs = os;
      brank = orank;
      bmeta = om;
      bslot = osl;
    }
  }
  if (lane == 0) {
    uint32_t bflags = bmeta >> 8, repeat = bmeta & 0xffu;
    ctl.win.score = bs;
    ctl.win.rank = brank;
    ctl.win.flags = bflags;
    ctl.win.node = brank == kRankNone ? -1 : __ldg(&p.s.rank_to_node[brank]);
    ctl.batch.valid = 0;
    if (brank != kRankNone) {
      const unsigned long long *slot = buf + (size_t)bslot * kSlotWords;
      for (int k = 0; k < 2; k++) {
        uint32_t f = (bflags >> (3 * k)) & 7u;
        double a = 0;
        if (f & WF_A_LT_MN) {  // a new global minimum: fetch its value
          unsigned long long lo, hi;
          {
            Spin spin;
            do {
              ld_relaxed_b128(slot + 2 + 2 * k, lo, hi);
            } while (((unsigned int)hi != tag) && !spin.expired(p, 2, (unsigned int)seq, (int)(bslot)));
          }
          a = __longlong_as_double((long long)lo);
        }
        if (f) track_decrease(ctl.trk[k], f, a);
      }
      if (repeat) {
        unsigned long long lo, hi;
        {
          Spin spin;
          do {
            ld_relaxed_b128(slot + 6, lo, hi);
          } while (((unsigned int)hi != tag) && !spin.expired(p, 3, (unsigned int)seq, (int)(bslot)));
        }
        ctl.batch.valid = 1;
        ctl.batch.node = ctl.win.node;
        ctl.batch.to_idle = (bflags & SLOT_TO_IDLE) ? 1 : 0;
        ctl.batch.left = (int)repeat;
        ctl.batch.idx = 0;
        ctl.batch.fl = lo;
      }
    }
    ctl.seq = seq + 1;
    ctl.n_delta = 0;
  }
  __syncwarp();
}

// min/max answer slots: four tagged words {value, [tag:32][count:32]} = gpu mn, gpu mx, cpu mn, cpu mx
__device__ void seq_gather_minmax(const ActionParams &p, Ctl &ctl) {
  const int lane = threadIdx.x & 31;
  const unsigned int seq = ctl.seq;
  unsigned long long *buf = p.mmbuf + (size_t)(seq & 1) * kMaxGrid * kSlotWords;
  const unsigned long long tag = seq;
  double gmn[2] = {DBL_MAX, DBL_MAX}, gmx[2] = {0, 0};
  long long cmn[2] = {0, 0}, cmx[2] = {0, 0};
  for (int cta = lane; cta < p.grid - 1; cta += 32) {
    const unsigned long long *slot = buf + (size_t)cta * kSlotWords;
    for (int k = 0; k < 2; k++) {
      unsigned long long lo, hi;
      {
        Spin spin;
        do {
          ld_relaxed_b128(slot + 4 * k, lo, hi);
        } while (((hi >> 32) != (tag & 0xffffffffu)) && !spin.expired(p, 4, (unsigned int)seq, (int)(cta)));
      }
      double v = __longlong_as_double((long long)lo);
      int cnt = (int)(hi & 0xffffffffu);
      if (cnt > 0) {
        if (cmn[k] == 0 || v < gmn[k]) {
          gmn[k] = v;
          cmn[k] = cnt;
        } else if (v == gmn[k])
          cmn[k] += cnt;
      }
      {
        Spin spin;
        do {
          ld_relaxed_b128(slot + 4 * k + 2, lo, hi);
        } while (((hi >> 32) != (tag & 0xffffffffu)) && !spin.expired(p, 5, (unsigned int)seq, (int)(cta)));
      }
      v = __longlong_as_double((long long)lo);
      cnt = (int)(hi & 0xffffffffu);
      if (cnt > 0) {
        if (cmx[k] == 0 || v > gmx[k]) {
          gmx[k] = v;
          cmx[k] = cnt;
        } else if (v == gmx[k])
          cmx[k] += cnt;
      }
    }
  }
  for (int k = 0; k < 2; k++)
    for (int o = 16; o > 0; o >>= 1) {
      double omn = __shfl_xor_sync(0xffffffffu, gmn[k], o);
      long long ocmn = __shfl_xor_sync(0xffffffffu, cmn[k], o);
      double omx = __shfl_xor_sync(0xffffffffu, gmx[k], o);
      long long ocmx = __shfl_xor_sync(0xffffffffu, cmx[k], o);
      if (ocmn > 0) {
        if (cmn[k] == 0 || omn < gmn[k]) {
          gmn[k] = omn;
          cmn[k] = ocmn;
        } else if (omn == gmn[k])
          cmn[k] += ocmn;
      }
      if (ocmx > 0) {
        if (cmx[k] == 0 || omx > gmx[k]) {
          gmx[k] = omx;
          cmx[k] = ocmx;
        } else if (omx == gmx[k])
          cmx[k] += ocmx;
      }
    }
  if (lane == 0) {
    for (int k = 0; k < 2; k++) {  // pack.go:66-86: min starts at MaxFloat64, max at 0
      ctl.trk[k].mn = cmn[k] > 0 ? gmn[k] : DBL_MAX;
      ctl.trk[k].mx = (cmx[k] > 0 && gmx[k] > 0) ? gmx[k] : 0.0;
      ctl.trk[k].cnt_mn = (int)cmn[k];
      ctl.trk[k].cnt_mx = (int)cmx[k];
      ctl.trk[k].dirty = 0;
    }
    ctl.seq = seq + 1;
    ctl.n_delta = 0;
  }
  __syncwarp();
}

// FLUSH issued by lane 0 alone from inside sequential code (delta list full; rare)
__device__ void seq_flush_deltas(Seq &q) {
  const ActionParams &p = *q.p;
  Ctl &c = *q.ctl;
  build_decision_words(c, DK_FLUSH, 0);
  unsigned long long *rec = p.dbuf + (size_t)(c.seq & 1) * kDecWords * 2;
  for (int i = 0; i < kDecWords; i++) st_relaxed_b128(rec + 2 * i, c.dw[i], (unsigned long long)c.seq);
  unsigned long long *buf = p.xbuf + (size_t)(c.seq & 1) * kMaxGrid * kSlotWords;
  const unsigned int tag = c.seq & 0xffffffu;
  for (int cta = 0; cta < p.grid - 1; cta++) {
    unsigned long long lo, hi;
    {
      Spin spin;
      do {
        ld_relaxed_b128(buf + (size_t)cta * kSlotWords, lo, hi);
      } while (((unsigned int)(hi >> 40) != tag) && !spin.expired(p, 6, (unsigned int)c.seq, (int)(cta)));
    }
  }
  c.seq++;
  c.n_delta = 0;
}

// =============================================================================================
// sequencer steps (lane 0)
// =============================================================================================
// InitializeWithJobs (input_jobs.go:21-68) in canonical order: leaf queues ascending, jobs of a queue in
// JobOrderFn order (the Go map order is unspecified).
__device__ void seq_init_job_order(Seq &q) {
  const DevSnap &s = *q.s;
  for (int qi = 0; qi < s.Q; qi++) {
    if (__ldg(&s.q_nchildren[qi]) != 0) continue;
    if (leaf_len(q, qi) == 0) continue;
    q.rp.qn_flags[qi] = QN_EXISTS;
    ensure_chain(q, qi);
    mark_ancestors(q, qi);
  }
}

// builds ctl.dec for task t of `job`; returns false when the task cannot be placed at all
__device__ bool seq_prepare_task(Seq &q, int t, int job) {
  const DevSnap &s = *q.s;
  Ctl &c = *q.ctl;
  double rq[KAI_MAX_RES];
  for (int r = 0; r < KAI_MAX_RES; r++) rq[r] = r < s.R ? __ldg(&s.t_req[(size_t)t * s.R + r]) : 0.0;
  int nominated = s.t_nominated ? __ldg(&s.t_nominated[t]) : -1;
  int pred_class = s.t_pred_class ? __ldg(&s.t_pred_class[t]) : -1;
  bool gpu_task = rq[KAI_RES_GPU] > 0;
  // predicates.go:196-200 -> capacity_policy.go:51-61 with node_info.go:734-744
  double creq[QR] = {rq[KAI_RES_CPU], rq[KAI_RES_MEM], gpu_task ? 1.0 : 0.0};
  if (over_capacity(q, job, creq)) return false;
  bool empty = !(rq[KAI_RES_GPU] > 0.01) && !(rq[KAI_RES_CPU] >= 10) && !(rq[KAI_RES_MEM] >= 10.0 * 1024 * 1024);
  for (int r = 3; r < s.R; r++)
    if (rq[r] >= 10) empty = false;
  int strategy = gpu_task ? q.cfg->gpu_placement : q.cfg->cpu_placement;
  Decision &d = c.dec;
  // same request/flags as the previous sweep and the owner vouched for more placements on the same node?
  bool same = c.batch.valid && c.batch.left > 0 && d.gpu_task == (int)gpu_task && d.nominated == nominated &&
              d.pred_class == pred_class && d.best_effort == (int)empty && d.strategy == strategy &&
              d.pipeline_only == 0;
  if (same)
    for (int r = 0; r < KAI_MAX_RES; r++)
      if (d.req[r] != rq[r]) same = false;
  c.use_batch = same ? 1 : 0;
  c.need_minmax = 0;
  d.task = t;
  if (same) return true;
  c.batch.valid = 0;
  for (int r = 0; r < KAI_MAX_RES; r++) d.req[r] = rq[r];
  d.gpu_task = gpu_task;
  d.res = gpu_task ? KAI_RES_GPU : KAI_RES_CPU;
  d.strategy = strategy;
  d.pipeline_only = 0;
  d.nominated = nominated;
  d.pred_class = pred_class;
  d.best_effort = empty;
  c.need_minmax = (d.strategy == KAI_PLACEMENT_BINPACK) && c.trk[gpu_task ? 0 : 1].dirty;
  return true;
}

__device__ void seq_apply_winner(Seq &q, int t) {
  Ctl &c = *q.ctl;
  q.sweeps++;
  q.nodes_scanned += q.s->N;
  if (c.win.node < 0) {
    c.item_ok = 0;
    return;
  }
  if (c.win.flags & SLOT_TO_IDLE)
    stmt_allocate(q, t, c.win.node, c.ctx_fresh != 0);
  else
    stmt_pipeline(q, t, c.win.node, c.ctx_fresh != 0);
  c.item_ok = 1;
}
__device__ void seq_apply_batched(Seq &q, int t) {
  Ctl &c = *q.ctl;
  Batch &b = c.batch;
  uint32_t f6 = (uint32_t)((b.fl >> (6 * b.idx)) & 0x3fu);
  for (int k = 0; k < 2; k++) {
    uint32_t f = (f6 >> (3 * k)) & 7u;
    if (f) track_decrease(c.trk[k], f, 0.0);
  }
  b.idx++;
  b.left--;
  if (b.to_idle)
    stmt_allocate(q, t, b.node, c.ctx_fresh != 0);
  else
    stmt_pipeline(q, t, b.node, c.ctx_fresh != 0);
  q.batched++;
  c.item_ok = 1;
}

__device__ void record_visit(Seq &q, int job, int outcome) {
  if (q.n_visits < q.visits_cap) {
    q.visits[q.n_visits].job = job;
    q.visits[q.n_visits].outcome = outcome;
  }
  q.n_visits++;
}

// =============================================================================================
// scanner CTA
// =============================================================================================
struct ScanShared {
  unsigned long long dw[kDecWords];
  Decision dec;
  Track trk[2];
  int kind, n_delta, batching;
  int2 delta[kMaxDelta];
  unsigned char mine[kMaxDelta];
  double dreq[kMaxDelta][KAI_MAX_RES];
};

__device__ void scanner_main(const ActionParams &p, unsigned char *smem, Cand *sh_warp, double *sh_d, int *sh_i,
                             ScanShared &sh, Tile &tile) {
  const DevSnap &s = p.s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int my = blockIdx.x - 1;
  if (tid == 0) {
    int npc = p.nodes_per_cta;
    unsigned char *ptr = smem;
    tile.npc = npc;
    tile.R = s.R;
    tile.base = p.node_base + my * npc;
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
  }
  __syncthreads();
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
  __syncthreads();
  unsigned int seq = p.seq0;
  for (;;) {
    // ---- wait for decision record `seq` ----
    if (warp == 0) {
      const unsigned long long *rec0 = p.dbuf + (size_t)(seq & 1) * kDecWords * 2;
      if (lane == 0) {  // one poller per CTA on word 0 keeps the record's L2 lines cool
        unsigned long long lo, hi;
        Spin spin;
        for (;;) {
          ld_relaxed_b128(rec0, lo, hi);
          if (hi == (unsigned long long)seq || spin.expired(p, 7, seq, 0)) break;
          __nanosleep(20);
        }
      }
      __syncwarp();
      if (lane < kDecWords) {  // every word is self-validating
        unsigned long long lo, hi;
        {
          Spin spin;
          do {
            ld_relaxed_b128(rec0 + 2 * lane, lo, hi);
          } while ((hi != (unsigned long long)seq) && !spin.expired(p, 8, (unsigned int)seq, (int)(lane)));
        }
        sh.dw[lane] = lo;
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long w0 = sh.dw[0];
      Decision &d = sh.dec;
      sh.kind = (int)(w0 & 0xff);
      d.res = (int)((w0 >> 8) & 0xff);
      d.strategy = (int)((w0 >> 16) & 0xff);
      unsigned int bits = (unsigned int)((w0 >> 24) & 0xff);
      sh.n_delta = (int)((w0 >> 32) & 0xffff);
      d.gpu_task = (bits & DB_GPU_TASK) ? 1 : 0;
      d.best_effort = (bits & DB_BEST_EFFORT) ? 1 : 0;
      d.pipeline_only = (bits & DB_PIPELINE_ONLY) ? 1 : 0;
      sh.batching = (bits & DB_BATCHING) ? 1 : 0;
      d.nominated = (int)(unsigned int)(sh.dw[1] & 0xffffffffu);
      d.pred_class = (int)(unsigned int)(sh.dw[1] >> 32);
      for (int r = 0; r < KAI_MAX_RES; r++) d.req[r] = __longlong_as_double((long long)sh.dw[2 + r]);
      for (int k = 0; k < 2; k++) {
        sh.trk[k].mn = __longlong_as_double((long long)sh.dw[10 + 2 * k]);
        sh.trk[k].mx = __longlong_as_double((long long)sh.dw[11 + 2 * k]);
        sh.trk[k].cnt_mn = (int)(unsigned int)(sh.dw[14 + k] & 0xffffffffu);
        sh.trk[k].cnt_mx = (int)(unsigned int)(sh.dw[14 + k] >> 32);
        sh.trk[k].dirty = (bits & (k == 0 ? DB_DIRTY0 : DB_DIRTY1)) ? 1 : 0;
      }
      int tk = d.res == KAI_RES_GPU ? 0 : 1;
      d.mn = sh.trk[tk].mn;
      d.mx = sh.trk[tk].mx;
      d.task = -1;
    }
    __syncthreads();
    // ---- apply the node deltas that belong to this tile (loads in parallel, application in list order) ----
    const int nd = sh.n_delta;
    if (nd > 0) {
      const unsigned long long *dl = p.delta + (size_t)(seq & 1) * kMaxDelta * 2;
      for (int e = tid; e < nd; e += blockDim.x) {
        unsigned long long lo, hi;
        {
          Spin spin;
          do {
            ld_relaxed_b128(dl + 2 * e, lo, hi);
          } while ((hi != (unsigned long long)seq) && !spin.expired(p, 9, (unsigned int)seq, (int)(e)));
        }
        int2 en = make_int2((int)(unsigned int)(lo & 0xffffffffu), (int)(unsigned int)(lo >> 32));
        sh.delta[e] = en;
        int node = en.x & 0x0fffffff;
        int ln = node - tile.base;
        bool mine = ln >= 0 && ln < tile.count;
        sh.mine[e] = mine ? 1 : 0;
        if (mine)
          for (int r = 0; r < s.R; r++) sh.dreq[e][r] = __ldg(&s.t_req[(size_t)en.y * s.R + r]);
      }
      __syncthreads();
      if (warp == 0 && lane < s.R) {
        for (int e = 0; e < nd; e++) {
          if (!sh.mine[e]) continue;
          int2 en = sh.delta[e];
          int ln = (en.x & 0x0fffffff) - tile.base;
          apply_delta_row(tile.I[lane * tile.npc + ln], tile.L[lane * tile.npc + ln], (en.x >> 28) & 7, sh.dreq[e][lane]);
        }
      }
      __syncthreads();
    }
    const int kind = sh.kind;
    if (kind == DK_DONE || ((volatile long long *)p.counters)[24] != 0) break;
    unsigned long long *slot = p.xbuf + (size_t)(seq & 1) * kMaxGrid * kSlotWords + (size_t)my * kSlotWords;
    if (kind == DK_SCAN) {
      Cand local = scan_tile(tile, sh.dec, s, sh_warp);
      if (tid == 0) publish_candidate(sh.trk, tile, sh.dec, local, slot, seq & 0xffffffu, sh.batching);
    } else if (kind == DK_MINMAX) {
      double mn[2] = {DBL_MAX, DBL_MAX}, mx[2] = {0, 0};
      for (int ln = tid; ln < tile.count; ln += blockDim.x)
        for (int k = 0; k < 2; k++) {
          int res = k == 0 ? KAI_RES_GPU : KAI_RES_CPU;
          double overall = k == 0 ? tile.Agpu[ln] : tile.Acpu[ln];
          if (overall == 0) continue;
          double cur = __dadd_rn(tile.I[res * tile.npc + ln], tile.L[res * tile.npc + ln]);
          if (cur < mn[k]) mn[k] = cur;
          if (cur > mx[k]) mx[k] = cur;
        }
      for (int k = 0; k < 2; k++)
        for (int o = 16; o > 0; o >>= 1) {
          mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
          mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
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
      int c[4] = {0, 0, 0, 0};
      for (int ln = tid; ln < tile.count; ln += blockDim.x)
        for (int k = 0; k < 2; k++) {
          int res = k == 0 ? KAI_RES_GPU : KAI_RES_CPU;
          double overall = k == 0 ? tile.Agpu[ln] : tile.Acpu[ln];
          if (overall == 0) continue;
          double cur = __dadd_rn(tile.I[res * tile.npc + ln], tile.L[res * tile.npc + ln]);
          if (cur == mn[k]) c[2 * k]++;
          if (cur == mx[k]) c[2 * k + 1]++;
        }
      for (int i = 0; i < 4; i++)
        for (int o = 16; o > 0; o >>= 1) c[i] += __shfl_xor_sync(0xffffffffu, c[i], o);
      if (lane == 0)
        for (int i = 0; i < 4; i++) sh_i[warp * 4 + i] = c[i];
      __syncthreads();
      if (tid == 0) {
        int tot[4] = {0, 0, 0, 0};
        for (int w = 0; w < nw; w++)
          for (int i = 0; i < 4; i++) tot[i] += sh_i[w * 4 + i];
        unsigned long long tag = seq;
        slot = p.mmbuf + (size_t)(seq & 1) * kMaxGrid * kSlotWords + (size_t)my * kSlotWords;
        st_relaxed_b128(slot + 0, (unsigned long long)__double_as_longlong(mn[0]), (tag << 32) | (unsigned int)tot[0]);
        st_relaxed_b128(slot + 2, (unsigned long long)__double_as_longlong(mx[0]), (tag << 32) | (unsigned int)tot[1]);
        st_relaxed_b128(slot + 4, (unsigned long long)__double_as_longlong(mn[1]), (tag << 32) | (unsigned int)tot[2]);
        st_relaxed_b128(slot + 6, (unsigned long long)__double_as_longlong(mx[1]), (tag << 32) | (unsigned int)tot[3]);
      }
    } else {  // DK_FLUSH: acknowledge
      if (tid == 0) {
        unsigned long long hi = ((unsigned long long)(seq & 0xffffffu) << 40) | (unsigned long long)kRankNone;
        st_relaxed_b128(slot, (unsigned long long)__double_as_longlong(-1.0), hi);
      }
    }
    seq++;
    __syncthreads();
  }
  // ---- DONE: write the tile back to the session tables ----
  for (int ln = tid; ln < tile.count; ln += blockDim.x) {
    int n = tile.base + ln;
    for (int r = 0; r < s.R; r++) {
      s.idle[(size_t)r * s.N + n] = tile.I[r * tile.npc + ln];
      s.rel[(size_t)r * s.N + n] = tile.L[r * tile.npc + ln];
    }
  }
}

// =============================================================================================
// sequencer CTA
// =============================================================================================
__device__ void sequencer_main(const ActionParams &p, unsigned char *smem, Ctl &ctl, Seq &seq) {
  const DevSnap &s = p.s;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    unsigned char *h = p.hot_in_smem ? smem : s.hot_global;
    auto take_from = [](unsigned char *&base, size_t bytes) {
      unsigned char *r = base;
      base += (bytes + 15) & ~(size_t)15;
      return r;
    };
    Replica &rp = seq.rp;
    rp.q_alloc = (double *)take_from(h, sizeof(double) * QR * s.Q);
    rp.q_alloc_np = (double *)take_from(h, sizeof(double) * QR * s.Q);
    rp.qkey = (QKey *)take_from(h, sizeof(QKey) * s.Q);
    rp.leaf_head = (int *)take_from(h, sizeof(int) * s.Q);
    rp.leaf_end = (int *)take_from(h, sizeof(int) * s.Q);
    rp.ovl_len = (int *)take_from(h, sizeof(int) * s.Q);
    rp.child_len = (int *)take_from(h, sizeof(int) * s.Q);
    rp.child_heap = (int *)take_from(h, sizeof(int) * s.Q);
    rp.root_heap = (int *)take_from(h, sizeof(int) * (s.n_top + 1));
    rp.qn_flags = (unsigned char *)take_from(h, s.Q);
    rp.touched = (unsigned int *)take_from(h, sizeof(unsigned int) * ((s.J + 31) / 32 + 1));
    // cold state = the session arrays themselves
    rp.t_status = s.t_status;
    rp.t_node = s.t_node;
    rp.t_node_status = s.t_node_status;
    rp.t_virtual = s.t_virtual;
    rp.ps_active_alloc = s.ps_cnt0;
    rp.ps_pending = s.ps_cnt0 + s.S;
    rp.ps_pipelined = s.ps_cnt0 + 2 * s.S;
    rp.j_req = s.j_req;
    rp.j_req_valid = s.j_req_valid;
    rp.j_key = s.j_key0;
    rp.leaf_heap = s.leaf_sorted;
    rp.ops = s.ops;
    rp.tta = s.tta;
    rp.ps_order = s.ps_order;
    seq.s = &p.s;
    seq.cfg = &p.cfg;
    seq.p = &p;
    seq.tile = nullptr;
    seq.ctl = &ctl;
    seq.n_ops = 0;
    seq.ops_cap = p.ops_cap;
    seq.root_len = 0;
    seq.batching = p.batching;
    seq.is_cta0 = true;
    seq.pods_placed = seq.pods_evicted = seq.sweeps = seq.nodes_scanned = seq.n_visits = 0;
    seq.minmax_exchanges = seq.batched = 0;
    seq.visits = p.visits;
    seq.visits_cap = p.visits_cap;
    seq.error = 0;
    seq.t_pop = seq.t_prep = seq.t_scan = seq.t_xchg = seq.t_apply = seq.t_finish = seq.t_init = 0;
    seq.t_key = seq.n_key = seq.t_tta = seq.t_heap = 0;
    ctl.trk[0].dirty = ctl.trk[1].dirty = 1;
    ctl.trk[0].mn = ctl.trk[1].mn = DBL_MAX;
    ctl.trk[0].mx = ctl.trk[1].mx = 0;
    ctl.trk[0].cnt_mn = ctl.trk[0].cnt_mx = ctl.trk[1].cnt_mn = ctl.trk[1].cnt_mx = 0;
    ctl.batch.valid = 0;
    for (int r = 0; r < KAI_MAX_RES; r++) ctl.dec.req[r] = 0;
    ctl.dec.pipeline_only = ctl.dec.res = ctl.dec.strategy = ctl.dec.gpu_task = ctl.dec.best_effort = 0;
    ctl.dec.nominated = ctl.dec.pred_class = -1;
    ctl.dec.task = -1;
    ctl.ctx_job = ctl.ctx_ps = -1;
    ctl.ctx_fresh = ctl.ctx_queue = ctl.ctx_preempt = ctl.ctx_base = 0;
    ctl.seq = p.seq0;
    ctl.n_delta = 0;
    ctl.stop = 0;
  }
  __syncthreads();
  long long tk0 = clock64();
  {
    Replica &rp = seq.rp;
    for (int i = tid; i < QR * s.Q; i += blockDim.x) {
      rp.q_alloc[i] = s.q_alloc[i];
      rp.q_alloc_np[i] = s.q_alloc_np[i];
    }
    for (int i = tid; i < s.Q; i += blockDim.x) {
      int b = s.q_job_begin[i];
      rp.leaf_head[i] = b;
      rp.leaf_end[i] = b + (s.q_nchildren[i] == 0 ? s.leaf_count[i] : 0);
      rp.ovl_len[i] = 0;
      rp.child_len[i] = 0;
      rp.qn_flags[i] = 0;
      rp.qkey[i].valid = 0;
    }
    for (int i = tid; i < (s.J + 31) / 32 + 1; i += blockDim.x) rp.touched[i] = 0;
  }
  __syncthreads();
  if (tid >= 32) return;  // the sequencer proper is warp 0
  if (lane == 0) {
    seq_init_job_order(seq);
    seq.t_init = clock64() - tk0;
  }
  __syncwarp();

  // ---- allocate action main loop (actions/allocate/allocate.go:46-111) ----
  for (;;) {
    if (lane == 0) {
      long long tk = clock64();
      int job = pop_next_job(seq);
      ctl.job = job;
      ctl.n_items = 0;
      ctl.job_ok = 0;
      if (job >= 0) {
        seq.n_ops = 0;
        long long tkt = clock64();
        seq.t_heap += tkt - tk;
        // job context: queue, preemptibility and (single-podset jobs) the podset counters
        const JobRec rec = s.jrec[job];
        ctl.ctx_job = job;
        ctl.ctx_queue = __ldg(&s.j_queue[job]);
        ctl.ctx_preempt = (__ldg(&s.j_flags[job]) & KAI_JOB_PREEMPTIBLE) ? 1 : 0;
        ctl.ctx_fresh = (!job_touched(seq, job) && rec.n_tta >= 0) ? 1 : 0;
        ctl.ctx_ps = -1;
        if (rec.n_podsets == 1) {
          if (!job_touched(seq, job)) {
            ctl.ctx_cnt[0] = rec.cnt[0];
            ctl.ctx_cnt[1] = rec.cnt[1];
            ctl.ctx_cnt[2] = rec.cnt[2];
          } else {
            for (int w = 0; w < 3; w++) ctl.ctx_cnt[w] = seq.rp.ps_active_alloc[(size_t)w * s.S + rec.ps0];
          }
          ctl.ctx_ps = rec.ps0;
        }
        // common/allocate.go:20-36 AllocateJob
        int n;
        double req[QR] = {0, 0, 0};
        if (ctl.ctx_fresh) {  // GetTasksToAllocate = tasks [tb, tb + n_tta), request sum precomputed
          n = rec.n_tta;
          ctl.ctx_base = rec.tb;
          for (int r = 0; r < QR; r++) req[r] = rec.req0[r];
          for (int k = 0; k < n; k++) prefetch_l1(s.t_req + (size_t)(rec.tb + k) * s.R);
        } else {
          n = tasks_to_allocate(seq, job, true, nullptr);
          ctl.ctx_base = -1;
          for (int k = 0; k < n; k++)
            for (int r = 0; r < QR; r++) req[r] = __dadd_rn(req[r], __ldg(&s.t_req[(size_t)seq.rp.tta[k] * s.R + r]));
        }
        seq.t_tta += clock64() - tkt;
        if (!over_capacity(seq, job, req)) {
          // tasks_to_allocate already emits tasks grouped in PodSetOrderFn order, which is the order
          // allocateSubGroupSetOnNodes/allocatePodSet visit them in (common/allocate.go:62-119)
          ctl.n_items = n;
          ctl.job_ok = 1;
        }
      }
      if (seq.error || ((volatile long long *)p.counters)[24] != 0) ctl.stop = 1;
      seq.t_pop += clock64() - tk;
    }
    __syncwarp();
    if (ctl.job < 0 || ctl.stop) break;
    bool job_success = ctl.job_ok != 0;
    if (job_success) {
      const int n_items = ctl.n_items;
      for (int k = 0; k < n_items; k++) {
        if (lane == 0) {
          long long tk = clock64();
          int t = ctl.ctx_base >= 0 ? ctl.ctx_base + k : seq.rp.tta[k];
          ctl.item_ok = seq_prepare_task(seq, t, ctl.job) ? 1 : 0;
          if (ctl.item_ok && ctl.use_batch) seq_apply_batched(seq, t);
          if (ctl.need_minmax) seq.minmax_exchanges++;
          seq.t_prep += clock64() - tk;
        }
        __syncwarp();
        if (!ctl.item_ok) {
          job_success = false;
          break;
        }
        if (ctl.use_batch) continue;  // placed without a sweep (same-node batching)
        long long tk1 = clock64();
        if (ctl.need_minmax) {
          seq_publish(p, ctl, DK_MINMAX);
          seq_gather_minmax(p, ctl);
        }
        seq_publish(p, ctl, DK_SCAN);
        seq_gather_candidates(p, ctl);
        if (lane == 0) {
          long long tk3 = clock64();
          seq_apply_winner(seq, ctl.dec.task);
          long long tk4 = clock64();
          seq.t_xchg += tk3 - tk1;
          seq.t_apply += tk4 - tk3;
        }
        __syncwarp();
        if (!ctl.item_ok) {
          job_success = false;
          break;
        }
      }
    }
    if (lane == 0) {
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
      if (ctl.ctx_ps >= 0)  // write the podset counters of the job back to the session state
        for (int w = 0; w < 3; w++) seq.rp.ps_active_alloc[(size_t)w * s.S + ctl.ctx_ps] = ctl.ctx_cnt[w];
      ctl.ctx_ps = -1;
      ctl.ctx_job = -1;
      ctl.ctx_fresh = 0;
      if (seq.error) ctl.stop = 1;
      seq.t_finish += clock64() - tk;
    }
    __syncwarp();
    if (ctl.stop) break;
  }
  // ---- DONE record (carries the last deltas), write back the hot per-queue state and the counters ----
  seq_publish(p, ctl, DK_DONE);
  for (int i = lane; i < QR * s.Q; i += 32) {
    s.q_alloc[i] = seq.rp.q_alloc[i];
    s.q_alloc_np[i] = seq.rp.q_alloc_np[i];
  }
  if (lane == 0) {
    p.counters[0] = seq.n_visits;
    p.counters[1] = seq.sweeps;
    p.counters[2] = seq.nodes_scanned;
    p.counters[3] = seq.pods_placed;
    p.counters[4] = seq.pods_evicted;
    p.counters[5] = seq.minmax_exchanges;
    p.counters[6] = seq.error;
    p.counters[7] = ctl.seq + 1;
    p.counters[8] = seq.t_init;
    p.counters[9] = seq.t_pop;
    p.counters[10] = seq.t_prep;
    p.counters[11] = seq.t_key;
    p.counters[12] = seq.t_xchg;
    p.counters[13] = seq.t_apply;
    p.counters[14] = seq.t_finish;
    p.counters[15] = seq.batched;
    p.counters[16] = seq.n_key;
    p.counters[17] = seq.t_tta;
    p.counters[18] = seq.t_heap;
  }
}

// =============================================================================================
// the kernel
// =============================================================================================
__global__ void __launch_bounds__(kThreads, 1) k_action(const __grid_constant__ ActionParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ Ctl ctl;
  __shared__ Tile tile;
  __shared__ Seq seq;
  __shared__ ScanShared scan_sh;
  __shared__ Cand sh_warp[kThreads / 32];
  __shared__ double sh_d[(kThreads / 32) * 4];
  __shared__ int sh_i[(kThreads / 32) * 4];
  if (blockIdx.x == 0)
    sequencer_main(p, smem, ctl, seq);
  else
    scanner_main(p, smem, sh_warp, sh_d, sh_i, scan_sh, tile);
}

}  // namespace kai
