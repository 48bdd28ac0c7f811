// kai_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see kai_oracle.h).
//
// A plain, sequential restatement of the reference's per-Session scheduling
// cycle (NVIDIA/KAI-Scheduler @72de64fc, pkg/scheduler).  Every function cites
// the reference file:line it follows (paths relative to pkg/scheduler/).  The
// reference is Go and cannot be compiled in this image (no Go toolchain, un-
// vendored k8s modules), so parity is pinned by the reference's own
// known-answer tests transcribed under tests/golden/ (tests/golden/README.md).
//
// Numeric contract: IEEE-754 binary64, no fused multiply-add (build with
// -ffp-contract=off), operation order exactly as in the Go sources.
//
// Where the reference iterates a Go map (unspecified order) the oracle uses
// ascending index order; SURVEY.md Appendix A.6 lists those places.
#include "kai_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <array>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int QR = KAI_QRES;
constexpr int kActiveUsed = KAI_POD_ALLOCATED | KAI_POD_PIPELINED | KAI_POD_BINDING | KAI_POD_BOUND |
                            KAI_POD_RUNNING | KAI_POD_RELEASING;  // pod_status.go:55
constexpr int kActiveAllocated = KAI_POD_ALLOCATED | KAI_POD_PIPELINED | KAI_POD_BINDING | KAI_POD_BOUND |
                                 KAI_POD_RUNNING;  // pod_status.go:56
constexpr int kAlive = kActiveAllocated | KAI_POD_PENDING | KAI_POD_GATED;                       // :57
constexpr int kAllocatedStatuses = KAI_POD_ALLOCATED | KAI_POD_BOUND | KAI_POD_BINDING | KAI_POD_RUNNING;  // :59

// ---------------------------------------------------------------------------
// container/heap (Go standard library, go1.24 src/container/heap/heap.go) —
// the algorithm behind scheduler_util.PriorityQueue (scheduler_util/priority_queue.go:50-118).
// Restated exactly because comparator inconsistencies make the pop order depend on it.
// ---------------------------------------------------------------------------
template <class T>
struct GoHeap {
  std::vector<T> items;
  std::function<bool(const T &, const T &)> less;
  int len() const { return (int)items.size(); }
  bool empty() const { return items.empty(); }
  void up(int j) {
    for (;;) {
      int i = (j - 1) / 2;  // parent
      if (i == j || !less(items[j], items[i])) break;
      std::swap(items[i], items[j]);
      j = i;
    }
  }
  bool down(int i0, int n) {
    int i = i0;
    for (;;) {
      int j1 = 2 * i + 1;
      if (j1 >= n || j1 < 0) break;
      int j = j1;
      int j2 = j1 + 1;
      if (j2 < n && less(items[j2], items[j1])) j = j2;
      if (!less(items[j], items[i])) break;
      std::swap(items[i], items[j]);
      i = j;
    }
    return i > i0;
  }
  void push(const T &x) {
    items.push_back(x);
    up(len() - 1);
  }
  T pop() {
    int n = len() - 1;
    std::swap(items[0], items[n]);
    down(0, n);
    T x = items.back();
    items.pop_back();
    return x;
  }
  void fix(int i) {
    if (!down(i, len())) up(i);
  }
  T remove(int i) {  // heap.Remove
    int n = len() - 1;
    if (n != i) {
      std::swap(items[i], items[n]);
      if (!down(i, n)) up(i);
    }
    T x = items.back();
    items.pop_back();
    return x;
  }
  const T &peek() const { return items[0]; }
};

// ---------------------------------------------------------------------------
// data model (api/**)
// ---------------------------------------------------------------------------
struct Share {  // plugins/proportion/resource_share/resource_share.go:12-21
  double deserved = 0, fair = 0, max_allowed = 0, oqw = 0, allocated = 0, alloc_np = 0, request = 0, usage = 0;
};

// resource_share.go:40-45
double requestable_share(const Share &s) {
  if (s.max_allowed == KAI_UNLIMITED) return s.request;
  return std::fmin(s.max_allowed, s.request);
}
// resource_share.go:51-61
double allocatable_share(const Share &s) {
  if (s.deserved == KAI_UNLIMITED) return s.max_allowed;
  double a = std::fmax(s.deserved, s.fair);
  if (s.max_allowed != KAI_UNLIMITED) a = std::fmin(s.max_allowed, a);
  return a;
}
// resource_quantities.go:81-97 compareQuantities
int compare_quantities(double q, double o) {
  if (q == KAI_UNLIMITED) return o == KAI_UNLIMITED ? 0 : 1;
  if (o == KAI_UNLIMITED) return -1;
  if (q > o) return 1;
  if (q < o) return -1;
  return 0;
}
// resource_quantities.go:57-64 LessEqual
bool q_less_equal(const double *a, const double *b) {
  for (int r = 0; r < QR; r++)
    if (compare_quantities(a[r], b[r]) > 0) return false;
  return true;
}

struct QueueAttr {  // resource_share/queue_resource_share.go:22-31
  int parent = -1, priority = 0, uid_rank = 0;
  int64_t creation = 0;
  Share s[QR];
  std::vector<int> children;
};

// queue_resource_share.go:142-166 GetDominantResourceShare
double dominant_share(const QueueAttr &q, const double *total) {
  double dom = 0.0;
  for (int r = 0; r < QR; r++) {
    double value;
    double allocatable = allocatable_share(q.s[r]);
    if (allocatable == KAI_UNLIMITED) allocatable = total[r];
    double allocated = q.s[r].allocated;
    if (allocatable == 0)
      value = allocated * 1000;  // noFairShareDrfMultiplier
    else
      value = allocated / allocatable;
    dom = std::fmax(dom, value);
  }
  return dom;
}

// ---------------------------------------------------------------------------
// plugins/proportion/resource_division/resource_division.go
// ---------------------------------------------------------------------------
// :283-290
bool is_queue_satisfied(const Share &s) {
  if (s.request <= s.fair) return true;
  if (s.max_allowed != KAI_UNLIMITED && s.max_allowed <= s.fair) return true;
  return false;
}
// :317-325
double remaining_requested(const Share &s) {
  double requested = requestable_share(s);
  if (requested < s.fair) return 0;
  return requested - s.fair;
}

struct RemReq {
  int q;
  double amount;
};

// :164-222 divideUpToFairShare — `group` in ascending index order (map order in Go)
double divide_up_to_fair_share(double total, double k, std::vector<QueueAttr> &Q, const std::vector<int> &group, int r,
                               std::map<int, double> &rr) {
  for (;;) {
    bool another = false;
    double round_amount = total;
    // :224-251 calcShareWeights, :307-315 getTotalWeightsForUnsatisfied
    double total_weights = 0;
    for (int q : group)
      if (remaining_requested(Q[q].s[r]) > 0) total_weights += Q[q].s[r].oqw;
    std::map<int, double> w;
    double wsum = 0.0;
    if (total_weights != 0) {
      for (int q : group) {
        const Share &s = Q[q].s[r];
        if (is_queue_satisfied(s)) continue;
        double n_weight = s.oqw / total_weights;
        double n_usage = s.usage;
        double t = n_weight - n_usage;
        double t2 = k * t;
        double sw = std::fmax(0.0, n_weight + t2);
        w[q] = sw;
        wsum += sw;
      }
    }
    if (wsum == 0) break;
    for (int q : group) {
      if (total == 0) break;
      Share &s = Q[q].s[r];
      if (is_queue_satisfied(s)) continue;
      double requested = remaining_requested(s);
      if (s.oqw == 0) continue;
      double qw = w.count(q) ? w[q] : 0.0;
      double nqw = qw / wsum;
      double fair = round_amount * nqw;
      // :264-281 getResourceToGiveInCurrentRound
      double give = 0;
      if (requested <= fair) {
        give = requested;
        rr.erase(q);
      } else {
        double rf = std::floor(fair);
        if (rf > 0) give = rf;
        if (fair - give > 0) rr[q] = fair - give;
      }
      if (give == 0) continue;
      s.fair += give;
      total -= give;
      another = another || requested < fair;
    }
    if (!another || total == 0) break;
  }
  return total;
}

double divide_over_quota_resource(double remaining, double k, std::vector<QueueAttr> &Q, const std::vector<int> &group, int r);

// :33-43 setResourceShare for one sibling group and one resource
double set_resource_share(double total, double k, std::vector<QueueAttr> &Q, const std::vector<int> &group, int r) {
  // :92-109 setDeservedResource
  double remaining = total;
  for (int q : group) {
    Share &s = Q[q].s[r];
    double deserved = s.deserved;
    if (deserved == KAI_UNLIMITED) deserved = total;
    double amount = std::fmin(deserved, requestable_share(s));
    s.fair += amount;
    remaining -= amount;
  }
  if (!(remaining > 0)) return 0;
  return divide_over_quota_resource(remaining, k, Q, group, r);
}

// :111-144 divideOverQuotaResource; :146-162 getQueuesByPriority (priorities descending)
double divide_over_quota_resource(double remaining, double k, std::vector<QueueAttr> &Q, const std::vector<int> &group, int r) {
  std::map<int, std::vector<int>, std::greater<int>> by_prio;
  for (int q : group) by_prio[Q[q].priority].push_back(q);
  std::map<int, std::map<int, double>, std::greater<int>> rem;
  for (auto &kv : by_prio) {
    std::map<int, double> rr;
    remaining = divide_up_to_fair_share(remaining, k, Q, kv.second, r, rr);
    rem[kv.first] = rr;
  }
  for (auto &kv : by_prio) {
    if (remaining <= 0) break;
    auto &rr = rem[kv.first];
    if (rr.empty()) continue;
    // :253-262 divideRemainingResource; order :335-357 (remaining desc, creation asc, uid asc)
    std::vector<RemReq> v;
    for (auto &e : rr) v.push_back({e.first, e.second});
    std::sort(v.begin(), v.end(), [&](const RemReq &a, const RemReq &b) {
      if (a.amount > b.amount) return true;
      if (a.amount < b.amount) return false;
      if (Q[a.q].creation != Q[b.q].creation) return Q[a.q].creation < Q[b.q].creation;
      return Q[a.q].uid_rank < Q[b.q].uid_rank;
    });
    size_t i = 0;
    while (!(remaining == 0) && i < v.size()) {
      double give = std::fmin(1.0, remaining);
      Q[v[i].q].s[r].fair += give;
      remaining -= give;
      i++;
    }
  }
  return remaining;
}

// plugins/topology/node_scoring.go:36-53: nodes of the i-th of n preferred-level domains (in sorted-tree order) score
// floor((i+1)/n * 10) * scores.Topology (plugins/scores/scores.go)
double topology_position_score(int i, int n) {
  double score = ((double)(i + 1) / (double)n) * 10;
  return std::floor(score) * 10000.0;
}

// ---------------------------------------------------------------------------
// plugins/proportion/queue_order/queue_order.go:19-73
// job_req: QuantifyResource(GetTasksToAllocateInitResource(job,..,false)) of the best pending job
// victims_alloc: Σ QuantifyResource(victim.Allocated) (reclaim victim queues), may be null
// ---------------------------------------------------------------------------
int queue_order_result(const QueueAttr &l, const QueueAttr &r, const double *l_req, const double *r_req,
                       const double *l_victims, const double *r_victims, const double *total) {
  // :87-100 prioritizeUnderUtilized — FairShare.Less(Allocated): strictly less on all resources
  auto over_utilized = [](const QueueAttr &q) {
    for (int i = 0; i < QR; i++)
      if (q.s[i].fair >= q.s[i].allocated) return false;
    return true;
  };
  bool lo = over_utilized(l), ro = over_utilized(r);
  if (!lo && ro) return -1;
  if (lo && !ro) return 1;
  // :102-128 prioritizeUnderQuotaWithJob
  double lw[QR], rw[QR], ld[QR], rd[QR];
  for (int i = 0; i < QR; i++) {
    lw[i] = l.s[i].allocated + l_req[i];
    rw[i] = r.s[i].allocated + r_req[i];
    ld[i] = l.s[i].deserved;
    rd[i] = r.s[i].deserved;
  }
  bool ls = q_less_equal(lw, ld), rs_ = q_less_equal(rw, rd);
  if (ls && !rs_) return -1;
  if (rs_ && !ls) return 1;
  // :75-85 prioritizePrioritized
  if (l.priority > r.priority) return -1;
  if (l.priority < r.priority) return 1;
  // :130-180 penalizeZeroShareWithJob
  auto violation = [](const QueueAttr &q, const double *with_job) {
    bool v = false;
    for (int i = 0; i < QR; i++) {
      if (allocatable_share(q.s[i]) != 0) continue;
      if (with_job[i] > 0) v = true;
    }
    return v;
  };
  bool lv = violation(l, lw), rv = violation(r, rw);
  if (lv && !rv) return 1;
  if (!lv && rv) return -1;
  // :182-201 prioritizeSmallerResourceShare; :242-273 calculateDominantResourceShareWithJob
  auto drf_with_job = [&](const QueueAttr &q, const double *req, const double *vict) {
    QueueAttr tmp = q;
    for (int i = 0; i < QR; i++) tmp.s[i].allocated += req[i];
    if (vict)
      for (int i = 0; i < QR; i++) tmp.s[i].allocated -= vict[i];
    return dominant_share(tmp, total);
  };
  double lsh = drf_with_job(l, l_req, l_victims), rsh = drf_with_job(r, r_req, r_victims);
  if (lsh < rsh) return -1;
  if (lsh > rsh) return 1;
  // :203-219
  lsh = dominant_share(l, total);
  rsh = dominant_share(r, total);
  if (lsh < rsh) return -1;
  if (lsh > rsh) return 1;
  // :221-233 prioritizeBasedOnAllocatableShare
  double la[QR], ra[QR];
  for (int i = 0; i < QR; i++) {
    la[i] = allocatable_share(l.s[i]);
    ra[i] = allocatable_share(r.s[i]);
  }
  if (!q_less_equal(ra, la) && q_less_equal(la, ra)) return -1;
  if (!q_less_equal(la, ra) && q_less_equal(ra, la)) return 1;
  // :235-240 prioritizeBasedOnCreationTime
  if (l.creation < r.creation) return -1;
  return 1;
}

// plugins/nodeplacement/pack.go:45-64
double binpack_score(double mn, double mx, double cur, double overall) {
  if (overall == 0) return 0.0;
  if (mx == 0) return 0.0;
  if (mn == mx) return 9.0;
  double t1 = cur - mn;
  double t2 = mx - mn;
  double t3 = t1 / t2;
  double t4 = 1 - t3;
  return 9 * t4;
}
// plugins/nodeplacement/spread.go:16-36
double spread_score(double non_allocated, double count) {
  if (count == 0) return 0;
  return non_allocated / count;
}

struct Task {
  int job = -1, podset = -1, status = 0, node = -1, order_rank = 0, nominated = -1, pred_class = -1;
  double req[KAI_MAX_RES] = {0};
  bool is_virtual = false;  // PodInfo.IsVirtualStatus
  // NodeInfo.PodInfos holds a CLONE of the task per node (node_info.go:400-402).  A task evicted from node A and
  // pipelined to node B in the same statement sits on both (Releasing on A, Pipelined on B); a victim that consolidation
  // moved and reclaim then evicts and re-places in a simulation sits on three, and so on: one entry per node, unbounded.
  std::vector<int> on_node, on_status;
  int find_on(int n) const {
    for (size_t e = 0; e < on_node.size(); e++)
      if (on_node[e] == n) return (int)e;
    return -1;
  }
};
struct PodSet {
  int job = -1, min_available = 0;
  std::vector<int> tasks;
};
// inner caches of a PodGroupInfo (job_info.go:98-101; invalidated on any status change :281-284)
struct TtaCache {
  bool tta_valid = false;
  std::vector<int> tta;
  bool tta_res_valid = false;
  double tta_res[QR] = {0, 0, 0};
};
struct Job {
  int queue = -1, priority = 0, order_rank = 0;
  bool preemptible = false;
  int signature = -1;  // GetSchedulingConstraintsSignature class (job_info.go:547-570); -1 = unique
  std::vector<int> podsets;
  TtaCache cache;
};
// PodGroupInfo.CloneWithTasks (job_info.go:477-510): a clone owns copies of the PodSets, so its per-podset counters,
// its PodStatusIndex, its Allocated sum and its inner caches are FROZEN at clone time (statement operations update
// the session's job, found by UID, never the clone), while the PodInfo objects it holds are the ones the statement
// mutates, so task statuses read through a clone are live.  Views with id < NJ are the session's jobs themselves.
struct View {
  int job = -1;
  std::vector<std::vector<int>> ps_tasks;  // members per podset (index = position in Job::podsets)
  std::vector<int> ps_min, ps_active_alloc, ps_active_used;
  int active_alloc_total = 0, n_pending = 0;
  double allocated[QR] = {0, 0, 0};
  TtaCache cache;
};

struct Op {  // framework/statement.go operations
  enum Kind { ALLOCATE, PIPELINE, EVICT, UNDO } kind;
  int task = -1;
  int prev_status = 0, prev_node = -1, next_node = -1;
  bool prev_virtual = false;
  int undo_index = -1;
};

struct QNode {  // actions/utils/job_order_by_queue.go:18-25 queueNode
  int queue = -1;
  bool is_leaf = false, needs_reorder = false;
  int parent = -1;  // index into nodes, -1 root level
  GoHeap<int> children;  // job ids (leaf) or QNode ids (non-leaf)
};

}  // namespace

struct kai_oracle {
  kai_config cfg;
  std::string err;
  int R = 4, N = 0, NQ = 0, NJ = 0, NS = 0, NT = 0, NPC = 0, mask_words = 0;
  std::vector<double> alloc, idle, rel;  // [R][N]
  std::vector<int> name_rank;
  std::vector<uint32_t> nflags;
  std::vector<double> gpu_count;
  std::vector<double> foreign;
  std::vector<QueueAttr> Q;
  std::vector<Job> J;
  std::vector<PodSet> PS;
  std::vector<Task> T;
  std::vector<uint32_t> pred_mask;
  double total[QR] = {0, 0, 0};
  bool loaded = false;
  bool use_signatures = false;  // SchedulerParams.UseSchedulingSignatures (options.go:120; tests: false)
  int n_threads = 1;

  // results
  std::vector<int32_t> r_task_node, r_task_status;
  std::vector<kai_job_visit> r_visits;
  std::vector<double> r_fair, r_alloc, r_alloc_np, r_request, r_idle, r_rel;
  double r_total[QR];
  int64_t pods_placed = 0, pods_evicted = 0;
  kai_stats stats;

  // statement
  std::vector<Op> ops;

  double &A(int r, int n) { return alloc[(size_t)r * N + n]; }
  double &I(int r, int n) { return idle[(size_t)r * N + n]; }
  double &L(int r, int n) { return rel[(size_t)r * N + n]; }

  // ---------------- PodInfo helpers (api/pod_info/pod_info.go) ----------------
  // :343-347 IsRequireAnyKindOfGPU (whole-GPU scope: GPUs() > 0)
  bool task_requires_gpu(const Task &t) const { return t.req[KAI_RES_GPU] > 0; }
  // resource_requirment.go:97-102 + base_resources.go:119-130 + gpu_resource_requirment.go:89-104
  bool task_req_is_empty(const Task &t) const {
    if (t.req[KAI_RES_GPU] > 0.01) return false;
    if (t.req[KAI_RES_CPU] >= 10 || t.req[KAI_RES_MEM] >= 10.0 * 1024 * 1024) return false;
    for (int r = 3; r < R; r++)
      if (t.req[r] >= 10) return false;
    return true;
  }
  // :518-521 ShouldAllocate
  bool should_allocate(const Task &t, bool real) const {
    return t.status == KAI_POD_PENDING || (!real && t.status == KAI_POD_RELEASING && t.is_virtual);
  }

  // ---------------- NodeInfo (api/node_info/node_info.go) ----------------
  // :361-382 isTaskAllocatableOnNonAllocatedResources -> :771-778 -> resource_requirment.go:126-140
  //  -> base_resources.go:90-105
  bool fits(const Task &t, int n, bool with_releasing) {
    for (int r = 0; r < R; r++) {
      double avail = I(r, n);
      if (with_releasing) avail = avail + L(r, n);
      if (r >= 3) {  // scalar resources: only requested ones are checked
        if (t.req[r] != 0 && t.req[r] > avail) return false;
      } else if (t.req[r] > avail)
        return false;
    }
    return true;
  }
  // :168-188 IsTaskAllocatable
  bool is_task_allocatable(const Task &t, int n) {
    if (task_req_is_empty(t)) return true;
    return fits(t, n, false);
  }
  // :190-206 IsTaskAllocatableOnReleasingOrIdle
  bool is_task_allocatable_releasing_or_idle(const Task &t, int n) { return fits(t, n, true); }
  // :697-702 IsCPUOnlyNode
  bool is_cpu_only_node(int n) { return !(nflags[n] & KAI_NODE_NOT_CPU_ONLY) && A(KAI_RES_GPU, n) <= 0; }

  // :457-493 addTaskResources (Used is not tracked: it never feeds a decision)
  void node_add_task(int ti) {
    Task &t = T[ti];
    int n = t.node;
    {
      int e = t.find_on(n);
      if (e < 0) {
        e = (int)t.on_node.size();
        t.on_node.push_back(n);
        t.on_status.push_back(t.status);
      }
      t.on_status[e] = t.status;
    }
    for (int r = 0; r < R; r++) {
      switch (t.status) {
        case KAI_POD_RELEASING:
          L(r, n) += t.req[r];
          I(r, n) -= t.req[r];
          break;
        case KAI_POD_PIPELINED:
          L(r, n) -= t.req[r];
          break;
        default:
          I(r, n) -= t.req[r];
      }
    }
  }
  // :515-551 removeTaskResources — uses the status of the clone stored on the node
  void node_remove_task(int ti, int n) {
    Task &t = T[ti];
    // NodeInfo.RemoveTask (:495-513) fails without touching anything when the pod is not on the node; callers that
    // continue with an add (UpdateTask in unevict, statement.go:171-183) then simply add it.  Happens when a task was
    // evicted, pipelined back onto its own node with updateTaskIfExistsOnNode and both are rolled back
    // (statement_checkpoint_test.go "rollback evict pipeline").
    if (t.find_on(n) < 0) {
      if (getenv("KAI_ORACLE_COUNT_MISSING")) fprintf(stderr, "[oracle] remove of a task that is not on the node\n");
      return;
    }
    for (int r = 0; r < R; r++) {
      switch (t.on_status[t.find_on(n)]) {
        case KAI_POD_RELEASING:
          L(r, n) -= t.req[r];
          I(r, n) += t.req[r];
          break;
        case KAI_POD_PIPELINED:
          L(r, n) += t.req[r];
          break;
        default:
          I(r, n) += t.req[r];
      }
    }
    {
      int e = t.find_on(n);
      t.on_node.erase(t.on_node.begin() + e);
      t.on_status.erase(t.on_status.begin() + e);
    }
  }

  // ---------------- PodGroupInfo ----------------
  void set_status(int ti, int status) {  // job_info.go:253-264 UpdateTaskStatus
    Task &t = T[ti];
    t.status = status;
    J[t.job].cache.tta_valid = false;
    J[t.job].cache.tta_res_valid = false;
  }
  // ---------------- views (see struct View) ----------------
  std::vector<View> views;
  bool is_clone(int v) const { return v >= NJ; }
  int vjob(int v) const { return v < NJ ? v : views[v - NJ].job; }
  TtaCache &vcache(int v) { return v < NJ ? J[v].cache : views[v - NJ].cache; }
  int v_nps(int v) const { return (int)J[vjob(v)].podsets.size(); }
  int v_min(int v, int k) const { return v < NJ ? PS[J[v].podsets[k]].min_available : views[v - NJ].ps_min[k]; }
  int v_active_alloc(int v, int k) const {
    return v < NJ ? podset_count(PS[J[v].podsets[k]], kActiveAllocated) : views[v - NJ].ps_active_alloc[k];
  }
  int v_active_used(int v, int k) const {
    return v < NJ ? podset_count(PS[J[v].podsets[k]], kActiveUsed) : views[v - NJ].ps_active_used[k];
  }
  const std::vector<int> &v_ps_tasks(int v, int k) const {
    return v < NJ ? PS[J[v].podsets[k]].tasks : views[v - NJ].ps_tasks[k];
  }
  std::vector<int> v_all_tasks(int v) const {  // GetAllPodsMap: canonical order = podset, then task index
    std::vector<int> out;
    for (int k = 0; k < v_nps(v); k++)
      for (int ti : v_ps_tasks(v, k)) out.push_back(ti);
    return out;
  }
  int v_active_alloc_total(int v) const {  // GetActiveAllocatedTasksCount (job_info.go:286-297)
    if (v >= NJ) return views[v - NJ].active_alloc_total;
    return job_count(J[v], kActiveAllocated);
  }
  int v_pending_count(int v) const { return v < NJ ? job_count(J[v], KAI_POD_PENDING) : views[v - NJ].n_pending; }
  void v_allocated(int v, double *out) const {  // PodGroupInfo.Allocated (job_info.go:245-250), cpu/mem/gpu
    if (v >= NJ) {
      for (int r = 0; r < QR; r++) out[r] += views[v - NJ].allocated[r];
      return;
    }
    for (int s2 : J[v].podsets)
      for (int ti : PS[s2].tasks)
        if (T[ti].status & kAllocatedStatuses)
          for (int r = 0; r < QR; r++) out[r] += T[ti].req[r];
  }
  // CloneWithTasks(tasks) of view `base` (the RootSubGroupSet clone keeps base's minAvailable values)
  int make_clone(int base, const std::vector<int> &tasks) {
    View c;
    c.job = vjob(base);
    const Job &j = J[c.job];
    int n = (int)j.podsets.size();
    c.ps_tasks.assign(n, {});
    c.ps_min.resize(n);
    c.ps_active_alloc.assign(n, 0);
    c.ps_active_used.assign(n, 0);
    for (int k = 0; k < n; k++) c.ps_min[k] = v_min(base, k);
    for (int ti : tasks) {
      int k = 0;
      while (j.podsets[k] != T[ti].podset) k++;
      c.ps_tasks[k].push_back(ti);
      if (T[ti].status & kActiveAllocated) {
        c.ps_active_alloc[k]++;
        c.active_alloc_total++;
      }
      if (T[ti].status & kActiveUsed) c.ps_active_used[k]++;
      if (T[ti].status == KAI_POD_PENDING) c.n_pending++;
      if (T[ti].status & kAllocatedStatuses)
        for (int r = 0; r < QR; r++) c.allocated[r] += T[ti].req[r];
    }
    for (auto &v : c.ps_tasks) std::sort(v.begin(), v.end());
    views.push_back(c);
    return NJ + (int)views.size() - 1;
  }
  int podset_count(const PodSet &ps, int mask) const {
    int c = 0;
    for (int ti : ps.tasks)
      if (T[ti].status & mask) c++;
    return c;
  }
  // subgroup_info/podset.go:114-120
  bool job_ready(const Job &j) const {
    for (int s : j.podsets) {
      const PodSet &ps = PS[s];
      int ready = podset_count(ps, kAlive) - podset_count(ps, KAI_POD_GATED);
      if (ready < ps.min_available) return false;
    }
    return true;
  }
  int job_count(const Job &j, int mask) const {
    int c = 0;
    for (int s : j.podsets) c += podset_count(PS[s], mask);
    return c;
  }

  // plugins/subgrouporder/subgroup_order.go:31-62 + framework/session_plugins.go:261-270 (name order = index order)
  bool podset_less_v(int v, int ka, int kb) const {
    int ln = v_active_alloc(v, ka), rn = v_active_alloc(v, kb);
    int lmin = v_min(v, ka), rmin = v_min(v, kb);
    bool lsat = ln >= lmin, rsat = rn >= rmin;
    if (!lsat && !rsat) return ka < kb;
    if (!lsat) return true;
    if (!rsat) return false;
    double lr = (double)ln / (double)lmin;
    double rr = (double)rn / (double)rmin;
    if (lr < rr) return true;
    if (rr < lr) return false;
    return ka < kb;
  }
  std::vector<int> ordered_podsets(int v) const {  // positions in Job::podsets, PodSetOrderFn order
    std::vector<int> sets(v_nps(v));
    for (int k = 0; k < (int)sets.size(); k++) sets[k] = k;
    std::sort(sets.begin(), sets.end(), [&](int a, int b) { return podset_less_v(v, a, b); });
    return sets;
  }

  // api/podgroup_info/allocation_info.go:27-54 GetTasksToAllocate
  const std::vector<int> &tasks_to_allocate(int v, bool real) {
    TtaCache &c = vcache(v);
    if (c.tta_valid) return c.tta;
    std::vector<int> out;
    std::vector<int> sets = ordered_podsets(v);
    // :165-177 getMaxNumSubGroupsToAllocate
    int unsat = 0;
    for (int k = 0; k < v_nps(v); k++)
      if (v_active_alloc(v, k) < v_min(v, k)) unsat++;
    int max_sets = unsat > 0 ? unsat : 1;
    int n_sets = 0;
    for (size_t i = 0; i < sets.size() && n_sets < max_sets; i++) {
      int k = sets[i];
      std::vector<int> cand;
      for (int ti : v_ps_tasks(v, k))
        if (should_allocate(T[ti], real)) cand.push_back(ti);
      if (cand.empty()) continue;
      std::sort(cand.begin(), cand.end(), [&](int a, int b) { return T[a].order_rank < T[b].order_rank; });
      // :144-153 getNumTasksToAllocate
      int n_alloc = v_active_alloc(v, k);
      int max_tasks;
      if (n_alloc >= v_min(v, k))
        max_tasks = std::min((int)cand.size(), 1);
      else
        max_tasks = v_min(v, k) - n_alloc;
      for (int k2 = 0; k2 < (int)cand.size() && k2 < max_tasks; k2++) out.push_back(cand[k2]);
      n_sets++;
    }
    TtaCache &c2 = vcache(v);
    c2.tta = out;
    c2.tta_valid = true;
    return c2.tta;
  }
  // allocation_info.go:87-113 GetTasksToAllocateInitResource, quantified (proportion/utils/utils.go:11-13)
  const double *tasks_to_allocate_init_resource(int v, bool real) {
    if (vcache(v).tta_res_valid) return vcache(v).tta_res;
    double acc[QR] = {0, 0, 0};
    std::vector<int> tta = tasks_to_allocate(v, real);
    for (int ti : tta)
      if (should_allocate(T[ti], real))
        for (int r = 0; r < QR; r++) acc[r] += T[ti].req[r];
    TtaCache &c = vcache(v);
    for (int r = 0; r < QR; r++) c.tta_res[r] = acc[r];
    c.tta_res_valid = true;
    return c.tta_res;
  }
  // api/podgroup_info/eviction_info.go:13-90 GetTasksToEvict: reverse podset order, reverse task order
  std::vector<int> tasks_to_evict(int v, bool &has_more) {
    std::vector<int> sets(v_nps(v));
    for (int k = 0; k < (int)sets.size(); k++) sets[k] = k;
    std::sort(sets.begin(), sets.end(), [&](int a, int b) { return podset_less_v(v, b, a); });
    int max_sets = (int)sets.size();  // :49-60 getNumOfSubGroupsToEvict
    for (int k = 0; k < v_nps(v); k++)
      if (v_active_alloc(v, k) > v_min(v, k)) {
        max_sets = 1;
        break;
      }
    std::vector<int> out;
    int n_sets = 0;
    for (size_t i = 0; i < sets.size() && n_sets < max_sets; i++) {
      int k = sets[i];
      std::vector<int> cand;
      for (int ti : v_ps_tasks(v, k))
        if (T[ti].status & kActiveAllocated) cand.push_back(ti);
      std::sort(cand.begin(), cand.end(), [&](int a, int b) { return T[a].order_rank > T[b].order_rank; });
      int n_alloc = v_active_alloc(v, k);  // :62-68 getMaxTasksToEvict
      int max_tasks = n_alloc > v_min(v, k) ? 1 : n_alloc;
      for (int k2 = 0; k2 < (int)cand.size() && k2 < max_tasks; k2++) out.push_back(cand[k2]);
      n_sets++;
    }
    has_more = (int)out.size() < v_active_alloc_total(v);
    return out;
  }
  bool has_tasks_to_allocate(int ji, bool real) const {  // allocation_info.go:18-25
    for (int s : J[ji].podsets)
      for (int ti : PS[s].tasks)
        if (should_allocate(T[ti], real)) return true;
    return false;
  }

  // ---------------- proportion: session open (proportion.go:242-423) ----------------
  void open_session() {
    // :252-288 setTotalResources
    for (int r = 0; r < QR; r++) total[r] = 0;
    for (int n = 0; n < N; n++) {
      if (!(nflags[n] & KAI_NODE_READY)) continue;
      for (int r = 0; r < QR; r++) {
        double v = A(r, n);
        if (!foreign.empty()) v -= foreign[(size_t)r * N + n];
        total[r] += v;
      }
    }
    // :347-401 updateQueuesCurrentResourceUsage
    for (auto &q : Q)
      for (int r = 0; r < QR; r++) q.s[r].allocated = q.s[r].alloc_np = q.s[r].request = q.s[r].fair = 0;
    for (int ji = 0; ji < NJ; ji++) {
      const Job &j = J[ji];
      for (int s : j.podsets)
        for (int ti : PS[s].tasks) {
          const Task &t = T[ti];
          if (t.status & kAllocatedStatuses) {
            for (int q = j.queue; q >= 0; q = Q[q].parent)
              for (int r = 0; r < QR; r++) {
                Q[q].s[r].allocated += t.req[r];
                Q[q].s[r].request += t.req[r];
                if (!j.preemptible) Q[q].s[r].alloc_np += t.req[r];
              }
          } else if (t.status == KAI_POD_PENDING) {
            for (int q = j.queue; q >= 0; q = Q[q].parent)
              for (int r = 0; r < QR; r++) Q[q].s[r].request += t.req[r];
          }
        }
    }
    // :403-423 setFairShare
    std::vector<int> top;
    for (int q = 0; q < NQ; q++)
      if (Q[q].parent < 0) top.push_back(q);
    set_fair_share_for_queues(total, top);
  }
  void set_fair_share_for_queues(const double *tot, const std::vector<int> &group) {
    if (group.empty()) return;
    for (int r = 0; r < QR; r++) set_resource_share(tot[r], cfg.k_value, Q, group, r);
    for (int q : group) {
      double fs[QR];
      for (int r = 0; r < QR; r++) fs[r] = Q[q].s[r].fair;
      set_fair_share_for_queues(fs, Q[q].children);
    }
  }

  // ---------------- capacity policy (plugins/proportion/capacity_policy) ----------------
  // max_allowed_check.go:16-66 + quota_check.go:25-77 through capacity_policy.go:63-74
  bool over_capacity(int ji, const double *req) {
    const Job &j = J[ji];
    for (int q = j.queue; q >= 0; q = Q[q].parent)
      for (int r = 0; r < QR; r++) {
        const Share &s = Q[q].s[r];
        if (s.max_allowed == KAI_UNLIMITED) continue;
        if (req[r] == 0) continue;
        if (s.max_allowed < s.allocated + req[r]) return true;
      }
    if (j.preemptible) return false;
    return non_preemptible_over_quota(ji, req);
  }
  bool non_preemptible_over_quota(int ji, const double *req) {
    const Job &j = J[ji];
    if (j.preemptible) return false;
    for (int q = j.queue; q >= 0; q = Q[q].parent)
      for (int r = 0; r < QR; r++) {
        const Share &s = Q[q].s[r];
        if (s.deserved == KAI_UNLIMITED) continue;
        if (req[r] == 0) continue;
        if (s.deserved < s.alloc_np + req[r]) return true;
      }
    return false;
  }

  // ---------------- event handlers (proportion.go:443-489) ----------------
  void queue_allocate(int ti, double sign) {
    const Task &t = T[ti];
    const Job &j = J[t.job];
    for (int q = j.queue; q >= 0; q = Q[q].parent)
      for (int r = 0; r < QR; r++) {
        if (sign > 0) {
          Q[q].s[r].allocated += t.req[r];
          if (!j.preemptible) Q[q].s[r].alloc_np += t.req[r];
        } else {
          Q[q].s[r].allocated -= t.req[r];
          if (!j.preemptible) Q[q].s[r].alloc_np -= t.req[r];
        }
      }
  }

  // ---------------- Statement (framework/statement.go) ----------------
  // :297-358 Allocate
  void stmt_allocate(int ti, int n) {
    Task &t = T[ti];
    Op op;
    op.kind = Op::ALLOCATE;
    op.task = ti;
    op.prev_status = t.status;
    op.prev_node = t.node;
    op.next_node = n;
    op.prev_virtual = t.is_virtual;
    set_status(ti, KAI_POD_ALLOCATED);
    t.node = n;
    node_add_task(ti);
    queue_allocate(ti, +1);
    ops.push_back(op);
    t.is_virtual = true;
  }
  // :392-427 unallocate
  void unallocate(int ti, bool prev_virtual) {
    Task &t = T[ti];
    set_status(ti, KAI_POD_PENDING);
    node_remove_task(ti, t.node);
    t.node = -1;
    t.is_virtual = prev_virtual;
    queue_allocate(ti, -1);
  }
  // :197-295 Pipeline
  void stmt_pipeline(int ti, int n, bool update_if_exists) {
    Task &t = T[ti];
    bool found_on_node = t.find_on(n) >= 0;
    if (found_on_node && !update_if_exists) {
      stmt_unevict(ti);
      return;
    }
    Op op;
    op.kind = Op::PIPELINE;
    op.task = ti;
    op.prev_status = t.status;
    op.prev_node = t.node;
    op.next_node = n;
    op.prev_virtual = t.is_virtual;
    set_status(ti, KAI_POD_PIPELINED);
    if (found_on_node) {
      node_remove_task(ti, n);  // node_info.go:570-576 UpdateTask
      t.node = n;
      node_add_task(ti);
    } else {
      t.node = n;
      node_add_task(ti);
    }
    queue_allocate(ti, +1);
    ops.push_back(op);
    t.is_virtual = true;
  }
  // :432-476 unpipeline
  void unpipeline(const Op &op) {
    Task &t = T[op.task];
    set_status(op.task, op.prev_status);
    int host = t.node;
    t.node = op.prev_node;
    t.is_virtual = op.prev_virtual;
    node_remove_task(op.task, host);
    queue_allocate(op.task, -1);
  }
  // :63-128 Evict
  void stmt_evict(int ti) {
    Task &t = T[ti];
    Op op;
    op.kind = Op::EVICT;
    op.task = ti;
    op.prev_status = t.status;
    op.prev_node = t.node;
    op.next_node = t.node;
    op.prev_virtual = t.is_virtual;
    set_status(ti, KAI_POD_RELEASING);
    node_remove_task(ti, t.node);  // UpdateTask
    node_add_task(ti);
    queue_allocate(ti, -1);
    ops.push_back(op);
    t.is_virtual = true;
  }
  // :156-195 unevict
  void unevict(const Op &op) {
    Task &t = T[op.task];
    set_status(op.task, op.prev_status);
    t.is_virtual = op.prev_virtual;
    {  // UpdateTask on the node the task was evicted from
      int keep = t.node;
      t.node = op.prev_node;
      node_remove_task(op.task, op.prev_node);
      node_add_task(op.task);
      t.node = keep;
    }
    queue_allocate(op.task, +1);
  }
  // :652-663 operationValid
  // decided by the FIRST undo operation that targets i (the reference scans from the start and returns at the first
  // match); the index is rebuilt lazily when the log changed since it was last used
  std::vector<int> first_undo;
  size_t first_undo_built = 0;
  void sync_first_undo() {
    if (first_undo_built > ops.size()) {
      first_undo.assign(ops.size(), -1);
      first_undo_built = 0;
    }
    first_undo.resize(ops.size(), -1);
    for (size_t u = first_undo_built; u < ops.size(); u++)
      if (ops[u].kind == Op::UNDO && first_undo[ops[u].undo_index] < 0) first_undo[ops[u].undo_index] = (int)u;
    first_undo_built = ops.size();
  }
  bool op_valid_rec(int i) const { return first_undo[i] < 0 ? true : !op_valid_rec(first_undo[i]); }
  bool op_valid(int i) {
    sync_first_undo();
    return op_valid_rec(i);
  }
  // :597-643 undoOperation
  void undo_operation(int index) {
    if (!op_valid(index)) return;
    Op op = ops[index];
    switch (op.kind) {
      case Op::EVICT:
        unevict(op);
        break;
      case Op::PIPELINE:
        unpipeline(op);
        break;
      case Op::ALLOCATE:
        unallocate(op.task, op.prev_virtual);
        break;
      case Op::UNDO:
        redo_operation(op.undo_index);
        break;
    }
    Op u;
    u.kind = Op::UNDO;
    u.undo_index = index;
    ops.push_back(u);
  }
  // the redoOperation closures of :607-628
  void redo_operation(int index) {
    Op op = ops[index];
    switch (op.kind) {
      case Op::EVICT:
        stmt_evict(op.task);
        break;
      case Op::PIPELINE:
        stmt_pipeline(op.task, op.next_node, true);
        break;
      case Op::ALLOCATE:
        stmt_allocate(op.task, op.next_node);
        break;
      case Op::UNDO:
        undo_operation(op.undo_index);
        break;
    }
  }
  // :478-481 Unevict -> :573-595 undoEarliestValidOperation
  void stmt_unevict(int ti) {
    for (int i = 0; i < (int)ops.size(); i++) {
      if (!op_valid(i)) continue;
      if (ops[i].kind != Op::EVICT || ops[i].task != ti) continue;
      undo_operation(i);
      return;
    }
  }
  int stmt_checkpoint() { return (int)ops.size(); }  // :44-46
  void stmt_rollback(int cp) {                       // :48-61
    for (int i = (int)ops.size() - 1; i >= cp; i--) undo_operation(i);
    ops.resize(cp);
    first_undo_built = (size_t)-1;
  }
  void stmt_discard() {  // :522-534
    for (int i = (int)ops.size() - 1; i >= 0; i--) undo_operation(i);
    ops.clear();
    first_undo_built = (size_t)-1;
  }
  // :483-520 ConvertAllAllocatedToPipelined
  void stmt_convert_all_allocated_to_pipelined(int ji) {
    size_t n0 = ops.size();
    for (size_t i = 0; i < n0; i++) {
      Op op = ops[i];
      if (op.kind != Op::ALLOCATE || T[op.task].job != ji) continue;
      int node = T[op.task].node;
      unallocate(op.task, true);
      stmt_pipeline(op.task, node, true);
    }
    std::vector<Op> keep;
    for (auto &op : ops)
      if (!(op.kind == Op::ALLOCATE && T[op.task].job == ji)) keep.push_back(op);
    ops = keep;
    first_undo_built = (size_t)-1;
  }
  // :536-571 Commit: allocate -> BindPod -> Binding (session.go:111-125); pipeline/evict keep their session status
  void stmt_commit() {
    for (int i = 0; i < (int)ops.size(); i++) {
      if (!op_valid(i)) continue;
      const Op &op = ops[i];
      if (op.kind == Op::ALLOCATE) {
        T[op.task].status = KAI_POD_BINDING;  // updatePodOnSession: node clone keeps accounting (default branch)
        if (T[op.task].find_on(T[op.task].node) >= 0) T[op.task].on_status[T[op.task].find_on(T[op.task].node)] = KAI_POD_BINDING;
        J[T[op.task].job].cache.tta_valid = J[T[op.task].job].cache.tta_res_valid = false;
        pods_placed++;
      } else if (op.kind == Op::PIPELINE) {
        pods_placed++;
      } else if (op.kind == Op::EVICT) {
        pods_evicted++;
      }
    }
    ops.clear();
    first_undo_built = (size_t)-1;
  }

  // ---------------- node ordering + fitting (framework/session.go:201-264,466-485) ----------------
  struct Best {
    double score;
    int rank;
    int node;
  };
  // returns best fitting node or -1.  node_set == nullptr means all nodes.
  int pick_node(int ti, const std::vector<int> *node_set) {
    const Task &t = T[ti];
    const int n_set = node_set ? (int)node_set->size() : N;
    stats.decisions++;
    stats.nodes_scanned += n_set;
    const bool gpu_task = task_requires_gpu(t);
    const int res = gpu_task ? KAI_RES_GPU : KAI_RES_CPU;  // nodeplacement.go:75-87
    const int strategy = gpu_task ? cfg.gpu_placement : cfg.cpu_placement;
    // pack.go:66-86 getMinMaxPerNode over the node set passed to allocateTask
    double mn = DBL_MAX, mx = 0;
    if (strategy == KAI_PLACEMENT_BINPACK && n_threads > 1 && n_set >= 4096) {
      pool_start(n_threads);
      std::vector<double> pmn(n_threads, DBL_MAX), pmx(n_threads, 0.0);
      std::function<void(int)> mm = [&](int w) {
        int k0 = (int)((int64_t)n_set * w / n_threads), k1 = (int)((int64_t)n_set * (w + 1) / n_threads);
        double a = DBL_MAX, b = 0;
        for (int k = k0; k < k1; k++) {
          int n = node_set ? (*node_set)[k] : k;
          if (A(res, n) == 0) continue;
          double cur = I(res, n) + L(res, n);
          if (cur < a) a = cur;
          if (cur > b) b = cur;
        }
        pmn[w] = a;
        pmx[w] = b;
      };
      pool_run(mm);
      for (int w = 0; w < n_threads; w++) {
        if (pmn[w] < mn) mn = pmn[w];
        if (pmx[w] > mx) mx = pmx[w];
      }
    } else if (strategy == KAI_PLACEMENT_BINPACK) {
      for (int k = 0; k < n_set; k++) {
        int n = node_set ? (*node_set)[k] : k;
        if (A(res, n) == 0) continue;
        double cur = I(res, n) + L(res, n);
        if (cur < mn) mn = cur;
        if (cur > mx) mx = cur;
      }
    }
    const uint32_t *mask = t.pred_class >= 0 ? &pred_mask[(size_t)t.pred_class * mask_words] : nullptr;
    auto sweep = [&](int k0, int k1) {
      Best b{-1.0, 0, -1};
      for (int k = k0; k < k1; k++) {
        int n = node_set ? (*node_set)[k] : k;
        // FittingNode (session.go:201-232): capacity part hoisted by the caller
        if (!is_task_allocatable_releasing_or_idle(t, n)) continue;
        if (mask && !((mask[n >> 5] >> (n & 31)) & 1u)) continue;
        // NodeOrderFn sum in plugin registration order (session_plugins.go:427-437)
        double score = 0.0;
        score += is_task_allocatable(t, n) ? 100.0 : 0.0;        // nodeavailability.go:29-40
        score += 0.0;                                            // gpusharingorder (whole GPUs only)
        score += (!gpu_task && is_cpu_only_node(n)) ? 10.0 : 0.0;  // resourcetype.go:29-41
        score += (t.nominated == n) ? 1000000.0 : 0.0;           // nominatednode.go:29-41
        double cur = I(res, n) + L(res, n);
        if (strategy == KAI_PLACEMENT_BINPACK)
          score += binpack_score(mn, mx, cur, A(res, n));
        else
          score += spread_score(cur, res == KAI_RES_GPU ? (double)(int64_t)gpu_count[n] : A(res, n));
        if (cur_scores) {  // topology/node_scoring.go:17-34: last NodeOrderFn; a node without an entry
          if ((*cur_scores)[n] < 0) continue;  // makes NodeOrderFn fail => the node is dropped (session.go:247-251)
          score += (*cur_scores)[n];
        }
        if (b.node < 0 || score > b.score || (score == b.score && name_rank[n] < b.rank)) {
          b.score = score;
          b.rank = name_rank[n];
          b.node = n;
        }
      }
      return b;
    };
    if (n_threads <= 1 || n_set < 4096) return sweep(0, n_set).node;
    // fan the sweep out over the worker pool (mirrors the reference's goroutine-per-node fan-out,
    // framework/session.go:243-262, with one slice per host thread)
    pool_start(n_threads);
    std::vector<Best> part(n_threads);
    std::function<void(int)> fn = [&](int w) {
      int k0 = (int)((int64_t)n_set * w / n_threads), k1 = (int)((int64_t)n_set * (w + 1) / n_threads);
      part[w] = sweep(k0, k1);
    };
    pool_run(fn);
    Best b{-1.0, 0, -1};
    for (auto &p : part) {
      if (p.node < 0) continue;
      if (b.node < 0 || p.score > b.score || (p.score == b.score && p.rank < b.rank)) b = p;
    }
    return b.node;
  }

  // ---- minimal spinning worker pool ----
  std::vector<std::thread> pool;
  std::atomic<uint64_t> pool_gen{0};
  std::atomic<int> pool_pending{0};
  std::atomic<bool> pool_stop{false};
  const std::function<void(int)> *pool_fn = nullptr;
  void pool_start(int n) {
    if ((int)pool.size() == n - 1) return;
    pool_shutdown();
    pool_stop = false;
    const uint64_t gen0 = pool_gen.load(std::memory_order_acquire);
    for (int w = 1; w < n; w++)
      pool.emplace_back([this, w, gen0] {
        uint64_t seen = gen0;
        for (;;) {
          uint64_t g;
          int idle = 0;
          while ((g = pool_gen.load(std::memory_order_acquire)) == seen) {
            if (pool_stop.load(std::memory_order_relaxed)) return;
            if (++idle > 2000) std::this_thread::yield();
          }
          seen = g;
          if (pool_stop.load(std::memory_order_acquire)) return;
          (*pool_fn)(w);
          pool_pending.fetch_sub(1, std::memory_order_acq_rel);
        }
      });
  }
  void pool_run(const std::function<void(int)> &fn) {
    pool_fn = &fn;
    pool_pending.store((int)pool.size(), std::memory_order_relaxed);
    pool_gen.fetch_add(1, std::memory_order_release);
    fn(0);
    while (pool_pending.load(std::memory_order_acquire) != 0) {
    }
  }
  void pool_shutdown() {
    pool_stop = true;
    pool_gen.fetch_add(1, std::memory_order_release);
    for (auto &t : pool) t.join();
    pool.clear();
  }
  ~kai_oracle() { pool_shutdown(); }


  // =====================================================================================================
  // plugins/topology: domain tree (topology_plugin.go:57-110), subSetNodesFn (job_filtering.go:34-112) and the
  // preferred-level node scores (node_scoring.go:36-53).  Scope: the constraint of the job's root SubGroupSet;
  // PodSets carry no constraint of their own.  Children of a domain that sortTree never reaches keep ascending
  // DomainID order (the reference: node map iteration order).
  // =====================================================================================================
  struct TopoDom {
    int level = -1;  // global level index, -1 = root
    int id = 0;
    std::vector<int> children, nodes;
    int alloc_pods = -1;               // allocatablePodsNotSet
    double free[KAI_MAX_RES] = {0};    // IdleOrReleasingResources
  };
  struct Topo {
    int lb = 0, le = 0;
    std::vector<TopoDom> doms;  // doms[0] = root
    std::vector<std::vector<int>> dom_at;  // [level - lb][id] -> index into doms
    std::vector<char> node_in;  // node carries every level label
  };
  std::vector<Topo> topos;
  std::vector<int> topo_level_begin, node_domain, job_topology, job_req_level, job_pref_level;
  // subGroupNodeScores of the job being allocated (topology_plugin.go:26-29): per SubGroup key (set id, or G + podset id)
  // a per-node table, -1 = no entry.  A task uses the table of its PodSet or of the nearest ancestor set that has one
  // (node_scoring.go:85-94 getRelevantNodeScores).
  std::map<int, std::vector<double>> sg_scores;
  const std::vector<double> *cur_scores = nullptr;
  // SubGroupSet tree, normalised at load: every job has a root set
  std::vector<int> set_parent, set_rank, job_root_set, ps_set;
  std::vector<std::vector<int>> set_children, set_podsets;
  std::vector<std::array<int, 3>> set_con, ps_con;  // (topology, required level, preferred level)
  std::vector<char> job_general;                    // nested sets or any topology constraint: walk the tree
  const std::vector<double> *scores_for_task(int ti) const {
    const int ps = T[ti].podset;
    auto it = sg_scores.find((int)set_parent.size() + ps);
    if (it != sg_scores.end()) return &it->second;
    for (int g = ps_set[ps]; g >= 0; g = set_parent[g]) {
      it = sg_scores.find(g);
      if (it != sg_scores.end()) return &it->second;
    }
    return nullptr;
  }
  int ND(int level, int n) const { return node_domain[(size_t)level * N + n]; }
  void build_topologies() {
    topos.clear();
    int nt = (int)topo_level_begin.size() - 1;
    for (int k = 0; k < nt; k++) {
      Topo tp;
      tp.lb = topo_level_begin[k];
      tp.le = topo_level_begin[k + 1];
      tp.doms.emplace_back();
      tp.node_in.assign(N, 0);
      tp.dom_at.assign(tp.le - tp.lb, {});
      for (int l = tp.lb; l < tp.le; l++) {
        int mx = -1;
        for (int n = 0; n < N; n++) mx = std::max(mx, ND(l, n));
        tp.dom_at[l - tp.lb].assign(mx + 1, -1);
      }
      for (int n = 0; n < N; n++) {
        bool in = tp.le > tp.lb;
        for (int l = tp.lb; l < tp.le; l++)
          if (ND(l, n) < 0) in = false;
        if (!in) continue;
        tp.node_in[n] = 1;
        tp.doms[0].nodes.push_back(n);
        int parent = 0;
        for (int l = tp.lb; l < tp.le; l++) {
          int &di = tp.dom_at[l - tp.lb][ND(l, n)];
          if (di < 0) {
            di = (int)tp.doms.size();
            tp.doms.emplace_back();
            tp.doms[di].level = l;
            tp.doms[di].id = ND(l, n);
            tp.doms[parent].children.push_back(di);
          }
          tp.doms[di].nodes.push_back(n);
          parent = di;
        }
      }
      for (auto &d : tp.doms)  // canonical child order where the reference has map order: ascending DomainID
        std::sort(d.children.begin(), d.children.end(), [&](int a, int b) { return tp.doms[a].id < tp.doms[b].id; });
      topos.push_back(tp);
    }
  }
  // job_filtering.go:445-486 getJobRatioToFreeResources on ResourceList quantities (Quantity.Value() of a milli
  // quantity rounds up: cpu in whole cores; memory in bytes; scalar resources other than pods like cpu)
  double job_ratio_to_free(const double *tasks_res, const TopoDom &d) const {
    double ratio = 0.0;
    bool empty = true;
    for (int r = 0; r < R; r++)
      if (tasks_res[r] > 0) empty = false;
    if (empty) return 0.0;
    if (tasks_res[KAI_RES_GPU] > 0) ratio = std::max(ratio, tasks_res[KAI_RES_GPU] / d.free[KAI_RES_GPU]);
    for (int r = 0; r < R; r++) {
      if (r == KAI_RES_GPU || r == 3) continue;  // gpus handled above; "pods" ignored for bin-packing
      int64_t tq = r == KAI_RES_MEM ? (int64_t)tasks_res[r] : (int64_t)std::ceil((double)(int64_t)tasks_res[r] / 1000.0);
      if (tq == 0) continue;
      int64_t fq = r == KAI_RES_MEM ? (int64_t)d.free[r] : (int64_t)std::ceil((double)(int64_t)d.free[r] / 1000.0);
      double rr = fq == 0 ? 1000.0 : (double)tq / (double)fq;
      ratio = std::max(ratio, rr);
    }
    return ratio;
  }
  bool check_job_domain_fit(const double *tasks_res, int tasks_count, const TopoDom &d) const {  // :302-320
    if (d.alloc_pods != -1) return d.alloc_pods >= tasks_count;
    return !(job_ratio_to_free(tasks_res, d) > 1.0);
  }
  void calc_subtree_free(Topo &tp, int di) {  // :191-211
    TopoDom &d = tp.doms[di];
    if (d.children.empty()) {
      for (int n : d.nodes)
        for (int r = 0; r < R; r++) {
          d.free[r] += I(r, n);
          d.free[r] += L(r, n);
        }
      return;
    }
    for (int c : d.children) {
      calc_subtree_free(tp, c);
      for (int r = 0; r < R; r++) tp.doms[di].free[r] += tp.doms[c].free[r];
    }
  }
  // :213-247 calcNodeAccommodation with the shared, growing list of test pods (k-th = (k-1)-th + max pod)
  int calc_subtree_allocatable(Topo &tp, int di, const double *max_pod, std::vector<std::vector<double>> &test_pods, int n_tasks) {
    TopoDom &d = tp.doms[di];
    d.alloc_pods = 0;
    if (d.children.empty()) {
      for (int n : d.nodes) {
        bool only_pods = true;  // maxPodResources.LessEqual(onePodOnly)
        for (int r = 0; r < R; r++)
          if (r == 3 ? max_pod[r] > 1 : max_pod[r] > 0) only_pods = false;
        if (only_pods) {
          d.alloc_pods += n_tasks;
          continue;
        }
        auto fits_pod = [&](const std::vector<double> &rq) {
          for (int r = 0; r < R; r++) {
            double avail = I(r, n) + L(r, n);
            if (r >= 3) {
              if (rq[r] != 0 && rq[r] > avail) return false;
            } else if (rq[r] > avail)
              return false;
          }
          return true;
        };
        int cnt = 0;
        for (auto &tpod : test_pods) {
          if (fits_pod(tpod))
            cnt++;
          else
            break;
        }
        if (cnt == (int)test_pods.size()) {
          for (;;) {
            std::vector<double> next = test_pods.back();
            for (int r = 0; r < R; r++) next[r] += max_pod[r];
            test_pods.push_back(next);
            if (fits_pod(next))
              cnt++;
            else
              break;
          }
        }
        d.alloc_pods += cnt;
      }
      return d.alloc_pods;
    }
    for (int c : d.children) {
      int a = calc_subtree_allocatable(tp, c, max_pod, test_pods, n_tasks);
      tp.doms[di].alloc_pods += a;
    }
    return tp.doms[di].alloc_pods;
  }
  void sort_tree(Topo &tp, int di, const double *tasks_res, int max_depth_level) {  // :396-420
    if (max_depth_level == -2) return;
    TopoDom &d = tp.doms[di];
    std::vector<std::pair<double, int>> keyed;
    for (int c : d.children) keyed.push_back({job_ratio_to_free(tasks_res, tp.doms[c]), c});
    std::stable_sort(keyed.begin(), keyed.end(), [&](const std::pair<double, int> &a, const std::pair<double, int> &b) {
      if (a.first != b.first) return a.first > b.first;
      return tp.doms[a.second].id < tp.doms[b.second].id;
    });
    for (size_t i = 0; i < keyed.size(); i++) tp.doms[di].children[i] = keyed[i].second;
    if (tp.doms[di].level == max_depth_level) return;
    std::vector<int> ch = tp.doms[di].children;
    for (int c : ch) sort_tree(tp, c, tasks_res, max_depth_level);
  }
  void level_domains(const Topo &tp, int di, int level, std::vector<int> &out) const {  // node_scoring.go:55-68
    if (tp.doms[di].level == level) {
      out.push_back(di);
      return;
    }
    for (int c : tp.doms[di].children) level_domains(tp, c, level, out);
  }
  // subSetNodesFn for the root SubGroupSet of the view's job.  ok = false: error / no node set (job fails).
  // `con` = the SubGroup's (topology, required, preferred); `key` identifies it in sg_scores; `under` = the PodSets below
  // it (positions in Job::podsets) for the active-pods restriction.
  std::vector<std::vector<int>> subset_nodes(int v, const std::array<int, 3> &con, int key, const std::vector<int> &under,
                                             const std::vector<int> &tasks, const std::vector<int> &node_set, bool &ok) {
    ok = true;
    const int k = con[0];
    if (k == -2) {  // requested topology does not exist
      return {};
    }
    if (k < 0 || tasks.empty()) return {node_set};
    Topo &tp = topos[k];
    const int req = con[1], pref = con[2];  // level index inside the topology, -1 none, -2 unknown
    // common.go:17-61 lowestCommonDomainID
    std::vector<int> valid;
    for (int n : node_set)
      if (tp.node_in[n]) valid.push_back(n);
    int dom = 0;
    for (int l = tp.lb; l < tp.le; l++) {
      bool all = !valid.empty();
      int value = valid.empty() ? -1 : ND(l, valid[0]);
      for (int n : valid)
        if (ND(l, n) != value) all = false;
      if (!all) break;
      dom = tp.dom_at[l - tp.lb][value];
      if (pref >= 0 && l - tp.lb == pref) break;
    }
    for (auto &d : tp.doms) {  // treeAllocatableCleanup
      d.alloc_pods = -1;
      for (int r = 0; r < KAI_MAX_RES; r++) d.free[r] = 0;
    }
    calc_subtree_free(tp, dom);
    // useRepresentorPodsAccounting (:503-527): all tasks use GPUs or none does (pods is the only scalar resource)
    int gpu_pods = 0;
    for (int ti : tasks)
      if (T[ti].req[KAI_RES_GPU] > 0) gpu_pods++;
    bool scalars_uniform = true;
    for (int r = 3; r < R; r++) {
      int c = 0;
      for (int ti : tasks)
        if (T[ti].req[r] != 0) c++;
      if (c != 0 && c != (int)tasks.size()) scalars_uniform = false;
    }
    if ((gpu_pods == (int)tasks.size() || gpu_pods == 0) && scalars_uniform) {
      std::vector<double> max_pod(R, 0.0);
      for (int ti : tasks)
        for (int r = 0; r < R; r++) max_pod[r] = std::max(max_pod[r], T[ti].req[r]);
      std::vector<std::vector<double>> test_pods{max_pod};
      calc_subtree_allocatable(tp, dom, max_pod.data(), test_pods, (int)tasks.size());
    }
    double tasks_res[KAI_MAX_RES] = {0};
    for (int ti : tasks)
      for (int r = 0; r < R; r++) tasks_res[r] += T[ti].req[r];
    const int tasks_count = (int)tasks.size();
    if (!check_job_domain_fit(tasks_res, tasks_count, tp.doms[dom])) return {};
    if (req == -2 || pref == -2 || (req < 0 && pref < 0)) {  // calculateRelevantDomainLevels errors
      ok = false;
      return {};
    }
    const int max_depth = pref >= 0 ? tp.lb + pref : tp.lb + req;
    sort_tree(tp, dom, tasks_res, max_depth);
    if (pref >= 0) {  // node_scoring.go:36-53 calculateNodeScores
      std::vector<double> &topo_scores = sg_scores[key];
      topo_scores.assign(N, -1.0);
      std::vector<int> lvl;
      level_domains(tp, dom, tp.lb + pref, lvl);
      for (size_t i = 0; i < lvl.size(); i++) {
        double normalized = topology_position_score((int)i, (int)lvl.size());
        for (int n : tp.doms[lvl[i]].nodes) topo_scores[n] = normalized;
      }
    }
    // :249-300 getJobAllocatableDomains; levels from the lowest up: start at preferred / required, stop after required
    std::vector<int> relevant;  // global level index, -1 = root
    {
      bool found_pref = false, found_req = false;
      for (int l = tp.le - 1; l >= tp.lb - 1; l--) {
        int li = l >= tp.lb ? l - tp.lb : -1;
        if (l >= tp.lb && pref >= 0 && li == pref) found_pref = true;
        if (l >= tp.lb && req >= 0 && li == req) found_req = true;
        if (found_pref || found_req) relevant.push_back(l >= tp.lb ? l : -1);
        if (found_req) break;
      }
    }
    std::vector<char> allowed(tp.doms.size(), 1);
    {
      bool has_active = false;
      for (int ps2 : under)
        if (v_active_alloc(v, ps2) > 0) has_active = true;
      if (has_active && req >= 0) {  // :269-300 getRelevantDomainsWithAllocatedPods
        std::fill(allowed.begin(), allowed.end(), 0);
        std::function<void(int)> mark = [&](int di) {
          allowed[di] = 1;
          for (int c : tp.doms[di].children) mark(c);
        };
        for (int di : tp.dom_at[req]) {
          if (di < 0) continue;
          bool has = false;
          for (int ps2 : under)
            for (int ti : v_ps_tasks(v, ps2))
              if ((T[ti].status & kActiveAllocated) && T[ti].node >= 0 && tp.node_in[T[ti].node] &&
                  tp.dom_at[req][ND(tp.lb + req, T[ti].node)] == di)
                has = true;
          if (has) mark(di);
        }
      }
    }
    std::vector<char> chosen(tp.doms.size(), 0);
    bool any = false;
    for (int l : relevant)
      for (size_t di = 0; di < tp.doms.size(); di++) {
        if (tp.doms[di].level != l || !allowed[di]) continue;
        if (!check_job_domain_fit(tasks_res, tasks_count, tp.doms[di])) continue;
        chosen[di] = 1;
        any = true;
      }
    if (!any) return {};
    // sortDomainInfos: reverse level order of the (sorted) tree from the root
    std::vector<std::vector<int>> levels;
    std::vector<int> cur{0};
    while (!cur.empty()) {
      levels.push_back(cur);
      std::vector<int> next;
      for (int di : cur)
        for (int c : tp.doms[di].children) next.push_back(c);
      cur = next;
    }
    std::vector<char> in_valid(N, 0);
    for (int n : valid) in_valid[n] = 1;
    std::vector<std::vector<int>> out;
    for (int li = (int)levels.size() - 1; li >= 0; li--)
      for (int di : levels[li]) {
        if (!chosen[di]) continue;
        std::vector<int> ns;
        for (int n : tp.doms[di].nodes)
          if (in_valid[n]) ns.push_back(n);
        out.push_back(ns);
      }
    return out;
  }

  // ---------------- actions/common/allocate.go ----------------
  // :121-174 allocateTask + allocateTaskToNode
  bool allocate_task(int ti, const std::vector<int> *node_set, bool pipeline_only) {
    const Task &t = T[ti];
    // predicates.go:196-200 -> capacity_policy.go:51-61 with node_info.go:734-744:
    // whole-GPU pods are checked with GPU = 1, CPU-only pods with GPU = 0; node independent.
    double req[QR] = {t.req[KAI_RES_CPU], t.req[KAI_RES_MEM], task_requires_gpu(t) ? 1.0 : 0.0};
    if (over_capacity(t.job, req)) return false;
    cur_scores = sg_scores.empty() ? nullptr : scores_for_task(ti);
    int n = pick_node(ti, node_set);
    cur_scores = nullptr;
    if (getenv("KAI_ORACLE_TRACE")) fprintf(stderr, "[solver]       task %d -> node %d (pipeline_only %d)\n", ti, n, (int)pipeline_only);
    if (n < 0) return false;
    if (!pipeline_only && is_task_allocatable(t, n))
      stmt_allocate(ti, n);
    else
      stmt_pipeline(ti, n, !pipeline_only);
    return true;
  }
  // :20-119 AllocateJob (flat root SubGroupSet; topology subsetting = single node set)
  // `v` is a view: the session's job or a clone (partial job representative of the solver)
  bool allocate_job(int v, const std::vector<int> *node_set, bool pipeline_only) {
    std::vector<int> tta = tasks_to_allocate(v, !pipeline_only);
    const int ji = vjob(v);
    // capacity_policy.go:26-36 IsJobOverQueueCapacity
    double req[QR] = {0, 0, 0};
    for (int ti : tta)
      for (int r = 0; r < QR; r++) req[r] += T[ti].req[r];
    if (over_capacity(ji, req)) return false;
    sg_scores.clear();  // PreJobAllocation (topology_plugin.go:52-55)
    auto on_nodes = [&](const std::vector<int> *ns) {  // flat job: the root's PodSets in PodSetOrderFn order
      int cp = stmt_checkpoint();
      std::vector<int> sets = ordered_podsets(v);
      for (int k : sets) {
        const int s = J[ji].podsets[k];
        int cp2 = stmt_checkpoint();
        bool ok = true;
        for (int ti : tta) {
          if (T[ti].podset != s) continue;
          if (!allocate_task(ti, ns, pipeline_only)) {
            ok = false;
            break;
          }
        }
        if (!ok) {
          stmt_rollback(cp2);
          stmt_rollback(cp);
          return false;
        }
      }
      return true;
    };
    if (job_general.empty() || !job_general[ji]) return on_nodes(node_set);
    std::vector<int> all;
    if (node_set)
      all = *node_set;
    else
      for (int n = 0; n < N; n++) all.push_back(n);
    return alloc_set(v, job_root_set[ji], tta, all, pipeline_only);
  }
  // allocate.go:36-83 allocateSubGroupSet / allocateSubGroupSetOnNodes / allocatePodSet on the SubGroupSet tree
  void podsets_under(int g, int ji, std::vector<int> &out) const {
    for (int ps : set_podsets[g])
      for (size_t k = 0; k < J[ji].podsets.size(); k++)
        if (J[ji].podsets[k] == ps) out.push_back((int)k);
    for (int c : set_children[g]) podsets_under(c, ji, out);
  }
  bool alloc_set(int v, int g, const std::vector<int> &tasks, const std::vector<int> &nodes, bool pipeline_only) {
    const int ji = vjob(v);
    std::vector<int> under;
    podsets_under(g, ji, under);
    bool ok = true;
    std::vector<std::vector<int>> node_sets = subset_nodes(v, set_con[g], g, under, tasks, nodes, ok);
    if (!ok) return false;
    for (auto &ns : node_sets) {
      int cp = stmt_checkpoint();
      if (set_on_nodes(v, g, tasks, ns, pipeline_only)) return true;
      stmt_rollback(cp);
    }
    return false;
  }
  bool set_on_nodes(int v, int g, const std::vector<int> &tasks, const std::vector<int> &ns, bool pipeline_only) {
    const int ji = vjob(v);
    for (int c : set_children[g]) {  // orderedSubGroupSets: by name
      std::vector<int> under;
      podsets_under(c, ji, under);
      std::vector<int> sub;
      for (int ti : tasks)
        for (int k : under)
          if (T[ti].podset == J[ji].podsets[k]) sub.push_back(ti);
      if (!alloc_set(v, c, sub, ns, pipeline_only)) return false;
    }
    std::vector<int> ks;  // orderedPodSets: the set's own PodSets in PodSetOrderFn order
    for (int ps : set_podsets[g])
      for (size_t k = 0; k < J[ji].podsets.size(); k++)
        if (J[ji].podsets[k] == ps) ks.push_back((int)k);
    std::sort(ks.begin(), ks.end(), [&](int a, int b) { return podset_less_v(v, a, b); });
    for (int k : ks) {
      const int ps = J[ji].podsets[k];
      std::vector<int> pt;
      for (int ti : tasks)
        if (T[ti].podset == ps) pt.push_back(ti);
      bool ok = true;
      std::vector<std::vector<int>> node_sets = subset_nodes(v, ps_con[ps], (int)set_parent.size() + ps, {k}, pt, ns, ok);
      if (!ok) return false;
      bool placed = false;
      for (auto &ns2 : node_sets) {
        int cp = stmt_checkpoint();
        bool all_ok = true;
        for (int ti : pt)
          if (!allocate_task(ti, &ns2, pipeline_only)) {
            all_ok = false;
            break;
          }
        if (all_ok) {
          placed = true;
          break;
        }
        stmt_rollback(cp);
      }
      if (!placed) return false;
    }
    return true;
  }
  // job_info.go:443-464 ShouldPipelineJob
  bool should_pipeline_job(int ji) const {
    for (int s : J[ji].podsets) {
      const PodSet &ps = PS[s];
      bool has_pipelined = false;
      int active = 0;
      for (int ti : ps.tasks) {
        if (T[ti].status == KAI_POD_PIPELINED)
          has_pipelined = true;
        else if (T[ti].status & kActiveAllocated)
          active++;
      }
      if (has_pipelined && active < ps.min_available) return true;
    }
    return false;
  }

  // ---------------- job ordering (actions/utils/job_order_by_queue.go) ----------------
  struct JobsOrder {
    kai_oracle *o;
    bool victim_queue = false;
    std::vector<QNode> nodes;
    std::vector<int> queue_node;  // queue -> QNode index or -1
    GoHeap<int> root;
    std::vector<std::vector<int>> popped_by_queue;

    // plugins/elastic/elastic.go:50-63 minAvailableState
    void min_available_state(int v, bool &below, bool &above, bool &exactly) const {
      exactly = true;
      for (int k = 0; k < o->v_nps(v); k++) {
        int n = o->v_active_alloc(v, k);
        if (n < o->v_min(v, k)) {
          below = true;
          above = false;
          exactly = false;
          return;
        }
        if (n > o->v_min(v, k)) exactly = false;
      }
      below = false;
      above = !exactly;
    }
    // framework/session_plugins.go:227-242 JobOrderFn = priority, elastic, creation, UID
    bool job_less(int l, int r) const {
      const Job &lj = o->J[o->vjob(l)], &rj = o->J[o->vjob(r)];
      if (lj.priority > rj.priority) return true;  // plugins/priority/priority.go:41-54
      if (lj.priority < rj.priority) return false;
      bool lb, la, le, rb, ra, re;
      min_available_state(l, lb, la, le);
      min_available_state(r, rb, ra, re);
      if (lb && !rb) return true;  // plugins/elastic/elastic.go:25-48
      if (le && ra) return true;
      if (!lb && rb) return false;
      if (la && re) return false;
      return lj.order_rank < rj.order_rank;
    }
    // :280-346 best job of a subtree and the queue comparator
    int best_job(int ni) const {
      const QNode &n = nodes[ni];
      if (n.is_leaf) return n.children.peek();
      return best_job(n.children.peek());
    }
    int leaf_of_best(int ni) const {
      const QNode &n = nodes[ni];
      if (n.is_leaf) return ni;
      return leaf_of_best(n.children.peek());
    }
    bool node_less(int l, int r) {
      if (nodes[l].children.empty()) return !victim_queue;
      if (nodes[r].children.empty()) return victim_queue;
      double lreq[QR] = {0, 0, 0}, rreq[QR] = {0, 0, 0}, lv[QR] = {0, 0, 0}, rv[QR] = {0, 0, 0};
      if (!victim_queue) {
        const double *a = o->tasks_to_allocate_init_resource(best_job(l), false);
        for (int i = 0; i < QR; i++) lreq[i] = a[i];
        const double *b = o->tasks_to_allocate_init_resource(best_job(r), false);
        for (int i = 0; i < QR; i++) rreq[i] = b[i];
      } else {
        victims_allocated(l, lv);
        victims_allocated(r, rv);
      }
      int res = queue_order_result(o->Q[nodes[l].queue], o->Q[nodes[r].queue], lreq, rreq,
                                   victim_queue ? lv : nullptr, victim_queue ? rv : nullptr, o->total);
      bool result = res < 0;
      return victim_queue ? !result : result;
    }
    // :338-346 getVictimsForQueue: popped jobs of the leaf + its next job; Σ job.Allocated (cpu, mem, gpu)
    void victims_allocated(int ni, double *out) {
      int leaf = leaf_of_best(ni);
      int q = nodes[leaf].queue;
      std::vector<int> v = popped_by_queue[q];
      if (!nodes[leaf].children.empty()) v.push_back(nodes[leaf].children.peek());
      for (int vi : v) o->v_allocated(vi, out);  // job_info.go:245-250 PodGroupInfo.Allocated
    }
    void init(kai_oracle *oracle, bool victims) {
      o = oracle;
      victim_queue = victims;
      nodes.clear();
      nodes.reserve(o->NQ);
      queue_node.assign(o->NQ, -1);
      popped_by_queue.assign(o->NQ, {});
      root = GoHeap<int>();
      root.less = [this](const int &a, const int &b) { return node_less(a, b); };
    }
    int make_node(int q, bool leaf) {
      nodes.emplace_back();
      int id = (int)nodes.size() - 1;
      QNode &n = nodes[id];
      n.queue = q;
      n.is_leaf = leaf;
      if (leaf)
        n.children.less = [this](const int &a, const int &b) { return victim_queue ? !job_less(a, b) : job_less(a, b); };
      else
        n.children.less = [this](const int &a, const int &b) { return node_less(a, b); };
      return id;
    }
    void mark_ancestors(int ni) {  // :246-250
      for (int c = ni; c >= 0; c = nodes[c].parent) nodes[c].needs_reorder = true;
    }
    // :135-175 ensureAncestorChainForPush.  `linked` tracks whether the node already sits in a parent heap
    std::vector<char> linked;
    void ensure_chain(int child) {
      int cq = nodes[child].queue;
      if ((int)linked.size() < (int)nodes.size()) linked.resize(nodes.size(), 0);
      if (o->Q[cq].parent < 0) {
        if (!linked[child]) {
          root.push(child);
          linked[child] = 1;
        }
        return;
      }
      int pq = o->Q[cq].parent;
      int pn = queue_node[pq];
      bool is_new = pn < 0;
      if (is_new) {
        pn = make_node(pq, false);
        queue_node[pq] = pn;
        if ((int)linked.size() < (int)nodes.size()) linked.resize(nodes.size(), 0);
      }
      if (!linked[child]) {
        nodes[child].parent = pn;
        nodes[pn].children.push(child);
        linked[child] = 1;
      }
      if (is_new) ensure_chain(pn);
    }
    void push_job(int ji) {  // :90-119 (ji is a view id)
      int q = o->J[o->vjob(ji)].queue;
      if (!o->Q[q].children.empty()) return;
      int leaf = queue_node[q];
      bool needs_linking = leaf < 0;
      if (needs_linking) {
        leaf = make_node(q, true);
        queue_node[q] = leaf;
      }
      nodes[leaf].children.push(ji);
      if (needs_linking) ensure_chain(leaf);
      mark_ancestors(leaf);
    }
    bool is_empty() const { return root.empty(); }
    int get_next_node(GoHeap<int> &pq) {  // :193-215
      for (;;) {
        if (pq.empty()) return -1;
        int ni = pq.peek();
        if (nodes[ni].needs_reorder) {
          pq.fix(0);
          nodes[ni].needs_reorder = false;
          continue;
        }
        if (nodes[ni].children.empty()) return -1;
        return ni;
      }
    }
    int traverse_to_leaf() {  // :178-190
      GoHeap<int> *pq = &root;
      for (;;) {
        int ni = get_next_node(*pq);
        if (ni < 0) return -1;
        if (nodes[ni].is_leaf) return ni;
        pq = &nodes[ni].children;
      }
    }
    void handle_pop(int ni) {  // :219-243
      if (nodes[ni].children.len() == 0) {
        if (nodes[ni].parent >= 0)
          nodes[nodes[ni].parent].children.pop();
        else
          root.pop();
        queue_node[nodes[ni].queue] = -1;
        linked[ni] = 0;
        if (nodes[ni].parent >= 0) handle_pop(nodes[ni].parent);
        return;
      }
      mark_ancestors(ni);
    }
    int pop_next_job() {  // :61-88
      if (is_empty()) return -1;
      int leaf = traverse_to_leaf();
      if (leaf < 0) return -1;
      int job = nodes[leaf].children.pop();
      if (victim_queue) popped_by_queue[nodes[leaf].queue].push_back(job);
      handle_pop(leaf);
      return job;
    }
  };

  // actions/utils/input_jobs.go:21-68 InitializeWithJobs.  The reference ranges over a Go map
  // (unspecified order); the canonical order used here and by the engine is: leaf queues in
  // ascending index, the jobs of a queue in JobOrderFn order.
  void init_jobs_order(JobsOrder &jo, bool filter_non_pending, bool filter_unready) {
    std::vector<std::vector<int>> by_queue(NQ);
    for (int ji = 0; ji < NJ; ji++) {
      const Job &j = J[ji];
      if (filter_unready && !job_ready(j)) continue;
      if (filter_non_pending && job_count(j, KAI_POD_PENDING) == 0) continue;
      if (j.queue < 0) continue;
      if (!Q[j.queue].children.empty()) continue;
      by_queue[j.queue].push_back(ji);
    }
    for (int q = 0; q < NQ; q++) {
      std::sort(by_queue[q].begin(), by_queue[q].end(), [&](int a, int b) { return jo.job_less(a, b); });
      for (int ji : by_queue[q]) jo.push_job(ji);
    }
  }


  // =====================================================================================================
  // Victim-selection solver shared by reclaim and consolidation
  // (actions/common/solvers/{job_solver,pod_scenario_builder,by_pod_solver}.go, scenario/*.go,
  //  accumulated_scenario_filters/idle_gpus/*.go, actions/common/action.go).
  // Go map iteration orders are resolved canonically: ascending node / job / queue index.
  // =====================================================================================================
  enum { SOLVER_RECLAIM = 0, SOLVER_CONSOLIDATION = 1, SOLVER_PREEMPT = 2 };

  // input_jobs.go:21-68 with an explicit job set (views) and the option flags
  struct OrderOpts {
    bool filter_unready = false, filter_non_pending = false, filter_non_preemptible = false,
         filter_non_active_allocated = false;
  };
  void init_jobs_order_views(JobsOrder &jo, const std::vector<int> &vs, const OrderOpts &op) {
    std::vector<std::vector<int>> by_queue(NQ);
    for (int v : vs) {
      const Job &j = J[vjob(v)];
      if (op.filter_unready && !job_ready(j)) continue;
      if (op.filter_non_pending && v_pending_count(v) == 0) continue;
      if (op.filter_non_preemptible && !j.preemptible) continue;
      if (op.filter_non_active_allocated) {
        bool active = false;
        for (int ti : v_all_tasks(v))
          if (T[ti].status & kActiveAllocated) active = true;
        if (!active) continue;
      }
      if (j.queue < 0) continue;
      if (!Q[j.queue].children.empty()) continue;
      by_queue[j.queue].push_back(v);
    }
    for (int q = 0; q < NQ; q++) {
      std::sort(by_queue[q].begin(), by_queue[q].end(),
                [&](int a, int b) { return jo.victim_queue ? jo.job_less(b, a) : jo.job_less(a, b); });
      for (int v : by_queue[q]) jo.push_job(v);
    }
  }

  struct Scenario {  // scenario/base_scenario.go + by_node_scenario.go
    int preemptor = -1;  // partial job representative (view)
    std::vector<int> pending_tasks, potential_tasks, recorded_jobs, recorded_tasks;
    std::map<int, std::vector<int>> victims;       // job -> victim tasks in append order (api.VictimInfo.Tasks)
    std::map<int, std::vector<int>> task_groups;   // job -> task-group representatives (views)
    std::map<int, std::vector<int>> jobs_by_node;  // node -> jobs with potential victims on it
  };
  void scenario_append_group(Scenario &sc, const std::vector<int> &tasks) {  // base_scenario.go:112-128
    int ji = T[tasks[0]].job;
    int group = make_clone(ji, tasks);
    sc.task_groups[ji].push_back(group);
    auto &vt = sc.victims[ji];
    vt.insert(vt.end(), tasks.begin(), tasks.end());
  }
  void scenario_add_potential(Scenario &sc, const std::vector<int> &tasks) {  // by_node_scenario.go:52-57
    if (tasks.empty()) return;
    sc.potential_tasks.insert(sc.potential_tasks.end(), tasks.begin(), tasks.end());
    scenario_append_group(sc, tasks);
    for (int ti : tasks) {
      auto &v = sc.jobs_by_node[T[ti].node];
      if (std::find(v.begin(), v.end(), T[ti].job) == v.end()) v.push_back(T[ti].job);
    }
  }
  void scenario_init(Scenario &sc, int partial, const std::vector<int> &recorded_jobs) {  // base_scenario.go:34-67
    sc = Scenario();
    sc.preemptor = partial;
    sc.pending_tasks = v_all_tasks(partial);
    sc.recorded_jobs = recorded_jobs;
    for (int rv : recorded_jobs) scenario_append_group(sc, v_all_tasks(rv));
    for (int rv : recorded_jobs)
      for (int ti : v_all_tasks(rv)) sc.recorded_tasks.push_back(ti);
  }
  std::vector<int> scenario_victims_from_node(const Scenario &sc, int node) {  // by_node_scenario.go:59-80
    std::vector<int> out;
    auto it = sc.jobs_by_node.find(node);
    if (it == sc.jobs_by_node.end()) return out;
    std::vector<int> jobs = it->second;
    std::sort(jobs.begin(), jobs.end());
    for (int ji : jobs) {
      auto g = sc.task_groups.find(ji);
      if (g == sc.task_groups.end()) continue;
      for (int group : g->second)
        for (int ti : v_all_tasks(group)) out.push_back(ti);
    }
    return out;
  }

  // accumulated_scenario_filters/idle_gpus/idle_gpus.go: top-k nodes by idle+releasing GPUs (+ GPUs freed by the
  // accumulated victims), greedy first-fit of the pending tasks' GPU requests sorted descending
  struct IdleGpusFilter {
    std::vector<double> idle;  // per node
    std::multiset<double, std::greater<double>> sorted;  // the same values, descending (the reference keeps its
                                                          // top-k list incrementally: orderedInsert, idle_gpus.go:210-247)
    int k = 0;
    size_t n_rec_done = 0, n_pot_done = 0;
    std::vector<char> seen;  // per task: already accounted as victim
    bool active = false;
  };
  void idle_filter_init(IdleGpusFilter &f, const Scenario &sc) {
    f.idle.assign(N, 0.0);
    for (int n = 0; n < N; n++) f.idle[n] = I(KAI_RES_GPU, n) + L(KAI_RES_GPU, n);
    f.sorted.clear();
    for (int n = 0; n < N; n++) f.sorted.insert(f.idle[n]);
    f.k = (int)sc.pending_tasks.size();
    f.seen.assign(NT, 0);
    f.active = true;
    idle_filter_update(f, sc);
  }
  void idle_filter_update(IdleGpusFilter &f, const Scenario &sc) {
    // recorded victims never change within a builder and potential victims are append-only: only new entries
    for (const std::vector<int> *lst : {&sc.recorded_tasks, &sc.potential_tasks})
      for (size_t i = (lst == &sc.recorded_tasks ? f.n_rec_done : f.n_pot_done); i < lst->size(); i++) {
        int ti = (*lst)[i];
        if (T[ti].node < 0 || f.seen[ti]) continue;
        f.seen[ti] = 1;
        if (N > 0) {
          f.sorted.erase(f.sorted.find(f.idle[T[ti].node]));
          f.idle[T[ti].node] += T[ti].req[KAI_RES_GPU];
          f.sorted.insert(f.idle[T[ti].node]);
        }
      }
    f.n_rec_done = sc.recorded_tasks.size();
    f.n_pot_done = sc.potential_tasks.size();
  }
  bool idle_filter_check(IdleGpusFilter &f, const Scenario &sc) {
    idle_filter_update(f, sc);
    std::vector<double> req;
    for (int ti : sc.pending_tasks) req.push_back(T[ti].req[KAI_RES_GPU]);
    std::sort(req.begin(), req.end(), std::greater<double>());
    std::vector<double> cap;
    for (auto it = f.sorted.begin(); it != f.sorted.end() && (int)cap.size() < f.k; ++it) cap.push_back(*it);
    return greedy_match_requirements(req, cap);
  }
  // idle_gpus/common.go:34-64 greedyMatchRequirements: requirements and holder capacities both sorted descending
  static bool greedy_match_requirements(const std::vector<double> &req, const std::vector<double> &cap) {
    std::vector<double> used(cap.size(), 0.0);
    for (double required : req) {
      if (required == 0) return true;
      bool matched = false;
      for (size_t h = 0; h < cap.size(); h++) {
        if (cap[h] < required) break;
        if (cap[h] - used[h] >= required) {
          used[h] += required;
          matched = true;
          break;
        }
      }
      if (!matched) return false;
    }
    return true;
  }

  struct SolveState {
    std::vector<int> recorded_jobs, recorded_tasks;
  };
  struct SolveResult {
    bool has = false, solved = false;
    std::vector<int> victim_tasks, victim_jobs;
  };
  int solver_kind = SOLVER_RECLAIM;
  int solver_reclaimer_job = -1;
  std::vector<QueueAttr> sim_queues;  // proportion.go:131-136 jobSimulationQueues

  // plugins/proportion/reclaimable/reclaimable.go:29-51
  static bool can_reclaim_from_share(const QueueAttr &q, const double *req, bool preemptible) {
    for (int r = 0; r < QR; r++)
      if (compare_quantities(q.s[r].allocated + req[r], q.s[r].fair) > 0) return false;
    if (preemptible) return true;
    for (int r = 0; r < QR; r++)
      if (compare_quantities(q.s[r].alloc_np + req[r], q.s[r].deserved) > 0) return false;
    return true;
  }
  bool can_reclaim_resources(int ji) {
    const Job &j = J[ji];
    return can_reclaim_from_share(Q[j.queue], tasks_to_allocate_init_resource(ji, false), j.preemptible);
  }
  // reclaimable.go:234-263 getLeveledQueues
  void leveled_queues(const std::vector<QueueAttr> &QS, int reclaimer_q, int reclaimee_q, int &a, int &b) {
    std::vector<int> pa, pb;
    for (int q = reclaimer_q; q >= 0; q = QS[q].parent) pa.insert(pa.begin(), q);
    for (int q = reclaimee_q; q >= 0; q = QS[q].parent) pb.insert(pb.begin(), q);
    size_t n = std::min(pa.size(), pb.size());
    a = b = -1;
    for (size_t i = 0; i < n; i++) {
      a = pa[i];
      b = pb[i];
      if (a != b) break;
    }
  }
  struct Quant {
    double v[QR];
  };
  static bool quant_le(const double *a, const double *b) {
    for (int r = 0; r < QR; r++)
      if (compare_quantities(a[r], b[r]) > 0) return false;
    return true;
  }
  // strategies/strategies.go:18-91
  // strategies/strategies.go:42-57 MaintainFairShareStrategy.Reclaimable
  static bool maintain_fair_share(const QueueAttr &reclaimee, const double *remaining) {
    double allocatable[QR];
    for (int r = 0; r < QR; r++) allocatable[r] = allocatable_share(reclaimee.s[r]);
    return !quant_le(remaining, allocatable);
  }
  // strategies/strategies.go:59-91 GuaranteeDeservedQuotaStrategy.Reclaimable + reclaimerWillGoOverQuota
  static bool guarantee_deserved_quota(const QueueAttr &reclaimer, const QueueAttr &reclaimee, const double *reclaimer_req,
                                       const double *remaining) {
    double want[QR], rdes[QR], deserved[QR];
    for (int r = 0; r < QR; r++) {
      want[r] = reclaimer.s[r].allocated + reclaimer_req[r];
      rdes[r] = reclaimer.s[r].deserved;
      deserved[r] = reclaimee.s[r].deserved;
    }
    if (!quant_le(want, rdes)) return false;
    if (quant_le(remaining, deserved)) return false;
    return true;
  }
  // strategies/strategies.go:15-30 FitsReclaimStrategy: any strategy accepts
  bool fits_reclaim_strategy(const std::vector<QueueAttr> &QS, const double *reclaimer_req, int reclaimer_q,
                             int reclaimee_q, const double *remaining) {
    return maintain_fair_share(QS[reclaimee_q], remaining) ||
           guarantee_deserved_quota(QS[reclaimer_q], QS[reclaimee_q], reclaimer_req, remaining);
  }
  static double saturation_ratio(double allocated, double fair) {  // reclaimable.go:222-232
    if (fair == 0) return allocated > 0 ? INFINITY : 0.0;
    if (fair == KAI_UNLIMITED) return 0.0;
    return allocated / fair;
  }
  // reclaimable.go:53-220 Reclaimable on the queue snapshot taken at OnJobSolutionStart
  bool reclaimable(const std::vector<QueueAttr> &QS, int reclaimer_q, bool reclaimer_preemptible,
                   const double *reclaimer_req, const std::map<int, std::vector<Quant>> &by_queue) {
    std::map<int, Quant> remaining;
    std::map<int, unsigned> involved;  // bit r set = resource r involved
    auto get_remaining = [&](int q) -> Quant & {
      auto it = remaining.find(q);
      if (it == remaining.end()) {
        Quant x;
        for (int r = 0; r < QR; r++) x.v[r] = QS[q].s[r].allocated;
        it = remaining.emplace(q, x).first;
      }
      return it->second;
    };
    auto involved_of = [&](const std::vector<Quant> &rs) {
      unsigned m = 0;
      for (const Quant &x : rs)
        for (int r = 0; r < QR; r++)
          if (x.v[r] > 0) m |= 1u << r;
      return m;
    };
    for (const auto &kv : by_queue) {
      int reclaimee_leaf = kv.first;
      int lq, eq;
      leveled_queues(QS, reclaimer_q, reclaimee_leaf, lq, eq);
      involved[reclaimee_leaf] = involved_of(kv.second);
      get_remaining(eq);
      for (const Quant &res : kv.second) {
        if (!fits_reclaim_strategy(QS, reclaimer_req, lq, eq, get_remaining(eq).v)) return false;
        for (int q = reclaimee_leaf; q >= 0; q = QS[q].parent) {  // subtractReclaimedResources
          Quant &rem = get_remaining(q);
          for (int r = 0; r < QR; r++) rem.v[r] -= res.v[r];
          if (involved.count(q))
            involved[q] |= involved[reclaimee_leaf];
          else
            involved[q] = involved[reclaimee_leaf];
        }
      }
    }
    // reclaimingQueuesRemainWithinBoundaries
    unsigned reclaimer_involved = 0;
    for (int r = 0; r < QR; r++)
      if (reclaimer_req[r] > 0) reclaimer_involved |= 1u << r;
    for (int rq = reclaimer_q; rq >= 0; rq = QS[rq].parent) {
      Quant mine;
      auto it = remaining.find(rq);
      if (it != remaining.end()) {
        for (int r = 0; r < QR; r++) it->second.v[r] += reclaimer_req[r];  // the map entry itself is updated
        mine = it->second;
      } else {
        for (int r = 0; r < QR; r++) mine.v[r] = QS[rq].s[r].allocated + reclaimer_req[r];
      }
      std::vector<int> sib_ids;
      for (const auto &kv : remaining) sib_ids.push_back(kv.first);
      for (int sib : sib_ids) {
        if (QS[sib].parent != QS[rq].parent || sib == rq) continue;
        const Quant &sr = remaining[sib];
        unsigned inv = (involved.count(sib) ? involved[sib] : 0u) | reclaimer_involved;
        for (int r = 0; r < QR; r++) {
          if (!(inv & (1u << r))) continue;
          double rf = QS[rq].s[r].fair, sf = QS[sib].s[r].fair;
          if (rf == KAI_UNLIMITED && sf == KAI_UNLIMITED) continue;
          double ratio_r = saturation_ratio(mine.v[r], rf), ratio_s = saturation_ratio(sr.v[r], sf);
          if (ratio_r > 1 && sf > 0 && ratio_r * cfg.saturation_multiplier >= ratio_s) return false;
        }
      }
      if (reclaimer_preemptible) continue;
      for (int r = 0; r < QR; r++)
        if (compare_quantities(QS[rq].s[r].alloc_np + reclaimer_req[r], QS[rq].s[r].deserved) > 0) return false;
    }
    return true;
  }
  // ---------------- plugins/minruntime ----------------
  std::vector<double> q_preempt_mrt, q_reclaim_mrt, j_last_start;  // seconds; < 0 = nil; last start <= 0 = nil
  std::vector<double> j_stale_since;                               // StalenessInfo.TimeStamp, <= 0 = nil
  double now_s = 0;
  // resolver.go:46-67 resolvePreemptMinRuntime: first value set on the queue or an ancestor, else the default
  double preempt_min_runtime(int q) const {
    for (int c = q; c >= 0; c = Q[c].parent)
      if (!q_preempt_mrt.empty() && q_preempt_mrt[c] >= 0) return q_preempt_mrt[c];
    return cfg.default_preempt_min_runtime_s;
  }
  double reclaim_min_runtime(int pq, int vq) const {  // resolver.go:69-187
    if (pq < 0 || vq < 0) return cfg.default_reclaim_min_runtime_s;
    auto set = [&](int q) { return !q_reclaim_mrt.empty() && q_reclaim_mrt[q] >= 0; };
    if (cfg.reclaim_resolve_method == KAI_RESOLVE_QUEUE) {  // :88-106
      for (int c = vq; c >= 0; c = Q[c].parent)
        if (set(c)) return q_reclaim_mrt[c];
      return cfg.default_reclaim_min_runtime_s;
    }
    std::vector<int> pp, vp;  // getQueueHierarchyPath: root first (:189-203)
    for (int c = pq; c >= 0; c = Q[c].parent) pp.insert(pp.begin(), c);
    for (int c = vq; c >= 0; c = Q[c].parent) vp.insert(vp.begin(), c);
    if (pp[0] != vp[0]) return set(vp[0]) ? q_reclaim_mrt[vp[0]] : cfg.default_reclaim_min_runtime_s;  // :122-131
    size_t lca = 0, n = std::min(pp.size(), vp.size());
    for (size_t i = 0; i < n; i++) {
      if (pp[i] != vp[i]) break;
      lca = i;
    }
    if (lca + 1 < vp.size()) lca++;  // the victim-side child of the common ancestor (:143-145)
    for (size_t i = lca + 1; i-- > 0;)
      if (set(vp[i])) return q_reclaim_mrt[vp[i]];
    return cfg.default_reclaim_min_runtime_s;
  }
  bool job_elastic(int ji) const {  // job_info.go:408-415, podset.go:127-129
    for (int ps : J[ji].podsets)
      if (PS[ps].min_available < (int)PS[ps].tasks.size()) return true;
    return false;
  }
  // minruntime.go:147-192 isReclaimMinRuntimeProtected / isPreemptMinRuntimeProtected
  bool minruntime_protected(bool reclaim, int pending_job, int victim) const {
    if (j_last_start.empty() || !(j_last_start[victim] > 0)) return false;
    double mrt = reclaim ? reclaim_min_runtime(J[pending_job].queue, J[victim].queue) : preempt_min_runtime(J[victim].queue);
    return now_s < j_last_start[victim] + mrt;
  }
  // minruntime.go:93-105 reclaimFilterFn / preemptFilterFn
  bool minruntime_filter(bool reclaim, int pending_job, int victim) const {
    if (job_elastic(victim)) return true;
    return !minruntime_protected(reclaim, pending_job, victim);
  }
  // minruntime.go:107-145 scenario validators + :206-229 validVictimForMinAvailable
  bool minruntime_validator(const Scenario &sc, bool reclaim) const {
    int pj = vjob(sc.preemptor);
    for (const auto &kv : sc.victims) {
      int vj = kv.first;
      if (!job_elastic(vj) || !minruntime_protected(reclaim, pj, vj)) continue;
      for (int ps : J[vj].podsets) {
        int victims = 0, running = 0;
        for (int t : kv.second)
          if (T[t].podset == ps) victims++;
        if (!victims) continue;
        for (int t : PS[ps].tasks)
          if (T[t].status & kActiveUsed) running++;
        if (PS[ps].min_available > running - victims) return false;
      }
    }
    return true;
  }

  // proportion.go:143-240 reclaimableFn / getVictimResources / splitVictimTasks / getResources
  bool reclaim_validator(const Scenario &sc) {
    const Job &rj = J[vjob(sc.preemptor)];
    const double *req = tasks_to_allocate_init_resource(sc.preemptor, false);
    std::map<int, std::vector<Quant>> by_queue;
    for (const auto &kv : sc.victims) {
      const Job &vj = J[kv.first];
      std::vector<Quant> res;
      std::vector<int> core, elastic;
      for (size_t k = 0; k < vj.podsets.size(); k++) {
        std::vector<int> sub;
        for (int ti : kv.second)
          if (T[ti].podset == vj.podsets[k]) sub.push_back(ti);
        if (sub.empty()) continue;
        int mn = PS[vj.podsets[k]].min_available;
        for (size_t i = 0; i < sub.size(); i++) ((int)i < mn ? core : elastic).push_back(sub[i]);
      }
      auto get_resources = [&](const std::vector<int> &tasks, Quant &out) {
        int n = 0;
        for (int r = 0; r < QR; r++) out.v[r] = 0;
        for (int ti : tasks) {
          if (cfg.allow_consolidating_reclaim && (T[ti].status & kActiveAllocated)) continue;
          n++;
          for (int r = 0; r < QR; r++) out.v[r] += T[ti].req[r];
        }
        return n > 0;
      };
      for (int ti : elastic) {
        Quant x;
        if (get_resources({ti}, x)) res.push_back(x);
      }
      Quant x;
      if (get_resources(core, x)) res.push_back(x);
      if (res.empty()) continue;
      auto &dst = by_queue[vj.queue];
      dst.insert(dst.end(), res.begin(), res.end());
    }
    return reclaimable(sim_queues, rj.queue, rj.preemptible, req, by_queue);
  }
  // consolidation.go:108-117 allPodsReallocated
  bool consolidation_validator(const Scenario &sc) {
    for (const auto &kv : sc.victims)
      for (int ti : kv.second)
        if (T[ti].status == KAI_POD_RELEASING) return false;
    return true;
  }

  // actions/common/action.go:67-122 GetJobsToAllocate + TryToVirtuallyAllocatePreemptorAndGetVictims
  bool try_virtually_allocate(const Scenario &sc, const std::vector<int> &nodes, const std::vector<int> &victim_tasks) {
    const int pj = vjob(sc.preemptor);
    std::vector<char> is_victim_job(NJ, 0), in_set(NJ, 0);
    std::vector<int> vs;
    for (int ji = 0; ji < NJ; ji++)
      if (job_count(J[ji], KAI_POD_PENDING) > 0) in_set[ji] = 1;
    for (int ti : victim_tasks) {
      in_set[T[ti].job] = 1;
      is_victim_job[T[ti].job] = 1;
    }
    in_set[pj] = 1;
    for (int ji = 0; ji < NJ; ji++)
      if (in_set[ji]) vs.push_back(ji == pj ? sc.preemptor : ji);
    JobsOrder jo;
    jo.init(this, false);
    init_jobs_order_views(jo, vs, OrderOpts());
    bool preemptor_allocated = false;
    while (!jo.is_empty()) {
      int v = jo.pop_next_job();
      if (v < 0) break;
      int ji = vjob(v);
      if (!is_victim_job[ji] && ji != pj) continue;
      if (getenv("KAI_ORACLE_TRACE")) fprintf(stderr, "[solver]     sim pops job %d%s\n", ji, ji == pj ? " (preemptor)" : "");
      tasks_to_allocate_init_resource(v, false);
      if (ji != pj) {
        allocate_job(v, &nodes, true);
        continue;
      }
      if (!allocate_job(v, &nodes, true)) return false;
      preemptor_allocated = true;
    }
    return preemptor_allocated;
  }

  // by_pod_solver.go:124-143,229-253 runSimulation + tryScenarioWithEvictedVictims + handleScenarioSolution
  SolveResult run_simulation(Scenario &sc, const std::vector<char> &feasible, const std::vector<int> &victim_tasks) {
    SolveResult res;
    std::vector<int> nodes;
    for (int n = 0; n < N; n++)
      if (feasible[n]) nodes.push_back(n);
    if (!try_virtually_allocate(sc, nodes, victim_tasks)) return res;  // has = false: keep looking
    std::vector<int> preempted, pipelined;
    for (int ti : victim_tasks) {
      if (T[ti].status == KAI_POD_RELEASING)
        preempted.push_back(ti);
      else if (T[ti].status == KAI_POD_PIPELINED)
        pipelined.push_back(ti);
    }
    res.has = true;
    // session_plugins.go:135-164: every registered validator must accept (reclaim: proportion + minruntime;
    // preempt: minruntime only; consolidation passes its own closure)
    bool valid = solver_kind == SOLVER_RECLAIM ? (reclaim_validator(sc) && minruntime_validator(sc, true))
                 : (solver_kind == SOLVER_CONSOLIDATION ? consolidation_validator(sc) : minruntime_validator(sc, false));
    if (!valid) {
      stmt_discard();
      return res;
    }
    res.victim_tasks = preempted;
    res.victim_tasks.insert(res.victim_tasks.end(), pipelined.begin(), pipelined.end());
    // getVictimJobsFromVictimTasks: the task-group representatives that hold the victim tasks
    std::map<int, std::vector<int>> groups;
    for (int ti : res.victim_tasks) {
      int ji = T[ti].job;
      bool exists = false;
      for (int g : groups[ji]) {
        std::vector<int> gt = v_all_tasks(g);
        if (std::find(gt.begin(), gt.end(), ti) != gt.end()) exists = true;
      }
      if (exists) continue;
      for (int g : sc.task_groups[ji]) {
        std::vector<int> gt = v_all_tasks(g);
        if (std::find(gt.begin(), gt.end(), ti) != gt.end()) {
          groups[ji].push_back(g);
          break;
        }
      }
    }
    for (auto &kv : groups) res.victim_jobs.insert(res.victim_jobs.end(), kv.second.begin(), kv.second.end());
    res.solved = true;
    return res;
  }

  // by_pod_solver.go:69-122,145-201 byPodSolver.solve
  SolveResult bypod_solve(Scenario &sc, std::vector<char> &feasible) {
    ops.clear();
    first_undo_built = (size_t)-1;  // session.Statement()
    for (int ti : sc.recorded_tasks) stmt_evict(ti);
    if (sc.potential_tasks.empty()) {
      if (!sc.recorded_tasks.empty()) {
        SolveResult r = run_simulation(sc, feasible, sc.recorded_tasks);
        if (r.has) return r;
      }
    } else {
      int latest = T[sc.potential_tasks.back()].job;
      std::vector<int> nodes;  // getNodesOfJob: distinct nodes of all pods of the job (canonical: ascending index)
      for (int s2 : J[latest].podsets)
        for (int ti : PS[s2].tasks)
          if (T[ti].node >= 0 && std::find(nodes.begin(), nodes.end(), T[ti].node) == nodes.end())
            nodes.push_back(T[ti].node);
      std::sort(nodes.begin(), nodes.end());
      for (int node : nodes) {
        int cp = stmt_checkpoint();
        std::vector<int> potential = scenario_victims_from_node(sc, node);
        for (int ti : potential) stmt_evict(ti);
        std::vector<int> added;
        for (int ti : potential)
          if (!feasible[T[ti].node]) {
            feasible[T[ti].node] = 1;
            added.push_back(T[ti].node);
          }
        std::vector<int> victim_tasks = sc.recorded_tasks;
        victim_tasks.insert(victim_tasks.end(), potential.begin(), potential.end());
        SolveResult r = run_simulation(sc, feasible, victim_tasks);
        if (r.has) return r;
        for (int n : added) feasible[n] = 0;
        stmt_rollback(cp);
      }
    }
    stmt_discard();
    SolveResult none;
    none.has = true;
    return none;
  }

  // the victims queue of the action (reclaim.go:121-143 / consolidation.go:119-157 + utils/action.go:20-52)
  void build_victims_queue(JobsOrder &jo, int pending_job) {
    jo.init(this, true);
    std::vector<int> vs;
    OrderOpts op;
    if (solver_kind == SOLVER_RECLAIM) {
      op.filter_non_preemptible = true;
      op.filter_non_active_allocated = true;
      for (int ji = 0; ji < NJ; ji++) {
        if (J[ji].queue == J[pending_job].queue) continue;
        if (!minruntime_filter(true, pending_job, ji)) continue;  // ssn.ReclaimVictimFilter (reclaim.go:134-136)
        vs.push_back(ji);
      }
    } else if (solver_kind == SOLVER_PREEMPT) {  // preempt.go:125-161 + utils/action.go:20-52
      for (int ji = 0; ji < NJ; ji++) {
        bool alive = false;
        for (int s2 : J[ji].podsets)
          for (int ti : PS[s2].tasks)
            if (T[ti].status & kAlive) alive = true;
        if (!alive) continue;
        if (!J[ji].preemptible) continue;
        if (J[ji].priority >= J[pending_job].priority) continue;
        if (J[ji].queue != J[pending_job].queue) continue;
        if (ji == pending_job) continue;
        if (job_count(J[ji], kActiveAllocated) == 0) continue;
        if (!minruntime_filter(false, pending_job, ji)) continue;  // ssn.PreemptVictimFilter (preempt.go:147-149)
        vs.push_back(ji);
      }
    } else {
      int counter = 0;
      for (int ji = 0; ji < NJ; ji++) {
        bool alive = false;
        for (int s2 : J[ji].podsets)
          for (int ti : PS[s2].tasks)
            if (T[ti].status & kAlive) alive = true;
        if (!alive) continue;
        if (!J[ji].preemptible) continue;
        if (ji == pending_job) continue;
        if (cfg.max_consolidation_preemptees != -1 && counter > cfg.max_consolidation_preemptees) continue;
        if (job_count(J[ji], kActiveAllocated) == 0) continue;
        counter++;
        vs.push_back(ji);
      }
    }
    init_jobs_order_views(jo, vs, op);
  }

  // job_solver.go:90-118 solvePartialJob + pod_scenario_builder.go
  SolveResult solve_partial(const SolveState &state, int pending_job, int partial, const std::vector<char> &base_feasible) {
    std::vector<char> feasible = base_feasible;
    for (int ti : state.recorded_tasks)
      if (T[ti].node >= 0) feasible[T[ti].node] = 1;
    Scenario sc;
    scenario_init(sc, partial, state.recorded_jobs);
    std::vector<char> recorded_set(NT, 0);
    for (int rv : state.recorded_jobs)
      for (int ti : v_all_tasks(rv)) recorded_set[ti] = 1;
    JobsOrder victims_queue;
    build_victims_queue(victims_queue, pending_job);
    IdleGpusFilter filter;
    idle_filter_init(filter, sc);
    bool first = true;
    for (;;) {
      bool need_add = !first;
      first = false;
      bool have = false;
      for (;;) {
        if (need_add) {
          bool added = false;
          while (!added) {  // GetNextScenario / addNextPotentialVictims
            if (victims_queue.is_empty()) break;
            int next = victims_queue.pop_next_job();
            if (next < 0) break;
            bool has_more = false;
            std::vector<int> tasks = tasks_to_evict(next, has_more);
            bool recorded_hit = false;
            for (int ti : tasks)
              if (recorded_set[ti]) recorded_hit = true;
            if (recorded_hit) {
              std::vector<int> remaining;
              for (int ti : v_all_tasks(next))
                if (!recorded_set[ti]) remaining.push_back(ti);
              if (!remaining.empty()) victims_queue.push_job(make_clone(next, remaining));
              continue;
            }
            if (has_more) {
              std::vector<int> remaining;
              for (int ti : v_all_tasks(next))
                if (std::find(tasks.begin(), tasks.end(), ti) == tasks.end()) remaining.push_back(ti);
              victims_queue.push_job(make_clone(next, remaining));
            }
            scenario_add_potential(sc, tasks);
            added = true;
          }
          if (!added) break;
        }
        if (idle_filter_check(filter, sc)) {
          have = true;
          break;
        }
        need_add = true;
      }
      if (!have) break;
      stats.kernel_launches++;  // scenarios simulated (metrics.IncScenarioSimulatedByAction)
      if (getenv("KAI_ORACLE_TRACE")) {
        fprintf(stderr, "[solver] job %d partial %zu tasks; recorded:", pending_job, sc.pending_tasks.size());
        for (int ti : sc.recorded_tasks) fprintf(stderr, " %d", ti);
        fprintf(stderr, " potential:");
        for (int ti : sc.potential_tasks) fprintf(stderr, " %d", ti);
        fprintf(stderr, "\n");
      }
      SolveResult r = bypod_solve(sc, feasible);
      if (getenv("KAI_ORACLE_TRACE")) {
        fprintf(stderr, "[solver]   -> solved %d victims:", (int)r.solved);
        for (int ti : r.victim_tasks) fprintf(stderr, " %d(st %d node %d)", ti, T[ti].status, T[ti].node);
        fprintf(stderr, "\n");
      }
      if (r.solved) return r;
    }
    SolveResult none;
    return none;
  }

  // job_solver.go:47-88 Solve + :120-148 getPartialJobRepresentative
  bool solve_job(int ji, const std::vector<char> &base_feasible) {
    SolveState state;
    int original_active = job_count(J[ji], kActiveUsed);
    std::vector<int> tta = tasks_to_allocate(ji, false);
    std::vector<int> pending;
    bool have_statement = false;
    for (size_t i = 0; i < tta.size(); i++) {
      pending.push_back(tta[i]);
      bool satisfactory = pending.size() == tta.size();
      int partial = make_clone(ji, pending);
      {
        View &pv = views[partial - NJ];
        for (size_t k = 0; k < pv.ps_tasks.size(); k++)
          if (!pv.ps_tasks[k].empty()) pv.ps_min[k] = (int)pv.ps_tasks[k].size();
      }
      SolveResult r = solve_partial(state, ji, partial, base_feasible);
      if (!r.solved) {
        have_statement = false;
        break;
      }
      if (!satisfactory) stmt_discard();
      have_statement = satisfactory;
      state.recorded_tasks = r.victim_tasks;
      state.recorded_jobs = r.victim_jobs;
    }
    int active = job_count(J[ji], kActiveUsed);
    bool solved = true;
    for (int s2 : J[ji].podsets)  // IsGangSatisfied
      if (podset_count(PS[s2], kActiveUsed) < PS[s2].min_available) solved = false;
    if (original_active >= active) solved = false;
    if (!have_statement) {
      ops.clear();
      first_undo_built = (size_t)-1;
    }
    return solved;
  }
  // actions/common/feasible_nodes.go:11-26
  std::vector<char> feasible_nodes_for_job(int ji) {
    std::vector<char> f(N, 1);
    for (int s2 : J[ji].podsets)
      for (int ti : PS[s2].tasks)
        if (!task_requires_gpu(T[ti])) return f;
    for (int n = 0; n < N; n++) f[n] = (I(KAI_RES_GPU, n) > 0 || L(KAI_RES_GPU, n) > 0) ? 1 : 0;
    return f;
  }

  // actions/common/minimal_job_comparison.go
  struct MinimalReps {
    std::map<int, int> rep;  // signature -> job
  };
  bool req_less_equal(int a, int b) const {  // ResourceRequirements.LessEqual (resource_requirment.go:126-140)
    for (int r = 0; r < R; r++) {
      if (r >= 3) {
        if (T[a].req[r] != 0 && T[a].req[r] > T[b].req[r]) return false;
      } else if (T[a].req[r] > T[b].req[r])
        return false;
    }
    return true;
  }
  std::vector<int> sorted_pending(int ji) const {  // extractSortedResourceRequests; sort.Slice == insertion sort for n <= 12
    std::vector<int> v;
    for (int s2 : J[ji].podsets)
      for (int ti : PS[s2].tasks)
        if (T[ti].status == KAI_POD_PENDING) v.push_back(ti);
    for (size_t i = 1; i < v.size(); i++)
      for (size_t j2 = i; j2 > 0 && req_less_equal(v[j2], v[j2 - 1]); j2--) std::swap(v[j2], v[j2 - 1]);
    return v;
  }
  bool easier_to_schedule(const MinimalReps &m, int ji) const {
    if (J[ji].signature < 0) return true;
    auto it = m.rep.find(J[ji].signature);
    if (it == m.rep.end()) return true;
    std::vector<int> a = sorted_pending(ji), b = sorted_pending(it->second);
    if (a.empty() || b.empty()) return false;
    if (b.size() > a.size()) return true;
    for (size_t i = 0; i < a.size(); i++) {
      if (i >= b.size()) return false;
      if (req_less_equal(a[i], b[i])) {
        if (req_less_equal(b[i], a[i])) continue;
        return true;
      }
    }
    return false;
  }
  void update_representative(MinimalReps &m, int ji) const {
    if (J[ji].signature < 0) return;
    auto it = m.rep.find(J[ji].signature);
    if (it != m.rep.end()) {
      std::vector<int> a = sorted_pending(ji), b = sorted_pending(it->second);
      bool smaller = !(a.empty() || b.empty()) && a.size() <= b.size();
      if (smaller)
        for (size_t i = 0; i < a.size(); i++)
          if (!req_less_equal(a[i], b[i])) smaller = false;
      if (!smaller) return;
    }
    m.rep[J[ji].signature] = ji;
  }

  // ---------------- actions/reclaim/reclaim.go:46-119 ----------------
  void run_reclaim() {
    solver_kind = SOLVER_RECLAIM;
    views.clear();
    JobsOrder jo;
    jo.init(this, false);
    init_jobs_order(jo, true, true);
    std::map<int, MinimalReps> failed_by_queue;
    while (!jo.is_empty()) {
      int ji = jo.pop_next_job();
      if (ji < 0) break;
      if (!can_reclaim_resources(ji)) continue;
      MinimalReps &reps = failed_by_queue[J[ji].queue];
      if (use_signatures && !easier_to_schedule(reps, ji)) continue;
      tasks_to_allocate_init_resource(ji, false);
      sim_queues = Q;  // OnJobSolutionStart
      std::vector<char> feasible = feasible_nodes_for_job(ji);
      ops.clear();
      first_undo_built = (size_t)-1;
      bool ok = solve_job(ji, feasible);
      if (ok) {
        stmt_commit();
        r_visits.push_back({ji, 1});
      } else {
        ops.clear();
        first_undo_built = (size_t)-1;
        update_representative(reps, ji);
        r_visits.push_back({ji, 0});
      }
    }
  }
  // ---------------- actions/stalegangeviction/stalegangeviction.go:29-95 ----------------
  void run_stale_gang_eviction() {
    if (cfg.staleness_grace_period_s < 0) return;  // :47-50 negative duration means no eviction
    for (int ji = 0; ji < NJ; ji++) {
      const Job &j = J[ji];
      // :42-57 a nil TimeStamp is stamped with time.Now(): zero time in stale state, which only a zero grace period lets
      // through; otherwise time.Since(TimeStamp) on the snapshot's single instant
      double in_stale = (!j_stale_since.empty() && j_stale_since[ji] > 0) ? now_s - j_stale_since[ji] : 0.0;
      if (in_stale < double(cfg.staleness_grace_period_s)) continue;
      if (job_count(j, KAI_POD_SUCCEEDED) > 0) continue;  // job_info.go:417-432 IsStale
      if (job_count(j, kActiveUsed) == 0) continue;
      bool stale = false;
      for (int s2 : j.podsets)
        if (podset_count(PS[s2], kActiveUsed) < PS[s2].min_available) stale = true;
      if (!stale) continue;
      for (int s2 : j.podsets)
        for (int ti : PS[s2].tasks) {
          if (!(T[ti].status & kActiveAllocated)) continue;
          // framework/session.go:127-150 Session.Evict: Releasing on the same node, deallocate handlers
          set_status(ti, KAI_POD_RELEASING);
          node_remove_task(ti, T[ti].node);
          node_add_task(ti);
          queue_allocate(ti, -1);
          pods_evicted++;
        }
      r_visits.push_back({ji, 1});
    }
  }
  // ---------------- actions/preempt/preempt.go:46-123 ----------------
  void run_preempt() {
    solver_kind = SOLVER_PREEMPT;
    views.clear();
    JobsOrder jo;
    jo.init(this, false);
    init_jobs_order(jo, true, true);
    std::map<int, MinimalReps> failed_by_queue;
    while (!jo.is_empty()) {
      int ji = jo.pop_next_job();
      if (ji < 0) break;
      MinimalReps &reps = failed_by_queue[J[ji].queue];
      if (use_signatures && !easier_to_schedule(reps, ji)) continue;
      tasks_to_allocate_init_resource(ji, false);
      double req[QR] = {0, 0, 0};
      for (int ti : tasks_to_allocate(ji, false))
        for (int r = 0; r < QR; r++) req[r] += T[ti].req[r];
      bool ok = false;
      ops.clear();
      first_undo_built = (size_t)-1;
      if (!non_preemptible_over_quota(ji, req)) {
        std::vector<char> feasible = feasible_nodes_for_job(ji);
        ok = solve_job(ji, feasible);
      }
      if (ok) {
        stmt_commit();
        r_visits.push_back({ji, 1});
      } else {
        ops.clear();
        first_undo_built = (size_t)-1;
        update_representative(reps, ji);
        r_visits.push_back({ji, 0});
      }
    }
  }
  // ---------------- actions/consolidation/consolidation.go:32-106 ----------------
  void run_consolidation() {
    solver_kind = SOLVER_CONSOLIDATION;
    views.clear();
    if (cfg.max_consolidation_preemptees == 0) return;
    JobsOrder jo;
    jo.init(this, false);
    {
      std::vector<int> vs;
      for (int ji = 0; ji < NJ; ji++) vs.push_back(ji);
      OrderOpts op;
      op.filter_non_pending = op.filter_unready = op.filter_non_preemptible = true;
      init_jobs_order_views(jo, vs, op);
    }
    MinimalReps reps;
    while (!jo.is_empty()) {
      int ji = jo.pop_next_job();
      if (ji < 0) break;
      if (use_signatures && !easier_to_schedule(reps, ji)) continue;
      tasks_to_allocate_init_resource(ji, false);
      // utils/action.go:130-160 IsEnoughGPUsAllocatableForJob
      double sum = 0, want = 0;
      for (int n = 0; n < N; n++)
        if (nflags[n] & KAI_NODE_READY) sum += I(KAI_RES_GPU, n), sum += L(KAI_RES_GPU, n);
      for (int ti : tasks_to_allocate(ji, false)) want += T[ti].req[KAI_RES_GPU];
      bool ok = false;
      ops.clear();
      first_undo_built = (size_t)-1;
      if (sum >= want) {
        std::vector<char> feasible = feasible_nodes_for_job(ji);
        ok = solve_job(ji, feasible);
      }
      if (ok) {
        stmt_commit();
        r_visits.push_back({ji, 1});
      } else {
        ops.clear();
        first_undo_built = (size_t)-1;
        update_representative(reps, ji);
        r_visits.push_back({ji, 0});
      }
    }
  }

  // ---------------- actions/allocate/allocate.go:46-111 ----------------
  void run_allocate() {
    JobsOrder jo;
    jo.init(this, false);
    init_jobs_order(jo, true, true);
    while (!jo.is_empty()) {
      int ji = jo.pop_next_job();
      if (ji < 0) break;
      ops.clear();
      first_undo_built = (size_t)-1;
      bool ok = allocate_job(ji, nullptr, false);
      if (ok) {
        if (should_pipeline_job(ji)) stmt_convert_all_allocated_to_pipelined(ji);
        stmt_commit();
        r_visits.push_back({ji, 1});
        if (has_tasks_to_allocate(ji, true)) {
          jo.push_job(ji);
          continue;
        }
      } else {
        stmt_discard();
        r_visits.push_back({ji, 0});
      }
    }
  }

  void fill_result(kai_result *out) {
    r_task_node.resize(NT);
    r_task_status.resize(NT);
    for (int t = 0; t < NT; t++) {
      r_task_node[t] = T[t].node;
      r_task_status[t] = T[t].status;
    }
    r_fair.assign((size_t)QR * NQ, 0);
    r_alloc.assign((size_t)QR * NQ, 0);
    r_alloc_np.assign((size_t)QR * NQ, 0);
    r_request.assign((size_t)QR * NQ, 0);
    for (int q = 0; q < NQ; q++)
      for (int r = 0; r < QR; r++) {
        r_fair[(size_t)r * NQ + q] = Q[q].s[r].fair;
        r_alloc[(size_t)r * NQ + q] = Q[q].s[r].allocated;
        r_alloc_np[(size_t)r * NQ + q] = Q[q].s[r].alloc_np;
        r_request[(size_t)r * NQ + q] = Q[q].s[r].request;
      }
    for (int r = 0; r < QR; r++) r_total[r] = total[r];
    r_idle = idle;
    r_rel = rel;
    memset(out, 0, sizeof(*out));
    out->n_tasks = NT;
    out->task_node = r_task_node.data();
    out->task_status = r_task_status.data();
    out->n_visits = (int)r_visits.size();
    out->visits = r_visits.data();
    out->n_queues = NQ;
    out->queue_fair_share = r_fair.data();
    out->queue_allocated = r_alloc.data();
    out->queue_allocated_non_preemptible = r_alloc_np.data();
    out->queue_request = r_request.data();
    out->total_resource = r_total;
    out->n_nodes = N;
    out->node_idle = r_idle.data();
    out->node_releasing = r_rel.data();
    out->pods_placed = pods_placed;
    out->pods_evicted = pods_evicted;
  }
};

extern "C" {

int kai_oracle_create(const kai_config *cfg, kai_oracle **out) {
  if (!cfg || !out || cfg->abi_version != KAI_ABI_VERSION) return KAI_ERR_INVALID;
  kai_oracle *o = new kai_oracle();
  o->cfg = *cfg;
  o->use_signatures = cfg->use_scheduling_signatures != 0;
  memset(&o->stats, 0, sizeof(o->stats));
  *out = o;
  return KAI_OK;
}

int kai_oracle_load_snapshot(kai_oracle *o, const kai_snapshot *s) {
  if (!o || !s || s->abi_version != KAI_ABI_VERSION) return KAI_ERR_INVALID;
  if (s->n_res < 4 || s->n_res > KAI_MAX_RES) {
    o->err = "n_res out of range";
    return KAI_ERR_INVALID;
  }
  o->R = s->n_res;
  o->N = s->n_nodes;
  o->NQ = s->n_queues;
  o->NJ = s->n_jobs;
  o->NS = s->n_podsets;
  o->NT = s->n_tasks;
  o->NPC = s->n_pred_classes;
  o->mask_words = (o->N + 31) / 32;
  size_t rn = (size_t)o->R * o->N;
  o->alloc.assign(s->node_allocatable, s->node_allocatable + rn);
  o->idle.assign(s->node_idle, s->node_idle + rn);
  o->rel.assign(s->node_releasing, s->node_releasing + rn);
  o->name_rank.assign(s->node_name_rank, s->node_name_rank + o->N);
  o->nflags.assign(s->node_flags, s->node_flags + o->N);
  o->gpu_count.resize(o->N);
  for (int n = 0; n < o->N; n++)
    o->gpu_count[n] = s->node_gpu_count ? s->node_gpu_count[n] : o->alloc[(size_t)KAI_RES_GPU * o->N + n];
  o->foreign.clear();
  if (s->node_foreign) o->foreign.assign(s->node_foreign, s->node_foreign + (size_t)3 * o->N);
  o->Q.assign(o->NQ, QueueAttr());
  for (int q = 0; q < o->NQ; q++) {
    QueueAttr &a = o->Q[q];
    a.parent = s->queue_parent[q];
    a.priority = s->queue_priority[q];
    a.creation = s->queue_creation[q];
    a.uid_rank = s->queue_uid_rank[q];
    for (int r = 0; r < QR; r++) {
      a.s[r].deserved = s->queue_deserved[(size_t)r * o->NQ + q];
      a.s[r].max_allowed = s->queue_limit[(size_t)r * o->NQ + q];
      a.s[r].oqw = s->queue_oqw[(size_t)r * o->NQ + q];
      a.s[r].usage = s->queue_usage ? s->queue_usage[(size_t)r * o->NQ + q] : 0.0;
    }
  }
  for (int q = 0; q < o->NQ; q++) {
    int p = o->Q[q].parent;
    if (p >= o->NQ || p == q) {
      o->err = "bad queue parent";
      return KAI_ERR_INVALID;
    }
    if (p >= 0) o->Q[p].children.push_back(q);
  }
  for (int t = 0; t < s->n_tasks; t++) {  // same refusal as the engine: shared-GPU pods are not restated here
    const double g = s->task_req[(size_t)t * s->n_res + KAI_RES_GPU];
    if (g != (double)(long long)g) {
      o->err = "fractional GPU request: GPU sharing is outside this oracle's scope";
      return KAI_ERR_UNSUPPORTED;
    }
  }
  o->now_s = s->now_s;
  o->q_preempt_mrt.clear();
  o->q_reclaim_mrt.clear();
  o->j_last_start.clear();
  o->j_stale_since.clear();
  if (s->queue_preempt_min_runtime_s) o->q_preempt_mrt.assign(s->queue_preempt_min_runtime_s, s->queue_preempt_min_runtime_s + o->NQ);
  if (s->queue_reclaim_min_runtime_s) o->q_reclaim_mrt.assign(s->queue_reclaim_min_runtime_s, s->queue_reclaim_min_runtime_s + o->NQ);
  if (s->job_last_start_s) o->j_last_start.assign(s->job_last_start_s, s->job_last_start_s + o->NJ);
  if (s->job_stale_since_s) o->j_stale_since.assign(s->job_stale_since_s, s->job_stale_since_s + o->NJ);
  o->J.assign(o->NJ, Job());
  o->PS.assign(o->NS, PodSet());
  o->T.assign(o->NT, Task());
  for (int j = 0; j < o->NJ; j++) {
    Job &jb = o->J[j];
    jb.queue = s->job_queue[j];
    jb.priority = s->job_priority[j];
    jb.order_rank = s->job_order_rank[j];
    jb.preemptible = (s->job_flags[j] & KAI_JOB_PREEMPTIBLE) != 0;
    jb.signature = s->job_signature ? s->job_signature[j] : -1;
    for (int ps = s->job_podset_begin[j]; ps < s->job_podset_begin[j + 1]; ps++) {
      jb.podsets.push_back(ps);
      o->PS[ps].job = j;
      o->PS[ps].min_available = s->podset_min_available[ps];
      for (int t = s->podset_task_begin[ps]; t < s->podset_task_begin[ps + 1]; t++) {
        o->PS[ps].tasks.push_back(t);
        Task &tk = o->T[t];
        tk.job = j;
        tk.podset = ps;
        tk.status = s->task_status[t];
        tk.node = s->task_node[t];
        for (int r = 0; r < o->R; r++) tk.req[r] = s->task_req[(size_t)t * o->R + r];
        tk.order_rank = s->task_order_rank[t];
        tk.nominated = s->task_nominated ? s->task_nominated[t] : -1;
        tk.pred_class = s->task_pred_class ? s->task_pred_class[t] : -1;
        bool on = (tk.status & kActiveUsed) && tk.node >= 0;
        tk.on_node.clear();
        tk.on_status.clear();
        if (on) {
          tk.on_node.push_back(tk.node);
          tk.on_status.push_back(tk.status);
        }
        if (!on && !(tk.status & kActiveUsed)) tk.node = -1;
      }
    }
  }
  o->pred_mask.clear();
  if (s->pred_mask && o->NPC > 0)
    o->pred_mask.assign(s->pred_mask, s->pred_mask + (size_t)o->NPC * o->mask_words);
  o->open_session();
  o->topo_level_begin.clear();
  o->node_domain.clear();
  o->job_topology.clear();
  o->job_req_level.clear();
  o->job_pref_level.clear();
  o->topos.clear();
  if (s->n_topologies > 0 && s->topology_level_begin && s->node_domain) {
    o->topo_level_begin.assign(s->topology_level_begin, s->topology_level_begin + s->n_topologies + 1);
    size_t nl = (size_t)o->topo_level_begin.back();
    o->node_domain.assign(s->node_domain, s->node_domain + nl * (size_t)o->N);
    o->build_topologies();
  }
  // SubGroupSet tree, normalised: with no tree in the snapshot every job has one root set holding all its PodSets
  {
    const int NJ = o->NJ, NS = o->NS;
    o->set_parent.clear();
    o->set_rank.clear();
    o->set_con.clear();
    o->job_root_set.assign(NJ, -1);
    o->ps_set.assign(NS, -1);
    o->ps_con.assign(NS, std::array<int, 3>{-1, -1, -1});
    if (s->job_sgs_begin && s->sgs_parent && s->podset_sgs) {
      const int G = s->n_subgroup_sets;
      for (int g = 0; g < G; g++) {
        o->set_parent.push_back(s->sgs_parent[g]);
        o->set_rank.push_back(s->sgs_name_rank ? s->sgs_name_rank[g] : g);
        o->set_con.push_back({s->sgs_topology ? s->sgs_topology[g] : -1, s->sgs_required_level ? s->sgs_required_level[g] : -1,
                              s->sgs_preferred_level ? s->sgs_preferred_level[g] : -1});
      }
      for (int j = 0; j < NJ; j++) o->job_root_set[j] = s->job_sgs_begin[j];
      for (int ps = 0; ps < NS; ps++) {
        o->ps_set[ps] = s->podset_sgs[ps];
        if (s->podset_topology)
          o->ps_con[ps] = {s->podset_topology[ps], s->podset_required_level ? s->podset_required_level[ps] : -1,
                           s->podset_preferred_level ? s->podset_preferred_level[ps] : -1};
      }
    } else {
      for (int j = 0; j < NJ; j++) {
        o->job_root_set[j] = (int)o->set_parent.size();
        o->set_parent.push_back(-1);
        o->set_rank.push_back(0);
        o->set_con.push_back({s->job_topology ? s->job_topology[j] : -1, s->job_required_level ? s->job_required_level[j] : -1,
                              s->job_preferred_level ? s->job_preferred_level[j] : -1});
        for (int ps : o->J[j].podsets) o->ps_set[ps] = o->job_root_set[j];
      }
    }
    const int G = (int)o->set_parent.size();
    o->set_children.assign(G, {});
    o->set_podsets.assign(G, {});
    for (int g = 0; g < G; g++)
      if (o->set_parent[g] >= 0) o->set_children[o->set_parent[g]].push_back(g);
    for (auto &ch : o->set_children)
      std::sort(ch.begin(), ch.end(), [&](int a, int b) { return o->set_rank[a] < o->set_rank[b]; });
    for (int ps = 0; ps < NS; ps++)
      if (o->ps_set[ps] >= 0) o->set_podsets[o->ps_set[ps]].push_back(ps);
    o->job_general.assign(NJ, 0);
    std::vector<int> job_of_root(G, -1);
    for (int j = 0; j < NJ; j++)
      if (o->job_root_set[j] >= 0) job_of_root[o->job_root_set[j]] = j;
    for (int g = 0; g < G; g++) {
      if (o->set_con[g][0] == -1 && o->set_parent[g] < 0) continue;
      int root = g;
      while (o->set_parent[root] >= 0) root = o->set_parent[root];
      if (job_of_root[root] >= 0) o->job_general[job_of_root[root]] = 1;
    }
    for (int ps = 0; ps < NS; ps++)
      if (o->ps_con[ps][0] != -1) o->job_general[o->PS[ps].job] = 1;
  }
  o->loaded = true;
  o->pods_placed = o->pods_evicted = 0;
  memset(&o->stats, 0, sizeof(o->stats));
  return KAI_OK;
}

int kai_oracle_run(kai_oracle *o, kai_action action, kai_result *out) {
  if (!o || !out) return KAI_ERR_INVALID;
  if (!o->loaded) return KAI_ERR_STATE;
  o->r_visits.clear();
  o->pods_placed = o->pods_evicted = 0;
  switch (action) {
    case KAI_ACTION_ALLOCATE:
      o->run_allocate();
      break;
    case KAI_ACTION_RECLAIM:
      o->run_reclaim();
      break;
    case KAI_ACTION_CONSOLIDATION:
      o->run_consolidation();
      break;
    case KAI_ACTION_PREEMPT:
      o->run_preempt();
      break;
    case KAI_ACTION_STALEGANGEVICTION:
      o->run_stale_gang_eviction();
      break;
    default:
      o->err = "action not implemented by the oracle";
      return KAI_ERR_UNSUPPORTED;
  }
  o->fill_result(out);
  return KAI_OK;
}

int kai_oracle_fair_share(kai_oracle *o, kai_result *out) {
  if (!o || !out) return KAI_ERR_INVALID;
  if (!o->loaded) return KAI_ERR_STATE;
  o->fill_result(out);
  return KAI_OK;
}

int kai_oracle_stats(kai_oracle *o, kai_stats *out) {
  if (!o || !out) return KAI_ERR_INVALID;
  *out = o->stats;
  out->algorithmic_bytes = o->stats.nodes_scanned * ((2 * o->R + 1) * 8 + 4);
  return KAI_OK;
}

int kai_oracle_set_threads(kai_oracle *o, int n) {
  if (!o || n < 1) return KAI_ERR_INVALID;
  o->n_threads = n;
  return KAI_OK;
}

void kai_oracle_destroy(kai_oracle *o) { delete o; }
const char *kai_oracle_last_error(const kai_oracle *o) { return o ? o->err.c_str() : "null oracle"; }

double kai_oracle_binpack_score(double mn, double mx, double cur, double overall) {
  return binpack_score(mn, mx, cur, overall);
}
double kai_oracle_spread_score(double non_allocated, double count) { return spread_score(non_allocated, count); }
double kai_oracle_topology_position_score(int i, int n) { return topology_position_score(i, n); }

double kai_oracle_set_resource_share(int n, double total, double k_value, const double *deserved,
                                     const double *limit, const double *oqw, const double *request,
                                     const double *usage, const int32_t *priority, const int64_t *creation,
                                     const int32_t *uid_rank, double *fair_share) {
  std::vector<QueueAttr> Q(n);
  std::vector<int> group;
  for (int i = 0; i < n; i++) {
    Q[i].priority = priority[i];
    Q[i].creation = creation[i];
    Q[i].uid_rank = uid_rank[i];
    Share &s = Q[i].s[0];
    s.deserved = deserved[i];
    s.max_allowed = limit[i];
    s.oqw = oqw[i];
    s.request = request[i];
    s.usage = usage ? usage[i] : 0;
    s.fair = fair_share[i];
    group.push_back(i);
  }
  double rem = set_resource_share(total, k_value, Q, group, 0);
  for (int i = 0; i < n; i++) fair_share[i] = Q[i].s[0].fair;
  return rem;
}

// divideOverQuotaResource (resource_division.go:111-144) alone, on FairShare values the caller has set; same arrays
double kai_oracle_divide_over_quota(int n, double amount, double k_value, const double *deserved, const double *limit,
                                    const double *oqw, const double *request, const double *usage,
                                    const int32_t *priority, const int64_t *creation, const int32_t *uid_rank,
                                    double *fair_share) {
  std::vector<QueueAttr> Q(n);
  std::vector<int> group;
  for (int i = 0; i < n; i++) {
    Q[i].priority = priority[i];
    Q[i].creation = creation[i];
    Q[i].uid_rank = uid_rank[i];
    Share &s = Q[i].s[0];
    s.deserved = deserved[i];
    s.max_allowed = limit[i];
    s.oqw = oqw[i];
    s.request = request[i];
    s.usage = usage ? usage[i] : 0;
    s.fair = fair_share[i];
    group.push_back(i);
  }
  double rem = divide_over_quota_resource(amount, k_value, Q, group, 0);
  for (int i = 0; i < n; i++) fair_share[i] = Q[i].s[0].fair;
  return rem;
}

// actions/common/minimal_job_comparison.go on two jobs of the loaded snapshot that share a signature:
// MinimalJobRepresentatives{representative}.IsEasierToSchedule(job) (:27-36,50-81)
int kai_oracle_job_easier_to_schedule(kai_oracle *o, int job, int representative) {
  kai_oracle::MinimalReps m;
  m.rep[o->J[representative].signature] = representative;
  return o->easier_to_schedule(m, job) ? 1 : 0;
}
// ... and UpdateRepresentative(job) (:38-48,83-105): 1 = the job replaced the representative
int kai_oracle_job_replaces_representative(kai_oracle *o, int job, int representative) {
  kai_oracle::MinimalReps m;
  m.rep[o->J[representative].signature] = representative;
  o->update_representative(m, job);
  return m.rep[o->J[representative].signature] == job ? 1 : 0;
}

// framework.Statement (statement.go) on the loaded snapshot: kinds[i] = 0 Evict(task), 1 Pipeline(task, node,
// updateTaskIfExistsOnNode = true), 2 Allocate(task, node), 3 undoOperation(index = task[i]), 4 Discard, 5 Pipeline with
// updateTaskIfExistsOnNode = false, 6 Rollback(checkpoint = task[i]); returns the operation count (= Checkpoint()) or a
// negative error; read the outcome with kai_oracle_fair_share
int kai_oracle_statement_exercise(kai_oracle *o, int n_ops, const int32_t *kinds, const int32_t *task, const int32_t *node) {
  if (!o || !o->loaded) return KAI_ERR_STATE;
  for (int i = 0; i < n_ops; i++) {
    switch (kinds[i]) {
      case 0:
        if (o->T[task[i]].node < 0) return KAI_ERR_INVALID;  // "node doesn't exist in session" (statement.go:70-74)
        o->stmt_evict(task[i]);
        break;
      case 1:
        o->stmt_pipeline(task[i], node[i], true);
        break;
      case 2:
        o->stmt_allocate(task[i], node[i]);
        break;
      case 3:
        o->undo_operation(task[i]);
        break;
      case 4:
        o->stmt_discard();
        break;
      case 5:  // Pipeline(task, node, updateTaskIfExistsOnNode = false)
        o->stmt_pipeline(task[i], node[i], false);
        break;
      case 6:  // Rollback(checkpoint = task[i]) (statement.go:48-61)
        o->stmt_rollback(task[i]);
        break;
      default:
        return KAI_ERR_INVALID;
    }
  }
  return (int)o->ops.size();  // = what Checkpoint() would return now
}

// scheduler_util.PriorityQueue (priority_queue.go:50-118) over container/heap, on ints with `<`: ops[i] = 0 push(vals[i])
// (with the max-size eviction heap.Remove(maxQueueSize) when max_size >= 0), 1 pop, 2 peek, 3 set items[0] = vals[i]
// and Fix(0), 4 len; out[i] receives the popped / peeked value or the length (INT32_MIN when the queue is empty)
void kai_oracle_priority_queue_exercise(int max_size, int n_ops, const int32_t *ops, const int32_t *vals, int32_t *out) {
  GoHeap<int> h;
  h.less = [](const int &a, const int &b) { return a < b; };
  for (int i = 0; i < n_ops; i++) {
    out[i] = INT32_MIN;
    switch (ops[i]) {
      case 0:
        h.push(vals[i]);
        if (max_size >= 0 && h.len() > max_size) h.remove(max_size);
        break;
      case 1:
        if (!h.empty()) out[i] = h.pop();
        break;
      case 2:
        if (!h.empty()) out[i] = h.peek();
        break;
      case 3:
        if (!h.empty()) {
          h.items[0] = vals[i];
          h.fix(0);
        }
        break;
      case 4:
        out[i] = h.len();
        break;
    }
  }
}

// idle_gpus/common.go:34-64 greedyMatchRequirements (both arrays sorted descending by the caller)
int kai_oracle_greedy_match(int n_req, const double *req, int n_holders, const double *capacity) {
  if (n_req < 0 || n_holders < 0 || (n_req && !req) || (n_holders && !capacity)) return KAI_ERR_INVALID;
  std::vector<double> r(req, req + n_req), c(capacity, capacity + n_holders);
  return kai_oracle::greedy_match_requirements(r, c) ? 1 : 0;
}

// podgroup_info.GetTasksToAllocate (allocation_info.go:27-54) of one job of the loaded snapshot: task indices in
// attempt order; real_allocation = the isRealAllocation argument; virtual_mask = tasks whose Releasing status is virtual
int kai_oracle_tasks_to_allocate(kai_oracle *o, int job, int real_allocation, int32_t *out, int cap) {
  if (!o || job < 0 || job >= o->NJ) return -1;
  o->vcache(job).tta_valid = false;
  const std::vector<int> &t = o->tasks_to_allocate(job, real_allocation != 0);
  for (size_t i = 0; i < t.size() && (int)i < cap; i++) out[i] = t[i];
  return (int)t.size();
}
int kai_oracle_set_task_virtual(kai_oracle *o, int task, int is_virtual) {  // PodInfo.IsVirtualStatus
  if (!o || task < 0 || task >= o->NT) return -1;
  o->T[task].is_virtual = is_virtual != 0;
  return 0;
}

// plugins/proportion/reclaimable/reclaimable.go on an explicit queue table (unit-level entry points for the reference's
// reclaimable_test.go).  share[q][r][5] = {Deserved, FairShare, Allocated, AllocatedNotPreemptible, MaxAllowed}, r in KAI_Q_*.
static void fill_queue_table(kai_oracle &o, int n, const int32_t *parent, const double *share, double saturation_multiplier) {
  o.cfg.saturation_multiplier = saturation_multiplier;
  o.NQ = n;
  o.Q.assign(n, QueueAttr());
  for (int q = 0; q < n; q++) {
    o.Q[q].parent = parent[q];
    for (int r = 0; r < QR; r++) {
      const double *x = share + ((size_t)q * QR + r) * 5;
      o.Q[q].s[r].deserved = x[0];
      o.Q[q].s[r].fair = x[1];
      o.Q[q].s[r].allocated = x[2];
      o.Q[q].s[r].alloc_np = x[3];
      o.Q[q].s[r].max_allowed = x[4];
    }
  }
}
// Reclaimable.CanReclaimResources (reclaimable.go:29-51) for one queue: share[3][4], req[3]
int kai_oracle_can_reclaim_resources(const double *share, const double *req, int preemptible) {
  QueueAttr q;
  for (int r = 0; r < QR; r++) {
    q.s[r].deserved = share[r * 4 + 0];
    q.s[r].fair = share[r * 4 + 1];
    q.s[r].allocated = share[r * 4 + 2];
    q.s[r].alloc_np = share[r * 4 + 3];
  }
  return kai_oracle::can_reclaim_from_share(q, req, preemptible != 0) ? 1 : 0;
}
// proportion.setFairShare (proportion.go:403-423) on an explicit queue tree: the level recursion over
// resource_division.SetResourcesShare.  in[q][3][4] = {Deserved, MaxAllowed, OverQuotaWeight, Request} per resource,
// fair_share[q][3] is written
int kai_oracle_set_fair_share_tree(int n_queues, const int32_t *parent, const int32_t *priority, const int64_t *creation,
                                   const int32_t *uid_rank, const double *in, const double *total, double k_value,
                                   double *fair_share) {
  kai_oracle o;
  o.cfg.k_value = k_value;
  o.NQ = n_queues;
  o.Q.assign(n_queues, QueueAttr());
  for (int q = 0; q < n_queues; q++) {
    o.Q[q].parent = parent[q];
    o.Q[q].priority = priority[q];
    o.Q[q].creation = creation[q];
    o.Q[q].uid_rank = uid_rank[q];
    for (int r = 0; r < QR; r++) {
      const double *x = in + ((size_t)q * QR + r) * 4;
      o.Q[q].s[r].deserved = x[0];
      o.Q[q].s[r].max_allowed = x[1];
      o.Q[q].s[r].oqw = x[2];
      o.Q[q].s[r].request = x[3];
    }
  }
  std::vector<int> top;
  for (int q = 0; q < n_queues; q++) {
    if (parent[q] >= 0)
      o.Q[parent[q]].children.push_back(q);
    else
      top.push_back(q);
  }
  o.set_fair_share_for_queues(total, top);
  for (int q = 0; q < n_queues; q++)
    for (int r = 0; r < QR; r++) fair_share[(size_t)q * QR + r] = o.Q[q].s[r].fair;
  return 0;
}

// capacity_policy.go:26-84 on an explicit queue tree (share[q][3][5] as above): mode 0 = resultsOverLimit +
// resultsWithNonPreemptibleOverQuota (IsJobOverQueueCapacity / IsTaskAllocationOnNodeOverCapacity), mode 1 = the quota
// check alone (IsNonPreemptibleJobOverQuota); returns IsSchedulable
int kai_oracle_capacity_schedulable(int n_queues, const int32_t *parent, const double *share, int queue, int preemptible,
                                    const double *req, int mode) {
  kai_oracle o;
  fill_queue_table(o, n_queues, parent, share, 1.0);
  o.NJ = 1;
  o.J.assign(1, Job());
  o.J[0].queue = queue;
  o.J[0].preemptible = preemptible != 0;
  return (mode == 0 ? o.over_capacity(0, req) : o.non_preemptible_over_quota(0, req)) ? 0 : 1;
}

int kai_oracle_node_entries(kai_oracle *o, int32_t *task, int32_t *node, int32_t *status, int cap) {
  if (!o) return -1;
  int n = 0;
  for (int t = 0; t < (int)o->T.size(); t++)
    for (size_t e = 0; e < o->T[t].on_node.size(); e++) {
      if (n < cap) {
        task[n] = t;
        node[n] = o->T[t].on_node[e];
        status[n] = o->T[t].on_status[e];
      }
      n++;
    }
  return n;
}
int kai_oracle_feasible_nodes(kai_oracle *o, int job, int32_t *out) {
  if (!o || job < 0 || job >= o->NJ) return KAI_ERR_INVALID;
  std::vector<char> f = o->feasible_nodes_for_job(job);
  for (int n = 0; n < o->N; n++) out[n] = f[n];
  return KAI_OK;
}
void kai_oracle_queue_attributes(const double *share, const double *total, double *out) {
  QueueAttr q;
  for (int r = 0; r < QR; r++) {
    const double *x = share + (size_t)r * 6;
    q.s[r].deserved = x[0];
    q.s[r].fair = x[1];
    q.s[r].allocated = x[2];
    q.s[r].alloc_np = x[3];
    q.s[r].max_allowed = x[4];
    q.s[r].request = x[5];
  }
  out[0] = dominant_share(q, total);
  for (int r = 0; r < QR; r++) {
    out[1 + r] = allocatable_share(q.s[r]);
    out[4 + r] = requestable_share(q.s[r]);
  }
}
int kai_oracle_compare_quantities(double a, double b) { return compare_quantities(a, b); }
int kai_oracle_quantities_relation(int kind, const double *a, const double *b) {
  if (kind == 1) return q_less_equal(a, b) ? 1 : 0;
  if (kind == 2) return q_less_equal(b, a) ? 0 : 1;  // :66-68 !other.LessEqual(rq)
  for (int r = 0; r < QR; r++)                        // :49-56 plain comparison, no "unlimited" handling
    if (a[r] >= b[r]) return 0;
  return 1;
}

// strategies.go: one reclaim strategy on two queue rows (share[3][5] each, as above) and a remaining share[3];
// strategy 0 = MaintainFairShareStrategy, 1 = GuaranteeDeservedQuotaStrategy
int kai_oracle_reclaim_strategy(int strategy, const double *reclaimer_share, const double *reclaimee_share,
                                const double *reclaimer_req, const double *remaining) {
  QueueAttr er, ee;
  for (int r = 0; r < QR; r++)
    for (int side = 0; side < 2; side++) {
      const double *x = (side ? reclaimee_share : reclaimer_share) + (size_t)r * 5;
      Share &sh = (side ? ee : er).s[r];
      sh.deserved = x[0];
      sh.fair = x[1];
      sh.allocated = x[2];
      sh.alloc_np = x[3];
      sh.max_allowed = x[4];
    }
  return (strategy == 0 ? kai_oracle::maintain_fair_share(ee, remaining)
                        : kai_oracle::guarantee_deserved_quota(er, ee, reclaimer_req, remaining))
             ? 1
             : 0;
}

// Reclaimable.Reclaimable (reclaimable.go:53-232): victims = (leaf queue, resources[3]) in the order given
int kai_oracle_reclaimable(int n_queues, const int32_t *parent, const double *share, double saturation_multiplier,
                           int reclaimer_queue, int preemptible, const double *req, int n_victims,
                           const int32_t *victim_queue, const double *victim_res) {
  kai_config cfg{};
  cfg.abi_version = KAI_ABI_VERSION;
  kai_oracle o;
  o.cfg = cfg;
  fill_queue_table(o, n_queues, parent, share, saturation_multiplier);
  std::map<int, std::vector<kai_oracle::Quant>> by_queue;
  for (int i = 0; i < n_victims; i++) {
    kai_oracle::Quant x;
    for (int r = 0; r < QR; r++) x.v[r] = victim_res[(size_t)i * QR + r];
    by_queue[victim_queue[i]].push_back(x);
  }
  return o.reclaimable(o.Q, reclaimer_queue, preemptible != 0, req, by_queue) ? 1 : 0;
}

// plugins/minruntime/resolver.go on the loaded snapshot's queue tree: reclaim != 0 -> getReclaimMinRuntime(method of
// the config, pending queue, victim queue), else getPreemptMinRuntime(victim queue); -1 = nil queue
double kai_oracle_min_runtime(kai_oracle *o, int reclaim, int pending_queue, int victim_queue) {
  return reclaim ? o->reclaim_min_runtime(pending_queue, victim_queue) : o->preempt_min_runtime(victim_queue);
}
int kai_oracle_min_runtime_protected(kai_oracle *o, int reclaim, int pending_job, int victim_job) {
  return o->minruntime_filter(reclaim != 0, pending_job, victim_job) ? 0 : 1;
}

int kai_oracle_queue_order(const double *l_share, const double *r_share, int l_priority, int r_priority,
                           int64_t l_creation, int64_t r_creation, const double *l_job_req,
                           const double *r_job_req, const double *total) {
  QueueAttr l, r;
  auto fill = [](QueueAttr &q, const double *s) {
    for (int i = 0; i < QR; i++) {
      q.s[i].deserved = s[i * 8 + 0];
      q.s[i].fair = s[i * 8 + 1];
      q.s[i].max_allowed = s[i * 8 + 2];
      q.s[i].oqw = s[i * 8 + 3];
      q.s[i].allocated = s[i * 8 + 4];
      q.s[i].alloc_np = s[i * 8 + 5];
      q.s[i].request = s[i * 8 + 6];
      q.s[i].usage = s[i * 8 + 7];
    }
  };
  fill(l, l_share);
  fill(r, r_share);
  l.priority = l_priority;
  r.priority = r_priority;
  l.creation = l_creation;
  r.creation = r_creation;
  return queue_order_result(l, r, l_job_req, r_job_req, nullptr, nullptr, total);
}

}  // extern "C"
