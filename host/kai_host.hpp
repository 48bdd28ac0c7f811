// kai_host.hpp — C++ host side above the C ABI, mirroring the reference's Action / Session surface.
//
// The reference's host code is Go; no Go toolchain exists in this image, so the shim a maintainer would write in Go
//  is written here in C++ with the reference's names and call sequence:
//
//   framework::Action            pkg/scheduler/framework/interface.go:41-47        Name(), Execute(ssn)
//   framework::RegisterAction    pkg/scheduler/framework/plugins.go:47-62          last registration wins
//   framework::GetAction         pkg/scheduler/conf_util/scheduler_conf_util.go:96-107
//   framework::Session           pkg/scheduler/framework/session.go:43-98          ClusterInfo, Statement(), Cache
//   framework::Statement         pkg/scheduler/framework/statement.go:36-663       Allocate / Pipeline / Evict / Commit
//   api::{NodeInfo,PodInfo,PodGroupInfo,QueueInfo,ClusterInfo}   pkg/scheduler/api/**
//   gpuengine::New(kind)         the replacement Actions: pack ssn.ClusterInfo -> kai_snapshot, kai_engine_run,
//                                replay the result through ssn.Statement()
//
// Header-only; links against libkaigpu.so (include/kai_engine.h).  Error behaviour: Action::Execute cannot return an
// error in the reference; here, as there, a failing engine call leaves the session untouched and is reported through
// Session::LastError() (the Go shim would run the stock action instead).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../include/kai_engine.h"

namespace kai_host {

namespace pod_status {  // pkg/scheduler/api/pod_status/pod_status.go:25-71 (same bit values as KAI_POD_*)
enum PodStatus : int {
  Pending = KAI_POD_PENDING, Gated = KAI_POD_GATED, Allocated = KAI_POD_ALLOCATED, Pipelined = KAI_POD_PIPELINED,
  Binding = KAI_POD_BINDING, Bound = KAI_POD_BOUND, Running = KAI_POD_RUNNING, Releasing = KAI_POD_RELEASING,
  Succeeded = KAI_POD_SUCCEEDED, Failed = KAI_POD_FAILED, Unknown = KAI_POD_UNKNOWN, Deleted = KAI_POD_DELETED,
};
inline bool IsActiveUsedStatus(int s) { return s & (Allocated | Pipelined | Binding | Bound | Running | Releasing); }
inline bool IsActiveAllocatedStatus(int s) { return s & (Allocated | Pipelined | Binding | Bound | Running); }
}  // namespace pod_status

namespace api {
using ResourceVector = std::vector<double>;  // resource_info/resource_vector.go: [cpu milli, memory, gpu, pods, ...]

struct PodInfo {  // pod_info/pod_info.go:70-112
  std::string UID, Job, SubGroupName, NodeName;
  int Status = pod_status::Pending;
  ResourceVector ResReq;
  long long OrderKey = 0;  // position under TaskOrderFn inside the job (priority, then UID)
  std::string NominatedNodeName;
  // required node affinity `kai.scheduler/type In NodeAffinityNames` (the only k8s Filter the reference's test DSL
  // produces, tasks_fake/tasks.go:98-116); the shim evaluates such node-local filters into pred_mask classes
  std::vector<std::string> NodeAffinityNames;
};
struct TopologyConstraintInfo {  // api/topology_info: empty Topology = no constraint
  std::string Topology, RequiredLevel, PreferredLevel;
};
struct PodSet {  // podgroup_info/subgroup_info/podset.go
  std::string Name;
  int MinAvailable = 1;
  TopologyConstraintInfo TopologyConstraint;
  std::string ParentSet;  // name of the SubGroupSet holding it ("" = the root)
};
struct SubGroupSet {  // podgroup_info/subgroup_info/subgroupset.go (flattened: Parent by name, "" = child of the root)
  std::string Name, Parent;
  TopologyConstraintInfo TopologyConstraint;
};
struct Topology {  // pkg/apis/kai/v1alpha1 Topology: Spec.Levels[].NodeLabel, top level first
  std::string Name;
  std::vector<std::string> Levels;
};
struct PodGroupInfo {  // podgroup_info/job_info.go:65-103
  std::string UID, Queue;
  int Priority = 0;
  bool Preemptible = true;
  long long CreationTimestamp = 0;
  std::vector<PodSet> PodSets;  // name order
  TopologyConstraintInfo RootTopologyConstraint;  // constraint of the RootSubGroupSet
  std::vector<SubGroupSet> SubGroupSets;          // nested sets below the root (none for most jobs)
  std::vector<std::shared_ptr<PodInfo>> Tasks;
  int SchedulingConstraintsSignature = -1;
  double LastStartTimestamp = -1;  // seconds on the session clock; <= 0 = nil (job_info.go:185-193)
  double StaleTimeStamp = -1;      // StalenessInfo.TimeStamp, same clock; <= 0 = nil (job_info.go:174-182)
};
struct NodeInfo {  // node_info/node_info.go:68-105
  std::string Name;
  ResourceVector Allocatable, Idle, Releasing;
  bool Ready = true, NotCpuOnly = false;
  std::map<std::string, std::string> Labels;
  std::map<std::string, std::shared_ptr<PodInfo>> PodInfos;
  // node_info.go:457-493 addTaskResources
  void AddTask(const std::shared_ptr<PodInfo> &t) {
    PodInfos[t->UID] = t;
    for (size_t r = 0; r < Idle.size(); r++) {
      if (t->Status == pod_status::Releasing) {
        Releasing[r] += t->ResReq[r];
        Idle[r] -= t->ResReq[r];
      } else if (t->Status == pod_status::Pipelined) {
        Releasing[r] -= t->ResReq[r];
      } else {
        Idle[r] -= t->ResReq[r];
      }
    }
  }
};
struct QueueInfo {  // queue_info/queue_info.go:32-43; quota / limit / over-quota weight per (cpu, memory, gpu)
  std::string UID, ParentQueue;
  int Priority = 100;
  long long CreationTimestamp = 0;
  double Deserved[3] = {-1, -1, -1}, Limit[3] = {-1, -1, -1}, OverQuotaWeight[3] = {1, 1, 1};
  double PreemptMinRuntime = -1, ReclaimMinRuntime = -1;  // seconds; < 0 = nil (*metav1.Duration)
};
struct ClusterInfo {  // cluster_info.go:43-64
  std::map<std::string, std::shared_ptr<NodeInfo>> Nodes;
  std::map<std::string, std::shared_ptr<PodGroupInfo>> PodGroupInfos;
  std::map<std::string, std::shared_ptr<QueueInfo>> Queues;
  std::vector<Topology> Topologies;
};
}  // namespace api

namespace framework {
enum ActionType { Allocate, Consolidation, Reclaim, Preempt, StaleGangEviction };
inline const char *ActionName(ActionType t) {
  static const char *n[] = {"allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"};
  return n[t];
}

struct Cache {  // the side effects the reference's cache sees (cache/cache.go Bind / Evict / TaskPipelined)
  int Binds = 0, Evictions = 0, Pipelines = 0;
};

class Session;
class Statement {  // framework/statement.go
 public:
  explicit Statement(Session *s) : ssn(s) {}
  void Allocate(const std::shared_ptr<api::PodInfo> &t, const std::string &host) { ops.push_back({0, t, host}); }
  void Pipeline(const std::shared_ptr<api::PodInfo> &t, const std::string &host) { ops.push_back({1, t, host}); }
  void Evict(const std::shared_ptr<api::PodInfo> &t) { ops.push_back({2, t, t->NodeName}); }
  void Commit();

 private:
  struct Op {
    int kind;
    std::shared_ptr<api::PodInfo> task;
    std::string host;
  };
  Session *ssn;
  std::vector<Op> ops;
};

class Session {  // framework/session.go:43-98
 public:
  api::ClusterInfo ClusterInfo;
  double Now = 0;  // the instant min-runtime windows are measured against (time.Now() in plugins/minruntime)
  Cache cache;
  kai_config Config{};
  Session() {
    Config.abi_version = KAI_ABI_VERSION;
    Config.k_value = 1.0;
    Config.saturation_multiplier = 1.0;
    Config.max_consolidation_preemptees = -1;
    Config.allow_consolidating_reclaim = 1;
    Config.shard_count = 1;
  }
  ~Session() {
    if (engine) kai_engine_destroy(engine);
  }
  Statement NewStatement() { return Statement(this); }
  const std::string &LastError() const { return last_error; }

  // engine state of this session: the snapshot is loaded by the first engine Action of the cycle
  kai_engine *engine = nullptr;
  bool snapshot_loaded = false;
  std::string last_error;
  // index maps of the packed snapshot (snapshot index -> session object)
  std::vector<std::shared_ptr<api::NodeInfo>> idx_nodes;
  std::vector<std::shared_ptr<api::PodGroupInfo>> idx_jobs;
  std::vector<std::shared_ptr<api::PodInfo>> idx_tasks;
  std::vector<int> task_job;
};

inline void Statement::Commit() {  // :536-571 with commitAllocate -> BindPod (session.go:111-125)
  for (auto &op : ops) {
    auto &t = op.task;
    auto &nodes = ssn->ClusterInfo.Nodes;
    switch (op.kind) {
      case 0:
        t->Status = pod_status::Binding;
        t->NodeName = op.host;
        ssn->cache.Binds++;
        break;
      case 1:
        t->Status = pod_status::Pipelined;
        t->NodeName = op.host;
        ssn->cache.Pipelines++;
        break;
      case 2:
        t->Status = pod_status::Releasing;
        ssn->cache.Evictions++;
        break;
    }
    (void)nodes;
  }
  ops.clear();
}

class Action {  // framework/interface.go:41-47
 public:
  virtual ~Action() = default;
  virtual ActionType Name() const = 0;
  virtual void Execute(Session &ssn) = 0;
};
inline std::map<std::string, std::shared_ptr<Action>> &actionMap() {
  static std::map<std::string, std::shared_ptr<Action>> m;
  return m;
}
inline void RegisterAction(std::shared_ptr<Action> a) { actionMap()[ActionName(a->Name())] = std::move(a); }  // plugins.go:47-62
inline std::shared_ptr<Action> GetAction(const std::string &name) {  // conf_util: "failed to find Action <name>"
  auto it = actionMap().find(name);
  return it == actionMap().end() ? nullptr : it->second;
}
}  // namespace framework

namespace gpuengine {
using namespace framework;

// ---- pack ssn.ClusterInfo into kai_snapshot ----
struct Packed {
  kai_snapshot c{};
  std::vector<double> alloc, idle, rel, qd, ql, qw, treq;
  std::vector<int32_t> name_rank, qparent, qprio, quid, jqueue, jprio, jorder, jpsb, psmin, pstb, tstatus, tnode, torder, tnom, jsig;
  std::vector<int32_t> level_begin, node_domain, jsgs, sgs_parent, sgs_rank, sgs_topo, sgs_req, sgs_pref, ps_sgs, ps_topo, ps_req, ps_pref;
  std::vector<uint32_t> nflags, jflags;
  std::vector<int64_t> qcreation;
  std::vector<double> q_preempt_mrt, q_reclaim_mrt, j_last_start, j_stale_since;
  std::vector<int32_t> tpred;
  std::vector<uint32_t> pred_mask;
};
template <class T, class K>
std::vector<int32_t> rank_of(const std::vector<T> &items, K key) {  // rank of every item under the key's `<`
  std::vector<int32_t> order(items.size()), rank(items.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key(items[a]) < key(items[b]); });
  for (size_t i = 0; i < order.size(); i++) rank[order[i]] = (int32_t)i;
  return rank;
}
inline void packSnapshot(Session &ssn, Packed &p) {
  auto &ci = ssn.ClusterInfo;
  ssn.idx_nodes.clear();
  ssn.idx_jobs.clear();
  ssn.idx_tasks.clear();
  p.q_preempt_mrt.clear();
  p.q_reclaim_mrt.clear();
  p.j_last_start.clear();
  p.j_stale_since.clear();
  p.tpred.clear();
  p.pred_mask.clear();
  std::map<std::vector<std::string>, int> pred_classes;
  ssn.task_job.clear();
  for (auto &kv : ci.Nodes) ssn.idx_nodes.push_back(kv.second);  // std::map: already in byte-wise name order
  const int N = (int)ssn.idx_nodes.size();
  const int R = N ? (int)ssn.idx_nodes[0]->Allocatable.size() : 4;
  p.alloc.assign((size_t)R * N, 0);
  p.idle.assign((size_t)R * N, 0);
  p.rel.assign((size_t)R * N, 0);
  p.name_rank.resize(N);
  p.nflags.resize(N);
  std::map<std::string, int> node_index;
  for (int n = 0; n < N; n++) {
    auto &nd = *ssn.idx_nodes[n];
    node_index[nd.Name] = n;
    p.name_rank[n] = n;  // session.go:480-485: Go string order == std::map order
    p.nflags[n] = (nd.Ready ? KAI_NODE_READY : 0) | (nd.NotCpuOnly ? KAI_NODE_NOT_CPU_ONLY : 0);
    for (int r = 0; r < R; r++) {
      p.alloc[(size_t)r * N + n] = nd.Allocatable[r];
      p.idle[(size_t)r * N + n] = nd.Idle[r];
      p.rel[(size_t)r * N + n] = nd.Releasing[r];
    }
  }
  // queues: leaves and inner queues alike, any order; parents by index
  std::vector<std::shared_ptr<api::QueueInfo>> queues;
  std::map<std::string, int> queue_index;
  for (auto &kv : ci.Queues) {
    queue_index[kv.first] = (int)queues.size();
    queues.push_back(kv.second);
  }
  const int Q = (int)queues.size();
  p.qparent.resize(Q);
  p.qprio.resize(Q);
  p.qcreation.resize(Q);
  p.qd.assign((size_t)3 * Q, 0);
  p.ql.assign((size_t)3 * Q, 0);
  p.qw.assign((size_t)3 * Q, 0);
  p.quid = rank_of(queues, [](const std::shared_ptr<api::QueueInfo> &q) { return q->UID; });
  for (int q = 0; q < Q; q++) {
    auto &qi = *queues[q];
    p.qparent[q] = qi.ParentQueue.empty() ? -1 : queue_index.at(qi.ParentQueue);
    p.qprio[q] = qi.Priority;
    p.qcreation[q] = qi.CreationTimestamp;
    p.q_preempt_mrt.push_back(qi.PreemptMinRuntime);
    p.q_reclaim_mrt.push_back(qi.ReclaimMinRuntime);
    for (int r = 0; r < 3; r++) {
      p.qd[(size_t)r * Q + q] = qi.Deserved[r];
      p.ql[(size_t)r * Q + q] = qi.Limit[r];
      p.qw[(size_t)r * Q + q] = qi.OverQuotaWeight[r];
    }
  }
  // jobs / podsets / tasks
  for (auto &kv : ci.PodGroupInfos) ssn.idx_jobs.push_back(kv.second);
  const int J = (int)ssn.idx_jobs.size();
  p.jorder = rank_of(ssn.idx_jobs, [](const std::shared_ptr<api::PodGroupInfo> &j) {
    return std::make_pair(j->CreationTimestamp, j->UID);  // session_plugins.go:235-241
  });
  p.jqueue.resize(J);
  p.jprio.resize(J);
  p.jflags.resize(J);
  p.jsig.resize(J);
  p.jpsb.assign(1, 0);
  p.pstb.assign(1, 0);
  for (int j = 0; j < J; j++) {
    auto &job = *ssn.idx_jobs[j];
    auto qi = queue_index.find(job.Queue);
    p.jqueue[j] = qi == queue_index.end() ? -1 : qi->second;
    p.jprio[j] = job.Priority;
    p.jflags[j] = job.Preemptible ? KAI_JOB_PREEMPTIBLE : 0;
    p.jsig[j] = job.SchedulingConstraintsSignature;
    p.j_last_start.push_back(job.LastStartTimestamp);
    p.j_stale_since.push_back(job.StaleTimeStamp);
    auto order = rank_of(job.Tasks, [](const std::shared_ptr<api::PodInfo> &t) { return std::make_pair(t->OrderKey, t->UID); });
    for (auto &ps : job.PodSets) {
      p.psmin.push_back(ps.MinAvailable);
      for (size_t k = 0; k < job.Tasks.size(); k++) {
        auto &t = job.Tasks[k];
        std::string sg = t->SubGroupName.empty() ? job.PodSets[0].Name : t->SubGroupName;
        if (sg != ps.Name) continue;
        ssn.idx_tasks.push_back(t);
        ssn.task_job.push_back(j);
        p.tstatus.push_back(t->Status);
        auto ni = node_index.find(t->NodeName);
        p.tnode.push_back(pod_status::IsActiveUsedStatus(t->Status) && ni != node_index.end() ? ni->second : -1);
        p.torder.push_back(order[k]);
        auto nom = node_index.find(t->NominatedNodeName);
        p.tnom.push_back(nom == node_index.end() ? -1 : nom->second);
        if (t->NodeAffinityNames.empty()) {
          p.tpred.push_back(-1);
        } else {  // one predicate class per distinct constraint; bit n = node n passes the filter
          auto it = pred_classes.find(t->NodeAffinityNames);
          if (it == pred_classes.end()) {
            const int words = (N + 31) / 32;
            p.pred_mask.resize(p.pred_mask.size() + words, 0u);
            uint32_t *row = p.pred_mask.data() + p.pred_mask.size() - words;
            for (int n = 0; n < N; n++) {
              auto &labels = ssn.idx_nodes[n]->Labels;
              auto lt = labels.find("kai.scheduler/type");
              const std::string &type = lt == labels.end() ? ssn.idx_nodes[n]->Name : lt->second;
              if (std::find(t->NodeAffinityNames.begin(), t->NodeAffinityNames.end(), type) != t->NodeAffinityNames.end())
                row[n >> 5] |= 1u << (n & 31);
            }
            it = pred_classes.emplace(t->NodeAffinityNames, (int)pred_classes.size()).first;
          }
          p.tpred.push_back(it->second);
        }
        for (int r = 0; r < R; r++) p.treq.push_back(t->ResReq[r]);
      }
      p.pstb.push_back((int32_t)p.tstatus.size());
    }
    p.jpsb.push_back((int32_t)p.psmin.size());
  }
  // Topology CRs -> per-level dense domain ids in DomainID order (plugins/topology/topology_structs.go:76-82)
  p.level_begin.assign(1, 0);
  std::vector<std::vector<std::string>> level_labels;
  for (auto &tp : ci.Topologies) {
    level_labels.push_back(tp.Levels);
    p.level_begin.push_back(p.level_begin.back() + (int32_t)tp.Levels.size());
  }
  p.node_domain.assign((size_t)p.level_begin.back() * N, -1);
  for (size_t k = 0; k < ci.Topologies.size(); k++)
    for (size_t li = 0; li < level_labels[k].size(); li++) {
      std::vector<std::string> ids(N);
      std::vector<char> has(N, 1);
      for (int n = 0; n < N; n++) {
        std::string id;
        for (size_t l2 = 0; l2 <= li; l2++) {
          auto it = ssn.idx_nodes[n]->Labels.find(level_labels[k][l2]);
          if (it == ssn.idx_nodes[n]->Labels.end()) {
            has[n] = 0;
            break;
          }
          id += (l2 ? "." : "") + it->second;
        }
        ids[n] = id;
      }
      std::vector<std::string> uniq;
      for (int n = 0; n < N; n++)
        if (has[n]) uniq.push_back(ids[n]);
      std::sort(uniq.begin(), uniq.end());
      uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
      for (int n = 0; n < N; n++)
        if (has[n])
          p.node_domain[(size_t)(p.level_begin[k] + li) * N + n] =
              (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), ids[n]) - uniq.begin());
    }
  auto constraint = [&](const api::TopologyConstraintInfo &tc, int32_t &topo, int32_t &req, int32_t &pref) {
    topo = req = pref = -1;
    if (tc.Topology.empty()) return;
    topo = -2;
    for (size_t k = 0; k < ci.Topologies.size(); k++)
      if (ci.Topologies[k].Name == tc.Topology) topo = (int32_t)k;
    if (topo < 0) return;
    auto level = [&](const std::string &name) -> int32_t {
      if (name.empty()) return -1;
      for (size_t l = 0; l < level_labels[topo].size(); l++)
        if (level_labels[topo][l] == name) return (int32_t)l;
      return -2;
    };
    req = level(tc.RequiredLevel);
    pref = level(tc.PreferredLevel);
  };
  // SubGroupSet tree: root first, then the job's nested sets; PodSets point at their set
  p.jsgs.assign(1, 0);
  {
    int ps_index = 0;
    for (int j = 0; j < J; j++) {
      auto &job = *ssn.idx_jobs[j];
      const int root = (int)p.sgs_parent.size();
      std::map<std::string, int> set_index;
      auto add_set = [&](const std::string &name, int parent, const api::TopologyConstraintInfo &tc) {
        int32_t t, r, q;
        constraint(tc, t, r, q);
        set_index[name] = (int)p.sgs_parent.size();
        p.sgs_parent.push_back(parent);
        p.sgs_topo.push_back(t);
        p.sgs_req.push_back(r);
        p.sgs_pref.push_back(q);
      };
      add_set("", -1, job.RootTopologyConstraint);
      for (auto &g : job.SubGroupSets) add_set(g.Name, g.Parent.empty() ? root : set_index.at(g.Parent), g.TopologyConstraint);
      std::vector<std::string> names;  // SubGroupSetOrderFn falls back to the name; the root is alone at its level
      for (auto &g : job.SubGroupSets) names.push_back(g.Name);
      std::sort(names.begin(), names.end());
      p.sgs_rank.push_back(0);
      for (auto &g : job.SubGroupSets)
        p.sgs_rank.push_back((int32_t)(std::lower_bound(names.begin(), names.end(), g.Name) - names.begin()));
      for (auto &ps : job.PodSets) {
        int32_t t, r, q;
        constraint(ps.TopologyConstraint, t, r, q);
        p.ps_sgs.push_back(ps.ParentSet.empty() ? root : set_index.at(ps.ParentSet));
        p.ps_topo.push_back(t);
        p.ps_req.push_back(r);
        p.ps_pref.push_back(q);
        ps_index++;
      }
      p.jsgs.push_back((int32_t)p.sgs_parent.size());
    }
  }
  kai_snapshot &c = p.c;
  c.abi_version = KAI_ABI_VERSION;
  c.n_res = R;
  c.n_nodes = N;
  c.n_queues = Q;
  c.n_jobs = J;
  c.n_podsets = (int32_t)p.psmin.size();
  c.n_tasks = (int32_t)p.tstatus.size();
  c.node_allocatable = p.alloc.data();
  c.node_idle = p.idle.data();
  c.node_releasing = p.rel.data();
  c.node_name_rank = p.name_rank.data();
  c.node_flags = p.nflags.data();
  c.queue_parent = p.qparent.data();
  c.queue_priority = p.qprio.data();
  c.queue_creation = p.qcreation.data();
  c.queue_uid_rank = p.quid.data();
  c.queue_deserved = p.qd.data();
  c.queue_limit = p.ql.data();
  c.queue_oqw = p.qw.data();
  c.job_queue = p.jqueue.data();
  c.job_priority = p.jprio.data();
  c.job_order_rank = p.jorder.data();
  c.job_flags = p.jflags.data();
  c.job_podset_begin = p.jpsb.data();
  c.podset_min_available = p.psmin.data();
  c.podset_task_begin = p.pstb.data();
  c.task_status = p.tstatus.data();
  c.task_node = p.tnode.data();
  c.task_req = p.treq.data();
  c.task_order_rank = p.torder.data();
  c.task_nominated = p.tnom.data();
  c.n_pred_classes = (int32_t)pred_classes.size();
  if (!pred_classes.empty()) {
    c.task_pred_class = p.tpred.data();
    c.pred_mask = p.pred_mask.data();
  }
  c.job_signature = p.jsig.data();
  c.now_s = ssn.Now;
  c.queue_preempt_min_runtime_s = p.q_preempt_mrt.data();
  c.queue_reclaim_min_runtime_s = p.q_reclaim_mrt.data();
  c.job_last_start_s = p.j_last_start.data();
  c.job_stale_since_s = p.j_stale_since.data();
  c.n_topologies = (int32_t)ci.Topologies.size();
  c.topology_level_begin = p.level_begin.data();
  c.node_domain = p.node_domain.data();
  c.n_subgroup_sets = (int32_t)p.sgs_parent.size();
  c.job_sgs_begin = p.jsgs.data();
  c.sgs_parent = p.sgs_parent.data();
  c.sgs_name_rank = p.sgs_rank.data();
  c.sgs_topology = p.sgs_topo.data();
  c.sgs_required_level = p.sgs_req.data();
  c.sgs_preferred_level = p.sgs_pref.data();
  c.podset_sgs = p.ps_sgs.data();
  c.podset_topology = p.ps_topo.data();
  c.podset_required_level = p.ps_req.data();
  c.podset_preferred_level = p.ps_pref.data();
}

class gpuAction : public Action {
 public:
  explicit gpuAction(ActionType t) : type(t) {}
  ActionType Name() const override { return type; }
  void Execute(Session &ssn) override {
    static const kai_action ids[] = {KAI_ACTION_ALLOCATE, KAI_ACTION_CONSOLIDATION, KAI_ACTION_RECLAIM, KAI_ACTION_PREEMPT,
                                     KAI_ACTION_STALEGANGEVICTION};
    ssn.last_error.clear();
    if (!ssn.engine && kai_engine_create(&ssn.Config, &ssn.engine) != KAI_OK) {
      ssn.last_error = "kai_engine_create failed (no CUDA device?)";  // Go shim: run the stock action instead
      ssn.engine = nullptr;
      return;
    }
    if (!ssn.snapshot_loaded) {  // first engine Action of the cycle: OpenSession hand-off
      Packed p;
      packSnapshot(ssn, p);
      if (kai_engine_load_snapshot(ssn.engine, &p.c) != KAI_OK) {
        ssn.last_error = kai_last_error(ssn.engine);
        return;
      }
      ssn.snapshot_loaded = true;
    }
    kai_result res{};
    if (kai_engine_run(ssn.engine, ids[type], &res) != KAI_OK) {
      ssn.last_error = kai_last_error(ssn.engine);
      return;
    }
    // Replay.  allocate: one Statement per committed visit, that job's newly placed
    // tasks in task order (allocate.go:63-72).  Solver actions and stalegangeviction: the run's victims are
    // evicted first (common.EvictAllPreemptees), then the pipelined pods, in one Statement per run — the result
    // carries final statuses, and the session state after Commit is what the stock actions leave.
    auto node_name = [&](int t) { return res.task_node[t] >= 0 ? ssn.idx_nodes[res.task_node[t]]->Name : std::string(); };
    auto changed = [&](int t) { return ssn.idx_tasks[t]->Status != res.task_status[t] || ssn.idx_tasks[t]->NodeName != node_name(t); };
    if (type == Allocate) {
      std::vector<char> done(res.n_tasks, 0);
      for (int v = 0; v < res.n_visits; v++) {
        if (!res.visits[v].outcome) continue;
        Statement stmt = ssn.NewStatement();
        for (int t = 0; t < res.n_tasks; t++) {
          if (done[t] || ssn.task_job[t] != res.visits[v].job || !changed(t)) continue;
          if (res.task_status[t] == pod_status::Binding)
            stmt.Allocate(ssn.idx_tasks[t], node_name(t));
          else if (res.task_status[t] == pod_status::Pipelined)
            stmt.Pipeline(ssn.idx_tasks[t], node_name(t));
          done[t] = 1;
        }
        stmt.Commit();
      }
    } else {
      Statement stmt = ssn.NewStatement();
      for (int t = 0; t < res.n_tasks; t++)
        if (changed(t) && res.task_status[t] == pod_status::Releasing) stmt.Evict(ssn.idx_tasks[t]);
      for (int t = 0; t < res.n_tasks; t++)
        if (changed(t) && res.task_status[t] == pod_status::Pipelined) stmt.Pipeline(ssn.idx_tasks[t], node_name(t));
      stmt.Commit();
    }
  }

 private:
  ActionType type;
};
inline std::shared_ptr<Action> New(ActionType t) { return std::make_shared<gpuAction>(t); }
// cmd/scheduler/app/server.go:167: call after actions.InitDefaultActions()
inline void RegisterAll() {
  for (ActionType t : {Allocate, Consolidation, Reclaim, Preempt, StaleGangEviction}) RegisterAction(New(t));
}
}  // namespace gpuengine

}  // namespace kai_host
