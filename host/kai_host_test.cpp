// kai_host_test — drives the C++ mirror of the reference's Action surface the way the reference's own tests do
// (pkg/scheduler/test_utils/test_utils.go:40-119 BuildSession + RunTests): reads a cluster description, builds a
// framework::Session, resolves the Actions from the registry by name, executes them in order and prints the tasks.
//
// input (text, one record per line; written by tests/test_host_cpp.py from the transcribed reference tables):
//   R <n_res>
//   queue <uid> <parent|-> <priority> <creation> <d0 d1 d2> <l0 l1 l2> <w0 w1 w2>
//   node <name> <allocatable x R>
//   job <uid> <queue> <priority> <preemptible 0/1> <creation>
//   podset <job> <name> <minAvailable> [<parent set|-> <topology|-> <required|-> <preferred|->]
//   set <job> <name> <parent set|-> <topology|-> <required|-> <preferred|->     (nested SubGroupSets, parents first)
//   rootconstraint <job> <topology> <required|-> <preferred|->
//   topology <name> <level label> ...        label <node> <key> <value>
//   task <job> <podset> <uid> <status> <node|-> <order key> <req x R>
//   actions <name> ...
// output: one line per task "<uid> <status> <node|->", then "cache <binds> <evictions> <pipelines>".
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "kai_host.hpp"

using namespace kai_host;

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <case file> | --registry\n", argv[0]);
    return 2;
  }
  gpuengine::RegisterAll();
  if (std::string(argv[1]) == "--registry") {  // no GPU needed: the registry resolves every default action name
    for (const char *n : {"allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"})
      printf("%s %s\n", n, framework::GetAction(n) ? "registered" : "missing");
    printf("unknown %s\n", framework::GetAction("unknown") ? "registered" : "missing");
    return 0;
  }
  std::ifstream in(argv[1]);
  framework::Session ssn;
  std::vector<std::string> actions;
  int R = 4;
  std::string line;
  std::vector<std::shared_ptr<api::PodInfo>> all_tasks;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string kind;
    ls >> kind;
    if (kind == "R") {
      ls >> R;
    } else if (kind == "queue") {
      auto q = std::make_shared<api::QueueInfo>();
      std::string parent;
      ls >> q->UID >> parent >> q->Priority >> q->CreationTimestamp;
      q->ParentQueue = parent == "-" ? "" : parent;
      for (double &v : q->Deserved) ls >> v;
      for (double &v : q->Limit) ls >> v;
      for (double &v : q->OverQuotaWeight) ls >> v;
      ssn.ClusterInfo.Queues[q->UID] = q;
    } else if (kind == "node") {
      auto n = std::make_shared<api::NodeInfo>();
      ls >> n->Name;
      n->Allocatable.resize(R);
      for (double &v : n->Allocatable) ls >> v;
      n->Idle = n->Allocatable;
      n->Releasing.assign(R, 0.0);
      ssn.ClusterInfo.Nodes[n->Name] = n;
    } else if (kind == "job") {
      auto j = std::make_shared<api::PodGroupInfo>();
      int pre;
      ls >> j->UID >> j->Queue >> j->Priority >> pre >> j->CreationTimestamp;
      j->Preemptible = pre != 0;
      ssn.ClusterInfo.PodGroupInfos[j->UID] = j;
    } else if (kind == "podset") {
      std::string job, parent, topo, req, pref;
      api::PodSet ps;
      ls >> job >> ps.Name >> ps.MinAvailable;
      if (ls >> parent >> topo >> req >> pref) {
        ps.ParentSet = parent == "-" ? "" : parent;
        if (topo != "-") ps.TopologyConstraint = {topo, req == "-" ? "" : req, pref == "-" ? "" : pref};
      }
      ssn.ClusterInfo.PodGroupInfos[job]->PodSets.push_back(ps);
    } else if (kind == "set") {
      std::string job, parent, topo, req, pref;
      api::SubGroupSet g;
      ls >> job >> g.Name >> parent >> topo >> req >> pref;
      g.Parent = parent == "-" ? "" : parent;
      if (topo != "-") g.TopologyConstraint = {topo, req == "-" ? "" : req, pref == "-" ? "" : pref};
      ssn.ClusterInfo.PodGroupInfos[job]->SubGroupSets.push_back(g);
    } else if (kind == "rootconstraint") {
      std::string job, topo, req, pref;
      ls >> job >> topo >> req >> pref;
      ssn.ClusterInfo.PodGroupInfos[job]->RootTopologyConstraint = {topo, req == "-" ? "" : req, pref == "-" ? "" : pref};
    } else if (kind == "topology") {
      api::Topology tp;
      ls >> tp.Name;
      std::string lv;
      while (ls >> lv) tp.Levels.push_back(lv);
      ssn.ClusterInfo.Topologies.push_back(tp);
    } else if (kind == "label") {
      std::string node, key, value;
      ls >> node >> key >> value;
      ssn.ClusterInfo.Nodes[node]->Labels[key] = value;
    } else if (kind == "task") {
      auto t = std::make_shared<api::PodInfo>();
      std::string node;
      ls >> t->Job >> t->SubGroupName >> t->UID >> t->Status >> node >> t->OrderKey;
      t->ResReq.resize(R);
      for (double &v : t->ResReq) ls >> v;
      t->NodeName = node == "-" ? "" : node;
      ssn.ClusterInfo.PodGroupInfos[t->Job]->Tasks.push_back(t);
      all_tasks.push_back(t);
      // nodes_fake/nodes.go:213-223: tasks in an active-used state are added to their node
      if (pod_status::IsActiveUsedStatus(t->Status) && ssn.ClusterInfo.Nodes.count(t->NodeName))
        ssn.ClusterInfo.Nodes[t->NodeName]->AddTask(t);
    } else if (kind == "affinity") {  // affinity <pod> <name>...
      std::string pod, name;
      ls >> pod;
      for (auto &t : all_tasks)
        if (t->UID == pod)
          while (ls >> name) t->NodeAffinityNames.push_back(name);
    } else if (kind == "conf") {  // conf <gpu placement> <cpu placement> <default reclaim s> <default preempt s> <method> <now>
      ls >> ssn.Config.gpu_placement >> ssn.Config.cpu_placement >> ssn.Config.default_reclaim_min_runtime_s >>
          ssn.Config.default_preempt_min_runtime_s >> ssn.Config.reclaim_resolve_method >> ssn.Now;
    } else if (kind == "queuemrt") {  // queuemrt <queue> <preempt s | -1> <reclaim s | -1>
      std::string q;
      ls >> q;
      ls >> ssn.ClusterInfo.Queues[q]->PreemptMinRuntime >> ssn.ClusterInfo.Queues[q]->ReclaimMinRuntime;
    } else if (kind == "jobstart") {  // jobstart <job> <last start s | -1>
      std::string j;
      ls >> j;
      ls >> ssn.ClusterInfo.PodGroupInfos[j]->LastStartTimestamp;
    } else if (kind == "jobstale") {  // jobstale <job> <stale since s | -1>
      std::string j;
      ls >> j;
      ls >> ssn.ClusterInfo.PodGroupInfos[j]->StaleTimeStamp;
    } else if (kind == "grace") {  // grace <GlobalDefaultStalenessGracePeriod s | -1>
      ls >> ssn.Config.staleness_grace_period_s;
    } else if (kind == "actions") {
      std::string a;
      while (ls >> a) actions.push_back(a);
    }
  }
  if (argc > 2 && std::string(argv[2]) == "--dump-packed") {  // no GPU: print what packSnapshot hands to the C ABI
    gpuengine::Packed p;
    gpuengine::packSnapshot(ssn, p);
    const kai_snapshot &c = p.c;
    const int L = c.n_topologies ? c.topology_level_begin[c.n_topologies] : 0;
    for (int l = 0; l < L; l++)
      for (int n = 0; n < c.n_nodes; n++)
        printf("node_domain %d %s %d\n", l, ssn.idx_nodes[n]->Name.c_str(), c.node_domain[(size_t)l * c.n_nodes + n]);
    for (int j = 0; j < c.n_jobs; j++) {
      for (int g = c.job_sgs_begin[j]; g < c.job_sgs_begin[j + 1]; g++)
        printf("set %s %d parent %d rank %d con %d %d %d\n", ssn.idx_jobs[j]->UID.c_str(), g - c.job_sgs_begin[j],
               c.sgs_parent[g] < 0 ? -1 : c.sgs_parent[g] - c.job_sgs_begin[j], c.sgs_name_rank[g], c.sgs_topology[g],
               c.sgs_required_level[g], c.sgs_preferred_level[g]);
      for (int ps = c.job_podset_begin[j]; ps < c.job_podset_begin[j + 1]; ps++)
        printf("podset %s %d min %d set %d con %d %d %d\n", ssn.idx_jobs[j]->UID.c_str(), ps - c.job_podset_begin[j],
               c.podset_min_available[ps], c.podset_sgs[ps] - c.job_sgs_begin[j], c.podset_topology[ps],
               c.podset_required_level[ps], c.podset_preferred_level[ps]);
    }
    for (int j = 0; j < c.n_jobs; j++)
      if (c.job_stale_since_s && c.job_stale_since_s[j] > 0)
        printf("stale %s %.17g now %.17g grace %d\n", ssn.idx_jobs[j]->UID.c_str(), c.job_stale_since_s[j], c.now_s,
               ssn.Config.staleness_grace_period_s);
    for (int t = 0; t < c.n_tasks; t++) {
      if (!c.task_pred_class || c.task_pred_class[t] < 0) continue;
      printf("pred %s %d", ssn.idx_tasks[t]->UID.c_str(), c.task_pred_class[t]);
      const int words = (c.n_nodes + 31) / 32;
      for (int w = 0; w < words; w++) printf(" %u", c.pred_mask[(size_t)c.task_pred_class[t] * words + w]);
      printf("\n");
    }
    return 0;
  }
  for (const std::string &name : actions) {
    auto action = framework::GetAction(name);  // scheduler.go:129-136 runOnce: for _, action := range actions
    if (!action) {
      fprintf(stderr, "failed to find Action %s\n", name.c_str());
      return 3;
    }
    action->Execute(ssn);
    if (!ssn.LastError().empty()) {
      fprintf(stderr, "engine error in %s: %s\n", name.c_str(), ssn.LastError().c_str());
      return 4;
    }
  }
  for (auto &t : all_tasks) printf("%s %d %s\n", t->UID.c_str(), t->Status, t->NodeName.empty() ? "-" : t->NodeName.c_str());
  printf("cache %d %d %d\n", ssn.cache.Binds, ssn.cache.Evictions, ssn.cache.Pipelines);
  return 0;
}
