// Host-only check of kai_topology.cuh: the incremental per-domain state (free sums, allocatable-pod classes, ratio
// memo) fed by node deltas must give the same subSetNodesFn answers as a state rebuilt from the node tables.
// Built and run by tests/test_topology_host.py (nvcc, no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../kai_scheduler_b200/csrc/kai_topology.cuh"

using namespace kai;
namespace kai {
void seq_flush_deltas(Seq &) {}
}

static unsigned long long rng_state = 0x0C41ULL;
static unsigned int rnd() {
  rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (unsigned int)(rng_state >> 33);
}

int main() {
  const int N = 600, R = 4, J = 1;
  std::vector<int> nd(3 * (size_t)N);
  for (int n = 0; n < N; n++) {
    int rack = n / 5, leaf = rack / 4, spine = leaf / 3;
    nd[n] = spine;
    nd[N + n] = leaf;
    nd[2 * N + n] = rack;
  }
  std::vector<int> lb = {0, 3}, jt(J, 0), jr(J, 1), jp(J, 2), jpb = {0, 1};
  kai_snapshot s{};
  s.n_nodes = N;
  s.n_res = R;
  s.n_topologies = 1;
  s.topology_level_begin = lb.data();
  s.node_domain = nd.data();
  s.n_jobs = J;
  s.n_podsets = 1;
  s.job_podset_begin = jpb.data();
  s.job_topology = jt.data();
  s.job_required_level = jr.data();
  s.job_preferred_level = jp.data();
  std::vector<double> mirror((size_t)N * 2 * R, 0.0);
  for (int n = 0; n < N; n++) {
    mirror[(size_t)n * 2 * R + 0] = 2e7;
    mirror[(size_t)n * 2 * R + 1] = 2e10;
    mirror[(size_t)n * 2 * R + 2] = 8;
    mirror[(size_t)n * 2 * R + 3] = 110;
  }
  const int T = 12;
  std::vector<double> req((size_t)T * R);
  std::vector<int> tps(T, 0);
  for (int i = 0; i < T; i++) {
    req[i * R] = 1000;
    req[i * R + 1] = 1e9;
    req[i * R + 2] = (i % 3 == 0) ? 4 : 2;
    req[i * R + 3] = 1;
  }
  TopologyHost live, fresh;
  live.build(&s);
  fresh.build(&s);
  live.mirror = fresh.mirror = mirror.data();
  live.t_req = fresh.t_req = req.data();
  live.t_podset = fresh.t_podset = tps.data();
  int checks = 0;
  for (int step = 0; step < 4000; step++) {
    // a random placement or release of 1..4 GPUs (+ cpu / memory / pod) on a random node, like a node delta
    int n = (int)(rnd() % N);
    double g = 1 + (rnd() % 4), sign = (rnd() & 1) ? -1.0 : 1.0;
    double *row = &mirror[(size_t)n * 2 * R];
    double before[KAI_MAX_RES], after[KAI_MAX_RES];
    double d[4] = {1000 * g, 1e9 * g, g, 1};
    bool ok = true;
    for (int r = 0; r < R; r++)
      if (row[r] + sign * d[r] < 0 || (r == 2 && row[r] + sign * d[r] > 8)) ok = false;
    if (!ok) continue;
    for (int r = 0; r < R; r++) {
      before[r] = row[r] + row[R + r];
      row[r] += sign * d[r];
      after[r] = row[r] + row[R + r];
    }
    live.node_changed(n, before, after);
    if (step % 7) continue;
    std::vector<int> tasks;
    int k = 1 + (int)(rnd() % T);
    for (int i = 0; i < k; i++) tasks.push_back((int)(rnd() % T));
    std::array<int, 3> con = {0, (int)(rnd() % 3), (rnd() & 1) ? 2 : -1};
    fresh.live = false;  // rebuild everything from the tables
    auto a = live.subset(con, tasks, nullptr, [](int) { return true; }, false, {}, true);
    auto b = fresh.subset(con, tasks, nullptr, [](int) { return true; }, false, {}, true);
    if (a.ok != b.ok || a.domains != b.domains || a.scores != b.scores || a.pref_level != b.pref_level) {
      printf("MISMATCH at step %d: %zu vs %zu domains\n", step, a.domains.size(), b.domains.size());
      return 1;
    }
    checks++;
  }
  printf("OK %d checks\n", checks);
  return 0;
}
