// Host-only check of the launch transport's host side (kai_host_seq.cuh): (1) a decision record is packed into the
// LaunchRec a k_record launch carries exactly as the scanners decode it (folded node deltas, repeat counts, extended
// entries), (2) the merged candidate lists of several GPUs are merged with the cut rule applied across ranks.
// Built and run by tests/test_launch_host.py (nvcc, no GPU needed: nothing is launched).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../kai_scheduler_b200/csrc/kai_host_seq.cuh"

using namespace kai;

static unsigned long long rng_state = 0x0C42ULL;
static unsigned int rnd() {
  rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (unsigned int)(rng_state >> 33);
}

static LaunchRec g_last;
static int g_launches = 0;
static bool fake_launch(void *, const LaunchRec &rec) {
  g_last = rec;
  g_launches++;
  return true;
}

#define CHECK(cond, ...)            \
  do {                              \
    if (!(cond)) {                  \
      printf("FAIL %s:%d ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);          \
      printf("\n");                 \
      return 1;                     \
    }                               \
  } while (0)

int main() {
  // ---------------------------------------------------------------- (1) record packing
  const int N = 64, T = 40, R = 4;
  std::vector<int> name_rank(N), rank_to_node(4096);
  for (int n = 0; n < N; n++) name_rank[n] = (n * 37) % N;  // a permutation (37 is odd, N a power of two)
  for (int i = 0; i < 4096; i++) rank_to_node[i] = i;
  std::vector<double> t_req((size_t)T * R);
  for (int t = 0; t < T; t++)
    for (int r = 0; r < R; r++) t_req[(size_t)t * R + r] = (double)(1 + (t / 4) % 3) * (r + 1);  // runs of 4 identical requests
  DevSnap s;
  memset(&s, 0, sizeof(s));
  s.R = R;
  s.N = N;
  s.T = T;
  s.name_rank = name_rank.data();
  s.t_req = t_req.data();
  HostBackend hb;
  std::vector<unsigned long long> delta_buf((size_t)2 * kMaxDelta * 2, 0);
  memset(&hb.ctl, 0, sizeof(hb.ctl));
  memset(&hb.seq, 0, sizeof(hb.seq));
  hb.seq.s = &s;
  hb.seq.ctl = &hb.ctl;
  hb.seq.delta_base = delta_buf.data();
  hb.seq.host_backend = &hb;
  hb.launch_mode = true;
  hb.launch_fn = &fake_launch;
  hb.rank_to_node = rank_to_node.data();
  hb.ctl.seq = 7;
  hb.ctl.dec.nominated = hb.ctl.dec.pred_class = -1;
  // what the scanners must see: (rank | code << 28, first task, repeat count)
  struct Want {
    unsigned int key, task;
    int count;
  };
  std::vector<Want> want;
  auto expect = [&](int node, int code, int t) {
    const unsigned int key = (unsigned int)(name_rank[node] | (code << 28));
    if (!want.empty() && code < ND_FEAS_SET && want.back().key == key && want.back().count < 255 && !(want.back().key & 0x80000000u)) {
      bool same = true;
      for (int r = 0; r < R; r++) same = same && t_req[(size_t)want.back().task * R + r] == t_req[(size_t)t * R + r];
      if (same) {
        want.back().count++;
        return;
      }
    }
    want.push_back({key, (unsigned int)t, 1});
  };
  // four identical pods on node 5 (fold to one entry, count 4), a different request on the same node, another node,
  // a removal, a feasible-set bit, an extended entry, then again node 5
  for (int t = 0; t < 4; t++) {
    emit_delta(hb.seq, 5, ND_ADD, t);
    expect(5, ND_ADD, t);
  }
  emit_delta(hb.seq, 5, ND_ADD, 4);
  expect(5, ND_ADD, 4);
  emit_delta(hb.seq, 9, ND_ADD_PIPELINED, 5);
  expect(9, ND_ADD_PIPELINED, 5);
  emit_delta(hb.seq, 9, ND_REM_PIPELINED, 5);
  expect(9, ND_REM_PIPELINED, 5);
  emit_delta(hb.seq, 11, ND_FEAS_SET, 0);
  want.push_back({(unsigned int)(name_rank[11] | (ND_FEAS_SET << 28)), 0u, 1});
  emit_ext(hb.seq, EXT_SCORE, 17u, 3u);
  want.push_back({0x80000000u | ((unsigned int)EXT_SCORE << 28) | 17u, 3u, 1});
  emit_delta(hb.seq, 5, ND_ADD, 8);
  want.push_back({(unsigned int)(name_rank[5] | (ND_ADD << 28)), 8u, 1});
  emit_delta(hb.seq, 5, ND_ADD, 9);  // same request as task 8: folds
  want.back().count++;
  for (int r = 0; r < KAI_MAX_RES; r++) hb.ctl.dec.req[r] = r < R ? t_req[r] : 0.0;
  hb.ctl.dec.gpu_task = 1;
  hb.ctl.dec.res = KAI_RES_GPU;
  hb.ctl.xbits = XB_SINGLE;
  hb.publish(DK_SCAN);
  CHECK(g_launches == 1, "one launch per record, got %d", g_launches);
  CHECK(g_last.seq == 7u && g_last.n_delta == (int)want.size(), "seq %u n_delta %d (want %zu)", g_last.seq, g_last.n_delta, want.size());
  CHECK((int)(g_last.dw[0] & 0xff) == DK_SCAN && (int)((g_last.dw[0] >> 32) & 0xffff) == (int)want.size(), "record word 0");
  CHECK(((unsigned int)(g_last.dw[0] >> 48) & XB_SINGLE) != 0, "xbits");
  for (size_t e = 0; e < want.size(); e++)
    CHECK(g_last.dkey[e] == want[e].key && g_last.dtask[e] == want[e].task && (int)g_last.dcount[e] == want[e].count - 1,
          "delta %zu: key %08x task %u count-1 %d, want %08x %u %d", e, g_last.dkey[e], g_last.dtask[e], (int)g_last.dcount[e], want[e].key,
          want[e].task, want[e].count - 1);
  // a full list flushes by a launch of its own and the sequence number moves on without waiting
  hb.ctl.seq = 8;
  hb.ctl.n_delta = 0;
  hb.ctl.last_dcount = 0;
  for (int i = 0; i < kMaxDelta + 3; i++) emit_delta(hb.seq, i % N, ND_ADD, (i * 5) % T);  // neighbours differ: no folding
  CHECK(g_launches == 2 && (int)(g_last.dw[0] & 0xff) == DK_FLUSH && g_last.n_delta == kMaxDelta, "flush launch: %d launches, kind %d, n_delta %d",
        g_launches, (int)(g_last.dw[0] & 0xff), g_last.n_delta);
  CHECK(hb.ctl.seq == 9u && hb.ctl.n_delta == 3, "after the flush: seq %u n_delta %d", hb.ctl.seq, hb.ctl.n_delta);

  // ---------------------------------------------------------------- (2) merging the GPUs' lists
  for (int trial = 0; trial < 200; trial++) {
    const int S = 1 + (int)(rnd() % 4);
    std::vector<unsigned long long> clist((size_t)S * 2 * kCListWords, 0);
    hb.h_clist = clist.data();
    hb.n_ranks = S;
    hb.failed = false;
    const unsigned int seq_no = 100 + (unsigned int)trial;
    hb.ctl.seq = seq_no;
    struct Ent {
      double score;
      unsigned int rank;
    };
    std::vector<Ent> all;
    bool have_cut = false;
    Ent cut{0, 0};
    auto before = [](const Ent &a, const Ent &b) { return a.score > b.score || (a.score == b.score && a.rank < b.rank); };
    std::vector<unsigned int> ranks(2048);
    for (unsigned int i = 0; i < 2048; i++) ranks[i] = i;
    for (int i = 2047; i > 0; i--) std::swap(ranks[i], ranks[rnd() % (i + 1)]);
    size_t next_rank = 0;
    for (int r = 0; r < S; r++) {
      const int n = (int)(rnd() % 40);
      std::vector<Ent> mine;
      for (int i = 0; i < n; i++) mine.push_back({(double)(rnd() % 5), ranks[next_rank++]});  // few score levels: many ties
      std::sort(mine.begin(), mine.end(), before);
      const bool more = (rnd() & 1) != 0;
      unsigned long long *cl = clist.data() + ((size_t)r * 2 + (seq_no & 1)) * kCListWords;
      for (int i = 0; i < n; i++) {
        unsigned long long *e = cl + 2 + (size_t)i * kCEntryWords;
        memcpy(&e[0], &mine[i].score, 8);
        e[1] = (unsigned long long)mine[i].rank | (2ull << 24) | ((unsigned long long)LF_TO_IDLE << 32);  // cap 3
        double v = 1.0 + i;
        for (int q = 0; q < 4; q++) memcpy(&e[2 + q], &v, 8);
      }
      cl[0] = (unsigned long long)(unsigned int)n | (more ? (1ull << 31) : 0ull);
      cl[1] = seq_no;
      for (auto &m : mine) all.push_back(m);
      if (more && n > 0 && (!have_cut || before(mine.back(), cut))) {
        have_cut = true;
        cut = mine.back();
      }
    }
    std::sort(all.begin(), all.end(), before);
    size_t valid = all.size();
    if (have_cut)
      for (size_t i = 0; i < all.size(); i++)
        if (before(cut, all[i])) {  // strictly worse than the cut row: an unseen row could sit before it
          valid = i;
          break;
        }
    hb.gather_list();
    CHECK(!hb.failed, "trial %d: wait failed", trial);
    CHECK(hb.list.size() == all.size(), "trial %d: %zu entries, want %zu", trial, hb.list.size(), all.size());
    CHECK(hb.list_valid == valid, "trial %d (S=%d): valid prefix %zu, want %zu", trial, S, hb.list_valid, valid);
    CHECK(hb.list_more == have_cut, "trial %d: more flag", trial);
    for (size_t i = 0; i < valid; i++)
      CHECK(hb.list[i].score == all[i].score && hb.list[i].rank == all[i].rank && hb.list[i].cap == 3 && hb.list[i].loaded &&
                hb.list[i].node == (int)all[i].rank,
            "trial %d entry %zu", trial, i);
    CHECK(hb.ctl.seq == seq_no + 1 && hb.ctl.n_delta == 0, "trial %d: sequence", trial);
  }
  printf("OK record packing (%zu deltas, flush) and 200 multi-GPU list merges\n", want.size());
  return 0;
}
