// kai_device.cuh — device-side data layout shared by the kernels of libkaigpu.so.
//
// HBM layout:
//   node tables   resource-major f64 [R][N]  (allocatable, idle, releasing) + name_rank/flags [N]
//   queue tables  resource-major f64 [3][Q]
//   task request  task-major f64 [T][R]
//   session state (one copy): task status/node/virtual, queue allocated, node idle/releasing
//   replica state (one copy per CTA of the persistent action kernel): the mutable part of the
//     session that the replicated sequencer (thread 0 of every CTA) updates in lock step
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/kai_engine.h"

namespace kai {

constexpr int QR = KAI_QRES;
constexpr int kThreads = 512;          // threads per CTA of the action kernel
constexpr int kMaxGrid = 1024;         // exchange slots per GPU
constexpr uint32_t kNoRank = 0xFFFFFFFFu;

constexpr int kActiveUsed = KAI_POD_ALLOCATED | KAI_POD_PIPELINED | KAI_POD_BINDING | KAI_POD_BOUND |
                            KAI_POD_RUNNING | KAI_POD_RELEASING;
constexpr int kActiveAllocated =
    KAI_POD_ALLOCATED | KAI_POD_PIPELINED | KAI_POD_BINDING | KAI_POD_BOUND | KAI_POD_RUNNING;
constexpr int kAlive = kActiveAllocated | KAI_POD_PENDING | KAI_POD_GATED;
constexpr int kAllocatedStatuses = KAI_POD_ALLOCATED | KAI_POD_BOUND | KAI_POD_BINDING | KAI_POD_RUNNING;

struct Op {  // framework/statement.go operations (allocate / pipeline / evict / undo)
  int kind, task, prev_status, prev_node, next_node, prev_virtual, undo_index, pad;
};
enum { OP_ALLOCATE = 0, OP_PIPELINE = 1, OP_EVICT = 2, OP_UNDO = 3 };

// queue-node flags of the job-order tree (actions/utils/job_order_by_queue.go:18-25)
enum { QN_EXISTS = 1, QN_LINKED = 2, QN_REORDER = 4 };

// Immutable (per cycle) device snapshot + session state pointers.  Passed by value to kernels.
struct DevSnap {
  int R, N, Q, J, S, T, NPC, mask_words;
  int n_top, max_job_tasks, max_job_podsets, n_levels;
  // nodes
  const double *alloc;      // [R][N]
  double *idle, *rel;       // [R][N] session state
  const int *name_rank;     // [N]
  const int *rank_to_node;  // [N]
  const uint32_t *nflags;   // [N]
  const double *gpu_count;  // [N]
  const double *foreign;    // [3][N] or null
  // queues
  const int *q_parent, *q_priority, *q_uid_rank, *q_nchildren;
  const long long *q_creation;
  const double *q_deserved, *q_limit, *q_oqw, *q_usage;  // [3][Q]
  double *q_fair, *q_request;                            // [3][Q] computed by open-session kernels
  double *q_alloc, *q_alloc_np;                          // [3][Q] session state
  const int *q_child_begin, *q_children;                 // CSR of children (ascending index)
  const int *top_queues;                                 // [n_top]
  const int *level_group_begin, *level_groups;           // fair-share: groups (parent queue or -1) per level
  const int *q_job_begin;                                // [Q+1] leaf-heap arena offsets
  const int *q_jobs_sorted;                              // [J] jobs grouped by queue, (priority desc, order_rank)
  // jobs
  const int *j_queue, *j_priority, *j_order_rank, *j_ps_begin;
  const uint32_t *j_flags;
  // podsets
  const int *ps_min, *ps_task_begin, *ps_job;
  const int *ps_sorted_tasks;  // [T] tasks of each podset in TaskOrderFn order
  // tasks
  const double *t_req;  // [T][R]
  const int *t_job, *t_podset, *t_nominated, *t_pred_class;
  int *t_status, *t_node, *t_node_status;  // session state
  unsigned char *t_virtual;                // session state
  const uint32_t *pred_mask;
  double *total;  // [3] device
};

// Per-CTA replica of the mutable session state (arrays live in one big arena per replica).
struct Replica {
  double *q_alloc, *q_alloc_np;  // [3][Q]
  int *t_status, *t_node, *t_node_status;
  unsigned char *t_virtual;
  int *ps_active_alloc;  // [S] tasks in an active-allocated status
  double *j_req;         // [J][3] cached GetTasksToAllocateInitResource
  unsigned char *j_req_valid;
  int *leaf_heap;  // [J]
  int *leaf_len;   // [Q]
  int *child_heap; // [Q]
  int *child_len;  // [Q]
  int *root_heap;  // [n_top]
  unsigned char *qn_flags;  // [Q]
  Op *ops;                  // [ops_cap]
  int *tta;                 // [max_job_tasks]
  int *ps_order;            // [max_job_podsets]
};

struct ActionParams {
  DevSnap s;
  kai_config cfg;
  int action;
  int grid;             // CTAs of this GPU
  int nodes_per_cta;    // node rows per CTA (tile height)
  int node_base;        // first node row of this GPU's shard
  int node_count;       // node rows of this GPU's shard
  unsigned char *replica_arena;  // grid * replica_bytes
  size_t replica_bytes;
  int ops_cap;
  unsigned long long *xbuf;  // exchange slots: [2][kMaxGrid][4] u64 (A word pair + B word pair)
  unsigned long long *mmbuf; // min/max exchange: [2][kMaxGrid][8] u64
  kai_job_visit *visits;     // [visits_cap]
  int visits_cap;
  long long *counters;  // [8]: n_visits, decisions, nodes_scanned, pods_placed, pods_evicted, minmax_exchanges, error
  unsigned int seq0;    // first exchange sequence number of this launch
};

}  // namespace kai
