/*
 * kai_oracle.h — C entry points of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * The oracle is a CPU restatement of the reference's per-Session scheduling
 * cycle (NVIDIA/KAI-Scheduler, pkg/scheduler).  It consumes the same
 * kai_snapshot / kai_config as libkaigpu.so (include/kai_engine.h) and returns
 * a kai_result with identical meaning, so that tests can compare the CUDA
 * engine with it field by field.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load this library.  The product (kai_scheduler_b200/) never does.
 */
#ifndef KAI_ORACLE_H_
#define KAI_ORACLE_H_
#include "../include/kai_engine.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kai_oracle kai_oracle;

int kai_oracle_create(const kai_config *cfg, kai_oracle **out);
int kai_oracle_load_snapshot(kai_oracle *o, const kai_snapshot *snap);
int kai_oracle_run(kai_oracle *o, kai_action action, kai_result *out);
int kai_oracle_fair_share(kai_oracle *o, kai_result *out);
int kai_oracle_stats(kai_oracle *o, kai_stats *out);
/* threads used for the per-task node sweep (mirrors the reference's
   goroutine-per-node fan-out, framework/session.go:243-262); 1 = scalar */
int kai_oracle_set_threads(kai_oracle *o, int n_threads);
void kai_oracle_destroy(kai_oracle *o);
const char *kai_oracle_last_error(const kai_oracle *o);

/* Known-answer hooks used by tests/ to pin the oracle against the
   reference's own unit tests. */
/* plugins/nodeplacement/pack.go:45-64 getScoreOfCurrentNode */
double kai_oracle_binpack_score(double min_alloc, double max_alloc, double cur, double node_overall);
/* plugins/nodeplacement/spread.go:16-36 */
double kai_oracle_spread_score(double non_allocated, double resource_count);
/* plugins/topology/node_scoring.go:36-53: score of the nodes of the i-th of n preferred-level domains */
double kai_oracle_topology_position_score(int i, int n);
/* plugins/proportion/resource_division/resource_division.go:33-43 setResourceShare
   on ONE sibling group and ONE resource; arrays of length n; fair_share is in/out.
   returns the remaining amount. */
double kai_oracle_set_resource_share(int n, double total, double k_value,
                                     const double *deserved, const double *limit,
                                     const double *oqw, const double *request,
                                     const double *usage, const int32_t *priority,
                                     const int64_t *creation, const int32_t *uid_rank,
                                     double *fair_share);
/* resource_division.go:111-144 divideOverQuotaResource alone (FairShare preset by the caller); same arrays */
double kai_oracle_divide_over_quota(int n, double amount, double k_value,
                                    const double *deserved, const double *limit,
                                    const double *oqw, const double *request,
                                    const double *usage, const int32_t *priority,
                                    const int64_t *creation, const int32_t *uid_rank,
                                    double *fair_share);
/* plugins/proportion/queue_order/queue_order.go:19-73 for two queues described by
   8-field ResourceShare rows [3][8] = {Deserved,FairShare,MaxAllowed,OverQuotaWeight,
   Allocated,AllocatedNotPreemptible,Request,Usage}; job requirement vectors [3].
   returns -1 (l first) or 1 (r first). */
int kai_oracle_queue_order(const double *l_share, const double *r_share,
                           int l_priority, int r_priority,
                           int64_t l_creation, int64_t r_creation,
                           const double *l_job_req, const double *r_job_req,
                           const double *total);

/* actions/common/minimal_job_comparison.go on two jobs of the loaded snapshot with equal job_signature:
   IsEasierToSchedule(job) against `representative`, and whether UpdateRepresentative(job) replaces it. */
int kai_oracle_job_easier_to_schedule(kai_oracle *o, int job, int representative);
int kai_oracle_job_replaces_representative(kai_oracle *o, int job, int representative);
/* framework.Statement operations on the loaded snapshot (statement.go): kinds 0 Evict(task), 1 Pipeline(task, node),
   2 Allocate(task, node), 3 undoOperation(index in task[]), 4 Discard, 5 Pipeline(update = false), 6 Rollback(checkpoint in
   task[]); returns the number of operations in the log (= Checkpoint()) or a negative error; the state is read with
   kai_oracle_fair_share. */
int kai_oracle_statement_exercise(kai_oracle *o, int n_ops, const int32_t *kinds, const int32_t *task, const int32_t *node);
/* scheduler_util.PriorityQueue (priority_queue.go:50-118) over the oracle's container/heap restatement, ints with `<`:
   ops 0 push(val) (+ max-size eviction when max_size >= 0), 1 pop, 2 peek, 3 items[0] = val; Fix(0), 4 len. */
void kai_oracle_priority_queue_exercise(int max_size, int n_ops, const int32_t *ops, const int32_t *vals, int32_t *out);
/* accumulated_scenario_filters/idle_gpus/common.go:34-64 greedyMatchRequirements; both arrays sorted descending */
int kai_oracle_greedy_match(int n_req, const double *req, int n_holders, const double *capacity);
/* podgroup_info.GetTasksToAllocate (allocation_info.go:27-54) of one job of the loaded snapshot: task indices in attempt
   order (returns the count); kai_oracle_set_task_virtual sets PodInfo.IsVirtualStatus of a task first if needed. */
int kai_oracle_tasks_to_allocate(kai_oracle *o, int job, int real_allocation, int32_t *out, int cap);
int kai_oracle_set_task_virtual(kai_oracle *o, int task, int is_virtual);
/* plugins/proportion/reclaimable/reclaimable.go:29-51 CanReclaimResources for one queue: share[3][4] =
   {Deserved, FairShare, Allocated, AllocatedNotPreemptible} per resource (cpu, memory, gpu), req[3]. */
int kai_oracle_can_reclaim_resources(const double *share, const double *req, int preemptible);
/* proportion.setFairShare (proportion.go:403-423) on an explicit queue tree: in[n_queues][3][4] = {Deserved, MaxAllowed,
   OverQuotaWeight, Request}; fair_share[n_queues][3] is written. */
int kai_oracle_set_fair_share_tree(int n_queues, const int32_t *parent, const int32_t *priority, const int64_t *creation,
                                   const int32_t *uid_rank, const double *in, const double *total, double k_value,
                                   double *fair_share);
/* capacity_policy.go:26-84 on an explicit queue tree (share[n_queues][3][5] as for kai_oracle_reclaimable): mode 0 =
   limit + non-preemptible quota checks, mode 1 = the quota check alone; returns IsSchedulable. */
int kai_oracle_capacity_schedulable(int n_queues, const int32_t *parent, const double *share, int queue, int preemptible,
                                    const double *req, int mode);
/* reclaimable/strategies/strategies.go: strategy 0 = MaintainFairShareStrategy (:42-57), 1 = GuaranteeDeservedQuota
   (:59-91) on two queue rows share[3][5] = {Deserved, FairShare, Allocated, AllocatedNotPreemptible, MaxAllowed}. */
int kai_oracle_reclaim_strategy(int strategy, const double *reclaimer_share, const double *reclaimee_share,
                                const double *reclaimer_req, const double *remaining);
/* reclaimable.go:53-232 Reclaimable on an explicit queue tree: share[n_queues][3][5] = the four fields above +
   MaxAllowed, victims =
   (leaf queue, resources[3]) in the given order. */
int kai_oracle_reclaimable(int n_queues, const int32_t *parent, const double *share, double saturation_multiplier,
                           int reclaimer_queue, int preemptible, const double *req, int n_victims,
                           const int32_t *victim_queue, const double *victim_res);
/* actions/common/feasible_nodes.go:11-26 FeasibleNodesForJob on the loaded snapshot: out[n_nodes] = 1 for kept nodes */
int kai_oracle_feasible_nodes(kai_oracle *o, int job, int32_t *out);
/* resource_share/{resource_share,queue_resource_share,resource_quantities}.go on one queue row share[3][6] = {Deserved,
   FairShare, Allocated, AllocatedNotPreemptible, MaxAllowed, Request}: out[0] = GetDominantResourceShare(total),
   out[1..3] = GetAllocatableShare per resource, out[4..6] = GetRequestableShare per resource. */
void kai_oracle_queue_attributes(const double *share, const double *total, double *out);
/* resource_quantities.go:81-97 compareQuantities; :48-79 Less (kind 0), LessEqual (1), LessInAtLeastOneResource (2) */
int kai_oracle_compare_quantities(double a, double b);
int kai_oracle_quantities_relation(int kind, const double *a, const double *b);
/* plugins/minruntime/resolver.go on the loaded snapshot's queue tree: getReclaimMinRuntime (method of the config)
   for (pending queue, victim queue) when reclaim != 0, else getPreemptMinRuntime(victim queue); -1 = nil queue. */
double kai_oracle_min_runtime(kai_oracle *o, int reclaim, int pending_queue, int victim_queue);
/* !reclaimFilterFn / !preemptFilterFn (minruntime.go:93-105): 1 = the victim job is withheld */
int kai_oracle_min_runtime_protected(kai_oracle *o, int reclaim, int pending_job, int victim_job);
/* NodeInfo.PodInfos of all nodes as (task, node, status-of-the-clone) triples (node_info.go:400-402): a task holds one
   entry per node it sits on.  Returns the number of entries (written up to cap). */
int kai_oracle_node_entries(kai_oracle *o, int32_t *task, int32_t *node, int32_t *status, int cap);

#ifdef __cplusplus
}
#endif
#endif
