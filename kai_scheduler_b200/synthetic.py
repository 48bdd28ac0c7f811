"""Synthetic cluster snapshots of the shapes BASELINE.json names.

Vectorised restatement of the reference's benchmark topology builders
(pkg/scheduler/actions/benchmark_test.go:199-244 createBenchmarkTopology,
:360-421 createBenchmarkTopologyWithManyQueues, :424-473 ...WithGangJobs) as
they come out of test_utils.BuildSession:

  nodes   "node-%d", 8 GPUs, 2e7 mCPU, 2e10 B, 110 pods          (nodes_fake/nodes.go:31-36,171-180)
  jobs    "job-%d", priority 50 (train, preemptible), queue i % numQueues, each task 1 GPU +
          1000 mCPU + 1e9 B + 1 pod; job i created now-(n-i) min  (jobs_fake/jobs.go:83,261-291)
  queues  "queue-%d" under "dept-%d" (i % numDepts, numDepts = ceil(numQueues/4)), deserved GPUs
          = total/numQueues, over-quota weight 1, CPU/memory unlimited; department over-quota
          weight = its deserved GPUs                               (test_utils_builder.go:96-225)

Snapshot queue order: leaf queues first (creation now+k min), then departments (now+d min).
"""
from __future__ import annotations

import numpy as np

from . import abi


def _name_rank(prefix: str, n: int) -> np.ndarray:
    """Rank of f"{prefix}{i}" in byte-wise ascending order (node-10 < node-2)."""
    names = np.array([f"{prefix}{i}" for i in range(n)], dtype=object)
    order = np.argsort(names, kind="stable")
    rank = np.empty(n, dtype=np.int32)
    rank[order] = np.arange(n, dtype=np.int32)
    return rank


def benchmark_snapshot(n_nodes: int, n_jobs: int, tasks_per_job: int = 1, n_queues: int = 4,
                       gpus_per_node: int = 8, mixed: bool = False, seed: int = 0x0C41,
                       named_depts: bool | None = None) -> abi.Snapshot:
    """createBenchmarkTopology{,WithManyQueues,WithGangJobs} scaled to any size.

    mixed=True draws per-job GPUs in {1,2,4,8}, CPU in {1000,2000,4000} m and memory in
    {1,2,4}e9 B from a seeded PRNG (SURVEY.md §8d config 3 request-mix variant; integer valued).
    """
    R = 4
    N = n_nodes
    alloc = np.empty((R, N), dtype=np.float64)
    alloc[0], alloc[1], alloc[2], alloc[3] = 2e7, 2e10, float(gpus_per_node), 110.0
    idle = alloc.copy()
    rel = np.zeros((R, N), dtype=np.float64)
    name_rank = _name_rank("node-", N)
    flags = np.full(N, abi.NODE_READY, dtype=np.uint32)

    total_gpus = float(N * gpus_per_node)
    if named_depts is None:
        named_depts = n_queues == 4
    if named_depts:  # createBenchmarkTopology / ...WithGangJobs: dept-a {q0,q1}, dept-b {q2,q3}
        n_depts = 2
        q_dept = np.array([0, 0, 1, 1], dtype=np.int32)
        dept_names = ["dept-a", "dept-b"]
    else:  # ...WithManyQueues
        n_depts = max(1, (n_queues + 3) // 4)
        q_dept = (np.arange(n_queues) % n_depts).astype(np.int32)
        dept_names = [f"dept-{i}" for i in range(n_depts)]
    Q = n_queues + n_depts
    parent = np.concatenate([q_dept + n_queues, np.full(n_depts, -1)]).astype(np.int32)
    prio = np.full(Q, 100, dtype=np.int32)
    creation = np.concatenate([np.arange(n_queues) * 60, np.arange(n_depts) * 60]).astype(np.int64)
    qnames = np.array([f"queue-{i}" for i in range(n_queues)] + dept_names, dtype=object)
    uid_rank = np.empty(Q, dtype=np.int32)
    uid_rank[np.argsort(qnames, kind="stable")] = np.arange(Q, dtype=np.int32)
    deserved = np.full((3, Q), -1.0)
    limit = np.full((3, Q), -1.0)
    oqw = np.ones((3, Q))
    per_queue = total_gpus / n_queues
    per_dept = total_gpus / n_depts
    deserved[2, :n_queues] = per_queue
    deserved[2, n_queues:] = per_dept
    oqw[2, :n_queues] = 1.0
    oqw[2, n_queues:] = per_dept

    J = n_jobs
    T = J * tasks_per_job
    job_queue = (np.arange(J) % n_queues).astype(np.int32)
    job_prio = np.full(J, 50, dtype=np.int32)
    # creation now-(n-i) min, UID = name: all distinct creation times -> rank = index
    job_order_rank = np.arange(J, dtype=np.int32)
    job_flags = np.full(J, abi.JOB_PREEMPTIBLE, dtype=np.uint32)
    job_podset_begin = np.arange(J + 1, dtype=np.int32)
    podset_min = np.full(J, tasks_per_job, dtype=np.int32)  # default podset, minAvailable = len(Tasks)
    podset_task_begin = (np.arange(J + 1) * tasks_per_job).astype(np.int32)
    task_status = np.full(T, abi.POD_PENDING, dtype=np.int32)
    task_node = np.full(T, -1, dtype=np.int32)
    req = np.empty((T, R), dtype=np.float64)
    if mixed:
        rng = np.random.default_rng(seed)
        g = rng.choice(np.array([1.0, 2.0, 4.0, 8.0]), size=J)
        c = rng.choice(np.array([1000.0, 2000.0, 4000.0]), size=J)
        m = rng.choice(np.array([1e9, 2e9, 4e9]), size=J)
        req[:, 0] = np.repeat(c, tasks_per_job)
        req[:, 1] = np.repeat(m, tasks_per_job)
        req[:, 2] = np.repeat(g, tasks_per_job)
    else:
        req[:, 0], req[:, 1], req[:, 2] = 1000.0, 1e9, 1.0
    req[:, 3] = 1.0
    # TaskOrderFn falls back to the pod UID "<job>-<k>" (byte-wise): rank of str(k) among 0..tasks_per_job-1
    k_rank = _name_rank("", tasks_per_job)
    task_order_rank = np.tile(k_rank, J).astype(np.int32)

    return abi.Snapshot(
        n_res=R, node_allocatable=alloc, node_idle=idle, node_releasing=rel, node_name_rank=name_rank,
        node_flags=flags, queue_parent=parent, queue_priority=prio, queue_creation=creation,
        queue_uid_rank=uid_rank, queue_deserved=deserved, queue_limit=limit, queue_oqw=oqw,
        job_queue=job_queue, job_priority=job_prio, job_order_rank=job_order_rank, job_flags=job_flags,
        job_podset_begin=job_podset_begin, podset_min_available=podset_min,
        podset_task_begin=podset_task_begin, task_status=task_status, task_node=task_node, task_req=req,
        task_order_rank=task_order_rank)


def cycle_snapshot(n_nodes: int, n_jobs: int, tasks_per_job: int = 4, n_queues: int = 1000, gpus_per_node: int = 4,
                   running_nodes: int = 8, victim_queues: int = 2) -> abi.Snapshot:
    """The allocate + reclaim cycle BASELINE.json's metric names, on the config-3 shape: the many-queues gang benchmark
    (createBenchmarkTopologyWithManyQueues / ...WithGangJobs, benchmark_test.go:360-473) on a cluster whose GPUs equal
    the pending demand, with `running_nodes` nodes already full of running 1-GPU Train pods that belong to
    `victim_queues` leaf queues without quota (deserved 0, over-quota weight 0; department "dept-victims", the
    over-quota shape of reclaim_benchmark_test.go:66-147).  `allocate` places every gang the free GPUs can take; the
    gangs left over sit in queues below their deserved quota and `reclaim` evicts the over-quota pods for them."""
    base = benchmark_snapshot(n_nodes=n_nodes, n_jobs=n_jobs, tasks_per_job=tasks_per_job, n_queues=n_queues,
                              gpus_per_node=gpus_per_node, named_depts=False)
    R, N = 4, n_nodes
    Q0 = int(base.queue_parent.shape[0])
    n_run = running_nodes * gpus_per_node
    # queues: [leaf queues | departments] of the base + [victim leaf queues | dept-victims]
    vq = np.arange(victim_queues) + Q0
    vdept = Q0 + victim_queues
    parent = np.concatenate([base.queue_parent, np.full(victim_queues, vdept), [-1]]).astype(np.int32)
    Q = int(parent.shape[0])
    prio = np.full(Q, 100, dtype=np.int32)
    creation = np.concatenate([base.queue_creation, (n_queues + np.arange(victim_queues)) * 60, [10 ** 6]]).astype(np.int64)
    n_depts = Q0 - n_queues
    qnames = np.array([f"queue-{i}" for i in range(n_queues)] + [f"dept-{i}" for i in range(n_depts)] +
                      [f"victims-{i}" for i in range(victim_queues)] + ["dept-victims"], dtype=object)
    uid_rank = np.empty(Q, dtype=np.int32)
    uid_rank[np.argsort(qnames, kind="stable")] = np.arange(Q, dtype=np.int32)
    deserved = np.full((3, Q), -1.0)
    limit = np.full((3, Q), -1.0)
    oqw = np.ones((3, Q))
    deserved[:, :Q0], oqw[:, :Q0] = base.queue_deserved, base.queue_oqw
    deserved[2, Q0:] = 0.0
    oqw[2, Q0:] = 0.0
    J0, T0 = int(base.job_queue.shape[0]), int(base.task_status.shape[0])
    J, T = J0 + n_run, T0 + n_run
    run_node = (np.arange(n_run) // gpus_per_node).astype(np.int32)  # node-0 .. node-(running_nodes-1), full
    req = np.concatenate([base.task_req, np.tile(np.array([[1000.0, 1e9, 1.0, 1.0]]), (n_run, 1))])
    idle = base.node_allocatable.copy()
    for r in range(R):
        np.subtract.at(idle[r], run_node, req[T0:, r])
    return abi.Snapshot(
        n_res=R, node_allocatable=base.node_allocatable, node_idle=idle, node_releasing=np.zeros((R, N)),
        node_name_rank=base.node_name_rank, node_flags=base.node_flags, queue_parent=parent, queue_priority=prio,
        queue_creation=creation, queue_uid_rank=uid_rank, queue_deserved=deserved, queue_limit=limit, queue_oqw=oqw,
        job_queue=np.concatenate([base.job_queue, vq[np.arange(n_run) % victim_queues]]).astype(np.int32),
        job_priority=np.full(J, 50, dtype=np.int32),
        # running jobs are older than every pending one (they started first)
        job_order_rank=np.concatenate([n_run + np.arange(J0), np.arange(n_run)]).astype(np.int32),
        job_flags=np.full(J, abi.JOB_PREEMPTIBLE, dtype=np.uint32), job_podset_begin=np.arange(J + 1, dtype=np.int32),
        podset_min_available=np.concatenate([base.podset_min_available, np.ones(n_run)]).astype(np.int32),
        podset_task_begin=np.concatenate([base.podset_task_begin, T0 + 1 + np.arange(n_run)]).astype(np.int32),
        task_status=np.concatenate([base.task_status, np.full(n_run, abi.POD_RUNNING)]).astype(np.int32),
        task_node=np.concatenate([base.task_node, run_node]).astype(np.int32), task_req=req,
        task_order_rank=np.concatenate([base.task_order_rank, np.zeros(n_run)]).astype(np.int32))


def reclaim_snapshot(n_nodes: int, running_per_node: int = 8, gpus_per_node: int = 8, victim_queues: int = 1,
                     reclaimer_jobs: int = 1, reclaimer_tasks: int | None = None, reclaimer_gpus: float = 8.0,
                     node_prefix: str = "node") -> abi.Snapshot:
    """Victim-selection workloads.

    Defaults = buildReclaimTopology of BenchmarkReclaimLargeJobs_<N>Node
    (pkg/scheduler/actions/integration_tests/reclaim/reclaim_benchmark_test.go:66-147): N nodes x 8 GPUs, N*8 running
    1-GPU Train jobs (job i on node i % N) in queue-0 (deserved 0, over-quota weight 0), one pending "very-large-job"
    of N/10 tasks x 8 GPUs in queue-1 (deserved all GPUs); no departments => the DSL's `default` department.
    `victim_queues` / `reclaimer_jobs` widen it towards BASELINE config 5 (running pods in several over-quota
    queues, many pending reclaimers).
    """
    R = 4
    N = n_nodes
    alloc = np.empty((R, N), dtype=np.float64)
    alloc[0], alloc[1], alloc[2], alloc[3] = 2e7, 2e10, float(gpus_per_node), 110.0
    name_rank = _name_rank(node_prefix, N)
    flags = np.full(N, abi.NODE_READY, dtype=np.uint32)
    n_run = N * running_per_node
    if reclaimer_tasks is None:
        reclaimer_tasks = max(1, N // 10)
    # queues: victim queues first, then the reclaimer queue, then the default department
    nq = victim_queues + 1
    Q = nq + 1
    parent = np.array([nq] * nq + [-1], dtype=np.int32)
    prio = np.full(Q, 100, dtype=np.int32)
    creation = np.concatenate([np.arange(nq) * 60, [0]]).astype(np.int64)
    qnames = np.array([f"queue-{i}" for i in range(nq)] + ["default"], dtype=object)
    uid_rank = np.empty(Q, dtype=np.int32)
    uid_rank[np.argsort(qnames, kind="stable")] = np.arange(Q, dtype=np.int32)
    deserved = np.full((3, Q), -1.0)
    limit = np.full((3, Q), -1.0)
    oqw = np.ones((3, Q))
    deserved[2, :victim_queues] = 0.0
    deserved[2, victim_queues] = float(N * gpus_per_node)
    oqw[2, :nq] = 0.0
    deserved[2, nq] = -1.0   # default department: DeservedGPUs -1, OverQuotaWeight = DeservedGPUs
    oqw[2, nq] = -1.0

    J = n_run + reclaimer_jobs
    T = n_run + reclaimer_jobs * reclaimer_tasks
    job_queue = np.concatenate([np.arange(n_run) % victim_queues, np.full(reclaimer_jobs, victim_queues)]).astype(np.int32)
    job_prio = np.full(J, 50, dtype=np.int32)
    job_order_rank = np.arange(J, dtype=np.int32)
    job_flags = np.full(J, abi.JOB_PREEMPTIBLE, dtype=np.uint32)
    job_podset_begin = np.arange(J + 1, dtype=np.int32)
    podset_min = np.concatenate([np.ones(n_run), np.full(reclaimer_jobs, reclaimer_tasks)]).astype(np.int32)
    podset_task_begin = np.concatenate([np.arange(n_run), n_run + np.arange(reclaimer_jobs + 1) * reclaimer_tasks]).astype(np.int32)
    task_status = np.concatenate([np.full(n_run, abi.POD_RUNNING), np.full(T - n_run, abi.POD_PENDING)]).astype(np.int32)
    task_node = np.concatenate([np.arange(n_run) % N, np.full(T - n_run, -1)]).astype(np.int32)
    req = np.empty((T, R), dtype=np.float64)
    req[:, 0], req[:, 1], req[:, 3] = 1000.0, 1e9, 1.0
    req[:n_run, 2] = 1.0
    req[n_run:, 2] = reclaimer_gpus
    task_order_rank = np.concatenate([np.zeros(n_run), np.tile(_name_rank("", reclaimer_tasks), reclaimer_jobs)]).astype(np.int32)
    idle = alloc.copy()
    for r in range(R):
        np.subtract.at(idle[r], task_node[:n_run], req[:n_run, r])
    rel = np.zeros((R, N), dtype=np.float64)
    return abi.Snapshot(
        n_res=R, node_allocatable=alloc, node_idle=idle, node_releasing=rel, node_name_rank=name_rank,
        node_flags=flags, queue_parent=parent, queue_priority=prio, queue_creation=creation,
        queue_uid_rank=uid_rank, queue_deserved=deserved, queue_limit=limit, queue_oqw=oqw,
        job_queue=job_queue, job_priority=job_prio, job_order_rank=job_order_rank, job_flags=job_flags,
        job_podset_begin=job_podset_begin, podset_min_available=podset_min,
        podset_task_begin=podset_task_begin, task_status=task_status, task_node=task_node, task_req=req,
        task_order_rank=task_order_rank)


def topology_snapshot(n_nodes: int, n_gangs: int, nodes_per_rack: int = 16, racks_per_leaf: int = 16, leaves_per_spine: int = 13,
                      gpus_per_node: int = 8, n_queues: int = 4, seed: int = 0x0C44, min_pods: int = 2, max_pods: int = 16,
                      running_fraction: float = 0.0) -> abi.Snapshot:
    """BASELINE config 4: nodes labelled on 3 tiers (spine / leaf / rack; Topology CR levels
    [spine, leaf, rack]), gangs of seeded 2..16 pods x 8 GPUs (node-exclusive) with
    topologyConstraint{requiredLevel = leaf | rack, preferredLevel = rack}.  `running_fraction` pre-fills that share of
    the nodes with one running 8-GPU pod each (so that domains differ in free capacity)."""
    base = benchmark_snapshot(n_nodes=n_nodes, n_jobs=1, tasks_per_job=1, n_queues=n_queues, gpus_per_node=gpus_per_node)
    rng = np.random.default_rng(seed)
    N, R = n_nodes, 4
    rack = np.arange(N) // nodes_per_rack
    leaf = rack // racks_per_leaf
    spine = leaf // leaves_per_spine
    # DomainID = joined label values; ids must follow ascending DomainID STRING order per level
    def dense(ids_str):
        uniq = sorted(set(ids_str))
        m = {d: i for i, d in enumerate(uniq)}
        return np.array([m[d] for d in ids_str], dtype=np.int32)
    spine_id = [f"s{spine[n]}" for n in range(N)]
    leaf_id = [f"s{spine[n]}.l{leaf[n]}" for n in range(N)]
    rack_id = [f"s{spine[n]}.l{leaf[n]}.r{rack[n]}" for n in range(N)]
    node_domain = np.stack([dense(spine_id), dense(leaf_id), dense(rack_id)]).astype(np.int32)
    sizes = rng.integers(min_pods, max_pods + 1, size=n_gangs)
    n_run = int(N * running_fraction)
    run_nodes = rng.choice(N, size=n_run, replace=False) if n_run else np.zeros(0, dtype=np.int64)
    J = n_gangs + n_run
    T = int(sizes.sum()) + n_run
    job_queue = (np.arange(J) % n_queues).astype(np.int32)
    podset_min = np.concatenate([sizes, np.ones(n_run, dtype=np.int64)]).astype(np.int32)
    podset_task_begin = np.concatenate([[0], np.cumsum(podset_min)]).astype(np.int32)
    task_status = np.concatenate([np.full(int(sizes.sum()), abi.POD_PENDING), np.full(n_run, abi.POD_RUNNING)]).astype(np.int32)
    task_node = np.concatenate([np.full(int(sizes.sum()), -1), run_nodes]).astype(np.int32)
    req = np.empty((T, R))
    req[:, 0], req[:, 1], req[:, 2], req[:, 3] = 1000.0, 1e9, float(gpus_per_node), 1.0
    order = np.concatenate([_name_rank("", int(k)) for k in podset_min]).astype(np.int32)
    idle = base.node_allocatable.copy()
    for r in range(R):
        np.subtract.at(idle[r], run_nodes, req[int(sizes.sum()):, r])
    required = rng.integers(1, 3, size=J).astype(np.int32)  # 1 = leaf, 2 = rack
    job_topology = np.concatenate([np.zeros(n_gangs), np.full(n_run, -1)]).astype(np.int32)
    return abi.Snapshot(
        n_res=R, node_allocatable=base.node_allocatable, node_idle=idle, node_releasing=np.zeros((R, N)),
        node_name_rank=base.node_name_rank, node_flags=base.node_flags, queue_parent=base.queue_parent,
        queue_priority=base.queue_priority, queue_creation=base.queue_creation, queue_uid_rank=base.queue_uid_rank,
        queue_deserved=base.queue_deserved, queue_limit=base.queue_limit, queue_oqw=base.queue_oqw,
        job_queue=job_queue, job_priority=np.full(J, 50, dtype=np.int32), job_order_rank=np.arange(J, dtype=np.int32),
        job_flags=np.full(J, abi.JOB_PREEMPTIBLE, dtype=np.uint32), job_podset_begin=np.arange(J + 1, dtype=np.int32),
        podset_min_available=podset_min, podset_task_begin=podset_task_begin, task_status=task_status,
        task_node=task_node, task_req=req, task_order_rank=order,
        topology_level_begin=np.array([0, 3], dtype=np.int32), node_domain=node_domain, job_topology=job_topology,
        job_required_level=np.where(job_topology >= 0, required, -1).astype(np.int32),
        job_preferred_level=np.where(job_topology >= 0, 2, -1).astype(np.int32))


# The BASELINE.json configs
CONFIGS = {
    # name: kwargs
    "config1": dict(n_nodes=100, n_jobs=500, tasks_per_job=1, n_queues=10),
    "config2": dict(n_nodes=10_000, n_jobs=40_000, tasks_per_job=1, n_queues=4),
    "config3": dict(n_nodes=50_000, n_jobs=50_000, tasks_per_job=4, n_queues=1000),
    "config3-mixed": dict(n_nodes=50_000, n_jobs=50_000, tasks_per_job=4, n_queues=1000, mixed=True),
}
# victim-selection workloads: BenchmarkReclaimLargeJobs_<N>Node (reference numbers in BASELINE.md) and a
# config-5-shaped full cycle (running pods in over-quota queues + pending reclaimers)
TOPOLOGY_CONFIGS = {
    "config4-small": dict(n_nodes=4096, n_gangs=600),
    "config4": dict(n_nodes=50_000, n_gangs=20_000),
}
RECLAIM_CONFIGS = {
    **{f"reclaim-large-{n}": dict(n_nodes=n) for n in (10, 50, 100, 200, 500, 1000)},
    "cycle5-small": dict(n_nodes=200, running_per_node=8, victim_queues=4, reclaimer_jobs=100, reclaimer_tasks=2,
                         reclaimer_gpus=4.0),
    # BASELINE config 5 ladder (reclaim + consolidate cycle: every node full of running over-quota pods, pending reclaimers)
    "cycle5-1000": dict(n_nodes=1000, running_per_node=8, victim_queues=4, reclaimer_jobs=250, reclaimer_tasks=2,
                        reclaimer_gpus=4.0),
    "cycle5-5000": dict(n_nodes=5000, running_per_node=8, victim_queues=4, reclaimer_jobs=500, reclaimer_tasks=2,
                        reclaimer_gpus=4.0),
}
# BASELINE.json's metric: one allocate + reclaim cycle over 50k nodes / 200k pending pods / 1k queues (+250 departments)
CYCLE_CONFIGS = {
    "config3-cycle": dict(n_nodes=50_000, n_jobs=50_000, tasks_per_job=4, n_queues=1000, gpus_per_node=4, running_nodes=8),
    "config3-cycle-small": dict(n_nodes=600, n_jobs=600, tasks_per_job=4, n_queues=40, gpus_per_node=4, running_nodes=4),
}
CONFIG_ACTIONS = {**{k: ["allocate", "reclaim"] for k in CYCLE_CONFIGS}, **{k: ["allocate"] for k in CONFIGS}, **{k: ["allocate"] for k in TOPOLOGY_CONFIGS},
                  **{k: ["reclaim"] for k in RECLAIM_CONFIGS},
                  **{k: ["allocate", "consolidation", "reclaim"] for k in ("cycle5-small", "cycle5-1000", "cycle5-5000")}}
# ms/op the reference publishes for BenchmarkReclaimLargeJobs (other hardware, includes BuildSession)
REFERENCE_PUBLISHED_MS = {"reclaim-large-10": 104.4, "reclaim-large-50": 130.2, "reclaim-large-100": 241.2,
                          "reclaim-large-200": 816.0, "reclaim-large-500": 8970.0}


def config_snapshot(name: str) -> abi.Snapshot:
    if name in RECLAIM_CONFIGS:
        return reclaim_snapshot(**RECLAIM_CONFIGS[name])
    if name in TOPOLOGY_CONFIGS:
        return topology_snapshot(**TOPOLOGY_CONFIGS[name])
    if name in CYCLE_CONFIGS:
        return cycle_snapshot(**CYCLE_CONFIGS[name])
    return benchmark_snapshot(**CONFIGS[name])
