"""plugins/proportion/resource_share/*_test.go on the oracle's restatement of ResourceShare / QueueResourceShare /
ResourceQuantities (CPU only).  Cases are transcribed by hand with the Go line of each table.

Rows are (cpu, memory, gpu); U = commonconstants.UnlimitedResourceQuantity.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_lib import lib  # noqa: E402

U = -1.0
NO_FAIR_SHARE_DRF_MULTIPLIER = 1000.0  # queue_resource_share.go:18


def _attributes(deserved=(0, 0, 0), fair=(0, 0, 0), allocated=(0, 0, 0), max_allowed=(U, U, U), request=(0, 0, 0), total=(0, 0, 0)):
    l = lib()
    dp = C.POINTER(C.c_double)
    l.kai_oracle_queue_attributes.argtypes = [dp, dp, dp]
    l.kai_oracle_queue_attributes.restype = None
    share = np.array([[deserved[r], fair[r], allocated[r], 0.0, max_allowed[r], request[r]] for r in range(3)], dtype=np.float64)
    tot = np.array(total, dtype=np.float64)
    out = np.zeros(7, dtype=np.float64)
    l.kai_oracle_queue_attributes(share.ctypes.data_as(dp), tot.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return {"dominant": out[0], "allocatable": tuple(out[1:4]), "requestable": tuple(out[4:7])}


# queue_attributes_test.go:42-203 GetRequestedResource: the same four cases for GPU, CPU and memory
@pytest.mark.parametrize("resource", [0, 1, 2])
@pytest.mark.parametrize("requested,max_allowed,expected", [(5, 6, 5), (5, 5, 5), (5, 3, 3), (5, U, 5)])
def test_get_requestable_share(resource, requested, max_allowed, expected):
    req, lim = [0, 0, 0], [U, U, U]
    req[resource], lim[resource] = requested, max_allowed
    assert _attributes(request=req, max_allowed=lim)["requestable"][resource] == expected


# queue_attributes_test.go:205-276 DominantResource (MaxAllowed unlimited on every resource)
@pytest.mark.parametrize("name,deserved,fair,allocated,total,expected", [
    ("CPU is most dominant", (1000, 1000, 1000), (1000, 1000, 1000), (500, 100, 200), (0, 0, 0), 500.0 / 1000.0),
    ("Memory is most dominant", (1000, 1000, 1000), (1000, 1000, 1000), (500, 600, 0), (0, 0, 0), 600.0 / 1000.0),
    ("GPU is most dominant", (1000, 1000, 1000), (1000, 1000, 1000), (500, 600, 700), (0, 0, 0), 700.0 / 1000.0),
    ("CPU allocated but not deserved", (0, 1000, 1000), (0, 1000, 1000), (500, 0, 700), (0, 0, 0), 500.0 * NO_FAIR_SHARE_DRF_MULTIPLIER),
    ("unlimited deserved GPU - CPU dominant", (1000, 1000, U), (500, 1000, 1500), (500, 0, 700), (10000, 10000, 2000), 500.0 / 1000.0),
    ("unlimited deserved GPU - GPU dominant", (10000, 1000, U), (500, 1000, 1500), (500, 0, 700), (10000, 10000, 2000), 700.0 / 2000.0),
])
def test_dominant_resource_share(name, deserved, fair, allocated, total, expected):
    assert _attributes(deserved=deserved, fair=fair, allocated=allocated, total=total)["dominant"] == expected


# queue_attributes_test.go:278-335 GetAllocatableShare
@pytest.mark.parametrize("name,deserved,fair,max_allowed,expected", [
    ("maxAllowed is limiting all resources", (1000, 1000, 1000), (1000, 1000, 1000), (500, 100, 200), (500, 100, 200)),
    ("maxAllowed is limiting some resources", (1000, 1000, 1000), (1500, 1500, 1500), (2000, 600, 2000), (1500, 600, 1500)),
    ("maxAllowed is limiting some, deserved is dominant in other", (2000, 1000, 1000), (1000, 1500, 1500), (3000, 600, 3000), (2000, 600, 1500)),
    ("maxAllowed is unlimited", (2000, 1000, 1000), (1000, 1500, 1500), (U, U, U), (2000, 1500, 1500)),
    ("deserved is unlimited maxAllowed is sometimes not", (U, U, U), (1000, 1500, 1500), (2000, U, 1000), (2000, U, 1000)),
])
def test_get_allocatable_share(name, deserved, fair, max_allowed, expected):
    assert _attributes(deserved=deserved, fair=fair, max_allowed=max_allowed)["allocatable"] == expected


# resource_share_test.go:44-82 on createResourceShare() = {Deserved 21, FairShare 22, MaxAllowed 10, Request 17, ...}
def test_resource_share_limited_and_unlimited():
    limited = _attributes(deserved=(21,) * 3, fair=(22,) * 3, allocated=(5,) * 3, max_allowed=(10,) * 3, request=(17,) * 3)
    assert limited["requestable"] == (10, 10, 10) and limited["allocatable"] == (10, 10, 10)  # :44-48, :57-61
    free = _attributes(deserved=(21,) * 3, fair=(22,) * 3, allocated=(5,) * 3, max_allowed=(U,) * 3, request=(17,) * 3)
    assert free["requestable"] == (17, 17, 17) and free["allocatable"] == (22, 22, 22)  # :50-55, :63-68


# resource_quantities_test.go:106-148 TestCompareResources
@pytest.mark.parametrize("a,b,expected", [(1.5, 2.5, -1), (2.5, 1.5, 1), (2.5, 2.5, 0), (U, 2.5, 1), (2.5, U, -1), (U, U, 0)])
def test_compare_quantities(a, b, expected):
    l = lib()
    l.kai_oracle_compare_quantities.argtypes = [C.c_double, C.c_double]
    assert l.kai_oracle_compare_quantities(a, b) == expected


# resource_quantities_test.go:64-104 Less / LessEqual / LessInAtLeastOneResource around (cpu, memory, gpu) = (111, 22, 0.5)
@pytest.mark.parametrize("kind,delta,expected", [
    (0, (1, 1, 0.1), True), (0, (1, 0, 0.1), False),          # :64-74 Less
    (1, (1, 1, 0), True), (1, (1, 1, -0.1), False),           # :76-86 LessEqual
    (2, (1, 0, -0.1), True), (2, (0, -1, -0.1), False),       # :88-98 LessInAtLeastOneResource
])
def test_quantities_relations(kind, delta, expected):
    l = lib()
    dp = C.POINTER(C.c_double)
    l.kai_oracle_quantities_relation.argtypes = [C.c_int, dp, dp]
    a = np.array([111.0, 22.0, 0.5])
    b = a + np.array(delta, dtype=np.float64)
    assert bool(l.kai_oracle_quantities_relation(kind, a.ctypes.data_as(dp), b.ctypes.data_as(dp))) is expected
