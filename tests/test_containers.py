"""libcpecan.so's own stList / stIntTuple (host/containers.c): result lists whose tuples are cut from one slab.

A caller of cPecan may do anything sonLib allows with the lists it gets back: destruct the list, destruct single tuples in any
order, pop, replace, reverse, sort.  The slab goes with its last tuple whichever way they go, nothing is freed twice (glibc aborts
on that), and nothing is left behind (resident memory does not grow over many rounds)."""
import ctypes as C
import os
import resource

import numpy as np
import pytest

import helpers

LIB = os.path.join(helpers.ROOT, "cpecan_b200", "lib", "libcpecan.so")


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(LIB):
        import __graft_entry__ as g

        g.build()
    lib = C.CDLL(LIB)
    for f in ("cpecan_tripleList_construct", "stList_get", "stList_pop", "stIntTuple_construct3"):
        getattr(lib, f).restype = C.c_void_p
    for f in ("stList_length", "stIntTuple_get"):
        getattr(lib, f).restype = C.c_int64
    lib.cpecan_tripleList_construct.argtypes = [C.c_void_p, C.c_int64]
    lib.stList_get.argtypes = [C.c_void_p, C.c_int64]
    lib.stList_set.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.stList_pop.argtypes = [C.c_void_p]
    lib.stList_length.argtypes = [C.c_void_p]
    lib.stList_destruct.argtypes = [C.c_void_p]
    lib.stList_reverse.argtypes = [C.c_void_p]
    lib.stList_sort.argtypes = [C.c_void_p, C.c_void_p]
    lib.stList_setDestructor.argtypes = [C.c_void_p, C.c_void_p]
    lib.stIntTuple_get.argtypes = [C.c_void_p, C.c_int64]
    lib.stIntTuple_destruct.argtypes = [C.c_void_p]
    lib.stIntTuple_construct3.argtypes = [C.c_int64, C.c_int64, C.c_int64]
    return lib


def make(L, n, seed=0):
    t = np.random.default_rng(seed).integers(0, 10 ** 7, size=(n, 3), dtype=np.int32)
    return L.cpecan_tripleList_construct(t.ctypes.data, n), t


def read(L, l):
    return [[L.stIntTuple_get(L.stList_get(l, i), k) for k in range(3)] for i in range(L.stList_length(l))]


def test_slab_list_holds_the_triples_and_survives_reordering(L):
    l, t = make(L, 500)
    assert read(L, l) == t.tolist()
    L.stList_reverse(l)
    assert read(L, l) == t[::-1].tolist()
    L.stList_sort(l, C.cast(L.stIntTuple_cmpFn, C.c_void_p))
    assert read(L, l) == sorted(t.tolist())
    L.stList_destruct(l)  # still exactly the slab's tuples: they go back together
    l, _ = make(L, 0)
    assert L.stList_length(l) == 0
    L.stList_destruct(l)


def test_every_way_of_giving_the_tuples_back(L):
    rng = np.random.default_rng(5)
    for round_ in range(30):
        n = int(rng.integers(1, 400))
        # a popped tuple is the caller's: destructed on its own, before or after the list
        l, _ = make(L, n, round_)
        popped = [L.stList_pop(l) for _ in range(int(rng.integers(0, n + 1)))]
        order = rng.permutation(len(popped))
        for k in order[: len(order) // 2]:
            L.stIntTuple_destruct(popped[k])
        L.stList_destruct(l)
        for k in order[len(order) // 2:]:
            L.stIntTuple_destruct(popped[k])
        # a replaced tuple: the old one is destructed by the caller, the new one (its own allocation) by the list
        l, _ = make(L, n, round_)
        for i in rng.choice(n, size=min(n, 5), replace=False):
            old = L.stList_get(l, int(i))
            L.stList_set(l, int(i), L.stIntTuple_construct3(1, 2, 3))
            L.stIntTuple_destruct(old)
        L.stList_destruct(l)
        # the list gives up ownership, the tuples are destructed one by one in any order
        l, _ = make(L, n, round_)
        items = [L.stList_get(l, i) for i in range(n)]
        L.stList_setDestructor(l, None)
        L.stList_destruct(l)
        for k in rng.permutation(n):
            L.stIntTuple_destruct(items[k])


def test_nothing_is_left_behind(L):
    """200 rounds of 200 000-tuple lists (6.4 MB of slab each) destructed in the three ways: resident memory stays where it was"""

    def rss_kb():
        with open("/proc/self/statm") as f:
            return int(f.read().split()[1]) * resource.getpagesize() // 1024

    def round_(k):
        l, _ = make(L, 200000, k)
        if k % 3 == 1:
            L.stIntTuple_destruct(L.stList_pop(l))
        if k % 3 == 2:
            L.stList_reverse(l)
        L.stList_destruct(l)

    for k in range(20):
        round_(k)
    before = rss_kb()
    for k in range(200):
        round_(k)
    assert rss_kb() - before < 64 * 1024, "resident memory grew by %d KB over 200 rounds" % (rss_kb() - before)
