"""Executable model of the level-parallel formulation of libstdc++'s introsort that the CUDA
kernel uses for RadioSaber's (rbg,slice) ordering (transport.cpp:351-376).

Test infrastructure: documents and checks the *algorithm* (DESIGN.md "sort") against the real
std::sort on CPU, so that the device code only has to implement these array operations.

Facts used (bits/stl_algo.h, g++ 13):
  * __introsort_loop recurses on [cut,last) and loops on [first,cut): the sub-ranges are
    disjoint, so all ranges of one recursion depth can be partitioned at the same time;
  * __unguarded_partition on the range (first,last) with pivot *first: the k-th element from
    the left that is "not before" the pivot (key <= p) is swapped with the k-th element from
    the right that the pivot is "not before" (key >= p) while the former lies left of the
    latter.  Both lists can be read off the array as it was before the partition started;
  * the cut is min(posL[K], posR[K-1]) where K is the number of swaps;
  * __final_insertion_sort is a stable sort of whatever is in the array by then.
"""
from __future__ import annotations

import numpy as np

THRESH = 16  # _S_threshold, stl_algo.h:1848


def _median_to_first(key, a, first, last):
    """__move_median_to_first(first, first+1, mid, last-1) with comp = key greater."""
    mid = first + (last - first) // 2
    pa, pb, pc = first + 1, mid, last - 1
    ka, kb, kc = key[a[pa]], key[a[pb]], key[a[pc]]
    if ka > kb:
        pick = pb if kb > kc else (pc if ka > kc else pa)
    elif ka > kc:
        pick = pa
    elif kb > kc:
        pick = pc
    else:
        pick = pb
    a[first], a[pick] = a[pick], a[first]


def _heap_sort(key, a, first, last):
    """std::__partial_sort(first,last,last) == __heap_select + __sort_heap with comp = key greater
    (stl_heap.h).  Sequential; only reached when the depth limit runs out."""
    seg = list(a[first:last])

    def before(x, y):
        return key[x] > key[y]

    def push_heap(hole, top, v):
        parent = (hole - 1) // 2
        while hole > top and before(seg[parent], v):
            seg[hole] = seg[parent]
            hole = parent
            parent = (hole - 1) // 2
        seg[hole] = v

    def adjust_heap(hole, length, v):
        top = hole
        child = hole
        while child < (length - 1) // 2:
            child = 2 * (child + 1)
            if before(seg[child], seg[child - 1]):
                child -= 1
            seg[hole] = seg[child]
            hole = child
        if (length & 1) == 0 and child == (length - 2) // 2:
            child = 2 * (child + 1)
            seg[hole] = seg[child - 1]
            hole = child - 1
        push_heap(hole, top, v)

    n = len(seg)
    if n >= 2:
        parent = (n - 2) // 2
        while True:
            adjust_heap(parent, n, seg[parent])
            if parent == 0:
                break
            parent -= 1
    lastp = n
    while lastp > 1:
        lastp -= 1
        v = seg[lastp]
        seg[lastp] = seg[0]
        adjust_heap(0, lastp, v)
    a[first:last] = seg


def level_parallel_introsort(keys, depth_limit=-1):
    """Returns perm with perm[i] = original index of the element that ends at position i
    (descending keys), computed level by level the way the kernel does."""
    key = np.asarray(keys)
    n = len(key)
    a = np.arange(n)
    if n == 0:
        return a
    if depth_limit < 0:
        depth_limit = 2 * (int(n).bit_length() - 1)
    segs = [(0, n)]
    level = 0
    while True:
        active = [(f, l) for (f, l) in segs if l - f > THRESH]
        if not active:
            break
        if level == depth_limit:
            for f, l in active:
                _heap_sort(key, a, f, l)
            break
        new = []
        for f, l in active:
            _median_to_first(key, a, f, l)
            p = key[a[f]]
            pos = np.arange(f + 1, l)
            k = key[a[pos]]
            posL = pos[k <= p]                 # ascending positions
            posR = pos[k >= p][::-1]           # descending positions
            m = min(len(posL), len(posR))
            swap = posL[:m] < posR[:m]
            K = int(swap.sum())                # monotone: true...false
            assert swap[:K].all()
            cut = int(posL[K]) if K < len(posL) else 10 ** 9
            if K >= 1:
                cut = min(cut, int(posR[K - 1]))
            lo, hi = posL[:K], posR[:K]
            a[lo], a[hi] = a[hi].copy(), a[lo].copy()
            new += [(f, cut), (cut, l)]
        segs = new
        level += 1
    # __final_insertion_sort: stable by key descending
    order = np.argsort(-key[a].astype(np.int64), kind="stable")
    return a[order]
