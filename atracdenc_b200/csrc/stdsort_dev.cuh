// stdsort_dev.cuh — libstdc++ (GCC 13) std::sort restated step by step.
//
// QuantMantisas (src/atrac/atrac_scale.cpp:78-82) orders its re-rounding candidates with
// std::sort and a comparator on |delta| only.  Candidates with equal |delta| are processed in
// whatever order the library's introsort leaves them in, and the re-rounding loop is order
// dependent, so a bit-exact encoder has to reproduce that order, ties included.  This is
// bits/stl_algo.h's algorithm (__introsort_loop / __unguarded_partition_pivot /
// __final_insertion_sort, threshold 16, heap-sort fallback at depth 2*floor(log2 n)) and
// bits/stl_heap.h's heap primitives, with the same comparisons in the same sequence.
// tests/tools/stdsort_check.cpp checks it against the real std::sort on the host.
#pragma once
#include "atde_cuda.h"

namespace atde {

struct SortCand {
    float delta;
    int idx;
};

ATDE_D bool cand_less(const SortCand& a, const SortCand& b) { return fabsf(a.delta) < fabsf(b.delta); }

ATDE_D void ss_swap(SortCand& a, SortCand& b) { const SortCand t = a; a = b; b = t; }

// __unguarded_linear_insert
ATDE_D void ss_unguarded_linear_insert(SortCand* a, int last)
{
    const SortCand val = a[last];
    int next = last - 1;
    while (cand_less(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

// __insertion_sort on [first, last)
ATDE_D void ss_insertion_sort(SortCand* a, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (cand_less(a[i], a[first])) {
            const SortCand val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];          // move_backward(first, i, i + 1)
            a[first] = val;
        } else {
            ss_unguarded_linear_insert(a, i);
        }
    }
}

// __adjust_heap + __push_heap, heap rooted at a[first], length len
ATDE_D void ss_adjust_heap(SortCand* a, int first, int hole, int len, SortCand value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (cand_less(a[first + child], a[first + child - 1])) child--;
        a[first + hole] = a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[first + hole] = a[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && cand_less(a[first + parent], value)) {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}

// __partial_sort(first, last, last) == __heap_select (make_heap only, middle == last) + __sort_heap
ATDE_D void ss_heap_sort(SortCand* a, int first, int last)
{
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            const SortCand v = a[first + parent];
            ss_adjust_heap(a, first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int l = last;
    while (l - first > 1) {
        --l;
        const SortCand v = a[l];                                        // __pop_heap(first, l, l)
        a[l] = a[first];
        ss_adjust_heap(a, first, 0, l - first, v);
    }
}

// __move_median_to_first(result, a, b, c)
ATDE_D void ss_median_to_first(SortCand* x, int result, int a, int b, int c)
{
    if (cand_less(x[a], x[b])) {
        if (cand_less(x[b], x[c])) ss_swap(x[result], x[b]);
        else if (cand_less(x[a], x[c])) ss_swap(x[result], x[c]);
        else ss_swap(x[result], x[a]);
    } else if (cand_less(x[a], x[c])) ss_swap(x[result], x[a]);
    else if (cand_less(x[b], x[c])) ss_swap(x[result], x[c]);
    else ss_swap(x[result], x[b]);
}

// __unguarded_partition(first, last, pivot)
ATDE_D int ss_unguarded_partition(SortCand* x, int first, int last, int pivot)
{
    for (;;) {
        while (cand_less(x[first], x[pivot])) ++first;
        --last;
        while (cand_less(x[pivot], x[last])) --last;
        if (!(first < last)) return first;
        ss_swap(x[first], x[last]);
        ++first;
    }
}

// std::sort(a, a + n, cmp)
ATDE_D void std_sort_cands(SortCand* a, int n)
{
    if (n <= 0) return;
    // __introsort_loop with its tail recursion on [cut, last) unrolled through an explicit stack
    int stack_first[32], stack_last[32], stack_depth[32];
    int sp = 0;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        const int first = stack_first[sp];
        int last = stack_last[sp];
        int depth = stack_depth[sp];
        // The library recurses on the RIGHT part first (depth-first) and then loops on the left
        // part.  Segments are disjoint, so the order in which they are finished does not change the
        // result; only the comparisons inside each segment matter.
        while (last - first > 16) {
            if (depth == 0) {
                ss_heap_sort(a, first, last);
                last = first;                      // segment done
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            ss_median_to_first(a, first, first + 1, mid, last - 1);
            const int cut = ss_unguarded_partition(a, first + 1, last, first);
            stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        ss_insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) ss_unguarded_linear_insert(a, i);
    } else {
        ss_insertion_sort(a, 0, n);
    }
}

} // namespace atde
