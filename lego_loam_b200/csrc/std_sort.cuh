// std_sort.cuh — libstdc++'s std::sort, move for move, on packed (value, ind) records.
//
// extractFeatures sorts every ring sector's cloudSmoothness records with std::sort and a comparator that looks at the
// curvature only (FA:57-61, FA:699).  std::sort is not stable, equal curvatures are common (ranges are quantised), and
// the greedy picks that follow take the records in sorted order - so WHICH of several equal-curvature points becomes
// a feature is decided by the exact sequence of swaps of libstdc++'s introsort (bits/stl_algo.h: __introsort_loop with
// __move_median_to_first + __unguarded_partition, threshold 16, __final_insertion_sort, heap sort on depth exhaustion).
// This header reproduces that sequence; tests/test_host_std_sort.py runs it on the host against libstdc++ itself.
//
// A record is one 64-bit word: curvature bits in the high half (curvature = d*d >= +0 and finite, so the unsigned order
// of the bits is the float order), point index in the low half; "less" looks at the high half only.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define LLB_HD
#else
#define LLB_HD __host__ __device__ __forceinline__
#endif

namespace llb {
namespace stdsort {

typedef unsigned long long rec_t;

LLB_HD bool less(rec_t a, rec_t b) { return (unsigned)(a >> 32) < (unsigned)(b >> 32); }
LLB_HD void swp(rec_t *a, rec_t *b) { rec_t t = *a; *a = *b; *b = t; }

LLB_HD void push_heap(rec_t *first, int hole, int top, rec_t value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

LLB_HD void adjust_heap(rec_t *first, int hole, int len, rec_t value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1])) child--;
        first[hole] = first[child]; hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1]; hole = child - 1;
    }
    push_heap(first, hole, top, value);
}

LLB_HD void heap_sort(rec_t *first, int len)
{   // std::__partial_sort(first, last, last): make_heap, then sort_heap
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            adjust_heap(first, parent, len, first[parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    while (len > 1) {
        --len;
        rec_t v = first[len]; first[len] = first[0];
        adjust_heap(first, 0, len, v);
    }
}

LLB_HD void median_to_first(rec_t *result, rec_t *a, rec_t *b, rec_t *c)
{
    if (less(*a, *b)) {
        if (less(*b, *c)) swp(result, b);
        else if (less(*a, *c)) swp(result, c);
        else swp(result, a);
    } else if (less(*a, *c)) swp(result, a);
    else if (less(*b, *c)) swp(result, c);
    else swp(result, b);
}

LLB_HD int unguarded_partition(rec_t *a, int first, int last, int pivot)
{
    const rec_t pv = a[pivot];                 // the pivot sits at `pivot` = first - 1 and is never swapped here
    for (;;) {
        while (less(a[first], pv)) ++first;
        --last;
        while (less(pv, a[last])) --last;
        if (!(first < last)) return first;
        swp(a + first, a + last);
        ++first;
    }
}

LLB_HD void unguarded_linear_insert(rec_t *a, int last)
{
    const rec_t val = a[last];
    int next = last - 1;
    while (less(val, a[next])) { a[last] = a[next]; last = next; --next; }
    a[last] = val;
}

LLB_HD void insertion_sort(rec_t *a, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(a[i], a[first])) {
            const rec_t val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];      // std::move_backward
            a[first] = val;
        } else unguarded_linear_insert(a, i);
    }
}

// std::sort(a, a + n); depth_limit < 0: 2 * floor(log2(n)) as std::sort sets it
LLB_HD void sort(rec_t *a, int n, int depth_limit = -1)
{
    if (n <= 0) return;
    if (depth_limit < 0) { int lg = 0; for (int m = n; m > 1; m >>= 1) lg++; depth_limit = 2 * lg; }
    // __introsort_loop: recursion on [cut, last), iteration on [first, cut) -> explicit stack of pending right parts
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    int first = 0, last = n, depth = depth_limit;
    for (;;) {
        while (last - first > 16) {
            if (depth == 0) { heap_sort(a + first, last - first); break; }
            --depth;
            const int mid = first + (last - first) / 2;
            median_to_first(a + first, a + first + 1, a + mid, a + last - 1);
            const int cut = unguarded_partition(a, first + 1, last, first);
            // the recursive call handles [cut, last) FIRST in libstdc++; the two ranges are disjoint, so the order in
            // which they are finished does not change the result: push the right part, continue with the left
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = cut;
        }
        if (sp == 0) break;
        sp--;
        first = stk_first[sp]; last = stk_last[sp]; depth = stk_depth[sp];
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) unguarded_linear_insert(a, i);
    } else insertion_sort(a, 0, n);
}

// ---- the same result from data-parallel steps (what the kernel's warps execute; this sequential form exists so that the
// ---- formulas can be checked on the host against libstdc++)
//
// __unguarded_partition in closed form.  With the pivot value pv, call position i a LEFT stopper if !(a[i] < pv) and a
// RIGHT stopper if !(pv < a[i]); L_0 < L_1 < ... are the left stoppers of [first, last) in increasing order, R_0 > R_1 >
// ... the right stoppers in decreasing order, both taken on the array as it is BEFORE the partition.  The sequential
// loop swaps L_k with R_k for k = 0 .. K-1, K = number of k with L_k < R_k (swaps only touch positions both pointers have
// passed, so the untouched stretch between them still shows the original stoppers), and returns
// cut = L_0 if K == 0, else min(L_K, R_{K-1}) (position R_{K-1} now holds a left stopper).
// Lbuf / Rbuf: scratch for the stopper positions.
LLB_HD int partition_closed(rec_t *a, int first, int last, int pivot, unsigned short *Lbuf, unsigned short *Rbuf)
{
    const rec_t pv = a[pivot];
    int nL = 0, nR = 0;
    for (int i = first; i < last; i++) if (!less(a[i], pv)) Lbuf[nL++] = (unsigned short)(i - first);
    for (int i = last - 1; i >= first; i--) if (!less(pv, a[i])) Rbuf[nR++] = (unsigned short)(i - first);
    int K = 0;
    const int m = nL < nR ? nL : nR;
    for (int k = 0; k < m; k++) if (Lbuf[k] < Rbuf[k]) K++;          // monotone: true for k < K only
    for (int k = 0; k < K; k++) swp(a + first + Lbuf[k], a + first + Rbuf[k]);
    if (K == 0) return first + Lbuf[0];
    const int rk = first + Rbuf[K - 1];
    return (K < nL && first + Lbuf[K] < rk) ? first + Lbuf[K] : rk;
}

// std::sort as: partitions (closed form) down to ranges of <= 16 records, then every such range sorted on its own.
// The final insertion sort of libstdc++ is a stable sort of the whole array; after the partitions the ranges are weakly
// ordered among themselves (left part <= pivot <= right part), so a stable sort of the whole = a stable sort of each.
LLB_HD void sort_closed(rec_t *a, int n, unsigned short *Lbuf, unsigned short *Rbuf, int depth_limit = -1)
{
    if (n <= 0) return;
    if (depth_limit < 0) { int lg = 0; for (int m = n; m > 1; m >>= 1) lg++; depth_limit = 2 * lg; }
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    int first = 0, last = n, depth = depth_limit;
    for (;;) {
        while (last - first > 16) {
            if (depth == 0) { heap_sort(a + first, last - first); last = first; break; }
            --depth;
            const int mid = first + (last - first) / 2;
            median_to_first(a + first, a + first + 1, a + mid, a + last - 1);
            const int cut = partition_closed(a, first + 1, last, first, Lbuf, Rbuf);
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = cut;
        }
        if (last - first > 1) insertion_sort(a, first, last);          // a leaf range
        if (sp == 0) break;
        sp--;
        first = stk_first[sp]; last = stk_last[sp]; depth = stk_depth[sp];
    }
}

}  // namespace stdsort
}  // namespace llb
