// Replay of libstdc++'s std::sort (introsort: median-of-3 quicksort to depth 2*lg(n), heapsort
// fallback, threshold-16 final insertion sort) on 32-bit packed elements.
//
// Why: the reference orders a tile's corners with std::sort on the response alone
// (lvt/src/lvt_image_features_handler.cpp:38-41).  Responses are small integers, ties are
// everywhere, std::sort is unstable, and the resulting permutation is the order in which the
// survivors are emitted (:72-80) -- i.e. it fixes every feature index downstream.  To be
// index-exact the GPU path has to walk the same comparison sequence, so this is the same
// algorithm (GCC 13 bits/stl_algo.h, bits/stl_heap.h) on elements (response << 24 | payload)
// compared by the response byte only, "greater" first.
//
// Sequential by nature: one thread runs it per tile while the rest of the CTA computes the
// suppression radii.  __host__ __device__ so that tests can check it against std::sort.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LVT_HD __host__ __device__
#else
#define LVT_HD
#endif

namespace lvtb
{
namespace isort
{

// comp(a, b): a sorts before b  <=>  response(a) > response(b)
LVT_HD inline bool before(uint32_t a, uint32_t b) { return (a >> 24) > (b >> 24); }

LVT_HD inline void swap_el(uint32_t *a, uint32_t *b)
{
    const uint32_t t = *a;
    *a = *b;
    *b = t;
}

LVT_HD inline void move_median_to_first(uint32_t *result, uint32_t *a, uint32_t *b, uint32_t *c)
{
    if (before(*a, *b))
    {
        if (before(*b, *c))
            swap_el(result, b);
        else if (before(*a, *c))
            swap_el(result, c);
        else
            swap_el(result, a);
    }
    else if (before(*a, *c))
        swap_el(result, a);
    else if (before(*b, *c))
        swap_el(result, c);
    else
        swap_el(result, b);
}

LVT_HD inline uint32_t *unguarded_partition(uint32_t *first, uint32_t *last, uint32_t *pivot)
{
    while (true)
    {
        while (before(*first, *pivot))
            ++first;
        --last;
        while (before(*pivot, *last))
            --last;
        if (!(first < last))
            return first;
        swap_el(first, last);
        ++first;
    }
}

LVT_HD inline void push_heap(uint32_t *first, long hole, long top, uint32_t value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && before(first[parent], value))
    {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

LVT_HD inline void adjust_heap(uint32_t *first, long hole, long len, uint32_t value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (before(first[child], first[child - 1]))
            child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap(first, hole, top, value);
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
LVT_HD inline void heap_sort(uint32_t *first, uint32_t *last)
{
    const long len = last - first;
    if (len >= 2)
    {
        long parent = (len - 2) / 2;
        while (true)
        {
            const uint32_t v = first[parent];
            adjust_heap(first, parent, len, v);
            if (parent == 0)
                break;
            parent--;
        }
    }
    while (last - first > 1)
    {
        --last;
        const uint32_t v = *last;
        *last = *first;
        adjust_heap(first, 0, last - first, v);
    }
}

LVT_HD inline void unguarded_linear_insert(uint32_t *last)
{
    const uint32_t val = *last;
    uint32_t *next = last - 1;
    while (before(val, *next))
    {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

LVT_HD inline void insertion_sort(uint32_t *first, uint32_t *last)
{
    if (first == last)
        return;
    for (uint32_t *i = first + 1; i != last; ++i)
    {
        if (before(*i, *first))
        {
            const uint32_t val = *i;
            for (uint32_t *p = i; p != first; --p)
                *p = *(p - 1);
            *first = val;
        }
        else
            unguarded_linear_insert(i);
    }
}

// std::sort(first, first + n, before)
LVT_HD inline void sort(uint32_t *first, long n)
{
    if (n <= 0)
        return;
    uint32_t *last = first + n;
    int lg = 0;
    for (long v = n; v > 1; v >>= 1)
        lg++;
    // explicit stack replaces the recursion on the right-hand part; sub-ranges are disjoint,
    // so the order in which they are processed does not change the result
    struct Range
    {
        uint32_t *first, *last;
        int depth;
    };
    Range stack[64];
    int sp = 0;
    stack[sp++] = Range{first, last, 2 * lg};
    while (sp > 0)
    {
        Range r = stack[--sp];
        while (r.last - r.first > 16)
        {
            if (r.depth == 0)
            {
                heap_sort(r.first, r.last);
                break;
            }
            --r.depth;
            uint32_t *mid = r.first + (r.last - r.first) / 2;
            move_median_to_first(r.first, r.first + 1, mid, r.last - 1);
            uint32_t *cut = unguarded_partition(r.first + 1, r.last, r.first);
            stack[sp++] = Range{cut, r.last, r.depth};
            r.last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16)
    {
        insertion_sort(first, first + 16);
        for (uint32_t *i = first + 16; i != last; ++i)
            unguarded_linear_insert(i);
    }
    else
        insertion_sort(first, last);
}

// The same sort as a level-synchronous schedule: the sub-ranges produced by a partition are disjoint
// and each carries its own depth budget, so all ranges of one recursion level can be partitioned
// concurrently; the final insertion sort never moves an element out of its <= 16-element leaf range
// (partitioning leaves keys(left) >= pivot >= keys(right), and insertion is stable), so it is an
// independent stable insertion sort per leaf.  This sequential version documents and tests the
// schedule; tile_kernel runs it with one thread per range.
struct LevelRange
{
    int first, last, depth;
};

// one step of __introsort_loop on r (size > 16): returns the two children, or heap-sorts in place
// and returns false when the depth budget is exhausted
LVT_HD inline bool split_range(uint32_t *a, const LevelRange &r, LevelRange &left, LevelRange &right)
{
    if (r.depth == 0)
    {
        heap_sort(a + r.first, a + r.last);
        return false;
    }
    uint32_t *first = a + r.first, *last = a + r.last;
    uint32_t *mid = first + (last - first) / 2;
    move_median_to_first(first, first + 1, mid, last - 1);
    uint32_t *cut = unguarded_partition(first + 1, last, first);
    left = LevelRange{r.first, (int)(cut - a), r.depth - 1};
    right = LevelRange{(int)(cut - a), r.last, r.depth - 1};
    return true;
}

// std::__unguarded_partition(first + 1, last, first) without the two-pointer loop (sequential
// statement of detect.cu's warp_partition; see there for the argument).  posL / posR: scratch.
LVT_HD inline int partition_lists(uint32_t *a, int first, int last, uint16_t *posL, uint16_t *posR)
{
    const uint32_t pv = a[first] >> 24;
    const int lo = first + 1, m = last - lo;
    int nL = 0, nR = 0;
    for (int p = 0; p < m; p++)
        if ((a[lo + p] >> 24) <= pv)
            posL[nL++] = (uint16_t)p;
    for (int p = m - 1; p >= 0; p--)
        if ((a[lo + p] >> 24) >= pv)
            posR[nR++] = (uint16_t)p;
    const int nmin = nL < nR ? nL : nR;
    int K = 0;
    while (K < nmin && posL[K] < posR[K])
    {
        swap_el(a + lo + posL[K], a + lo + posR[K]);
        K++;
    }
    const int prevR = K > 0 ? (int)posR[K - 1] : m;
    return (K < nL && (int)posL[K] < prevR) ? lo + posL[K] : lo + prevR;
}

// q0, q1: scratch for n / 8 + 2 ranges each; ranges of at least coop_min elements use partition_lists
LVT_HD inline void sort_levels(uint32_t *a, int n, LevelRange *q0, LevelRange *q1, uint16_t *posL = nullptr,
                               uint16_t *posR = nullptr, int coop_min = 1 << 30)
{
    if (n <= 1)
        return;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1)
        lg++;
    LevelRange *cur = q0, *nxt = q1;
    int ncur = 1;
    cur[0] = LevelRange{0, n, 2 * lg};
    while (ncur > 0)
    {
        int nnxt = 0;
        for (int i = 0; i < ncur; i++)
        {
            const LevelRange r = cur[i];
            if (r.last - r.first <= 16)
            {
                insertion_sort(a + r.first, a + r.last); // leaf
                continue;
            }
            LevelRange l, rr;
            if (r.last - r.first >= coop_min && r.depth != 0)
            {
                uint32_t *f = a + r.first, *e = a + r.last;
                move_median_to_first(f, f + 1, f + (e - f) / 2, e - 1);
                const int cut = partition_lists(a, r.first, r.last, posL, posR);
                nxt[nnxt++] = LevelRange{r.first, cut, r.depth - 1};
                nxt[nnxt++] = LevelRange{cut, r.last, r.depth - 1};
            }
            else if (split_range(a, r, l, rr))
            {
                nxt[nnxt++] = l;
                nxt[nnxt++] = rr;
            }
        }
        LevelRange *t = cur;
        cur = nxt;
        nxt = t;
        ncur = nnxt;
    }
}

} // namespace isort
} // namespace lvtb
