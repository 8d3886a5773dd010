// dsc_device.cuh -- device-side building blocks of the DSC image-energy kernels.
//
// Arithmetic contract: everything here is FP64 with the reference's exact operation order and
// NO fused multiply-add (the library is compiled with -fmad=false): sample positions, volumes,
// areas, barycentric tests and z of a ray hit all feed integer decisions (voxel index, sample
// level, Gauss set, inside test, nearbyint) that must be bit-identical to the reference's x86-64
// build, which has no FMA contraction.  double division and sqrt are IEEE
// round-to-nearest on the device, so they agree as well.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define DSC_HD __host__ __device__ __forceinline__
#define DSC_D __device__ __forceinline__
#else
#define DSC_HD inline
#define DSC_D inline
#endif

namespace dsc {

struct V3 { double x, y, z; };

DSC_HD V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
DSC_HD V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
DSC_HD V3 mul(V3 a, double k) { return {a.x * k, a.y * k, a.z * k}; }
DSC_HD V3 divs(V3 a, double k) { return {a.x / k, a.y / k, a.z / k}; }
// cross product, operand order of CGLA/ArithVec3Float.h:46-52
DSC_HD V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// dot = std::inner_product from 0, left to right (CGLA/ArithVec.h:423-426); 0 + x is exact
DSC_HD double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
DSC_HD double length(V3 a) { return sqrt(dot(a, a)); }
DSC_HD V3 normalize(V3 a) { return divs(a, length(a)); }

// Util::signed_volume / volume / area (is_mesh/util.h:80-96)
DSC_HD double signed_volume(V3 a, V3 b, V3 c, V3 d) { return dot(sub(a, d), cross(sub(b, d), sub(c, d))) / 6.; }
DSC_HD double tet_volume(V3 a, V3 b, V3 c, V3 d) { return fabs(signed_volume(a, b, c, d)); }
DSC_HD double tri_area(V3 a, V3 b, V3 c) { return 0.5 * length(cross(sub(b, a), sub(c, a))); }

// Sample level of image3d::get_tetra_intensity (DEMO/image3d.cpp:355-357): upper_bound over
// {1,8,27,...,729} minus one, clamped at 0  =>  n with n^3 <= v < (n+1)^3, n in [1,9].
// Returns n-1.  NaN volume falls through every "<" and lands on the last level, as upper_bound does.
DSC_HD int tet_level(double v)
{
    int idx = 0;
#pragma unroll
    for (int n = 1; n <= 9; n++) {
        if (v < (double)(n * n * n)) break;
        idx++;
    }
    idx -= 1;
    return idx < 0 ? 0 : idx;
}

// Gauss set of gaussian_points::get_point(double area) (DEMO/gaussian_points.h:97-101).  The
// reference indexes past its 7 sets for area >= 999999 (latent UB); clamped to the last set.
DSC_HD int gauss_set(double a)
{
    const double thres[7] = {6, 10, 20, 30, 50, 80, 999999};
    int r = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) {
        if (a < thres[k]) break;
        r++;
    }
    return r < 7 ? r : 6;
}

// ---- z-ray hit sort: libstdc++'s std::sort restated ------------------------------------------
// A hit packs (z << 1) | b_in in one int32.  The reference sorts with a comparator that looks at z
// only (DEMO/segment_function.cpp:1595-1598, 1703), so elements with equal z are "equivalent" and
// their final order is whatever libstdc++'s introsort produces; the duplicate-erase loop that
// follows depends on that order.  This is the same algorithm, step for step:
// introsort loop with depth limit 2*floor(log2 n), median-of-three moved to first, unguarded
// partition, threshold 16, heap-sort fallback, final (un)guarded insertion sort.
DSC_HD bool hit_less(int a, int b) { return (a >> 1) < (b >> 1); }

DSC_HD void hs_swap(int *v, int i, int j) { int t = v[i]; v[i] = v[j]; v[j] = t; }

DSC_HD void hs_push_heap(int *v, int first, int hole, int top, int value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && hit_less(v[first + parent], value)) {
        v[first + hole] = v[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[first + hole] = value;
}

DSC_HD void hs_adjust_heap(int *v, int first, int hole, int len, int value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (hit_less(v[first + child], v[first + child - 1])) child--;
        v[first + hole] = v[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[first + hole] = v[first + child - 1];
        hole = child - 1;
    }
    hs_push_heap(v, first, hole, top, value);
}

// std::__partial_sort(first, last, last): make_heap then sort_heap (heap_select's loop is empty)
DSC_HD void hs_heap_sort(int *v, int first, int last)
{
    int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            int value = v[first + parent];
            hs_adjust_heap(v, first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        int value = v[last];
        v[last] = v[first];
        hs_adjust_heap(v, first, 0, last - first, value);
    }
}

DSC_HD void hs_unguarded_linear_insert(int *v, int last)
{
    int val = v[last];
    int next = last - 1;
    while (hit_less(val, v[next])) {
        v[last] = v[next];
        last = next;
        --next;
    }
    v[last] = val;
}

DSC_HD void hs_insertion_sort(int *v, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (hit_less(v[i], v[first])) {
            int val = v[i];
            for (int j = i; j > first; --j) v[j] = v[j - 1]; // move_backward(first, i, i+1)
            v[first] = val;
        } else {
            hs_unguarded_linear_insert(v, i);
        }
    }
}

DSC_HD void std_sort_hits(int *v, int n)
{
    if (n <= 0) return;
    // __introsort_loop with an explicit stack for the recursion on the right part; ranges are
    // disjoint, so the order in which they are finished does not change the result.
    int stk_first[40], stk_last[40], stk_depth[40];
    int sp = 0;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                hs_heap_sort(v, first, last);
                break;
            }
            --depth;
            // __unguarded_partition_pivot
            int mid = first + (last - first) / 2;
            {   // __move_median_to_first(first, first+1, mid, last-1)
                int a = first + 1, b = mid, c = last - 1;
                if (hit_less(v[a], v[b])) {
                    if (hit_less(v[b], v[c])) hs_swap(v, first, b);
                    else if (hit_less(v[a], v[c])) hs_swap(v, first, c);
                    else hs_swap(v, first, a);
                } else if (hit_less(v[a], v[c])) hs_swap(v, first, a);
                else if (hit_less(v[b], v[c])) hs_swap(v, first, c);
                else hs_swap(v, first, b);
            }
            int lo = first + 1, hi = last;
            while (true) { // __unguarded_partition(first+1, last, pivot = first)
                while (hit_less(v[lo], v[first])) ++lo;
                --hi;
                while (hit_less(v[first], v[hi])) --hi;
                if (!(lo < hi)) break;
                hs_swap(v, lo, hi);
                ++lo;
            }
            int cut = lo;
            if (sp < 40) { stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++; }
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        hs_insertion_sort(v, 0, 16);
        for (int i = 16; i != n; ++i) hs_unguarded_linear_insert(v, i);
    } else {
        hs_insertion_sort(v, 0, n);
    }
}

// Adjacent-duplicate removal with the reference's skip-after-erase behaviour
// (DEMO/segment_function.cpp:1706-1719).  Returns the new length.
DSC_HD int erase_adjacent_duplicates(int *v, int n)
{
    int p = 1;
    while (p < n) {
        if (v[p] == v[p - 1]) { // same z and same b_in  <=>  same packed word
            for (int j = p; j + 1 < n; j++) v[j] = v[j + 1];
            n--;
            if (p == n) break;
        }
        p++;
    }
    return n;
}

} // namespace dsc
