// host_builders_experiment.hpp -- TEST-ONLY experiment (tests/emul, RT_B200_BUILDER=sweep|sah): host-side hierarchy builders
// used to measure, on the CPU replay, how much traversal work a better tree than the Karras LBVH would save
// (profiles/r01_bvh_quality_experiment.txt: sweep +4.8 % node visits, binned SAH -7.5 % on C3 -- not adopted).  They
// produce exactly what karras_thread produces -- KarrasNode{left,right,lo,hi} over contiguous ranges of an ORDERED leaf
// array, parent arrays, root = node 0 -- so everything downstream (bottom-up refit with SAH leaf collapse, 4-wide BFS
// emission, traversal) is unchanged.  Any binary tree fits that form if the leaf order is its own depth-first order.
//
//   sweep : leaves stay in Morton order; every split is the best surface-area-heuristic cut ALONG that order
//   sah   : top-down binned SAH (16 bins x 3 axes) over the faces, leaf order = the tree's depth-first order
//
// Node numbering: the split after ordered leaf s (left = [lo..s], right = [s+1..hi]) gets id s -- every gap between two
// adjacent leaves is split exactly once in the whole tree, so ids are 0..n-2 -- with the root's id and 0 swapped.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../distributed-raytracer_b200/csrc/rt_bvh.cuh"

struct RtHostTree {
    KarrasNode *kn; int *parent_int, *parent_leaf; int root_s;
    int id_of(int s) const { return s == root_s ? 0 : (s == 0 ? root_s : s); }
    // records the node that splits [lo..hi] after s; returns its id
    int emit(int lo, int hi, int s, int parent, bool is_left) {
        const int id = id_of(s);
        if (parent >= 0) {
            parent_int[id] = parent;
            if (is_left) kn[parent].left = id; else kn[parent].right = id;
        } else parent_int[id] = -1;
        kn[id].lo = lo; kn[id].hi = hi;
        if (s == lo) { kn[id].left = ~lo; parent_leaf[lo] = id; }
        if (s + 1 == hi) { kn[id].right = ~hi; parent_leaf[hi] = id; }
        return id;
    }
};

// sorted_box[i] = box of the i-th leaf in Morton order.  n >= 2.
inline void rt_sweep_build_host(const BoxF *sorted_box, int n, KarrasNode *kn, int *parent_int, int *parent_leaf) {
    struct Item { int lo, hi, parent; bool is_left; };
    std::vector<Item> stack;
    std::vector<float> right_area((size_t)n);
    // work guard: equal boxes (duplicate faces) could make every split maximally lopsided and the sweeps quadratic; past
    // 64 n log2 n box unions the remaining nodes are split at the median
    double work = 0.0, budget = 64.0 * (double)n * (double)(64 - rt_clz64((unsigned long long)n));
    auto best_split = [&](int lo, int hi) {
        const int mid = lo + (hi - lo) / 2;
        work += (double)(hi - lo + 1);
        if (work > budget) return mid;
        BoxF acc = sorted_box[hi];
        right_area[hi] = box_half_area(acc);
        for (int k = hi - 1; k > lo; k--) { acc = box_union(acc, sorted_box[k]); right_area[k] = box_half_area(acc); }
        BoxF left = sorted_box[lo];
        float best = 0.0f;
        int best_s = lo;
        for (int s = lo; s < hi; s++) {                  // left = [lo..s], right = [s+1..hi]
            if (s > lo) left = box_union(left, sorted_box[s]);
            const float c = box_half_area(left) * (float)(s - lo + 1) + right_area[s + 1] * (float)(hi - s);
            const int dm = s > mid ? s - mid : mid - s, dbm = best_s > mid ? best_s - mid : mid - best_s;
            if (s == lo || c < best || (c == best && dm < dbm)) { best = c; best_s = s; }       // ties: the more balanced split
        }
        return best_s;
    };
    RtHostTree t{ kn, parent_int, parent_leaf, best_split(0, n - 1) };
    stack.push_back({ 0, n - 1, -1, false });
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const int s = (it.parent < 0) ? t.root_s : best_split(it.lo, it.hi);
        const int id = t.emit(it.lo, it.hi, s, it.parent, it.is_left);
        if (s > it.lo) stack.push_back({ it.lo, s, id, true });
        if (s + 1 < it.hi) stack.push_back({ s + 1, it.hi, id, false });
    }
}

// box[f] = box of face f (original order).  order[] receives the leaf order (a permutation of 0..n-1).  n >= 2.
inline void rt_sah_build_host(const BoxF *box, int n, uint32_t *order, KarrasNode *kn, int *parent_int, int *parent_leaf) {
    constexpr int NB = 16;
    for (int i = 0; i < n; i++) order[i] = (uint32_t)i;
    auto centroid = [&](uint32_t f, int a) { return 0.5f * (box[f].lo[a] + box[f].hi[a]); };
    // returns s in [lo, hi-1]: order[lo..s] left, order[s+1..hi] right (order is partitioned in place)
    auto split = [&](int lo, int hi) {
        const int cnt = hi - lo + 1;
        if (cnt == 2) return lo;
        float clo[3] = { HUGE_VALF, HUGE_VALF, HUGE_VALF }, chi[3] = { -HUGE_VALF, -HUGE_VALF, -HUGE_VALF };
        for (int k = lo; k <= hi; k++)
            for (int a = 0; a < 3; a++) { float c = centroid(order[k], a); clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c); }
        float best = HUGE_VALF; int best_axis = -1, best_bin = 0;
        for (int a = 0; a < 3; a++) {
            const float ext = chi[a] - clo[a];
            if (!(ext > 0.0f)) continue;
            BoxF bb[NB]; int bc[NB];
            for (int b = 0; b < NB; b++) { bc[b] = 0; for (int q = 0; q < 3; q++) { bb[b].lo[q] = HUGE_VALF; bb[b].hi[q] = -HUGE_VALF; } }
            const float scale = (float)NB / ext;
            for (int k = lo; k <= hi; k++) {
                const uint32_t f = order[k];
                int b = (int)((centroid(f, a) - clo[a]) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                bc[b]++; bb[b] = box_union(bb[b], box[f]);
            }
            float ra[NB]; int rc[NB];
            BoxF acc = bb[NB - 1]; int c = bc[NB - 1];
            ra[NB - 1] = c ? box_half_area(acc) : 0.0f; rc[NB - 1] = c;
            for (int b = NB - 2; b >= 1; b--) { acc = box_union(acc, bb[b]); c += bc[b]; ra[b] = c ? box_half_area(acc) : 0.0f; rc[b] = c; }
            BoxF l = bb[0]; int lc = bc[0];
            for (int b = 0; b < NB - 1; b++) {           // split between bin b and b+1
                if (b > 0) { l = box_union(l, bb[b]); lc += bc[b]; }
                if (lc == 0 || rc[b + 1] == 0) continue;
                const float cost = box_half_area(l) * (float)lc + ra[b + 1] * (float)rc[b + 1];
                if (cost < best) { best = cost; best_axis = a; best_bin = b; }
            }
        }
        if (best_axis < 0) return lo + (cnt - 1) / 2;    // all centroids equal: median of the current order
        const float scale = (float)NB / (chi[best_axis] - clo[best_axis]);
        auto mid = std::stable_partition(order + lo, order + hi + 1, [&](uint32_t f) {
            int b = (int)((centroid(f, best_axis) - clo[best_axis]) * scale);
            b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
            return b <= best_bin;
        });
        return (int)(mid - order) - 1;
    };
    struct Item { int lo, hi, parent; bool is_left; int s; };
    std::vector<Item> stack;
    const int root_s = split(0, n - 1);
    RtHostTree t{ kn, parent_int, parent_leaf, root_s };
    stack.push_back({ 0, n - 1, -1, false, root_s });
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const int s = it.parent < 0 ? it.s : split(it.lo, it.hi);
        const int id = t.emit(it.lo, it.hi, s, it.parent, it.is_left);
        if (s > it.lo) stack.push_back({ it.lo, s, id, true, 0 });
        if (s + 1 < it.hi) stack.push_back({ s + 1, it.hi, id, false, 0 });
    }
}
