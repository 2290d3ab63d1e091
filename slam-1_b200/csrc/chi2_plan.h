// chi2_plan.h -- the shape of numpy's pairwise summation for a float64 vector of k elements, laid out for the wide chi-square
// scan (bow.cu): leaves (runs of <= 128 elements) and the inner nodes of  sum(n) = sum(n2) + sum(n - n2),  n2 = n / 2 rounded
// down to a multiple of 8, sorted by height.  Plain C++ (no CUDA): tests/test_chi2_plan.py compiles it with g++ and checks a
// level-by-level evaluation against the recursion itself.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

struct Chi2Pair { int x, y; };      // same layout as CUDA's int2

struct Chi2Levels {
    int n_levels;
    int start[25];          // inner nodes [start[h], start[h + 1]) have height h + 1
};

struct Chi2Plan {
    // table[0 .. n_leaves)                  = (offset, length) of every leaf, in element order
    // table[n_leaves .. n_leaves + n_inner) = (left slot, right slot) of every inner node, sorted by height; value slot of
    //                                         leaf l is l, of inner node i is n_leaves + i; the root is the last one
    std::vector<Chi2Pair> table;
    int n_leaves = 0, n_inner = 0;
    Chi2Levels levels = {};
    bool ok = true;         // false: deeper than Chi2Levels can describe
};

inline Chi2Plan chi2_plan(int k)
{
    struct Inner { int left, right, height; };      // children as ids: >= 0 leaf, < 0 inner node ~id
    struct Build {
        std::vector<Chi2Pair> leaves;
        std::vector<Inner> inner;
        int run(int off, int n, int &height)
        {
            if (n <= 128) {
                leaves.push_back({off, n});
                height = 0;
                return (int)leaves.size() - 1;
            }
            int n2 = n / 2;
            n2 -= n2 % 8;
            int hl = 0, hr = 0;
            const int l = run(off, n2, hl);
            const int r = run(off + n2, n - n2, hr);
            height = (hl > hr ? hl : hr) + 1;
            inner.push_back({l, r, height});
            return ~((int)inner.size() - 1);
        }
    } b;
    int root_height = 0;
    b.run(0, k, root_height);
    Chi2Plan p;
    p.n_leaves = (int)b.leaves.size();
    p.n_inner = (int)b.inner.size();
    p.table = b.leaves;
    if (root_height > 24) {
        p.ok = false;
        return p;
    }
    // inner nodes by height (stable): position of creation-order node i in that order
    std::vector<int> order(b.inner.size()), pos(b.inner.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return b.inner[a].height < b.inner[c].height; });
    for (size_t i = 0; i < order.size(); ++i) pos[order[i]] = (int)i;
    auto slot_of = [&](int id) { return id >= 0 ? id : p.n_leaves + pos[~id]; };
    p.levels.n_levels = root_height;
    p.levels.start[0] = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        const Inner &nd = b.inner[order[i]];
        p.table.push_back({slot_of(nd.left), slot_of(nd.right)});
        p.levels.start[nd.height] = (int)i + 1;          // end of this level = start of the next
    }
    return p;
}
