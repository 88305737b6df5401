// sx_push.cu -- f1 (host part): the push phase of tree basis identification on the sparse support.
//
// Replaces the loop of `push_tree_to_bfs` (reference tree_BI.py:81-113), which keeps the tree solution in
// a dense S x D array and takes np.argmax over whole rows and columns.  The flows live only on the tree
// arcs and on the 4-cycle corners the pushes create, so this walks per-row / per-column entry lists:
// O(N + pushes * degree) work and memory instead of O(S * D).  It is sequential by nature (every push
// depends on the previous one) and runs on the host; the tree primal flows it starts from come from the
// device (sx_tree_flows).  Arithmetic is the reference's, operation for operation: theta is the first
// minimum of (-x[I1,J1], x[I1,J2], x[I2,J1]), the four corners are updated with one add / subtract each,
// np.argmax = first index of the (positive) maximum.
#include <algorithm>
#include <unordered_map>
#include <vector>

#include "sx_common.cuh"

namespace {

struct PushEntry {
    int    i, j;
    double v;
};

struct PushState {
    long long                          D;
    std::vector<PushEntry>             entries;
    std::unordered_map<long long, int> index;          // arc id -> entry
    std::vector<std::vector<int>>      row, col;       // entry numbers per row / per column

    double get(int i, int j) const {
        auto it = index.find((long long)i * D + j);
        return it == index.end() ? 0.0 : entries[it->second].v;
    }
    void put(int i, int j, double v) {
        const long long key = (long long)i * D + j;
        auto it = index.find(key);
        if (it != index.end()) { entries[it->second].v = v; return; }
        const int e = (int)entries.size();
        entries.push_back({i, j, v});
        index.emplace(key, e);
        row[i].push_back(e);
        col[j].push_back(e);
    }
    // np.argmax over the dense row (by_row) or column: first index of the maximum, which must be positive
    // (the reference asserts it right after, tree_BI.py:93).  -1 if there is no positive entry.
    int argmax(const std::vector<int> &list, bool by_row) const {
        int    best = -1;
        double best_val = 0.0;
        for (int e : list) {
            const int    idx = by_row ? entries[e].j : entries[e].i;
            const double val = entries[e].v;
            if (val > best_val || (val == best_val && best >= 0 && idx < best)) { best = idx; best_val = val; }
        }
        return best;
    }
};

}  // namespace

extern "C" int sx_push_tree_h(const int64_t *tree_h, const double *flow_h, int64_t n_tree, int64_t S, int64_t D,
                              int64_t *pos_arc_h, int64_t cap, int64_t *n_pos_h, int64_t *push_iter_h) {
    if (n_tree < 0 || S <= 0 || D <= 0 || cap < 0 || !n_pos_h || !push_iter_h) return SX_ERR_INVALID;
    if (n_tree > 0 && (!tree_h || !flow_h)) return SX_ERR_INVALID;
    if (cap > 0 && !pos_arc_h) return SX_ERR_INVALID;
    if (S >= (1ll << 31) || D >= (1ll << 31)) return SX_ERR_TOO_LARGE;
    try {
    PushState st;
    st.D = D;
    st.row.resize((size_t)S);
    st.col.resize((size_t)D);
    st.entries.reserve((size_t)n_tree * 2);
    st.index.reserve((size_t)n_tree * 2);
    std::vector<long long> negative;
    for (int64_t t = 0; t < n_tree; ++t) {
        const long long k = tree_h[t];
        if (k < 0 || k >= S * D) return SX_ERR_INVALID;
        st.put((int)(k / D), (int)(k % D), flow_h[t]);
        if (flow_h[t] < 0) negative.push_back(k);
    }
    // np.where(tree_solution < 0) lists the negative flows in row-major order, fixed before any push
    std::sort(negative.begin(), negative.end());
    long long pushes = 0;
    for (long long k : negative) {
        const int I1 = (int)(k / D), J1 = (int)(k % D);
        if (st.get(I1, J1) >= 0) continue;               // an earlier push already repaired it
        int J2 = st.argmax(st.row[I1], true);
        int I2 = st.argmax(st.col[J1], false);
        while (st.get(I1, J1) < 0) {
            if (J2 < 0 || I2 < 0) return SX_ERR_PUSH_ASSERT;
            const double x11 = st.get(I1, J1), x12 = st.get(I1, J2), x21 = st.get(I2, J1), x22 = st.get(I2, J2);
            if (!(x21 > 0 && x12 > 0) || x22 != 0) return SX_ERR_PUSH_ASSERT;      // tree_BI.py:93-94
            const double cand[3] = {-x11, x12, x21};
            int flag = 0;                                 // np.argmin: first minimum
            if (cand[1] < cand[flag]) flag = 1;
            if (cand[2] < cand[flag]) flag = 2;
            const double theta = cand[flag];
            st.put(I1, J1, x11 + theta);
            st.put(I2, J1, x21 - theta);
            st.put(I1, J2, x12 - theta);
            st.put(I2, J2, x22 + theta);
            if (flag == 1) J2 = st.argmax(st.row[I1], true);
            else if (flag == 2) I2 = st.argmax(st.col[J1], false);
            ++pushes;
        }
    }
    long long n_pos = 0;
    for (const PushEntry &e : st.entries)
        if (e.v > 0) {
            if (n_pos < cap) pos_arc_h[n_pos] = (long long)e.i * D + e.j;
            ++n_pos;
        }
    *n_pos_h = n_pos;
    *push_iter_h = pushes;
    return n_pos <= cap ? SX_OK : SX_ERR_WORKSPACE;
    } catch (...) {          // host allocation failure: nothing may propagate through the C ABI
        return SX_ERR_TOO_LARGE;
    }
}
