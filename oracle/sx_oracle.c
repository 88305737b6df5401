/*
 * sx_oracle.c -- CPU restatement (plain C) of the sequential parts of the
 * reference's network-crossover hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker*: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (smart-crossover_b200/) never links or calls it.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/src/smart_crossover/).  The arithmetic of the reference
 * lives in NumPy/SciPy (numpy pinned 1.21.3, scipy pinned 1.7.3 in the
 * reference's environment.yml:10-12; 2.3.5 / 1.18.1 installed here); what is
 * restated below is the published algorithm of those routines, pinned by the
 * golden vectors in tests/golden/ (generated from the reference itself by
 * tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int64_t uf_find(int64_t *parent, int64_t x)
{
    while (parent[x] != x) {
        parent[x] = parent[parent[x]];
        x = parent[x];
    }
    return x;
}

/*
 * Kruskal over a pre-sorted arc order.
 *
 * Restates network_methods/tree_BI.py:32-59 -> scipy.sparse.csgraph.
 * minimum_spanning_tree on the negated weights: edges are visited in the order
 * of a STABLE argsort of -w over CSR order (= arc id order for the complete
 * bipartite OT graph, tree_BI.py:45-50), an edge is kept when its endpoints are
 * in different components, and the scan stops once N-1 edges are kept.
 *
 * Endpoints: if tail == NULL the OT convention is used, arc k = (k / D, S + k % D)
 * (formats.py:156-160, net_manager.py:366); otherwise tail[k], head[k].
 * Returns the number of tree arcs written to tree_out (unsorted, acceptance order).
 */
int64_t sxo_kruskal(const int64_t *order, int64_t n_order,
                    const int64_t *tail, const int64_t *head,
                    int64_t S, int64_t D, int64_t N, int64_t *tree_out)
{
    int64_t *parent = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
    int64_t cnt = 0;
    if (!parent) return -1;
    for (int64_t i = 0; i < N; ++i) parent[i] = i;
    for (int64_t p = 0; p < n_order && cnt < N - 1; ++p) {
        int64_t k = order[p];
        int64_t a = tail ? tail[k] : k / D;
        int64_t b = head ? head[k] : S + k % D;
        int64_t ra = uf_find(parent, a), rb = uf_find(parent, b);
        if (ra != rb) {
            parent[ra] = rb;
            tree_out[cnt++] = k;
        }
    }
    free(parent);
    return cnt;
}

/*
 * Node potentials of a spanning tree.
 *
 * the reference takes duals from the LP solver
 * (solver_caller/gurobi.py:157-159 used at network_methods/algorithms.py:132);
 * the restated definition is  B^T y[:-1] = c[tree],  y[root] = 0  with
 * B = A[:-1, tree] (tree_BI.py:74).  Each column of A has +1 at node `plus`
 * and -1 at node `minus`, so every tree arc gives y[plus] - y[minus] = cost.
 *   OT  (formats.py:156-158): plus = S + j, minus = i.
 *   MCF (scripts/min2mcf.py:36-37): plus = tail, minus = head.
 * Solved by a breadth-first walk from the root.  Returns 0, or -2 when the
 * arcs do not form a spanning tree of the N nodes.
 */
int sxo_tree_potentials(const int64_t *plus, const int64_t *minus,
                        const double *cost, int64_t T, int64_t N,
                        int64_t root, double *y)
{
    if (T != N - 1) return -2;
    int64_t *deg = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    int64_t *adj = (int64_t *)malloc(sizeof(int64_t) * (size_t)(2 * T + 1));
    int64_t *queue = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
    char *seen = (char *)calloc((size_t)N, 1);
    int rc = 0;
    for (int64_t t = 0; t < T; ++t) { deg[plus[t] + 1]++; deg[minus[t] + 1]++; }
    for (int64_t i = 0; i < N; ++i) deg[i + 1] += deg[i];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
    memcpy(fill, deg, sizeof(int64_t) * (size_t)N);
    for (int64_t t = 0; t < T; ++t) { adj[fill[plus[t]]++] = t; adj[fill[minus[t]]++] = t; }
    int64_t qh = 0, qt = 0;
    y[root] = 0.0; seen[root] = 1; queue[qt++] = root;
    while (qh < qt) {
        int64_t v = queue[qh++];
        for (int64_t q = deg[v]; q < deg[v + 1]; ++q) {
            int64_t t = adj[q];
            int64_t w;
            double yw;
            if (plus[t] == v) { w = minus[t]; yw = y[v] - cost[t]; }   /* y[minus] = y[plus] - c */
            else              { w = plus[t];  yw = y[v] + cost[t]; }   /* y[plus]  = y[minus] + c */
            if (seen[w]) continue;
            seen[w] = 1; y[w] = yw; queue[qt++] = w;
        }
    }
    if (qt != N) rc = -2;
    free(deg); free(adj); free(queue); free(seen); free(fill);
    return rc;
}

/*
 * Per-node flow sums for the MCF flow indicators.
 *
 * Restates network_methods/net_manager.py:171-176:  f_1 = A_barplus @ x_hat,
 * f_2 = A_barminus @ x_hat with SciPy's csr_matvec, i.e. for every node a
 * sequential sum starting at 0.0 over its incident arcs in ascending arc id.
 * node_ptr/node_arc/node_val are the CSR arrays of A (formats.py:118); flip[k]
 * is 1 where the arc was reversed (x > u/2, net_manager.py:166,169).
 */
void sxo_mcf_node_sums(const int64_t *node_ptr, const int64_t *node_arc,
                       const double *node_val, const unsigned char *flip,
                       const double *x_hat, int64_t N, double *f1, double *f2)
{
    for (int64_t v = 0; v < N; ++v) {
        double a = 0.0, b = 0.0;
        for (int64_t q = node_ptr[v]; q < node_ptr[v + 1]; ++q) {
            int64_t k = node_arc[q];
            double s = flip[k] ? -node_val[q] : node_val[q];
            if (s > 0) a += s * x_hat[k];
            else if (s < 0) b += (-s) * x_hat[k];
        }
        f1[v] = a; f2[v] = b;
    }
}

/*
 * Tree primal flows + push to a basic feasible solution (OT only).
 *
 * Restates network_methods/tree_BI.py:62-114.  `flow` is the dense S x D
 * scratch (tree_BI.py:77-79) already holding the tree solution; the loop below
 * is tree_BI.py:81-110 verbatim in C: negative entries are visited in row-major
 * order (np.where), J2/I2 are first-occurrence argmax over the row / column.
 * Returns push_iter, or -1 if one of the reference's asserts (tree_BI.py:93-94)
 * would fire.
 */
static int64_t argmax_row(const double *flow, int64_t D, int64_t i)
{
    int64_t best = 0; const double *r = flow + i * D;
    for (int64_t j = 1; j < D; ++j) if (r[j] > r[best]) best = j;
    return best;
}
static int64_t argmax_col(const double *flow, int64_t S, int64_t D, int64_t j)
{
    int64_t best = 0;
    for (int64_t i = 1; i < S; ++i) if (flow[i * D + j] > flow[best * D + j]) best = i;
    return best;
}
int64_t sxo_push_tree(double *flow, int64_t S, int64_t D)
{
    int64_t n = S * D, nneg = 0, push_iter = 0;
    int64_t *neg = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t k = 0; k < n; ++k) if (flow[k] < 0) neg[nneg++] = k;
    for (int64_t q = 0; q < nneg; ++q) {
        int64_t I1 = neg[q] / D, J1 = neg[q] % D;
        int64_t J2 = argmax_row(flow, D, I1);
        int64_t I2 = argmax_col(flow, S, D, J1);
        while (flow[I1 * D + J1] < 0) {
            if (!(flow[I2 * D + J1] > 0 && flow[I1 * D + J2] > 0)) { free(neg); return -1; }
            if (!(flow[I2 * D + J2] == 0)) { free(neg); return -1; }
            double c0 = -flow[I1 * D + J1], c1 = flow[I1 * D + J2], c2 = flow[I2 * D + J1];
            double theta = c0; int flag = 0;              /* np.min / np.argmin: first minimum */
            if (c1 < theta) { theta = c1; flag = 1; }
            if (c2 < theta) { theta = c2; flag = 2; }
            flow[I1 * D + J1] += theta;
            flow[I2 * D + J1] -= theta;
            flow[I1 * D + J2] -= theta;
            flow[I2 * D + J2] += theta;
            if (flag == 1) J2 = argmax_row(flow, D, I1);
            else if (flag == 2) I2 = argmax_col(flow, S, D, J1);
            push_iter++;
        }
    }
    free(neg);
    return push_iter;
}
