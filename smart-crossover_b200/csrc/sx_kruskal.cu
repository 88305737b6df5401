// sx_kruskal.cu -- K2: spanning-tree basis identification (sm_100a).
//
// Replaces `sp.csgraph.minimum_spanning_tree(-w)` + `np.flatnonzero` of the reference
// (tree_BI.py:32-59): Kruskal over the arcs in `korder` (descending weight, ties by ascending
// arc id), keep an arc when its end nodes are in different components, stop at N-1 arcs.
//
// Parallel but bit-exact: the position in `korder` is a strict total order, so the maximum
// spanning forest is unique and any correct algorithm returns the reference's arc set
// (SURVEY.md H2).  One persistent cooperative kernel walks `korder` in growing chunks (N/2 arcs first at
// large N, doubling); per chunk
//   phase A  filter: drop arcs whose ends already share a root (lock-free union-find, path
//            halving), compact the survivors, and let every root take the atomicMin of the
//            survivor positions incident to it (epoch-tagged so the table never needs a reset);
//   phase B  hook: an arc that is the minimum of one of its roots is a tree arc (cut property
//            on the contracted graph); hook that root under the other (smaller id wins when
//            both picked the same arc) and record the arc;
// and repeats until the chunk has no survivor.  Work is atomics / L2 bound, not HBM bound.
#include "sx_common.cuh"
#include "sx_gridbar.cuh"

namespace sx {

constexpr int kKrThreads = 256;
constexpr unsigned long long kEpochMax = (1ull << 24) - 1;

struct KrParams {
    const uint32_t *korder;
    long long       n;
    const int32_t  *tail, *head;
    long long       S, D, N;
    int            *parent;               // N
    unsigned long long *best;             // N, epoch-tagged minimum survivor position per root
    uint32_t       *list[2];              // survivor positions (capacity list_cap)
    int2           *roots[2];             // (root_u, root_v) of each survivor
    long long       list_cap;
    long long       first_chunk;          // arcs of the first chunk (doubling afterwards)
    long long      *tree_out;             // capacity N-1, pre-filled with a sentinel by the host
    unsigned long long *ctr;              // [0] tree count, [1],[2] survivor counts of list 0/1
};

// SINGLE: the whole forest lives in the shared memory of one CTA (plain loads); otherwise in global memory,
// read through L2 because other SMs rewrite it.
template <bool SINGLE>
__device__ __forceinline__ int kr_find(int *parent, int x) {
    for (;;) {
        const int p = SINGLE ? parent[x] : __ldcg(parent + x);
        if (p == x) return x;
        const int gp = SINGLE ? parent[p] : __ldcg(parent + p);
        if (gp == p) return p;
        parent[x] = gp;   // path halving; racing writers only ever store an ancestor
        x = gp;
    }
}

__device__ __forceinline__ void kr_endpoints(const KrParams &p, uint32_t e, int &u, int &v) {
    if (p.tail) { u = p.tail[e]; v = p.head[e]; }
    else { const long long i = (long long)e / p.D; u = (int)i; v = (int)(p.S + ((long long)e - i * p.D)); }
}

template <bool SINGLE>
__device__ __forceinline__ void kr_sync(GridBarrier *bar) {
    if (SINGLE) __syncthreads(); else grid_barrier(bar);
}

// The whole algorithm; `parent` / `best` are the forest and the per-root proposals (global memory, or the
// shared memory of the only CTA when SINGLE: every barrier is then a __syncthreads and every find a few
// shared-memory reads -- the 784 x 784 tree drops from 0.18 ms of grid-wide barriers to a few tens of us).
template <bool SINGLE>
__device__ __forceinline__ unsigned long long kr_ld(const unsigned long long *q) {
    return SINGLE ? *reinterpret_cast<const volatile unsigned long long *>(q) : __ldcg(q);
}

// The two survivor lists (position in the order + the roots found for its ends) of a run.
struct KrLists {
    uint32_t *q[2];
    int2     *roots[2];
    long long cap;
};

// `ctr`: [0] tree count, [1], [2] survivor counts of the two lists (shared memory when SINGLE).
// LSMEM: the survivor lists live in shared memory too (plain loads).
template <bool SINGLE, bool LSMEM>
__device__ __forceinline__ void kruskal_body(const KrParams &p, int *parent, unsigned long long *best,
                                             unsigned long long *ctr, const KrLists &L) {
    GridBarrier *bar = reinterpret_cast<GridBarrier *>(p.ctr + 4);
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz  = (long long)gridDim.x * blockDim.x;
    for (long long v = gtid; v < p.N; v += gsz) { parent[v] = (int)v; best[v] = ~0ull; }
    kr_sync<SINGLE>(bar);

    long long pos = 0;
    long long csize = p.first_chunk;
    if (csize > L.cap) csize = L.cap;
    unsigned long long epoch = 0;
    const unsigned long long want = (unsigned long long)(p.N - 1);
    bool done = false;

    while (pos < p.n && !done) {
        long long cend = pos + csize;
        if (cend > p.n) cend = p.n;
        bool first = true;
        int cur = 0;                       // list[cur] holds the survivors of the previous round
        for (;;) {
            ++epoch;
            const unsigned long long tag = (kEpochMax - epoch) << 40;
            const int nxt = cur ^ 1;
            // (selected without indexing the pointer tables, which would put them in local memory)
            uint32_t *q_cur = cur ? L.q[1] : L.q[0], *q_nxt = cur ? L.q[0] : L.q[1];
            int2 *r_cur = cur ? L.roots[1] : L.roots[0], *r_nxt = cur ? L.roots[0] : L.roots[1];
            const long long cnt_in = first ? (cend - pos) : (long long)kr_ld<SINGLE>(&ctr[1 + cur]);
            // ---- phase A: filter + propose ----
            for (long long base = (long long)blockIdx.x * blockDim.x; base < cnt_in; base += gsz) {
                const long long idx = base + threadIdx.x;
                bool alive = false;
                uint32_t q = 0;
                int ru = 0, rv = 0;
                if (idx < cnt_in) {
                    int u, v;
                    if (first) {
                        q = (uint32_t)(pos + idx);
                        kr_endpoints(p, p.korder[q], u, v);
                    } else {
                        // a survivor: continue from the roots found last round (same components, no second
                        // trip to the order and the arc's end points)
                        q = LSMEM ? q_cur[idx] : __ldcg(&q_cur[idx]);
                        const int2 r = LSMEM ? r_cur[idx] : __ldcg(&r_cur[idx]);
                        u = r.x; v = r.y;
                    }
                    ru = kr_find<SINGLE>(parent, u);
                    rv = kr_find<SINGLE>(parent, v);
                    alive = ru != rv;
                }
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                if (m) {
                    unsigned long long slot0 = 0;
                    if (lane_id() == (unsigned)(__ffs(m) - 1))
                        slot0 = atomicAdd(&ctr[1 + nxt], (unsigned long long)__popc(m));
                    slot0 = __shfl_sync(0xffffffffu, slot0, __ffs(m) - 1);
                    if (alive) {
                        const long long slot = (long long)slot0 + __popc(m & ((1u << lane_id()) - 1u));
                        q_nxt[slot] = q;
                        r_nxt[slot] = make_int2(ru, rv);
                        atomicMin(&best[ru], tag | q);
                        atomicMin(&best[rv], tag | q);
                    }
                }
            }
            kr_sync<SINGLE>(bar);
            const long long alive_n = (long long)kr_ld<SINGLE>(&ctr[1 + nxt]);
            if (alive_n == 0) break;
            // ---- phase B: hook the winners ----
            if (gtid == 0) ctr[1 + cur] = 0;   // becomes the append counter of the next round
            for (long long idx = gtid; idx < alive_n; idx += gsz) {
                const uint32_t q = LSMEM ? q_nxt[idx] : __ldcg(&q_nxt[idx]);
                const int2 r = LSMEM ? r_nxt[idx] : __ldcg(&r_nxt[idx]);
                const unsigned long long key = tag | q;
                const bool su = (SINGLE ? best[r.x] : __ldcg(&best[r.x])) == key;
                const bool sv = (SINGLE ? best[r.y] : __ldcg(&best[r.y])) == key;
                if (su || sv) {
                    const unsigned long long t = atomicAdd(&ctr[0], 1ull);
                    if (t < want) p.tree_out[t] = (long long)p.korder[q];
                    if (su && sv) {
                        const int hi = r.x > r.y ? r.x : r.y, lo = r.x > r.y ? r.y : r.x;
                        parent[hi] = lo;
                    } else if (su) {
                        parent[r.x] = r.y;
                    } else {
                        parent[r.y] = r.x;
                    }
                }
            }
            kr_sync<SINGLE>(bar);
            if (kr_ld<SINGLE>(&ctr[0]) >= want) { done = true; break; }
            first = false;
            cur = nxt;
        }
        // the break on alive_n == 0 leaves ctr[1+nxt] == 0 and ctr[1+cur] possibly stale: clear both
        kr_sync<SINGLE>(bar);
        if (gtid == 0) { ctr[1] = 0; ctr[2] = 0; }
        kr_sync<SINGLE>(bar);
        pos = cend;
        csize *= 2;
        if (csize > L.cap) csize = L.cap;
    }
}

__global__ void __launch_bounds__(kKrThreads) kruskal_kernel(KrParams p) {
    const KrLists L{{p.list[0], p.list[1]}, {p.roots[0], p.roots[1]}, p.list_cap};
    kruskal_body<false, false>(p, p.parent, p.best, p.ctr, L);
}

constexpr int kKrSmallThreads = 1024;
constexpr long long kKrSmallN = 16000;      // 12 bytes of shared memory per node
__global__ void __launch_bounds__(kKrSmallThreads) kruskal_small_kernel(KrParams p, long long lists_in_smem) {
    extern __shared__ __align__(16) unsigned char kr_raw[];
    unsigned long long *best = reinterpret_cast<unsigned long long *>(kr_raw);
    int *parent = reinterpret_cast<int *>(kr_raw + (size_t)p.N * sizeof(unsigned long long));
    __shared__ unsigned long long s_ctr[4];
    if (threadIdx.x < 4) s_ctr[threadIdx.x] = 0;
    __syncthreads();
    if (lists_in_smem > 0) {
        // small forests leave room for the survivor lists as well: a round then touches global memory only for the
        // first visit of an arc and for the tree arcs it records
        unsigned char *base = kr_raw + (((size_t)p.N * 12 + 15) & ~(size_t)15);
        int2 *r0 = reinterpret_cast<int2 *>(base), *r1 = r0 + lists_in_smem;
        uint32_t *q0 = reinterpret_cast<uint32_t *>(r1 + lists_in_smem), *q1 = q0 + lists_in_smem;
        const KrLists L{{q0, q1}, {r0, r1}, lists_in_smem};
        kruskal_body<true, true>(p, parent, best, s_ctr, L);
    } else {
        const KrLists L{{p.list[0], p.list[1]}, {p.roots[0], p.roots[1]}, p.list_cap};
        kruskal_body<true, false>(p, parent, best, s_ctr, L);
    }
    __syncthreads();
    if (threadIdx.x == 0) p.ctr[0] = s_ctr[0];          // read by kr_count_kernel
}

__global__ void kr_fill_kernel(long long *tree_out, long long cap, unsigned long long *ctr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) tree_out[i] = 0x7fffffffffffffffll;
    if (i < 8) ctr[i] = 0;                     // [0..2] counters, [4] grid barrier
}
__global__ void kr_count_kernel(const unsigned long long *ctr, long long cap, long long *n_tree_out) {
    const unsigned long long c = ctr[0];
    *n_tree_out = (long long)(c < (unsigned long long)cap ? c : (unsigned long long)cap);
}

// First chunk of the order in quarters of N (16 = 4 N); 0 = chosen by N in sx_kruskal.
static int g_kr_first_chunk_q = 0;

static long long kr_list_cap(long long N, long long n) {
    long long cap = 8 * N;
    if (cap < (1ll << 22)) cap = 1ll << 22;
    if (cap > n) cap = n;
    if (cap < 1) cap = 1;
    return cap;
}

}  // namespace sx

using namespace sx;

extern "C" int sx_kruskal_set_tuning(int first_chunk_quarters_of_N) {
    if (first_chunk_quarters_of_N < 0 || first_chunk_quarters_of_N > 64) return SX_ERR_INVALID;
    g_kr_first_chunk_q = first_chunk_quarters_of_N;
    return SX_OK;
}

extern "C" size_t sx_kruskal_workspace_bytes(int64_t N, int64_t n) {
    if (N < 0 || n < 0) return 0;
    const size_t cap = (size_t)kr_list_cap(N, n);
    const size_t T = (size_t)(N > 0 ? N : 1);
    return carve_bytes(N, 4) + carve_bytes(N, 8) + 2 * carve_bytes(cap, 4) + 2 * carve_bytes(cap, 8) +
           carve_bytes(8, 8) + carve_bytes(T, 8) + carve_bytes(T, 4) + sx_argsort_workspace_bytes((int64_t)T) + 256;
}

extern "C" int sx_kruskal(const uint32_t *korder, int64_t n, const int32_t *tail, const int32_t *head,
                          int64_t S, int64_t D, int64_t N, int64_t *tree_out, int64_t *n_tree_out,
                          void *ws, size_t ws_bytes, void *stream) {
    if (n < 0 || N < 0 || !tree_out || !n_tree_out) return SX_ERR_INVALID;
    if ((tail == nullptr) != (head == nullptr)) return SX_ERR_INVALID;
    if (!tail && (S <= 0 || D <= 0 || S + D != N || n > S * D)) return SX_ERR_INVALID;   // n < S * D: a prefix of the order
    if (N >= (1ll << 31) || n >= (1ll << 32)) return SX_ERR_TOO_LARGE;
    if (n > 0 && !korder) return SX_ERR_INVALID;
    if (!ws || ws_bytes < sx_kruskal_workspace_bytes(N, n)) return SX_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long tcap = N > 1 ? N - 1 : 0;

    KrParams p;
    Carver cv(ws);
    p.korder = korder; p.n = n; p.tail = tail; p.head = head; p.S = S; p.D = D; p.N = N;
    p.parent = cv.take<int>(N);
    p.best = cv.take<unsigned long long>(N);
    p.list_cap = kr_list_cap(N, n);
    {
        // measured (first chunk 4 N / 2 N / N / N/2 / N/4): 1 M nodes, 10 M arcs 3.12 / 1.71 / 1.22 / 1.19 / 1.29 ms --
        // once the first N/2 arcs have built the giant component, later chunks die in one filter pass instead of
        // surviving several hooking rounds; 40 K nodes (OT 20 000^2) 0.32 ms at 4 N, 0.30 at 2 N, 0.38 at 8 N
        const int q = g_kr_first_chunk_q > 0 ? g_kr_first_chunk_q : (N >= (1ll << 18) ? 2 : (N >= (1ll << 14) ? 8 : 16));
        p.first_chunk = (N * q + 3) / 4;
        if (p.first_chunk < 1024) p.first_chunk = 1024;
    }
    p.list[0] = cv.take<uint32_t>(p.list_cap); p.list[1] = cv.take<uint32_t>(p.list_cap);
    p.roots[0] = cv.take<int2>(p.list_cap);    p.roots[1] = cv.take<int2>(p.list_cap);
    p.ctr = cv.take<unsigned long long>(8);
    const size_t T = (size_t)(N > 0 ? N : 1);
    long long *raw_tree = cv.take<long long>(T);
    uint32_t *perm = cv.take<uint32_t>(T);
    void *sort_ws = cv.base + cv.off;
    const size_t sort_ws_bytes = ws_bytes - cv.off;
    p.tree_out = raw_tree;

    kr_fill_kernel<<<(int)((tcap + 8 + 255) / 256), 256, 0, st>>>(raw_tree, tcap, p.ctr);
    SX_LAUNCH_CHECK();
    if (n > 0 && tcap > 0 && N <= kKrSmallN) {
        // shared memory: 12 bytes per node, and -- when at least 4 096 survivors fit beside them -- both lists
        constexpr size_t kBudget = 200 * 1024;
        const size_t forest = (((size_t)N * 12 + 15) & ~(size_t)15) + 16;
        long long lists = forest < kBudget ? (long long)((kBudget - forest) / 24) : 0;
        if (lists < 4096) lists = 0;
        if (lists > p.list_cap) lists = p.list_cap;
        const size_t smem = forest + (size_t)lists * 24;
        SX_CUDA(cudaFuncSetAttribute(kruskal_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBudget + 1024));
        kruskal_small_kernel<<<1, kKrSmallThreads, smem, st>>>(p, lists);
        SX_LAUNCH_CHECK();
    } else if (n > 0 && tcap > 0) {
        int dev = 0, sms = 0, per_sm = 0;
        SX_CUDA(cudaGetDevice(&dev));
        SX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kruskal_kernel, kKrThreads, 0));
        if (per_sm < 1) return SX_ERR_NO_DEVICE;
        if (per_sm > 4) per_sm = 4;
        // the kernel is a chain of grid-wide barriers: with little work per round (small N) fewer, fatter
        // arrivals make each barrier cheaper (measured: 0.42 -> 0.34 ms at N = 40 000; the other way at N = 1e6)
        if (N <= (1ll << 18)) per_sm = 1;
        void *args[] = {&p};
        SX_CUDA(cudaLaunchCooperativeKernel((void *)kruskal_kernel, dim3(sms * per_sm), dim3(kKrThreads), args, 0, st));
    }
    kr_count_kernel<<<1, 1, 0, st>>>(p.ctr, tcap, (long long *)n_tree_out);
    SX_LAUNCH_CHECK();
    if (tcap > 0) {
        // ascending arc ids (np.flatnonzero order, tree_BI.py:56-57); sentinels sort last
        const long long id_bound = tail ? (1ll << 32) - 1 : S * D;      // korder may be a prefix: ids are not < n
        int bits = 1;
        while (bits < 63 && (1ll << bits) <= id_bound) ++bits;
        // the low `bits` bits of the sentinel are all ones (>= n > any arc id), so it still sorts last
        int rc = sx_argsort_u64((const unsigned long long *)raw_tree, tcap, bits, perm,
                                (unsigned long long *)tree_out, sort_ws, sort_ws_bytes, st);
        if (rc != SX_OK) return rc;
    }
    return SX_OK;
}
