// sx_sort.cu -- K1c: stable LSD radix argsort of 64-bit keys with 32-bit arc ids (sm_100a),
// plus the two orders the reference derives from it.
//
// Replaces `np.argsort(flow_indicators)` (reference net_manager.py:184,379; run stable per
// north_star) and the stable argsort inside scipy.sparse.csgraph.minimum_spanning_tree
// (tree_BI.py:53).  HBM-bound: per 8-bit pass an upsweep (digit histogram per block, 8 B/key),
// a single-block scan, and a downsweep (stable in-tile ranking with warp match + shuffles,
// shared-memory exchange, coalesced run-wise scatter; 12 B read + 12 B written per key).
#include "sx_common.cuh"

namespace sx {

constexpr int kRsThreads = 512;
constexpr int kRsItems   = 8;
constexpr int kRsTile    = kRsThreads * kRsItems;   // 4096 keys
constexpr int kRsWarps   = kRsThreads / 32;
constexpr int kRsMaxGrid = kNumSMs * 6;

enum KeySource { kFromBuffer = 0, kFromF64 = 1, kFromU64 = 2 };

struct RsSrc {
    const unsigned long long *keys;   // kFromBuffer / kFromU64
    const double             *f64;    // kFromF64
    const uint32_t           *vals;   // kFromBuffer
};
struct RsDst {
    unsigned long long *keys;         // may be null on the last pass
    double             *f64;          // last pass of an f64 sort: sorted keys as doubles (may be null)
    uint32_t           *vals;
};

template <int SRC>
__device__ __forceinline__ unsigned long long rs_load_key(const RsSrc &s, long long i) {
    if (SRC == kFromF64) return f64_to_sort_key(s.f64[i]);
    return s.keys[i];
}

// ---- upsweep: per-block digit histogram ------------------------------------------------
template <int SRC>
__global__ void __launch_bounds__(kRsThreads)
rs_upsweep_kernel(RsSrc src, long long n, int shift, long long tiles_per_block, uint32_t *hist, int grid) {
    __shared__ uint32_t h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const long long beg = (long long)blockIdx.x * tiles_per_block * kRsTile;
    long long end = beg + tiles_per_block * kRsTile;
    if (end > n) end = n;
    for (long long base = beg; base < end; base += kRsThreads * 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long i = base + (long long)q * kRsThreads + threadIdx.x;
            const bool ok = i < end;
            const unsigned d = ok ? (unsigned)((rs_load_key<SRC>(src, i) >> shift) & 0xff) : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (ok && (int)lane_id() == __ffs(peers) - 1) atomicAdd(&h[d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    if (threadIdx.x < 256) hist[(size_t)threadIdx.x * grid + blockIdx.x] = h[threadIdx.x];
}

// ---- scan: exclusive prefix over hist[digit][block] (digit-major) -----------------------
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *hist, int count) {
    __shared__ uint32_t warp_tot[32];
    const int per = (count + 1023) / 1024;
    const int beg = threadIdx.x * per;
    int end = beg + per;
    if (end > count) end = count;
    uint32_t sum = 0;
    for (int i = beg; i < end; ++i) sum += hist[i];
    // block exclusive scan of `sum`
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += t;
    }
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = warp_tot[threadIdx.x], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if ((int)lane_id() >= o) wi += t;
        }
        warp_tot[threadIdx.x] = wi - w;
    }
    __syncthreads();
    uint32_t run = warp_tot[threadIdx.x >> 5] + (incl - sum);
    for (int i = beg; i < end; ++i) {
        uint32_t v = hist[i];
        hist[i] = run;
        run += v;
    }
}

// ---- downsweep: stable scatter ------------------------------------------------------------
struct RsSmem {
    unsigned long long keys[kRsTile];
    uint32_t           vals[kRsTile];
    uint32_t           cnt[kRsWarps][256];
    uint32_t           digit_base[256];
    uint32_t           tile_cnt[256];
    uint32_t           tile_start[256];
    uint32_t           warp_tot[8];
};

template <int SRC, bool LAST_F64>
__global__ void __launch_bounds__(kRsThreads)
rs_downsweep_kernel(RsSrc src, RsDst dst, long long n, int shift, long long tiles_per_block,
                    const uint32_t *hist, int grid) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem &sm = *reinterpret_cast<RsSmem *>(rs_raw);
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    if (threadIdx.x < 256) sm.digit_base[threadIdx.x] = hist[(size_t)threadIdx.x * grid + blockIdx.x];

    const long long tile0 = (long long)blockIdx.x * tiles_per_block;
    for (long long tl = tile0; tl < tile0 + tiles_per_block; ++tl) {
        const long long tbase = tl * kRsTile;
        if (tbase >= n) break;
        const long long valid = (n - tbase < kRsTile) ? (n - tbase) : kRsTile;
        // zero the per-warp counters
        for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&sm.cnt[0][0])[i] = 0;
        unsigned long long key[kRsItems];
        uint32_t           val[kRsItems];
#pragma unroll
        for (int r = 0; r < kRsItems; ++r) {
            const long long li = (long long)warp * (32 * kRsItems) + r * 32 + lane;
            const long long gi = tbase + li;
            if (li < valid) {
                key[r] = rs_load_key<SRC>(src, gi);
                val[r] = (SRC == kFromBuffer) ? src.vals[gi] : (uint32_t)gi;
            } else {
                key[r] = ~0ull;
                val[r] = 0xffffffffu;
            }
        }
        __syncthreads();
        uint32_t rank[kRsItems];
#pragma unroll
        for (int r = 0; r < kRsItems; ++r) {
            const unsigned d = (unsigned)((key[r] >> shift) & 0xff);
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) {
                old = sm.cnt[warp][d];
                sm.cnt[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[r] = old + __popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        // per-digit exclusive scan over warps, and the tile's digit totals
        if (threadIdx.x < 256) {
            uint32_t s = 0;
#pragma unroll
            for (int w = 0; w < kRsWarps; ++w) {
                uint32_t t = sm.cnt[w][threadIdx.x];
                sm.cnt[w][threadIdx.x] = s;
                s += t;
            }
            sm.tile_cnt[threadIdx.x] = s;
            // exclusive scan over the 256 digits (8 warps x 32)
            uint32_t incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) sm.warp_tot[warp] = incl;
            sm.tile_start[threadIdx.x] = incl - s;   // warp-local exclusive
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t add = 0;
            for (int w = 0; w < warp; ++w) add += sm.warp_tot[w];
            sm.tile_start[threadIdx.x] += add;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kRsItems; ++r) {
            const unsigned d = (unsigned)((key[r] >> shift) & 0xff);
            const uint32_t pos = sm.tile_start[d] + sm.cnt[warp][d] + rank[r];
            sm.keys[pos] = key[r];
            sm.vals[pos] = val[r];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kRsItems; ++k) {
            const int i = threadIdx.x + k * kRsThreads;
            if (i < valid) {
                const unsigned long long kk = sm.keys[i];
                const unsigned d = (unsigned)((kk >> shift) & 0xff);
                const size_t g = (size_t)sm.digit_base[d] + (uint32_t)(i - sm.tile_start[d]);
                if (LAST_F64) {
                    if (dst.f64) dst.f64[g] = sort_key_to_f64(kk);
                } else if (dst.keys) {
                    dst.keys[g] = kk;
                }
                dst.vals[g] = sm.vals[i];
            }
        }
        __syncthreads();
        if (threadIdx.x < 256) sm.digit_base[threadIdx.x] += sm.tile_cnt[threadIdx.x];
    }
}

struct RsPlan {
    long long tiles, tiles_per_block;
    int       grid;
};
static RsPlan rs_plan(long long n) {
    RsPlan p;
    p.tiles = (n + kRsTile - 1) / kRsTile;
    if (p.tiles < 1) p.tiles = 1;
    p.tiles_per_block = (p.tiles + kRsMaxGrid - 1) / kRsMaxGrid;
    p.grid = (int)((p.tiles + p.tiles_per_block - 1) / p.tiles_per_block);
    return p;
}

template <int SRC, bool LAST_F64>
static int rs_pass(const RsSrc &src, const RsDst &dst, long long n, int shift, const RsPlan &pl,
                   uint32_t *hist, cudaStream_t st) {
    rs_upsweep_kernel<SRC><<<pl.grid, kRsThreads, 0, st>>>(src, n, shift, pl.tiles_per_block, hist, pl.grid);
    SX_LAUNCH_CHECK();
    rs_scan_kernel<<<1, 1024, 0, st>>>(hist, 256 * pl.grid);
    SX_LAUNCH_CHECK();
    auto kern = rs_downsweep_kernel<SRC, LAST_F64>;
    SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
    kern<<<pl.grid, kRsThreads, sizeof(RsSmem), st>>>(src, dst, n, shift, pl.tiles_per_block, hist, pl.grid);
    SX_LAUNCH_CHECK();
    return SX_OK;
}

// Generic driver: `passes` 8-bit passes; the last pass writes to the caller's outputs.
static int rs_sort(const double *key_f64, const unsigned long long *key_u64, long long n, int passes,
                   uint32_t *order_out, double *sorted_f64, unsigned long long *sorted_u64,
                   void *ws, size_t ws_bytes, cudaStream_t st) {
    if (n >= (1ll << 32)) return SX_ERR_TOO_LARGE;
    if (ws_bytes < sx_argsort_workspace_bytes(n) || !ws) return SX_ERR_WORKSPACE;
    if (n == 0) return SX_OK;
    const RsPlan pl = rs_plan(n);
    Carver cv(ws);
    unsigned long long *kbuf[2] = {cv.take<unsigned long long>(n), cv.take<unsigned long long>(n)};
    uint32_t *vbuf[2] = {cv.take<uint32_t>(n), cv.take<uint32_t>(n)};
    uint32_t *hist = cv.take<uint32_t>((size_t)256 * kRsMaxGrid);
    const bool f64 = key_f64 != nullptr;
    for (int ps = 0; ps < passes; ++ps) {
        const bool first = ps == 0, last = ps == passes - 1;
        RsSrc src{first ? key_u64 : kbuf[(ps - 1) & 1], key_f64, first ? nullptr : vbuf[(ps - 1) & 1]};
        RsDst dst{last ? sorted_u64 : kbuf[ps & 1], last ? sorted_f64 : nullptr, last ? order_out : vbuf[ps & 1]};
        int rc;
        const int shift = 8 * ps;
        if (first && f64)       rc = last ? rs_pass<kFromF64, true>(src, dst, n, shift, pl, hist, st)
                                          : rs_pass<kFromF64, false>(src, dst, n, shift, pl, hist, st);
        else if (first)         rc = rs_pass<kFromU64, false>(src, dst, n, shift, pl, hist, st);
        else if (last && f64)   rc = rs_pass<kFromBuffer, true>(src, dst, n, shift, pl, hist, st);
        else                    rc = rs_pass<kFromBuffer, false>(src, dst, n, shift, pl, hist, st);
        if (rc != SX_OK) return rc;
    }
    return SX_OK;
}

// ---- queue / Kruskal order -------------------------------------------------------------
__global__ void queue_from_order_kernel(const uint32_t *__restrict__ order, long long n, long long *queue) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
         q += (long long)gridDim.x * blockDim.x)
        queue[q] = (long long)order[n - 1 - q];
}

constexpr int kKoThreads = 1024;
constexpr int kKoItems   = 4;
constexpr int kKoTile    = kKoThreads * kKoItems;

__device__ __forceinline__ bool ko_is_head(const double *key, long long p) {
    return p == 0 || key[p] != key[p - 1];
}

// pass 1: first / last run head inside each block
__global__ void __launch_bounds__(kKoThreads)
ko_block_heads_kernel(const double *__restrict__ key, long long n, long long *first_head, long long *last_head) {
    __shared__ long long s_min[32], s_max[32];
    const long long base = (long long)blockIdx.x * kKoTile;
    long long mn = n, mx = -1;
#pragma unroll
    for (int q = 0; q < kKoItems; ++q) {
        const long long p = base + (long long)q * kKoThreads + threadIdx.x;
        if (p < n && ko_is_head(key, p)) {
            mn = p < mn ? p : mn;
            mx = p > mx ? p : mx;
        }
    }
    mn = warp_min(mn);
    mx = -warp_min(-mx);
    if (lane_id() == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kKoThreads / 32; ++w) {
            mn = s_min[w] < mn ? s_min[w] : mn;
            mx = s_max[w] > mx ? s_max[w] : mx;
        }
        first_head[blockIdx.x] = mn;
        last_head[blockIdx.x]  = mx;
    }
}

// pass 2: carries across blocks.  Block 0: exclusive forward max-scan of last_head (init -1)
// -> carry_left.  Block 1: exclusive backward min-scan of first_head (init n) -> carry_right,
// done as a forward max-scan of the negated, reversed sequence.
__global__ void __launch_bounds__(1024)
ko_carry_kernel(const long long *first_head, const long long *last_head, long long nb,
                long long n, long long *carry_left, long long *carry_right) {
    __shared__ long long warp_max[32];
    const bool rev = blockIdx.x == 1;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    auto get = [&](long long i) { return rev ? -first_head[nb - 1 - i] : last_head[i]; };
    const long long init = rev ? -n : -1;
    const long long per = (nb + 1023) / 1024;
    const long long beg = (long long)threadIdx.x * per;
    long long end = beg + per;
    if (end > nb) end = nb;
    long long m = init;
    for (long long i = beg; i < end; ++i) { long long v = get(i); m = v > m ? v : m; }
    long long incl = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = t > incl ? t : incl;
    }
    long long excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = init;
    if (lane == 31) warp_max[warp] = incl;
    __syncthreads();
    long long run = excl;
    for (int w = 0; w < warp; ++w) run = warp_max[w] > run ? warp_max[w] : run;
    for (long long i = beg; i < end; ++i) {
        const long long v = get(i);
        if (rev) carry_right[nb - 1 - i] = -run; else carry_left[i] = run;
        run = v > run ? v : run;
    }
}

// pass 3: run bounds per element and scatter into Kruskal order
//   b(p) = last head <= p,  e(p) = first head > p (or n);  q = (n - e) + (p - b)
__global__ void __launch_bounds__(kKoThreads)
ko_scatter_kernel(const double *__restrict__ key, const uint32_t *__restrict__ order, long long n,
                  const long long *carry_left, const long long *carry_right, uint32_t *korder) {
    __shared__ long long s_fwd[32], s_bwd[32];
    const long long base = (long long)blockIdx.x * kKoTile;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    // blocked arrangement: thread t owns kKoItems consecutive positions
    const long long p0 = base + (long long)threadIdx.x * kKoItems;
    bool head[kKoItems + 1];
#pragma unroll
    for (int q = 0; q <= kKoItems; ++q) {
        const long long p = p0 + q;
        head[q] = (p < n) ? ko_is_head(key, p) : (p == n);   // position n acts as a sentinel head
    }
    // forward: last head <= p   (thread-local, then warp / block max-scan)
    long long lastb[kKoItems];
    long long run = -1;
#pragma unroll
    for (int q = 0; q < kKoItems; ++q) {
        if (head[q] && p0 + q < n) run = p0 + q;
        lastb[q] = run;
    }
    long long incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = t > incl ? t : incl;
    }
    long long excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = -1;
    if (lane == 31) s_fwd[warp] = incl;
    // backward: first head > p
    long long nexte[kKoItems];
    long long runb = n + 1;   // "none inside this thread's tail"
#pragma unroll
    for (int q = kKoItems - 1; q >= 0; --q) {
        if (head[q + 1]) runb = p0 + q + 1;
        nexte[q] = runb;
    }
    // what the thread exposes to lower threads: first head at position >= p0 + 1 ... careful:
    // lower threads need the first head > their p, i.e. at positions >= p0 (their q+1 covers p0 only
    // for the last item, already handled through head[kKoItems]); so expose first head in (p0, p0+items].
    long long inclb = runb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_down_sync(0xffffffffu, inclb, o);
        if (lane + o < 32) inclb = t < inclb ? t : inclb;
    }
    long long exclb = __shfl_down_sync(0xffffffffu, inclb, 1);
    if (lane == 31) exclb = n + 1;
    if (lane == 0) s_bwd[warp] = inclb;
    __syncthreads();
    long long pre = carry_left[blockIdx.x];
    for (int w = 0; w < warp; ++w) pre = s_fwd[w] > pre ? s_fwd[w] : pre;
    pre = excl > pre ? excl : pre;
    long long suf = n + 1;
    for (int w = kKoThreads / 32 - 1; w > warp; --w) suf = s_bwd[w] < suf ? s_bwd[w] : suf;
    suf = exclb < suf ? exclb : suf;
    if (suf > n) suf = carry_right[blockIdx.x];
#pragma unroll
    for (int q = 0; q < kKoItems; ++q) {
        const long long p = p0 + q;
        if (p < n) {
            const long long b = lastb[q] >= 0 ? lastb[q] : pre;
            long long e = nexte[q] <= n ? nexte[q] : suf;
            const long long dstq = (n - e) + (p - b);
            korder[dstq] = order[p];
        }
    }
}

}  // namespace sx

using namespace sx;

extern "C" size_t sx_argsort_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return 2 * carve_bytes((size_t)n, 8) + 2 * carve_bytes((size_t)n, 4) +
           carve_bytes((size_t)256 * kRsMaxGrid, 4) + 256;
}

extern "C" int sx_argsort_f64(const double *key, int64_t n, uint32_t *order_asc_out, double *sorted_key_out,
                              void *ws, size_t ws_bytes, void *stream) {
    if (n < 0 || (n > 0 && (!key || !order_asc_out))) return SX_ERR_INVALID;
    return rs_sort(key, nullptr, n, 8, order_asc_out, sorted_key_out, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int sx_argsort_u64(const unsigned long long *key, int64_t n, int key_bits, uint32_t *order_asc_out,
                              unsigned long long *sorted_key_out, void *ws, size_t ws_bytes, void *stream) {
    if (n < 0 || key_bits < 1 || key_bits > 64 || (n > 0 && (!key || !order_asc_out))) return SX_ERR_INVALID;
    int passes = (key_bits + 7) / 8;
    if (passes < 2) passes = 2;   // the first pass always writes to the ping-pong buffers
    return rs_sort(nullptr, key, n, passes, order_asc_out, nullptr, sorted_key_out, ws, ws_bytes,
                   (cudaStream_t)stream);
}

extern "C" int sx_queue_from_order(const uint32_t *order_asc, int64_t n, int64_t *queue_out, void *stream) {
    if (n < 0 || (n > 0 && (!order_asc || !queue_out))) return SX_ERR_INVALID;
    if (n == 0) return SX_OK;
    long long grid = (n + 255) / 256;
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    queue_from_order_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(order_asc, n, (long long *)queue_out);
    SX_LAUNCH_CHECK();
    return SX_OK;
}

extern "C" size_t sx_kruskal_order_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    const size_t nb = ((size_t)n + kKoTile - 1) / kKoTile + 1;
    return 4 * carve_bytes(nb, 8) + 256;
}

extern "C" int sx_kruskal_order(const double *sorted_key, const uint32_t *order_asc, int64_t n,
                                uint32_t *korder_out, void *ws, size_t ws_bytes, void *stream) {
    if (n < 0 || (n > 0 && (!sorted_key || !order_asc || !korder_out))) return SX_ERR_INVALID;
    if (n == 0) return SX_OK;
    if (!ws || ws_bytes < sx_kruskal_order_workspace_bytes(n)) return SX_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long nb = (n + kKoTile - 1) / kKoTile;
    Carver cv(ws);
    long long *first_head = cv.take<long long>(nb + 1), *last_head = cv.take<long long>(nb + 1);
    long long *carry_left = cv.take<long long>(nb + 1), *carry_right = cv.take<long long>(nb + 1);
    ko_block_heads_kernel<<<(int)nb, kKoThreads, 0, st>>>(sorted_key, n, first_head, last_head);
    SX_LAUNCH_CHECK();
    ko_carry_kernel<<<2, 1024, 0, st>>>(first_head, last_head, nb, n, carry_left, carry_right);
    SX_LAUNCH_CHECK();
    ko_scatter_kernel<<<(int)nb, kKoThreads, 0, st>>>(sorted_key, order_asc, n, carry_left, carry_right, korder_out);
    SX_LAUNCH_CHECK();
    return SX_OK;
}
