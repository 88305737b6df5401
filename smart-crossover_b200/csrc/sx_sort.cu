// sx_sort.cu -- K1c: stable LSD radix argsort of 64-bit keys with 32-bit arc ids (sm_100a),
// plus the two orders the reference derives from it.
//
// Replaces `np.argsort(flow_indicators)` (reference net_manager.py:184,379; run stable per
// north_star) and the stable argsort inside scipy.sparse.csgraph.minimum_spanning_tree
// (tree_BI.py:53).  HBM-bound: per 8-bit pass an upsweep (digit histogram per block, 8 B/key),
// a single-block scan, and a downsweep (stable in-tile ranking with warp match + shuffles,
// shared-memory exchange, coalesced run-wise scatter; 12 B read + 12 B written per key).
#include "sx_common.cuh"
#include "sx_gridbar.cuh"

namespace sx {

constexpr int kRsItems   = 8;      // keys per thread in the downsweep; tile = threads * 8
constexpr int kRsMaxGrid = 148 * 6;   // cap on the upsweep grid (sizes the histogram workspace), not an SM count

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

// Lanes of the warp holding the same 8-bit digit, from 8 ballots.  The hardware MATCH.ANY
// instruction takes time proportional to the number of distinct values in the warp (32 for the
// random low mantissa bytes); this form costs the same for every distribution.
__device__ __forceinline__ unsigned match_digit(unsigned d) {
    unsigned peers = 0xffffffffu;
    // per bit: test (one LOP3 with a predicate result), ballot, select, one three-input logic op.  Written
    // in PTX because the C++ form is canonicalised into a shift, two predicate set-ups and the same tail.
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        asm("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 t, bal, m;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
            "selp.b32 m, 0xffffffff, 0, p;\n\t"
            "lop3.b32 %0, %0, bal, m, 0x90;\n\t"               // peers & ~(bal ^ m): lanes whose bit equals mine
            "}" : "+r"(peers) : "r"(d), "r"(1u << b));
    }
    return peers;
}

template <int SRC>
__device__ __forceinline__ unsigned long long rs_load_key(const RsSrc &s, long long i) {
    if (SRC == kFromF64) return f64_to_sort_key(s.f64[i]);
    return s.keys[i];
}

// Lowest byte of f64_to_sort_key(v) from the raw bits b of v: for an ordinary value it is the low byte of
// b, complemented when v is negative; zeros, infinities and NaNs take the full conversion.
__device__ __forceinline__ unsigned f64_bits_low_digit(unsigned long long b) {
    const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
    const unsigned h2 = hi << 1;
    if (h2 >= 0xffe00000u || (h2 | lo) == 0u)
        return (unsigned)f64_to_sort_key(__longlong_as_double((long long)b)) & 0xffu;
    return (lo ^ (unsigned)((int)hi >> 31)) & 0xffu;
}

// ---- upsweep: per-block digit histogram ------------------------------------------------
// Every lane owns a private 16-bit counter per digit (cnt[warp][digit][lane], 64 KB per CTA), so
// counting a key is a plain shared-memory read-modify-write: no atomics (shared atomics cost ~2
// cycles per lane on this part and made the histogram 5x slower than the memory system), no
// warp matching.  A lane sees at most tiles_per_block * 2048 / 128 < 65536 keys.
constexpr int kUpThreads = 128;
constexpr int kUpWarps   = kUpThreads / 32;
constexpr int kUpUnroll  = 8;
constexpr size_t kUpSmem = (size_t)kUpWarps * 256 * 32 * sizeof(uint16_t);

template <int SRC>
__global__ void __launch_bounds__(kUpThreads)
rs_upsweep_kernel(RsSrc src, long long n, int shift, long long keys_per_block, uint32_t *hist, int grid) {
    extern __shared__ __align__(16) uint16_t up_cnt[];
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(up_cnt);
        for (int i = threadIdx.x; i < (int)(kUpSmem / 4); i += kUpThreads) z[i] = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint16_t *mine = up_cnt + (size_t)warp * 256 * 32 + lane;
    const long long beg = (long long)blockIdx.x * keys_per_block;
    long long end = beg + keys_per_block;
    if (end > n) end = n;
    const bool raw_f64 = SRC == kFromF64 && shift == 0;      // the first pass of an f64 sort: low byte only
    const void *base_ptr = (SRC == kFromF64) ? (const void *)src.f64 : (const void *)src.keys;
    const bool vec = (reinterpret_cast<uintptr_t>(base_ptr) & 15) == 0;   // beg is even
    for (long long base = beg; base < end; base += (long long)kUpThreads * 2 * kUpUnroll) {
        unsigned long long k0[kUpUnroll], k1[kUpUnroll];
#pragma unroll
        for (int u = 0; u < kUpUnroll; ++u) {
            const long long i = base + 2ll * (u * kUpThreads + threadIdx.x);
            k0[u] = k1[u] = 0;
            if (i + 1 < end && vec) {
                if (SRC == kFromF64) {
                    const double2 q = *reinterpret_cast<const double2 *>(src.f64 + i);
                    if (raw_f64) {      // digit taken straight from the double's bits below
                        k0[u] = (unsigned long long)__double_as_longlong(q.x);
                        k1[u] = (unsigned long long)__double_as_longlong(q.y);
                    } else {
                        k0[u] = f64_to_sort_key(q.x); k1[u] = f64_to_sort_key(q.y);
                    }
                } else {
                    const ulonglong2 q = *reinterpret_cast<const ulonglong2 *>(src.keys + i);
                    k0[u] = q.x; k1[u] = q.y;
                }
            } else {
                if (i < end) k0[u] = raw_f64 ? (unsigned long long)__double_as_longlong(src.f64[i]) : rs_load_key<SRC>(src, i);
                if (i + 1 < end) k1[u] = raw_f64 ? (unsigned long long)__double_as_longlong(src.f64[i + 1]) : rs_load_key<SRC>(src, i + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < kUpUnroll; ++u) {
            const long long i = base + 2ll * (u * kUpThreads + threadIdx.x);
            if (raw_f64) {
                if (i < end) mine[f64_bits_low_digit(k0[u]) * 32] += 1;
                if (i + 1 < end) mine[f64_bits_low_digit(k1[u]) * 32] += 1;
            } else {
                if (i < end) mine[((k0[u] >> shift) & 0xff) * 32] += 1;
                if (i + 1 < end) mine[((k1[u] >> shift) & 0xff) * 32] += 1;
            }
        }
    }
    __syncthreads();
    // reduce the 4 x 32 private copies of every digit: a warp sums one (warp, digit) row per step
    for (int d = warp; d < 256; d += kUpWarps) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kUpWarps; ++w) s += up_cnt[((size_t)w * 256 + d) * 32 + lane];
        s = warp_sum(s);
        if (lane == 0) hist[(size_t)d * grid + blockIdx.x] = s;
    }
}

// ---- scan: exclusive prefix over hist[digit][block] (digit-major), one CTA ----
// Warp w owns one contiguous chunk (a multiple of 128 entries) and walks it with coalesced 128-bit
// accesses, 128 entries per step: first the chunk totals (independent loads), one block-wide barrier to
// turn the 32 totals into chunk offsets, then the walk again with a shuffle scan per step and the next
// step's load already in flight.  The first version (8192 entries per sweep, three barriers each) took
// 20 us for the 40 K counters of a 640 K-key sort and 135 us for the 227 K counters of a large one.
// count is a multiple of 256.
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *hist, int count) {
    __shared__ uint32_t warp_tot[32];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int chunk = (count + 31) / 32;
    chunk = (chunk + 127) & ~127;
    const int beg = warp * chunk;
    int end = beg + chunk;
    if (end > count) end = count;                            // count % 128 == 0: every step is whole
    uint4 *h4 = reinterpret_cast<uint4 *>(hist);
    uint32_t sum = 0;
#pragma unroll 4
    for (int base = beg; base < end; base += 128) {
        const uint4 q = h4[(base >> 2) + lane];
        sum += q.x + q.y + q.z + q.w;
    }
    sum = warp_sum(sum);
    if (lane == 0) warp_tot[warp] = sum;
    __syncthreads();
    const uint32_t wt = warp_tot[lane];                      // every warp scans the 32 chunk totals
    uint32_t wincl = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += t;
    }
    uint32_t carry = __shfl_sync(0xffffffffu, wincl - wt, warp);
    uint4 q = make_uint4(0, 0, 0, 0);
    if (beg < end) q = h4[(beg >> 2) + lane];
    for (int base = beg; base < end; base += 128) {
        uint4 nq = make_uint4(0, 0, 0, 0);
        if (base + 128 < end) nq = h4[((base + 128) >> 2) + lane];
        const uint32_t s4 = q.x + q.y + q.z + q.w;
        uint32_t incl = s4;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t run = carry + incl - s4;
        uint4 o4;
        o4.x = run; run += q.x;
        o4.y = run; run += q.y;
        o4.z = run; run += q.z;
        o4.w = run;
        h4[(base >> 2) + lane] = o4;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        q = nq;
    }
}

// ---- downsweep: stable scatter ------------------------------------------------------------
template <int THREADS>
struct RsSmem {
    unsigned long long keys[THREADS * kRsItems];
    uint32_t           vals[THREADS * kRsItems];
    uint32_t           cnt[THREADS / 32][256];
    uint32_t           digit_base[256];
    uint32_t           tile_cnt[256];
    uint32_t           tile_start[256];
    uint32_t           warp_tot[8];
};

// Stable in-tile ranking of the tile's keys by their digit: on return rank[r] is the key's position among the
// keys of its warp with the same digit, sm.cnt[w][d] the number of such keys in earlier warps, sm.tile_cnt[d] the
// tile's digit histogram and sm.tile_start[d] its exclusive scan.  sm.cnt must be zero and visible on entry.
template <int THREADS>
__device__ __forceinline__ void rs_rank_tile(RsSmem<THREADS> &sm, const unsigned long long (&key)[kRsItems], int shift,
                                             uint32_t (&rank)[kRsItems]) {
    constexpr int kWarps = THREADS / 32;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const unsigned d = (unsigned)((key[r] >> shift) & 0xff);
        const unsigned peers = match_digit(d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        // (a shared-memory atomic by the leader instead of load + store, with no __syncwarp between the rounds,
        // measured the same: 27.6 vs 27.2 ms at 2.7e8 keys)
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
        for (int w = 0; w < kWarps; ++w) {
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
}

// Exchange through shared memory: the tile's keys grouped by digit, stable.
template <int THREADS>
__device__ __forceinline__ void rs_exchange_tile(RsSmem<THREADS> &sm, const unsigned long long (&key)[kRsItems],
                                                 const uint32_t (&val)[kRsItems], const uint32_t (&rank)[kRsItems],
                                                 int shift) {
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const unsigned d = (unsigned)((key[r] >> shift) & 0xff);
        const uint32_t pos = sm.tile_start[d] + sm.cnt[warp][d] + rank[r];
        sm.keys[pos] = key[r];
        sm.vals[pos] = val[r];
    }
}

// Run-coalesced scatter of the exchanged tile to its global positions sm.digit_base[d] + (index within the tile's
// digit run).
template <int THREADS, bool LAST_F64>
__device__ __forceinline__ void rs_writeout_tile(RsSmem<THREADS> &sm, int shift, long long valid, const RsDst &dst) {
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        const int i = threadIdx.x + k * THREADS;
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
}

template <int THREADS, bool LAST_F64>
__device__ __forceinline__ void rs_scatter_tile(RsSmem<THREADS> &sm, const unsigned long long (&key)[kRsItems],
                                                const uint32_t (&val)[kRsItems], const uint32_t (&rank)[kRsItems],
                                                int shift, long long valid, const RsDst &dst) {
    rs_exchange_tile<THREADS>(sm, key, val, rank, shift);
    __syncthreads();
    rs_writeout_tile<THREADS, LAST_F64>(sm, shift, valid, dst);
}

// The tile loop is software pipelined (below).  Shapes measured with the pipelined loop at 4e8 keys: 1024 threads
// x 1 CTA/SM 28.9 ms, 512 x 2 32.7, 384 x 3 38.5 (56 registers leave no room to hold the prefetched tile), 256 x 4
// 43.5; before pipelining 512 x 2 took 39.8 ms on the same box class.
template <int THREADS, int SRC, bool LAST_F64>
__global__ void __launch_bounds__(THREADS, THREADS == 384 ? 3 : (THREADS == 512 ? 2 : (THREADS == 256 ? 4 : 1)))
rs_downsweep_kernel(RsSrc src, RsDst dst, long long n, int shift, long long tiles_per_block,
                    const uint32_t *hist, int grid) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    constexpr int kTile = THREADS * kRsItems, kWarps = THREADS / 32;
    RsSmem<THREADS> &sm = *reinterpret_cast<RsSmem<THREADS> *>(rs_raw);
    const int warp = threadIdx.x >> 5, lane = lane_id();
    if (threadIdx.x < 256) sm.digit_base[threadIdx.x] = hist[(size_t)threadIdx.x * grid + blockIdx.x];

    // Software pipeline over the CTA's tiles: once a tile's keys sit in the shared-memory exchange, its registers
    // are free, so the NEXT tile's keys are requested before the write-out of this one -- the load latency at the
    // top of a tile was 25 % of all stall samples (ncu source view).
    const long long tile0 = (long long)blockIdx.x * tiles_per_block;
    long long tile_end = tile0 + tiles_per_block;
    if (tile_end * kTile > n) tile_end = (n + kTile - 1) / kTile;
    if (tile0 >= tile_end) return;
    const int li0 = warp * (32 * kRsItems) + lane;
    unsigned long long key[kRsItems];
    uint32_t           val[kRsItems];
    auto load_tile = [&](long long tl) {
        const long long tbase = tl * kTile;
        const long long valid = (n - tbase < kTile) ? (n - tbase) : kTile;
        if (valid == kTile) {                                    // every tile but the last: no bounds checks
#pragma unroll
            for (int r = 0; r < kRsItems; ++r) {
                const long long gi = tbase + li0 + r * 32;
                key[r] = rs_load_key<SRC>(src, gi);
                val[r] = (SRC == kFromBuffer) ? src.vals[gi] : (uint32_t)gi;
            }
        } else {
#pragma unroll
            for (int r = 0; r < kRsItems; ++r) {
                const int li = li0 + r * 32;
                const long long gi = tbase + li;
                if (li < valid) {
                    key[r] = rs_load_key<SRC>(src, gi);
                    val[r] = (SRC == kFromBuffer) ? src.vals[gi] : (uint32_t)gi;
                } else {
                    key[r] = ~0ull;
                    val[r] = 0xffffffffu;
                }
            }
        }
    };
    for (int i = threadIdx.x; i < kWarps * 256; i += THREADS) (&sm.cnt[0][0])[i] = 0;
    load_tile(tile0);
    __syncthreads();
    for (long long tl = tile0; tl < tile_end; ++tl) {
        const long long tbase = tl * kTile;
        const long long valid = (n - tbase < kTile) ? (n - tbase) : kTile;
        uint32_t rank[kRsItems];
        rs_rank_tile<THREADS>(sm, key, shift, rank);
        rs_exchange_tile<THREADS>(sm, key, val, rank, shift);
        __syncthreads();
        // key / val / rank are dead, the counters have been read: fetch the next tile and clear the counters for it
        if (tl + 1 < tile_end) load_tile(tl + 1);
        for (int i = threadIdx.x; i < kWarps * 256; i += THREADS) (&sm.cnt[0][0])[i] = 0;
        rs_writeout_tile<THREADS, LAST_F64>(sm, shift, valid, dst);
        __syncthreads();
        if (threadIdx.x < 256) sm.digit_base[threadIdx.x] += sm.tile_cnt[threadIdx.x];
    }
}

// ---- small inputs: one CTA, every pass in shared memory ------------------------------------------------------
// Up to 8192 keys (the half-edges and tree arcs of a 784 x 784 instance, all of a 40 x 40 one): one tile of
// 1024 threads x 8 keys, ranked per 8-bit digit exactly like a downsweep tile and exchanged through shared
// memory; nothing touches global memory between the load and the final store.  ~3 us per pass instead of
// three launches.
constexpr int kSmallSortMax = 8192;
constexpr int kSmallThreads = 1024;

template <int SRC>
__global__ void __launch_bounds__(kSmallThreads)
rs_small_sort_kernel(RsSrc src, int n, int passes, uint32_t *order_out, double *sorted_f64,
                     unsigned long long *sorted_u64) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem<kSmallThreads> &sm = *reinterpret_cast<RsSmem<kSmallThreads> *>(rs_raw);
    constexpr int kWarps = kSmallThreads / 32;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int li0 = warp * (32 * kRsItems) + lane;
    unsigned long long key[kRsItems];
    uint32_t           val[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const int li = li0 + r * 32;
        key[r] = li < n ? rs_load_key<SRC>(src, li) : ~0ull;      // padding: all-ones, last by stability
        val[r] = li < n ? (uint32_t)li : 0xffffffffu;
    }
    for (int ps = 0; ps < passes; ++ps) {
        const int shift = 8 * ps;
        for (int i = threadIdx.x; i < kWarps * 256; i += kSmallThreads) (&sm.cnt[0][0])[i] = 0;
        __syncthreads();
        uint32_t rank[kRsItems];
        rs_rank_tile<kSmallThreads>(sm, key, shift, rank);
#pragma unroll
        for (int r = 0; r < kRsItems; ++r) {
            const unsigned d = (unsigned)((key[r] >> shift) & 0xff);
            const uint32_t pos = sm.tile_start[d] + sm.cnt[warp][d] + rank[r];
            sm.keys[pos] = key[r];
            sm.vals[pos] = val[r];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kRsItems; ++r) {
            key[r] = sm.keys[li0 + r * 32];
            val[r] = sm.vals[li0 + r * 32];
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const int li = li0 + r * 32;
        if (li < n) {
            order_out[li] = val[r];
            if (sorted_f64) sorted_f64[li] = sort_key_to_f64(key[r]);
            if (sorted_u64) sorted_u64[li] = key[r];
        }
    }
}

// ---- mid-size inputs: every pass of the sort in ONE cooperative launch ---------------------------------------
// One tile of 4096 keys per CTA, all CTAs resident (up to ~1.2 M keys).  Per 8-bit pass: rank the tile in shared
// memory, publish its digit histogram, grid barrier, every CTA reads the (256 x tiles) table through L2 and
// derives its own digit bases (digits before mine over all tiles + my digit over earlier tiles), scatter, grid
// barrier.  16 barriers instead of 24 launches: the 614 656 keys of a 784 x 784 instance are launch-latency
// bound (0.25 ms for ~2 us of memory traffic per pass).

struct RsCoop {
    const double *key_f64;                 // first pass source (one of the two)
    const unsigned long long *key_u64;
    unsigned long long *kbuf[2];
    uint32_t *vbuf[2];
    uint32_t *hist;                        // 256 x gridDim.x, digit-major
    GridBarrier *bar;
    long long n;
    int passes;
    uint32_t *order_out;
    double *sorted_f64;
    unsigned long long *sorted_u64;
};

// kCoopThreads = 512 (4 096-key tiles, two CTAs per SM) or 1 024 (8 192-key tiles, one per SM: half as many
// arrivals at the barriers and rows in the table -- taken from 64 tiles up).
template <bool F64, int kCoopThreads>
__global__ void __launch_bounds__(kCoopThreads, kCoopThreads == 512 ? 2 : 1) rs_coop_sort_kernel(RsCoop c) {
    constexpr int kCoopTile = kCoopThreads * kRsItems;
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem<kCoopThreads> &sm = *reinterpret_cast<RsSmem<kCoopThreads> *>(rs_raw);
    constexpr int kWarps = kCoopThreads / 32;
    constexpr int kPartHalf = (kCoopThreads / 64) * 256;            // uint32 words of one half of the partial-sum scratch
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int grid = gridDim.x;
    const long long tbase = (long long)blockIdx.x * kCoopTile;
    const long long valid = (c.n - tbase < kCoopTile) ? (c.n - tbase) : kCoopTile;
    const int li0 = warp * (32 * kRsItems) + lane;
    for (int ps = 0; ps < c.passes; ++ps) {
        const int shift = 8 * ps;
        const bool last = ps == c.passes - 1;
        // ping-pong buffers selected without indexing the parameter arrays (keeps them out of local memory)
        const unsigned long long *src_k = (ps & 1) ? c.kbuf[0] : c.kbuf[1];
        const uint32_t *src_v = (ps & 1) ? c.vbuf[0] : c.vbuf[1];
        unsigned long long *dst_k = (ps & 1) ? c.kbuf[1] : c.kbuf[0];
        uint32_t *dst_v = (ps & 1) ? c.vbuf[1] : c.vbuf[0];
        for (int i = threadIdx.x; i < kWarps * 256; i += kCoopThreads) (&sm.cnt[0][0])[i] = 0;
        unsigned long long key[kRsItems];
        uint32_t           val[kRsItems];
#pragma unroll
        for (int r = 0; r < kRsItems; ++r) {
            const int li = li0 + r * 32;
            const long long gi = tbase + li;
            key[r] = ~0ull;
            val[r] = 0xffffffffu;
            if (li < valid) {
                if (ps == 0) {
                    key[r] = F64 ? f64_to_sort_key(c.key_f64[gi]) : c.key_u64[gi];
                    val[r] = (uint32_t)gi;
                } else {
                    key[r] = __ldcg(src_k + gi);
                    val[r] = __ldcg(src_v + gi);
                }
            }
        }
        __syncthreads();
        uint32_t rank[kRsItems];
        rs_rank_tile<kCoopThreads>(sm, key, shift, rank);
        // the padding of the last tile sits in digit 0xff (all-ones keys): it is not part of the histogram
        if (threadIdx.x < 256) {
            uint32_t cnt = sm.tile_cnt[threadIdx.x];
            if (threadIdx.x == 255) cnt -= (uint32_t)(kCoopTile - valid);
            c.hist[(size_t)blockIdx.x * 256 + threadIdx.x] = cnt;               // tile-major: a row is 1 KB
        }
        grid_barrier(c.bar);
        // Every CTA reads the whole (tiles x 256) table: 64 threads cover a row with 128-bit loads, the 8
        // thread groups take rows t = g, g + 8, ... (all loads independent, several rows in flight per thread).
        // tot[d] = sum over all tiles, pre[d] = sum over the tiles before this one.
        {
            const int q = threadIdx.x & 63, g = threadIdx.x >> 6;                 // digits 4q..4q+3, row group g
            uint4 tot = make_uint4(0, 0, 0, 0), pre = make_uint4(0, 0, 0, 0);
            const uint4 *h4 = reinterpret_cast<const uint4 *>(c.hist);
#pragma unroll 4
            for (int t = g; t < grid; t += kCoopThreads / 64) {
                const uint4 v = __ldcg(h4 + (size_t)t * 64 + q);
                tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
                if (t < (int)blockIdx.x) { pre.x += v.x; pre.y += v.y; pre.z += v.z; pre.w += v.w; }
            }
            // partial sums of the 8 groups: sm.keys (free until the scatter) as 2 x 8 x 256 uint32
            uint32_t *part = reinterpret_cast<uint32_t *>(sm.keys);
            reinterpret_cast<uint4 *>(part + g * 256)[q] = tot;
            reinterpret_cast<uint4 *>(part + kPartHalf + g * 256)[q] = pre;
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            const uint32_t *part = reinterpret_cast<const uint32_t *>(sm.keys);
            uint32_t tot = 0, pre = 0;
#pragma unroll
            for (int g = 0; g < kCoopThreads / 64; ++g) { tot += part[g * 256 + threadIdx.x]; pre += part[kPartHalf + g * 256 + threadIdx.x]; }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) sm.warp_tot[warp] = incl;
            sm.digit_base[threadIdx.x] = incl - tot + pre;                     // warp-local exclusive + earlier tiles
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t add = 0;
            for (int w = 0; w < warp; ++w) add += sm.warp_tot[w];
            sm.digit_base[threadIdx.x] += add;
        }
        __syncthreads();
        RsDst dst{last ? c.sorted_u64 : dst_k, last ? c.sorted_f64 : nullptr, last ? c.order_out : dst_v};
        if (last && F64) rs_scatter_tile<kCoopThreads, true>(sm, key, val, rank, shift, valid, dst);
        else             rs_scatter_tile<kCoopThreads, false>(sm, key, val, rank, shift, valid, dst);
        if (!last) grid_barrier(c.bar);
    }
}

struct RsPlan {
    long long tiles, tiles_per_block;
    int       grid, threads;
};
static int g_rs_threads = 0;    // 0 = by size: 1024-thread tiles from 2^26 keys up, 512 below (sx_sort_set_tuning overrides)
static int g_rs_coop = 1;      // mid-size inputs take the single cooperative launch (sx_sort_set_tuning(0 / 1) switches it)
static RsPlan rs_plan(long long n) {
    RsPlan p;
    // 8192-key tiles (1024 threads, one CTA per SM) win on large inputs: half the barriers per key and twice as
    // long runs in the scatter (28.9 vs 32.7 ms at 4e8 keys, 7.8 vs 8.6 at 1e8); 512 x 2 CTAs/SM wins below
    // (1.08 vs 1.31 ms at 1e7, where a pass is latency bound)
    p.threads = g_rs_threads ? g_rs_threads : (n >= (1ll << 26) ? 1024 : 512);
    const long long tile = (long long)p.threads * kRsItems;
    p.tiles = (n + tile - 1) / tile;
    if (p.tiles < 1) p.tiles = 1;
    // One wave of downsweep CTAs (2 per SM) is enough up to ~64 M keys and keeps the (256 x grid) counter table --
    // which a single CTA scans between the sweeps -- small: 31 -> ~10 us per pass at 10 M keys.
    long long max_grid = n < (1ll << 26) ? (p.threads == 1024 ? 1ll : 2ll) * num_sms() : kRsMaxGrid;
    if (max_grid > kRsMaxGrid) max_grid = kRsMaxGrid;
    p.tiles_per_block = (p.tiles + max_grid - 1) / max_grid;
    p.grid = (int)((p.tiles + p.tiles_per_block - 1) / p.tiles_per_block);
    return p;
}

template <int SRC, bool LAST_F64>
static int rs_pass(const RsSrc &src, const RsDst &dst, long long n, int shift, const RsPlan &pl,
                   uint32_t *hist, cudaStream_t st) {
    auto up = rs_upsweep_kernel<SRC>;
    SX_CUDA(cudaFuncSetAttribute(up, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpSmem));
    up<<<pl.grid, kUpThreads, kUpSmem, st>>>(src, n, shift, pl.tiles_per_block * pl.threads * kRsItems, hist, pl.grid);
    SX_LAUNCH_CHECK();
    rs_scan_kernel<<<1, 1024, 0, st>>>(hist, 256 * pl.grid);
    SX_LAUNCH_CHECK();
    if (pl.threads == 384) {
        auto kern = rs_downsweep_kernel<384, SRC, LAST_F64>;
        SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem<384>)));
        kern<<<pl.grid, 384, sizeof(RsSmem<384>), st>>>(src, dst, n, shift, pl.tiles_per_block, hist, pl.grid);
    } else if (pl.threads == 256) {
        auto kern = rs_downsweep_kernel<256, SRC, LAST_F64>;
        SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem<256>)));
        kern<<<pl.grid, 256, sizeof(RsSmem<256>), st>>>(src, dst, n, shift, pl.tiles_per_block, hist, pl.grid);
    } else if (pl.threads == 1024) {
        auto kern = rs_downsweep_kernel<1024, SRC, LAST_F64>;
        SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem<1024>)));
        kern<<<pl.grid, 1024, sizeof(RsSmem<1024>), st>>>(src, dst, n, shift, pl.tiles_per_block, hist, pl.grid);
    } else {
        auto kern = rs_downsweep_kernel<512, SRC, LAST_F64>;
        SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem<512>)));
        kern<<<pl.grid, 512, sizeof(RsSmem<512>), st>>>(src, dst, n, shift, pl.tiles_per_block, hist, pl.grid);
    }
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
    if (n <= kSmallSortMax) {
        const size_t smem = sizeof(RsSmem<kSmallThreads>);
        RsSrc src{key_u64, key_f64, nullptr};
        if (key_f64) {
            SX_CUDA(cudaFuncSetAttribute(rs_small_sort_kernel<kFromF64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rs_small_sort_kernel<kFromF64><<<1, kSmallThreads, smem, st>>>(src, (int)n, passes, order_out, sorted_f64, sorted_u64);
        } else {
            SX_CUDA(cudaFuncSetAttribute(rs_small_sort_kernel<kFromU64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rs_small_sort_kernel<kFromU64><<<1, kSmallThreads, smem, st>>>(src, (int)n, passes, order_out, sorted_f64, sorted_u64);
        }
        SX_LAUNCH_CHECK();
        return SX_OK;
    }
    const RsPlan pl = rs_plan(n);
    Carver cv(ws);
    unsigned long long *kbuf[2] = {cv.take<unsigned long long>(n), cv.take<unsigned long long>(n)};
    uint32_t *vbuf[2] = {cv.take<uint32_t>(n), cv.take<uint32_t>(n)};
    uint32_t *hist = cv.take<uint32_t>((size_t)256 * kRsMaxGrid);
    GridBarrier *bar = cv.take<GridBarrier>(1);
    const bool f64 = key_f64 != nullptr;
    {
        // mid-size: one tile per CTA, all passes in one cooperative launch, if every tile can be resident
        static int coop_max[64][2] = {};                       // [device][0: 512-thread CTAs, 1: 1024-thread CTAs]
        int dev = 0;
        SX_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && coop_max[dev][0] == 0) {
            auto probe = [&](auto ka, auto kb, int threads, size_t smem, int &out) -> int {
                int occ_a = 0, occ_b = 0;
                SX_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                SX_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                SX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, ka, threads, smem));
                SX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, kb, threads, smem));
                const int occ = occ_a < occ_b ? occ_a : occ_b;
                out = occ > 0 ? occ * num_sms() : -1;
                return SX_OK;
            };
            int rc = probe(rs_coop_sort_kernel<true, 512>, rs_coop_sort_kernel<false, 512>, 512, sizeof(RsSmem<512>), coop_max[dev][0]);
            if (rc != SX_OK) return rc;
            rc = probe(rs_coop_sort_kernel<true, 1024>, rs_coop_sort_kernel<false, 1024>, 1024, sizeof(RsSmem<1024>), coop_max[dev][1]);
            if (rc != SX_OK) return rc;
        }
        const long long tiles512 = (n + 512 * kRsItems - 1) / (512 * kRsItems);
        const int big = tiles512 > 64 ? 1 : 0;
        const int threads = big ? 1024 : 512;
        const long long tiles = (n + (long long)threads * kRsItems - 1) / ((long long)threads * kRsItems);
        if (dev >= 0 && dev < 64 && g_rs_coop && tiles <= coop_max[dev][big] && tiles <= kRsMaxGrid) {
            RsCoop c{key_f64, key_u64, {kbuf[0], kbuf[1]}, {vbuf[0], vbuf[1]}, hist, bar, n, passes, order_out, sorted_f64, sorted_u64};
            SX_CUDA(cudaMemsetAsync(bar, 0, sizeof(GridBarrier), st));
            void *args[] = {(void *)&c};
            void *kern = big ? (f64 ? (void *)rs_coop_sort_kernel<true, 1024> : (void *)rs_coop_sort_kernel<false, 1024>)
                             : (f64 ? (void *)rs_coop_sort_kernel<true, 512> : (void *)rs_coop_sort_kernel<false, 512>);
            const size_t smem = big ? sizeof(RsSmem<1024>) : sizeof(RsSmem<512>);
            SX_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)tiles), dim3(threads), args, smem, st));
            return SX_OK;
        }
    }
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

// pass 1: first / last run head inside each block.  Position p is a head when key[p] != key[p - 1]; the
// previous key comes from the neighbouring lane (one extra load per warp and row, not per thread).
__global__ void __launch_bounds__(kKoThreads)
ko_block_heads_kernel(const double *__restrict__ key, long long n, long long *first_head, long long *last_head) {
    __shared__ long long s_min[32], s_max[32];
    const long long base = (long long)blockIdx.x * kKoTile;
    double k[kKoItems], prev0[kKoItems];
#pragma unroll
    for (int q = 0; q < kKoItems; ++q) {
        const long long p = base + (long long)q * kKoThreads + threadIdx.x;
        k[q] = p < n ? __ldcs(key + p) : 0.0;
        prev0[q] = (lane_id() == 0 && p > 0 && p < n) ? key[p - 1] : 0.0;
    }
    long long mn = n, mx = -1;
#pragma unroll
    for (int q = 0; q < kKoItems; ++q) {
        const long long p = base + (long long)q * kKoThreads + threadIdx.x;
        double pv = __shfl_up_sync(0xffffffffu, k[q], 1);
        if (lane_id() == 0) pv = prev0[q];
        if (p < n && (p == 0 || k[q] != pv)) {
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
    // blocked arrangement: thread t owns kKoItems consecutive positions; kk[j] = key[p0 - 1 + j]
    const long long p0 = base + (long long)threadIdx.x * kKoItems;
    static_assert(kKoItems == 4, "vector loads below assume 4 items per thread");
    double kk[kKoItems + 2];
    uint32_t ord[kKoItems];
    if (p0 + kKoItems <= n && ((reinterpret_cast<uintptr_t>(key) | reinterpret_cast<uintptr_t>(order)) & 15) == 0) {
        const double2 a = __ldcs(reinterpret_cast<const double2 *>(key + p0));
        const double2 b = __ldcs(reinterpret_cast<const double2 *>(key + p0 + 2));
        kk[1] = a.x; kk[2] = a.y; kk[3] = b.x; kk[4] = b.y;
        const uint4 o = __ldcs(reinterpret_cast<const uint4 *>(order + p0));
        ord[0] = o.x; ord[1] = o.y; ord[2] = o.z; ord[3] = o.w;
    } else {
#pragma unroll
        for (int q = 0; q < kKoItems; ++q) {
            kk[q + 1] = p0 + q < n ? key[p0 + q] : 0.0;
            ord[q] = p0 + q < n ? order[p0 + q] : 0u;
        }
    }
    kk[0] = (p0 > 0 && p0 <= n) ? key[p0 - 1] : 0.0;
    kk[kKoItems + 1] = p0 + kKoItems < n ? key[p0 + kKoItems] : 0.0;
    bool head[kKoItems + 1];
#pragma unroll
    for (int q = 0; q <= kKoItems; ++q) {
        const long long p = p0 + q;
        head[q] = (p < n) ? (p == 0 || kk[q + 1] != kk[q]) : (p == n);   // position n acts as a sentinel head
    }
    // the tile's arcs land (tie runs that cross the tile boundary aside) in the mirrored range
    // [lo, lo + len): they are staged in shared memory and written out coalesced
    __shared__ uint32_t s_out[kKoTile];
    const long long len = n - base < kKoTile ? n - base : kKoTile;
    const long long lo = n - base - len;
#pragma unroll
    for (int q = 0; q < kKoItems; ++q) s_out[threadIdx.x + q * kKoThreads] = 0xffffffffu;   // no arc id (ids < n < 2^32 - 1)
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
            const long long r = dstq - lo;
            if (r >= 0 && r < len) s_out[r] = ord[q];
            else korder[dstq] = ord[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kKoItems; ++q) {
        const int i = threadIdx.x + q * kKoThreads;
        const uint32_t v = s_out[i];
        if (i < len && v != 0xffffffffu) korder[lo + i] = v;
    }
}

// First position of the tie run that holds position `pos` of the ascending keys (lower bound of key[pos]).
__global__ void ko_run_start_kernel(const double *__restrict__ key, long long pos, long long *out) {
    const double v = key[pos];
    long long lo = 0, hi = pos;
    while (lo < hi) {
        const long long mid = lo + (hi - lo) / 2;
        if (key[mid] < v) lo = mid + 1; else hi = mid;
    }
    *out = lo;
}

}  // namespace sx

using namespace sx;

extern "C" int sx_sort_set_tuning(int downsweep_threads) {
    if (downsweep_threads == 0 || downsweep_threads == 1) { g_rs_coop = downsweep_threads; return SX_OK; }
    if (downsweep_threads == 2) { g_rs_threads = 0; return SX_OK; }        // back to the size-based choice
    if (downsweep_threads != 256 && downsweep_threads != 384 && downsweep_threads != 512 && downsweep_threads != 1024) return SX_ERR_INVALID;
    g_rs_threads = downsweep_threads;
    return SX_OK;
}

extern "C" size_t sx_argsort_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return 2 * carve_bytes((size_t)n, 8) + 2 * carve_bytes((size_t)n, 4) +
           carve_bytes((size_t)256 * kRsMaxGrid, 4) + 512;
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
    if (grid > num_sms() * 16) grid = num_sms() * 16;
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

// Head of the Kruskal order from the ascending sort: the tie runs that cover the last T positions, flipped.
// The slice starts at a run head, so it is flipped exactly like the whole array would be.
extern "C" int sx_kruskal_order_head(const double *sorted_key, const uint32_t *order_asc, int64_t n, int64_t T,
                                     int64_t T_cap, uint32_t *korder_out, int64_t *n_head_h, void *ws,
                                     size_t ws_bytes, void *stream) {
    if (n <= 0 || T <= 0 || T_cap < T || !sorted_key || !order_asc || !korder_out || !n_head_h) return SX_ERR_INVALID;
    if (!ws || ws_bytes < sx_kruskal_order_workspace_bytes(T_cap) + 256) return SX_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    long long start = 0;
    if (T < n) {
        long long *start_d = (long long *)ws;
        ko_run_start_kernel<<<1, 1, 0, st>>>(sorted_key, n - T, start_d);
        SX_LAUNCH_CHECK();
        SX_CUDA(cudaMemcpyAsync(&start, start_d, sizeof(start), cudaMemcpyDeviceToHost, st));
        SX_CUDA(cudaStreamSynchronize(st));
    }
    const long long m = n - start;
    if (m > T_cap) { *n_head_h = -1; return SX_OK; }
    *n_head_h = m;
    return sx_kruskal_order(sorted_key + start, order_asc + start, m, korder_out, (char *)ws + 256, ws_bytes - 256, stream);
}
