// sx_price.cu -- K4: column-generation pricing pass (sm_100a).
//
// Dense OT:  rc_ij = fl(M_ij - fl(y_dst[j] - y_src[i]))  over a row slab of the fp64 cost
// matrix, replacing `self.mcf.c - self.mcf.A.T @ y` + `np.all(rc >= -tol)`
// (reference net_manager.py:474-497; association fixed by SciPy's csc_matvec).
// Arc list:  rc_k = c_k - (y[tail_k] - y[head_k]), negated where vbasis_k == -2
// (net_manager.py:293-319).
//
// The pass is HBM-bound (8 B per arc, no reuse, no tensor cores).  The main kernel is a
// persistent, warp-specialised TMA pipeline: one producer lane streams ROWS x 256 fp64
// boxes of M into a STAGES-deep shared-memory ring (cp.async.bulk.tensor + mbarrier
// complete_tx), 8 consumer warps read them back with 128-bit shared loads, keep the sink
// potentials of their two columns in registers and fuse: violator count, min reduced cost,
// and compaction of the violating (rc, arc id) pairs for the top-k selection (sx_topk.cu).
#include <cuda.h>
#include <math.h>

#include "sx_common.cuh"

namespace sx {

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA (cp.async.bulk.tensor)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

// ---------------------------------------------------------------------------------------
// Candidate compaction shared by all pricing kernels.
// A warp reserves slots for all violators of its tile with ONE atomicAdd on
// header->n_violating (which is also the exact violator count); once the buffer is full
// it stops reserving and only counts (flushed at kernel end).
// ---------------------------------------------------------------------------------------
struct CandSink {
    sx_price_header *hdr;
    double          *rc;
    int64_t         *id;
    long long        cap;
};

struct WarpTally {
    unsigned long long deferred = 0;   // violators counted but not reserved (lane 0 only)
    bool               full     = false;
};

// Reserve `n` (warp-uniform) slots; returns the base slot or -1 if nothing can be written.
__device__ __forceinline__ long long warp_reserve(const CandSink &sink, WarpTally &tally, unsigned n) {
    if (tally.full || sink.cap == 0) {
        if (lane_id() == 0) tally.deferred += n;
        return -1;
    }
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(&sink.hdr->n_violating, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned long long)sink.cap) {
        tally.full = true;
        return -1;
    }
    return (long long)base;
}
__device__ __forceinline__ void cand_store(const CandSink &sink, long long slot, double rc, long long id) {
    if (slot < sink.cap) {
        sink.rc[slot] = rc;
        sink.id[slot] = id;
    }
}
__device__ __forceinline__ void warp_flush(const CandSink &sink, const WarpTally &tally) {
    if (lane_id() == 0 && tally.deferred) atomicAdd(&sink.hdr->n_violating, tally.deferred);
}

// Block-level min -> one atomicMin per CTA.
__device__ __forceinline__ void block_min_commit(double tmin, sx_price_header *hdr, long long *smem_scratch,
                                                 int n_warps, int warp) {
    long long k = f64_to_min_key(tmin);
    k = warp_min(k);
    if (lane_id() == 0) smem_scratch[warp] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = smem_scratch[0];
        for (int w = 1; w < n_warps; ++w) m = smem_scratch[w] < m ? smem_scratch[w] : m;
        atomicMin(&hdr->min_rc_key, m);
    }
}

// ---------------------------------------------------------------------------------------
// K4a, variant 0: TMA pipeline
// ---------------------------------------------------------------------------------------
constexpr int kBoxCols      = 256;                  // fp64 elements per box row (TMA max box dim)
constexpr int kConsWarps    = 8;                    // 4 warps across the 256 columns x 2 row halves
constexpr int kConsThreads  = kConsWarps * 32;
constexpr int kTmaThreads   = kConsThreads + 32;    // + 1 producer warp

struct DenseParams {
    const double *y_src;   // S_loc
    const double *y_dst;   // D
    long long     S_loc, D, row0;
    double        thr;     // -tol
    CandSink      sink;
    double       *rc_out;  // optional
    long long     ld_out;
    long long     n_col_blocks, n_row_tiles;
};

template <int ROWS, int STAGES, bool WRITE_RC>
__global__ void __launch_bounds__(kTmaThreads, 1)
price_dense_tma_kernel(const __grid_constant__ CUtensorMap tmap, const DenseParams p) {
    constexpr int      kRowsPerHalf = ROWS / 2;
    constexpr uint32_t kStageBytes  = ROWS * kBoxCols * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double   *stage_base = reinterpret_cast<double *>(smem_raw);
    uint64_t *full_bar   = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * kStageBytes);
    uint64_t *empty_bar  = full_bar + STAGES;
    long long *scratch   = reinterpret_cast<long long *>(empty_bar + STAGES);

    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // contiguous, balanced range of tiles; tile t = row_tile * n_col_blocks + col_block so a
    // CTA walks along the rows of one row tile (16 sequential DRAM streams per CTA)
    const long long total = p.n_row_tiles * p.n_col_blocks;
    const long long t_beg = total * (long long)blockIdx.x / (long long)gridDim.x;
    const long long t_end = total * (long long)(blockIdx.x + 1) / (long long)gridDim.x;

    if (warp == kConsWarps) {
        // ===== producer warp: one elected lane issues the TMA loads =====
        if (lane_id() == 0) {
            const uint64_t pol = l2_evict_first_policy();
            long long rt = t_beg / p.n_col_blocks;
            long long cb = t_beg - rt * p.n_col_blocks;
            uint32_t  it = 0;
            for (long long t = t_beg; t < t_end; ++t, ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
                mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                tma_load_2d(stage_base + (size_t)s * ROWS * kBoxCols, &tmap, &full_bar[s],
                            (int)(cb * kBoxCols), (int)(rt * ROWS), pol);
                if (++cb == p.n_col_blocks) { cb = 0; ++rt; }
            }
        }
    } else {
        // ===== consumer warps =====
        const int c    = threadIdx.x & 127;           // column pair inside the box
        const int half = threadIdx.x >> 7;             // row half
        long long rt = t_beg / p.n_col_blocks;
        long long cb = t_beg - rt * p.n_col_blocks;
        double    tmin = INFINITY;
        WarpTally tally;
        double    u[kRowsPerHalf];
        long long rt_loaded = -1;
        // software-prefetched sink potentials of this thread's two columns
        auto load_v = [&](long long cbx, double &a, double &b) {
            const long long j = cbx * kBoxCols + 2 * c;
            a = (j < p.D) ? __ldg(p.y_dst + j) : 0.0;
            b = (j + 1 < p.D) ? __ldg(p.y_dst + j + 1) : 0.0;
        };
        double v0n = 0.0, v1n = 0.0;
        if (t_beg < t_end) load_v(cb, v0n, v1n);
        uint32_t it = 0;
        for (long long t = t_beg; t < t_end; ++t, ++it) {
            const int    s  = it % STAGES;
            const double v0 = v0n, v1 = v1n;
            const long long j0 = cb * kBoxCols + 2 * c;
            const long long i0 = rt * ROWS + (long long)half * kRowsPerHalf;
            long long cb_next = cb + 1, rt_next = rt;
            if (cb_next == p.n_col_blocks) { cb_next = 0; ++rt_next; }
            if (t + 1 < t_end) load_v(cb_next, v0n, v1n);
            if (rt != rt_loaded) {
#pragma unroll
                for (int r = 0; r < kRowsPerHalf; ++r)
                    u[r] = (i0 + r < p.S_loc) ? __ldg(p.y_src + i0 + r) : 0.0;
                rt_loaded = rt;
            }
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            const double2 *tile = reinterpret_cast<const double2 *>(stage_base + (size_t)s * ROWS * kBoxCols) +
                                  (size_t)half * kRowsPerHalf * (kBoxCols / 2) + c;
            double2 m[kRowsPerHalf];
#pragma unroll
            for (int r = 0; r < kRowsPerHalf; ++r) m[r] = tile[(size_t)r * (kBoxCols / 2)];
            // all shared reads of this stage are in registers: release the slot
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(&empty_bar[s]);

            const bool ok0 = j0 < p.D, ok1 = j0 + 1 < p.D;
            double   rc0[kRowsPerHalf], rc1[kRowsPerHalf];
            unsigned nviol = 0;
#pragma unroll
            for (int r = 0; r < kRowsPerHalf; ++r) {
                const bool rok = i0 + r < p.S_loc;
                const double a = m[r].x - (v0 - u[r]);
                const double b = m[r].y - (v1 - u[r]);
                rc0[r] = (rok && ok0) ? a : INFINITY;
                rc1[r] = (rok && ok1) ? b : INFINITY;
                tmin = fmin(tmin, fmin(rc0[r], rc1[r]));
                nviol += __popc(__ballot_sync(0xffffffffu, rc0[r] < p.thr)) +
                         __popc(__ballot_sync(0xffffffffu, rc1[r] < p.thr));
            }
            if (WRITE_RC) {
#pragma unroll
                for (int r = 0; r < kRowsPerHalf; ++r) {
                    if (i0 + r < p.S_loc) {
                        double *o = p.rc_out + (i0 + r) * p.ld_out + j0;
                        if (ok0) o[0] = rc0[r];
                        if (ok1) o[1] = rc1[r];
                    }
                }
            }
            if (nviol) {   // warp-uniform, rare
                long long slot = warp_reserve(p.sink, tally, nviol);
                if (slot >= 0) {
                    const unsigned lt = (1u << lane_id()) - 1u;
#pragma unroll
                    for (int r = 0; r < kRowsPerHalf; ++r) {
                        const long long gid = (p.row0 + i0 + r) * p.D + j0;
                        const unsigned b0 = __ballot_sync(0xffffffffu, rc0[r] < p.thr);
                        if (rc0[r] < p.thr) cand_store(p.sink, slot + __popc(b0 & lt), rc0[r], gid);
                        slot += __popc(b0);
                        const unsigned b1 = __ballot_sync(0xffffffffu, rc1[r] < p.thr);
                        if (rc1[r] < p.thr) cand_store(p.sink, slot + __popc(b1 & lt), rc1[r], gid + 1);
                        slot += __popc(b1);
                    }
                }
            }
            cb = cb_next;
            rt = rt_next;
        }
        warp_flush(p.sink, tally);
        // stash the per-thread min for the block reduction below
        long long k = warp_min(f64_to_min_key(tmin));
        if (lane_id() == 0) scratch[warp] = k;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = scratch[0];
        for (int w = 1; w < kConsWarps; ++w) m = scratch[w] < m ? scratch[w] : m;
        atomicMin(&p.sink.hdr->min_rc_key, m);
    }
}

// ---------------------------------------------------------------------------------------
// K4a, variants 1/2: direct global loads (no shared-memory staging).  VEC = 128-bit loads
// (needs 16 B alignment and even ld); otherwise 64-bit loads, any alignment.  Same tile
// walk and same fused epilogue as the TMA kernel; used as the fallback for odd leading
// dimensions and as the non-TMA comparison point in bench sweeps.
// ---------------------------------------------------------------------------------------
constexpr int kDirectRows    = 8;
constexpr int kDirectThreads = 128;

// streaming loads: read-only path, no L1 allocation, L2 evict-first (each cost is read once)
__device__ __forceinline__ double2 ldg_stream_v2(const double *p, uint64_t pol) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double *p, uint64_t pol) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}

template <bool VEC, bool WRITE_RC>
__global__ void __launch_bounds__(kDirectThreads)
price_dense_direct_kernel(const double *__restrict__ M, long long ld, const DenseParams p) {
    __shared__ long long scratch[kDirectThreads / 32];
    const int warp = threadIdx.x >> 5;
    const long long total = p.n_row_tiles * p.n_col_blocks;
    const uint64_t pol = l2_evict_first_policy();
    double    tmin = INFINITY;
    WarpTally tally;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const long long rt = t / p.n_col_blocks;
        const long long cb = t - rt * p.n_col_blocks;
        const long long i0 = rt * kDirectRows;
        // VEC: columns (2c, 2c+1); scalar: columns (c, c + 128) so each load is coalesced
        const long long j0 = cb * kBoxCols + (VEC ? 2 * threadIdx.x : threadIdx.x);
        const long long j1 = VEC ? j0 + 1 : j0 + 128;
        const bool ok0 = j0 < p.D, ok1 = j1 < p.D;
        const double v0 = ok0 ? __ldg(p.y_dst + j0) : 0.0;
        const double v1 = ok1 ? __ldg(p.y_dst + j1) : 0.0;
        double a[kDirectRows], b[kDirectRows];
#pragma unroll
        for (int r = 0; r < kDirectRows; ++r) {
            const bool rok = i0 + r < p.S_loc;
            const double *src = M + (i0 + r) * ld;
            if (VEC) {
                if (rok && ok1) { double2 q = ldg_stream_v2(src + j0, pol); a[r] = q.x; b[r] = q.y; }
                else { a[r] = (rok && ok0) ? ldg_stream(src + j0, pol) : 0.0; b[r] = 0.0; }
            } else {
                a[r] = (rok && ok0) ? ldg_stream(src + j0, pol) : 0.0;
                b[r] = (rok && ok1) ? ldg_stream(src + j1, pol) : 0.0;
            }
        }
        double   rc0[kDirectRows], rc1[kDirectRows];
        unsigned nviol = 0;
#pragma unroll
        for (int r = 0; r < kDirectRows; ++r) {
            const bool rok = i0 + r < p.S_loc;
            const double ui = rok ? __ldg(p.y_src + i0 + r) : 0.0;
            const double x0 = a[r] - (v0 - ui);
            const double x1 = b[r] - (v1 - ui);
            rc0[r] = (rok && ok0) ? x0 : INFINITY;
            rc1[r] = (rok && ok1) ? x1 : INFINITY;
            tmin = fmin(tmin, fmin(rc0[r], rc1[r]));
            nviol += __popc(__ballot_sync(0xffffffffu, rc0[r] < p.thr)) +
                     __popc(__ballot_sync(0xffffffffu, rc1[r] < p.thr));
            if (WRITE_RC && rok) {
                if (ok0) p.rc_out[(i0 + r) * p.ld_out + j0] = rc0[r];
                if (ok1) p.rc_out[(i0 + r) * p.ld_out + j1] = rc1[r];
            }
        }
        if (nviol) {
            long long slot = warp_reserve(p.sink, tally, nviol);
            if (slot >= 0) {
                const unsigned lt = (1u << lane_id()) - 1u;
#pragma unroll
                for (int r = 0; r < kDirectRows; ++r) {
                    const long long gbase = (p.row0 + i0 + r) * p.D;
                    const unsigned b0 = __ballot_sync(0xffffffffu, rc0[r] < p.thr);
                    if (rc0[r] < p.thr) cand_store(p.sink, slot + __popc(b0 & lt), rc0[r], gbase + j0);
                    slot += __popc(b0);
                    const unsigned b1 = __ballot_sync(0xffffffffu, rc1[r] < p.thr);
                    if (rc1[r] < p.thr) cand_store(p.sink, slot + __popc(b1 & lt), rc1[r], gbase + j1);
                    slot += __popc(b1);
                }
            }
        }
    }
    warp_flush(p.sink, tally);
    block_min_commit(tmin, p.sink.hdr, scratch, kDirectThreads / 32, warp);
}

// ---------------------------------------------------------------------------------------
// K4b: arc-list pricing
// ---------------------------------------------------------------------------------------
constexpr int kArcThreads = 256;
constexpr int kArcPerThread = 4;

__global__ void __launch_bounds__(kArcThreads)
price_arcs_kernel(const double *__restrict__ c, const int32_t *__restrict__ tail,
                  const int32_t *__restrict__ head, const int8_t *__restrict__ vbasis,
                  const double *__restrict__ y, long long E, long long id0, double thr,
                  CandSink sink, double *rc_out) {
    __shared__ long long scratch[kArcThreads / 32];
    const int warp = threadIdx.x >> 5;
    double    tmin = INFINITY;
    WarpTally tally;
    const long long chunk = (long long)kArcThreads * kArcPerThread;
    const long long n_chunks = (E + chunk - 1) / chunk;
    const uint64_t pol = l2_evict_first_policy();
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        double   rc[kArcPerThread];
        unsigned nviol = 0;
#pragma unroll
        for (int q = 0; q < kArcPerThread; ++q) {
            const long long k = ch * chunk + (long long)q * kArcThreads + threadIdx.x;
            double v = INFINITY;
            if (k < E) {
                const double ck = ldg_stream(c + k, pol);
                const double yt = __ldg(y + tail[k]);
                const double yh = __ldg(y + head[k]);
                v = ck - (yt - yh);
                if (vbasis != nullptr && vbasis[k] == -2) v = -v;
                if (rc_out != nullptr) rc_out[k] = v;
            }
            rc[q] = v;
            tmin = fmin(tmin, v);
            nviol += __popc(__ballot_sync(0xffffffffu, v < thr));
        }
        if (nviol) {
            long long slot = warp_reserve(sink, tally, nviol);
            if (slot >= 0) {
                const unsigned lt = (1u << lane_id()) - 1u;
#pragma unroll
                for (int q = 0; q < kArcPerThread; ++q) {
                    const long long k = ch * chunk + (long long)q * kArcThreads + threadIdx.x;
                    const unsigned b = __ballot_sync(0xffffffffu, rc[q] < thr);
                    if (rc[q] < thr) cand_store(sink, slot + __popc(b & lt), rc[q], id0 + k);
                    slot += __popc(b);
                }
            }
        }
    }
    warp_flush(sink, tally);
    block_min_commit(tmin, sink.hdr, scratch, kArcThreads / 32, warp);
}

__global__ void header_reset_kernel(sx_price_header *h) {
    h->n_violating = 0ull;
    h->min_rc_key  = 0x7fffffffffffffffll;
    h->n_priced    = 0ull;
    h->reserved    = 0ull;
}
__global__ void header_add_priced_kernel(sx_price_header *h, unsigned long long n) {
    atomicAdd(&h->n_priced, n);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// tunables of the TMA variant (overridable for bench sweeps through sx_price_set_tuning)
static int g_tma_rows = 16, g_tma_stages = 6, g_ctas_per_sm_direct = 8;

template <int ROWS, int STAGES, bool WRITE_RC>
static int launch_tma(const CUtensorMap &map, const DenseParams &p, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * ROWS * kBoxCols * sizeof(double) + 2 * STAGES * sizeof(uint64_t) +
                            kConsWarps * sizeof(long long) + 128;
    auto kern = price_dense_tma_kernel<ROWS, STAGES, WRITE_RC>;
    static bool attr_set = false;
    if (!attr_set) {
        SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long total = p.n_row_tiles * p.n_col_blocks;
    int grid = (int)(total < kNumSMs ? total : kNumSMs);
    if (grid < 1) grid = 1;
    kern<<<grid, kTmaThreads, smem, st>>>(map, p);
    SX_LAUNCH_CHECK();
    return SX_OK;
}

template <bool WRITE_RC>
static int dispatch_tma(const CUtensorMap &map, const DenseParams &p, cudaStream_t st) {
    if (g_tma_rows == 8 && g_tma_stages == 8) return launch_tma<8, 8, WRITE_RC>(map, p, st);
    if (g_tma_rows == 8 && g_tma_stages == 12) return launch_tma<8, 12, WRITE_RC>(map, p, st);
    if (g_tma_rows == 16 && g_tma_stages == 4) return launch_tma<16, 4, WRITE_RC>(map, p, st);
    if (g_tma_rows == 32 && g_tma_stages == 3) return launch_tma<32, 3, WRITE_RC>(map, p, st);
    return launch_tma<16, 6, WRITE_RC>(map, p, st);
}

}  // namespace sx

using namespace sx;

extern "C" int sx_price_set_tuning(int tma_rows, int tma_stages, int direct_ctas_per_sm) {
    if (tma_rows > 0) g_tma_rows = tma_rows;
    if (tma_stages > 0) g_tma_stages = tma_stages;
    if (direct_ctas_per_sm > 0) g_ctas_per_sm_direct = direct_ctas_per_sm;
    return SX_OK;
}

extern "C" int sx_price_header_reset(sx_price_header *header, void *stream) {
    if (!header) return SX_ERR_INVALID;
    header_reset_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(header);
    SX_LAUNCH_CHECK();
    return SX_OK;
}

extern "C" int sx_price_dense_ot(const double *M, int64_t ld, int64_t row0, int64_t S_loc, int64_t D,
                                 const double *y_src, const double *y_dst, double tol,
                                 sx_price_header *header, double *cand_rc, int64_t *cand_id,
                                 int64_t cand_cap, double *rc_out, int64_t ld_out, int variant,
                                 void *stream) {
    if (!M || !y_src || !y_dst || !header || S_loc < 0 || D <= 0 || ld < D || row0 < 0 || cand_cap < 0)
        return SX_ERR_INVALID;
    if (cand_cap > 0 && (!cand_rc || !cand_id)) return SX_ERR_INVALID;
    if (rc_out && ld_out < D) return SX_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (S_loc == 0) return SX_OK;
    const bool aligned = ((uintptr_t)M % 16 == 0) && (ld % 2 == 0);
    if (variant < 0) variant = aligned ? 0 : 2;
    if ((variant == 0 || variant == 1) && !aligned) return SX_ERR_UNALIGNED;
    if (variant == 0 && (D >= (1ll << 31) || S_loc >= (1ll << 31))) variant = 1;

    DenseParams p;
    p.y_src = y_src; p.y_dst = y_dst; p.S_loc = S_loc; p.D = D; p.row0 = row0; p.thr = -tol;
    p.sink.hdr = header; p.sink.rc = cand_rc; p.sink.id = (int64_t *)cand_id; p.sink.cap = cand_cap;
    p.rc_out = rc_out; p.ld_out = ld_out;
    p.n_col_blocks = (D + kBoxCols - 1) / kBoxCols;

    int rc = SX_OK;
    if (variant == 0) {
        EncodeTiledFn enc = get_encode_fn();
        if (!enc) return SX_ERR_NO_DEVICE;
        const int rows = g_tma_rows;
        p.n_row_tiles = (S_loc + rows - 1) / rows;
        CUtensorMap map;
        cuuint64_t gdim[2]    = {(cuuint64_t)D, (cuuint64_t)S_loc};
        cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
        cuuint32_t box[2]     = {(cuuint32_t)kBoxCols, (cuuint32_t)rows};
        cuuint32_t estr[2]    = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)M, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return SX_ERR_CUDA; }
        rc = rc_out ? dispatch_tma<true>(map, p, st) : dispatch_tma<false>(map, p, st);
    } else {
        p.n_row_tiles = (S_loc + kDirectRows - 1) / kDirectRows;
        const long long total = p.n_row_tiles * p.n_col_blocks;
        long long grid = (long long)kNumSMs * g_ctas_per_sm_direct;
        if (grid > total) grid = total;
        if (variant == 1) {
            if (rc_out) price_dense_direct_kernel<true, true><<<(int)grid, kDirectThreads, 0, st>>>(M, ld, p);
            else        price_dense_direct_kernel<true, false><<<(int)grid, kDirectThreads, 0, st>>>(M, ld, p);
        } else {
            if (rc_out) price_dense_direct_kernel<false, true><<<(int)grid, kDirectThreads, 0, st>>>(M, ld, p);
            else        price_dense_direct_kernel<false, false><<<(int)grid, kDirectThreads, 0, st>>>(M, ld, p);
        }
        SX_LAUNCH_CHECK();
    }
    if (rc != SX_OK) return rc;
    header_add_priced_kernel<<<1, 1, 0, st>>>(header, (unsigned long long)(S_loc * D));
    SX_LAUNCH_CHECK();
    return SX_OK;
}

extern "C" int sx_price_arcs(const double *c, const int32_t *tail, const int32_t *head,
                             const int8_t *vbasis, const double *y, int64_t E, int64_t id0, double tol,
                             sx_price_header *header, double *cand_rc, int64_t *cand_id,
                             int64_t cand_cap, double *rc_out, void *stream) {
    if (!c || !tail || !head || !y || !header || E < 0 || cand_cap < 0) return SX_ERR_INVALID;
    if (cand_cap > 0 && (!cand_rc || !cand_id)) return SX_ERR_INVALID;
    if (E == 0) return SX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CandSink sink{header, cand_rc, (int64_t *)cand_id, cand_cap};
    const long long chunk = (long long)kArcThreads * kArcPerThread;
    long long n_chunks = (E + chunk - 1) / chunk;
    long long grid = (long long)kNumSMs * 8;
    if (grid > n_chunks) grid = n_chunks;
    price_arcs_kernel<<<(int)grid, kArcThreads, 0, st>>>(c, tail, head, vbasis, y, E, id0, -tol, sink, rc_out);
    SX_LAUNCH_CHECK();
    header_add_priced_kernel<<<1, 1, 0, st>>>(header, (unsigned long long)E);
    SX_LAUNCH_CHECK();
    return SX_OK;
}
