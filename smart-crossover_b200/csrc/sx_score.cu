// sx_score.cu -- K1a / K1b: per-arc flow-ratio scores (sm_100a).
//
// K1a replaces `np.maximum(X / s[:,None], X / d[None,:])` (reference net_manager.py:377-378).
// K1b replaces the SciPy sparse pipeline of `MCFManagerStd.get_sorted_flows`
// (net_manager.py:165-182).  Both must be bit-exact with NumPy/SciPy: correctly rounded
// IEEE fp64 division, NumPy `maximum` semantics, no FMA contraction (built with -fmad=false)
// and, for K1b, per-node sums taken sequentially in ascending arc id (csr_matvec order).
#include <math.h>

#include "sx_common.cuh"

namespace sx {

// NumPy: maximum(a, b) = isnan(a) ? a : (a > b ? a : b)
__device__ __forceinline__ double np_maximum(double a, double b) {
    return (a != a) ? a : (a > b ? a : b);
}

// max(x / s, x / d) with ONE division when that is provably the same double: correctly rounded
// division is monotone in the divisor, so for finite x >= 0 and s, d > 0 the larger quotient is
// x / min(s, d), bit for bit (equal quotients are the same value either way).  Anything else
// (negative, NaN or infinite x, non-positive or NaN marginals) takes the literal two-division form.
__device__ __forceinline__ double ot_score(double x, double s, double d) {
    // x is +0, a positive subnormal or a positive finite number  <=>  its high word is below 0x7ff00000
    // (one integer compare; -0.0 takes the literal form below, which gives the same -0.0)
    if ((unsigned)__double2hiint(x) < 0x7ff00000u && s > 0.0 && d > 0.0) return x / (s < d ? s : d);
    return np_maximum(x / s, x / d);
}

constexpr int kScThreads = 256;
constexpr int kScCols    = 4;   // columns per thread, strided by the block width (coalesced)

__global__ void __launch_bounds__(kScThreads)
score_ot_kernel(const double *__restrict__ x, const double *__restrict__ s, const double *__restrict__ d,
                long long S, long long D, long long col_blocks, double *__restrict__ out) {
    const long long tiles = S * col_blocks;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long i = t / col_blocks;
        const long long j0 = (t - i * col_blocks) * (kScThreads * kScCols) + threadIdx.x;
        const double si = __ldg(s + i);
        double xv[kScCols], dv[kScCols];
#pragma unroll
        for (int q = 0; q < kScCols; ++q) {
            const long long j = j0 + (long long)q * kScThreads;
            if (j < D) { xv[q] = x[i * D + j]; dv[q] = __ldg(d + j); }
        }
#pragma unroll
        for (int q = 0; q < kScCols; ++q) {
            const long long j = j0 + (long long)q * kScThreads;
            if (j < D) out[i * D + j] = ot_score(xv[q], si, dv[q]);
        }
    }
}

// Vector form: two adjacent columns per 128-bit access (needs even D and 16-byte aligned x / out),
// kScCols pairs per thread.  Optionally accumulates the 4096-bin histogram of the scores' top 12 key
// bits (sign + exponent) that sx_kruskal_prefix would otherwise compute with a pass of its own: the
// scores are already in registers here.  Shared-memory privatised (two sub-histograms, run-length
// cached), flushed once per CTA.
constexpr int kScBins = 4096;

template <bool HIST, int MIN_CTAS>
__global__ void __launch_bounds__(kScThreads, MIN_CTAS)
score_ot_vec_kernel(const double *__restrict__ x, const double *__restrict__ s, const double *__restrict__ d,
                    long long S, long long D, long long col_blocks, double *__restrict__ out,
                    unsigned *__restrict__ hist) {
    __shared__ unsigned sh[HIST ? 2 : 1][HIST ? kScBins : 1];
    unsigned *mine = sh[HIST ? (threadIdx.x & 1) : 0];
    unsigned run_d = 0xffffffffu, run_c = 0;
    if (HIST) {
        for (int i = threadIdx.x; i < 2 * kScBins; i += kScThreads) (&sh[0][0])[i] = 0;
        __syncthreads();
    }
    auto count = [&](double v) {
        // top 12 bits of the order-preserving image; for a value with the sign bit clear that is its
        // sign + exponent field with the top bit set (a positive NaN lands in bin 4095 either way)
        const int hi = __double2hiint(v);
        const unsigned dg = hi >= 0 ? ((unsigned)hi >> 20) | 0x800u : (unsigned)(f64_to_sort_key(v) >> 52);
        if (dg == run_d) { ++run_c; return; }
        if (run_c) atomicAdd(&mine[run_d], run_c);
        run_d = dg; run_c = 1;
    };
    const long long D2 = D / 2;
    // tile t = (row i, column block cb); the walk t += gridDim.x is kept as (i, cb) so that no 64-bit
    // division is needed per tile
    long long i = (long long)blockIdx.x / col_blocks;
    long long cb = (long long)blockIdx.x - i * col_blocks;
    const long long step_i = (long long)gridDim.x / col_blocks, step_cb = (long long)gridDim.x - step_i * col_blocks;
    for (; i < S; ) {
        const long long p0 = cb * (kScThreads * kScCols) + threadIdx.x;   // pair index in the row
        const double si = __ldg(s + i);
        const double2 *xr = reinterpret_cast<const double2 *>(x + i * D);
        double2 *orow = reinterpret_cast<double2 *>(out + i * D);
        double2 xv[kScCols], dv[kScCols];
#pragma unroll
        for (int q = 0; q < kScCols; ++q) {
            const long long pj = p0 + (long long)q * kScThreads;
            if (pj < D2) { xv[q] = __ldcs(xr + pj); dv[q] = __ldg(reinterpret_cast<const double2 *>(d) + pj); }
        }
#pragma unroll
        for (int q = 0; q < kScCols; ++q) {
            const long long pj = p0 + (long long)q * kScThreads;
            if (pj < D2) {
                const double2 r = make_double2(ot_score(xv[q].x, si, dv[q].x), ot_score(xv[q].y, si, dv[q].y));
                orow[pj] = r;
                if (HIST) { count(r.x); count(r.y); }
            }
        }
        i += step_i;
        cb += step_cb;
        if (cb >= col_blocks) { cb -= col_blocks; ++i; }
    }
    if (HIST) {
        if (run_c) atomicAdd(&mine[run_d], run_c);
        __syncthreads();
        for (int b = threadIdx.x; b < kScBins; b += kScThreads) {
            const unsigned c = sh[0][b] + sh[1][b];
            if (c) atomicAdd(&hist[b], c);
        }
    }
}

// ---- K1b ----------------------------------------------------------------------------------
// x_hat and the reversal flag, net_manager.py:166-168 evaluated literally:
//   x_hat = x * (~mask) + u * mask - x * mask ;  x_hat[(x < 0) | (x > u)] = 0
__global__ void mcf_xhat_kernel(const double *__restrict__ x, const double *__restrict__ u, long long E,
                                double *__restrict__ xhat, int8_t *__restrict__ flip) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < E;
         k += (long long)gridDim.x * blockDim.x) {
        const double xv = x[k], uv = u[k];
        const bool   m  = xv > uv / 2;
        const double nm = m ? 0.0 : 1.0, mm = m ? 1.0 : 0.0;
        double xh = (xv * nm + uv * mm) - xv * mm;
        if ((xv < 0) || (xv > uv)) xh = 0.0;
        xhat[k] = xh;
        flip[k] = m ? 1 : 0;
    }
}

// per node: f1 = sum over incident arcs with A_bar = +1, f2 = with A_bar = -1, both sequential
// from 0.0 in ascending arc id (net_manager.py:171-175); f_inv = 1 / max(f1, f2) or 0 (:176-177).
__global__ void mcf_node_kernel(const long long *__restrict__ node_ptr, const int32_t *__restrict__ node_arc,
                                const int8_t *__restrict__ node_sign, const int8_t *__restrict__ flip,
                                const double *__restrict__ xhat, long long N, double *__restrict__ finv) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < N;
         v += (long long)gridDim.x * blockDim.x) {
        double f1 = 0.0, f2 = 0.0;
        const long long q1 = node_ptr[v + 1];
        for (long long q = node_ptr[v]; q < q1; ++q) {
            const int32_t k = node_arc[q];
            int sg = node_sign[q];
            if (flip[k]) sg = -sg;
            const double xh = xhat[k];
            if (sg > 0) f1 += (double)sg * xh;
            else if (sg < 0) f2 += (double)(-sg) * xh;
        }
        const double f = np_maximum(f1, f2);
        finv[v] = (f != 0) ? 1.0 / f : 0.0;
    }
}

// per arc: max over its end nodes of |f_inv[node] * x_hat| (net_manager.py:178-182)
__global__ void mcf_arc_kernel(const int32_t *__restrict__ tail, const int32_t *__restrict__ head,
                               const double *__restrict__ finv, const double *__restrict__ xhat, long long E,
                               double *__restrict__ out) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < E;
         k += (long long)gridDim.x * blockDim.x) {
        const double xh = xhat[k];
        const int32_t t = tail[k], h = head[k];
        double r = 0.0;
        if (t >= 0) r = np_maximum(r, fabs(__ldg(finv + t) * xh));
        if (h >= 0) r = np_maximum(r, fabs(__ldg(finv + h) * xh));
        out[k] = r;
    }
}

static int grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace sx

using namespace sx;

static int g_score_min_ctas = 4;     // resident CTAs per SM the vector kernel is compiled for (3: 78 registers, 4: 64; 4 measured 8-18 % faster)
extern "C" int sx_score_set_tuning(int min_ctas_per_sm) {
    if (min_ctas_per_sm != 3 && min_ctas_per_sm != 4) return SX_ERR_INVALID;
    g_score_min_ctas = min_ctas_per_sm;
    return SX_OK;
}

extern "C" int sx_score_ot(const double *x, const double *s, const double *d, int64_t S, int64_t D,
                           double *score_out, uint32_t *hist12_out, void *stream) {
    if (S < 0 || D < 0) return SX_ERR_INVALID;
    if (S == 0 || D == 0) return SX_OK;
    if (!x || !s || !d || !score_out) return SX_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (D % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(score_out) |
                                       reinterpret_cast<uintptr_t>(d)) & 15) == 0;
    if (vec) {
        const long long col_blocks = (D / 2 + kScThreads * kScCols - 1) / (kScThreads * kScCols);
        const long long tiles = S * col_blocks;
        const long long grid = tiles < (long long)num_sms() * 8 ? tiles : (long long)num_sms() * 8;
        auto launch = [&](auto kern) { kern<<<(int)grid, kScThreads, 0, st>>>(x, s, d, S, D, col_blocks, score_out, hist12_out); };
        if (g_score_min_ctas >= 4) { if (hist12_out) launch(score_ot_vec_kernel<true, 4>); else launch(score_ot_vec_kernel<false, 4>); }
        else { if (hist12_out) launch(score_ot_vec_kernel<true, 3>); else launch(score_ot_vec_kernel<false, 3>); }
        SX_LAUNCH_CHECK();
        return SX_OK;
    }
    const long long col_blocks = (D + kScThreads * kScCols - 1) / (kScThreads * kScCols);
    long long tiles = S * col_blocks;
    long long grid = tiles < (long long)num_sms() * 16 ? tiles : (long long)num_sms() * 16;
    score_ot_kernel<<<(int)grid, kScThreads, 0, st>>>(x, s, d, S, D, col_blocks, score_out);
    SX_LAUNCH_CHECK();
    if (hist12_out) return sx_hist12_f64(score_out, S * D, hist12_out, stream);      // unaligned shapes: separate pass
    return SX_OK;
}

extern "C" size_t sx_score_mcf_workspace_bytes(int64_t N, int64_t E) {
    if (N < 0 || E < 0) return 0;
    return carve_bytes((size_t)E, 8) + carve_bytes((size_t)E, 1) + carve_bytes((size_t)N, 8) + 256;
}

extern "C" int sx_score_mcf(const double *x, const double *u, const int32_t *tail, const int32_t *head,
                            const int64_t *node_ptr, const int32_t *node_arc, const int8_t *node_sign,
                            int64_t N, int64_t E, double *score_out, void *ws, size_t ws_bytes,
                            void *stream) {
    if (N < 0 || E < 0) return SX_ERR_INVALID;
    if (E == 0) return SX_OK;
    if (!x || !u || !tail || !head || !node_ptr || !node_arc || !node_sign || !score_out) return SX_ERR_INVALID;
    if (N >= (1ll << 31) || E >= (1ll << 31)) return SX_ERR_TOO_LARGE;
    if (!ws || ws_bytes < sx_score_mcf_workspace_bytes(N, E)) return SX_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(ws);
    double *xhat = cv.take<double>(E);
    int8_t *flip = cv.take<int8_t>(E);
    double *finv = cv.take<double>(N);
    mcf_xhat_kernel<<<grid_for(E, 256), 256, 0, st>>>(x, u, E, xhat, flip);
    SX_LAUNCH_CHECK();
    mcf_node_kernel<<<grid_for(N, 128), 128, 0, st>>>((const long long *)node_ptr, node_arc, node_sign, flip, xhat, N, finv);
    SX_LAUNCH_CHECK();
    mcf_arc_kernel<<<grid_for(E, 256), 256, 0, st>>>(tail, head, finv, xhat, E, score_out);
    SX_LAUNCH_CHECK();
    return SX_OK;
}
