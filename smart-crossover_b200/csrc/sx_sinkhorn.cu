// sx_sinkhorn.cu -- interior-point warm start for optimal transport: entropic Sinkhorn iterations in
// the log domain on the dense cost matrix (sm_100a).
//
// The reference's experiment driver produces the first-order point it then crosses over with POT:
// `sinkhorn_x = sinkhorn(ot.s, ot.d, ot.M, reg=10, numItermax=1000)` (scripts/run_network_crossover.py:96),
// i.e. Sinkhorn-Knopp  u = a / (K v),  v = b / (K^T u),  K = exp(-M / reg),  X = diag(u) K diag(v),
// stopping when the column-marginal error drops below stopThr (checked every 10 iterations).  POT is a
// third-party package that is not in the reference tree (and not installable offline); this file
// restates the same iteration (same order: v first, from u = 1 / S) with potentials f = reg log u,
// g = reg log v, which is the identical map in exact arithmetic and does not overflow for small reg:
//     g_j = reg log b_j - reg LSE_i((f_i - M_ij) / reg)
//     f_i = reg log a_i - reg LSE_j((g_j - M_ij) / reg)
//     X_ij = exp((f_i + g_j - M_ij) / reg)
// Streaming passes over M: one per half-iteration (8 B per arc) with an online (running max, scaled
// sum) log-sum-exp taken in groups of 8 arcs per thread, plus one read + one write for X at the end.
// The column-marginal error of POT's test is free: with the new f and the old g the column sums of X
// are b_j exp((g_old_j - g_new_j) / reg), so the error of iteration k is known during the column pass
// of iteration k + 1 (the loop then stops with that pass's g, one half-step further than POT).
#include <math.h>

#include "sx_common.cuh"
#include "sx_expm.cuh"

namespace sx {

// The two streaming passes were instruction-issue and fp64-pipe bound, not HBM bound (round-2 capture of
// the first version, profiles/r02_sinkhorn_raw.csv: 84 % issue slots, 91 warp instructions per arc, fp64
// pipe 52 %, DRAM 33 %): one libm exp per arc plus a second one whenever ANY lane of the warp met a new
// running maximum.  They now take the arcs in groups of kSkGroup per thread -- loads first (kSkGroup
// independent 8-byte loads in flight), one shared maximum per group, one rescale of the running sum per
// group, and the sign-aware 11-operation exp of sx_expm.cuh -- ~17 fp64 operations per arc and no
// divergent branch.
constexpr int kSkGroup = 8;

__constant__ double c_exp_tab[kExpTabSize] = {SX_EXP_TAB_VALUES};

__device__ __forceinline__ void load_exp_tab(double *s_tab) {      // all threads of the CTA call this
    if (threadIdx.x < kExpTabSize) s_tab[threadIdx.x] = c_exp_tab[threadIdx.x];
    __syncthreads();
}

struct Lse {          // log-sum-exp accumulator: sum of exp(t_k) = s * exp(m)
    double m, s;
};
// Accumulation works on the unscaled w_k = potential - cost (t_k = w_k / reg): `wm` is the running
// maximum of w, acc.m = wm / reg rounded once, and every exponent is fma(w_k, 1 / reg, -acc.m).
__device__ __forceinline__ void lse_add_group(Lse &a, double &wm, const double (&w)[kSkGroup], double inv,
                                              const double *tab) {
    double gm = w[0];
#pragma unroll
    for (int k = 1; k < kSkGroup; ++k) gm = w[k] > gm ? w[k] : gm;
    gm = gm > wm ? gm : wm;
    if (gm == -INFINITY) return;                      // only masked slots so far
    const double q = gm * inv;
    double s = a.s * exp_nonpos(a.m - q, tab);        // a.m = -inf: exp gives 0 (and a.s is 0)
    const double nq = -q;
#pragma unroll
    for (int k = 0; k < kSkGroup; ++k) s += exp_nonpos(fma(w[k], inv, nq), tab);   // masked slots (-inf) add 0
    wm = gm;
    a.m = q;
    a.s = s;
}
__device__ __forceinline__ void lse_merge(Lse &a, const Lse &b) {
    if (b.m <= a.m) a.s += b.s * exp(b.m - a.m);
    else { a.s = a.s * exp(a.m - b.m) + b.s; a.m = b.m; }
}

// f update: one warp per row, lanes stride the columns (coalesced), kSkGroup * 32 columns per step.
__global__ void __launch_bounds__(256)
sk_row_kernel(const double *__restrict__ M, long long ld, long long S, long long D, const double *__restrict__ g,
              const double *__restrict__ log_a, double reg, double *__restrict__ f) {
    __shared__ double s_tab[kExpTabSize];
    load_exp_tab(s_tab);
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = lane_id();
    const double inv = 1.0 / reg;
    constexpr long long kStep = (long long)kSkGroup * 32;
    for (long long i = warp; i < S; i += n_warps) {
        const double *row = M + i * ld;
        Lse acc{-INFINITY, 0.0};
        double wm = -INFINITY;
        long long j0 = 0;
        double nx[kSkGroup];                          // the next group's costs, in flight during this group's math
        if (kStep <= D) {
#pragma unroll
            for (int k = 0; k < kSkGroup; ++k) nx[k] = __ldcs(row + lane + 32 * k);
        }
        for (; j0 + kStep <= D; j0 += kStep) {
            double w[kSkGroup];
#pragma unroll
            for (int k = 0; k < kSkGroup; ++k) w[k] = g[j0 + lane + 32 * k] - nx[k];
            if (j0 + 2 * kStep <= D) {
#pragma unroll
                for (int k = 0; k < kSkGroup; ++k) nx[k] = __ldcs(row + j0 + kStep + lane + 32 * k);
            }
            lse_add_group(acc, wm, w, inv, s_tab);
        }
        if (j0 < D) {
            double w[kSkGroup];
#pragma unroll
            for (int k = 0; k < kSkGroup; ++k) {
                const long long j = j0 + lane + 32 * k;
                w[k] = j < D ? g[j] - __ldcs(row + j) : -INFINITY;
            }
            lse_add_group(acc, wm, w, inv, s_tab);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Lse other{__shfl_xor_sync(0xffffffffu, acc.m, o), __shfl_xor_sync(0xffffffffu, acc.s, o)};
            if (other.s > 0.0) { if (acc.s > 0.0) lse_merge(acc, other); else acc = other; }
        }
        if (lane == 0) f[i] = reg * (log_a[i] - (acc.m + log(acc.s)));
    }
}

// g update, stage 1: a CTA covers 256 columns x `rows_per_chunk` rows; partial LSE per (chunk, column).
__global__ void __launch_bounds__(256)
sk_col_partial_kernel(const double *__restrict__ M, long long ld, long long S, long long D,
                      const double *__restrict__ f, double reg, long long rows_per_chunk,
                      double *__restrict__ pm, double *__restrict__ ps) {
    __shared__ double s_tab[kExpTabSize];
    load_exp_tab(s_tab);
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long chunk = blockIdx.y;
    const long long i0 = chunk * rows_per_chunk, i1 = i0 + rows_per_chunk < S ? i0 + rows_per_chunk : S;
    if (j >= D) return;
    const double inv = 1.0 / reg;
    const double *col = M + j;
    Lse acc{-INFINITY, 0.0};
    double wm = -INFINITY;
    long long i = i0;
    double nx[kSkGroup];                              // the next group's costs, in flight during this group's math
    if (i + kSkGroup <= i1) {
#pragma unroll
        for (int k = 0; k < kSkGroup; ++k) nx[k] = __ldcs(col + (i + k) * ld);
    }
    for (; i + kSkGroup <= i1; i += kSkGroup) {
        double w[kSkGroup];
#pragma unroll
        for (int k = 0; k < kSkGroup; ++k) w[k] = f[i + k] - nx[k];
        if (i + 2 * kSkGroup <= i1) {
#pragma unroll
            for (int k = 0; k < kSkGroup; ++k) nx[k] = __ldcs(col + (i + kSkGroup + k) * ld);
        }
        lse_add_group(acc, wm, w, inv, s_tab);
    }
    if (i < i1) {
        double w[kSkGroup];
#pragma unroll
        for (int k = 0; k < kSkGroup; ++k) w[k] = i + k < i1 ? f[i + k] - __ldcs(col + (i + k) * ld) : -INFINITY;
        lse_add_group(acc, wm, w, inv, s_tab);
    }
    pm[chunk * D + j] = acc.m;
    ps[chunk * D + j] = acc.s;
}

// g update, stage 2: merge the chunks of a column; accumulate POT's error sum_j (b_j (exp((g_old - g_new)/reg) - 1))^2.
__global__ void __launch_bounds__(256)
sk_col_merge_kernel(long long D, long long n_chunks, const double *__restrict__ pm, const double *__restrict__ ps,
                    const double *__restrict__ log_b, const double *__restrict__ b, double reg,
                    double *__restrict__ g, double *err2) {
    __shared__ double s_err[8];
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    double e2 = 0.0;
    if (j < D) {
        Lse acc{-INFINITY, 0.0};
        for (long long c = 0; c < n_chunks; ++c) {
            const Lse other{pm[c * D + j], ps[c * D + j]};
            if (other.s > 0.0) { if (acc.s > 0.0) lse_merge(acc, other); else acc = other; }
        }
        const double g_new = reg * (log_b[j] - (acc.m + log(acc.s)));
        const double e = b[j] * (exp((g[j] - g_new) / reg) - 1.0);
        e2 = e * e;
        g[j] = g_new;
    }
    e2 = warp_sum(e2);
    if (lane_id() == 0) s_err[threadIdx.x >> 5] = e2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_err[w];
        if (err2 != nullptr && t != 0.0) atomicAdd(err2, t);
    }
}

__global__ void sk_log_kernel(const double *__restrict__ a, long long n, double *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = log(a[i]);
}

__global__ void sk_fill_kernel(double *out, long long n, double v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = v;
}

// X_ij = exp((f_i + g_j - M_ij) / reg): a CTA per row (grid-stride), threads stride the columns.
__global__ void __launch_bounds__(256)
sk_plan_kernel(const double *__restrict__ M, long long ld, long long S, long long D, const double *__restrict__ f,
               const double *__restrict__ g, double reg, double *__restrict__ X) {
    const double inv = 1.0 / reg;
    for (long long i = blockIdx.x; i < S; i += gridDim.x) {
        const double fi = f[i];
        const double *row = M + i * ld;
        double *out = X + i * D;
        for (long long j = threadIdx.x; j < D; j += blockDim.x) out[j] = exp((fi + g[j] - __ldcs(row + j)) * inv);
    }
}

constexpr long long kSkRowsPerChunk = 128;

}  // namespace sx

using namespace sx;

extern "C" size_t sx_sinkhorn_workspace_bytes(int64_t S, int64_t D) {
    if (S < 0 || D < 0) return 0;
    const size_t chunks = ((size_t)S + kSkRowsPerChunk - 1) / kSkRowsPerChunk + 1;
    return carve_bytes((size_t)S, 8) + carve_bytes((size_t)D, 8) + 2 * carve_bytes(chunks * (size_t)D, 8) +
           carve_bytes(1, 8) + 256;
}

extern "C" int sx_sinkhorn_ot(const double *M, int64_t ld, int64_t S, int64_t D, const double *a, const double *b,
                              double reg, int64_t max_iter, double stop_thr, int64_t check_every, double *f,
                              double *g, double *x_out, int64_t *iters_h, double *err_h, void *ws, size_t ws_bytes,
                              void *stream) {
    if (!M || !a || !b || !f || !g || S <= 0 || D <= 0 || ld < D || !(reg > 0.0) || max_iter < 0) return SX_ERR_INVALID;
    if (!ws || ws_bytes < sx_sinkhorn_workspace_bytes(S, D)) return SX_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (check_every <= 0) check_every = 10;
    Carver cv(ws);
    double *log_a = cv.take<double>(S), *log_b = cv.take<double>(D);
    const long long n_chunks = (S + kSkRowsPerChunk - 1) / kSkRowsPerChunk;
    double *pm = cv.take<double>((size_t)n_chunks * D), *ps = cv.take<double>((size_t)n_chunks * D);
    double *err2 = cv.take<double>(1);
    sk_log_kernel<<<num_sms(), 256, 0, st>>>(a, S, log_a);
    sk_log_kernel<<<num_sms(), 256, 0, st>>>(b, D, log_b);
    SX_LAUNCH_CHECK();
    SX_CUDA(cudaMemsetAsync(g, 0, sizeof(double) * (size_t)D, st));
    sk_fill_kernel<<<num_sms(), 256, 0, st>>>(f, S, -reg * log((double)S));     // u = 1 / S
    SX_LAUNCH_CHECK();
    long long row_grid = (S * 32 + 255) / 256;
    if (row_grid > num_sms() * 16) row_grid = num_sms() * 16;
    const dim3 col_grid((unsigned)((D + 255) / 256), (unsigned)n_chunks);
    int64_t it = 0;
    double err = INFINITY;
    for (; it < max_iter; ++it) {
        // the column pass of iteration `it` also yields the marginal error of iteration it - 1
        const bool check = stop_thr > 0.0 && it > 0 && ((it - 1) % check_every == 0);
        if (check) SX_CUDA(cudaMemsetAsync(err2, 0, sizeof(double), st));
        sk_col_partial_kernel<<<col_grid, 256, 0, st>>>(M, ld, S, D, f, reg, kSkRowsPerChunk, pm, ps);
        SX_LAUNCH_CHECK();
        sk_col_merge_kernel<<<(int)((D + 255) / 256), 256, 0, st>>>(D, n_chunks, pm, ps, log_b, b, reg, g,
                                                                    check ? err2 : nullptr);
        SX_LAUNCH_CHECK();
        if (check) {
            double e2 = 0.0;
            SX_CUDA(cudaMemcpyAsync(&e2, err2, sizeof(double), cudaMemcpyDeviceToHost, st));
            SX_CUDA(cudaStreamSynchronize(st));
            err = sqrt(e2);
            if (err < stop_thr) break;
        }
        sk_row_kernel<<<(int)row_grid, 256, 0, st>>>(M, ld, S, D, g, log_a, reg, f);
        SX_LAUNCH_CHECK();
    }
    if (x_out) {
        sk_plan_kernel<<<num_sms() * 8, 256, 0, st>>>(M, ld, S, D, f, g, reg, x_out);
        SX_LAUNCH_CHECK();
    }
    if (iters_h) *iters_h = it;
    if (err_h) *err_h = err;
    return SX_OK;
}
