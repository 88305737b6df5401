// sx_api.cu -- library-level entry points of libsxcross and the host-buffer pricing call.
#include <math.h>

#include "sx_common.cuh"

namespace sx {
int g_last_cuda_error = 0;

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
}
using namespace sx;

extern "C" int sx_abi_version(void) { return SX_ABI_VERSION; }

extern "C" const char *sx_error_string(int code) {
    switch (code) {
        case SX_OK: return "ok";
        case SX_ERR_INVALID: return "invalid argument";
        case SX_ERR_CUDA: return "CUDA runtime error (see sx_last_cuda_error)";
        case SX_ERR_WORKSPACE: return "workspace too small";
        case SX_ERR_TOO_LARGE: return "problem exceeds the 32-bit arc / node id range of this build";
        case SX_ERR_NOT_SPANNING: return "arcs do not form a spanning tree";
        case SX_ERR_UNALIGNED: return "cost matrix not 16-byte aligned or odd leading dimension";
        case SX_ERR_NO_DEVICE: return "no usable sm_100 device / driver entry point";
        case SX_ERR_PEER_TIMEOUT: return "a peer GPU did not deliver its block in time";
        case SX_ERR_PUSH_ASSERT: return "push phase: no positive flow to push against (reference assertion, tree_BI.py:93-94)";
        default: return "unknown error";
    }
}

extern "C" int sx_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" double sx_key_to_f64(long long key) { return min_key_to_f64(key); }

namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) { p = nullptr; return cuda_fail(e); }
        return SX_OK;
    }
};
}  // namespace

// One pricing pass driven from HOST buffers (what a ctypes binding inside the reference's
// `OTManager.check_optimality_condition`, net_manager.py:485-497, would call with NumPy arrays).
extern "C" int sx_price_dense_ot_h(const double *M_h, const double *M_dev, int64_t S, int64_t D,
                                   const double *y_h, double tol, int64_t K,
                                   unsigned long long *n_violating_h, double *min_rc_h,
                                   double *topk_rc_h, int64_t *topk_id_h, int64_t *topk_n_h) {
    if ((!M_h && !M_dev) || !y_h || S <= 0 || D <= 0 || K < 0 || !n_violating_h || !min_rc_h) return SX_ERR_INVALID;
    if (K > 0 && (!topk_rc_h || !topk_id_h || !topk_n_h)) return SX_ERR_INVALID;
    cudaStream_t st = nullptr;
    SX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};

    DevBuf m_buf, y_buf, hdr_buf, sel_buf, crc_buf, cid_buf, orc_buf, oid_buf, on_buf, ws_buf;
    const double *M = M_dev;
    int rc;
    if (!M) {
        if ((rc = m_buf.alloc((size_t)S * D * sizeof(double))) != SX_OK) return rc;
        SX_CUDA(cudaMemcpyAsync(m_buf.p, M_h, (size_t)S * D * sizeof(double), cudaMemcpyHostToDevice, st));
        M = (const double *)m_buf.p;
    }
    long long cap = K > 0 ? (K * 64 > (1ll << 20) ? K * 64 : (1ll << 20)) : 0;
    if (cap > S * D) cap = S * D;
    if ((rc = y_buf.alloc((size_t)(S + D) * sizeof(double))) != SX_OK) return rc;
    if ((rc = hdr_buf.alloc(sizeof(sx_price_header))) != SX_OK) return rc;
    SX_CUDA(cudaMemcpyAsync(y_buf.p, y_h, (size_t)(S + D) * sizeof(double), cudaMemcpyHostToDevice, st));
    sx_price_header *hdr_d = (sx_price_header *)hdr_buf.p;
    if (K > 0) {
        if ((rc = sel_buf.alloc(sx_select_state_bytes())) != SX_OK) return rc;
        if ((rc = orc_buf.alloc((size_t)K * 8)) != SX_OK) return rc;
        if ((rc = oid_buf.alloc((size_t)K * 8)) != SX_OK) return rc;
        if ((rc = on_buf.alloc(8)) != SX_OK) return rc;
    }
    sx_select_state *sel_d = (sx_select_state *)sel_buf.p;
    sx_price_header hdr;
    for (;;) {
        size_t wsb = 0;
        if (K > 0) {
            if ((rc = crc_buf.alloc((size_t)cap * 8)) != SX_OK) return rc;
            if ((rc = cid_buf.alloc((size_t)cap * 8)) != SX_OK) return rc;
            wsb = sx_topk_workspace_bytes(cap, K);
            if ((rc = ws_buf.alloc(wsb)) != SX_OK) return rc;
        }
        if ((rc = sx_price_pass_begin(hdr_d, sel_d, K, st)) != SX_OK) return rc;
        rc = sx_price_dense_ot(M, D, 0, S, D, (const double *)y_buf.p, (const double *)y_buf.p + S, tol, hdr_d, sel_d,
                               (double *)crc_buf.p, (int64_t *)cid_buf.p, cap, nullptr, 0, -1, st);
        if (rc != SX_OK) return rc;
        if (K > 0) {
            rc = sx_topk_select((const double *)crc_buf.p, (const int64_t *)cid_buf.p, cap, sel_d, hdr_d, K,
                                (double *)orc_buf.p, (int64_t *)oid_buf.p, (int64_t *)on_buf.p, ws_buf.p, wsb, st);
            if (rc != SX_OK) return rc;
        }
        SX_CUDA(cudaMemcpyAsync(&hdr, hdr_d, sizeof(hdr), cudaMemcpyDeviceToHost, st));
        SX_CUDA(cudaStreamSynchronize(st));
        if (K > 0 && (hdr.status & SX_STATUS_CAND_OVERFLOW) && cap < S * D) {
            // the pruned candidate list still outgrew its buffer (ties): enlarge it and price again
            cap = cap * 4 < S * D ? cap * 4 : S * D;
            cudaFree(crc_buf.p); crc_buf.p = nullptr;
            cudaFree(cid_buf.p); cid_buf.p = nullptr;
            cudaFree(ws_buf.p); ws_buf.p = nullptr;
            continue;
        }
        if (K > 0 && (hdr.status & SX_STATUS_NEED_SORTED)) {
            rc = sx_topk_select_sorted((const double *)crc_buf.p, (const int64_t *)cid_buf.p, cap, sel_d, hdr_d, K,
                                       (double *)orc_buf.p, (int64_t *)oid_buf.p, (int64_t *)on_buf.p, ws_buf.p, wsb,
                                       st);
            if (rc != SX_OK) return rc;
        }
        break;
    }
    *n_violating_h = hdr.n_violating;
    *min_rc_h = min_key_to_f64(hdr.min_rc_key);
    if (K > 0) {
        SX_CUDA(cudaMemcpyAsync(topk_rc_h, orc_buf.p, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
        SX_CUDA(cudaMemcpyAsync(topk_id_h, oid_buf.p, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
        SX_CUDA(cudaMemcpyAsync(topk_n_h, on_buf.p, 8, cudaMemcpyDeviceToHost, st));
        SX_CUDA(cudaStreamSynchronize(st));
    }
    return SX_OK;
}
