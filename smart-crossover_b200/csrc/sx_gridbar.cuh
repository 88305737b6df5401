// sx_gridbar.cuh -- reusable grid-wide barrier for cooperatively launched kernels (all CTAs resident).
// 8 bytes of device memory {arrivals, generation}, zero before the first use.  The last CTA to arrive
// resets the arrival count and bumps the generation the others spin on, so the barrier can be used any
// number of times and nothing depends on the grid size of earlier launches.
#pragma once
#include "sx_common.cuh"

namespace sx {

struct GridBarrier {
    unsigned int cnt;
    unsigned int gen;
};

__device__ __forceinline__ unsigned gb_ld_volatile(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned gb_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void grid_barrier(GridBarrier *b) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = gb_ld_volatile(&b->gen);             // read before arriving
        __threadfence();
        if (atomicAdd(&b->cnt, 1u) == gridDim.x - 1u) {
            *reinterpret_cast<volatile unsigned *>(&b->cnt) = 0u;
            __threadfence();
            atomicAdd(&b->gen, 1u);
        } else {
            while (gb_ld_acquire(&b->gen) == gen) {}
        }
        __threadfence();
    }
    __syncthreads();
}

}  // namespace sx
