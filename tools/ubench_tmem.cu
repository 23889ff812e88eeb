// ubench_tmem.cu -- B200 tensor-memory microbenchmark for the "thread owns a TMEM lane" access pattern:
// round-trip latency and per-SM throughput of tcgen05.ld / tcgen05.st (32x32b.x4 = one (S_P, S_Q) pair of
// doubles per thread) against the same pattern on shared memory (LDS.128 / STS.128, conflict-free stride).
// Informs the TMEM-resident S tile of the step kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_tmem ubench_tmem.cu && ./ubench_tmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void lds4(uint32_t a, uint32_t (&r)[4]) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts4(uint32_t a, const uint32_t (&r)[4]) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: dependent chain  ld -> wait -> +1 -> st -> wait  (latency of one round trip)
// MODE 1: 16 independent loads in flight, one wait, 16 stores, one wait (throughput)
template <int MODE, bool TMEM>
__global__ void __launch_bounds__(128, 1) k_rt(uint32_t* out, int iters, long long* cycles) {
    __shared__ uint32_t slot;
    extern __shared__ __align__(16) uint8_t sm[];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);          // this warp's lane quarter
    const uint32_t row = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 33 * 16;   // 33 x 16 B per thread: conflict-free
    uint32_t v[4] = {(uint32_t)threadIdx.x, 1u, 2u, 3u};
    for (int c = 0; c < 32; ++c) {                                        // initialise 32 pairs per thread
        if (TMEM) tmem_st4(base + 4 * c, v); else sts4(row + 16 * c, v);
    }
    if (TMEM) wait_st();
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                uint32_t r[4];
                if (TMEM) { tmem_ld4(base + 4 * c, r); wait_ld(); }
                else lds4(row + 16 * c, r);
                r[0] += acc; acc = r[1] + r[0];
                if (TMEM) { tmem_st4(base + 4 * ((c + 1) & 31), r); wait_st(); }
                else sts4(row + 16 * ((c + 1) & 31), r);
            }
        } else {
            uint32_t r[16][4];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (TMEM) tmem_ld4(base + 4 * c, r[c]);
                else lds4(row + 16 * c, r[c]);
            }
            if (TMEM) wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) { r[c][0] += acc; acc += r[c][1]; }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (TMEM) tmem_st4(base + 4 * (c + 16), r[c]);
                else sts4(row + 16 * (c + 16), r[c]);
            }
            if (TMEM) wait_st();
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE, bool TMEM>
static void run(const char* what, int warps_note) {
    uint32_t* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000, smem = 128 * 33 * 16;
    cudaFuncSetAttribute(k_rt<MODE, TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_rt<MODE, TMEM><<<148, 128, smem>>>(out, 10, cyc);
    k_rt<MODE, TMEM><<<148, 128, smem>>>(out, iters, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / iters / 16.0;
    if (MODE == 0) printf("%-34s %7.1f clk per (load -> store) round trip, 4 warps per SM   [%s]\n", what, per, cudaGetErrorString(e));
    else printf("%-34s %7.1f clk per (16-B load + 16-B store) per warp; %5.1f B/clk/SM both ways, 4 warps per SM   [%s]\n", what, per,
                4.0 * 32 * 32 / per, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
    (void)warps_note;
}

int main() {
    run<0, false>("shared memory, dependent chain", 4);
    run<0, true>("tensor memory, dependent chain", 4);
    run<1, false>("shared memory, 16 in flight", 4);
    run<1, true>("tensor memory, 16 in flight", 4);
    return 0;
}
