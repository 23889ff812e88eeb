// ubench_ifetch.cu -- how many instructions per clock can an SM FETCH when its warps walk straight-line code that does
// not fit the instruction caches?  (k_env_t's tile loop is ~98 KB of straight-line code, 8 single-warp CTAs per SM, each
// at its own phase.)  Body = groups of 8 independent FFMAs (ILP 8: one warp alone can issue ~1 per clock); sizes 8 KB ..
// 192 KB; w single-warp CTAs per SM, started out of phase; or ONE CTA of w warps that meets at a barrier every SYNC groups
// (shared instruction stream).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_ifetch ubench_ifetch.cu
#include <cstdio>
#include <cuda_runtime.h>

#define G8 asm volatile("fma.rn.f32 %0, %0, %8, %9;\n\tfma.rn.f32 %1, %1, %8, %9;\n\tfma.rn.f32 %2, %2, %8, %9;\n\tfma.rn.f32 %3, %3, %8, %9;\n\t" \
                        "fma.rn.f32 %4, %4, %8, %9;\n\tfma.rn.f32 %5, %5, %8, %9;\n\tfma.rn.f32 %6, %6, %8, %9;\n\tfma.rn.f32 %7, %7, %8, %9;" \
                        : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]) : "f"(c1), "f"(c2));
#define R2(x) x x
#define R4(x) R2(R2(x))
#define R16(x) R4(R4(x))
#define R64(x) R4(R16(x))

// GROUPS64 blocks of 64 groups (= 512 instructions = 8 KB) per loop trip
template <int BLOCKS, bool SYNC>
__global__ void k_body(float* out, int iters, long long* cyc, int stagger) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0f + threadIdx.x * 1e-3f + i;
    const float c1 = 0.999f, c2 = 1e-3f;
    if (stagger > 0) {                                   // out of phase: CTA k of an SM starts k * stagger ns later
        const unsigned ns = (blockIdx.x / 148) * stagger;
        if (ns) __nanosleep(ns);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < BLOCKS; ++b) {
            R64(G8)
            if (SYNC) __syncthreads();
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

template <int BLOCKS, bool SYNC>
void run(int warps, bool one_cta, float* d_out, long long* d_cyc, long long* h_cyc) {
    const int iters = 4000 / BLOCKS + 4;
    const int grid = one_cta ? 148 : 148 * warps, block = one_cta ? 32 * warps : 32;
    k_body<BLOCKS, SYNC><<<grid, block>>>(d_out, 2, d_cyc, 0);
    k_body<BLOCKS, SYNC><<<grid, block>>>(d_out, iters, d_cyc, one_cta ? 0 : 700);
    cudaDeviceSynchronize();
    const int nw = 148 * warps;
    cudaMemcpy(h_cyc, d_cyc, 8 * nw, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < nw; ++i) mean += (double)h_cyc[i]; mean /= nw;
    const double instr = (double)iters * BLOCKS * 512;
    printf("body %4d KB  %s warps/SM=%2d%s: %.2f clk per instruction per warp, %.2f instr/clk/SM\n", BLOCKS * 8,
           one_cta ? "one CTA " : "1-warp CTAs", warps, SYNC ? " barrier every 8 KB" : "", mean / instr, warps * instr / mean);
}

int main() {
    float* d_out; long long *d_cyc, *h_cyc = new long long[148 * 32];
    cudaMalloc(&d_out, 4 * 148 * 1024 * 4); cudaMalloc(&d_cyc, 8 * 148 * 32);
    for (int w : {1, 2, 4, 8}) {
        run<1, false>(w, false, d_out, d_cyc, h_cyc);
        run<3, false>(w, false, d_out, d_cyc, h_cyc);
        run<6, false>(w, false, d_out, d_cyc, h_cyc);
        run<12, false>(w, false, d_out, d_cyc, h_cyc);
        run<24, false>(w, false, d_out, d_cyc, h_cyc);
    }
    for (int w : {2, 4, 8}) {
        run<12, false>(w, true, d_out, d_cyc, h_cyc);
        run<12, true>(w, true, d_out, d_cyc, h_cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
