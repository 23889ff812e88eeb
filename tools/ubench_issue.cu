// ubench_issue.cu -- does a half-rate (fp64) or quarter-rate (XU: F2F) instruction block the
// SMSP's dispatch port for more than one cycle?  Times independent-chain mixes on 1 warp per SMSP.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_issue ubench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
template <int MODE>
__global__ void k_mix(double* out, int iters, long long* cycles) {
    double d[ILP]; float f[ILP]; float g[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i] = 1.0 + threadIdx.x * 1e-3 + i; f[i] = 1.0f + i; g[i] = 0.5f + i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0 || MODE == 1 || MODE == 4) asm volatile("fma.rn.f64 %0, %0, 0d3FEFFBE76C8B4396, 0d3F50624DD2F1A9FC;" : "+d"(d[i]));
                if (MODE == 1 || MODE == 3 || MODE == 5) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(f[i]));
                if (MODE == 2 || MODE == 3 || MODE == 4) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(g[i])); }
                if (MODE == 6) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(g[i]) : "d"(d[i])); }
                if (MODE == 7) { asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(f[i])); asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(g[i])); asm volatile("fma.rn.f64 %0, %0, 0d3FEFFBE76C8B4396, 0d3F50624DD2F1A9FC;" : "+d"(d[i])); }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i] + f[i] + g[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, int n_per_group, int warps, double* d_out, long long* d_cyc) {
    const int iters = 2000;
    k_mix<MODE><<<148, 32 * warps>>>(d_out, 10, d_cyc);
    k_mix<MODE><<<148, 32 * warps>>>(d_out, iters, d_cyc);
    cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double groups = (double)iters * 4 * ILP;           // per warp
    printf("%-28s warps/SM=%2d: %.2f clk per group of %d instr (per SMSP: %.2f clk/group/warp)\n", name, warps,
           (double)cyc / groups, n_per_group, (double)cyc / groups / ((warps + 3) / 4));
}

int main() {
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 8 * 148 * 1024); cudaMalloc(&d_cyc, 8);
    for (int w : {4, 8, 16}) {
        run<0>("DFMA", 1, w, d_out, d_cyc);
        run<5>("FFMA", 1, w, d_out, d_cyc);
        run<1>("DFMA + FFMA", 2, w, d_out, d_cyc);
        run<7>("DFMA + 2 FFMA", 3, w, d_out, d_cyc);
        run<2>("F2F.F64.F32", 1, w, d_out, d_cyc);
        run<6>("F2F.F32.F64", 1, w, d_out, d_cyc);
        run<3>("F2F.F64.F32 + FFMA", 2, w, d_out, d_cyc);
        run<4>("DFMA + F2F.F64.F32", 2, w, d_out, d_cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
