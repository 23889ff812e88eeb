// ubench_ffma2.cu -- issue rate and dependent latency of fma.rn.f32x2 (FFMA2, sm_100+) against scalar fma.rn.f32.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_ffma2 ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int ILP>
__global__ void k(float* out, int iters, long long* cyc) {
    float f[2 * ILP];
    unsigned long long p[ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) f[i] = 1.0f + threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(f[2 * i]), "f"(f[2 * i + 1]));
    unsigned long long c1, c2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(c1) : "f"(0.999f));
    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(1e-3f));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(f[i]));
                else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c1), "l"(c2));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); s += a + b + f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int warps, float* d_out, long long* d_cyc) {
    const int iters = 4000;
    k<MODE, ILP><<<148, 32 * warps>>>(d_out, 10, d_cyc);
    k<MODE, ILP><<<148, 32 * warps>>>(d_out, iters, d_cyc);
    cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-8s ILP %d warps/SM %2d: %.2f clk per instruction per warp (%.2f instr/clk/SMSP)\n", name, ILP, warps,
           (double)c / (iters * 8.0 * ILP), (iters * 8.0 * ILP) * ((warps + 3) / 4) / (double)c);
}
int main() {
    float* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 4 * 148 * 1024); cudaMalloc(&d_cyc, 8);
    for (int w : {4, 8, 16}) {
        run<0, 1>("FFMA", w, d_out, d_cyc); run<1, 1>("FFMA2", w, d_out, d_cyc);
        run<0, 2>("FFMA", w, d_out, d_cyc); run<1, 2>("FFMA2", w, d_out, d_cyc);
        run<0, 8>("FFMA", w, d_out, d_cyc); run<1, 8>("FFMA2", w, d_out, d_cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
