// ubench_fp64.cu -- B200 fp64 pipe microbenchmark: dependent-issue latency and per-SM throughput
// of DFMA as a function of resident warps and independent chains per thread.  Informs how much
// ILP x TLP the DistFlow sweep needs to keep the fp64 pipe busy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64 ubench_fp64.cu && ./ubench_fp64
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b, long long* cycles) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// float <-> double conversion chain (the reciprocal-seed path): asm volatile keeps every conversion
__global__ void k_cvt(double* out, int iters, long long* cycles) {
    double x = 1.0 + threadIdx.x * 1e-3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            float f;
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x));
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x) : "f"(f));
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// dependent fp32 FMA chain
__global__ void k_ffma(double* out, int iters, long long* cycles) {
    float x = 1.0f + threadIdx.x * 1e-3f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(x));
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// dependent shared-memory load chain (pointer chase), 8-byte
__global__ void k_lds(double* out, int iters, long long* cycles) {
    __shared__ long long next[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) next[i] = (i + 33) & 1023;
    __syncthreads();
    long long p = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) p = next[p];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (double)p;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, double* d_out, long long* d_cyc) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2000;
    // one CTA per SM with warps_per_sm warps
    k_dfma<ILP><<<sms, 32 * warps_per_sm>>>(d_out, 10, 0.999, 0.001, d_cyc);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_dfma<ILP><<<sms, 32 * warps_per_sm>>>(d_out, iters, 0.999, 0.001, d_cyc);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double n_inst = (double)iters * 16 * ILP;           // per warp
    const double lanes = n_inst * 32 * warps_per_sm / (double)cyc;   // fp64 lane-ops per clk per SM
    printf("DFMA warps/SM=%2d ILP=%d: %.2f clk per dependent step, %.1f lane-ops/clk/SM, %.3f ms, %.2f TFLOP/s\n",
           warps_per_sm, ILP, (double)cyc / (iters * 16.0), lanes, ms,
           2.0 * n_inst * 32 * warps_per_sm * sms / (ms * 1e-3) / 1e12);
}

int main() {
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 8 * 148 * 1024 * 4); cudaMalloc(&d_cyc, 8);
    const int ws[] = {1, 4, 8, 12, 16, 32};
    for (int w : ws) { run<1>(w, d_out, d_cyc); run<2>(w, d_out, d_cyc); run<4>(w, d_out, d_cyc); run<8>(w, d_out, d_cyc); }
    long long cyc;
    k_cvt<<<148, 32>>>(d_out, 1000, d_cyc); cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("F2F.F32.F64 + F2F.F64.F32 dependent pair: %.2f clk per pair\n", (double)cyc / 16000.0);
    k_ffma<<<148, 32>>>(d_out, 1000, d_cyc); cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("FFMA dependent chain: %.2f clk per op\n", (double)cyc / 16000.0);
    k_lds<<<148, 32>>>(d_out, 1000, d_cyc); cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.64 dependent chain: %.2f clk per load\n", (double)cyc / 16000.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
