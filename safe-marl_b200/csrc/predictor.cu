// predictor.cu -- safety-signal voltage predictor + safety penalty on tcgen05 tensor cores,
// writing straight into a device-resident replay ring (BASELINE config 4).
//
//   fp_predictor_load / fp_predict   replace `model.predict` of the reference's regressor
//       (safety_signal/train_safety_signal_model.py:33-46,73: MinMax-scaled X[66] -> 33 linear
//        outputs; interleaved input layout [P1,Q1,...,P33,Q33], data_generation.py:48-50) and the
//       slack penalty of the safety layer at fixed actions
//       (madrl/models/safemaddpg.py:205-229,264-277: 1000 * sum(max(0, v_min - V) + max(0, V - v_max)))
//   fp_replay_*                      replace TransReplayBuffer (utils/replay_buffer.py:3-30)
//
// Kernel k_predict (one persistent CTA of 128 threads per SM, one 128-env tile per iteration):
//   1. TMA: `cp.async.bulk` brings the tile's X rows (128 x 66 fp32 = 33 792 contiguous bytes)
//      into a double-buffered staging area, completion on an mbarrier; the copy of tile i+1
//      overlaps the MMA + epilogue of tile i.
//   2. 3xTF32 split: thread t owns row t; x = hi + lo with hi exactly representable in TF32
//      (low 13 mantissa bits cleared) and lo = x - hi (exact).  hi and lo are written in the
//      UMMA K-major no-swizzle ("interleave") layout: core matrices of 8 rows x 16 bytes,
//      [k-chunk][row][4 floats], SBO = 128 B, LBO = rows * 16 B.
//   3. tcgen05.mma (kind::tf32, M=128, N=48, K=8 per instruction; 9 k-steps x 3 products
//      hi*hi + lo*hi + hi*lo), issued by ONE thread, accumulating fp32 in TMEM; completion via
//      tcgen05.commit -> mbarrier.  The weights (hi and lo parts, same layout) sit in shared
//      memory for the lifetime of the CTA.
//   4. Epilogue: tcgen05.ld (32x32b.x16) hands every thread the 48 accumulators of its env;
//      add the bias, evaluate the slack penalty in fp64, stage Vhat in shared memory and store
//      coalesced rows into the replay ring at (pos + env) mod capacity (and/or a dense output).
// By shape (K = 66, N = 33: ~11 flop/B) the op is HBM-bound; the tensor pipe is nearly idle by
// construction and the point of tcgen05 here is to take the contraction off the fp32 pipe.
#include <cmath>
#include <cstring>

#include "flex_kernels.cuh"
#include "predictor.cuh"

// implemented in flex_api.cu
PredictorState* fp_internal_predictor(FpHandle* h);
int fp_internal_fail(FpHandle* h, int code, const char* msg);
void fp_internal_count_launch(FpHandle* h, int k);
int fp_internal_device(FpHandle* h);

namespace {

constexpr int N_IN = 66, N_OUT = 33;
constexpr uint32_t A_BYTES = PRED_KC * PRED_M * 16;          // 36 864
constexpr uint32_t B_BYTES = PRED_KC * PRED_N * 16;          // 13 824
constexpr uint32_t STAGE_BYTES = PRED_M * N_IN * 4;          // 33 792
constexpr uint32_t OFF_AHI = 0, OFF_ALO = OFF_AHI + A_BYTES, OFF_BHI = OFF_ALO + A_BYTES, OFF_BLO = OFF_BHI + B_BYTES;
constexpr uint32_t OFF_STAGE = OFF_BLO + B_BYTES;            // two stages
constexpr uint32_t OFF_OUT = OFF_STAGE + 2 * STAGE_BYTES;    // [128][33] fp32
constexpr uint32_t OFF_BIAS = OFF_OUT + PRED_M * N_OUT * 4;     // 16 896 B: keeps OFF_PEN 8-byte aligned
constexpr uint32_t OFF_PEN = OFF_BIAS + PRED_N * 4;          // [2][128] fp64 penalty partials (column halves)
constexpr uint32_t OFF_BAR = OFF_PEN + 2 * PRED_M * 8;       // 3 mbarriers
constexpr uint32_t OFF_TMEM = OFF_BAR + 3 * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16;
constexpr uint32_t TMEM_COLS = 64;                           // power of two >= 48

struct PredParams {
    const float* X; int64_t n;
    const float* B; const float* bias;
    float* vhat; double* penalty;                            // dense outputs (may be null)
    float* ring_vhat; float* ring_pen; int64_t ring_cap, ring_pos;   // replay sink (may be null)
    double v_min, v_max, w;
    float lo_f, hi_f;                                        // smallest float >= v_min, largest float <= v_max
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a descriptor or protocol mistake must surface as a trap, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle: start address, LBO (between the two
// 16-byte K chunks of one MMA), SBO (between 8-row groups), version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor: D fp32, A/B tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PRED_N >> 3) << 17) | ((uint32_t)(PRED_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

constexpr int PRED_THREADS = 2 * PRED_M;     // two threads per env row: K halves in the split, column halves in the epilogue

__global__ void __launch_bounds__(PRED_THREADS, 1) k_predict(const PredParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* A_hi = reinterpret_cast<float*>(smem + OFF_AHI);
    float* A_lo = reinterpret_cast<float*>(smem + OFF_ALO);
    float* Bsm = reinterpret_cast<float*>(smem + OFF_BHI);              // hi then lo, contiguous
    float* out = reinterpret_cast<float*>(smem + OFF_OUT);
    float* bias = reinterpret_cast<float*>(smem + OFF_BIAS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const uint32_t bar_full0 = smem_u32(smem + OFF_BAR), bar_mma = bar_full0 + 16;
    double* pen_part = reinterpret_cast<double*>(smem + OFF_PEN);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (PRED_M - 1), half = tid >> 7;               // env row of the tile, which half of the work
    const int64_t n_tiles = (prm.n + PRED_M - 1) / PRED_M;

    // ---- one-time setup: barriers, TMEM, weights, zero K padding
    if (tid == 0) {
        mbar_init(bar_full0, 1); mbar_init(bar_full0 + 8, 1); mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < (int)(2 * B_BYTES / 4); i += PRED_THREADS) Bsm[i] = prm.B[i];
    if (tid < PRED_N) bias[tid] = prm.bias[tid];
    if (half == 0) {   // columns 66..71 of every row stay zero for the lifetime of the CTA
        float* hz = A_hi + ((N_IN >> 2) * PRED_M + row) * 4;            // chunk 16: k = 64..67
        float* lz = A_lo + ((N_IN >> 2) * PRED_M + row) * 4;
        hz[2] = hz[3] = 0.0f; lz[2] = lz[3] = 0.0f;
        float4* h4 = reinterpret_cast<float4*>(A_hi + ((PRED_KC - 1) * PRED_M + row) * 4);   // chunk 17: k = 68..71
        float4* l4 = reinterpret_cast<float4*>(A_lo + ((PRED_KC - 1) * PRED_M + row) * 4);
        *h4 = make_float4(0.f, 0.f, 0.f, 0.f); *l4 = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto rows_of = [&](int64_t tl) -> int { const int64_t r = prm.n - tl * PRED_M; return (int)(r < PRED_M ? r : PRED_M); };
    auto tma_ok = [&](int64_t tl) -> bool { return ((rows_of(tl) * N_IN * 4) & 15) == 0; };
    auto issue = [&](int64_t tl, int s) {                               // thread 0 only
        const uint32_t bytes = (uint32_t)(rows_of(tl) * N_IN * 4);
        const uint32_t bar = bar_full0 + 8 * s;
        mbar_expect_tx(bar, bytes);
        tma_load_1d(smem_u32(smem + OFF_STAGE + s * STAGE_BYTES), prm.X + tl * (int64_t)(PRED_M * N_IN), bytes, bar);
    };

    uint32_t ph_full[2] = {0u, 0u}, ph_mma = 0u;
    int64_t tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles && tma_ok(tile)) issue(tile, 0);
    for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const int64_t next = tile + gridDim.x;
        const int rows = rows_of(tile);
        float* stage = reinterpret_cast<float*>(smem + OFF_STAGE + s * STAGE_BYTES);
        if (tid == 0 && next < n_tiles && tma_ok(next)) issue(next, s ^ 1);   // overlaps this tile's MMA + epilogue
        if (tma_ok(tile)) {
            mbar_wait(bar_full0 + 8 * s, ph_full[s]);
            ph_full[s] ^= 1u;
        } else {                                                        // odd-sized last tile: plain loads
            const float* src = prm.X + tile * (int64_t)(PRED_M * N_IN);
            for (int i = tid; i < rows * N_IN; i += PRED_THREADS) stage[i] = src[i];
            __syncthreads();
        }

        // ---- 3xTF32 split into the UMMA layout (two threads per row: pairs 0..16 / 17..32)
        if (row < rows) {
            const float2* src = reinterpret_cast<const float2*>(stage + row * N_IN);
            const int j0 = half ? 17 : 0, j1 = half ? N_IN / 2 : 17;
#pragma unroll
            for (int jj = 0; jj < 17; ++jj) {
                const int j = j0 + jj;
                if (j < j1) {
                    const float2 x = src[j];
                    const float hx = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                    const float hy = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                    const int k = 2 * j, off = ((k >> 2) * PRED_M + row) * 4 + (k & 3);
                    *reinterpret_cast<float2*>(A_hi + off) = make_float2(hx, hy);
                    *reinterpret_cast<float2*>(A_lo + off) = make_float2(x.x - hx, x.y - hy);
                }
            }
        }
        fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core
        tc_fence_before();
        __syncthreads();

        // ---- MMA: one thread issues 9 k-steps x 3 products into TMEM
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(Bsm), b_lo = b_hi + B_BYTES;
#pragma unroll
            for (int ks = 0; ks < PRED_K / 8; ++ks) {
                const uint32_t ao = ks * 2 * (PRED_M * 16), bo = ks * 2 * (PRED_N * 16);
                const uint64_t dah = umma_desc(a_hi + ao, PRED_M * 16, 128), dal = umma_desc(a_lo + ao, PRED_M * 16, 128);
                const uint64_t dbh = umma_desc(b_hi + bo, PRED_N * 16, 128), dbl = umma_desc(b_lo + bo, PRED_N * 16, 128);
                umma_tf32(tmem_base, dah, dbh, ks > 0 ? 1u : 0u);
                umma_tf32(tmem_base, dal, dbh, 1u);
                umma_tf32(tmem_base, dah, dbl, 1u);
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tc_fence_after();

        // ---- epilogue: TMEM lane = env row; warps 0-3 take columns 0..16, warps 4-7 columns 17..32
        // (a warp may only touch the TMEM lane quarter 32 * (warp % 4))
        {
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (half ? 17u : 0u);
            uint32_t r0[16], r1 = 0u;
            tmem_ld16(taddr, r0);
            if (half == 0) tmem_ld1(taddr + 16, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            double pen = 0.0;
            if (row < rows) {
                const int c0 = half ? 17 : 0;
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    if (i < 16 || half == 0) {
                        const float V = __uint_as_float(i < 16 ? r0[i] : r1) + bias[c0 + i];
                        out[row * N_OUT + c0 + i] = V;
                        // in-limit voltages (the common case) contribute exactly zero: the fp64 terms are
                        // only evaluated outside the conservative fp32 bounds lo_f >= v_min, hi_f <= v_max
                        if (V < prm.lo_f || V > prm.hi_f) {
                            const double under = prm.v_min - (double)V, over = (double)V - prm.v_max;
                            pen = pen + ((under > 0.0 ? under : 0.0) + (over > 0.0 ? over : 0.0));
                        }
                    }
                }
            }
            pen_part[half * PRED_M + row] = pen;
        }
        tc_fence_before();
        __syncthreads();
        const int64_t e = tile * PRED_M + row;
        if (half == 0 && row < rows) {
            const double pen = prm.w * (pen_part[row] + pen_part[PRED_M + row]);
            if (prm.penalty != nullptr) prm.penalty[e] = pen;
            if (prm.ring_pen != nullptr) {               // pos < cap and n <= cap: one conditional subtraction wraps
                const int64_t rr = prm.ring_pos + e;
                prm.ring_pen[rr >= prm.ring_cap ? rr - prm.ring_cap : rr] = (float)pen;
            }
        }

        // ---- coalesced row stores: dense output and/or replay ring (wraps at capacity)
        // (the tile's rows are contiguous in the dense output; in the ring they are contiguous up to
        //  one wrap, so element i of the tile lands at i + base, minus the ring size past the end)
        const int64_t e0 = tile * PRED_M;
        const int64_t ring_elems = prm.ring_cap * N_OUT, ring_base = (prm.ring_pos + e0) * N_OUT;
        const int n_el = rows * N_OUT;
        float* dense = prm.vhat != nullptr ? prm.vhat + e0 * N_OUT : nullptr;
        // 16-byte stores where the destination allows it: a full tile is 1056 float4 (e0 * 33 * 4 bytes is
        // a multiple of 16); the ring segment qualifies when it does not wrap and starts 16-byte aligned
        const bool ring_vec = prm.ring_vhat != nullptr && ((ring_base & 3) == 0) && (ring_base + n_el <= ring_elems);
        if ((n_el & 3) == 0 && (dense == nullptr || ((reinterpret_cast<uintptr_t>(dense) & 15) == 0)) &&
            (prm.ring_vhat == nullptr || (ring_vec && ((reinterpret_cast<uintptr_t>(prm.ring_vhat) & 15) == 0)))) {
            const float4* o4 = reinterpret_cast<const float4*>(out);
            float4* d4 = reinterpret_cast<float4*>(dense);
            float4* r4 = prm.ring_vhat != nullptr ? reinterpret_cast<float4*>(prm.ring_vhat + ring_base) : nullptr;
            for (int i = tid; i < (n_el >> 2); i += PRED_THREADS) {
                const float4 v = o4[i];
                if (d4 != nullptr) d4[i] = v;
                if (r4 != nullptr) r4[i] = v;
            }
        } else {
            for (int i = tid; i < n_el; i += PRED_THREADS) {
                const float v = out[i];
                if (dense != nullptr) dense[i] = v;
                if (prm.ring_vhat != nullptr) {
                    int64_t o = ring_base + i;
                    o = o >= ring_elems ? o - ring_elems : o;
                    prm.ring_vhat[o] = v;
                }
            }
        }
        // the next iteration's transform does not touch `out`; its epilogue is two barriers away
    }

    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------- replay ring
// n consecutive ring rows starting at `pos` are contiguous in memory up to one wrap, so both
// copies are flat, fully coalesced and free of integer division: element i <-> ring element
// pos * width + i, minus the ring size past the end.
__global__ void k_replay_write(float* __restrict__ dst, int64_t cap, int width, int64_t pos, int64_t n,
                               const float* __restrict__ src) {
    const int64_t total = n * width, ring = cap * width, base = pos * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = base + i;
        o = o >= ring ? o - ring : o;
        dst[o] = src[i];
    }
}

__global__ void k_replay_gather(const float* __restrict__ field, int64_t cap, int width, int64_t first, int64_t batch,
                                float* __restrict__ out) {
    const int64_t total = batch * width, ring = cap * width, base = first * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = base + i;
        o = o >= ring ? o - ring : o;
        out[i] = field[o];
    }
}

int grid_for_elems(int64_t total) {
    int64_t g = (total + 255) / 256;
    if (g < 1) g = 1;
    return (int)(g > 148 * 16 ? 148 * 16 : g);
}

thread_local std::string g_replay_err;
int rfail(FpReplay* r, int code, const std::string& msg) {
    if (r) r->err = msg; else g_replay_err = msg;
    return code;
}

}  // namespace

void predictor_free(PredictorState* p) {
    if (!p) return;
    cudaFree(p->d_B); cudaFree(p->d_bias);
    p->d_B = nullptr; p->d_bias = nullptr; p->loaded = 0;
}

extern "C" {

int fp_predictor_load(FpHandle* h, int32_t n_in, int32_t n_out, const double* h_A, const double* h_c, double v_min,
                      double v_max, double slack_weight) {
    if (!h) return FP_EINVAL;
    if (n_in != N_IN || n_out != N_OUT || !h_A || !h_c)
        return fp_internal_fail(h, FP_EINVAL, "fp_predictor_load: expects a 66 -> 33 affine model (2*n_bus inputs, n_bus outputs)");
    PredictorState* p = fp_internal_predictor(h);
    cudaSetDevice(fp_internal_device(h));
    predictor_free(p);
    // weights: fp64 -> (hi, lo) TF32-representable pair, UMMA K-major layout [part][k-chunk][n][4], zero padded
    std::vector<float> B(2 * PRED_KC * PRED_N * 4, 0.0f), bias(PRED_N, 0.0f);
    for (int n = 0; n < N_OUT; ++n) {
        for (int k = 0; k < N_IN; ++k) {
            const double a = h_A[(size_t)n * N_IN + k];
            float hi = (float)a;
            uint32_t u; std::memcpy(&u, &hi, 4); u &= 0xFFFFE000u; std::memcpy(&hi, &u, 4);
            const float lo = (float)(a - (double)hi);
            const size_t off = ((size_t)(k >> 2) * PRED_N + n) * 4 + (k & 3);
            B[off] = hi; B[(size_t)PRED_KC * PRED_N * 4 + off] = lo;
        }
        bias[n] = (float)h_c[n];
    }
    if (cudaMalloc(&p->d_B, B.size() * 4) != cudaSuccess || cudaMalloc(&p->d_bias, bias.size() * 4) != cudaSuccess)
        return fp_internal_fail(h, FP_ENOMEM, "fp_predictor_load: cudaMalloc failed");
    cudaMemcpy(p->d_B, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    cudaError_t e = cudaFuncSetAttribute(k_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return fp_internal_fail(h, FP_ECUDA, cudaGetErrorString(e));
    p->n_in = n_in; p->n_out = n_out; p->v_min = v_min; p->v_max = v_max; p->slack_weight = slack_weight;
    p->loaded = 1;
    return FP_OK;
}

int fp_predict(FpHandle* h, int64_t n, const float* d_X, float* d_vhat, double* d_penalty, FpReplay* sink,
               int32_t f_vhat, int32_t f_penalty, int64_t pos, void* stream) {
    if (!h) return FP_EINVAL;
    PredictorState* p = fp_internal_predictor(h);
    if (!p->loaded) return fp_internal_fail(h, FP_ESTATE, "fp_predict: call fp_predictor_load first");
    if (n < 1 || !d_X) return fp_internal_fail(h, FP_EINVAL, "fp_predict: bad arguments");
    if (((uintptr_t)d_X & 15) != 0) return fp_internal_fail(h, FP_EINVAL, "fp_predict: X must be 16-byte aligned (TMA)");
    PredParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.X = d_X; prm.n = n; prm.B = p->d_B; prm.bias = p->d_bias; prm.vhat = d_vhat; prm.penalty = d_penalty;
    prm.v_min = p->v_min; prm.v_max = p->v_max; prm.w = p->slack_weight;
    prm.lo_f = (float)p->v_min; if ((double)prm.lo_f < p->v_min) prm.lo_f = std::nextafterf(prm.lo_f, INFINITY);
    prm.hi_f = (float)p->v_max; if ((double)prm.hi_f > p->v_max) prm.hi_f = std::nextafterf(prm.hi_f, -INFINITY);
    prm.ring_cap = 1; prm.ring_pos = 0;
    if (sink) {
        const int nf = (int)sink->widths.size();
        if (n > sink->capacity || pos < 0 || pos >= sink->capacity)
            return fp_internal_fail(h, FP_EINVAL, "fp_predict: rows do not fit the replay ring");
        if (f_vhat >= 0) {
            if (f_vhat >= nf || sink->widths[f_vhat] != N_OUT) return fp_internal_fail(h, FP_EINVAL, "fp_predict: Vhat field must have width 33");
            prm.ring_vhat = sink->d_fields[f_vhat];
        }
        if (f_penalty >= 0) {
            if (f_penalty >= nf || sink->widths[f_penalty] != 1) return fp_internal_fail(h, FP_EINVAL, "fp_predict: penalty field must have width 1");
            prm.ring_pen = sink->d_fields[f_penalty];
        }
        prm.ring_cap = sink->capacity; prm.ring_pos = pos;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + PRED_M - 1) / PRED_M;
    const int grid = (int)(tiles < sms ? tiles : sms);
    k_predict<<<grid, PRED_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(prm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fp_internal_fail(h, FP_ECUDA, cudaGetErrorString(e));
    fp_internal_count_launch(h, 1);
    return FP_OK;
}

// ------------------------------------------------------------------ replay ring (TransReplayBuffer)
int fp_replay_create(int64_t capacity, int32_t n_fields, const int32_t* widths, int device, FpReplay** out) {
    if (!out || capacity < 1 || n_fields < 1 || n_fields > 64 || !widths) return rfail(nullptr, FP_EINVAL, "fp_replay_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return rfail(nullptr, FP_ECUDA, "fp_replay_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return rfail(nullptr, FP_EINVAL, "fp_replay_create: bad device index");
    FpReplay* r = new FpReplay();
    r->capacity = capacity; r->device = device;
    cudaSetDevice(device);
    for (int f = 0; f < n_fields; ++f) {
        if (widths[f] < 1) { fp_replay_destroy(r); return rfail(nullptr, FP_EINVAL, "fp_replay_create: field width must be >= 1"); }
        float* p = nullptr;
        if (cudaMalloc(&p, (size_t)capacity * widths[f] * 4) != cudaSuccess) { fp_replay_destroy(r); return rfail(nullptr, FP_ENOMEM, "fp_replay_create: cudaMalloc failed"); }
        cudaMemset(p, 0, (size_t)capacity * widths[f] * 4);
        r->widths.push_back(widths[f]); r->d_fields.push_back(p);
    }
    *out = r;
    return FP_OK;
}

int fp_replay_destroy(FpReplay* r) {
    if (!r) return FP_OK;
    cudaSetDevice(r->device);
    for (float* p : r->d_fields) cudaFree(p);
    delete r;
    return FP_OK;
}

const char* fp_replay_last_error(const FpReplay* r) { return r ? r->err.c_str() : g_replay_err.c_str(); }
int64_t fp_replay_len(const FpReplay* r) { return r ? r->len : 0; }
int64_t fp_replay_capacity(const FpReplay* r) { return r ? r->capacity : 0; }
int fp_replay_clear(FpReplay* r) { if (!r) return FP_EINVAL; r->head = 0; r->len = 0; return FP_OK; }     // :29-30

// add_experience (:23-27) for n rows at once: the oldest rows are overwritten when the ring is full
int fp_replay_reserve(FpReplay* r, int64_t n, int64_t* pos) {
    if (!r || !pos) return FP_EINVAL;
    if (n < 1 || n > r->capacity) return rfail(r, FP_EINVAL, "fp_replay_reserve: 1 <= n <= capacity required");
    *pos = r->head;
    r->head = (r->head + n) % r->capacity;
    r->len = (r->len + n < r->capacity) ? r->len + n : r->capacity;
    return FP_OK;
}

int fp_replay_write(FpReplay* r, int32_t field, int64_t pos, int64_t n, const float* d_src, void* stream) {
    if (!r) return FP_EINVAL;
    if (field < 0 || field >= (int)r->widths.size() || !d_src || n < 1 || n > r->capacity || pos < 0 || pos >= r->capacity)
        return rfail(r, FP_EINVAL, "fp_replay_write: bad arguments");
    const int w = r->widths[field];
    k_replay_write<<<grid_for_elems(n * w), 256, 0, (cudaStream_t)stream>>>(r->d_fields[field], r->capacity, w, pos, n, d_src);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FP_OK : rfail(r, FP_ECUDA, cudaGetErrorString(e));
}

// get_batch (:14-21): `batch` consecutive rows from logical index `start` (0 = oldest)
int fp_replay_sample(FpReplay* r, int64_t start, int64_t batch, float* const* d_out, void* stream) {
    if (!r || !d_out) return FP_EINVAL;
    if (batch < 1 || start < 0 || start + batch > r->len) return rfail(r, FP_EINVAL, "fp_replay_sample: window outside the stored rows");
    const int64_t first = ((r->head - r->len + start) % r->capacity + r->capacity) % r->capacity;
    for (size_t f = 0; f < r->widths.size(); ++f) {
        if (!d_out[f]) continue;
        const int w = r->widths[f];
        k_replay_gather<<<grid_for_elems(batch * w), 256, 0, (cudaStream_t)stream>>>(r->d_fields[f], r->capacity, w, first, batch, d_out[f]);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FP_OK : rfail(r, FP_ECUDA, cudaGetErrorString(e));
}

int fp_replay_field_ptr(FpReplay* r, int32_t field, float** d_ptr) {
    if (!r || !d_ptr || field < 0 || field >= (int)r->widths.size()) return FP_EINVAL;
    *d_ptr = r->d_fields[field];
    return FP_OK;
}

}  // extern "C"
