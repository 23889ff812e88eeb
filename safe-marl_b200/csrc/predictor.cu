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
// Kernel k_predict: one persistent CTA per SM (17 warps), 128-env tiles, WARP-SPECIALISED so that the
// phases of consecutive tiles overlap.  The op is HBM-bound by shape (K = 66, N = 33: ~11 flop/B); the
// job of the kernel is to keep five tiles of X in flight per SM while everything else hides underneath,
// and to keep shared-memory traffic (the second bound: 128 B/clk/SM) to the X tile and the Vhat tile.
//   TMA/MMA warp 16 (one elected thread)
//     1. TMA: `cp.async.bulk` brings a tile's X rows (128 x 66 fp32 = 33 792 contiguous bytes) into one
//        of FIVE staging buffers, completion on an mbarrier; a buffer is refilled (five tiles ahead) as
//        soon as every producer has read it.
//     3. tcgen05.mma (kind::tf32, M=128, N=48, K=8 per instruction; 9 k-steps x 3 products
//        hi*hi + lo*hi + hi*lo) with the A OPERAND IN TMEM and the weights (hi and lo parts, UMMA K-major
//        no-swizzle layout: [k-chunk][n][4 floats], SBO = 128 B, LBO = 48 * 16 B) in shared memory for the
//        lifetime of the CTA; fp32 accumulation in one of TWO TMEM accumulators; completion via
//        tcgen05.commit -> mbarrier.
//   producer warps 0-7 (two threads per env row of the tile: K halves)
//     2. 3xTF32 split: x = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared)
//        and lo = x - hi (exact).  A thread owns a row, and a TMEM lane IS a row: `tcgen05.st.32x32b.x4`
//        writes hi and lo straight into one of TWO A buffers in TMEM (lane = env row, column = k;
//        144 columns per buffer), so A never touches shared memory (it would cost 70 KB of stores and
//        110 KB of operand reads per tile there) and the split of tile i+1 overlaps the MMAs of tile i.
//   epilogue warps 8-15 (two threads per env row: column halves; TMEM lane quarter = warp % 4)
//     4. tcgen05.ld (32x32b.x16 + .x1) hands every thread its 17 / 16 accumulators of its env and frees
//        the accumulator for the tile after next; add the bias, evaluate the slack penalty (fp64, only
//        for rows that leave the limits), stage Vhat in shared memory; ONE bulk store per destination
//        (TMA engine) writes the tile into the replay ring at (pos + env) mod capacity and/or a dense
//        output (a wrapping or unaligned segment falls back to a coalesced store loop).
// TMEM map (512 columns): accumulators at 0 and 64, A buffers at 128 and 272 (hi 72 | lo 72 columns).
// The weights arrive by one bulk copy that overlaps the first X tiles (own mbarrier).  The groups meet
// only through mbarriers (full[5], a_ready[2], mma_done[2], tmem_free[2]); the epilogue synchronises
// internally with a named barrier; every wait is bounded (trap, never a hung GPU).
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
constexpr uint32_t B_BYTES = PRED_KC * PRED_N * 16;          // 13 824
constexpr uint32_t STAGE_BYTES = PRED_M * N_IN * 4;          // 33 792
constexpr uint32_t OFF_BHI = 0, OFF_BLO = OFF_BHI + B_BYTES;
constexpr int N_STAGES = 5;                                  // X tiles in flight per SM (5 x 33 KB covers the HBM latency)
constexpr uint32_t OFF_STAGE = OFF_BLO + B_BYTES;
constexpr uint32_t OFF_OUT = OFF_STAGE + N_STAGES * STAGE_BYTES;   // [128][33] fp32
constexpr uint32_t OFF_BIAS = OFF_OUT + PRED_M * N_OUT * 4;
constexpr uint32_t OFF_BAR = OFF_BIAS + PRED_N * 4;          // full[5], mma_done[2], tmem_free[2], weights, a_ready[2]
constexpr uint32_t OFF_PEN = OFF_BAR + 12 * 8;                // [2][128] fp64 penalty partials (column halves)
constexpr uint32_t OFF_TMEM = OFF_PEN + 2 * PRED_M * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16;
constexpr uint32_t ACC_COLS = 64;                            // one accumulator (48 columns used)
constexpr uint32_t A_COL0 = 2 * ACC_COLS;                    // A operand in TMEM: [2 buffers][hi 72 | lo 72] columns, lane = env row
constexpr uint32_t A_COLS = 2 * PRED_K;
constexpr uint32_t TMEM_COLS = 512;                          // 2 accumulators + 2 A buffers = 416 columns -> the whole TMEM (one CTA per SM)
static_assert(SMEM_BYTES <= 227 * 1024, "predictor tile buffers exceed the shared memory of one SM");

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

// A operand from TMEM (lane = row, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
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
// Bulk shared -> global stores (TMA engine)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <int ID, int THREADS>
__device__ __forceinline__ void group_sync() {               // named barrier of one warp group
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(THREADS) : "memory");
}
__device__ __forceinline__ bool elect_one() {                // one lane of a converged warp (lets ptxas keep the
    uint32_t p;                                              // descriptors of the MMA on the uniform datapath)
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(p));
    return p != 0u;
}

constexpr int PROD_THREADS = 2 * PRED_M;     // warps 0-7: TMA + split (two threads per row) + MMA issue (warp 0)
constexpr int EPI_THREADS = 2 * PRED_M;      // warps 8-15: epilogue (two threads per row: column halves)
constexpr int MMA_WARP = (PROD_THREADS + EPI_THREADS) / 32;   // warp 16: TMA + tcgen05.mma issue
constexpr int PRED_THREADS = PROD_THREADS + EPI_THREADS + 32;

__global__ void __launch_bounds__(PRED_THREADS, 1) k_predict(const PredParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* Bsm = reinterpret_cast<float*>(smem + OFF_BHI);              // hi then lo, contiguous
    float* out = reinterpret_cast<float*>(smem + OFF_OUT);
    float* bias = reinterpret_cast<float*>(smem + OFF_BIAS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    double* pen_part = reinterpret_cast<double*>(smem + OFF_PEN);
    const uint32_t bar_full0 = smem_u32(smem + OFF_BAR), bar_mma0 = bar_full0 + 8 * N_STAGES, bar_free0 = bar_mma0 + 16,
                   bar_w = bar_free0 + 16, bar_a0 = bar_w + 8;
    const int tid = threadIdx.x, warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);   // warp-uniform for the compiler
    const int row = tid & (PRED_M - 1);                                 // env row of the tile (both groups)
    const int64_t n_tiles = (prm.n + PRED_M - 1) / PRED_M;

    auto rows_of = [&](int64_t tl) -> int { const int64_t r = prm.n - tl * PRED_M; return (int)(r < PRED_M ? r : PRED_M); };
    auto tma_ok = [&](int64_t tl) -> bool { return ((rows_of(tl) * N_IN * 4) & 15) == 0; };
    // this CTA's tiles: blockIdx.x + k * gridDim.x, k = 0 .. n_my - 1 (the grid never exceeds the tile count)
    const int n_my = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    auto tile_of = [&](int k) -> int64_t { return (int64_t)blockIdx.x + (int64_t)k * gridDim.x; };

    auto issue = [&](int k) {                                           // one thread only (MMA warp)
        const int64_t tl = tile_of(k);
        if (!tma_ok(tl)) return;                                        // odd-sized last tile: plain loads by the producers
        const uint32_t bytes = (uint32_t)(rows_of(tl) * N_IN * 4);
        const int s = k % N_STAGES;
        const uint32_t bar = bar_full0 + 8 * s;
        mbar_expect_tx(bar, bytes);
        tma_load_1d(smem_u32(smem + OFF_STAGE + s * STAGE_BYTES), prm.X + tl * (int64_t)(PRED_M * N_IN), bytes, bar);
    };
    // ---- one-time setup: barriers, TMEM, weights; the first TMA loads are issued before anything else
    if (tid == 0) {
        for (int s = 0; s < N_STAGES; ++s) mbar_init(bar_full0 + 8 * s, 1);
        mbar_init(bar_mma0, 1); mbar_init(bar_mma0 + 8, 1);
        mbar_init(bar_free0, EPI_THREADS); mbar_init(bar_free0 + 8, EPI_THREADS);
        mbar_init(bar_w, 1);
        mbar_init(bar_a0, PROD_THREADS); mbar_init(bar_a0 + 8, PROD_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, 2 * B_BYTES);                             // weights: one bulk copy, awaited before the first MMA
        tma_load_1d(smem_u32(Bsm), prm.B, 2 * B_BYTES, bar_w);
        for (int k = 0; k < N_STAGES && k < n_my; ++k) issue(k);        // the first X tiles travel during the rest of the setup
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < PRED_N) bias[tid] = prm.bias[tid];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // one tile's MMAs: 9 k-steps x 3 products (hi*hi + lo*hi + hi*lo), A from TMEM buffer `ab`
    auto mma_tile = [&](uint32_t acc, uint32_t ab) {
        const uint32_t b_hi = smem_u32(Bsm), b_lo = b_hi + B_BYTES;
        const uint32_t t_hi = tmem_base + A_COL0 + ab * A_COLS, t_lo = t_hi + PRED_K;
#pragma unroll
        for (int ks = 0; ks < PRED_K / 8; ++ks) {
            const uint32_t bo = ks * 2 * (PRED_N * 16);
            const uint64_t dbh = umma_desc(b_hi + bo, PRED_N * 16, 128), dbl = umma_desc(b_lo + bo, PRED_N * 16, 128);
            umma_tf32_ts(acc, t_hi + 8 * ks, dbh, ks > 0 ? 1u : 0u);
            umma_tf32_ts(acc, t_lo + 8 * ks, dbh, 1u);
            umma_tf32_ts(acc, t_hi + 8 * ks, dbl, 1u);
        }
    };

    if (warp == MMA_WARP) {
        // =============================================================== MMA warp: TMA issue, tcgen05.mma issue
        for (int k = 0; k < n_my; ++k) {
            const int b = k & 1;
            mbar_wait(bar_a0 + 8 * b, (uint32_t)((k >> 1) & 1));        // A buffer b holds tile k; its staging buffer is read
            if (k >= 2) mbar_wait(bar_free0 + 8 * b, (uint32_t)(((k >> 1) - 1) & 1));   // epilogue of tile k - 2 drained the accumulator
            if (k == 0) mbar_wait(bar_w, 0u);                           // the weights have landed
            tc_fence_after();
            __syncwarp();
            if (elect_one()) {
                mma_tile(tmem_base + (uint32_t)b * ACC_COLS, (uint32_t)b);
                umma_commit(bar_mma0 + 8 * b);
                if (k + N_STAGES < n_my) issue(k + N_STAGES);           // refill the staging buffer, N_STAGES tiles ahead -- AFTER the MMAs:
                                                                        // issuing the bulk copy first delays them by its issue latency
                                                                        // (k_policy: 315 -> 266 us from this order alone)
            }
            __syncwarp();
        }
    } else if (warp < PROD_THREADS / 32) {
        // =============================================================== producers: 3xTF32 split -> A operand in TMEM
        // thread = (env row, K half); a warp writes the TMEM lane quarter 32 * (warp % 4) = its rows
        const int half = tid >> 7;
        for (int k = 0; k < n_my; ++k) {
            const int s = k % N_STAGES, b = k & 1;
            const int64_t tile = tile_of(k);
            const int rows = rows_of(tile);
            float* stage = reinterpret_cast<float*>(smem + OFF_STAGE + s * STAGE_BYTES);
            if (tma_ok(tile)) {
                mbar_wait(bar_full0 + 8 * s, (uint32_t)((k / N_STAGES) & 1));
            } else {
                const float* src = prm.X + tile * (int64_t)(PRED_M * N_IN);
                for (int i = tid; i < rows * N_IN; i += PROD_THREADS) stage[i] = src[i];
                group_sync<1, PROD_THREADS>();
            }
            // the MMAs of tile k - 2 have read A buffer b
            if (k >= 2) mbar_wait(bar_mma0 + 8 * b, (uint32_t)(((k >> 1) - 1) & 1));
            tc_fence_after();
            const float2* src = reinterpret_cast<const float2*>(stage + row * N_IN);
            const uint32_t t_hi = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + A_COL0 + (uint32_t)b * A_COLS, t_lo = t_hi + PRED_K;
            // rows past the end of a partial tile split whatever the staging buffer holds: never stored
#pragma unroll
            for (int cc = 0; cc < 9; ++cc) {
                const int c = (half ? 9 : 0) + cc;                      // 4-k chunk; 16 holds k = 64, 65 and zeros, 17 zeros
                float2 x = make_float2(0.0f, 0.0f), y = make_float2(0.0f, 0.0f);
                if (2 * c < N_IN / 2) x = src[2 * c];
                if (2 * c + 1 < N_IN / 2) y = src[2 * c + 1];
                const float hx = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), hy = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                const float hz = __uint_as_float(__float_as_uint(y.x) & 0xFFFFE000u), hw = __uint_as_float(__float_as_uint(y.y) & 0xFFFFE000u);
                tmem_st4(t_hi + 4 * c, hx, hy, hz, hw);
                tmem_st4(t_lo + 4 * c, x.x - hx, x.y - hy, y.x - hz, y.y - hw);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(bar_a0 + 8 * b);
        }
    } else if (warp < (PROD_THREADS + EPI_THREADS) / 32) {
        // =============================================================== epilogue: TMEM -> bias, penalty -> ring
        const int etid = tid - PROD_THREADS, ehalf = etid >> 7;         // two threads per env row: column halves
        const double neg_vmax = -prm.v_max;
        for (int k = 0; k < n_my; ++k) {
            const int b = k & 1;
            const int64_t tile = tile_of(k);
            const int rows = rows_of(tile);
            mbar_wait(bar_mma0 + 8 * b, (uint32_t)((k >> 1) & 1));
            tc_fence_after();
            // TMEM lane = env row; a warp may only touch the lane quarter 32 * (warp % 4): warps 8-11 take
            // columns 0..16 of their quarter, warps 12-15 columns 17..32
            const int c0 = ehalf ? 17 : 0;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)b * ACC_COLS + (uint32_t)c0;
            uint32_t r0[16], r1 = 0u;
            tmem_ld16(taddr, r0);
            if (ehalf == 0) tmem_ld1(taddr + 16, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(bar_free0 + 8 * b);              // the accumulator may be overwritten (tile k + 2)
            if (etid == 0) bulk_wait_read();             // the previous tile's bulk stores have read `out`
            group_sync<2, EPI_THREADS>();                // ... and so have its store loops: `out` is free
            double pen = 0.0;
            if (row < rows) {
                float V[17];
                float vmn = INFINITY, vmx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    if (i < 16 || ehalf == 0) {
                        V[i] = __uint_as_float(i < 16 ? r0[i < 16 ? i : 0] : r1) + bias[c0 + i];
                        out[row * N_OUT + c0 + i] = V[i];
                        vmn = fminf(vmn, V[i]); vmx = fmaxf(vmx, V[i]);
                    }
                }
                // in-limit voltages (the common case) contribute exactly zero, and the fp32 tests are EXACT
                // (lo_f is the smallest float >= v_min, hi_f the largest <= v_max, V is a float): the fp64 terms
                // are only evaluated for a row that leaves the limits, selected per column by the fp32 test
                if (vmn < prm.lo_f || vmx > prm.hi_f) {
#pragma unroll
                    for (int i = 0; i < 17; ++i) {
                        if (i < 16 || ehalf == 0) {
                            // v_min - V or V - v_max as ONE fused multiply-add with selected operands (the same
                            // single rounding as the subtraction): a third of the fp64 work of computing both
                            const bool un = V[i] < prm.lo_f, ov = V[i] > prm.hi_f;
                            const double sg = un ? -1.0 : (ov ? 1.0 : 0.0);
                            const double cc = un ? prm.v_min : (ov ? neg_vmax : 0.0);
                            pen = pen + fma(sg, (double)V[i], cc);
                        }
                    }
                }
            }
            pen_part[ehalf * PRED_M + row] = pen;
            fence_proxy_async();                                 // `out` is read by the bulk stores (async proxy)
            group_sync<2, EPI_THREADS>();                        // the tile's Vhat rows and penalty halves are staged
            const int64_t e = tile * PRED_M + row;
            if (ehalf == 0 && row < rows) {
                const double pw = prm.w * (pen_part[row] + pen_part[PRED_M + row]);
                if (prm.penalty != nullptr) prm.penalty[e] = pw;
                if (prm.ring_pen != nullptr) {               // pos < cap and n <= cap: one conditional subtraction wraps
                    const int64_t rr = prm.ring_pos + e;
                    prm.ring_pen[rr >= prm.ring_cap ? rr - prm.ring_cap : rr] = (float)pw;
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
                // one bulk store (TMA engine) per destination: the staged tile is the contiguous image of its rows
                if (etid == 0) {
                    if (dense != nullptr) bulk_s2g(dense, out, (uint32_t)n_el * 4u);
                    if (prm.ring_vhat != nullptr) bulk_s2g(prm.ring_vhat + ring_base, out, (uint32_t)n_el * 4u);
                    bulk_commit();
                }
            } else {
                for (int i = etid; i < n_el; i += EPI_THREADS) {
                    const float v = out[i];
                    if (dense != nullptr) dense[i] = v;
                    if (prm.ring_vhat != nullptr) {
                        int64_t o = ring_base + i;
                        o = o >= ring_elems ? o - ring_elems : o;
                        prm.ring_vhat[o] = v;
                    }
                }
            }
        }
        if (etid == 0) bulk_wait_all();                  // shared memory must outlive the bulk stores
    }

    tc_fence_before();
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
