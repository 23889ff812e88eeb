// flex_common.cuh -- device-side structures and the warp-per-env DistFlow sweep.
//
// One warp owns one environment; lane k owns line k (the line feeding the bus whose DFS
// pre-order rank is k+1).  Pre-order numbering makes every subtree a contiguous lane range
// [k, end_k], so
//   * the backward sweep  (utils/pf.py:65-83: P_ij = p_j + sum_children(P_jk + R_jk l_jk))
//     is an inclusive warp scan followed by one indexed shuffle:  P_k = p_k + (S[end_k] - S[k]);
//   * the forward sweep   (utils/pf.py:90-94: v_j = v_i - 2(R P + X Q) - |z|^2 l)
//     is a root-path sum by pointer jumping over the statically known 2^r-th ancestors;
//   * the current update  (utils/pf.py:85-88: l v_j = P^2 + Q^2) is one IEEE fp64 divide/lane.
//
// All arithmetic is fp64 with a FIXED operation order (compiled with -fmad=false; fused
// multiply-adds are written explicitly as fma()) so that oracle/c/flex_oracle.c can mirror
// it bit for bit -- that is what makes the voltage-violation masks bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flexgpu.h"

#define FULL 0xffffffffu
#define FP_NL 32            // lanes == max lines
#define ANC_NONE 32u

// Scalar configuration, passed by value as a kernel parameter (uniform constant-bank reads).
struct DevCfg {
    int32_t nb, nl, na, history, episode_limit, raw_actions, pf_max_iter, obs_w;
    double pf_tol, v_min, v_max, e_min, e_max, p_ch_max, p_dis_max, eta_ch, eta_dis, inv_eta_dis;
    double mpr, kappa, pv_cost, ess_cost, discomfort_coeff, voltage_coeff, delta_t, fail_penalty;
    double e_next_lb, slack_pen;
    int32_t slack_viol, pad_;
};

// Per-lane topology tables.  Lives in global memory (handle-owned); every CTA stages it into
// shared memory once and each lane then keeps its own hot entries (R, X, |z|^2, end, anc) in
// registers for the whole grid-stride loop.
struct DevTopo {
    double R[FP_NL], X[FP_NL], Z2[FP_NL], imax2[FP_NL];
    int32_t end[FP_NL];        // last lane of this lane's subtree
    uint32_t anc[FP_NL];       // 5 x 6-bit: lane of the 2^r-th ancestor line, ANC_NONE if none
    int32_t col[FP_NL];        // dataset column (= bus position - 1) of this lane's bus
    int32_t agent[FP_NL];      // agent whose building sits on this lane's bus, else -1
    int32_t agent_lane[8];     // lane of agent i's bus
    int32_t agent_col[8];      // dataset column of agent i's bus
};

struct LaneTopo {
    double R, X, Z2;
    int32_t end;
    uint32_t anc;
};

__device__ __forceinline__ LaneTopo load_lane_topo(const DevTopo& t, int lane) {
    LaneTopo l;
    l.R = t.R[lane]; l.X = t.X[lane]; l.Z2 = t.Z2[lane];
    l.end = t.end[lane]; l.anc = t.anc[lane];
    return l;
}

__device__ __forceinline__ void stage_topo(DevTopo* s_topo, const DevTopo* __restrict__ g_topo) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g_topo);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_topo);
    for (int i = threadIdx.x; i < (int)(sizeof(DevTopo) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ double clipd(double x, double lo, double hi) {
    // numpy.clip == minimum(maximum(x, lo), hi)
    double t = (x > lo) ? x : lo;
    return (t < hi) ? t : hi;
}

// Inclusive Hillis-Steele scan over the warp (lane order).
__device__ __forceinline__ double warp_scan_incl(double s, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double y = __shfl_up_sync(FULL, s, d);
        if (lane >= d) s = s + y;
    }
    return s;
}

// Butterfly sum; every lane ends with the same value.
__device__ __forceinline__ double warp_sum_xor(double s) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s = s + __shfl_xor_sync(FULL, s, d);
    return s;
}

struct SweepOut {
    double P, Q, ell, v;   // receiving-end flows, squared current, squared voltage of this lane
    int iters;
    bool ok;
};

// Backward/forward sweep from a flat start (l = 0, v = 1).  p, q: net consumption at this
// lane's bus.  Lanes >= nl must pass p = q = 0 and a LaneTopo with R = X = Z2 = 0, end = lane,
// anc = none (the host builds the tables that way).
__device__ __forceinline__ SweepOut distflow_sweep(const LaneTopo& t, double p, double q, int lane,
                                                  double tol, int max_iter, bool inject_fail) {
    SweepOut o;
    double ell = 0.0, v_old = 1.0, P = 0.0, Q = 0.0, v = 1.0;
    bool conv = false, bad = false;
    int it = 0;
    while (it < max_iter) {
        ++it;
        // ---- backward: subtree sums of x = p + R*l (sending-end contributions)
        double SP = warp_scan_incl(fma(t.R, ell, p), lane);
        double SQ = warp_scan_incl(fma(t.X, ell, q), lane);
        double SPe = __shfl_sync(FULL, SP, t.end);
        double SQe = __shfl_sync(FULL, SQ, t.end);
        P = p + (SPe - SP);
        Q = q + (SQe - SQ);
        // ---- forward: v = 1 - sum over the root path of the per-line drops
        double d = t.R * P;
        d = fma(t.X, Q, d);
        d = fma(t.Z2, ell, d + d);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            uint32_t a = (t.anc >> (6 * r)) & 63u;
            double y = __shfl_sync(FULL, d, a & 31u);
            if (a < ANC_NONE) d = d + y;
        }
        v = 1.0 - d;
        // ---- current: l = (P^2 + Q^2) / v
        double s = P * P;
        s = fma(Q, Q, s);
        ell = s / v;
        bad = __any_sync(FULL, !(v > 0.0));
        conv = __all_sync(FULL, fabs(v - v_old) <= tol);
        v_old = v;
        if (bad || conv) break;
    }
    o.P = P; o.Q = Q; o.ell = ell; o.v = v; o.iters = it;
    o.ok = conv && !bad && !inject_fail;
    return o;
}

// ------------------------------------------------------------------ Philox4x32-10
struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
        k0 += W0; k1 += W1;
    }
    return c;
}

// 53-bit uniform in [0,1) from two 32-bit words (same construction as numpy's random_sample).
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
