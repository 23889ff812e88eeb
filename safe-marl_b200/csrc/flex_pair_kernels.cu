// flex_pair_kernels.cu -- pair-per-env kernels (sm_100a) for the IEEE 33-bus shape: the
// throughput path of the thread variant.
//
//   k_env_p<STEP>   replaces FlexibilityProvisionEnv.step           (flexibility_provision_env.py:241-356)
//   k_env_p<RESET>  replaces reset()/manual_reset()                  (:74-155, :157-239)
//   k_power_flow_p  replaces power_flow_solver_simplified, batched   (utils/pf.py:115-192)
//
// Mapping: TWO adjacent lanes own one environment; a warp owns a tile of 16 consecutive envs.
//   * role 0 (even lane) walks the main feeder (lines 0..16 in DFS pre-order), role 1 (odd lane)
//     the three laterals (C: 28..31 off bus 2, B: 25..27 off bus 3, A: 17..24 off bus 6), both in
//     17 lock-step "steps" of the same instruction stream.  What differs between the roles is
//     DATA -- per-lane base pointers into the line-constant table and the S tile, a handful of
//     selects at the three chain heads -- so there is no divergence.  Role 1 pads its 15 lines
//     with two zero-impedance steps (first and last), which places every lateral head one step
//     after the main-feeder line whose bus it leaves: the parent voltage arrives by one shuffle.
//   * The arithmetic per line and its order are exactly those of the one-thread-per-env kernel
//     (flex_thread_kernels.cu: one-pass S/W form of the DistFlow fixed point), so the same CPU
//     mirror (oracle/c/flex_oracle.c, variant 0) reproduces every bit.
//   * Why: an env's state (S: 64 doubles, l: 32 doubles) caps an SM at ~256 resident envs.  With
//     one thread per env that is 8 warps of 255 registers whose 32-line dependent chains leave
//     the fp64 pipe idle 3/4 of the time (a lone warp needs 25.6 us per tile, two per scheduler
//     37.9 us: latency-bound).  Two lanes per env give the same 256 envs as 16 warps of <= 128
//     registers with chains half as long: twice the warps to hide each latency behind.
#include "flex_kernels.cuh"
#include "flex_env_math.cuh"

#include <cmath>
#include <type_traits>
#include <utility>

namespace {

constexpr int NS = 17;        // steps per pass (lines per role, padded)
constexpr int NPOS = 2 * NS;  // line positions per env: role 0 at 0..16, role 1 at 17..33
constexpr int PROW = 2 * NPOS;  // doubles per env row of the S tile: 34 (S_P, S_Q) pairs.  34 = 2 (mod 8) 16-byte
                                // units per env and 17 = 1 (mod 8) between the roles: the eight lanes of a
                                // quarter warp hit eight different 16-byte bank groups -> conflict-free LDS.128
constexpr int VROW = 33;      // row stride of the V tile (doubles)
constexpr int TILE = 16;      // envs per warp
constexpr int CTA_WARPS = 2;

// role 1's DFS lane at each step (role 0's lane at step s is s); -1 = padding step
struct PairMap {
    static constexpr int8_t K1[NS] = {-1, 28, 29, 30, 31, 25, 26, 27, 17, 18, 19, 20, 21, 22, 23, 24, -1};
};
template <int S> constexpr int COL0 = Ieee33Tree::COL[S];
template <int S> constexpr int COL1 = PairMap::K1[S] >= 0 ? Ieee33Tree::COL[PairMap::K1[S]] : 0;
template <int S> constexpr bool PAD1 = PairMap::K1[S] < 0;

__host__ __device__ constexpr int warp_smem_doubles_p() { return TILE * PROW + TILE * VROW; }
__host__ __device__ constexpr int cta_smem_doubles_p() { return CTA_WARPS * warp_smem_doubles_p() + 4 * NPOS; }

__device__ __forceinline__ uint64_t pack2(int32_t lo, int32_t hi) {
    return (uint64_t)(uint32_t)lo | ((uint64_t)(uint32_t)hi << 32);
}
__device__ __forceinline__ uint64_t d2u(double x) { return (uint64_t)__double_as_longlong(x); }
__device__ __forceinline__ double u2d(uint64_t x) { return __longlong_as_double((long long)x); }
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// exchange with the partner lane; m = the lanes executing this code (both lanes of a pair always do)
__device__ __forceinline__ double shfl_pair(unsigned m, double x) { return __shfl_xor_sync(m, x, 1); }

// Line constants of this lane's step S: volatile broadcast loads exactly where they are used (see
// flex_thread_kernels.cu: left to itself the compiler hoists the 96 loop-invariant doubles out of
// the iteration loop and spills them).  ltp is the lane's role-offset pointer into the table
// [34 positions] x (R, X, |z|^2/2, Imax^2).
template <int S>
__device__ __forceinline__ void line_rx(const double* ltp, double& R, double& X) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(ltp + 4 * S);
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(R), "=d"(X) : "r"(a));
}
template <int S>
__device__ __forceinline__ double line_z2h(const double* ltp) {
    return *(reinterpret_cast<const volatile double*>(ltp) + 4 * S + 2);
}
template <int S>
__device__ __forceinline__ double line_imax2(const double* ltp) {
    return *(reinterpret_cast<const volatile double*>(ltp) + 4 * S + 3);
}

template <int V> struct IntC { static constexpr int value = V; };
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(IntC<Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// Per-lane state of a solve.
//   ell       squared currents of this lane's 17 steps (registers)
//   U, Uq     role 0: [0] = loss total of the whole tree (what the main feeder's first line starts
//                       from); role 1: loss totals of its chains C, B, A (heads at steps 1, 5, 8)
//   X, Xq     role 0: the partner's chain totals C, B, A, subtracted where each lateral leaves the
//                       main feeder (steps 1, 2, 5); role 1: zero
struct PState {
    double ell[NS];
    double U[3], Uq[3], X[3], Xq[3];
    int iters; bool conv, bad;
};

// Carried state of a pass: W (P and Q parts) and the squared voltage of the previous line; the
// main feeder's squared voltages at the three branch buses (steps 0, 1, 4).
struct PCarry { double w, wq, v, vs0, vs1, vs4, acc, accq, un0, un0q, un1, un1q; };

// Flows and squared voltage of this lane's line at step S (utils/pf.py:65-83, :90-94); same
// arithmetic as t_line in flex_thread_kernels.cu.
template <int S>
__device__ __forceinline__ void p_line(unsigned m, bool role, const double2* row2, const double* ltp, double eo, const PState& st,
                                       PCarry& cy, double& P, double& Q, double& v, double& R, double& X) {
    double w, wq, vp;
    if constexpr (S == 0) {                      // role 0: head of the main feeder; role 1: padding
        w = role ? 0.0 : st.U[0]; wq = role ? 0.0 : st.Uq[0]; vp = 1.0;
    } else if constexpr (S == 1 || S == 5 || S == 8) {
        // role 1: head of lateral C / B / A -- starts from its chain total and the voltage of the
        // branch bus (one shuffle from the partner); role 0: next line of the main feeder
        constexpr int HI = (S == 1) ? 0 : (S == 5) ? 1 : 2;
        const double vb = shfl_pair(m, S == 1 ? cy.vs0 : S == 5 ? cy.vs1 : cy.vs4);
        double wc = cy.w, wqc = cy.wq;
        if constexpr (S == 1) { wc = wc - st.X[0]; wqc = wqc - st.Xq[0]; }     // lateral C leaves bus 2
        if constexpr (S == 5) { wc = wc - st.X[2]; wqc = wqc - st.Xq[2]; }     // lateral A leaves bus 6
        w = role ? st.U[HI] : wc; wq = role ? st.Uq[HI] : wqc; vp = role ? vb : cy.v;
    } else {
        w = cy.w; wq = cy.wq; vp = cy.v;
        if constexpr (S == 2) { w = w - st.X[1]; wq = wq - st.Xq[1]; }         // lateral B leaves bus 3
    }
    line_rx<S>(ltp, R, X);
    w = fma(-R, eo, w); wq = fma(-X, eo, wq);
    cy.w = w; cy.wq = wq;
    const double2 s = row2[S];
    P = s.x + w; Q = s.y + wq;
    double g = R * P;
    g = fma(X, Q, g);
    g = fma(line_z2h<S>(ltp), eo, g);
    v = fma(-2.0, g, vp);
    cy.v = v;
    if constexpr (S == 0) cy.vs0 = v;
    if constexpr (S == 1) cy.vs1 = v;
    if constexpr (S == 4) cy.vs4 = v;
}

// Steps [S0, S0 + B) of one pass, emitted stage by stage (see t_pass_batch in
// flex_thread_kernels.cu).  Loss totals: one running accumulator per lane; role 1 closes a chain
// where the next one starts (steps 5 and 8).
template <int S0, int B>
__device__ __forceinline__ void p_pass_batch(unsigned m, bool role, const double2* row2, const double* ltp, PState& st, PCarry& cy,
                                             int32_t& dmax, bool& bad) {
    double v[B], s[B], R[B], X[B];
    static_for<B>([&](auto j) {
        constexpr int J = decltype(j)::value, S = S0 + J;
        if constexpr (S < NS) {
            double P, Q;
            p_line<S>(m, role, row2, ltp, st.ell[S], st, cy, P, Q, v[J], R[J], X[J]);
            const double t = P * P;
            s[J] = fma(Q, Q, t);
            bad = bad || sqv_bad(v[J]);
        }
    });
    // division-free relaxed update of the current row (same arithmetic as t_pass_batch)
    double e[B], d[B], rt[B];
    static_for<B>([&](auto j) {
        constexpr int J = decltype(j)::value, S = S0 + J;
        if constexpr (S < NS) { e[J] = fma(-v[J], st.ell[S], s[J]); d[J] = 1.0 - v[J]; rt[J] = 2.0 - v[J]; }
    });
    static_for<B>([&](auto j) {
        constexpr int J = decltype(j)::value, S = S0 + J;
        if constexpr (S < NS) rt[J] = fma(d[J], rt[J], 1.0);
    });
    static_for<B>([&](auto j) {
        constexpr int J = decltype(j)::value, S = S0 + J;
        if constexpr (S < NS) {
            const double en = fma(e[J], rt[J], st.ell[S]);
            const int32_t dh = __double2hiint(e[J]) & 0x7FFFFFFF;
            dmax = dh > dmax ? dh : dmax;
            st.ell[S] = en;
            if constexpr (S == 5) { cy.un0 = cy.acc; cy.un0q = cy.accq; cy.acc = role ? 0.0 : cy.acc; cy.accq = role ? 0.0 : cy.accq; }
            if constexpr (S == 8) { cy.un1 = cy.acc; cy.un1q = cy.accq; cy.acc = role ? 0.0 : cy.acc; cy.accq = role ? 0.0 : cy.accq; }
            cy.acc = fma(R[J], en, cy.acc);
            cy.accq = fma(X[J], en, cy.accq);
        }
    });
}

#ifndef FP_PAIR_BATCH
#define FP_PAIR_BATCH 6
#endif
constexpr int PB = FP_PAIR_BATCH;

template <int S0>
__device__ __forceinline__ void p_pass_from(unsigned m, bool role, const double2* row2, const double* ltp, PState& st, PCarry& cy,
                                            int32_t& dmax, bool& bad) {
    p_pass_batch<S0, PB>(m, role, row2, ltp, st, cy, dmax, bad);
    if constexpr (S0 + PB < NS) p_pass_from<S0 + PB>(m, role, row2, ltp, st, cy, dmax, bad);
}

__device__ __forceinline__ void carry_init(PCarry& cy) {
    cy.w = cy.wq = 0.0; cy.v = cy.vs0 = cy.vs1 = cy.vs4 = 1.0;
    cy.acc = cy.accq = cy.un0 = cy.un0q = cy.un1 = cy.un1q = 0.0;
}

// New loss totals after a pass.  Role 1 keeps its chain totals (C, B, A); role 0 receives them
// from the partner and forms the tree total in the mirror's order ((main + A) + B) + C.
__device__ __forceinline__ void p_totals(unsigned m, bool role, PState& st, const PCarry& cy) {
    const double c = cy.un0, cq = cy.un0q, b = cy.un1, bq = cy.un1q, a = cy.acc, aq = cy.accq;   // role 1's chains
    const double tc = shfl_pair(m, c), tcq = shfl_pair(m, cq), tb = shfl_pair(m, b), tbq = shfl_pair(m, bq),
                 ta = shfl_pair(m, a), taq = shfl_pair(m, aq);
    if (role) {
        st.U[0] = c; st.Uq[0] = cq; st.U[1] = b; st.Uq[1] = bq; st.U[2] = a; st.Uq[2] = aq;
    } else {
        st.X[0] = tc; st.Xq[0] = tcq; st.X[1] = tb; st.Xq[1] = tbq; st.X[2] = ta; st.Xq[2] = taq;
        st.U[0] = ((cy.acc + ta) + tb) + tc;
        st.Uq[0] = ((cy.accq + taq) + tbq) + tcq;
    }
}

// S_k: subtree sums of the net injections, children before parents, in place in the tile row.
// The sums of the lateral heads (role 1: steps 8, 5, 1) reach the main feeder's branch buses
// (role 0: steps 4, 1, 0) by shuffle, and are added before the next line's carry -- the mirror's
// order S_k = (p_k + [laterals]) + S_{k+1}.
template <int S>
__device__ __forceinline__ void p_setup_from(unsigned m, bool role, double2* row2, double& c, double& cq, double& ha, double& haq,
                                             double& hb, double& hbq, double& hc, double& hcq) {
    const double2 pq = row2[S];
    double tp = pq.x, tq = pq.y;
    if constexpr (S == 4 || S == 1 || S == 0) {
        const double sl = shfl_pair(m, S == 4 ? ha : S == 1 ? hb : hc), slq = shfl_pair(m, S == 4 ? haq : S == 1 ? hbq : hcq);
        if (!role) { tp = tp + sl; tq = tq + slq; }
    }
    if constexpr (S < NS - 1) {
        // role 0: line S+1 is always the child of line S; role 1: not at the leaves of A (15), B (7),
        // C (4) and not at the padding step 0
        constexpr bool leaf1 = (S == 15 || S == 7 || S == 4 || S == 0);
        if constexpr (leaf1) { if (!role) { tp = tp + c; tq = tq + cq; } }
        else { tp = tp + c; tq = tq + cq; }
    }
    row2[S] = make_double2(tp, tq);
    c = tp; cq = tq;
    if constexpr (S == 8) { ha = tp; haq = tq; }
    if constexpr (S == 5) { hb = tp; hbq = tq; }
    if constexpr (S == 1) { hc = tp; hcq = tq; }
    if constexpr (S > 0) p_setup_from<S - 1>(m, role, row2, c, cq, ha, haq, hb, hbq, hc, hcq);
}

// First half of a solve: S setup + the fixed-point passes (see t_iterate).  The two lanes of a pair
// always take the same branches (valid / active are per env), so the pair shuffles inside divergent
// code name exactly the lanes that execute them.
__device__ __forceinline__ void p_iterate(bool role, double2* row2, const double* ltp, PState& st, double tol, int max_iter,
                                          bool valid) {
    bool active = valid;
    st.conv = false; st.bad = false; st.iters = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { st.U[i] = st.Uq[i] = st.X[i] = st.Xq[i] = 0.0; }
#pragma unroll
    for (int k = 0; k < NS; ++k) st.ell[k] = 0.0;
    const unsigned mv = __ballot_sync(FULL, valid);
    if (valid) {
        double c = 0.0, cq = 0.0, ha = 0.0, haq = 0.0, hb = 0.0, hbq = 0.0, hc = 0.0, hcq = 0.0;
        p_setup_from<NS - 1>(mv, role, row2, c, cq, ha, haq, hb, hbq, hc, hcq);
    }
    const int32_t tol_hi = __double2hiint(tol);
    for (int it = 1; it <= max_iter; ++it) {
        const unsigned ma = __ballot_sync(FULL, active);
        if (ma == 0u) break;
        if (active) {
            int32_t dmax = 0;
            PCarry cy;
            carry_init(cy);
            p_pass_from<0>(ma, role, row2, ltp, st, cy, dmax, st.bad);
            p_totals(ma, role, st, cy);
            dmax = max(dmax, __shfl_xor_sync(ma, dmax, 1));
            st.bad = (__shfl_xor_sync(ma, (int)st.bad, 1) != 0) || st.bad;
            st.iters = it;
            const bool cv = dmax < tol_hi;
            if (st.bad || cv) { active = false; st.conv = cv; }
        }
    }
}

struct PSolve { int iters; bool ok; uint32_t vm, lm; };

// Final pass (see t_final_batch): (P, Q) replace (S_P, S_Q) in the tile row, V = sqrt(v) goes to
// the env's voltage row (bus order), violation masks on the fly.
__device__ __forceinline__ double sqrt_normal(double v) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
    double e = y * y;
    e = fma(v, -e, 1.0);
    const double p = fma(e, 0.375, 0.5);
    e = y * e;
    y = fma(p, e, y);
    const double g = v * y;
    const double h = 0.5 * y;
    const double d = fma(-g, g, v);
    return fma(d, h, g);
}

template <int S0, int B>
__device__ __forceinline__ void p_final_batch(unsigned m, bool role, const DevCfg* c, bool any_imax, double2* row2,
                                              const double* ltp, const PState& st, PCarry& cy, double* vrow, bool& bad,
                                              uint32_t& vm, uint32_t& lm) {
    double v[B], V[B];
    static_for<B>([&](auto j) {
        constexpr int J = decltype(j)::value, S = S0 + J;
        if constexpr (S < NS) {
            double P, Q, R, X;
            p_line<S>(m, role, row2, ltp, st.ell[S], st, cy, P, Q, v[J], R, X);
            row2[S] = make_double2(P, Q);
            bad = bad || sqv_bad(v[J]);
        }
    });
    static_for<B>([&](auto j) {
        constexpr int J = decltype(j)::value, S = S0 + J;
        if constexpr (S < NS) V[J] = sqrt_normal(v[J]);                       // pf.py:108
    });
    static_for<B>([&](auto j) {
        constexpr int J = decltype(j)::value, S = S0 + J;
        if constexpr (S < NS) {
            const bool real = !(role && PAD1<S>);
            const int col = role ? COL1<S> : COL0<S>;
            if (real) {
                vrow[col + 1] = V[J];
                if (c != nullptr) {
                    if ((V[J] > c->v_max) || (V[J] < c->v_min)) vm |= 1u << col;
                    if (any_imax && (st.ell[S] > line_imax2<S>(ltp))) lm |= 1u << col;   // utils/opf.py:124-126
                }
            }
        }
    });
}
template <int S0>
__device__ __forceinline__ void p_final_from(unsigned m, bool role, const DevCfg* c, bool any_imax, double2* row2,
                                             const double* ltp, const PState& st, PCarry& cy, double* vrow, bool& bad,
                                             uint32_t& vm, uint32_t& lm) {
    p_final_batch<S0, PB>(m, role, c, any_imax, row2, ltp, st, cy, vrow, bad, vm, lm);
    if constexpr (S0 + PB < NS) p_final_from<S0 + PB>(m, role, c, any_imax, row2, ltp, st, cy, vrow, bad, vm, lm);
}

// Second half of a solve.  Both lanes of the pair return the same PSolve (masks and flags merged).
__device__ __forceinline__ PSolve p_finish(bool role, const DevCfg* c, bool any_imax, double2* row2, const double* ltp,
                                           double* vrow, PState& st, bool valid) {
    PSolve s; s.vm = 0u; s.lm = 0u;
    bool bad = st.bad;
    const unsigned mv = __ballot_sync(FULL, valid);
    if (valid) {
        PCarry cy;
        carry_init(cy);
        if (!role) vrow[0] = 1.0;                     // slack: sqrt(Vsqr = 1), pf.py:51-53
        p_final_from<0>(mv, role, c, any_imax, row2, ltp, st, cy, vrow, bad, s.vm, s.lm);
        s.vm |= __shfl_xor_sync(mv, s.vm, 1);
        s.lm |= __shfl_xor_sync(mv, s.lm, 1);
        bad = (__shfl_xor_sync(mv, (int)bad, 1) != 0) || bad;
    }
    s.iters = st.iters; s.ok = st.conv && !bad;
    return s;
}

// Cooperative store of the V tile rows [16][VROW] (first `w` columns) to g[(e0 + j) * w + c] for the
// envs whose bit is set in emask: one coalesced pass over the contiguous chunk.
__device__ __forceinline__ void store_rows16(const double* tile, double* __restrict__ g, int64_t e0, int w, uint32_t emask,
                                             int lane) {
    double* dst = g + e0 * w;
    for (int i = lane; i < TILE * w; i += 32) {
        const int j = i / w, c = i - j * w;
        if ((emask >> j) & 1u) dst[i] = tile[j * VROW + c];
    }
}

// Voltage penalty (:685) of the buses flagged in vm, summed in DFS lane order; rebuild = true
// re-derives the mask from rolled-back voltages (both off the common path).
__device__ __forceinline__ double voltage_penalty(const PairTopo& T, const DevCfg& c, const double* vrow, int nl,
                                                  bool rebuild, uint32_t& vm) {
    double vpen = 0.0;
    if (rebuild) vm = 0u;
    else if (vm == 0u) return vpen;
    for (int k = 0; k < nl; ++k) {
        const int col = T.col_of_lane[k];
        const double V = vrow[col + 1];
        const double over = V - c.v_max, under = c.v_min - V;
        if ((over > 0.0) || (under > 0.0)) {
            vm |= 1u << col;
            vpen = vpen + c.voltage_coeff * ((over > under) ? over : under);
        }
    }
    return vpen;
}

// Tile-wide staging of the line-constant table (once per CTA).
__device__ __forceinline__ void stage_line_table(double* lt, const PairTopo& T) {
    for (int i = threadIdx.x; i < NPOS; i += blockDim.x) {
        double2* l2 = reinterpret_cast<double2*>(lt) + 2 * i;
        l2[0] = make_double2(T.R[i], T.X[i]);
        l2[1] = make_double2(T.Z2h[i], T.imax2[i]);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------- env kernel
template <int MODE>
__global__ void __launch_bounds__(32 * CTA_WARPS, 8) k_env_p(const EnvParamsP prm) {
    extern __shared__ double smem[];
    const EnvParams& q = prm.e;
    const PairTopo& T = prm.t;
    const DevCfg& c = q.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool role = (lane & 1) != 0;
    const int el = lane >> 1;                                             // env of the tile this lane works for
    const int nl = c.nl, na = c.na, nb = c.nb;
    double* lt = smem + CTA_WARPS * warp_smem_doubles_p();
    stage_line_table(lt, T);
    double* st_tile = smem + warp * warp_smem_doubles_p();
    double* vt_tile = st_tile + TILE * PROW;
    double2* row2e = reinterpret_cast<double2*>(st_tile + el * PROW);     // this env's 34 pairs
    double2* row2 = row2e + (role ? NS : 0);                              // this lane's 17 pairs
    const double* ltp = lt + (role ? 4 * NS : 0);
    double* vrow = vt_tile + el * VROW;
    const int my_pos = (lane < nl) ? T.pos_of_col[lane] : 0;              // S-tile pair of dataset column `lane`
    const bool any_imax = T.any_imax != 0;
    double stat_acc = 0.0;

    const int64_t n_tiles = (q.tile_end > q.tile_begin) ? 2 * q.tile_end : ((q.n + TILE - 1) / TILE);
    const int64_t w_stride = (int64_t)gridDim.x * CTA_WARPS;
    for (int64_t tile = 2 * q.tile_begin + (int64_t)blockIdx.x * CTA_WARPS + warp; tile < n_tiles; tile += w_stride) {
        const int64_t e0 = tile * TILE, e = e0 + el;
        const bool valid = (e < q.n) && (q.mask == nullptr || q.mask[e] != 0);
        const bool owner = valid && !role;                                // the lane that runs the env's scalar logic
        const uint32_t vmask_w = __ballot_sync(FULL, valid);
        if (vmask_w == 0u) continue;
        uint64_t* rec = q.rec + (valid ? e : e0) * FP_REC_STRIDE;

        // ------------------------------------------------------------ L2 prefetch of this warp's next tile
        uint64_t time_next = 0ull;
        bool have_next = false;
        if (MODE == MODE_STEP && !role) {
            const int64_t en = (tile + w_stride) * TILE + el;
            if (tile + w_stride < n_tiles && en < q.n) {
                const uint64_t* rn = q.rec + en * FP_REC_STRIDE;
                time_next = __ldg(rn + FP_REC_TIME);
                prefetch_l2(reinterpret_cast<const char*>(q.actions) + en * (int64_t)na * (q.act_f64 ? 32 : 16));
                have_next = true;
            }
        }

        // ------------------------------------------------------------ per-env record + inputs (owner lane)
        int32_t start = 0, steps = 1, hist_n = 0, episode = 0;
        double a[FP_MAX_AGENTS][4], e_clip[FP_MAX_AGENTS], e_init[FP_MAX_AGENTS];
        double cum = 0.0;
#pragma unroll
        for (int i = 0; i < FP_MAX_AGENTS; ++i) { a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.0; e_clip[i] = e_init[i] = 0.0; }
        if (owner) {
            if (MODE == MODE_STEP) {
                const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(rec);
                uint64_t r[FP_REC_STRIDE];
#pragma unroll
                for (int i = 0; i < 13; i += 2) { const ulonglong2 t2 = r2[i >> 1]; r[i] = t2.x; r[i + 1] = t2.y; }
#pragma unroll
                for (int i = 0; i < FP_MAX_AGENTS; ++i)
                    if (i < na) { e_init[i] = u2d(r[FP_REC_E_INIT + i]); e_clip[i] = u2d(r[FP_REC_E_CUR + i]); }
                cum = u2d(r[FP_REC_CUM]);
                start = (int32_t)(uint32_t)r[FP_REC_TIME]; steps = (int32_t)(r[FP_REC_TIME] >> 32);
                hist_n = (int32_t)(uint32_t)r[FP_REC_HIST]; episode = (int32_t)(r[FP_REC_HIST] >> 32);
                if (q.act_f64) {
                    const double2* a2 = reinterpret_cast<const double2*>(q.actions) + e * (2 * na);
#pragma unroll
                    for (int i = 0; i < FP_MAX_AGENTS; ++i)
                        if (i < na) { const double2 lo = a2[2 * i], hi = a2[2 * i + 1]; a[i][0] = lo.x; a[i][1] = lo.y; a[i][2] = hi.x; a[i][3] = hi.y; }
                } else {                                             // fp32 actions widen exactly (quirk Q6)
                    const float4* a4 = reinterpret_cast<const float4*>(q.actions) + e * na;
#pragma unroll
                    for (int i = 0; i < FP_MAX_AGENTS; ++i)
                        if (i < na) { const float4 t4 = a4[i]; a[i][0] = (double)t4.x; a[i][1] = (double)t4.y; a[i][2] = (double)t4.z; a[i][3] = (double)t4.w; }
                }
            } else {
                episode = (int32_t)(rec[FP_REC_HIST] >> 32);
                if (q.random) {
                    // counter = (global env id lo, hi, episode, block): same blocks as the other kernels
                    const uint64_t gid = (uint64_t)(q.env_offset + e);
                    const uint32_t k0 = (uint32_t)q.seed, k1 = (uint32_t)(q.seed >> 32);
                    U4 ctr; ctr.x = (uint32_t)gid; ctr.y = (uint32_t)(gid >> 32); ctr.z = (uint32_t)episode;
                    ctr.w = 15u;
                    U4 rr = philox4x32_10(ctr, k0, k1);
                    start = (int32_t)(u53(rr.x, rr.y) * (double)q.start_range);
                    const double lo = 0.9 * (c.e_max / 2), hi = 1.1 * (c.e_max / 2);          // :100
#pragma unroll
                    for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                        if (i < na) {
                            ctr.w = 10u + i; rr = philox4x32_10(ctr, k0, k1);
                            e_init[i] = lo + (hi - lo) * u53(rr.x, rr.y);
                            ctr.w = 2u * i; rr = philox4x32_10(ctr, k0, k1);
                            a[i][0] = u53(rr.x, rr.y); a[i][1] = u53(rr.z, rr.w);
                            ctr.w = 2u * i + 1u; rr = philox4x32_10(ctr, k0, k1);
                            a[i][2] = u53(rr.x, rr.y); a[i][3] = u53(rr.z, rr.w);
                        }
                    }
                } else {
                    start = q.start[e];
#pragma unroll
                    for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                        if (i < na) {
                            e_init[i] = q.e0[e * na + i];
                            const double* ap = q.a0 + (e * na + i) * 4;
                            a[i][0] = ap[0]; a[i][1] = ap[1]; a[i][2] = ap[2]; a[i][3] = ap[3];
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < FP_MAX_AGENTS; ++i) e_clip[i] = e_init[i];     // reset clips against E0 (:130)
            }
        }
        // Quirk Q1: the row in force is max(steps-1, 1); the row loaded after the solve is `steps`.
        const int32_t row = start + ((MODE == MODE_STEP && steps > 1) ? min(steps - 1, c.episode_limit + c.history) : 1);

        // ------------------------------------------------------------ gather the profile rows
        // one coalesced 256-byte row per instruction (a different dataset row per env), written
        // asynchronously into the (p, q) pairs of the S tile; padding pairs are zeroed
#pragma unroll
        for (int j = 0; j < TILE; ++j) {
            const int32_t rj = __shfl_sync(FULL, row, 2 * j);
            if (((vmask_w >> (2 * j)) & 1u) && lane < nl) {
                cp_async8(st_tile + j * PROW + 2 * my_pos, q.P + (int64_t)rj * nl + lane);
                cp_async8(st_tile + j * PROW + 2 * my_pos + 1, q.Q + (int64_t)rj * nl + lane);
            }
        }
        if (role) { row2[0] = make_double2(0.0, 0.0); row2[NS - 1] = make_double2(0.0, 0.0); }
        double pv[FP_MAX_AGENTS], price = 0.0;
#pragma unroll
        for (int i = 0; i < FP_MAX_AGENTS; ++i) pv[i] = 0.0;
        if (owner) {
            const double2* pv2 = reinterpret_cast<const double2*>(q.PVP + (int64_t)row * FP_PVP_STRIDE);
            const double2 p01 = __ldg(pv2), p23 = __ldg(pv2 + 1), p45 = __ldg(pv2 + 2);
            pv[0] = p01.x; pv[1] = p01.y; pv[2] = p23.x; pv[3] = p23.y; pv[4] = p45.x; price = p45.y;
        }
        cp_async_wait_all();
        __syncwarp();
        if (MODE == MODE_STEP && have_next) {
            const int32_t sn = (int32_t)(uint32_t)time_next, tn = (int32_t)(time_next >> 32);
            const int64_t rown = (int64_t)sn + ((tn > 1) ? min(tn - 1, c.episode_limit + c.history) : 1);
            const char* pp = reinterpret_cast<const char*>(q.P + rown * nl);
            const char* qp = reinterpret_cast<const char*>(q.Q + rown * nl);
            prefetch_l2(pp); prefetch_l2(qp);
            if (nl > 16) { prefetch_l2(pp + 128); prefetch_l2(qp + 128); }
            prefetch_l2(q.PVP + rown * FP_PVP_STRIDE);
        }

        // ------------------------------------------------------------ actions -> setpoints -> injections
        double s_pred[FP_MAX_AGENTS], s_ch[FP_MAX_AGENTS], s_dis[FP_MAX_AGENTS], s_qpv[FP_MAX_AGENTS], e_next[FP_MAX_AGENTS];
        double rev = 0.0, der = 0.0, ess = 0.0, disc = 0.0;
        bool e_bad = false;
#pragma unroll
        for (int i = 0; i < FP_MAX_AGENTS; ++i) { s_pred[i] = s_ch[i] = s_dis[i] = s_qpv[i] = e_next[i] = 0.0; }
        if (owner) {
            const bool scale = (MODE == MODE_RESET) || !c.raw_actions;
#pragma unroll
            for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                if (i < na) {
                    const int ap = T.agent_pos[i];
                    const double2 pq = row2e[ap];
                    const double pload = pq.x;
                    const Setpoint sp = apply_actions(c, scale, a[i][0], a[i][1], a[i][2], a[i][3], pload, pv[i], e_clip[i]);
                    // net consumption at the building's bus, balance rows utils/pf.py:65-83
                    row2e[ap] = make_double2((((pload - sp.pred) - pv[i]) + sp.ch) - sp.dis, pq.y - sp.qpv);
                    s_pred[i] = sp.pred; s_ch[i] = sp.ch; s_dis[i] = sp.dis; s_qpv[i] = sp.qpv;
                    // ESS update utils/pf.py:96-98 with E_init (quirk Q2) and delta_t
                    e_next[i] = e_init[i] + c.delta_t * (c.eta_ch * sp.ch - c.inv_eta_dis * sp.dis);
                    e_bad = e_bad || (e_next[i] < c.e_next_lb);        // E_next in NonNegativeReals (pf.py:46)
                    if (MODE == MODE_STEP) {                           // reward terms (:681-684), left to right
                        const double t0 = price * sp.pred, t1 = c.pv_cost * sp.qpv, t2 = c.ess_cost * (sp.ch + sp.dis),
                                     t3 = c.discomfort_coeff * (sp.pred * sp.pred);
                        if (i == 0) { rev = t0; der = t1; ess = t2; disc = t3; }
                        else { rev = rev + t0; der = der + t1; ess = ess + t2; disc = disc + t3; }
                    }
                }
            }
            // parked in the env's voltage row (unused until the final pass) for the iteration
#pragma unroll
            for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                vrow[i] = s_pred[i]; vrow[5 + i] = s_ch[i]; vrow[10 + i] = s_dis[i]; vrow[15 + i] = s_qpv[i];
                vrow[20 + i] = e_next[i];
            }
            vrow[25] = rev; vrow[26] = der; vrow[27] = ess; vrow[28] = disc; vrow[29] = cum; vrow[30] = price;
            vrow[31] = u2d(pack2(start, steps)); vrow[32] = u2d(pack2(hist_n, episode));
        }
        __syncwarp();                                                  // the partner reads the modified injections

        // ------------------------------------------------------------ power flow
        PState st;
        p_iterate(role, row2, ltp, st, c.pf_tol, c.pf_max_iter, valid);
        if (owner) {
            const volatile double* park = vrow;
#pragma unroll
            for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                s_pred[i] = park[i]; s_ch[i] = park[5 + i]; s_dis[i] = park[10 + i]; s_qpv[i] = park[15 + i];
                e_next[i] = park[20 + i];
            }
            rev = park[25]; der = park[26]; ess = park[27]; disc = park[28]; cum = park[29]; price = park[30];
            const uint64_t t0 = d2u(park[31]), t1 = d2u(park[32]);
            start = (int32_t)(uint32_t)t0; steps = (int32_t)(t0 >> 32);
            hist_n = (int32_t)(uint32_t)t1; episode = (int32_t)(t1 >> 32);
        }
        __syncwarp();                                                  // parked values are out before V rows come in
        const PSolve sv = p_finish(role, &c, any_imax, row2, ltp, vrow, st, valid);
        __syncwarp();                                                  // both lanes' V entries and (P, Q) pairs are in
        const bool inject = owner && (q.inject != nullptr) && (q.inject[e] != 0);
        const bool ok = __shfl_sync(FULL, (int)(sv.ok && !inject && !e_bad), lane & ~1) != 0;    // the owner's verdict, for both lanes

        if (owner && ok && q.pfl != nullptr) {                         // optional line-flow dump (parity/debug)
            double* pf = q.pfl + e * nl; double* qf = q.qfl + e * nl; double* lf = q.isq + e * nl;
            for (int k = 0; k < nl; ++k) { const double2 f = row2e[T.pos_of_lane[k]]; pf[T.col_of_lane[k]] = f.x; qf[T.col_of_lane[k]] = f.y; }
            (void)lf;                                                  // currents: below, from the registers of both lanes
        }
        if (valid && ok && q.isq != nullptr) {                         // each lane dumps the currents it holds
            double* lf = q.isq + e * nl;
            static_for<NS>([&](auto j) {
                constexpr int S = decltype(j)::value;
                if (!(role && PAD1<S>)) lf[role ? COL1<S> : COL0<S>] = st.ell[S];
            });
        }
        if (MODE == MODE_STEP && owner && !ok) {
            // roll back to the last valid state (:318-328): voltages, setpoints, reward terms
            const double* Vold = q.V + e * nb;
            for (int b = 0; b < nb; ++b) vrow[b] = Vold[b];
            const double* sprow = q.setp + e * 4 * na;
#pragma unroll
            for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                if (i < na) {
                    s_pred[i] = sprow[i]; s_ch[i] = sprow[na + i]; s_dis[i] = sprow[2 * na + i]; s_qpv[i] = sprow[3 * na + i];
                    const double t0 = price * s_pred[i], t1 = c.pv_cost * s_qpv[i], t2 = c.ess_cost * (s_ch[i] + s_dis[i]),
                                 t3 = c.discomfort_coeff * (s_pred[i] * s_pred[i]);
                    if (i == 0) { rev = t0; der = t1; ess = t2; disc = t3; }
                    else { rev = rev + t0; der = der + t1; ess = ess + t2; disc = disc + t3; }
                }
            }
        }

        // ------------------------------------------------------------ constraint masks, penalty
        uint32_t vm = sv.vm;
        const uint32_t lm = ok ? sv.lm : 0u;
        double vpen = 0.0;
        if (owner) {
            vpen = voltage_penalty(T, c, vrow, nl, MODE == MODE_STEP && !ok, vm);
            vpen = vpen + c.slack_pen;
        }
        const uint64_t vmask = ((uint64_t)vm << 1) | (uint64_t)(c.slack_viol & 1);
        const int vcount = __popc(vm) + (c.slack_viol & 1);

        // ------------------------------------------------------------ reward, bookkeeping, write back
        double reward_info = 0.0, reward = 0.0;
        bool done = false;
        if (owner) {
            uint64_t r[FP_REC_STRIDE];
#pragma unroll
            for (int i = 0; i < FP_REC_STRIDE; ++i) r[i] = 0ull;
            if (MODE == MODE_STEP) {
                reward_info = (((rev - der) - ess) - disc) - vpen;                       // :686
                reward = ok ? reward_info : (reward_info - c.fail_penalty);              // :336
                const int steps_new = steps + 1;                                         // :342
                done = (steps_new >= c.episode_limit) || !ok;                            // :345-348
                if (q.info != nullptr) {                 // info['reward'] is pre-penalty (:697), cumulative before adding (:703)
                    double2* io = reinterpret_cast<double2*>(q.info + e * FP_INFO_STRIDE);
                    io[0] = make_double2(reward_info, rev); io[1] = make_double2(der, ess);
                    io[2] = make_double2(disc, vpen); io[3] = make_double2(cum, ok ? 0.0 : 1.0);
                }
                q.reward[e] = reward;
                q.done[e] = done ? 1 : 0;
                // success: E_cur <- E_next; failure: E_cur stays (rolled back).  E_init <- E_cur (:354)
#pragma unroll
                for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                    if (i < na) {
                        const uint64_t en = ok ? d2u(e_next[i]) : rec[FP_REC_E_CUR + i];
                        r[FP_REC_E_INIT + i] = en; r[FP_REC_E_CUR + i] = en;
                    }
                }
                r[FP_REC_CUM] = d2u(cum + reward);                                       // :343
                r[FP_REC_TIME] = pack2(start, steps_new);
                r[FP_REC_HIST] = pack2(hist_n, episode);
                r[FP_REC_COUNTS] = pack2(vcount, (done ? FP_FLAG_DONE : 0) | (ok ? 0 : FP_FLAG_FAILED));
            } else {
#pragma unroll
                for (int i = 0; i < FP_MAX_AGENTS; ++i) {
                    if (i < na) {
                        r[FP_REC_E_INIT + i] = d2u(e_init[i]);                            // stays E0 (Q2)
                        r[FP_REC_E_CUR + i] = d2u(ok ? e_next[i] : e_init[i]);            // :147
                    }
                }
                r[FP_REC_CUM] = 0ull;                                                     // :77
                r[FP_REC_TIME] = pack2(start, 1);                                         // :76
                r[FP_REC_HIST] = pack2(0, episode + 1);                                   // :79-80
                r[FP_REC_COUNTS] = pack2(vcount, ok ? 0 : FP_FLAG_RESET_FAILED);
            }
            r[FP_REC_VMASK] = vmask;
            r[FP_REC_LINES] = pack2((int32_t)lm, sv.iters);
            ulonglong2* r2 = reinterpret_cast<ulonglong2*>(rec);
#pragma unroll
            for (int i = 0; i < FP_REC_STRIDE / 2; ++i) r2[i] = make_ulonglong2(r[2 * i], r[2 * i + 1]);
            if (ok || MODE == MODE_RESET) {
                double* so = q.setp + e * 4 * na;
                if (na == FP_MAX_AGENTS) {               // 160-byte row: ten 16-byte stores
                    double2* s2 = reinterpret_cast<double2*>(so);
                    s2[0] = make_double2(s_pred[0], s_pred[1]); s2[1] = make_double2(s_pred[2], s_pred[3]);
                    s2[2] = make_double2(s_pred[4], s_ch[0]); s2[3] = make_double2(s_ch[1], s_ch[2]);
                    s2[4] = make_double2(s_ch[3], s_ch[4]); s2[5] = make_double2(s_dis[0], s_dis[1]);
                    s2[6] = make_double2(s_dis[2], s_dis[3]); s2[7] = make_double2(s_dis[4], s_qpv[0]);
                    s2[8] = make_double2(s_qpv[1], s_qpv[2]); s2[9] = make_double2(s_qpv[3], s_qpv[4]);
                } else {
#pragma unroll
                    for (int i = 0; i < FP_MAX_AGENTS; ++i)
                        if (i < na) { so[i] = s_pred[i]; so[na + i] = s_ch[i]; so[2 * na + i] = s_dis[i]; so[3 * na + i] = s_qpv[i]; }
                }
            }
        }
        // voltages: coalesced rows; a failed step keeps the old row (the tile holds it), a failed
        // reset leaves the stored voltages untouched
        __syncwarp();
        const uint32_t wv2 = __ballot_sync(FULL, owner && (ok || MODE == MODE_STEP));     // bit 2j = env j
        uint32_t emask = 0u;
#pragma unroll
        for (int j = 0; j < TILE; ++j) emask |= ((wv2 >> (2 * j)) & 1u) << j;
        store_rows16(vt_tile, q.V, e0, nb, emask, lane);

        if (MODE == MODE_STEP && q.stats_partial != nullptr) {
#pragma unroll
            for (int s = 0; s < 12; ++s) {
                double x;
                if (s == FP_INFO_REWARD) x = reward_info;
                else if (s == FP_INFO_REVENUE) x = rev;
                else if (s == FP_INFO_DER_COST) x = der;
                else if (s == FP_INFO_ESS_COST) x = ess;
                else if (s == FP_INFO_DISCOMFORT) x = disc;
                else if (s == FP_INFO_VOLTAGE_PENALTY) x = vpen;
                else if (s == FP_INFO_CUMULATIVE) x = cum;
                else if (s == FP_INFO_SOLVER_FAILED) x = ok ? 0.0 : 1.0;
                else if (s == 8) x = (double)vcount;
                else if (s == 9) x = 1.0;
                else if (s == 10) x = done ? 1.0 : 0.0;
                else x = (double)__popc(lm);
                x = warp_sum_xor(owner ? x : 0.0);
                if (lane == s) stat_acc += x;
            }
        }
        __syncwarp();                                                  // tiles are reused by the next tile's loads
    }

    if (MODE == MODE_STEP && q.stats_partial != nullptr && lane < FP_NSTATS)
        atomicAdd(&q.stats_partial[((int64_t)blockIdx.x * CTA_WARPS + warp) * FP_NSTATS + lane], stat_acc);   // this warp owns the row
}

// ---------------------------------------------------------------------------- power flow only
__global__ void __launch_bounds__(32 * CTA_WARPS, 8) k_power_flow_p(const PfParamsP prm) {
    extern __shared__ double smem[];
    const PfParams& q = prm.p;
    const PairTopo& T = prm.t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool role = (lane & 1) != 0;
    const int el = lane >> 1;
    const int nl = q.nl, nb = q.nl + 1;
    double* lt = smem + CTA_WARPS * warp_smem_doubles_p();
    stage_line_table(lt, T);
    double* st_tile = smem + warp * warp_smem_doubles_p();
    double* vt_tile = st_tile + TILE * PROW;
    double2* row2 = reinterpret_cast<double2*>(st_tile + el * PROW) + (role ? NS : 0);
    const double* ltp = lt + (role ? 4 * NS : 0);
    double* vrow = vt_tile + el * VROW;
    const int my_pos = (lane < nl) ? T.pos_of_col[lane] : 0;
    const int64_t n_tiles = (q.n + TILE - 1) / TILE;
    const int64_t w_stride = (int64_t)gridDim.x * CTA_WARPS;
    for (int64_t tile = (int64_t)blockIdx.x * CTA_WARPS + warp; tile < n_tiles; tile += w_stride) {
        const int64_t e0 = tile * TILE, e = e0 + el;
        const bool valid = e < q.n;
        const uint32_t wm2 = __ballot_sync(FULL, valid);
        uint32_t emask = 0u;
#pragma unroll
        for (int j = 0; j < TILE; ++j) emask |= ((wm2 >> (2 * j)) & 1u) << j;
#pragma unroll
        for (int j = 0; j < TILE; ++j) {
            if (((emask >> j) & 1u) && lane < nl) {
                cp_async8(st_tile + j * PROW + 2 * my_pos, q.p + (e0 + j) * nl + lane);
                cp_async8(st_tile + j * PROW + 2 * my_pos + 1, q.q + (e0 + j) * nl + lane);
            }
        }
        if (role) { row2[0] = make_double2(0.0, 0.0); row2[NS - 1] = make_double2(0.0, 0.0); }
        cp_async_wait_all();
        __syncwarp();
        PState st;
        p_iterate(role, row2, ltp, st, q.tol, q.max_iter, valid);
        const PSolve sv = p_finish(role, nullptr, false, row2, ltp, vrow, st, valid);
        __syncwarp();
        store_rows16(vt_tile, q.V, e0, nb, emask, lane);
        // line flows leave straight from the (P, Q) pairs of the S tile, the currents through the V tile
        if (q.Pl != nullptr || q.Ql != nullptr) {
#pragma unroll 4
            for (int j = 0; j < TILE; ++j) {
                if (((emask >> j) & 1u) && lane < nl) {
                    if (q.Pl != nullptr) q.Pl[(e0 + j) * nl + lane] = st_tile[j * PROW + 2 * my_pos];
                    if (q.Ql != nullptr) q.Ql[(e0 + j) * nl + lane] = st_tile[j * PROW + 2 * my_pos + 1];
                }
            }
        }
        if (q.Isq != nullptr) {
            __syncwarp();
            if (valid) {
                static_for<NS>([&](auto j) {
                    constexpr int S = decltype(j)::value;
                    if (!(role && PAD1<S>)) vrow[role ? COL1<S> : COL0<S>] = st.ell[S];
                });
            }
            __syncwarp(); store_rows16(vt_tile, q.Isq, e0, nl, emask, lane);
        }
        if (valid && !role) {
            if (q.iters != nullptr) q.iters[e] = sv.iters;
            if (q.fail != nullptr) q.fail[e] = sv.ok ? 0 : 1;
        }
        __syncwarp();
    }
}

}  // namespace

// ---------------------------------------------------------------------------- launchers
size_t pair_kernel_smem_bytes() { return (size_t)cta_smem_doubles_p() * sizeof(double); }

cudaError_t pair_kernels_configure() {
    const int bytes = (int)pair_kernel_smem_bytes();
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_env_p<MODE_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_env_p<MODE_RESET>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_power_flow_p, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    return cudaSuccess;
}

int pair_kernel_max_grid(int mode) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t bytes = pair_kernel_smem_bytes();
    cudaError_t err;
    if (mode == MODE_STEP) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_env_p<MODE_STEP>, 32 * CTA_WARPS, bytes);
    else if (mode == MODE_RESET) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_env_p<MODE_RESET>, 32 * CTA_WARPS, bytes);
    else err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_power_flow_p, 32 * CTA_WARPS, bytes);
    if (err != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * sms;
}

cudaError_t launch_env_p(int mode, const EnvParamsP& prm, int grid, cudaStream_t st) {
    const size_t bytes = pair_kernel_smem_bytes();
    if (mode == MODE_STEP) k_env_p<MODE_STEP><<<grid, 32 * CTA_WARPS, bytes, st>>>(prm);
    else k_env_p<MODE_RESET><<<grid, 32 * CTA_WARPS, bytes, st>>>(prm);
    return cudaGetLastError();
}

cudaError_t launch_power_flow_p(const PfParamsP& prm, int grid, cudaStream_t st) {
    k_power_flow_p<<<grid, 32 * CTA_WARPS, pair_kernel_smem_bytes(), st>>>(prm);
    return cudaGetLastError();
}

// Pair tables from the thread tables of a feeder with the IEEE 33-bus shape (SHAPE_IEEE33).
void pair_topo_from(const ThreadTopo& t, PairTopo& p) {
    int pos_of_lane[FP_NL];
    for (int k = 0; k < NS; ++k) pos_of_lane[k] = k;
    for (int s = 0; s < NS; ++s)
        if (PairMap::K1[s] >= 0) pos_of_lane[PairMap::K1[s]] = NS + s;
    for (int i = 0; i < NPOS; ++i) { p.R[i] = 0.0; p.X[i] = 0.0; p.Z2h[i] = 0.0; p.imax2[i] = INFINITY; }
    for (int k = 0; k < t.nl; ++k) {
        const int i = pos_of_lane[k];
        p.R[i] = t.R[k]; p.X[i] = t.X[k]; p.Z2h[i] = t.Z2h[k]; p.imax2[i] = t.imax2[k];
        p.pos_of_lane[k] = (int8_t)i;
        p.col_of_lane[k] = t.col[k];
        p.pos_of_col[t.col[k]] = (int8_t)i;
    }
    for (int i = 0; i < 8; ++i) p.agent_pos[i] = (int8_t)pos_of_lane[t.agent_lane[i] >= 0 ? t.agent_lane[i] : 0];
    p.any_imax = t.any_imax;
    p.nl = t.nl;
}
