// flex_api.cu -- C-ABI host layer of libflexgpu (see include/flexgpu.h).
//
// Owns the per-GPU handle: configuration, the lane-ordered topology tables (the host-side
// equivalent of utils/create_net.py:8-38 plus the DFS pre-order the kernels need), the
// device-resident profile dataset and the per-env state arrays.  No torch types cross this
// boundary; every pointer is a plain host or device pointer.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <cstdint>
#include <string>
#include <vector>

#include "flex_kernels.cuh"
#include "predictor.cuh"

#define FP_HOST_CHUNKS 4      // fp_step_host can pipeline the batch in up to this many chunks (copy engines || SMs)
#define FP_HOST_STREAMS 3
#define FP_HOST_CHUNKS_DEFAULT 4
#define FP_HOST_GRAPHS 8

struct FpHandle {
    FpConfig cfg;
    DevCfg dc;
    DevTopo topo;
    ThreadTopo tt;               // thread-per-env tables (by-value kernel parameter)
    int variant = FP_VARIANT_THREAD, shape = SHAPE_RUNTIME;
    int64_t n = 0;
    int device = 0;
    int64_t T = 0;
    int32_t start_range = 0;
    // device buffers
    DevTopo* d_topo = nullptr;
    double *d_P = nullptr, *d_Q = nullptr, *d_PVP = nullptr, *d_PQD = nullptr, *d_OBS = nullptr;
    uint64_t* d_rec = nullptr;
    double *d_V = nullptr, *d_setp = nullptr, *d_hist = nullptr;
    // fp32 window ring of the observation history (fp_get_obs_view): [N][na][3H][6], last written slot, validity
    float* d_obsm = nullptr; int obs_q = 0; bool obsm_valid = false;
    // env-minor fp32 ring [H][na][6][n_pad] (fp_step_ring / fp_obs_ring): slot of the newest push, validity
    float* d_obsr = nullptr; int ring_q = 0; bool obsr_valid = false; int64_t n_pad = 0;
    double *d_pfl = nullptr, *d_qfl = nullptr, *d_isq = nullptr;
    double* d_stats_partial = nullptr;
    double safety_sP[8] = {0}, safety_sQ[8] = {0}, safety_b[8] = {0}, safety_vmin = 0, safety_vmax = 0, safety_w = 0; int safety_loaded = 0;
    uint8_t* d_retry_mask = nullptr; void* d_retry_count = nullptr;     // fp_reset_random_retry: device-built retry mask
    int stats_rows = 0, stats_cap = 0;     // rows of one launch's statistics block
    const uint8_t* d_inject = nullptr;
    int keep_flows = 0;
    // staging for the host-buffer entry points
    void* d_act_stage = nullptr; float* d_act_xlat = nullptr; double* d_reward_stage = nullptr; uint8_t* d_done_stage = nullptr;
    double* d_info_stage = nullptr;
    int grid_step = 0, grid_reset = 0, grid_pf = 0, grid_obs = 0;
    cudaStream_t host_streams[FP_HOST_STREAMS] = {nullptr, nullptr, nullptr};
    cudaEvent_t host_ev_in = nullptr, host_ev_out[FP_HOST_STREAMS] = {nullptr, nullptr, nullptr};
    int fuse_obs = 0;            // set around the fp_step of fp_step_obs: the step kernel pushes the observation
    int keep_hist = 1;           // fp_set_obs_history: 0 = the fused ring step does not maintain the fp64 history ring
    bool hist_valid = true;      // false after a ring-only step: restored from the env-minor ring on demand (hist_ensure)
    int host_chunks = 0;         // 0 = not decided yet (FLEXGPU_HOST_CHUNKS or the default)
    // the chunk pipeline of fp_step_host as instantiated CUDA graphs, one per set of (pinned) host
    // buffers: one launch per step instead of five API calls per chunk
    cudaStream_t host_origin = nullptr;
    struct HostGraph { cudaGraphExec_t exec = nullptr; std::vector<char> key; };
    HostGraph host_graphs[FP_HOST_GRAPHS];
    int host_graph_next = 0;
    int64_t launches = 0;
    std::string err;
    PredictorState pred;
};

static thread_local std::string g_create_err;

static int fail(FpHandle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

// Every entry point runs on the handle's device whatever the caller's current device is (two handles on two
// GPUs in one process); cudaSetDevice on the current device is a no-op.
#define USE_DEVICE(h)                                                                            \
    do {                                                                                         \
        cudaError_t e_ = cudaSetDevice((h)->device);                                             \
        if (e_ != cudaSuccess) return fail((h), FP_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e_)); \
    } while (0)
#define CUDA_TRY(h, expr)                                                              \
    do {                                                                               \
        cudaError_t e_ = (expr);                                                       \
        if (e_ != cudaSuccess)                                                         \
            return fail((h), FP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------ topology preprocessing
// Pre-order numbering (children visited in ascending bus position) + per-lane tables.
static int build_topology(const FpConfig& c, DevTopo& t, ThreadTopo& tt, int& shape, std::string& err) {
    const int nb = c.n_bus, nl = nb - 1;
    if (nb < 2 || nb > FP_MAX_BUS) { err = "n_bus must be in [2, 33]"; return FP_EINVAL; }
    if (c.n_agents < 1 || c.n_agents > FP_MAX_AGENTS) { err = "n_agents must be in [1, 5]"; return FP_EINVAL; }
    if (c.parent[0] != -1) { err = "slack bus must be position 0 (parent[0] == -1)"; return FP_EINVAL; }
    std::vector<std::vector<int>> kids(nb);
    for (int b = 1; b < nb; ++b) {
        int p = c.parent[b];
        if (p < 0 || p >= nb || p == b) { err = "parent[] out of range"; return FP_EINVAL; }
        kids[p].push_back(b);                          // ascending b by construction
    }
    std::vector<int> lane_of(nb, -1), bus_of(FP_NL, -1), size(nb, 1);
    // iterative pre-order
    std::vector<int> stack;
    int next = 0;
    for (int i = (int)kids[0].size() - 1; i >= 0; --i) stack.push_back(kids[0][i]);
    std::vector<int> order;
    while (!stack.empty()) {
        int b = stack.back(); stack.pop_back();
        if (lane_of[b] >= 0) { err = "topology has a cycle"; return FP_EINVAL; }
        lane_of[b] = next; bus_of[next] = b; ++next;
        order.push_back(b);
        for (int i = (int)kids[b].size() - 1; i >= 0; --i) stack.push_back(kids[b][i]);
    }
    if (next != nl) { err = "topology is not a connected radial tree rooted at the slack bus"; return FP_EINVAL; }
    for (int i = nl - 1; i >= 0; --i) {                // subtree sizes, children before parents
        int b = order[i];
        if (c.parent[b] != 0) size[c.parent[b]] += size[b];
    }
    std::memset(&t, 0, sizeof(t));
    for (int k = 0; k < FP_NL; ++k) {
        t.end[k] = k; t.col[k] = 0; t.agent[k] = -1;
        t.imax2[k] = std::numeric_limits<double>::infinity();
        uint32_t anc = 0;
        for (int r = 0; r < 5; ++r) anc |= (ANC_NONE << (6 * r));
        t.anc[k] = anc;
    }
    for (int k = 0; k < nl; ++k) {
        int b = bus_of[k];
        if (!(c.r[b] >= 0.0) || !std::isfinite(c.x[b])) { err = "line impedance must be finite"; return FP_EINVAL; }
        t.R[k] = c.r[b]; t.X[k] = c.x[b];
        t.Z2[k] = c.r[b] * c.r[b] + c.x[b] * c.x[b];   // utils/pf.py:93 R**2 + X**2
        t.imax2[k] = (c.imax[b] > 0.0) ? c.imax[b] * c.imax[b] : std::numeric_limits<double>::infinity();
        t.end[k] = k + size[b] - 1;
        t.col[k] = b - 1;
        uint32_t anc = 0;
        for (int r = 0; r < 5; ++r) {
            int a = b;
            for (int s = 0; s < (1 << r) && a > 0; ++s) a = c.parent[a];
            uint32_t al = (a > 0) ? (uint32_t)lane_of[a] : ANC_NONE;
            anc |= (al << (6 * r));
        }
        t.anc[k] = anc;
    }
    for (int i = 0; i < c.n_agents; ++i) {
        int b = c.agent_bus[i];
        if (b < 1 || b >= nb) { err = "agent_bus must be a non-slack bus position"; return FP_EINVAL; }
        if (t.agent[lane_of[b]] >= 0) { err = "two agents on the same bus"; return FP_EINVAL; }
        t.agent[lane_of[b]] = i;
        t.agent_lane[i] = lane_of[b];
        t.agent_col[i] = b - 1;
    }
    // ---- thread-per-env tables: parent lanes -> carry/slot tables (same constexpr derivation
    //      the built-in shapes use at compile time)
    std::memset(&tt, 0, sizeof(tt));
    int8_t par_lane[FP_NL];
    for (int k = 0; k < FP_NL; ++k) par_lane[k] = -1;
    for (int k = 0; k < nl; ++k) {
        int b = bus_of[k];
        par_lane[k] = (int8_t)((c.parent[b] > 0) ? lane_of[c.parent[b]] : -1);
    }
    const TreeTables tb = derive_tree_tables(par_lane, nl);
    bool any_imax = false;
    for (int k = 0; k < FP_NL; ++k) {
        tt.R[k] = t.R[k]; tt.X[k] = t.X[k]; tt.Z2h[k] = 0.5 * t.Z2[k]; tt.imax2[k] = t.imax2[k];
        tt.Rf[k] = (float)tt.R[k]; tt.Xf[k] = (float)tt.X[k]; tt.Z2hf[k] = (float)tt.Z2h[k];
        any_imax = any_imax || (k < nl && std::isfinite(t.imax2[k]));
        tt.par_src[k] = tb.par_src[k]; tt.own_slot[k] = tb.own_slot[k]; tt.dep_slot[k] = tb.dep_slot[k];
        tt.dep_first[k] = tb.dep_first[k]; tt.next_is_child[k] = tb.next_is_child[k];
        tt.chain_of[k] = tb.chain_of[k]; tt.attach_mask[k] = tb.attach_mask[k];
        tt.col[k] = (int8_t)t.col[k];
    }
    for (int c = 0; c < FP_MAX_CHAINS; ++c) tt.child_mask[c] = tb.child_mask[c];
    tt.n_chains = tb.n_chains;
    for (int k = 0; k < nl; ++k) tt.lane_of_col[t.col[k]] = (int8_t)k;
    for (int i = 0; i < 8; ++i) { tt.agent_lane[i] = (int8_t)t.agent_lane[i]; tt.agent_col[i] = (int8_t)t.agent_col[i]; }
    tt.nl = nl; tt.n_slots = tb.n_slots; tt.any_imax = any_imax ? 1 : 0;
    shape = thread_shape_of(tt, par_lane);
    if (c.n_agents != FP_MAX_AGENTS) shape = SHAPE_RUNTIME;      // the built-in shape unrolls the reference's five buildings
    return FP_OK;
}

static void fill_devcfg(const FpConfig& c, DevCfg& d) {
    std::memset(&d, 0, sizeof(d));
    d.nb = c.n_bus; d.nl = c.n_bus - 1; d.na = c.n_agents; d.history = c.history;
    d.episode_limit = c.episode_limit; d.raw_actions = c.raw_actions; d.pf_max_iter = c.pf_max_iter;
    d.obs_w = 6 * c.history;
    d.pf_tol = c.pf_tol; d.v_min = c.v_min; d.v_max = c.v_max; d.e_min = c.e_min; d.e_max = c.e_max;
    d.p_ch_max = c.p_ch_max; d.p_dis_max = c.p_dis_max; d.eta_ch = c.eta_ch; d.eta_dis = c.eta_dis;
    d.inv_eta_ch = 1.0 / c.eta_ch;
    d.inv_eta_dis = 1.0 / c.eta_dis;                   // python evaluates (1 / eta_dis) first (:634, pf.py:97)
    d.mpr = c.max_power_reduction; d.kappa = c.kappa; d.pv_cost = c.pv_cost; d.ess_cost = c.ess_cost;
    d.discomfort_coeff = c.discomfort_coeff; d.voltage_coeff = c.voltage_coeff; d.delta_t = c.delta_t;
    d.fail_penalty = c.fail_penalty; d.e_next_lb = c.e_next_lb;
    // slack bus: V = sqrt(1) = 1 (pf.py:51-53); its penalty term is a constant
    const double over = 1.0 - c.v_max, under = c.v_min - 1.0;
    d.slack_viol = (over > 0.0 || under > 0.0) ? 1 : 0;
    d.slack_pen = d.slack_viol ? c.voltage_coeff * ((over > under) ? over : under) : 0.0;
    d.pf_f32 = c.pf_f32_passes;
}

static int grid_for(int64_t n, int cap) {
    int64_t ctas = (n + (FP_CTA_THREADS / 32) - 1) / (FP_CTA_THREADS / 32);
    if (ctas < 1) ctas = 1;
    return (int)((ctas < cap) ? ctas : cap);
}

extern "C" {

int fp_create(const FpConfig* cfg, int64_t n_envs, int device, FpHandle** out) {
    if (!cfg || !out || n_envs < 1) return fail(nullptr, FP_EINVAL, "fp_create: bad arguments");
    if (cfg->history < 1 || cfg->episode_limit < 2 || cfg->pf_max_iter < 1 || !(cfg->pf_tol > 0.0) ||
        cfg->pf_f32_passes < 0 || cfg->pf_f32_passes >= cfg->pf_max_iter ||
        !(cfg->eta_ch > 0.0) || !(cfg->eta_dis > 0.0))
        return fail(nullptr, FP_EINVAL, "fp_create: bad scalar configuration");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(nullptr, FP_ECUDA, "fp_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, FP_EINVAL, "fp_create: bad device index");
    FpHandle* h = new (std::nothrow) FpHandle();
    if (!h) return fail(nullptr, FP_ENOMEM, "fp_create: out of host memory");
    h->cfg = *cfg; h->n = n_envs; h->device = device;
    std::string err;
    int rc = build_topology(*cfg, h->topo, h->tt, h->shape, err);
    if (rc != FP_OK) { delete h; return fail(nullptr, rc, "fp_create: " + err); }
    if (cfg->variant != FP_VARIANT_THREAD && cfg->variant != FP_VARIANT_WARP) {
        delete h; return fail(nullptr, FP_EINVAL, "fp_create: unknown kernel variant");
    }
    h->variant = cfg->variant;
    if (h->variant == FP_VARIANT_THREAD && (h->tt.n_slots > FP_MAX_SLOTS || h->tt.n_chains > FP_MAX_CHAINS)) {
        delete h; return fail(nullptr, FP_EINVAL, "fp_create: more than 8 branching buses or 16 laterals; use FP_VARIANT_WARP");
    }
    fill_devcfg(*cfg, h->dc);
    const int nb = cfg->n_bus, na = cfg->n_agents, H = cfg->history;
#define CREATE_TRY(expr)                                                                   \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            std::string m = std::string("fp_create: " #expr ": ") + cudaGetErrorString(e_); \
            fp_destroy(h);                                                                 \
            return fail(nullptr, e_ == cudaErrorMemoryAllocation ? FP_ENOMEM : FP_ECUDA, m); \
        }                                                                                  \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(cudaMalloc(&h->d_topo, sizeof(DevTopo)));
    CREATE_TRY(cudaMemcpy(h->d_topo, &h->topo, sizeof(DevTopo), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc(&h->d_rec, (size_t)n_envs * FP_REC_STRIDE * 8));
    CREATE_TRY(cudaMemset(h->d_rec, 0, (size_t)n_envs * FP_REC_STRIDE * 8));
    CREATE_TRY(cudaMalloc(&h->d_V, (size_t)n_envs * nb * 8));
    CREATE_TRY(cudaMemset(h->d_V, 0, (size_t)n_envs * nb * 8));
    CREATE_TRY(cudaMalloc(&h->d_setp, (size_t)n_envs * 4 * na * 8));
    CREATE_TRY(cudaMemset(h->d_setp, 0, (size_t)n_envs * 4 * na * 8));
    CREATE_TRY(cudaMalloc(&h->d_hist, (size_t)n_envs * H * FP_HIST_SLOT * 8));
    CREATE_TRY(cudaMemset(h->d_hist, 0, (size_t)n_envs * H * FP_HIST_SLOT * 8));
    {   // observation / state kernels are streaming kernels: a full SM worth of warps (8 CTAs of 256 threads)
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        h->grid_obs = grid_for(n_envs, 8 * (sms > 0 ? sms : 1));
    }
    if (h->variant == FP_VARIANT_THREAD) {
        CREATE_TRY(thread_kernels_configure(FP_MAX_SLOTS));
        const int64_t tiles = (n_envs + 31) / 32;
        const int cap_step = thread_kernel_max_grid(MODE_STEP, h->tt.n_slots, h->shape);
        const int cap_reset = thread_kernel_max_grid(MODE_RESET, h->tt.n_slots, h->shape);
        h->grid_step = (int)((tiles < cap_step) ? tiles : cap_step);
        if (const char* gs = std::getenv("FLEXGPU_GRID_STEP")) {      // tuning knob: persistent CTAs of the step kernel
            const int g = std::atoi(gs);
            if (g > 0 && g <= cap_step) h->grid_step = (int)((tiles < g) ? tiles : g);
        }
        h->grid_reset = (int)((tiles < cap_reset) ? tiles : cap_reset);
        h->grid_pf = thread_kernel_max_grid(MODE_PF, h->tt.n_slots, h->shape);
        h->stats_cap = cap_step;
        h->stats_rows = cap_step * (1 + FP_HOST_CHUNKS);     // + one block of rows per chunk of the pipelined host path
    } else {
        h->grid_step = grid_for(n_envs, max_resident_grid(MODE_STEP));
        h->grid_reset = grid_for(n_envs, max_resident_grid(MODE_RESET));
        h->grid_pf = max_resident_grid(MODE_PF);
        h->stats_rows = max_resident_grid(MODE_STEP); h->stats_cap = h->stats_rows;
    }
    CREATE_TRY(cudaMalloc(&h->d_stats_partial, (size_t)h->stats_rows * FP_NSTATS * 8));
    CREATE_TRY(cudaMemset(h->d_stats_partial, 0, (size_t)h->stats_rows * FP_NSTATS * 8));
#undef CREATE_TRY
    *out = h;
    return FP_OK;
}

int fp_destroy(FpHandle* h) {
    if (!h) return FP_OK;
    cudaSetDevice(h->device);
    predictor_free(&h->pred);
    cudaFree(h->d_topo); cudaFree(h->d_P); cudaFree(h->d_Q); cudaFree(h->d_PVP); cudaFree(h->d_PQD); cudaFree(h->d_OBS);
    cudaFree(h->d_rec); cudaFree(h->d_V); cudaFree(h->d_setp); cudaFree(h->d_hist); cudaFree(h->d_obsm); cudaFree(h->d_obsr);
    cudaFree(h->d_pfl); cudaFree(h->d_qfl); cudaFree(h->d_isq); cudaFree(h->d_stats_partial); cudaFree(h->d_retry_mask); cudaFree(h->d_retry_count);
    cudaFree(h->d_act_stage); cudaFree(h->d_act_xlat); cudaFree(h->d_reward_stage); cudaFree(h->d_done_stage); cudaFree(h->d_info_stage);
    for (int i = 0; i < FP_HOST_STREAMS; ++i) {
        if (h->host_streams[i]) cudaStreamDestroy(h->host_streams[i]);
        if (h->host_ev_out[i]) cudaEventDestroy(h->host_ev_out[i]);
    }
    if (h->host_ev_in) cudaEventDestroy(h->host_ev_in);
    for (auto& g : h->host_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->host_origin) cudaStreamDestroy(h->host_origin);
    delete h;
    return FP_OK;
}

const char* fp_last_error(const FpHandle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
int64_t fp_n_envs(const FpHandle* h) { return h ? h->n : 0; }
int64_t fp_launch_count(const FpHandle* h) { return h ? h->launches : 0; }
int32_t fp_sizeof_config(void) { return (int32_t)sizeof(FpConfig); }

int fp_load_profiles(FpHandle* h, const double* h_P, const double* h_Q, const double* h_PV,
                     const double* h_price, int64_t T) {
    if (!h) return FP_EINVAL;
    if (!h_P || !h_Q || !h_PV || !h_price) return fail(h, FP_EINVAL, "fp_load_profiles: null array");
    const int nl = h->dc.nl, na = h->dc.na;
    // the reference slices episode_limit + history + 1 rows per episode (:478)
    const int64_t need = (int64_t)h->cfg.episode_limit + h->cfg.history + 1;
    if (T < need) return fail(h, FP_EINVAL, "fp_load_profiles: fewer rows than one episode slice");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaFree(h->d_P); cudaFree(h->d_Q); cudaFree(h->d_PVP); cudaFree(h->d_PQD); cudaFree(h->d_OBS);
    h->d_P = h->d_Q = h->d_PVP = h->d_PQD = h->d_OBS = nullptr;
    double *d_pv = nullptr, *d_price = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->d_P, (size_t)T * nl * 8));
    CUDA_TRY(h, cudaMalloc(&h->d_Q, (size_t)T * nl * 8));
    CUDA_TRY(h, cudaMalloc(&h->d_PVP, (size_t)T * FP_PVP_STRIDE * 8));
    CUDA_TRY(h, cudaMalloc(&d_pv, (size_t)T * na * 8));
    CUDA_TRY(h, cudaMalloc(&d_price, (size_t)T * 8));
    CUDA_TRY(h, cudaMemcpy(h->d_P, h_P, (size_t)T * nl * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_Q, h_Q, (size_t)T * nl * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(d_pv, h_PV, (size_t)T * na * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(d_price, h_price, (size_t)T * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(h, launch_pack_pvp(d_pv, d_price, na, T, h->d_PVP, 0));
    h->launches++;
    CUDA_TRY(h, cudaMalloc(&h->d_OBS, (size_t)T * FP_OBS_STRIDE * 8));
    CUDA_TRY(h, launch_pack_obsrow(h->d_P, h->d_Q, h->d_PVP, h->topo.agent_col, na, nl, T, h->d_OBS, 0));
    h->launches++;
    if (h->variant == FP_VARIANT_THREAD) {   // (p, q) pairs in DFS lane order for the thread kernels
        CUDA_TRY(h, cudaMalloc(&h->d_PQD, (size_t)T * nl * 16));
        CUDA_TRY(h, launch_pack_pq(h->d_P, h->d_Q, h->tt, T, h->d_PQD, 0));
        h->launches++;
    }
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(d_pv); cudaFree(d_price);
    h->T = T;
    h->start_range = (int32_t)((T - need + 1 < 2147483647LL) ? (T - need + 1) : 2147483647LL);
    return FP_OK;
}

static void fill_env_params(FpHandle* h, EnvParams& p) {
    std::memset(&p, 0, sizeof(p));
    p.c = h->dc; p.topo = h->d_topo; p.n = h->n;
    p.P = h->d_P; p.Q = h->d_Q; p.PVP = h->d_PVP; p.PQD = h->d_PQD; p.OBSROW = h->d_OBS;
    p.rec = h->d_rec; p.V = h->d_V; p.setp = h->d_setp;
    if (h->keep_flows) { p.pfl = h->d_pfl; p.qfl = h->d_qfl; p.isq = h->d_isq; }
    p.inject = h->d_inject;
    p.start_range = h->start_range;
}

// whole-tile bulk stores (thread kernels) need 16-byte aligned output arrays
static int bulk_io_ok(const EnvParams& p) {
    auto al16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15u) == 0; };
    return (al16(p.rec) && al16(p.setp) && al16(p.V) && al16(p.reward) && al16(p.done) && (p.info == nullptr || al16(p.info))) ? 1 : 0;
}

// translate_action's constants exactly as torch forms them (utils/util.py:124-128): the clamp bounds
// and `low` become fp32 scalars, `high - low` is a Python (double) difference turned into an fp32 scalar
static void set_translate(const FpHandle* h, EnvParams& p) {
    p.act_translate = 1;
    p.act_lo = (float)h->cfg.action_low; p.act_hi = (float)h->cfg.action_high;
    p.act_span = (float)(h->cfg.action_high - h->cfg.action_low);
}

static cudaError_t launch_env_any(FpHandle* h, int mode, const EnvParams& p, cudaStream_t st) {
    const int grid = (mode == MODE_STEP) ? h->grid_step : h->grid_reset;
    if (h->variant == FP_VARIANT_THREAD) {
        EnvParamsT pt;
        pt.e = p; pt.t = h->tt;
        if (mode == MODE_STEP) pt.e.bulk_io = bulk_io_ok(p);
        return launch_env_t(mode, h->shape, pt, grid, st);
    }
    return launch_env(mode, p, grid, st);
}

int fp_reset(FpHandle* h, const int32_t* d_start, const double* d_e0, const double* d_a0,
             const uint8_t* d_mask, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_reset: call fp_load_profiles first");
    if (!d_start || !d_e0 || !d_a0) return fail(h, FP_EINVAL, "fp_reset: null input");
    EnvParams p; fill_env_params(h, p);
    p.start = d_start; p.e0 = d_e0; p.a0 = d_a0; p.mask = d_mask; p.random = 0;
    p.inject = nullptr;
    CUDA_TRY(h, launch_env_any(h, MODE_RESET, p, (cudaStream_t)stream));
    h->launches++;
    if (h->d_obsm && h->obsm_valid) {              // restart the zero padding of the reset envs' observation windows
        CUDA_TRY(h, launch_obsm_clear(h->d_obsm, d_mask, h->n, h->dc.na * 3 * h->dc.history * 6, (cudaStream_t)stream));
        h->launches++;
    }
    if (h->d_obsr && h->obsr_valid) {
        CUDA_TRY(h, launch_obsr_clear(h->d_obsr, d_mask, h->n, h->n_pad, h->dc.history * h->dc.na * 6, (cudaStream_t)stream));
        h->launches++;
    }
    return FP_OK;
}

// One launch of the random reset for the envs of d_mask (+ the observation rings' restart).
static int reset_random_once(FpHandle* h, uint64_t seed, int64_t env_offset, const uint8_t* d_mask, void* stream) {
    EnvParams p; fill_env_params(h, p);
    p.mask = d_mask; p.random = 1; p.seed = seed; p.env_offset = env_offset;
    p.inject = nullptr;
    CUDA_TRY(h, launch_env_any(h, MODE_RESET, p, (cudaStream_t)stream));
    h->launches++;
    if (h->d_obsm && h->obsm_valid) {
        CUDA_TRY(h, launch_obsm_clear(h->d_obsm, d_mask, h->n, h->dc.na * 3 * h->dc.history * 6, (cudaStream_t)stream));
        h->launches++;
    }
    if (h->d_obsr && h->obsr_valid) {
        CUDA_TRY(h, launch_obsr_clear(h->d_obsr, d_mask, h->n, h->n_pad, h->dc.history * h->dc.na * 6, (cudaStream_t)stream));
        h->launches++;
    }
    return FP_OK;
}

int fp_reset_random(FpHandle* h, uint64_t seed, int64_t env_offset, const uint8_t* d_mask, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_reset_random: call fp_load_profiles first");
    return reset_random_once(h, seed, env_offset, d_mask, stream);
}

// The reference redraws until the initial power flow is solvable (`while not solvable`, :82-153).  Here: the random
// reset, then `retries` more launches for the envs (of d_mask) whose FP_FLAG_RESET_FAILED is set -- the retry mask is
// built on the device, so nothing synchronises with the host; a launch whose mask is empty costs a few microseconds.
// Every attempt draws from the env's next Philox episode counter.  d_failed (may be NULL) receives the number of envs
// still flagged after the last attempt (int32 on the device).
int fp_reset_random_retry(FpHandle* h, uint64_t seed, int64_t env_offset, const uint8_t* d_mask, int32_t retries,
                          int32_t* d_failed, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_reset_random_retry: call fp_load_profiles first");
    if (retries < 0 || retries > 64) return fail(h, FP_EINVAL, "fp_reset_random_retry: 0 <= retries <= 64");
    int rc = reset_random_once(h, seed, env_offset, d_mask, stream);
    if (rc != FP_OK) return rc;
    if (retries > 0 || d_failed) {
        if (!h->d_retry_mask) CUDA_TRY(h, cudaMalloc(&h->d_retry_mask, (size_t)h->n + 4));
        int32_t* cnt = reinterpret_cast<int32_t*>(h->d_retry_count);
        if (!cnt) { CUDA_TRY(h, cudaMalloc(&h->d_retry_count, 4)); cnt = reinterpret_cast<int32_t*>(h->d_retry_count); }
        for (int r = 0; r <= retries; ++r) {
            CUDA_TRY(h, launch_reset_failed_mask(h->d_rec, d_mask, h->n, h->d_retry_mask, cnt, (cudaStream_t)stream));
            h->launches++;
            if (r == retries) break;
            rc = reset_random_once(h, seed, env_offset, h->d_retry_mask, stream);
            if (rc != FP_OK) return rc;
        }
        if (d_failed) CUDA_TRY(h, cudaMemcpyAsync(d_failed, cnt, 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return FP_OK;
}

int fp_step(FpHandle* h, const void* d_actions, int act_dtype, double* d_reward, uint8_t* d_done,
            double* d_info, const uint8_t* d_mask, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_step: call fp_load_profiles first");
    if (!d_actions || !d_reward || !d_done) return fail(h, FP_EINVAL, "fp_step: null input/output");
    if (act_dtype != FP_F32 && act_dtype != FP_F64 && act_dtype != FP_F32_POLICY) return fail(h, FP_EINVAL, "fp_step: bad action dtype");
    EnvParams p; fill_env_params(h, p);
    p.actions = d_actions; p.act_f64 = (act_dtype == FP_F64);
    if (act_dtype == FP_F32_POLICY) {
        if (h->variant == FP_VARIANT_THREAD) set_translate(h, p);      // fused into the step kernel
        else {                                                                     // other variants: elementwise pre-pass
            const size_t cnt = (size_t)h->n * h->dc.na * 4;
            if (!h->d_act_xlat) CUDA_TRY(h, cudaMalloc(&h->d_act_xlat, cnt * 4));
            EnvParams t; set_translate(h, t);
            CUDA_TRY(h, launch_translate_actions((const float*)d_actions, h->d_act_xlat, (int64_t)cnt, t.act_lo, t.act_hi, t.act_span,
                                                 (cudaStream_t)stream));
            h->launches++;
            p.actions = h->d_act_xlat;
        }
    }
    p.reward = d_reward; p.done = d_done; p.info = d_info; p.mask = d_mask;
    p.stats_partial = h->d_stats_partial;
    if (h->fuse_obs) { p.obs_push = 1; p.obs_q = h->ring_q; p.hist = h->keep_hist ? h->d_hist : nullptr; p.obsr = h->d_obsr; p.n_pad = h->n_pad; }
    CUDA_TRY(h, launch_env_any(h, MODE_STEP, p, (cudaStream_t)stream));
    h->launches++;
    return FP_OK;
}

// The chunk pipeline of fp_step_host, enqueued behind everything already on `st` (no synchronisation).
static int enqueue_host_chunks(FpHandle* h, const void* h_actions, int act_dtype, double* h_reward, uint8_t* h_done,
                               double* h_info, int n_chunks, cudaStream_t st) {
    const size_t n = (size_t)h->n, na = (size_t)h->dc.na;
    const size_t abpe = na * 4 * (act_dtype == FP_F64 ? 8 : 4);          // action bytes per env
    const int64_t tiles = ((int64_t)n + 31) / 32;
    // everything already queued on the caller's stream happens before the chunks
    cudaStream_t s_in = h->host_streams[0], s_k = h->host_streams[1], s_out = h->host_streams[2];
    CUDA_TRY(h, cudaEventRecord(h->host_ev_in, st));
    for (int i = 0; i < FP_HOST_STREAMS; ++i) CUDA_TRY(h, cudaStreamWaitEvent(h->host_streams[i], h->host_ev_in, 0));
    const int cap = h->stats_cap;                                       // statistics rows of one launch
    const int cap_grid = cap;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t t0 = tiles * c / n_chunks, t1 = tiles * (c + 1) / n_chunks;
        const size_t e0 = (size_t)t0 * 32, e1 = ((size_t)t1 * 32 < n) ? (size_t)t1 * 32 : n, ne = e1 - e0;
        // host -> device copies run back to back on their own stream (the link never idles) ...
        CUDA_TRY(h, cudaMemcpyAsync((char*)h->d_act_stage + e0 * abpe, (const char*)h_actions + e0 * abpe, ne * abpe,
                                    cudaMemcpyHostToDevice, s_in));
        CUDA_TRY(h, cudaEventRecord(h->host_ev_out[0], s_in));
        // ... the kernel of a chunk follows its copy ...
        CUDA_TRY(h, cudaStreamWaitEvent(s_k, h->host_ev_out[0], 0));
        EnvParams p; fill_env_params(h, p);
        p.actions = h->d_act_stage; p.act_f64 = (act_dtype == FP_F64);
        if (act_dtype == FP_F32_POLICY) set_translate(h, p);
        p.reward = h->d_reward_stage; p.done = h->d_done_stage; p.info = h_info ? h->d_info_stage : nullptr;
        p.stats_partial = h->d_stats_partial + (size_t)(1 + c) * cap * FP_NSTATS;
        p.tile_begin = t0; p.tile_end = t1;
        const int grid = (int)((t1 - t0 < cap_grid) ? (t1 - t0) : cap_grid);
        {
            EnvParamsT pt; pt.e = p; pt.t = h->tt;
            pt.e.bulk_io = bulk_io_ok(p);
            CUDA_TRY(h, launch_env_t(MODE_STEP, h->shape, pt, grid, s_k));
        }
        h->launches++;
        CUDA_TRY(h, cudaEventRecord(h->host_ev_out[1], s_k));
        // ... and its results leave on the other copy engine while the next chunk computes
        CUDA_TRY(h, cudaStreamWaitEvent(s_out, h->host_ev_out[1], 0));
        CUDA_TRY(h, cudaMemcpyAsync(h_reward + e0, h->d_reward_stage + e0, ne * 8, cudaMemcpyDeviceToHost, s_out));
        CUDA_TRY(h, cudaMemcpyAsync(h_done + e0, h->d_done_stage + e0, ne, cudaMemcpyDeviceToHost, s_out));
        if (h_info) CUDA_TRY(h, cudaMemcpyAsync(h_info + e0 * FP_INFO_STRIDE, h->d_info_stage + e0 * FP_INFO_STRIDE,
                                                ne * FP_INFO_STRIDE * 8, cudaMemcpyDeviceToHost, s_out));
    }
    for (int i = 0; i < FP_HOST_STREAMS; ++i) {
        CUDA_TRY(h, cudaEventRecord(h->host_ev_out[i], h->host_streams[i]));
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->host_ev_out[i], 0));
    }
    return FP_OK;
}

// Host-buffer step.  The batch is cut into chunks of 32-env tiles that flow through three internal
// streams (copy in / kernels / copy out): the host->device copies run back to back, the kernel of
// chunk c overlaps the copy of chunk c+1 and the device->host copies of chunk c-1 (both copy
// engines and the SMs busy at once).  Chunks touch disjoint envs and disjoint statistics rows, so
// results are identical to one full-batch launch.  With pinned host buffers the whole pipeline is
// one instantiated CUDA graph per buffer set.
int fp_step_host(FpHandle* h, const void* h_actions, int act_dtype, double* h_reward, uint8_t* h_done,
                 double* h_info, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_step_host: call fp_load_profiles first");
    if (!h_actions || !h_reward || !h_done) return fail(h, FP_EINVAL, "fp_step_host: null buffer");
    if (act_dtype != FP_F32 && act_dtype != FP_F64 && act_dtype != FP_F32_POLICY) return fail(h, FP_EINVAL, "fp_step_host: bad action dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->n, na = (size_t)h->dc.na;
    if (!h->d_act_stage) {
        CUDA_TRY(h, cudaMalloc(&h->d_act_stage, n * na * 4 * 8));
        CUDA_TRY(h, cudaMalloc(&h->d_reward_stage, n * 8));
        CUDA_TRY(h, cudaMalloc(&h->d_done_stage, n));
        CUDA_TRY(h, cudaMalloc(&h->d_info_stage, n * FP_INFO_STRIDE * 8));
    }
    const size_t abpe = na * 4 * (act_dtype == FP_F64 ? 8 : 4);          // action bytes per env
    const int64_t tiles = ((int64_t)n + 31) / 32;
    if (h->host_chunks == 0) {
        const char* ev = std::getenv("FLEXGPU_HOST_CHUNKS");
        int k = ev ? std::atoi(ev) : FP_HOST_CHUNKS_DEFAULT;
        h->host_chunks = (k < 1) ? 1 : ((k > FP_HOST_CHUNKS) ? FP_HOST_CHUNKS : k);
    }
    const int n_chunks = h->host_chunks;
    if (h->variant != FP_VARIANT_THREAD || n_chunks == 1 || tiles < 4 * n_chunks) {  // small batch / warp variant: one shot
        CUDA_TRY(h, cudaMemcpyAsync(h->d_act_stage, h_actions, n * abpe, cudaMemcpyHostToDevice, st));
        int rc = fp_step(h, h->d_act_stage, act_dtype, h->d_reward_stage, h->d_done_stage,
                         h_info ? h->d_info_stage : nullptr, nullptr, stream);
        if (rc != FP_OK) return rc;
        CUDA_TRY(h, cudaMemcpyAsync(h_reward, h->d_reward_stage, n * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(h, cudaMemcpyAsync(h_done, h->d_done_stage, n, cudaMemcpyDeviceToHost, st));
        if (h_info) CUDA_TRY(h, cudaMemcpyAsync(h_info, h->d_info_stage, n * FP_INFO_STRIDE * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(h, cudaStreamSynchronize(st));
        return FP_OK;
    }
    if (!h->host_ev_in) {
        CUDA_TRY(h, cudaSetDevice(h->device));
        for (int i = 0; i < FP_HOST_STREAMS; ++i) {
            CUDA_TRY(h, cudaStreamCreateWithFlags(&h->host_streams[i], cudaStreamNonBlocking));
            CUDA_TRY(h, cudaEventCreateWithFlags(&h->host_ev_out[i], cudaEventDisableTiming));
        }
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->host_ev_in, cudaEventDisableTiming));
    }
    // Pinned buffers: the whole chunk pipeline is one instantiated graph per buffer set.
    auto pinned = [](const void* x) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, x) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool all_pinned = pinned(h_actions) && pinned(h_reward) && pinned(h_done) && (!h_info || pinned(h_info));
    // Zero-copy: pinned host memory is part of the unified address space, so the step kernel reads the
    // actions of tile i+1 straight over PCIe while it sweeps tile i (its input stage is asynchronous and
    // one tile ahead) and stores reward / done / info straight into the host buffers -- no staging copies,
    // no chunking, one launch.  (thread variant, fp32 actions: their rows are staged by LDGSTS.)
    // FLEXGPU_HOST_ZEROCOPY: 2 (default) = inputs and outputs, 1 = inputs only, 0 = staged chunk pipeline below
    // (measured at 131 072 envs: 294 / 310 / 310 us per step; the 10.5 MB of actions alone take 240 us by DMA)
    static const int zc_mode = std::getenv("FLEXGPU_HOST_ZEROCOPY") ? std::atoi(std::getenv("FLEXGPU_HOST_ZEROCOPY")) : 2;
    if (all_pinned && zc_mode > 0 && act_dtype != FP_F64) {
        double* out_r = h_reward; uint8_t* out_d = h_done; double* out_i = h_info;
        if (zc_mode == 1) { out_r = h->d_reward_stage; out_d = h->d_done_stage; out_i = h_info ? h->d_info_stage : nullptr; }   // inputs only
        int rc = fp_step(h, h_actions, act_dtype, out_r, out_d, out_i, nullptr, stream);
        if (rc != FP_OK) return rc;
        if (zc_mode == 1) {
            CUDA_TRY(h, cudaMemcpyAsync(h_reward, h->d_reward_stage, n * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(h, cudaMemcpyAsync(h_done, h->d_done_stage, n, cudaMemcpyDeviceToHost, st));
            if (h_info) CUDA_TRY(h, cudaMemcpyAsync(h_info, h->d_info_stage, n * FP_INFO_STRIDE * 8, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(h, cudaStreamSynchronize(st));
        return FP_OK;
    }
    const bool use_graph = all_pinned && !std::getenv("FLEXGPU_NO_HOST_GRAPH");
    if (!use_graph) {
        int rc = enqueue_host_chunks(h, h_actions, act_dtype, h_reward, h_done, h_info, n_chunks, st);
        if (rc != FP_OK) return rc;
        CUDA_TRY(h, cudaStreamSynchronize(st));
        return FP_OK;
    }
    struct Key { EnvParams p; const void* a; void* r; void* d; void* i; int dtype, chunks; } key;
    std::memset(&key, 0, sizeof(key));
    fill_env_params(h, key.p);
    key.a = h_actions; key.r = h_reward; key.d = h_done; key.i = h_info; key.dtype = act_dtype; key.chunks = n_chunks;
    FpHandle::HostGraph* g = nullptr;
    for (auto& c : h->host_graphs)
        if (c.exec && c.key.size() == sizeof(key) && std::memcmp(c.key.data(), &key, sizeof(key)) == 0) { g = &c; break; }
    if (!g) {
        g = &h->host_graphs[h->host_graph_next];
        h->host_graph_next = (h->host_graph_next + 1) % FP_HOST_GRAPHS;
        if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
        if (!h->host_origin) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->host_origin, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        CUDA_TRY(h, cudaStreamBeginCapture(h->host_origin, cudaStreamCaptureModeThreadLocal));
        const int64_t launches0 = h->launches;
        int rc = enqueue_host_chunks(h, h_actions, act_dtype, h_reward, h_done, h_info, n_chunks, h->host_origin);
        h->launches = launches0;                                         // counted per graph launch below
        cudaError_t ce = cudaStreamEndCapture(h->host_origin, &graph);
        if (rc != FP_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        CUDA_TRY(h, ce);
        ce = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        CUDA_TRY(h, ce);
        g->key.assign(reinterpret_cast<const char*>(&key), reinterpret_cast<const char*>(&key) + sizeof(key));
    }
    CUDA_TRY(h, cudaGraphLaunch(g->exec, st));
    h->launches += n_chunks;
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return FP_OK;
}

static void fill_obs_params(FpHandle* h, ObsParams& p, void* out, int push) {
    std::memset(&p, 0, sizeof(p));
    p.c = h->dc; p.n = h->n; p.P = h->d_P; p.Q = h->d_Q; p.PVP = h->d_PVP; p.OBSROW = h->d_OBS;
    p.rec = h->d_rec; p.V = h->d_V; p.hist = h->d_hist;
    for (int i = 0; i < 8; ++i) p.agent_col[i] = h->topo.agent_col[i];
    p.out = out; p.push = push;
}

// The fp64 history ring after ring-only steps (fp_set_obs_history(h, 0)): restored from the env-minor ring, which is
// valid whenever the history is not.
static int hist_ensure(FpHandle* h, cudaStream_t st) {
    if (h->hist_valid) return FP_OK;
    if (!h->d_obsr || !h->obsr_valid) return fail(h, FP_ESTATE, "observation history lost: neither the fp64 ring nor the env-minor ring is valid");
    ObsParams p; fill_obs_params(h, p, nullptr, 0);
    CUDA_TRY(h, launch_hist_from_ring(p, h->d_obsr, h->n_pad, h->ring_q, st));
    h->launches++;
    h->hist_valid = true;
    return FP_OK;
}

int fp_set_obs_history(FpHandle* h, int keep) {
    if (!h) return FP_EINVAL;
    h->keep_hist = keep ? 1 : 0;
    return FP_OK;
}

int fp_get_obs(FpHandle* h, void* d_out, int dtype, int push, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_get_obs: call fp_load_profiles first");
    if (!d_out || (dtype != FP_F32 && dtype != FP_F64)) return fail(h, FP_EINVAL, "fp_get_obs: bad arguments");
    { const int rc = hist_ensure(h, (cudaStream_t)stream); if (rc != FP_OK) return rc; }
    ObsParams p; fill_obs_params(h, p, d_out, push ? 1 : 0);
    CUDA_TRY(h, launch_obs(p, dtype == FP_F64, h->grid_obs, (cudaStream_t)stream));
    h->launches++;
    if (push) { h->obsm_valid = false; h->obsr_valid = false; }   // the fp32 rings missed this push: rebuilt on their next use
    return FP_OK;
}

// Window ring bookkeeping: make the ring usable (allocate / rebuild) and, for a push, advance the write
// slot (compacting at the end of the ring).  Returns the slot the push writes.
static int obs_ring_prepare(FpHandle* h, bool push, cudaStream_t st) {
    const int H = h->dc.history, na = h->dc.na;
    const int64_t per_env = (int64_t)na * 3 * H * 6;
    { const int rc = hist_ensure(h, st); if (rc != FP_OK) return rc; }   // the rebuild and the push below use the fp64 ring
    ObsParams p; fill_obs_params(h, p, nullptr, 1);
    if (!h->d_obsm) {
        CUDA_TRY(h, cudaSetDevice(h->device));
        CUDA_TRY(h, cudaMalloc(&h->d_obsm, (size_t)h->n * per_env * 4));
        h->obsm_valid = false;
    }
    if (!h->obsm_valid) {                          // (re)build from the fp64 history ring: state after a push at w = H - 1
        CUDA_TRY(h, launch_obsm_rebuild(p, h->d_obsm, st));
        h->launches++;
        h->obs_q = H - 1; h->obsm_valid = true;
    }
    if (push) {
        if (h->obs_q == 3 * H - 1) {               // end of the ring: the last H-1 entries move to the front
            CUDA_TRY(h, launch_obsm_compact(h->d_obsm, h->n * na, H, st));
            h->launches++;
            h->obs_q = H - 2;
        }
        h->obs_q += 1;
    }
    return FP_OK;
}

int fp_get_obs_view(FpHandle* h, int push, float** d_view, int64_t* env_pitch, int64_t* agent_pitch, void* stream) {
    if (!h || !d_view) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_get_obs_view: call fp_load_profiles first");
    const int H = h->dc.history, na = h->dc.na;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = obs_ring_prepare(h, push != 0, st);
    if (rc != FP_OK) return rc;
    if (push) {
        ObsParams p; fill_obs_params(h, p, nullptr, 1);
        CUDA_TRY(h, launch_obs_push(p, h->d_obsm, h->obs_q, h->grid_obs, st));
        h->launches++;
        h->obsr_valid = false;
    }
    *d_view = h->d_obsm + (int64_t)(h->obs_q - H + 1) * 6;    // slots w-H+1 .. w: oldest .. newest
    if (env_pitch) *env_pitch = (int64_t)na * 3 * H * 6;
    if (agent_pitch) *agent_pitch = (int64_t)3 * H * 6;
    return FP_OK;
}

/* step + the pushing get_obs that follows it in the rollout loop, dense-view form: two launches (the one-launch
 * form pushes into the env-minor ring, fp_step_ring) */
int fp_step_obs(FpHandle* h, const void* d_actions, int act_dtype, double* d_reward, uint8_t* d_done, double* d_info,
                const uint8_t* d_mask, float** d_view, int64_t* env_pitch, int64_t* agent_pitch, void* stream) {
    if (!h || !d_view) return FP_EINVAL;
    int rc = fp_step(h, d_actions, act_dtype, d_reward, d_done, d_info, d_mask, stream);
    if (rc != FP_OK) return rc;
    return fp_get_obs_view(h, 1, d_view, env_pitch, agent_pitch, stream);
}

// Env-minor ring bookkeeping: allocate / rebuild from the fp64 history ring; for a push, advance the slot.
static int obs_envminor_prepare(FpHandle* h, bool push, cudaStream_t st) {
    const int H = h->dc.history, na = h->dc.na;
    if (!h->d_obsr) {
        CUDA_TRY(h, cudaSetDevice(h->device));
        h->n_pad = (h->n + 31) / 32 * 32;
        CUDA_TRY(h, cudaMalloc(&h->d_obsr, (size_t)H * na * 6 * h->n_pad * 4));
        h->obsr_valid = false;
    }
    if (!h->obsr_valid) {
        if (!h->hist_valid) return fail(h, FP_ESTATE, "observation history lost: neither the fp64 ring nor the env-minor ring is valid");
        ObsParams p; fill_obs_params(h, p, nullptr, 1);
        CUDA_TRY(h, launch_obsr_rebuild(p, h->d_obsr, h->n_pad, st));
        h->launches++;
        h->ring_q = H - 1; h->obsr_valid = true;
    }
    if (push) h->ring_q = (h->ring_q + 1) % H;
    return FP_OK;
}

int fp_obs_ring(FpHandle* h, float** d_ring, int32_t* slot, int64_t* n_pad, void* stream) {
    if (!h || !d_ring) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_obs_ring: call fp_load_profiles first");
    int rc = obs_envminor_prepare(h, false, (cudaStream_t)stream);
    if (rc != FP_OK) return rc;
    *d_ring = h->d_obsr;
    if (slot) *slot = h->ring_q;
    if (n_pad) *n_pad = h->n_pad;
    return FP_OK;
}

int fp_step_ring(FpHandle* h, const void* d_actions, int act_dtype, double* d_reward, uint8_t* d_done, double* d_info,
                 const uint8_t* d_mask, float** d_ring, int32_t* slot, int64_t* n_pad, void* stream) {
    if (!h || !d_ring) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_step_ring: call fp_load_profiles first");
    // one launch for the built-in feeder shape; the run-time-table kernels, the warp variant and masked steps push
    // through the generic path (fp64 history ring) and rebuild the ring from it: same contents
    const bool fuse = (h->variant == FP_VARIANT_THREAD) && d_mask == nullptr && h->shape == SHAPE_IEEE33;
    if (!fuse) {
        int rc = fp_step(h, d_actions, act_dtype, d_reward, d_done, d_info, d_mask, stream);
        if (rc != FP_OK) return rc;
        float* view = nullptr;
        rc = fp_get_obs_view(h, 1, &view, nullptr, nullptr, stream);
        if (rc != FP_OK) return rc;
        return fp_obs_ring(h, d_ring, slot, n_pad, stream);
    }
    int rc = obs_envminor_prepare(h, true, (cudaStream_t)stream);
    if (rc != FP_OK) return rc;
    h->fuse_obs = 1;
    rc = fp_step(h, d_actions, act_dtype, d_reward, d_done, d_info, nullptr, stream);
    h->fuse_obs = 0;
    if (rc != FP_OK) return rc;
    if (!h->keep_hist) h->hist_valid = false;      // ring-only stepping: the fp64 history ring missed this push (hist_ensure)
    h->obsm_valid = false;                         // the env-major window ring missed this push
    *d_ring = h->d_obsr;
    if (slot) *slot = h->ring_q;
    if (n_pad) *n_pad = h->n_pad;
    return FP_OK;
}

int fp_obs_ring_reset_push(FpHandle* h, const uint8_t* d_mask, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_obs_ring_reset_push: call fp_load_profiles first");
    int rc = obs_envminor_prepare(h, false, (cudaStream_t)stream);
    if (rc != FP_OK) return rc;
    ObsParams p; fill_obs_params(h, p, nullptr, 1);
    CUDA_TRY(h, launch_obsr_reset_push(p, h->d_obsr, h->n_pad, h->ring_q, d_mask, (cudaStream_t)stream));
    h->launches++;
    h->obsm_valid = false;
    return FP_OK;
}

int fp_obs_ring_gather(FpHandle* h, float* d_out, void* stream) {
    if (!h || !d_out) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_obs_ring_gather: call fp_load_profiles first");
    int rc = obs_envminor_prepare(h, false, (cudaStream_t)stream);
    if (rc != FP_OK) return rc;
    CUDA_TRY(h, launch_obsr_gather(h->d_obsr, d_out, h->n, h->n_pad, h->dc.na, h->dc.history, h->ring_q, (cudaStream_t)stream));
    h->launches++;
    return FP_OK;
}

// Safety layer (madrl/models/safemaddpg.py:176-299) as a batched closed-form projection: see k_safety_project.
int fp_safety_load(FpHandle* h, const double* h_coef, const double* h_intercept, double v_min, double v_max, double slack_weight) {
    if (!h) return FP_EINVAL;
    if (!h_coef || !h_intercept) return fail(h, FP_EINVAL, "fp_safety_load: null model");
    if (!(slack_weight > 0.0) || !(v_min < v_max)) return fail(h, FP_EINVAL, "fp_safety_load: bad limits / weight");
    const int nb = h->dc.nb, na = h->dc.na;
    // safemaddpg.py:182-184: W_P = coef[:, :nb], W_Q = coef[:, nb:]; :266,272 multiply their ROW SUMS by the bus's own P_net / Q_net
    for (int i = 0; i < na; ++i) {
        const int bus = h->topo.agent_col[i] + 1;                 // bus position of agent i's building
        double sp = 0.0, sq = 0.0;
        for (int j = 0; j < nb; ++j) { sp += h_coef[(size_t)bus * 2 * nb + j]; sq += h_coef[(size_t)bus * 2 * nb + nb + j]; }
        h->safety_sP[i] = sp; h->safety_sQ[i] = sq; h->safety_b[i] = h_intercept[bus];
    }
    h->safety_vmin = v_min; h->safety_vmax = v_max; h->safety_w = slack_weight; h->safety_loaded = 1;
    return FP_OK;
}

int fp_safety_project(FpHandle* h, const void* d_actions, int act_dtype, float* d_adjusted, int32_t type_major, double* d_slack,
                      uint8_t* d_intervened, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->safety_loaded) return fail(h, FP_ESTATE, "fp_safety_project: call fp_safety_load first");
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_safety_project: call fp_load_profiles first");
    if (!d_actions || !d_adjusted || (act_dtype != FP_F32 && act_dtype != FP_F64)) return fail(h, FP_EINVAL, "fp_safety_project: bad arguments");
    SafetyParams p;
    std::memset(&p, 0, sizeof(p));
    fill_obs_params(h, p.o, nullptr, 0);
    for (int i = 0; i < h->dc.na; ++i) { p.sP[i] = h->safety_sP[i]; p.sQ[i] = h->safety_sQ[i]; p.b[i] = h->safety_b[i]; }
    p.v_min = h->safety_vmin; p.v_max = h->safety_vmax; p.w = h->safety_w;
    p.actions = d_actions; p.act_f64 = (act_dtype == FP_F64);
    p.out = d_adjusted; p.type_major = type_major; p.slack = d_slack; p.intervened = d_intervened;
    CUDA_TRY(h, launch_safety_project(p, (cudaStream_t)stream));
    h->launches++;
    return FP_OK;
}

int fp_get_state(FpHandle* h, void* d_out, int dtype, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (!h->d_P) return fail(h, FP_ESTATE, "fp_get_state: call fp_load_profiles first");
    if (!d_out || (dtype != FP_F32 && dtype != FP_F64)) return fail(h, FP_EINVAL, "fp_get_state: bad arguments");
    ObsParams p; fill_obs_params(h, p, d_out, 0);
    CUDA_TRY(h, launch_state(p, dtype == FP_F64, h->grid_obs, (cudaStream_t)stream));
    h->launches++;
    return FP_OK;
}

int fp_state_ptrs(FpHandle* h, void** d_rec, void** d_voltage, void** d_setpoint, void** d_pflow,
                  void** d_qflow, void** d_isq) {
    if (!h) return FP_EINVAL;
    if (d_rec) *d_rec = h->d_rec;
    if (d_voltage) *d_voltage = h->d_V;
    if (d_setpoint) *d_setpoint = h->d_setp;
    if (d_pflow) *d_pflow = h->d_pfl;
    if (d_qflow) *d_qflow = h->d_qfl;
    if (d_isq) *d_isq = h->d_isq;
    return FP_OK;
}

// The fp64 observation-history ring (the source of truth of get_obs, quirk Q7): [N][history][FP_HIST_SLOT] doubles.
// With write != 0 the caller is about to overwrite it (state restore): the derived fp32 rings are rebuilt on next use.
int fp_history_ptr(FpHandle* h, void** d_hist, int64_t* doubles_per_env, int32_t write) {
    if (!h || !d_hist) return FP_EINVAL;
    if (write) h->hist_valid = true;               // about to be overwritten as a whole
    else { const int rc = hist_ensure(h, nullptr); if (rc != FP_OK) return rc; cudaStreamSynchronize(nullptr); }
    *d_hist = h->d_hist;
    if (doubles_per_env) *doubles_per_env = (int64_t)h->dc.history * FP_HIST_SLOT;
    if (write) { h->obsm_valid = false; h->obsr_valid = false; }
    return FP_OK;
}

int fp_set_keep_flows(FpHandle* h, int keep) {
    if (!h) return FP_EINVAL;
    if (keep && !h->d_pfl) {
        const size_t bytes = (size_t)h->n * h->dc.nl * 8;
        CUDA_TRY(h, cudaSetDevice(h->device));
        CUDA_TRY(h, cudaMalloc(&h->d_pfl, bytes));
        CUDA_TRY(h, cudaMalloc(&h->d_qfl, bytes));
        CUDA_TRY(h, cudaMalloc(&h->d_isq, bytes));
        CUDA_TRY(h, cudaMemset(h->d_pfl, 0, bytes));
        CUDA_TRY(h, cudaMemset(h->d_qfl, 0, bytes));
        CUDA_TRY(h, cudaMemset(h->d_isq, 0, bytes));
    }
    h->keep_flows = keep ? 1 : 0;
    return FP_OK;
}

int fp_power_flow(FpHandle* h, int64_t n, const double* d_p, const double* d_q, double* d_V, double* d_Pl,
                  double* d_Ql, double* d_Isq, int32_t* d_iters, uint8_t* d_fail, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    if (n < 1 || !d_p || !d_q || !d_V) return fail(h, FP_EINVAL, "fp_power_flow: bad arguments");
    PfParams p;
    p.topo = h->d_topo; p.n = n; p.nl = h->dc.nl; p.max_iter = h->dc.pf_max_iter; p.tol = h->dc.pf_tol; p.n32 = h->dc.pf_f32; p.pad_ = 0;
    p.p = d_p; p.q = d_q; p.V = d_V; p.Pl = d_Pl; p.Ql = d_Ql; p.Isq = d_Isq; p.iters = d_iters; p.fail = d_fail;
    if (h->variant == FP_VARIANT_THREAD) {
        PfParamsT pt;
        pt.p = p; pt.t = h->tt;
        const int64_t tiles = (n + 31) / 32;
        CUDA_TRY(h, launch_power_flow_t(h->shape, pt, (int)((tiles < h->grid_pf) ? tiles : h->grid_pf), (cudaStream_t)stream));
    } else {
        CUDA_TRY(h, launch_power_flow(p, grid_for(n, h->grid_pf), (cudaStream_t)stream));
    }
    h->launches++;
    return FP_OK;
}

int fp_stats_read(FpHandle* h, double* d_out, void* stream) {
    if (!h || !d_out) return FP_EINVAL;
    USE_DEVICE(h);
    CUDA_TRY(h, launch_stats_fold(h->d_stats_partial, h->stats_rows, d_out, (cudaStream_t)stream));
    h->launches++;
    return FP_OK;
}

int fp_stats_reset(FpHandle* h, void* stream) {
    if (!h) return FP_EINVAL;
    USE_DEVICE(h);
    CUDA_TRY(h, cudaMemsetAsync(h->d_stats_partial, 0, (size_t)h->stats_rows * FP_NSTATS * 8, (cudaStream_t)stream));
    return FP_OK;
}

int fp_inject_failure(FpHandle* h, const uint8_t* d_mask) {
    if (!h) return FP_EINVAL;
    h->d_inject = d_mask;
    return FP_OK;
}

}  // extern "C"

// accessors used by the predictor / replay translation unit
PredictorState* fp_internal_predictor(FpHandle* h) { return &h->pred; }
int fp_internal_fail(FpHandle* h, int code, const char* msg) { return fail(h, code, msg); }
void fp_internal_count_launch(FpHandle* h, int k) { h->launches += k; }
int fp_internal_device(FpHandle* h) { return h->device; }
