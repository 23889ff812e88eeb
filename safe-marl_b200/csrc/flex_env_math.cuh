// flex_env_math.cuh -- per-agent arithmetic of the env step, shared by the warp-per-env and the
// thread-per-env kernels (action scaling, ESS anti-simultaneity and energy clipping).
#pragma once
#include "flex_common.cuh"

// Setpoints of one agent (executed by the agent's lane).
struct Setpoint { double pred, ch, dis, qpv; };

// flexibility_provision_env.py:628-661 (no delta_t here -- quirk Q3)
__device__ __forceinline__ void ess_energy_clip(const DevCfg& c, double& ch, double& dis, double e_now) {
    ch = clipd(ch, 0.0, c.p_ch_max);
    dis = clipd(dis, 0.0, c.p_dis_max);
    double e_next = (e_now + c.eta_ch * ch) - c.inv_eta_dis * dis;
    if (e_next > c.e_max) {
        double excess = e_next - c.e_max;
        double t = excess / c.eta_ch;
        if (ch > t) {
            ch = ch - t;
        } else {
            dis = dis + (excess - ch * c.eta_ch) * c.eta_dis;
            ch = 0.0;
        }
    } else if (e_next < c.e_min) {
        double lack = c.e_min - e_next;
        double t = lack * c.eta_dis;
        if (dis > t) {
            dis = dis - t;
        } else {
            ch = ch + (lack - dis / c.eta_dis) / c.eta_ch;
            dis = 0.0;
        }
    }
    ch = clipd(ch, 0.0, c.p_ch_max);
    dis = clipd(dis, 0.0, c.p_dis_max);
}

// :262-293 (step) / :113-130 (reset): raw action -> applied setpoints, everything that does not
// need the building's load: `pred` returns the clipped reduction FRACTION (:676-677).
__device__ __forceinline__ Setpoint apply_actions_frac(const DevCfg& c, bool scale, double a0, double a1,
                                                       double a2, double a3, double ppv, double e_clip) {
    double pct, ch, dis, qpv;
    if (scale) {
        pct = c.mpr * a0;                                   // :278
        ch = c.p_ch_max * a1;                               // :279
        dis = c.p_dis_max * a2;                             // :280
        double lim = c.kappa * ppv;                         // :623
        double lo = -lim;
        qpv = clipd(lo + a3 * (lim - lo), lo, lim);         // :626
    } else {                                                // 'safemaddpg' branch :268-274
        pct = a0; ch = a1; dis = a2; qpv = a3;
    }
    pct = clipd(pct, 0.0, c.mpr);                           // :676-677
    if (ch > 0.0 && dis > 0.0) {                            // :663-674
        if (ch > dis) { ch = ch - dis; dis = 0.0; }
        else          { dis = dis - ch; ch = 0.0; }
    }
    ess_energy_clip(c, ch, dis, e_clip);                    // :289-290
    Setpoint s;
    s.pred = pct;
    s.ch = ch; s.dis = dis; s.qpv = qpv;
    return s;
}

__device__ __forceinline__ Setpoint apply_actions(const DevCfg& c, bool scale, double a0, double a1,
                                                  double a2, double a3, double pload, double ppv,
                                                  double e_clip) {
    Setpoint s = apply_actions_frac(c, scale, a0, a1, a2, a3, ppv, e_clip);
    s.pred = pload * s.pred;                                // :293
    return s;
}
