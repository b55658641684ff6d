// models.cuh -- per-step recurrences of the seven SDE models.
//
// One set of step functions serves both the fed-draw kernels (parity with the
// numba originals to <= 1e-12) and the Philox kernels (throughput): what is
// parity-checked is what is benchmarked.  Reference lines per function below
// are relative to /root/reference/src/optpricing/techniques/kernels/.
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/optmc.h"
#include "philox.cuh"

namespace optmc {

// Everything a kernel needs, precomputed once on the host in the reference's
// own expression order (so drift/compensator constants match its doubles).
struct Consts {
    int kind;
    int beta_mode;            // SABR: 0 general pow, 1 beta==1, 2 beta==0, 3 beta==0.5
    double x0;                // log_s0, or s0 for SABR kinds        (monte_carlo.py:217-220)
    double v0;
    double dt, sqrt_dt, inv_sqrt_dt;
    double r, q;
    // one-factor log models
    double sigma;
    double drift;             // (r - q - 0.5 sigma^2 - compensator) dt   (mc_kernels.py:29,104,304)
    // Heston / Bates
    double rqc;               // r - q - compensator
    double rqc_dt;            // rqc * dt
    double neg_half_dt;       // -0.5 dt
    double kappa_theta_dt;    // kappa theta dt
    double one_minus_kappa_dt;
    double n_rqc_dt;          // n_steps * rqc_dt   (deferred-drift kernels)
    double n_drift;           // n_steps * drift
    double xi;                // vol_of_vol
    // SABR
    double alpha, beta;
    double neg_half_alpha2_dt;
    double log_v0;            // ln v0 (SABR kinds, log-vol form; -600 stands for v0 <= 0)
    double pw[4];             // (1 + rho)^beta = 1 + r' (pw0 + r' (pw1 + r' (pw2 + r' pw3))), r' = -2 rho
    // correlation of the kernel itself (mc_kernels.py:59,67)
    double rho, rho_bar;
    // Philox mode: (z1, cz) = M (a, b) for iid N(0,1) a, b; sqrt(dt) and, for
    // Heston/Bates, vol_of_vol are folded in (see make_consts).
    double m11, m12, m21, m22;
    // jumps
    double lam_dt;            // lambda dt (per Poisson draw; x2 when two steps share a draw)
    double mu_j, sigma_j;
    double p_up, inv_eta1, inv_eta2;
    uint32_t p_up_bits;
    // Dupire: sigma(t_i, S) table, rows = steps, columns = nodes equally spaced in log S
    const double *lv_table;
    int lv_n_s;
    double lv_lo, lv_inv_h, lv_max_u;
    double rq;                // r - q
    double scen[3];           // spot factors of the scenario kernels (k_european_scen)
};

__host__ __device__ constexpr bool kind_two_factor(int k) {
    return k == OPTMC_HESTON || k == OPTMC_BATES || k == OPTMC_SABR || k == OPTMC_SABR_JUMP;
}
__host__ __device__ constexpr bool kind_jumps(int k) {
    return k == OPTMC_MERTON || k == OPTMC_BATES || k == OPTMC_SABR_JUMP || k == OPTMC_KOU;
}
__host__ __device__ constexpr bool kind_sabr(int k) {
    return k == OPTMC_SABR || k == OPTMC_SABR_JUMP;
}

// N spot scenarios of one path on ONE set of draws (SABR kinds: the vol path does not depend
// on S, so bumped-spot paths -- the up / base / down re-pricings of GreekMixin.delta and
// .gamma, greek_mixin.py:59-122 -- share everything but the spot recurrence)
template <int N>
struct Scen {
    double a[N];
    __device__ __forceinline__ Scen &operator*=(double f) {
#pragma unroll
        for (int k = 0; k < N; ++k) a[k] *= f;
        return *this;
    }
    __device__ __forceinline__ Scen &operator+=(double f) {
#pragma unroll
        for (int k = 0; k < N; ++k) a[k] += f;
        return *this;
    }
};
template <class X> struct scen_count { static constexpr int value = 1; };
template <int N> struct scen_count<Scen<N>> { static constexpr int value = N; };

template <class X>
struct StateT {
    X x;       // log S (S for SABR kinds); with DEFER: the stochastic part only
    double v;  // variance (Heston/Bates, stored before the floor) or vol (SABR)
    double w;  // DEFER only: running sum of v_pos (Heston/Bates); sum of log-jumps (Merton/Kou)
    double j;  // Bates terminal kernels: the path's summed log-jumps (written once, after the loop)
};
using State = StateT<double>;

// DEFER (terminal-value kernels of the log models): the deterministic drift is
// added once after the last step instead of every step,
//   one-factor:  log S_T = log S_0 + n drift + sigma sum(dw)            [+ jumps]
//   Heston:      log S_T = log S_0 + n (r-q-c) dt - dt/2 sum(v+) + sum(sqrt(v+) z1)
// which is the reference recurrence re-associated (agreement ~ n * 1e-16).  The
// path kernels need log S at every step and use the literal form.

// bsm_kernel mc_kernels.py:30-32; merton :105-106; kou :306-307 (diffusion part)
template <bool DEFER>
__device__ __forceinline__ void step_one_factor(State &s, double dw, const Consts &c) {
    if constexpr (DEFER)
        s.x += dw;
    else
        s.x += fma(c.sigma, dw, c.drift);
}

// heston_kernel mc_kernels.py:62-77, bates_kernel :151-164.
// z1 = dw1; czx = vol_of_vol * (rho*dw1 + rho_bar*dw2).
// s.v holds the variance BEFORE the floor of :76-77; flooring on read is the same
// thing for every step (v0 >= 0 is enforced on the host, see make_consts).
template <bool DEFER>
__device__ __forceinline__ void step_heston(State &s, double z1, double czx, const Consts &c) {
    double vp = relu64(s.v);                       // v_pos = max(v, 0)
    double sq = sqrt_nonneg(vp);                   // v_sqrt, < 1 ulp: the fed-draw parity step
    if constexpr (DEFER) {
        s.x = fma(sq, z1, s.x);
        s.w += vp;
    } else {
        s.x += fma(sq, z1, fma(vp, c.neg_half_dt, c.rqc_dt));
    }
    s.v = fma(sq, czx, fma(vp, c.one_minus_kappa_dt, c.kappa_theta_dt));
}

// The same step on the device RNG, from the parts of a Box-Muller pair (philox.cuh:
// z1 = sqrt(L) t1, czx = sqrt(L) t2): the step uses the normals only as sqrt(v+) z, and
// sqrt(v+) sqrt(L) = sqrt(v+ L) is ONE root per path instead of one per path plus the
// pair's radius.  The partner path passes (-t1, -t2).
template <bool DEFER>
__device__ __forceinline__ void step_heston_parts(State &s, double L, double t1, double t2,
                                                  const Consts &c) {
    double vp = relu64(s.v);
    // v_sqrt * radius, to the accuracy the Box-Muller radius always had here (one Newton
    // correction, rel. 1e-12: a perturbation of a N(0,1) draw no price can see); the
    // fed-draw parity kernels take step_heston with its < 1 ulp root.
    // The 1e-290 keeps the seed finite where v+ L is 0 (v+ = 0 is common when Feller's
    // condition fails): the root comes out at 1e-145 instead of 0, for one FMA where a
    // product plus an integer clamp stood.
    double rq = sqrt_pos_fast(fma(vp, L, C_TINY));
    if constexpr (DEFER) {
        s.x = fma(rq, t1, s.x);
        s.w += vp;
    } else {
        s.x += fma(rq, t1, fma(vp, c.neg_half_dt, c.rqc_dt));
    }
    s.v = fma(rq, t2, fma(vp, c.one_minus_kappa_dt, c.kappa_theta_dt));
}

template <int KIND, bool DEFER, class S>
__device__ __forceinline__ void finish_deferred(S &s, const Consts &c) {
    if constexpr (DEFER) {
        if constexpr (KIND == OPTMC_HESTON || KIND == OPTMC_BATES)
            s.x = c.x0 + (c.n_rqc_dt + fma(s.w, c.neg_half_dt, s.x));
        else {
            s.v = s.x;                                           // keep the Brownian sum W (see k_sigma_sweep)
            s.x = (c.x0 + fma(c.sigma, s.x, c.n_drift)) + s.w;   // s.w: sum of log-jumps
        }
    }
}

// dupire_kernel mc_kernels.py:351-357: local_vol = vol_surface(i dt, exp(log_s));
// log_s += (r - q - local_vol^2 / 2) dt + local_vol dw.  The surface is the table
// described in optmc.h, linear in log S between nodes, flat outside.
__device__ __forceinline__ void step_dupire(State &s, double dw, int i, const Consts &c) {
    const double *row = c.lv_table + (size_t)i * c.lv_n_s;
    double u = fmin(fmax((s.x - c.lv_lo) * c.lv_inv_h, 0.0), c.lv_max_u);
    int j = (int)u;                                   // lv_max_u < n_s - 1: j + 1 is a node
    double f = u - (double)j;
    double a = __ldg(row + j), b = __ldg(row + j + 1);
    double vol = fma(f, b - a, a);
    s.x += fma(vol, dw, fma(-0.5 * vol, vol, c.rq) * c.dt);
}

// s_pos ** beta (mc_kernels.py:211): exp(beta ln s) on the table-driven helpers,
// with the exact special cases beta in {1, 0, 1/2}.
template <class V>
__device__ __forceinline__ double sabr_pow(double s_pos, const Consts &c, V t) {
    switch (c.beta_mode) {
        case 1: return s_pos;
        case 2: return 1.0;
        case 3: return sqrt_nonneg(s_pos);
        default: return fast_exp(c.beta * fast_log(s_pos, t), t);
    }
}

// sabr_kernel mc_kernels.py:204-216, sabr_jump_kernel :251-258.
// z1 = dw1; cz = rho*dw1 + rho_bar*dw2.
// LOGV (device-RNG kernels): the vol recurrence v *= exp(a + alpha cz) is carried as
// ln v += a + alpha cz, and the diffusion coefficient v s^beta = exp(ln v + beta ln s)
// costs ONE exponential per step instead of two (the same reals; rounding differs by
// ~1e-16 per step, so the fed-draw parity kernels keep the literal form).
// max(a, 1e-8) and max(a, -650) in ONE integer-pipe instruction each, like relu64
// (fp64 max is DSETP + two selects here).  Doubles order like their high words:
// signed for the positive floor, unsigned -- where every negative sorts above every
// positive, by magnitude -- for the negative one.  The result can differ from the
// exact max by the low word of `a` (relative 2^-20) only when `a` is within one
// high-word step of the bound: the floor 1e-8 and the underflow guard -650 are
// arbitrary at that level.
__device__ __forceinline__ double floor_1e8(double a) {
    return __hiloint2double(max(__double2hiint(a), 0x3E45798E), __double2loint(a));
}
__device__ __forceinline__ double floor_neg650(double a) {
    return __hiloint2double((int)min((unsigned)__double2hiint(a), 0xC0845000u), __double2loint(a));
}

// LOGV: 0 = literal form; 1 = log-vol form, beta_mode read at run time; 2 = log-vol form
// for a beta that is neither 0 nor 1, chosen by the host -- the step loop then carries one
// code path instead of three and no switch.
// the spot recurrence of one SABR step, given the (log-)vol BEFORE its own update
template <int LOGV, class V>
__device__ __forceinline__ void sabr_spot_update(double &x, double v, double z1, const Consts &c,
                                                 V t) {
    if constexpr (LOGV != 0) {
        const double s_pos = floor_1e8(x);
        double coef;
        if constexpr (LOGV == 2) {
            coef = fast_exp(floor_neg650(fma(c.beta, fast_log(s_pos, t), v)), t);
        } else {
            switch (c.beta_mode) {
                case 1: coef = fast_exp(floor_neg650(v), t) * s_pos; break;
                case 2: coef = fast_exp(floor_neg650(v), t); break;
                default: coef = fast_exp(floor_neg650(fma(c.beta, fast_log(s_pos, t), v)), t);
            }
        }
        x += fma(coef, z1, c.rqc_dt * s_pos);
    } else {
        double s_pos = fmax(x, 1e-8);
        double v_pos = relu64(v);
        x += fma(v_pos * sabr_pow(s_pos, c, t), z1, c.rqc_dt * s_pos);
    }
}

// s^beta for a positive normal s from the tables of the scenario kernel: with s = 2^e m,
// m = c_i (1 + rho) for the log table's node c_i (|rho| <= 2^-10),
//   s^beta = 2^(e beta) c_i^beta (1 + rho)^beta,
// two 8-byte lookups (beta is known at launch; the host expands the binomial series in
// r' = -2 rho to degree 4: the next term is below 1e-17) and 7 fp64-pipe instructions, where
// exp(beta ln s) costs 17.  Exponents outside -64 .. 63 are clamped (s is floored at 1e-8).
template <class V>
__device__ __forceinline__ double pow_beta(double s_pos, const Consts &c, V t) {
    const int hi = __double2hiint(s_pos);
    const double m = __hiloint2double(mantissa_hi(hi, t.k3ff), __double2loint(s_pos));
    const double2 tb = lds_f64x2(t.log_at((uint32_t)hi));
    const double r = fma(m, tb.x, 2.0);
    double p = fma(r, c.pw[3], c.pw[2]);
    p = fma(p, r, c.pw[1]);
    p = fma(p, r, c.pw[0]);
    p = fma(p, r, 1.0);
    const double tc = lds_f64(t.pw_c + (((uint32_t)hi >> 7) & (511u << 4)));
    const int e = min(max((hi >> 20) - (1023 - POW_E_BIAS), 0), POW_E_N - 1);
    const double te = lds_f64(t.pw_e + (uint32_t)e * 8u);
    return (te * tc) * p;
}

template <int LOGV, class S, class V>
__device__ __forceinline__ void step_sabr(S &s, double z1, double cz, const Consts &c, V t) {
    constexpr int NS = scen_count<decltype(s.x)>::value;
    if constexpr (NS == 1) {
        sabr_spot_update<LOGV>(s.x, s.v, z1, c, t);
    } else if constexpr (LOGV == 3) {
        // spot scenarios, beta neither 0 nor 1: ONE exponential for the vol state they share,
        // s^beta per spot from the beta-specific tables (pow_beta)
        const double vz = fast_exp(floor_neg650(s.v), t) * z1;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const double s_pos = floor_1e8(s.x.a[k]);
            s.x.a[k] += fma(vz, pow_beta(s_pos, c, t), c.rqc_dt * s_pos);
        }
    } else {
#pragma unroll
        for (int k = 0; k < NS; ++k) sabr_spot_update<LOGV>(s.x.a[k], s.v, z1, c, t);
    }
    if constexpr (LOGV != 0)
        s.v += fma(c.alpha, cz, c.neg_half_alpha2_dt);
    else
        s.v = s.v * fast_exp(fma(c.alpha, cz, c.neg_half_alpha2_dt), t);
}

// Fed mode: the kernel's own correlation step applied to the given increments.
__device__ __forceinline__ void mix_fed(int kind, double dw1, double dw2, const Consts &c,
                                        double &z1, double &cz) {
    z1 = dw1;
    cz = fma(c.rho, dw1, c.rho_bar * dw2);          // mc_kernels.py:67
    if (kind == OPTMC_HESTON || kind == OPTMC_BATES) cz *= c.xi;
}

template <int KIND, bool DEFER = false, int LOGV = 0, class S = State, class V = RngView>
__device__ __forceinline__ void step_diffusion(S &s, double z1, double cz, const Consts &c,
                                               V t, int i) {
    if constexpr (KIND == OPTMC_DUPIRE)
        step_dupire(s, z1, i, c);
    else if constexpr (KIND == OPTMC_HESTON || KIND == OPTMC_BATES)
        step_heston<DEFER>(s, z1, cz, c);
    else if constexpr (kind_sabr(KIND))
        step_sabr<LOGV>(s, z1, cz, c, t);
    else
        step_one_factor<DEFER>(s, z1, c);
}

// Initial state.  With DEFER the log-spot accumulator starts at 0 and log S_0 is
// added by finish_deferred; jumps (additive in log space) accumulate in x too.
template <int KIND, bool DEFER, int LOGV = 0, class S = State>
__device__ __forceinline__ void init_state(S &s, const Consts &c) {
    constexpr int NS = scen_count<decltype(s.x)>::value;
    if constexpr (NS == 1) {
        s.x = (DEFER && !kind_sabr(KIND)) ? 0.0 : c.x0;
    } else {
        static_assert(kind_sabr(KIND), "spot scenarios: SABR kinds");
#pragma unroll
        for (int k = 0; k < NS; ++k) s.x.a[k] = c.x0 * c.scen[k];
    }
    s.v = (LOGV != 0 && kind_sabr(KIND)) ? c.log_v0 : c.v0;
    s.w = 0.0;
}

// Apply one jump of size J (Merton/Bates/Kou: additive in log space,
// mc_kernels.py:115,174,323; SABR-jump: multiplicative on S, :274).
template <int KIND, bool DEFER = false, class S = State>
__device__ __forceinline__ void apply_jump(S &s, double J) {
    if constexpr (KIND == OPTMC_SABR_JUMP)
        s.x *= exp(J);     // rare path: CUDA math library
    else if constexpr (DEFER && (KIND == OPTMC_MERTON || KIND == OPTMC_KOU))
        s.w += J;      // s.x holds the unscaled Brownian sum here
    else
        s.x += J;
}

__device__ __forceinline__ double terminal_spot(int kind, const State &s) {
    return kind_sabr(kind) ? s.x : exp(s.x);        // monte_carlo.py:257
}

// Spot after a step, for the path kernels (path_kernels.py:34: paths[:, i+1] = exp(log_s))
template <class V>
__device__ __forceinline__ double step_spot(int kind, const State &s, V t) {
    return kind_sabr(kind) ? s.x : fast_exp(s.x, t);
}

}  // namespace optmc
