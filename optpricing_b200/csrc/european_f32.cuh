// european_f32.cuh -- the optional single-precision mode of the European kernels
// (BASELINE north_star: "fp64 (with an optional fp32 mode)").
//
// Same structure as european.cuh -- one thread per draw (a path and its antithetic
// partner), Philox4x32-10 in registers, nothing read from HBM, five doubles per
// block out -- but the step loop runs on the fp32 pipe:
//   * a Philox call yields TWO Box-Muller pairs (32-bit radius uniform, 24 of them
//     significant: |z| <= 5.9; 32-bit angle), i.e. two steps of a two-factor model
//     or four of a one-factor model;
//   * log / exp / sin / cos / sqrt / rsqrt are the SFU approximations
//     (__logf, __expf, __sincosf, sqrt.approx);
//   * the path state and the running sums of the deferred-drift form are floats;
//     the once-per-path finishing (drift, exp, payoff) and the payoff moments are
//     fp64, so the only single-precision error is in the accumulated increments
//     (~1e-6 of log S_T over 252 steps: a few 1e-5 relative on a price).
// The recurrences are the ones of models.cuh (same reference lines).  Not offered:
// fed draws (parity is an fp64 statement), the LSMC path store, Dupire.
#pragma once
#include "european.cuh"

namespace optmc {

struct StateF {
    float x, v, w;
};

struct ConstsF {
    float sqrt_dt, m11, m12, m21, m22;
    float one_minus_kappa_dt, kappa_theta_dt;
    float alpha, beta, neg_half_alpha2_dt, log_v0, rqc_dt;
    int beta_mode;
};

__device__ __forceinline__ float sqrt_approx(float a) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

// Two N(0,1) from two Philox words.
__device__ __forceinline__ void normal_pair_f32(uint32_t wr, uint32_t wa, float &z1, float &z2) {
    const float u = ((float)(wr >> 8) + 0.5f) * 0x1.0p-24f;               // exact, in (0, 1)
    const float rad = sqrt_approx(fmaxf(-2.0f * __logf(u), 0.0f));   // __logf is an approximation near 1
    float s, c;
    __sincosf((float)(int32_t)wa * (3.14159265358979f * 0x1.0p-31f), &s, &c);   // angle in [-pi, pi)
    z1 = rad * c;
    z2 = rad * s;
}

template <int KIND>
__device__ __forceinline__ void step_f32(StateF &s, float z1, float z2, const ConstsF &c) {
    if constexpr (KIND == OPTMC_HESTON || KIND == OPTMC_BATES) {
        // heston_kernel mc_kernels.py:62-77, deferred-drift form (models.cuh:step_heston)
        const float vp = fmaxf(s.v, 0.0f);
        const float sq = sqrt_approx(vp);
        s.x = fmaf(sq, z1, s.x);
        s.w += vp;
        s.v = fmaf(sq, z2, fmaf(vp, c.one_minus_kappa_dt, c.kappa_theta_dt));
    } else if constexpr (kind_sabr(KIND)) {
        // sabr_kernel mc_kernels.py:204-216, log-vol form (models.cuh:step_sabr<true>)
        const float s_pos = fmaxf(s.x, 1e-8f);
        float coef;
        switch (c.beta_mode) {
            case 1: coef = __expf(s.v) * s_pos; break;
            case 2: coef = __expf(s.v); break;
            default: coef = __expf(fmaf(c.beta, __logf(s_pos), s.v));
        }
        s.x += fmaf(coef, z1, c.rqc_dt * s_pos);
        s.v += fmaf(c.alpha, z2, c.neg_half_alpha2_dt);
    } else {
        s.x += z1;                       // one-factor, deferred drift: the Brownian sum only
    }
}

template <int KIND, bool ANTI>
__global__ void __launch_bounds__(EURO_THREADS, 4) k_european_f32(const EuroArgs A, const ConstsF F) {
    __shared__ double scratch[5 * 32];
    constexpr bool TWO = kind_two_factor(KIND);
    constexpr bool JUMPS = kind_jumps(KIND);
    constexpr bool STEP_JUMPS = KIND == OPTMC_SABR_JUMP;       // jumps act on S between steps
    const Consts &c = A.c;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        const uint64_t id = (uint64_t)(A.draw_offset + g);
        const uint32_t id_lo = (uint32_t)id, id_hi = (uint32_t)(id >> 32);
        StateF sa, sb;
        sa.x = sb.x = kind_sabr(KIND) ? (float)c.x0 : 0.0f;
        sa.v = sb.v = kind_sabr(KIND) ? F.log_v0 : (float)c.v0;
        sa.w = sb.w = 0.0f;
        auto advance = [&](uint32_t wr, uint32_t wa, int i) {
            float a, b;
            normal_pair_f32(wr, wa, a, b);
            if constexpr (TWO) {
                const float z1 = fmaf(F.m12, b, F.m11 * a), z2 = fmaf(F.m22, b, F.m21 * a);
                step_f32<KIND>(sa, z1, z2, F);
                if constexpr (ANTI) step_f32<KIND>(sb, -z1, -z2, F);
            } else {
                // the sqrt(dt) sigma scale is applied once per path by the fp64 finish
                step_f32<KIND>(sa, a, 0.0f, F);
                if constexpr (ANTI) step_f32<KIND>(sb, -a, 0.0f, F);
                if (i + 1 < A.n_steps) {
                    step_f32<KIND>(sa, b, 0.0f, F);
                    if constexpr (ANTI) step_f32<KIND>(sb, -b, 0.0f, F);
                }
            }
        };
        constexpr int SPP = TWO ? 1 : 2;                         // steps per pair
        if constexpr (STEP_JUMPS) {
            for (int i = 0; i < A.n_steps; ++i) {
                const uint4 w = philox4x32_10(make_uint4(id_lo, id_hi, (uint32_t)i, 0u), A.key);
                advance(w.x, w.y, i);
                if (__builtin_expect(poisson_maybe(w.w, A.pois_p0_bits[0]), 0)) {
                    double2 J = jumps_slow<KIND, ANTI>(w.w, A.pois_mu[0], A.pois_p0[0], A.key, id_lo,
                                                       id_hi, (uint32_t)i, c);
                    sa.x *= __expf((float)J.x);                  // mc_kernels.py:274
                    if constexpr (ANTI) sb.x *= __expf((float)J.y);
                }
            }
        } else {
            uint32_t call = 0;
            for (int i = 0; i < A.n_steps; i += 2 * SPP, ++call) {
                const uint4 w = philox4x32_10(make_uint4(id_lo, id_hi, call, 0x40u), A.key);
                advance(w.x, w.y, i);
                if (i + SPP < A.n_steps) advance(w.z, w.w, i + SPP);
            }
        }
        // ---- once per path, in fp64 -------------------------------------------------
        double xa, xb = 0.0;
        if constexpr (kind_sabr(KIND)) {
            xa = (double)sa.x;
            xb = (double)sb.x;
        } else {
            double ja = 0.0, jb = 0.0;
            if constexpr (JUMPS) {                               // whole-path jump draw (european.cuh)
                int n = 0;
                for (int ch = 0; ch < A.pois_chunks; ++ch) {
                    const uint4 e = philox4x32_10(
                        make_uint4(id_lo, id_hi, 0xFFFFFFFFu - (uint32_t)ch, 0xFEu), A.key);
                    if (poisson_maybe(e.x, A.pois_p0_path_bits))
                        n += poisson_slow(e.x, e.y, A.pois_mu_path, A.pois_p0_path);
                }
                if (n > 0) {
                    double2 J = jump_sum<KIND, ANTI>(n, A.key, id_lo, id_hi, 0xFFFFFFFFu, c);
                    ja = J.x;
                    jb = J.y;
                }
            }
            if constexpr (TWO) {         // log S_T = log S_0 + n (r-q-c) dt - dt/2 sum v+ + sum sqrt(v+) z1
                xa = c.x0 + (c.n_rqc_dt + fma((double)sa.w, c.neg_half_dt, (double)sa.x)) + ja;
                xb = c.x0 + (c.n_rqc_dt + fma((double)sb.w, c.neg_half_dt, (double)sb.x)) + jb;
            } else {                     // log S_T = log S_0 + n drift + sigma sqrt(dt) sum z + jumps
                const double sc = c.sigma * c.sqrt_dt;
                xa = (c.x0 + fma(sc, (double)sa.x, c.n_drift)) + ja;
                xb = (c.x0 + fma(sc, (double)sb.x, c.n_drift)) + jb;
            }
        }
        const double ST = kind_sabr(KIND) ? xa : exp(xa);
        double pay = fmax(A.is_call ? ST - A.strike : A.strike - ST, 0.0);
        double ctl = ST;
        if (A.terminal) A.terminal[g] = ST;
        if constexpr (ANTI) {
            const double STb = kind_sabr(KIND) ? xb : exp(xb);
            pay = 0.5 * (pay + fmax(A.is_call ? STb - A.strike : A.strike - STb, 0.0));
            ctl = 0.5 * (ST + STb);
            if (A.terminal) A.terminal[A.n_local + g] = STb;
        }
        acc[0] += pay;
        acc[1] = fma(pay, pay, acc[1]);
        acc[2] += ctl;
        acc[3] = fma(ctl, ctl, acc[3]);
        acc[4] = fma(pay, ctl, acc[4]);
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        double *p = A.partials + (size_t)blockIdx.x * NPART;
#pragma unroll
        for (int k = 0; k < 5; ++k) p[k] = acc[k];
    }
}

}  // namespace optmc
