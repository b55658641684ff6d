// european.cuh -- European Monte Carlo on the device RNG.
//
// One thread simulates one draw: a path and (if antithetic) its negated
// partner, sharing the Brownian draws as the reference's second pass does
// (monte_carlo.py:247-255).  State lives in registers for all n_steps; nothing
// is read from HBM, and only 5 doubles per block are written.
#pragma once
#include "models.cuh"
#include "reduce.cuh"

namespace optmc {

constexpr int EURO_THREADS = 256;
#ifndef OPTMC_EURO_MIN_BLOCKS
#define OPTMC_EURO_MIN_BLOCKS 3   // resident 256-thread blocks per SM the register budget is
                                  // sized for (85 regs): no constant reloads in the step loop
#endif
constexpr int NPART = 8;  // doubles per block partial (5 used)

struct EuroArgs {
    Consts c;
    PhiloxKey key;
    int64_t draw_offset;
    int64_t n_local;
    int n_steps;
    int is_call;
    double strike;
    // Poisson constants for a draw covering 1 step ([0]) or 2 steps ([1])
    double pois_mu[2], pois_p0[2];
    uint32_t pois_p0_bits[2];
    double *partials;   // [gridDim.x][NPART]
    double *terminal;   // optional [n_local * (anti ? 2 : 1)]
};

template <int KIND, int NIMPL>
__device__ __forceinline__ double jump_size(uint4 w, const Consts &c, RngView tab) {
    if constexpr (KIND == OPTMC_KOU) {
        // up w.p. p_up: +Exp(eta1), else -Exp(eta2)   (mc_kernels.py:314-317)
        double e;
        if constexpr (NIMPL == 0)
            e = -log(u01_open(w.x, w.y));
        else
            e = 0.5 * neg2_log_u(w.x, w.y, tab);
        return w.w < c.p_up_bits ? e * c.inv_eta1 : -e * c.inv_eta2;
    } else {
        double z1, z2;
        normal_pair<NIMPL>(w.x, w.y, w.z, z1, z2, tab);
        return fma(c.sigma_j, z1, c.mu_j);           // N(mu_j, sigma_j)  (mc_kernels.py:111)
    }
}

// Jumps of one Poisson draw (rare path; the caller has already seen w >= p0_bits).
// Counts are shared by the antithetic partner, sizes are fresh draws for it,
// as in the reference's second pass (monte_carlo.py:247-255).
template <int KIND, bool ANTI, int NIMPL, bool DEFER>
__device__ __noinline__ void jumps_slow(State &sa, State &sb, uint32_t w, double mu, double p0,
                                        const PhiloxKey &key, uint64_t id, uint32_t step,
                                        const Consts &c, RngView tab) {
    uint4 e = philox_draw(key, id, step, 0xFFu);
    int n = poisson_slow(w, e.x, mu, p0);
    for (int j = 0; j < n; ++j) {
        apply_jump<KIND, DEFER>(sa, jump_size<KIND, NIMPL>(
                                        philox_draw(key, id, step, 1u | ((uint32_t)j << 8)), c, tab));
        if constexpr (ANTI)
            apply_jump<KIND, DEFER>(sb, jump_size<KIND, NIMPL>(
                                            philox_draw(key, id, step, 2u | ((uint32_t)j << 8)), c,
                                            tab));
    }
}

// Simulate one draw over all steps.  SINK(step_index, sa, sb) is called after
// every completed step when PER_STEP (LSMC path store), never otherwise.
template <int KIND, bool ANTI, int NIMPL, bool PER_STEP, class Sink>
__device__ __forceinline__ void simulate_draw(const EuroArgs &A, uint64_t id, State &sa, State &sb,
                                              RngView tab, Sink &&sink) {
    const Consts &c = A.c;
    constexpr bool TWO = kind_two_factor(KIND);
    constexpr bool JUMPS = kind_jumps(KIND);
    // terminal-only kernels of the log models defer the deterministic drift
    constexpr bool DEFER = !PER_STEP && !kind_sabr(KIND);
    init_state<KIND, DEFER>(sa, c);
    init_state<KIND, DEFER>(sb, c);
    // A Box-Muller pair feeds one step of a two-factor model, or two
    // consecutive steps of a one-factor model (terminal-only kernels).
    constexpr int SPC = (TWO || PER_STEP) ? 1 : 2;
    const uint32_t id_lo = keep_u32((uint32_t)id), id_hi = keep_u32((uint32_t)(id >> 32));
    for (int i = 0; i < A.n_steps; i += SPC) {
        uint4 w = philox4x32_10(make_uint4(id_lo, id_hi, (uint32_t)i, 0u), A.key);
        double a, b;
        normal_pair<NIMPL>(w.x, w.y, w.z, a, b, tab);
        if constexpr (TWO) {
            double z1 = fma(c.m12, b, c.m11 * a);
            double cz = fma(c.m22, b, c.m21 * a);
            step_diffusion<KIND, DEFER>(sa, z1, cz, c);
            if constexpr (ANTI) step_diffusion<KIND, DEFER>(sb, -z1, -cz, c);
        } else {
            double dw = c.sqrt_dt * a;
            step_diffusion<KIND, DEFER>(sa, dw, 0.0, c);
            if constexpr (ANTI) step_diffusion<KIND, DEFER>(sb, -dw, 0.0, c);
            if constexpr (SPC == 2) {
                if (i + 1 < A.n_steps) {
                    double dw2 = c.sqrt_dt * b;
                    step_diffusion<KIND, DEFER>(sa, dw2, 0.0, c);
                    if constexpr (ANTI) step_diffusion<KIND, DEFER>(sb, -dw2, 0.0, c);
                }
            }
        }
        if constexpr (JUMPS) {
            // One Poisson draw for the steps this iteration covered; for the
            // additive-in-log one-factor models the split between the two
            // steps does not affect S_T.
            const int two = (SPC == 2 && i + 1 < A.n_steps) ? 1 : 0;
            if (poisson_maybe(w.w, A.pois_p0_bits[two]))
                jumps_slow<KIND, ANTI, NIMPL, DEFER>(sa, sb, w.w, A.pois_mu[two], A.pois_p0[two],
                                                     A.key, id, (uint32_t)i, c, tab);
        }
        if constexpr (PER_STEP) sink(i, sa, sb);
    }
    finish_deferred<KIND, DEFER>(sa, c);
    if constexpr (ANTI) finish_deferred<KIND, DEFER>(sb, c);
}

struct NoSink {
    __device__ __forceinline__ void operator()(int, const State &, const State &) const {}
};

template <int KIND, bool ANTI, int NIMPL>
__global__ void __launch_bounds__(EURO_THREADS, OPTMC_EURO_MIN_BLOCKS) k_european(const EuroArgs A) {
    __shared__ RngTables tab;
    RngView tv{0u};
    __shared__ double scratch[5 * 32];
    if constexpr (NIMPL != 0) {
        tv = load_rng_tables(&tab);
        __syncthreads();
    }
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        State sa, sb;
        simulate_draw<KIND, ANTI, NIMPL, false>(A, (uint64_t)(A.draw_offset + g), sa, sb, tv,
                                                NoSink{});
        double ST = terminal_spot(KIND, sa);
        double pay = A.is_call ? ST - A.strike : A.strike - ST;   // monte_carlo.py:110-114
        pay = fmax(pay, 0.0);
        double ctl = ST;
        if (A.terminal) A.terminal[g] = ST;
        if constexpr (ANTI) {
            double STb = terminal_spot(KIND, sb);
            double payb = A.is_call ? STb - A.strike : A.strike - STb;
            pay = 0.5 * (pay + fmax(payb, 0.0));
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

// Fixed-order sum of the per-block partials -> out[0..5] (n, sums).
__global__ void __launch_bounds__(256) k_finish_moments(const double *partials, int n_blocks,
                                                        double n_samples, double *out) {
    __shared__ double scratch[5 * 32];
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[k] += partials[(size_t)b * NPART + k];
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        out[0] = n_samples;
#pragma unroll
        for (int k = 0; k < 5; ++k) out[1 + k] = acc[k];
    }
}

// Strike sweep over stored terminal spots: block (x = path chunk, y = contract).
constexpr int SWEEP_MAX = 64;
struct SweepArgs {
    const double *terminal;
    int64_t n_local;
    int antithetic;
    int n_contracts;          // <= SWEEP_MAX per launch; passed by value, no staging copy
    double strikes[SWEEP_MAX];
    unsigned char is_call[SWEEP_MAX];
    double *partials;         // [n_contracts][gridDim.x][NPART]
};

__global__ void __launch_bounds__(256) k_payoff_sweep(const SweepArgs A) {
    __shared__ double scratch[5 * 32];
    const int cidx = blockIdx.y;
    const double K = A.strikes[cidx];
    const bool call = A.is_call[cidx] != 0;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        double ST = A.terminal[g];
        double pay = fmax(call ? ST - K : K - ST, 0.0);
        double ctl = ST;
        if (A.antithetic) {
            double STb = A.terminal[A.n_local + g];
            pay = 0.5 * (pay + fmax(call ? STb - K : K - STb, 0.0));
            ctl = 0.5 * (ST + STb);
        }
        acc[0] += pay;
        acc[1] = fma(pay, pay, acc[1]);
        acc[2] += ctl;
        acc[3] = fma(ctl, ctl, acc[3]);
        acc[4] = fma(pay, ctl, acc[4]);
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        double *p = A.partials + ((size_t)cidx * gridDim.x + blockIdx.x) * NPART;
#pragma unroll
        for (int k = 0; k < 5; ++k) p[k] = acc[k];
    }
}

__global__ void __launch_bounds__(256) k_finish_sweep(const double *partials, int n_blocks,
                                                      double n_samples, double *out) {
    // one block per contract
    __shared__ double scratch[5 * 32];
    const double *mine = partials + (size_t)blockIdx.x * n_blocks * NPART;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[k] += mine[(size_t)b * NPART + k];
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        double *o = out + (size_t)blockIdx.x * OPTMC_NMOM;
        o[0] = n_samples;
#pragma unroll
        for (int k = 0; k < 5; ++k) o[1 + k] = acc[k];
    }
}

}  // namespace optmc
