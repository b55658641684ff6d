// european.cuh -- European Monte Carlo on the device RNG.
//
// One thread simulates one draw: a path and (if antithetic) its negated
// partner, sharing the Brownian draws as the reference's second pass does
// (monte_carlo.py:247-255).  State lives in registers for all n_steps; nothing
// is read from HBM, and only 5 doubles per block are written.
#pragma once
#include <type_traits>

#include "models.cuh"
#include "reduce.cuh"

namespace optmc {

#define OPTMC_PRAGMA_(x) _Pragma(#x)
#define OPTMC_PRAGMA(x) OPTMC_PRAGMA_(x)
#ifdef OPTMC_UNROLL
#define OPTMC_UNROLL_STEPS OPTMC_PRAGMA(unroll OPTMC_UNROLL)
#else
#define OPTMC_UNROLL_STEPS
#endif

constexpr int EURO_THREADS = 256;
constexpr int SCHED_SLOTS = 8;    // scheduled jump steps a path keeps in registers
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
    double pois_mu_path, pois_p0_path;   // lambda T / pois_chunks and exp(-that): whole-path count
    uint32_t pois_p0_path_bits;          // floor(exp(-lambda T / pois_chunks) 2^32)
    int pois_chunks;                     // 1 unless lambda T > 16 (terminal kernels only)
    double *partials;   // [gridDim.x][NPART]
    double *terminal;   // optional [n_local * (anti ? 2 : 1)]
    double *aux;        // optional, one-factor log models: [W | J | J partner], n_local each
};

// N(0,1) draws from one Philox call for the jump sizes: x always; y is independent of x
// only for the one-factor kinds (the two-factor kinds' second output is the mixed one).
// NIMPL 1: the block's tables -- the first output of the step loop's own normal generator
// is sqrt(dt) N(0,1) for every kind (the two-factor mixing row (m11, m12) has norm sqrt(dt)
// in both correlation modes).
template <int KIND, int NIMPL, class V>
__device__ __forceinline__ double2 unit_normals(uint4 w, const Consts &c, V tab) {
    double z1, z2;
    if constexpr (NIMPL == 0) {
        normal_pair_lib2(w.x, w.y, z1, z2);
    } else {
        normal_pair_fast<kind_two_factor(KIND)>(w.x, w.y, z1, z2, tab);
        z1 *= c.inv_sqrt_dt;
        z2 *= c.inv_sqrt_dt;
    }
    return make_double2(z1, z2);
}

// One jump size.  NIMPL 0: CUDA math library.  NIMPL 1: the block's tables -- the first
// output of the step loop's own normal generator is sqrt(dt) N(0,1) for every kind (the
// two-factor mixing row (m11, m12) has norm sqrt(dt) in both correlation modes), and
// -ln u comes from the log table.  With lambda T of order one most warps reach this
// epilogue, so it is not as rare as a single step's jump.
template <int KIND, int NIMPL, class V>
__device__ __forceinline__ double jump_size(uint4 w, const Consts &c, V tab) {
    if constexpr (KIND == OPTMC_KOU) {
        // up w.p. p_up: +Exp(eta1), else -Exp(eta2)   (mc_kernels.py:314-317)
        double e;
        if constexpr (NIMPL == 0) e = -log(u01_open(w.x, w.y));
        else e = 0.5 * neg2_log_u(w.x | 1u, w.y, tab);
        return w.w < c.p_up_bits ? e * c.inv_eta1 : -e * c.inv_eta2;
    } else {
        return fma(c.sigma_j, unit_normals<KIND, NIMPL>(w, c, tab).x, c.mu_j);   // N(mu_j, sigma_j)  (mc_kernels.py:111)
    }
}

// Jumps of one Poisson draw (rare path; the caller has already seen w >= p0_bits).
// Counts are shared by the antithetic partner, sizes are fresh draws for it, as in
// the reference's second pass (monte_carlo.py:247-255).  Returns the
// summed log-jumps (x: the draw, y: its partner) so that no path state has to be
// passed by reference.  Inlined on purpose: an out-of-line call would force the
// kernel parameters (constants, Philox key) to be copied to local memory so that
// their address can be taken.
template <int KIND, bool ANTI, int NIMPL = 0, class V = RngView>
__device__ __forceinline__ double2 jump_sum(int n, const PhiloxKey &key, uint32_t id_lo,
                                            uint32_t id_hi, uint32_t step, const Consts &c,
                                            V tab = V{}) {
    double2 acc = make_double2(0.0, 0.0);
    if constexpr (KIND != OPTMC_KOU) {
        // n iid N(mu_J, sigma_J^2) sizes sum to N(n mu_J, n sigma_J^2): ONE normal per path and
        // draw whatever n is -- the same law as the reference's n draws (mc_kernels.py:111-115);
        // the partner's is independent of it (the second output of the pair for the one-factor
        // kinds, a second Philox call for the two-factor ones)
        if (n > 0) {
            const double nd = (double)n, sd = c.sigma_j * sqrt(nd);
            const double2 z = unit_normals<KIND, NIMPL>(
                philox4x32_10(make_uint4(id_lo, id_hi, step, 1u), key), c, tab);
            acc.x = fma(sd, z.x, nd * c.mu_j);
            if constexpr (ANTI) {
                double zb = z.y;
                if constexpr (kind_two_factor(KIND) || NIMPL == 0)
                    zb = unit_normals<KIND, NIMPL>(
                             philox4x32_10(make_uint4(id_lo, id_hi, step, 2u), key), c, tab).x;
                acc.y = fma(sd, zb, nd * c.mu_j);
            }
        }
    } else {
        for (int j = 0; j < n; ++j) {
            acc.x += jump_size<KIND, NIMPL>(
                philox4x32_10(make_uint4(id_lo, id_hi, step, 1u | ((uint32_t)j << 8)), key), c, tab);
            if constexpr (ANTI)
                acc.y += jump_size<KIND, NIMPL>(
                    philox4x32_10(make_uint4(id_lo, id_hi, step, 2u | ((uint32_t)j << 8)), key), c,
                    tab);
        }
    }
    return acc;
}

template <int KIND, bool ANTI, int NIMPL = 0, class V = RngView>
__device__ __forceinline__ double2 jumps_slow(uint32_t w, double mu, double p0, const PhiloxKey &key,
                                           uint32_t id_lo, uint32_t id_hi, uint32_t step,
                                           const Consts &c, V tab = V{}) {
    uint4 e = philox4x32_10(make_uint4(id_lo, id_hi, step, 0xFFu), key);
    return jump_sum<KIND, ANTI, NIMPL>(poisson_slow(w, e.x, mu, p0), key, id_lo, id_hi, step, c, tab);
}

// Simulate one draw over all steps.  SINK(step_index, sa, sb) is called after
// every completed step when PER_STEP (LSMC path store), never otherwise.
//
// SCHED (kernels whose jumps must act inside the step loop: the path kernels of the jump
// models and European SABR-jump): instead of one Poisson(lambda dt) draw per step, the path draws
// N ~ Poisson(lambda T) once and the N step indices floor(U n_steps) -- the arrival
// times of a Poisson process given their number are iid uniform, so the per-step counts
// have exactly the law of the reference's independent Poisson(lambda dt) draws
// (monte_carlo.py:238-240).  The step loop then needs no Poisson word (the packed
// Philox layout applies, one-factor path kernels use both normals of a pair).  Up to eight
// indices are kept sorted in four registers, and the path's NEXT jump -- its step and what it
// does to the draw and to its partner -- is kept ready and applied branch-free at every step;
// a path with more than eight jumps (P = 2e-4 at lambda T = 2) recomputes its indices at every
// step.  The host selects SCHED while lambda T <= 2.
// (Until late in round 2 the terminal-only SABR-jump kernel ran groups of four steps, jump-free
// groups through a copy of the body without jump code: the same speed for one spot, 15 % slower
// with three spot scenarios side by side -- profiles/r02_experiments.md.)
template <int KIND, bool ANTI, int NIMPL, bool PER_STEP, bool SCHED, bool BGEN = false, class S,
          class V, class Sink>
__device__ __forceinline__ void simulate_draw(const EuroArgs &A, uint64_t id, S &sa, S &sb,
                                              V tab, Sink &&sink) {
    const Consts &c = A.c;
    constexpr bool TWO = kind_two_factor(KIND);
    constexpr bool JUMPS = kind_jumps(KIND);
    // Terminal-value kernels of the models whose jumps are additive in log space
    // and independent of the state (Merton, Bates, Kou): the sum of the per-step
    // Poisson(lambda dt) counts is Poisson(lambda T) and the jump sizes are iid, so
    // the whole jump contribution to log S_T is drawn once after the last step --
    // the same law as the reference's per-step draws, without a divergent branch
    // inside the step loop.  SABR-jump (jumps act on S between Euler steps) and the
    // path kernels keep jumps inside the loop.
    constexpr bool JUMPS_AT_END = JUMPS && !PER_STEP && KIND != OPTMC_SABR_JUMP;
    constexpr bool STEP_JUMPS = JUMPS && !JUMPS_AT_END;
    constexpr bool SCHEDULED = STEP_JUMPS && SCHED;       // pre-drawn jump schedule
    constexpr bool LEGACY = STEP_JUMPS && !SCHED;         // one Poisson word per step
    // terminal-only kernels of the log models defer the deterministic drift
    constexpr bool DEFER = !PER_STEP && !kind_sabr(KIND) && KIND != OPTMC_DUPIRE;
    // one-factor log models with deferred drift: the partner's Brownian sum is minus the
    // path's, so the step loop carries one accumulator for both
    constexpr bool MIRROR = DEFER && !TWO;
    // SABR kinds with the hand-written maths: log-vol form, one exponential per step
    // (BGEN: the host has seen that beta is neither 0 nor 1 -- no switch in the step)
    // (3: spot scenarios with a general beta -- power tables, see step_sabr)
    constexpr int LOGV = (kind_sabr(KIND) && NIMPL == 1)
                             ? (BGEN ? (scen_count<decltype(sa.x)>::value > 1 ? 3 : 2) : 1) : 0;
    init_state<KIND, DEFER, LOGV>(sa, c);
    init_state<KIND, DEFER, LOGV>(sb, c);
    sa.j = sb.j = 0.0;
    // A Box-Muller pair feeds one step of a two-factor model, or two consecutive
    // steps of a one-factor model (one step when every step needs its own Poisson word).
    constexpr int SPC = (TWO || LEGACY) ? 1 : 2;
    const uint32_t id_lo = keep_u32((uint32_t)id), id_hi = keep_u32((uint32_t)(id >> 32));

    // ---- jump schedule -------------------------------------------------------------
    int n_j = 0, k_j = 0;                       // jumps of this path, ordinal of the next one
    uint32_t sch[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};   // 8 sorted 16-bit step indices
    auto sched_index = [&](int j) -> uint32_t {
        const uint4 u = philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFCu, (uint32_t)j), A.key);
        return __umulhi(u.x, (uint32_t)A.n_steps);          // floor(U n_steps)
    };
    // The NEXT jump of the path is kept ready -- its step and what it does
    // to the state of the draw and of its partner (the log-jump, or the factor exp(J) for
    // SABR-jump) -- and every step applies it branch-free (+0 / x1 when it is not due).  The
    // Philox calls and size transforms of a jump then run once per jump, in a pass shared by
    // all jumping lanes of the warp for the first one, instead of in a divergent branch at
    // every step at which some lane of the warp jumps (27 % of the steps at lambda dt = 0.01).
    uint32_t pend_step = 0xFFFFFFFFu;
    double pend_a = 0.0, pend_b = 0.0;
    auto fetch_pending = [&]() {                 // requires k_j < n_j <= SCHED_SLOTS
        pend_step = sch[0] & 0xFFFFu;
        sch[0] = (sch[0] >> 16) | (sch[1] << 16);
        sch[1] = (sch[1] >> 16) | (sch[2] << 16);
        sch[2] = (sch[2] >> 16) | (sch[3] << 16);
        sch[3] = (sch[3] >> 16) | 0xFFFF0000u;
        // sizes: fresh draws for the partner, as in the reference's second pass
        pend_a = jump_size<KIND, NIMPL>(
            philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFDu, 1u | ((uint32_t)k_j << 8)), A.key), c, tab);
        if constexpr (ANTI)
            pend_b = jump_size<KIND, NIMPL>(
                philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFDu, 2u | ((uint32_t)k_j << 8)), A.key), c,
                tab);
        if constexpr (KIND == OPTMC_SABR_JUMP) {      // multiplicative on S (mc_kernels.py:274)
            pend_a = exp(pend_a);
            if constexpr (ANTI) pend_b = exp(pend_b);
        }
        ++k_j;
    };
    auto apply_pending = [&](bool due) {
        if constexpr (KIND == OPTMC_SABR_JUMP) {
            sa.x *= due ? pend_a : 1.0;
            if constexpr (ANTI) sb.x *= due ? pend_b : 1.0;
        } else {
            sa.x += due ? pend_a : 0.0;                   // additive in log space (:115,174,323)
            if constexpr (ANTI) sb.x += due ? pend_b : 0.0;
        }
    };
    if constexpr (SCHEDULED) {
        const uint4 e = philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFFu, 0xFEu), A.key);
        if (__builtin_expect(poisson_maybe(e.x, A.pois_p0_path_bits), 0)) {
            const uint4 x = philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFFu, 0xFFu), A.key);
            n_j = poisson_slow(e.x, x.x, A.pois_mu_path, A.pois_p0_path);
            if (n_j <= SCHED_SLOTS) {
                // insertion sort of the n_j indices into 8 ascending 16-bit slots (0xFFFF = none)
                uint32_t v[SCHED_SLOTS];
#pragma unroll
                for (int j = 0; j < SCHED_SLOTS; ++j) v[j] = 0xFFFFu;
                for (int j = 0; j < n_j; ++j) {
                    uint32_t x = sched_index(j);
#pragma unroll
                    for (int q = 0; q < SCHED_SLOTS; ++q) {        // bubble x through the sorted prefix
                        const uint32_t lo = min(v[q], x);
                        x = max(v[q], x);
                        v[q] = lo;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) sch[q] = v[2 * q] | (v[2 * q + 1] << 16);
                if (n_j > 0) fetch_pending();
            }
        }
    }
    auto one_jump = [&](int ordinal) {
        // sizes: fresh draws for the partner, as in the reference's second pass
        const double Ja = jump_size<KIND, NIMPL>(
            philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFDu, 1u | ((uint32_t)ordinal << 8)), A.key),
            c, tab);
        apply_jump<KIND, DEFER>(sa, Ja);
        if constexpr (ANTI) {
            const double Jb = jump_size<KIND, NIMPL>(
                philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFDu, 2u | ((uint32_t)ordinal << 8)),
                              A.key), c, tab);
            apply_jump<KIND, DEFER>(sb, Jb);
        }
    };
    // after the diffusion part of step i: scheduled jumps of that step, then the sink.
    auto after_step = [&](int i) {
        if constexpr (SCHEDULED) {
            static_assert(!DEFER, "kernels with jumps inside the loop carry the drift in the state");
            const bool due = pend_step == (uint32_t)i;
            apply_pending(due);
            if (__builtin_expect(due, 0)) {
                pend_step = 0xFFFFFFFFu;
                while (k_j < n_j) {                  // the path's next jump; may fall on this step too
                    fetch_pending();
                    if (pend_step != (uint32_t)i) break;
                    apply_pending(true);
                    pend_step = 0xFFFFFFFFu;
                }
            }
            if (__builtin_expect(n_j > SCHED_SLOTS, 0)) {
                for (int j = 0; j < n_j; ++j)
                    if (sched_index(j) == (uint32_t)i) one_jump(j);
            }
        }
        if constexpr (PER_STEP) sink(i, sa, sb);
    };

    // One Box-Muller pair (two Philox words, philox.cuh) and the SPC steps it feeds,
    // starting at step i.
    // `whole` (a std::bool_constant): the caller knows that step i + 1 exists too
    auto advance = [&](uint32_t wa, uint32_t wb, int i, auto whole) {
        // (z1, z2): two-factor models -- the mixed, scaled increments (dW_S, xi-weighted
        // dW_v); one-factor models -- sqrt(dt) N(0,1) for this step and the next.
        // Heston / Bates on the hand-written maths: the parts of the pair, one root per path
        if constexpr (NIMPL == 1 && (KIND == OPTMC_HESTON || KIND == OPTMC_BATES)) {
            double L, t1, t2;
            normal_parts_fast<true>(wa, wb, L, t1, t2, tab);
            step_heston_parts<DEFER>(sa, L, t1, t2, c);
            if constexpr (ANTI) step_heston_parts<DEFER>(sb, L, -t1, -t2, c);
            if constexpr (!LEGACY) after_step(i);
            return;
        }
        double z1, z2;
        if constexpr (NIMPL == 0) {
            double a, b;
            normal_pair_lib2(wa, wb, a, b);
            if constexpr (TWO) {
                z1 = fma(c.m12, b, c.m11 * a);
                z2 = fma(c.m22, b, c.m21 * a);
            } else {
                z1 = c.sqrt_dt * a;
                z2 = c.sqrt_dt * b;
            }
        } else {
            normal_pair_fast<TWO>(wa, wb, z1, z2, tab);
        }
        if constexpr (TWO) {
            step_diffusion<KIND, DEFER, LOGV>(sa, z1, z2, c, tab, i);
            if constexpr (ANTI) step_diffusion<KIND, DEFER, LOGV>(sb, -z1, -z2, c, tab, i);
            if constexpr (!LEGACY) after_step(i);
        } else {
            step_diffusion<KIND, DEFER>(sa, z1, 0.0, c, tab, i);
            if constexpr (ANTI && !MIRROR) step_diffusion<KIND, DEFER>(sb, -z1, 0.0, c, tab, i);
            if constexpr (!LEGACY) after_step(i);
            if constexpr (SPC == 2) {
                if (decltype(whole)::value || i + 1 < A.n_steps) {
                    step_diffusion<KIND, DEFER>(sa, z2, 0.0, c, tab, i + 1);
                    if constexpr (ANTI && !MIRROR)
                        step_diffusion<KIND, DEFER>(sb, -z2, 0.0, c, tab, i + 1);
                    after_step(i + 1);
                }
            }
        }
    };
    constexpr std::true_type WHOLE{};
    constexpr std::false_type RAGGED{};
    // Without a Poisson word per step all 128 bits of a Philox call are used: call g feeds
    // pairs 2g = (x, y) and 2g + 1 = (z, w), see philox_pair_words.
    if constexpr (SCHEDULED && PER_STEP) {
        // one copy of the step body (it carries the inlined jump code; for the path kernels the
        // whole-group loops of the diffusion kinds measured the same or 5 % slower, with 768 or
        // 1024 threads too: profiles/r02_experiments.md): the four words of a
        // call are handed out pair by pair
        uint32_t call = 0;
        uint4 a = make_uint4(0u, 0u, 0u, 0u);
        bool second = false;
#pragma unroll 1
        for (int i = 0; i < A.n_steps; i += SPC, second = !second) {
            uint32_t wa, wb;
            if (!second) {
                a = philox4x32_10(make_uint4(id_lo, id_hi, call++, 0u), A.key);
                wa = a.x; wb = a.y;
            } else {
                wa = a.z; wb = a.w;
            }
            advance(wa, wb, i, RAGGED);
        }
    } else if constexpr (!LEGACY && TWO) {
        // Whole calls (two steps) without a bounds test: with the branch-free logarithm the
        // body is one basic block.  (Issuing the next call's Philox rounds ahead of the
        // current steps, whole or in two halves, measured 5 - 7 % slower: ptxas then keeps the
        // round keys in vector registers and reloads constants -- profiles/r02_experiments.md.)
        uint32_t call = 0;
        int i = 0;
        OPTMC_UNROLL_STEPS
        for (; i + 2 <= A.n_steps; i += 2, ++call) {
            const uint4 a = philox4x32_10(make_uint4(id_lo, id_hi, call, 0u), A.key);
            advance(a.x, a.y, i, RAGGED);
            advance(a.z, a.w, i + 1, RAGGED);
        }
        if (i < A.n_steps) {
            const uint4 a = philox4x32_10(make_uint4(id_lo, id_hi, call, 0u), A.key);
            advance(a.x, a.y, i, RAGGED);
        }
    } else if constexpr (!LEGACY) {
        // one-factor models: whole groups of four steps without a single bounds test, then
        // the ragged rest (n_steps % 4 steps) through one more call
        constexpr int GS = 2 * SPC;
        uint32_t call = 0;
        int i = 0;
        OPTMC_UNROLL_STEPS
        for (; i + GS <= A.n_steps; i += GS, ++call) {
            const uint4 a = philox4x32_10(make_uint4(id_lo, id_hi, call, 0u), A.key);
            advance(a.x, a.y, i, WHOLE);
            advance(a.z, a.w, i + SPC, WHOLE);
        }
        if (i < A.n_steps) {
            const uint4 a = philox4x32_10(make_uint4(id_lo, id_hi, call, 0u), A.key);
            advance(a.x, a.y, i, RAGGED);
            if (i + SPC < A.n_steps) advance(a.z, a.w, i + SPC, RAGGED);
        }
    } else {
        OPTMC_UNROLL_STEPS
        for (int i = 0; i < A.n_steps; ++i) {
            const uint4 w = philox4x32_10(make_uint4(id_lo, id_hi, (uint32_t)i, 0u), A.key);
            advance(w.x, w.y, i, RAGGED);
            if (__builtin_expect(poisson_maybe(w.w, A.pois_p0_bits[0]), 0)) {
                double2 J = jumps_slow<KIND, ANTI, NIMPL>(w.w, A.pois_mu[0], A.pois_p0[0], A.key,
                                                          id_lo, id_hi, (uint32_t)i, c, tab);
                // additive in log space (mc_kernels.py:115,174,323); multiplicative on S for
                // SABR-jump (:274): the product of exp(J_j) is exp of the sum
                apply_jump<KIND, DEFER>(sa, J.x);
                if constexpr (ANTI) apply_jump<KIND, DEFER>(sb, J.y);
            }
            if constexpr (PER_STEP) sink(i, sa, sb);
        }
    }
    if constexpr (JUMPS_AT_END) {
        // N ~ Poisson(lambda T) as the sum of pois_chunks independent Poisson(lambda T / chunks)
        // draws (one chunk unless lambda T > 16: the CDF walk of poisson_slow stays short and
        // exp(-mu) far from underflow whatever the intensity), then the summed sizes of N jumps
        int n = 0;
        for (int ch = 0; ch < A.pois_chunks; ++ch) {
            const uint4 e = philox4x32_10(make_uint4(id_lo, id_hi, 0xFFFFFFFFu - (uint32_t)ch, 0xFEu), A.key);
            if (poisson_maybe(e.x, A.pois_p0_path_bits))
                n += poisson_slow(e.x, e.y, A.pois_mu_path, A.pois_p0_path);
        }
        if (n > 0) {
            double2 J = jump_sum<KIND, ANTI, NIMPL>(n, A.key, id_lo, id_hi, 0xFFFFFFFFu, c, tab);
            apply_jump<KIND, DEFER>(sa, J.x);
            if constexpr (ANTI) apply_jump<KIND, DEFER>(sb, J.y);
            if constexpr (KIND == OPTMC_BATES) {
                sa.j = J.x;
                sb.j = J.y;
            }
        }
    }
    finish_deferred<KIND, DEFER>(sa, c);
    if constexpr (ANTI) {
        if constexpr (MIRROR) sb.x = -sa.v;       // sa.v: the Brownian sum W kept by finish_deferred(sa)
        finish_deferred<KIND, DEFER>(sb, c);
    }
}

struct NoSink {
    template <class S>
    __device__ __forceinline__ void operator()(int, const S &, const S &) const {}
};

// Resident 256-thread blocks per SM the register budget is sized for.  Heston / Bates: three
// (85 registers; with four the step loop reloads 19 constants per four steps and runs 1 %
// slower).  The one-factor models and the SABR kinds: four (63 registers, no spills; 2 - 3 %
// faster than three, profiles/r01_sweep_launch_geometry.log).
__host__ __device__ constexpr int euro_min_blocks(int kind, int nimpl, bool sched) {
#ifdef OPTMC_EURO_MIN_BLOCKS_ALL
    return OPTMC_EURO_MIN_BLOCKS_ALL;
#else
    // (the library-maths variants and the SABR-jump kernels -- inlined jump code; the scheduled
    // one keeps its jump schedule and pending jump in registers -- spill at 64 registers)
    return (kind == OPTMC_HESTON || kind == OPTMC_BATES || kind == OPTMC_DUPIRE || nimpl == 0 ||
            kind == OPTMC_SABR_JUMP)
               ? OPTMC_EURO_MIN_BLOCKS : 4;
#endif
}

// What one simulated draw adds to the block's five running sums (and to the optional
// terminal / aux arrays): payoff of the draw and its partner, the control (mean terminal spot).
template <int KIND, bool ANTI>
__device__ __forceinline__ void accumulate_draw(const EuroArgs &A, int64_t g, const State &sa,
                                                const State &sb, double (&acc)[5]) {
    if constexpr (!kind_two_factor(KIND) && KIND != OPTMC_DUPIRE) {
        // log S_T = log S_0 + n drift(sigma) + sigma W + J: W and J do not depend on sigma,
        // so sigma-bumped prices (vega) can be read off this one simulation
        if (A.aux) {
            A.aux[g] = sa.v;
            A.aux[A.n_local + g] = sa.w;
            if constexpr (ANTI) A.aux[2 * A.n_local + g] = sb.w;
        }
    }
    if constexpr (KIND == OPTMC_BATES) {
        // the summed log-jumps of the draw and its partner (replay tests: the diffusion part of
        // a path is a function of the device normals, the jump part is this one number)
        if (A.aux) {
            A.aux[g] = 0.0;
            A.aux[A.n_local + g] = sa.j;
            if constexpr (ANTI) A.aux[2 * A.n_local + g] = sb.j;
        }
    }
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

template <int KIND, bool ANTI, int NIMPL, bool SCHED = false, bool BGEN = false>
__global__ void __launch_bounds__(EURO_THREADS, euro_min_blocks(KIND, NIMPL, SCHED))
k_european(const EuroArgs A) {
    __shared__ RngShared<kind_two_factor(KIND)> tab;
    __shared__ double scratch[5 * 32];
    const RngView tv = fill_rng_shared(&tab, kind_two_factor(KIND) ? A.c.m11 : A.c.sqrt_dt, A.c.m12,
                                       A.c.m21, A.c.m22);
    __syncthreads();
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        State sa, sb;
        simulate_draw<KIND, ANTI, NIMPL, false, SCHED, BGEN>(A, (uint64_t)(A.draw_offset + g), sa, sb, tv,
                                                       NoSink{});
        accumulate_draw<KIND, ANTI>(A, g, sa, sb, acc);
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        double *p = A.partials + (size_t)blockIdx.x * NPART;
#pragma unroll
        for (int k = 0; k < 5; ++k) p[k] = acc[k];
    }
}

// The same kernel on lane-replicated tables (philox.cuh: WideLayout / WideExpLayout): ONE
// block per SM -- 768 threads for the two-factor kinds (85 registers), 1024 for the one-factor
// models (64) -- whose warps share 160 - 224 KB of (nearly) conflict-free tables in dynamic
// shared memory.
#ifndef OPTMC_SJ_THREADS
#define OPTMC_SJ_THREADS 640
#endif
__host__ __device__ constexpr int wide_threads(int kind, bool sched = false) {
#ifdef OPTMC_WIDE_THREADS
    return OPTMC_WIDE_THREADS;
#else
    // (SABR-jump on the jump schedule -- the schedule and the pending jump on top of the SABR
    // state: 15.0 ms at 2^24 x 252 with 640 threads, 15.4 with 512, 15.3 with 768)
    return kind == OPTMC_SABR_JUMP && sched ? OPTMC_SJ_THREADS : kind_two_factor(kind) ? 768 : 1024;
#endif
}
template <int KIND>
using wide_layout_t = std::conditional_t<kind_sabr(KIND), WideExpLayout, WideLayout>;
template <int KIND>
constexpr uint32_t wide_smem_bytes() { return wide_layout_t<KIND>::bytes(kind_two_factor(KIND)); }

// THREADS: the launch bound.  Heston / Bates also exist with 1024 threads (64 registers, ~2 %
// slower per draw): one draw per thread per "row" of sm_count * THREADS draws, so a shard of
// 2^20 draws -- 2^24 paths over 8 GPUs -- is 9.2 rows of 768 threads (a tenth row runs mostly
// idle) but 6.9 rows of 1024; the host takes the cheaper one (optmc.cu: launch_euro_wide).
template <int KIND, bool ANTI, bool SCHED = false, bool BGEN = false,
          int THREADS = wide_threads(KIND, SCHED)>
__global__ void __launch_bounds__(THREADS, 1) k_european_wide(const EuroArgs A) {
    static_assert(KIND != OPTMC_DUPIRE, "Dupire keeps the compact kernel");
    extern __shared__ __align__(16) unsigned char wide_smem[];
    __shared__ double scratch[5 * 32];
    const auto tv = fill_rng_tables<wide_layout_t<KIND>, kind_two_factor(KIND)>(
        wide_smem, kind_two_factor(KIND) ? A.c.m11 : A.c.sqrt_dt, A.c.m12, A.c.m21, A.c.m22);
    __syncthreads();
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        State sa, sb;
        simulate_draw<KIND, ANTI, 1, false, SCHED, BGEN>(A, (uint64_t)(A.draw_offset + g), sa, sb,
                                                         tv, NoSink{});
        accumulate_draw<KIND, ANTI>(A, g, sa, sb, acc);
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        double *p = A.partials + (size_t)blockIdx.x * NPART;
#pragma unroll
        for (int k = 0; k < 5; ++k) p[k] = acc[k];
    }
}

// Spot scenarios of the SABR kinds on ONE set of draws: NS bumped spots -- the
// re-pricings GreekMixin.delta / .gamma make on common random numbers (greek_mixin.py:59-122)
// -- advance side by side; Philox, the Box-Muller pair, the vol path and the jump draws are
// shared, only the spot recurrence (one logarithm and one exponential per scenario and step) is
// repeated.  Same draws, same arithmetic per scenario as k_european_wide: the scenario prices
// equal separate simulations with the bumped spot to rounding.
// partials: [NS][gridDim.x][NPART], the layout k_finish_sweep adds up.
// 512 threads = 128 registers (three spot states for path and partner); on the jump schedule
// (schedule + pending jump in registers) 384 threads = 168 registers: C5 45.1 ms, 448 threads
// 51.0, 512 threads 45.3 (both spill)
#ifndef OPTMC_SCEN_SCHED_THREADS
#define OPTMC_SCEN_SCHED_THREADS 384
#endif
__host__ __device__ constexpr int scen_threads(bool sched) { return sched ? OPTMC_SCEN_SCHED_THREADS : 512; }
template <int KIND, bool ANTI, int NS, bool SCHED = false, bool BGEN = false>
__global__ void __launch_bounds__(scen_threads(SCHED), 1) k_european_scen(const EuroArgs A) {
    static_assert(kind_sabr(KIND), "spot scenarios share the vol path: SABR kinds");
    extern __shared__ __align__(16) unsigned char wide_smem[];
    __shared__ double scratch[5 * 32];     // (the tables take 221 of the 227 KB a block can have)
    auto tv = fill_rng_tables<ScenLayout, true>(wide_smem, A.c.m11, A.c.m12, A.c.m21, A.c.m22);
    if constexpr (BGEN) {
        // power tables behind the generator's: c_i^beta for the 512 nodes of the log table (its
        // entries hold -2 / c_i), two copies, and 2^(e beta) for the exponents -64 .. 63
        double *pc = reinterpret_cast<double *>(wide_smem + ScenLayout::bytes(true));
        double *pe = reinterpret_cast<double *>(wide_smem + ScenLayout::bytes(true) + POW_C_BYTES);
        for (int i = threadIdx.x; i < 512; i += blockDim.x) {
            const double v = pow(-2.0 / g_rng_base.logt[i].x, A.c.beta);
            pc[2 * i] = v;
            pc[2 * i + 1] = v;
        }
        for (int i = threadIdx.x; i < POW_E_N; i += blockDim.x)
            pe[i] = exp2(A.c.beta * (double)(i - POW_E_BIAS));
        tv.pw_c = keep_u32(tv.base + ScenLayout::bytes(true) + (tv.l2 >> 1));
        tv.pw_e = keep_u32(tv.base + ScenLayout::bytes(true) + POW_C_BYTES);
    }
    __syncthreads();
    double acc[5 * NS];
#pragma unroll
    for (int k = 0; k < 5 * NS; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        StateT<Scen<NS>> sa, sb;
        simulate_draw<KIND, ANTI, 1, false, SCHED, BGEN>(A, (uint64_t)(A.draw_offset + g), sa, sb,
                                                         tv, NoSink{});
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const double ST = sa.x.a[k];
            double pay = fmax(A.is_call ? ST - A.strike : A.strike - ST, 0.0);
            double ctl = ST;
            if constexpr (ANTI) {
                const double STb = sb.x.a[k];
                pay = 0.5 * (pay + fmax(A.is_call ? STb - A.strike : A.strike - STb, 0.0));
                ctl = 0.5 * (ST + STb);
            }
            acc[5 * k] += pay;
            acc[5 * k + 1] = fma(pay, pay, acc[5 * k + 1]);
            acc[5 * k + 2] += ctl;
            acc[5 * k + 3] = fma(ctl, ctl, acc[5 * k + 3]);
            acc[5 * k + 4] = fma(pay, ctl, acc[5 * k + 4]);
        }
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double a5[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) a5[j] = acc[5 * k + j];
        block_sum<5>(a5, scratch);
        if (threadIdx.x == 0) {
            double *p = A.partials + ((size_t)k * gridDim.x + blockIdx.x) * NPART;
#pragma unroll
            for (int j = 0; j < 5; ++j) p[j] = a5[j];
        }
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

// max(x, 0) on the integer pipe: both words AND-ed with the complement of the sign (one shift,
// two logic instructions, nothing on the fp64 pipe; written as a comparison the compiler emits
// fmax()'s NaN-quieting sequence).  -0 and negatives give +0; a NaN stays a NaN.
__device__ __forceinline__ double relu_bits(double x) {
    const int hi = __double2hiint(x), keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(x) & keep);
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

// Each thread reads a terminal pair once and prices SWEEP_G contracts on it (block.y =
// group of SWEEP_G contracts): the terminal array crosses L2 n_contracts / SWEEP_G times
// instead of n_contracts times.  The control sums do not depend on the contract.
constexpr int SWEEP_G = 8;

__global__ void __launch_bounds__(256) k_payoff_sweep(const SweepArgs A) {
    constexpr int NACC = 3 * SWEEP_G + 2;
    __shared__ double scratch[NACC * 32];
    const int c0 = blockIdx.y * SWEEP_G;
    // payoff = max(sgn (S - K), 0) with sgn = +-1 as ONE fused multiply-add per spot: sgn S is
    // exact, so the single rounding is that of S - K or K - S -- the same bits as the two-step
    // form, and the maximum on the integer pipe (relu_bits): 31 instructions per pair and contract
    // with fmax(sgn (S - K), 0) before, 13 now.
    double nsK[SWEEP_G], sgn[SWEEP_G];
#pragma unroll
    for (int j = 0; j < SWEEP_G; ++j) {
        const int cidx = min(c0 + j, A.n_contracts - 1);
        sgn[j] = A.is_call[cidx] ? 1.0 : -1.0;
        nsK[j] = -sgn[j] * A.strikes[cidx];
    }
    const double half = A.antithetic ? 0.5 : 1.0;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        const double ST = A.terminal[g];
        const double STb = A.antithetic ? A.terminal[A.n_local + g] : ST;
        const double ctl = A.antithetic ? 0.5 * (ST + STb) : ST;
        acc[3 * SWEEP_G] += ctl;
        acc[3 * SWEEP_G + 1] = fma(ctl, ctl, acc[3 * SWEEP_G + 1]);
        // sums of q = pay + pay_partner (q = pay without antithetic partners); the pair mean
        // 0.5 q enters the moments through exact scalings by 0.5 and 0.25 after the loop
#pragma unroll
        for (int j = 0; j < SWEEP_G; ++j) {
            double q = relu_bits(fma(sgn[j], ST, nsK[j]));
            if (A.antithetic) q += relu_bits(fma(sgn[j], STb, nsK[j]));
            acc[3 * j] += q;
            acc[3 * j + 1] = fma(q, q, acc[3 * j + 1]);
            acc[3 * j + 2] = fma(q, ctl, acc[3 * j + 2]);
        }
    }
#pragma unroll
    for (int j = 0; j < SWEEP_G; ++j) {
        acc[3 * j] *= half;
        acc[3 * j + 1] *= half * half;
        acc[3 * j + 2] *= half;
    }
    block_sum<NACC>(acc, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < SWEEP_G; ++j) {
            if (c0 + j < A.n_contracts) {
                double *p = A.partials + ((size_t)(c0 + j) * gridDim.x + blockIdx.x) * NPART;
                p[0] = acc[3 * j];
                p[1] = acc[3 * j + 1];
                p[2] = acc[3 * SWEEP_G];
                p[3] = acc[3 * SWEEP_G + 1];
                p[4] = acc[3 * j + 2];
            }
        }
    }
}

// Sigma sweep over the stored (W, J) of a one-factor log-Euler simulation:
// scenario s has log S_T = a0[s] + sigma[s] * (+-W) + J with a0 = log S_0 + n drift(sigma).
// Same block layout and partials as k_payoff_sweep (block.y = scenario).
constexpr int SCEN_MAX = 16;
struct SigmaSweepArgs {
    const double *aux;
    int64_t n_local;
    int antithetic;
    int n_scen;
    double a0[SCEN_MAX], sigma[SCEN_MAX], strikes[SCEN_MAX];
    unsigned char is_call[SCEN_MAX];
    double *partials;         // [n_scen][gridDim.x][NPART]
};

__global__ void __launch_bounds__(256) k_sigma_sweep(const SigmaSweepArgs A) {
    __shared__ double scratch[5 * 32];
    const int sidx = blockIdx.y;
    const double a0 = A.a0[sidx], sg = A.sigma[sidx], K = A.strikes[sidx];
    const bool call = A.is_call[sidx] != 0;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        const double W = A.aux[g];
        double ST = exp((a0 + sg * W) + A.aux[A.n_local + g]);
        double pay = fmax(call ? ST - K : K - ST, 0.0);
        double ctl = ST;
        if (A.antithetic) {
            double STb = exp((a0 - sg * W) + A.aux[2 * A.n_local + g]);
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
        double *p = A.partials + ((size_t)sidx * gridDim.x + blockIdx.x) * NPART;
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
