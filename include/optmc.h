/*
 * optmc.h -- C ABI of liboptmc.so, the B200 (sm_100a) Monte Carlo engine that
 * replaces the numba hot path of Diljit22/optpricing.
 *
 * The reference has no FFI today: its "operator API" for this path is the set
 * of Python functions below, and every entry point here names the one it
 * replaces (paths relative to /root/reference/src/optpricing/techniques/).
 * INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (optmc_last_error
 *     gives the text); nothing throws across the ABI;
 *   - one process drives one GPU: a ctx is bound to one device; calls on one
 *     ctx are not re-entrant; distinct ctx objects are independent;
 *   - *_dev pointers are device memory owned by the caller (in the Python
 *     host layer: torch tensors); `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream).  The library itself allocates only a
 *     few KB of scratch at ctx creation;
 *   - functions whose name ends in _async only enqueue work on `stream`; the
 *     others synchronise the stream once before returning and write their
 *     results to host memory.
 *   - all floating point is IEEE binary64 (the step loop of the European kernels
 *     has an optional binary32 mode, optmc_sim.precision).
 */
#ifndef OPTMC_H
#define OPTMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct optmc_ctx optmc_ctx;

/* Dispatch order of monte_carlo.py:130-197 / american_monte_carlo.py:87-150 */
enum optmc_model_kind {
    OPTMC_BSM = 0,       /* kernels/mc_kernels.py:18-34   path_kernels.py:17-35   */
    OPTMC_HESTON = 1,    /* mc_kernels.py:43-79           path_kernels.py:43-77   */
    OPTMC_MERTON = 2,    /* mc_kernels.py:88-118          path_kernels.py:85-121  */
    OPTMC_BATES = 3,     /* mc_kernels.py:127-177         path_kernels.py:130-180 */
    OPTMC_SABR = 4,      /* mc_kernels.py:186-218         path_kernels.py:189-217 */
    OPTMC_SABR_JUMP = 5, /* mc_kernels.py:227-276         path_kernels.py:225-274 */
    OPTMC_KOU = 6,       /* mc_kernels.py:285-329         path_kernels.py:277-324 */
    OPTMC_DUPIRE = 7     /* mc_kernels.py:340-359 (European only; see lv_* below)  */
};

/* The reference's kernel_params dict (monte_carlo.py:128-197), flattened.
   Fields a model does not use are ignored. */
typedef struct optmc_model {
    int32_t kind;       /* enum optmc_model_kind */
    int32_t reserved;
    double r, q;        /* risk-free rate, dividend yield */
    double sigma;       /* BSM, Merton, Kou */
    double v0;          /* Heston/Bates initial variance; SABR initial vol */
    double kappa, theta, rho, vol_of_vol; /* Heston/Bates; rho also SABR */
    double alpha, beta; /* SABR */
    double lambda, mu_j, sigma_j;         /* Merton, Bates, SABR-jump */
    double p_up, eta1, eta2;              /* Kou */
    /* Dupire local volatility.  The reference's dupire_kernel calls a Python
       vol_surface_func(t_i, S) from inside the step loop (mc_kernels.py:351-357),
       which cannot run on a device (nor, in fact, under numba's nopython mode).
       Here the host samples that callable once per pricing into a device table
       lv_table_dev[i * lv_n_s + j] = vol_surface(i * dt, exp(lv_log_s_lo + j / lv_inv_h)),
       i < lv_n_t (>= n_steps: exactly the times the kernel evaluates), and the
       step interpolates linearly in log S (flat beyond the ends). */
    const double *lv_table_dev;
    int32_t lv_n_s, lv_n_t;
    double lv_log_s_lo, lv_inv_h;
} optmc_model;

/* How the two Brownian drivers of a variance model are mixed (Philox mode). */
enum optmc_correlation {
    /* dw1 = rho*a + rho_bar*b, dw2 = a, then the kernel's own
       rho*dw1 + rho_bar*dw2: what MonteCarloTechnique does
       (monte_carlo.py:222-231 followed by mc_kernels.py:67). */
    OPTMC_CORR_REFERENCE_EUROPEAN = 0,
    /* dw1 = a, dw2 = b (american_monte_carlo.py:181-185): corr(dW_S, dW_v) = rho */
    OPTMC_CORR_INDEPENDENT_DRAWS = 1
};

/* One simulation job (monte_carlo.py:199-215). */
typedef struct optmc_sim {
    double s0;            /* Stock.spot */
    double maturity;      /* Option.maturity; dt = maturity / n_steps */
    int64_t n_draws;      /* draws of the WHOLE job: n_paths/2 if antithetic else n_paths
                             (num_draws, monte_carlo.py:210) */
    int64_t draw_offset;  /* first draw this call simulates (multi-GPU partition) */
    int64_t n_local;      /* draws this call simulates: [draw_offset, draw_offset+n_local) */
    int32_t n_steps;
    int32_t antithetic;   /* 1: each draw also yields its negated partner */
    uint64_t seed;        /* Philox4x32-10 key; the stream of draw g depends only on (seed, g) */
    int32_t correlation;  /* enum optmc_correlation */
    int32_t precision;    /* enum optmc_precision: arithmetic of the step loop (optmc_european*) */
} optmc_sim;

/* OPTMC_FP64 (default): everything in IEEE binary64.  OPTMC_FP32: the optional
   single-precision mode -- path state, normals and per-step arithmetic in binary32
   with the SFU approximations of log/exp/sin/cos/sqrt; the once-per-path finish
   (drift, exp, payoff) and the moments stay binary64.  Statistical agreement only;
   not available for Dupire, fed draws or the Longstaff-Schwartz path store. */
enum optmc_precision { OPTMC_FP64 = 0, OPTMC_FP32 = 1 };

/* Per contract, over this call's samples (a sample = one draw: the mean of the
   antithetic pair when antithetic, else one path; payoff p and control c = S_T
   are undiscounted):
     m[0] = n samples   m[1] = sum p   m[2] = sum p^2
     m[3] = sum c       m[4] = sum c^2 m[5] = sum p*c
   Sums over disjoint draw ranges (ranks) add. */
#define OPTMC_NMOM 6

int optmc_version(void);
int optmc_ctx_create(int device, optmc_ctx **out);
void optmc_ctx_destroy(optmc_ctx *ctx);
const char *optmc_last_error(const optmc_ctx *ctx); /* ctx may be NULL: last create error */

/*
 * European pricing, on-device RNG.  Replaces the whole of
 * MonteCarloTechnique._simulate_sde_path + payoff + mean
 * (monte_carlo.py:199-257 and :110-115): no increments are materialised.
 * Writes OPTMC_NMOM doubles per contract.
 *   optmc_european_async: moments_dev is device memory; nothing is synchronised.
 *   optmc_european:       moments_host is host memory; returns after one sync.
 * terminal_dev (optional, may be NULL): receives S_T, n_local doubles for the
 * draws followed (if antithetic) by n_local doubles for their partners -- the
 * order of np.concatenate([ST, ST_anti]) (monte_carlo.py:255).
 */
int optmc_european_async(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                         int32_t n_contracts, const double *strikes,
                         const int32_t *is_call, double *moments_dev,
                         double *terminal_dev, void *stream);
int optmc_european(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                   int32_t n_contracts, const double *strikes, const int32_t *is_call,
                   double *moments_host, double *terminal_dev, void *stream);

/*
 * Spot scenarios on common draws for the SABR kinds.  GreekMixin.delta / .gamma
 * (techniques/base/greek_mixin.py:59-122) re-price 2 / 3 times with the spot bumped, on common
 * random numbers.  For SABR and SABR-jump the vol path does not depend on the spot, so the
 * bumped paths are carried side by side through ONE simulation: n_scen (2 or 3) spots
 * s0 * spot_factors[k] share Philox, the normals, the vol path and the jump draws.  Writes
 * OPTMC_NMOM doubles per scenario, equal (to rounding) to optmc_european_async called per
 * scenario with s0 * spot_factors[k] and the same seed.
 */
int optmc_european_scen_async(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                              int32_t n_scen, const double *spot_factors, double strike,
                              int32_t is_call, double *moments_dev, void *stream);

/*
 * Fed-draw kernels: the drop-in for the 7 terminal kernels of mc_kernels.py and
 * the 7 path kernels of path_kernels.py, same inputs in the same C-order
 * layouts, all in device memory:
 *   dw1_dev        double[n_paths][n_steps]   (`dw` for one-factor models)
 *   dw2_dev        double[n_paths][n_steps]   (two-factor models, else NULL)
 *   counts_dev     int64 [n_paths][n_steps]   (jump models, else NULL)
 *   jumps_dev      double[n_jumps]  jump sizes in the order the reference pulls
 *                  them from its RNG: by step, then path, then jump
 *                  (Merton/Bates/Kou: additive log-jumps; SABR-jump: the
 *                  N(mu_j, sigma_j) draws, exponentiated here).  A path whose counts ask
 *                  for sizes beyond n_jumps comes out NaN; nothing is read out of bounds
 *   x0             log_s0, or s0 for the SABR kinds (monte_carlo.py:217-220)
 *   sign           +1.0, or -1.0 for the antithetic second pass
 *                  (monte_carlo.py:247-254) without touching the buffers
 *   terminal_dev   double[n_paths] log S_T (S_T for SABR kinds) or NULL
 *   paths_dev      double[n_paths][n_steps+1] spot matrix or NULL
 *   path_sum_mode  1: jumps of one path-step are summed, then added
 *                  (path_kernels.py, kou_kernel); 0: added one by one
 *   workspace_dev  optmc_fed_workspace_bytes() bytes, 8-byte aligned
 */
int64_t optmc_fed_workspace_bytes(int32_t kind, int64_t n_paths, int32_t n_steps);
int optmc_fed_async(optmc_ctx *ctx, const optmc_model *model, double x0, double dt,
                    int64_t n_paths, int32_t n_steps, const double *dw1_dev,
                    const double *dw2_dev, const int64_t *counts_dev,
                    const double *jumps_dev, int64_t n_jumps, double sign,
                    int32_t path_sum_mode, double *terminal_dev, double *paths_dev,
                    void *workspace_dev, int64_t workspace_bytes, void *stream);

/*
 * Strike sweep over stored terminal spots: moments for n_contracts payoffs of
 * the same maturity from one simulation (BASELINE config 3).  terminal_dev is
 * laid out as optmc_european writes it.
 */
int optmc_payoff_sweep_async(optmc_ctx *ctx, const double *terminal_dev, int64_t n_local,
                             int32_t antithetic, int32_t n_contracts,
                             const double *strikes, const int32_t *is_call,
                             double *moments_dev, void *stream);

/*
 * Vega from one simulation for the one-factor log-Euler models
 * (BSM, Merton, Kou), whose terminal value is
 *     log S_T = log S_0 + n drift(sigma) + sigma W + J,
 * W = the path's Brownian sum, J = its summed log-jumps, neither depending on sigma.
 *   optmc_european_aux_async  = optmc_european_async that also stores, per draw,
 *                               aux_dev[g] = W, aux_dev[n_local + g] = J and (antithetic)
 *                               aux_dev[2 n_local + g] = J of the partner, whose W is -W
 *                               (BSM, Merton, Kou; Bates: W slot 0, J = the path's summed
 *                               log-jumps -- with the device normals of optmc_normals that is
 *                               everything a path's terminal value is a function of)
 *   optmc_sigma_sweep_async   payoff moments of up to 16 scenarios (a0 = log S_0 + n drift,
 *                             sigma, strike, is_call) over the stored (W, J): what
 *                             GreekMixin.vega's two re-pricings on common random numbers
 *                             compute (greek_mixin.py:125-175), without re-simulating
 */
int optmc_european_aux_async(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                             int32_t n_contracts, const double *strikes, const int32_t *is_call,
                             double *moments_dev, double *terminal_dev, double *aux_dev,
                             void *stream);
int optmc_sigma_sweep_async(optmc_ctx *ctx, const double *aux_dev, int64_t n_local,
                            int32_t antithetic, int32_t n_scenarios, const double *a0,
                            const double *sigma, const double *strikes, const int32_t *is_call,
                            double *moments_dev, void *stream);

/*
 * Single-step terminal samplers: the pure-Levy and exact-sampler branches of
 * MonteCarloTechnique.price (monte_carlo.py:103-106) and _simulate_levy_terminal
 * (:259-280), which in the reference call each model's numpy/scipy sampler on
 * the host.  The host layer folds the model parameters into p[] (the laws, in
 * the reference's own parameterisation):
 *   OPTMC_LEVY_VG          models/vg.py:136-168          p = {T/nu, nu, theta, sigma}
 *                          x = theta g + sigma sqrt(g) z,  g = p[1] Gamma(p[0])
 *   OPTMC_LEVY_NIG         models/nig.py:136-169         p = {mu_ig, beta}
 *                          x = beta w + sqrt(w) z,  w ~ inverse Gaussian(mean mu_ig, shape 1)
 *   OPTMC_LEVY_CGMY1       models/cgmy.py:134-178 (Y=1)  p = {C T, 1/M, 1/G}
 *                          x = p[1] Gamma(p[0]) - p[2] Gamma'(p[0])
 *   OPTMC_LEVY_HYPERBOLIC  models/hyperbolic.py:123-152  p = {b_gig, s_gig, beta, loc, scale,
 *                          p_gig - 1, mode m, u-, u+, log g(m)} (optmc_levy_gig_setup fills 5..9)
 *                          x = loc + scale (beta v + sqrt(v) z),  v = s_gig GIG(p_gig, b_gig)
 *   OPTMC_LEVY_CEV         models/cev.py:49-96           p = {df, nc, k, 1/(2(1-gamma))}, df > 1
 *                          S_T = (ncx2(df, nc) / k)^p[3]            (no drift, no exp)
 *   OPTMC_LEVY_LOGNORMAL   models/cev.py:81-85 (gamma -> 1)  p = {drift, sigma sqrt(T)}
 *                          S_T = s0 exp(p[0] + p[1] z)
 * Levy kinds: S_T = s0 exp(drift + x), drift = (r - q) T - log Re phi_raw(-i)  (:276-280).
 * Samples [offset, offset + n_local) of one seed are drawn (disjoint Philox
 * counters per sample: any partition over ranks gives the same samples).
 * Moments as optmc_european_async (control = S_T); terminal_dev (optional,
 * required for n_contracts > 1) receives the n_local terminal spots.
 */
enum optmc_levy_kind {
    OPTMC_LEVY_VG = 0, OPTMC_LEVY_NIG = 1, OPTMC_LEVY_CGMY1 = 2, OPTMC_LEVY_HYPERBOLIC = 3,
    OPTMC_LEVY_CEV = 4, OPTMC_LEVY_LOGNORMAL = 5
};
#define OPTMC_LEVY_NPARAM 12

typedef struct optmc_levy {
    int32_t kind;       /* enum optmc_levy_kind */
    int32_t reserved;
    double p[OPTMC_LEVY_NPARAM];
} optmc_levy;

/* Bounding rectangle of the ratio-of-uniforms GIG sampler (Hoermann & Leydold 2014,
   mode-shift branch; needs p_gig >= 1 or b_gig > 1): reads p[0] (b_gig) and p[5]
   (p_gig - 1), writes p[6..9]. */
int optmc_levy_gig_setup(optmc_levy *levy);

int optmc_levy_terminal_async(optmc_ctx *ctx, const optmc_levy *levy, double s0, double drift,
                              uint64_t seed, int64_t offset, int64_t n_local,
                              int32_t n_contracts, const double *strikes, const int32_t *is_call,
                              double *moments_dev, double *terminal_dev, void *stream);

/*
 * Longstaff-Schwartz, replacing AmericanMonteCarloTechnique.price
 * (american_monte_carlo.py:52-73).
 *
 * Path store: step-major, store_dev[(i-1) * ld + j] = spot of local path j at
 * date i (i = 1..n_steps), j < n_paths_local = n_local * (antithetic ? 2 : 1),
 * partners at j + n_local; ld >= n_paths_local, a multiple of 2.
 *
 *   optmc_lsmc_paths_async   simulate (Philox) and fill the store; replaces
 *                            _simulate_full_sde_paths (:152-207) + path_kernels.py
 *   optmc_lsmc_load_async    fill the store from a reference-layout matrix
 *                            paths_dev[n_paths][n_steps+1] (column 0 ignored)
 *   optmc_lsmc_step_async    one fused sweep for exercise date `date`
 *                            (call with date = n_steps-1, ..., 1, 0):
 *                              date == n_steps-1: cashflow = payoff at T      (american_mc_kernels.py:48-51)
 *                              otherwise: exercise decision of date+1 from the
 *                                normal equations in gram_dev slot [date+1] --
 *                                exercise where intrinsic > fitted continuation (:72-80);
 *                                a rank-deficient or non-positive-definite system
 *                                skips the decision like the bare `except` of :81-82
 *                              cashflow *= discount                            (:54)
 *                              date >= 1: in-the-money power sums sum x^k (k <= 2d)
 *                                and sum y x^k (k <= d) of x = S/K - 1 into slot
 *                                [date] (:57-69; X'X and X'y of :72 are read off them)
 *                              date == 0: slot [0] = n, sum cf, sum cf^2       (:84)
 *                            With several ranks, all-reduce slot [date] between calls.
 *   optmc_lsmc               the whole loop for one GPU, one sync, host result;
 *                            by default a single persistent launch (one block
 *                            per SM, a grid barrier per date) running the same
 *                            per-path arithmetic as optmc_lsmc_step_async
 *                            out_host[0..2] = n, sum, sum of squares of the
 *                            discounted cashflows
 *
 * gram_dev: n_steps slots of OPTMC_LSMC_GRAM doubles:
 *   [0] n_itm, [1..2d] sum x^k, then [16 + k] sum y x^k (k = 0..d).  degree <= 7.
 */
#define OPTMC_LSMC_GRAM 24
#define OPTMC_LSMC_MAX_DEGREE 7
#define OPTMC_MAX_PEERS 8
#define OPTMC_IPC_HANDLE_BYTES 64

int optmc_lsmc_paths_async(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                           double *store_dev, int64_t ld, void *stream);
int optmc_lsmc_load_async(optmc_ctx *ctx, const double *paths_dev, int64_t n_paths,
                          int32_t n_steps, double *store_dev, int64_t ld, void *stream);
int optmc_lsmc_step_async(optmc_ctx *ctx, const double *store_dev, int64_t ld, int64_t n_paths,
                          int32_t n_steps, int32_t date, double strike, int32_t is_call,
                          double discount, int32_t degree, double *cashflow_dev,
                          double *gram_dev, void *stream);
int optmc_lsmc(optmc_ctx *ctx, const double *store_dev, int64_t ld, int64_t n_paths,
               int32_t n_steps, double strike, int32_t is_call, double r, double dt,
               int32_t degree, double *cashflow_dev, double *gram_dev, double *out_host,
               void *stream);

/*
 * Multi-GPU Longstaff-Schwartz without a collective library on the data path
 * (SURVEY 8 e2): one process per GPU, each owning its shard of the path store.
 *   optmc_lsmc_peer_export   writes the CUDA IPC handle (OPTMC_IPC_HANDLE_BYTES) of
 *                            this ctx's exchange buffer
 *   optmc_lsmc_peer_attach   handles = world * OPTMC_IPC_HANDLE_BYTES bytes, rank order
 *                            (all-gathered by the host layer); maps every peer's buffer
 *                            (NVLink peer access).  world <= OPTMC_MAX_PEERS.
 *                            Call with world = 1 to detach.
 * After attach and optmc_set_int(ctx, "lsmc_use_peers", 1), optmc_lsmc runs the
 * persistent kernel on every rank at once (set it back to 0 for a local call): per
 * exercise date each rank's power sums cross NVLink as 24 flag-carrying 16-byte
 * stores into every peer's buffer and every rank adds the rows in rank order, so
 * all ranks solve bit-identical normal equations; out_host holds the GLOBAL
 * (n, sum, sum of squares).  Ranks must make the same sequence of attached
 * optmc_lsmc calls (SPMD); a missing rank ends in a bounded time-out error, not a hang.
 */
int optmc_lsmc_peer_export(optmc_ctx *ctx, void *handle_out);
int optmc_lsmc_peer_attach(optmc_ctx *ctx, int32_t rank, int32_t world, const void *handles);

/*
 * Read n (<= 384) doubles of device memory -- payoff moments -- into host memory through the
 * context's pinned buffer: one asynchronous copy on `stream`, one synchronisation.  What
 * optmc_european does at its end, for callers of the *_async entry points.
 */
int optmc_read_doubles(optmc_ctx *ctx, const double *src_dev, int32_t n, double *out_host,
                       void *stream);

/*
 * The all-reduce that ends a partitioned pricing (the payoff moments, 48 bytes
 * per contract), as ONE kernel of this library on the pricing's stream instead of a library
 * collective: every rank stores its n doubles into every attached rank's exchange buffer
 * (peer stores over NVLink, flag-carrying 16-byte slots) and adds the world's rows from its
 * local copy in rank order -- bit-identical sums on every rank.  vals_dev is reduced in place.
 * Needs optmc_lsmc_peer_attach (the same exported buffer); ranks must make the same sequence
 * of calls (SPMD).  A rank that never arrives ends the others in a wall-clock time-out
 * (20 s): the values come back as NaN 0x7FF8DEADDEADDEAD and the next synchronising call fails.
 */
int optmc_peer_allreduce_async(optmc_ctx *ctx, double *vals_dev, int32_t n, void *stream);
/*
 * optmc_european for one rank of a partitioned pricing: simulation of this rank's draws, the
 * all-reduce of the moments over the attached ranks (optmc_peer_allreduce_async) and the read
 * into host memory in ONE call and one synchronisation -- what MonteCarloTechnique.price runs
 * per rank under an NCCL process group (monte_carlo.py:65-116 has no counterpart: it is
 * single-process).  moments_host receives the GLOBAL moments, identical on every rank.
 */
int optmc_european_allreduce(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                             int32_t n_contracts, const double *strikes, const int32_t *is_call,
                             double *moments_host, double *terminal_dev, void *stream);

/*
 * Diagnostics used by the tests and by bench.py's roofline.
 *   optmc_philox_raw     out_dev[4*i..] = Philox4x32-10(counter = (lo(i), hi(i), c2, c3), key = seed)
 *   optmc_normals        out_dev[2*i], out_dev[2*i+1] = the two N(0,1) of Box-Muller pair
 *                        `step` of draw i: Philox call c2 = g of stream 0 feeds the pairs 2g
 *                        (words x, y) and 2g + 1 (words z, w); from the two words (wa, wb) of a
 *                        pair, u = ((wa : wb[13:0]) + 1/2) 2^-46, angle = 2 pi ((wb >> 14) + 1/2)
 *                        / 2^18, (z1, z2) = sqrt(-2 ln u) (cos, sin)(angle).  A pair drives one
 *                        step of a two-factor model or two consecutive steps of a one-factor
 *                        model (the per-step jump kernels -- jump path kernels with lambda T >
 *                        2 -- use one call per step instead: words x, y for the pair, w for the
 *                        Poisson draw).  These are the normals the step loops of
 *                        optmc_european / optmc_lsmc_paths consume (tests/test_gpu_replay.py).
 *   optmc_fp64_peak      runs a register-resident DFMA loop on every SM for
 *                        ~`iters` iterations and returns fp64 FMA instructions
 *                        per second (thread-level): the measured denominator of
 *                        the fp64-pipe roofline
 */
int optmc_philox_raw(optmc_ctx *ctx, uint64_t seed, int64_t n, uint32_t c2, uint32_t c3,
                     uint32_t *out_dev, void *stream);
int optmc_normals(optmc_ctx *ctx, uint64_t seed, int64_t n, int32_t step, double *out_dev,
                  void *stream);
int optmc_fp64_peak(optmc_ctx *ctx, int64_t iters, double *fma_per_sec, double *elapsed_ms);

/* Tunables: "normals_impl" (0 = CUDA math library Box-Muller, 1 = hand-written,
   default 1), "euro_blocks_per_sm" (0 = occupancy API decides), "euro_threads"
   (threads per block of the European kernel: 64, 128 or 256), "lsmc_mode"
   (optmc_lsmc: 1 = every exercise date in one persistent cooperative launch,
   default; 0 = one optmc_lsmc_step_async launch per date), "lsmc_use_peers"
   (1 = the next optmc_lsmc calls are joint calls of the attached ranks, see
   optmc_lsmc_peer_attach; 0 = local), "jump_schedule" (kernels whose jumps act
   inside the step loop -- the path kernels of the jump models, optmc_lsmc_paths_async,
   and European SABR-jump, optmc_european*: 1 = draw the path's Poisson(lambda T) count
   and its jump steps once, default while lambda T <= 2; 0 = one Poisson(lambda dt) draw
   per step; two samplers of the same law), "fed_stream" (optmc_fed_async: 1 = one thread streams one row, default;
   0 = the shared-memory tiled kernel), "euro_wide" (1, default: European jobs of at least one row of draws run on
   lane-replicated tables, one block per SM; 0 = always the compact kernel), "lsmc_ring" (sweep of the
   persistent Longstaff-Schwartz kernel fed by a cp.async ring in shared memory: 0 = never,
   1 = always, 2 = from 2^19 paths per GPU, default), "lsmc_timing_all" (diagnostics: every
   block stamps, see "lsmc_timing"). */
int optmc_set_int(optmc_ctx *ctx, const char *key, int64_t value);

/* Pointer-valued options: "lsmc_timing" (diagnostics: device memory for
   (n_steps + 1) * 4 uint64 that block 0 of the persistent Longstaff-Schwartz kernel
   fills with clock64 stamps per date; NULL = off -- tools/lsmc_phases.py; with
   "lsmc_timing_all" = 1 every block stamps and the buffer is gridDim.x times as large,
   [block][n_steps + 1][4] -- tools/lsmc_block_phases.py). */
int optmc_set_ptr(optmc_ctx *ctx, const char *key, void *value);

/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
int64_t optmc_launch_count(const optmc_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* OPTMC_H */
