// lsmc.cuh -- Longstaff-Schwartz backward induction on a step-major path store.
//
// Replaces longstaff_schwartz_pricer (kernels/american_mc_kernels.py:36-84) and
// the materialised (n_paths, n_steps+1) matrix of
// AmericanMonteCarloTechnique._simulate_full_sde_paths (american_monte_carlo.py:152-207).
//
// Store layout: store[(i-1) * ld + j] = spot of path j at exercise date i
// (i = 1..n_steps).  Every per-date pass is a unit-stride sweep over paths.
//
// Regression: the reference fits raw monomials [1, S, .., S^d] by inv(X'X) X'y
// on in-the-money paths.  X'X is a Hankel matrix of power sums, so the kernels
// accumulate sum x^k (k <= 2d) and sum y x^k (k <= d) of the shifted variable
// x = S/K - 1 (same polynomial space, hence the same fitted values, but
// cond ~1e6 instead of ~1e20 at d = 3), and a (d+1)x(d+1) equilibrated Cholesky
// is solved redundantly by every block.
#pragma once
#include "european.cuh"

namespace optmc {

constexpr int LSMC_THREADS = 256;

// ---- path store fill (Philox) ---------------------------------------------
template <class V>
struct StoreSink {
    double *store;
    int64_t ld, g, n_local;
    int kind;
    bool anti;
    V tab;
    __device__ __forceinline__ void operator()(int i, const State &sa, const State &sb) const {
        double *row = store + (size_t)i * ld;
        row[g] = step_spot(kind, sa, tab);                      // path_kernels.py:34
        if (anti) row[n_local + g] = step_spot(kind, sb, tab);
    }
};

template <int KIND, bool ANTI, int NIMPL, bool SCHED = false, bool BGEN = false>
__global__ void __launch_bounds__(EURO_THREADS) k_lsmc_paths(const EuroArgs A, double *store,
                                                             int64_t ld) {
    __shared__ RngShared<kind_two_factor(KIND)> tab;
    const RngView tv = fill_rng_shared(&tab, kind_two_factor(KIND) ? A.c.m11 : A.c.sqrt_dt, A.c.m12,
                                       A.c.m21, A.c.m22);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        State sa, sb;
        StoreSink<RngView> sink{store, ld, g, A.n_local, KIND, ANTI, tv};
        simulate_draw<KIND, ANTI, NIMPL, true, SCHED, BGEN>(A, (uint64_t)(A.draw_offset + g), sa, sb,
                                                            tv, sink);
    }
}

// The same fill on lane-replicated tables, one block per SM (european.cuh: k_european_wide).
// Every kind takes the layout with the exp table: the log models exponentiate each stored spot.
#ifndef OPTMC_PATHS_SCHED_THREADS
#define OPTMC_PATHS_SCHED_THREADS 640
#endif
__host__ __device__ constexpr int paths_wide_threads(int kind, bool sched) {
    // (jump schedule: 640 threads measured 3 % faster than 512 for Merton / Kou / Bates, 2 % slower for SABR-jump)
    return sched ? (kind == OPTMC_SABR_JUMP ? 512 : OPTMC_PATHS_SCHED_THREADS) : kind_two_factor(kind) ? 768 : 1024;
}
template <int KIND>
constexpr uint32_t paths_wide_smem_bytes() { return WideExpLayout::bytes(kind_two_factor(KIND)); }

template <int KIND, bool ANTI, bool SCHED = false, bool BGEN = false>
__global__ void __launch_bounds__(paths_wide_threads(KIND, SCHED), 1)
k_lsmc_paths_wide(const EuroArgs A, double *store, int64_t ld) {
    extern __shared__ __align__(16) unsigned char wide_smem[];
    const auto tv = fill_rng_tables<WideExpLayout, kind_two_factor(KIND)>(
        wide_smem, kind_two_factor(KIND) ? A.c.m11 : A.c.sqrt_dt, A.c.m12, A.c.m21, A.c.m22);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        State sa, sb;
        StoreSink<RngViewS> sink{store, ld, g, A.n_local, KIND, ANTI, tv};
        simulate_draw<KIND, ANTI, 1, true, SCHED, BGEN>(A, (uint64_t)(A.draw_offset + g), sa, sb, tv,
                                                        sink);
    }
}

// ---- path store fill from a reference-layout matrix -----------------------
// paths[n_paths][n_steps+1] (row-major) -> store[(i-1)*ld + p], i = 1..n_steps
__global__ void __launch_bounds__(256) k_lsmc_load(const double *paths, int64_t n_paths,
                                                   int n_steps, double *store, int64_t ld) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int i0 = blockIdx.y * 32;                            // date index - 1
    for (int rr = ty; rr < 32; rr += 8) {
        int64_t p = p0 + rr;
        int i = i0 + tx;
        tile[rr][tx] = (p < n_paths && i < n_steps) ? paths[(size_t)p * (n_steps + 1) + 1 + i] : 0.0;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        int i = i0 + rr;
        int64_t p = p0 + tx;
        if (p < n_paths && i < n_steps) store[(size_t)i * ld + p] = tile[tx][rr];
    }
}

// Solve the (d+1)x(d+1) normal equations built from the power sums.
// Returns false when the date must be skipped: fewer in-the-money paths than
// basis functions, or the equilibrated Cholesky meets a non-positive pivot --
// the counterpart of the LinAlgError the reference swallows (:71-82).
// Fully unrolled for the degree, so the factor lives in registers, and every
// division by a pivot is a multiplication by its reciprocal square root: one
// thread solves the system on the critical path between two exercise dates.
// 1/sqrt(a) for a > 0 (normal range): MUFU.RSQ64H seed (2^-22) + two Newton steps, ~1 ulp.
// The library rsqrt() costs ~300 dependent cycles a call; eight of them sat on the critical
// path between two exercise dates.
__device__ __forceinline__ double rsqrt_newton(double a) {
    double y, h;
    rsqrt_seed(a, y, h);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double e = fma(-(a * y), y, 1.0);      // 1 - a y^2
        y = fma(h, e, y);                            // y + (y / 2) e
        h = 0.5 * y;
    }
    return y;
}

template <int DEG>
__device__ __forceinline__ bool lsmc_solve(const double *gram, double *beta) {
    constexpr int m = DEG + 1;
    const double *P = gram;        // P[k] = sum x^k (P[0] = n_itm)
    const double *Q = gram + 16;   // Q[k] = sum y x^k
    if (!(P[0] >= (double)m)) return false;
    double L[m][m], d[m], rhs[m], inv[m];
#pragma unroll
    for (int a = 0; a < m; ++a) {
        double diag = P[2 * a];
        if (!(diag > 0.0)) return false;
        d[a] = rsqrt_newton(diag);
    }
#pragma unroll
    for (int a = 0; a < m; ++a) {
        rhs[a] = Q[a] * d[a];
#pragma unroll
        for (int b = 0; b <= a; ++b) L[a][b] = P[a + b] * d[a] * d[b];
    }
#pragma unroll
    for (int j = 0; j < m; ++j) {
        double s = L[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[j][k], L[j][k], s);
        if (!(s > 1e-13)) return false;
        inv[j] = rsqrt_newton(s);
#pragma unroll
        for (int i = j + 1; i < m; ++i) {
            double t = L[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L[i][k], L[j][k], t);
            L[i][j] = t * inv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < m; ++i) {
        double t = rhs[i];
#pragma unroll
        for (int k = 0; k < i; ++k) t = fma(-L[i][k], rhs[k], t);
        rhs[i] = t * inv[i];
    }
#pragma unroll
    for (int i = m - 1; i >= 0; --i) {
        double t = rhs[i];
#pragma unroll
        for (int k = i + 1; k < m; ++k) t = fma(-L[k][i], rhs[k], t);
        rhs[i] = t * inv[i];
    }
#pragma unroll
    for (int a = 0; a < m; ++a) beta[a] = rhs[a] * d[a];
    return true;
}

// ---- one fused pass per exercise date ---------------------------------------
// Call for date i (i = n_steps-1 ... 0) does, for every path, in one sweep:
//   1. (i == n_steps-1)  y = payoff at T                       american_mc_kernels.py:48-51
//      (otherwise)       y = cashflow; exercise decision of date i+1 with the
//                        regression solved from gram_in (the all-reduced power
//                        sums of date i+1): y = intrinsic where intrinsic > fit   :72-80
//   2. y *= exp(-r dt)                                          :54 (and :84 for i == 0)
//   3. (i >= 1) cashflow = y; in-the-money paths of date i add to the power
//      sums sum x^k, sum y x^k of x = S_i/K - 1                 :57-69
//      (i == 0) accumulate sum y, sum y^2 instead: the final mean :84
// so the cashflow vector is read and written once per date and row i+1, read by
// the previous call, is normally still in the 126 MB L2.  Per-block partials;
// the last block to finish adds them in fixed order into gram_out.
template <int DEG>
struct LsmcAcc {
    static constexpr int NP = 2 * DEG + 1, NQ = DEG + 1, NV = NP + NQ;
    double a[NV];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < NV; ++k) a[k] = 0.0;
    }
};

// a > 0 && a > c ? a : alt with ONE select: the second comparison takes the first one's
// predicate as its AND input (the compiler emits a select per condition otherwise: 4 more
// instructions per path in a loop that is bound by instruction issue).
__device__ __forceinline__ double select_if_positive_and_above(double a, double c, double alt) {
    double r;
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.gt.f64 p, %1, 0d0000000000000000;\n\t"
        "setp.gt.and.f64 q, %1, %2, p;\n\t"
        "selp.f64 %0, %1, %3, q;\n\t}"
        : "=d"(r)
        : "d"(a), "d"(c), "d"(alt));
    return r;
}

// One path of the fused sweep (see k_lsmc_step).  Returns the discounted
// cashflow to store.
// Branch-free on purpose: within a warp some lanes are almost always in the
// money, so a branch saves nothing, while straight-line code lets the compiler
// interleave the independent paths a thread handles per trip.  Out-of-the-money
// paths enter the power sums with weight 0.
template <int DEG, bool INIT, bool FINAL>
__device__ __forceinline__ double lsmc_path(double y_in, double s_dec, double s_acc, bool use_fit,
                                            const double *beta, double K, double inv_K, bool call,
                                            double disc, LsmcAcc<DEG> &acc) {
    double y;
    // intrinsic value sgn (S - K) as ONE fused multiply-add, sgn = +-1: the products are exact,
    // so the single rounding is that of S - K or K - S -- the same bits as the two-sided form,
    // without computing both sides and selecting (2 DADD + 2 FSEL per use)
    const double sgn = call ? 1.0 : -1.0, nsK = call ? -K : K;
    const double iv_dec = fma(sgn, s_dec, nsK);
    if constexpr (INIT) {
        y = fmax(iv_dec, 0.0);                                        // american_mc_kernels.py:48-51
    } else {
        const double x = fma(s_dec, inv_K, -1.0);
        double cont = beta[DEG];
#pragma unroll
        for (int k = DEG - 1; k >= 0; --k) cont = fma(cont, x, beta[k]);   // X @ beta, :73
        // in the money (:62) and intrinsic above the fitted continuation (:75-80)
        y = use_fit ? select_if_positive_and_above(iv_dec, cont, y_in) : y_in;
    }
    y *= disc;                                                        // :54 / :84
    if constexpr (FINAL) {
        acc.a[0] += 1.0;
        acc.a[1] += y;
        acc.a[2] = fma(y, y, acc.a[2]);
    } else {
        constexpr int NP = LsmcAcc<DEG>::NP, NQ = LsmcAcc<DEG>::NQ;
        const double iv = fma(sgn, s_acc, nsK);                       // :57-60
        const double w = iv > 0.0 ? 1.0 : 0.0;
        const double x = fma(s_acc, inv_K, -1.0);
        double pw[NP];
        pw[0] = w;
        pw[1] = w * x;
#pragma unroll
        for (int k = 2; k < NP; ++k) pw[k] = pw[k / 2] * pw[k - k / 2];   // w^2 = w: still w x^k
#pragma unroll
        for (int k = 0; k < NP; ++k) acc.a[k] += pw[k];
#pragma unroll
        for (int k = 0; k < NQ; ++k) acc.a[NP + k] = fma(y, pw[k], acc.a[NP + k]);
    }
    return y;
}

template <int DEG, bool INIT, bool FINAL>
__global__ void __launch_bounds__(LSMC_THREADS)
k_lsmc_step(const double *__restrict__ row_dec,   // row of date i+1
            const double *__restrict__ row_acc,   // row of date i (unused when FINAL)
            int64_t n, double K, double inv_K, int is_call, double disc,
            const double *__restrict__ gram_in, double *__restrict__ cashflow,
            double *partials, unsigned int *ticket, double *gram_out) {
    constexpr int NV = LsmcAcc<DEG>::NV, NP = LsmcAcc<DEG>::NP;
    __shared__ double scratch[(NV > LSMC_THREADS / 32 ? NV : LSMC_THREADS / 32) * 32];
    __shared__ double beta[8];
    __shared__ bool ok, last;
    if (threadIdx.x == 0) ok = INIT ? false : lsmc_solve<DEG>(gram_in, beta);
    __syncthreads();
    const bool use_fit = ok, call = is_call != 0;
    LsmcAcc<DEG> acc;
    acc.clear();
    // two paths per thread per trip: 16-byte loads and stores (rows are 16 B aligned:
    // ld is even and the store base comes from the allocator)
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double2 *dec2 = reinterpret_cast<const double2 *>(row_dec);
    const double2 *acc2 = reinterpret_cast<const double2 *>(row_acc);
    double2 *cf2 = reinterpret_cast<double2 *>(cashflow);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += stride) {
        double2 sd = (INIT || use_fit) ? __ldcs(dec2 + q) : make_double2(0.0, 0.0);
        double2 y = INIT ? make_double2(0.0, 0.0) : cf2[q];
        double2 sa = FINAL ? make_double2(0.0, 0.0) : acc2[q];
        y.x = lsmc_path<DEG, INIT, FINAL>(y.x, sd.x, sa.x, use_fit, beta, K, inv_K, call, disc, acc);
        y.y = lsmc_path<DEG, INIT, FINAL>(y.y, sd.y, sa.y, use_fit, beta, K, inv_K, call, disc, acc);
        if (!FINAL) cf2[q] = y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {             // odd tail
        int64_t p = n - 1;
        double y = lsmc_path<DEG, INIT, FINAL>(INIT ? 0.0 : cashflow[p], row_dec[p],
                                               FINAL ? 0.0 : row_acc[p], use_fit, beta, K, inv_K,
                                               call, disc, acc);
        if (!FINAL) cashflow[p] = y;
    }
    block_sum<NV>(acc.a, scratch);
    double *mine = partials + (size_t)blockIdx.x * OPTMC_LSMC_GRAM;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) mine[k] = acc.a[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        // Fixed-order sum of the per-block partials by the last block to finish:
        // lane = entry, warp w takes blocks w, w+8, ...; eight loads in flight.
        __threadfence();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int NW = LSMC_THREADS / 32;
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (lane < NV) {
            for (unsigned int b = warp; b < gridDim.x; b += NW * 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    unsigned int bb = b + NW * j;
                    if (bb < gridDim.x)
                        s[j] += __ldcg(partials + (size_t)bb * OPTMC_LSMC_GRAM + lane);
                }
            }
        }
        scratch[warp * 32 + lane] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
        __syncthreads();
        if (warp == 0 && lane < NV) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) tot += scratch[w * 32 + lane];
            gram_out[lane < NP ? lane : 16 + (lane - NP)] = tot;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}


// ---- the whole induction in ONE launch (single GPU) --------------------------
// k_lsmc_step costs one launch, one memset and one serial "last block adds the
// partials" tail per exercise date; at 2^22 paths that overhead is as large as
// the sweep itself.  k_lsmc_persistent keeps one block resident on every SM for
// all dates (cooperative launch):
//   * each block owns a fixed slice of paths: its cashflows are touched by no
//     other SM and stay in L2, its rows stream from HBM once per date;
//   * the per-date exchange needs no separate barrier: every block publishes its
//     NV partial sums as self-validating 16-byte slots (lo, tag, hi, tag) -- the
//     flag-in-data scheme of NCCL's LL protocol -- and every block polls all G
//     rows and adds them in the same fixed order, so all blocks solve identical
//     normal equations one memory round trip after the last block has published.
// Per-path arithmetic is lsmc_path, the same code k_lsmc_step runs.
constexpr int LSMC_P_STRIDE = 24;           // slots per block row (NV <= 23)
// One 512-thread block per SM: 128 registers per thread hold the 3d+2 running sums
// next to the loads in flight without spilling (1024 threads x 64 registers
// spilled in the sweep and measured 7 % slower at degree 3).
#ifndef OPTMC_LSMC_THREADS
#define OPTMC_LSMC_THREADS 512
#endif
__host__ __device__ constexpr int lsmc_p_threads(int) { return OPTMC_LSMC_THREADS; }
#ifndef OPTMC_LSMC_SLICE_ALIGN
#define OPTMC_LSMC_SLICE_ALIGN 8      // pairs: block slices start on 128-byte lines
#endif
#ifndef OPTMC_LSMC_UNROLL
#define OPTMC_LSMC_UNROLL 4
#endif
// trips in flight per thread
__host__ __device__ constexpr int lsmc_p_unroll(int deg) { return deg <= 3 ? OPTMC_LSMC_UNROLL : 2; }

struct LsmcPersistArgs {
    const double *store;
    int64_t ld, n;
    int n_steps;
    double K, inv_K, disc;
    int is_call;
    double *cashflow;
    int prefetch_next_row;     // L2-prefetch the next date's row during the exchange
    uint4 *slots;              // [2][gridDim.x][LSMC_P_STRIDE], flag-carrying partials
    uint32_t nonce;            // per-launch tag prefix
    // multi-GPU (world > 1): rank totals travel over NVLink as the same kind of slots.
    // peer_rank_slots[r] is rank r's buffer [2 call parity][2 epoch parity][world][LSMC_P_STRIDE]
    // mapped into this process (CUDA IPC); [rank] is the local one.
    int rank, world;
    uint32_t peer_nonce;       // per distributed call, identical on every rank
    uint4 *peer_rank_slots[OPTMC_MAX_PEERS];
    uint4 *local_rank_slots;   // = peer_rank_slots[rank]
    unsigned int *error;       // set to 1 if an exchange times out
    double *gram;              // [n_steps][OPTMC_LSMC_GRAM], written by block 0
    unsigned long long *timing;   // optional [n_steps + 1][4] clock64 stamps of block 0 (SM cycles)
    int timing_all;               // every block stamps: [gridDim.x][n_steps + 1][4]
};

// SM cycle counter: block 0 stays on one SM for the whole launch, so differences of its
// stamps are exact phase durations in SM clocks (%globaltimer only ticks every ~0.26 us).
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    return (unsigned long long)clock64();
}

__device__ __forceinline__ void ll_store(uint4 *slot, double v, uint32_t tag) {
    uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    // Two 64-bit exchanges, performed AT the L2 slice: the pollers see the slot ~0.7 us
    // earlier than after a posted st.volatile.v4 (poll + sum 1.9 -> 1.1 us per date at 2^20
    // paths, tools/lsmc_phases.py).  Each half carries its own tag, so the pair need not
    // be atomic together.
    unsigned long long a = ((unsigned long long)tag << 32) | lo, b = ((unsigned long long)tag << 32) | hi;
    atomicExch(reinterpret_cast<unsigned long long *>(slot), a);
    atomicExch(reinterpret_cast<unsigned long long *>(slot) + 1, b);
}

// the same store towards a peer GPU's memory (system scope: it crosses NVLink)
__device__ __forceinline__ void ll_store_sys(uint4 *slot, double v, uint32_t tag) {
    uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    // a posted 16-byte store: system-scope exchanges over NVLink measured 2 % slower (2 GPUs)
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"(lo), "r"(tag),
                 "r"(hi), "r"(tag)
                 : "memory");
}

__device__ __forceinline__ uint4 ll_peek(const uint4 *slot) {
    uint4 w;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                 : "l"(slot)
                 : "memory");
    return w;
}

// Wall clock (ns).  %globaltimer ticks every ~0.26 us: fine for a time-out, too coarse for the
// phase stamps above.
__device__ __forceinline__ unsigned long long wall_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// How long a block waits for another block -- or another PROCESS's kernel on a peer GPU --
// before it gives up: wall-clock, generous (first-call module loads, a rank delayed by the
// host), and independent of clocks and of how fast a poll is.
#ifndef OPTMC_LL_TIMEOUT_NS
#define OPTMC_LL_TIMEOUT_NS 20000000000ull
#endif
// A block that has given up keeps publishing -- correctly tagged, so nobody waits for it --
// but publishes POISON instead of sums computed from garbage: whoever consumes it gives up
// too, so every block of every rank ends with the error flag set instead of some of them
// silently returning numbers built on a missing contribution.
constexpr uint32_t LL_POISON_HI = 0x7FF8DEADu, LL_POISON_LO = 0xDEADDEADu;
__device__ __forceinline__ double ll_poison() {
    return __hiloint2double((int)LL_POISON_HI, (int)LL_POISON_LO);
}

// `w` is a first look at the slot; spin until both halves carry `tag`.  False on time-out or
// poison (a lost block or rank, never the normal path); `dead`: this thread has already given
// up once and does not wait again.
__device__ __forceinline__ bool ll_wait(const uint4 *slot, uint32_t tag, uint4 w, double &v,
                                        bool dead = false) {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (w.y != tag || w.w != tag) {
        if (dead) return false;
        if ((++spins & 0x3FFu) == 0u) {
            const unsigned long long now = wall_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > OPTMC_LL_TIMEOUT_NS) return false;
        }
        w = ll_peek(slot);
    }
    if (w.z == LL_POISON_HI && w.x == LL_POISON_LO) return false;
    v = __hiloint2double((int)w.z, (int)w.x);
    return true;
}

// One date's sweep over this block's slice of path pairs [q0, q1): U trips in
// flight per thread, i.e. up to 3 U 16-byte loads before the first use.
// Measured alternatives that were slower at 2^22 x 50 (profiles/r01_lsmc_persistent.md):
// shared-memory-resident cashflows (+3 %), and a TMA (cp.async.bulk + mbarrier)
// ring for the two store rows (2.07 ms vs 1.06 ms: without per-load eviction
// hints the rows and the cashflows no longer stay in L2).
template <int DEG, bool INIT, bool FINAL>
__device__ __forceinline__ void lsmc_sweep(const double *__restrict__ row_dec,
                                           const double *__restrict__ row_acc,
                                           double *__restrict__ cashflow, int64_t q0, int64_t q1,
                                           bool use_fit, const double *beta, double K, double inv_K,
                                           bool call, double disc, LsmcAcc<DEG> &acc) {
    constexpr int T = lsmc_p_threads(DEG), U = lsmc_p_unroll(DEG);
    const double2 *dec2 = reinterpret_cast<const double2 *>(row_dec);
    const double2 *acc2 = reinterpret_cast<const double2 *>(row_acc);
    double2 *cf2 = reinterpret_cast<double2 *>(cashflow);
    const double2 zero = make_double2(0.0, 0.0);
    const bool need_dec = INIT || use_fit;
    int64_t qa = q0 + threadIdx.x;
    // whole groups of U trips: no bounds tests, no zero fills
    for (; qa + (int64_t)(U - 1) * T < q1; qa += (int64_t)U * T) {
        double2 sd[U], sa[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = qa + (int64_t)u * T;
            sd[u] = need_dec ? __ldcs(dec2 + q) : zero;
            y[u] = !INIT ? cf2[q] : zero;
            sa[u] = !FINAL ? acc2[q] : zero;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = qa + (int64_t)u * T;
            y[u].x = lsmc_path<DEG, INIT, FINAL>(y[u].x, sd[u].x, sa[u].x, use_fit, beta, K, inv_K,
                                                 call, disc, acc);
            y[u].y = lsmc_path<DEG, INIT, FINAL>(y[u].y, sd[u].y, sa[u].y, use_fit, beta, K, inv_K,
                                                 call, disc, acc);
            if (!FINAL) cf2[q] = y[u];
        }
    }
    // the ragged rest of the slice, one trip at a time
#pragma unroll 1
    for (; qa < q1; qa += T) {
        const double2 sd = need_dec ? __ldcs(dec2 + qa) : zero;
        double2 y = !INIT ? cf2[qa] : zero;
        const double2 sa = !FINAL ? acc2[qa] : zero;
        y.x = lsmc_path<DEG, INIT, FINAL>(y.x, sd.x, sa.x, use_fit, beta, K, inv_K, call, disc, acc);
        y.y = lsmc_path<DEG, INIT, FINAL>(y.y, sd.y, sa.y, use_fit, beta, K, inv_K, call, disc, acc);
        if (!FINAL) cf2[qa] = y;
    }
}

// ---- the same sweep fed by an asynchronous-copy ring ------------------------------------
// Loads held in registers are issued in a burst and then waited for: the four warps of a
// sub-partition, in step with each other, all stall at once (ncu: long_scoreboard 4.6 cycles
// per issue, issue slots 39 % busy).  Here every thread keeps S trips in flight as
// cp.async (LDGSTS) copies into its own slots of a shared-memory ring -- no registers, no
// burst: one new trip is requested per trip consumed -- and reads a trip back with three
// LDS.128 once cp.async.wait_group says it has landed.  Only the owning thread touches a slot,
// so no barrier is needed.  L2 policy per stream, as before: the decision row is read for the
// last time (evict-first), the cashflows live in L2 for the whole launch (evict-last).
#ifndef OPTMC_LSMC_ASYNC_STAGES
#define OPTMC_LSMC_ASYNC_STAGES 4            // ring stages of the RING instantiation (a power of two)
#endif
#ifndef OPTMC_LSMC_ASYNC_TRIPS
#define OPTMC_LSMC_ASYNC_TRIPS 1             // trips (pairs of paths) per ring stage
#endif
// Measured at degree 3 (tools/lsmc_phases.py, us per date, register-staged -> ring of 4, block
// slices on 128-byte lines): 2^22 paths 16.6 -> 13.8, 2^21 9.7 -> 9.3, 2^20 6.9 -> 6.8, 2^19
// 6.1 -> 5.4, 2^17 4.25 -> 4.35: the host takes the ring from 2^19 paths per GPU up (optmc.cu).
// (With slices starting 16 bytes into a 32-byte sector the ring LOST at 2^20 and 2^21 paths:
// every other block's sweep took 7-9 us instead of 5.8 and the exchange waited for them.)
__host__ __device__ constexpr bool lsmc_has_ring(int deg) { return deg <= 3; }
__host__ __device__ constexpr size_t lsmc_ring_bytes(int deg) {
    return (size_t)OPTMC_LSMC_ASYNC_STAGES * OPTMC_LSMC_ASYNC_TRIPS * 3u * (size_t)lsmc_p_threads(deg) * 16u;
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void *src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
                 "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Request the first S - 1 stages of a sweep.  They depend on nothing the exchange produces, so
// the persistent kernel issues them for the NEXT date before it publishes its sums: the first
// trips' HBM latency then overlaps the ~4 us of exchange and solve.  (The decision row is
// requested whether or not the fit will be used.)
template <int DEG, bool INIT, bool FINAL, int S>
__device__ __forceinline__ void lsmc_ring_prologue(const double *__restrict__ row_dec,
                                                   const double *__restrict__ row_acc,
                                                   const double *__restrict__ cashflow, int64_t q0,
                                                   int64_t q1, uint32_t ring, uint64_t pol_first,
                                                   uint64_t pol_last) {
    constexpr int T = lsmc_p_threads(DEG), U = OPTMC_LSMC_ASYNC_TRIPS;
    const double2 *dec2 = reinterpret_cast<const double2 *>(row_dec);
    const double2 *acc2 = reinterpret_cast<const double2 *>(row_acc);
    const double2 *cf2 = reinterpret_cast<const double2 *>(cashflow);
    const uint32_t mine = ring + threadIdx.x * 16u;
    const int64_t qa = q0 + threadIdx.x;
    // Rolled on purpose: unrolled, ptxas 12.9 encoded these hinted copies of one instantiation as
    // `LDGSTS [R+UR0], desc[UR1]` with neither uniform register ever written -- "an illegal
    // instruction was encountered" at run time.  tests/test_host_api.py scans the built library
    // for that encoding.
#pragma unroll 1
    for (int s = 0; s < S - 1; ++s) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = qa + (int64_t)(s * U + u) * T;
            if (q < q1) {
                const uint32_t base = mine + (uint32_t)(((s * U + u) * 3) * T) * 16u;
                cp_async16_hint(base, dec2 + q, pol_first);
                if (!INIT) cp_async16_hint(base + (uint32_t)T * 16u, cf2 + q, pol_last);
                if (!FINAL) cp_async16(base + 2u * (uint32_t)T * 16u, acc2 + q);
            }
        }
        cp_async_commit();
    }
}

template <int DEG, bool INIT, bool FINAL, int S>
__device__ __forceinline__ void lsmc_sweep_async(const double *__restrict__ row_dec,
                                                 const double *__restrict__ row_acc,
                                                 double *__restrict__ cashflow, int64_t q0,
                                                 int64_t q1, bool use_fit, const double *beta,
                                                 double K, double inv_K, bool call, double disc,
                                                 LsmcAcc<DEG> &acc, uint32_t ring, uint64_t pol_first,
                                                 uint64_t pol_last, bool prologue_done) {
    static_assert(S >= 2 && (S & (S - 1)) == 0, "stages: a power of two");
    constexpr int T = lsmc_p_threads(DEG), U = OPTMC_LSMC_ASYNC_TRIPS;
    const double2 *dec2 = reinterpret_cast<const double2 *>(row_dec);
    const double2 *acc2 = reinterpret_cast<const double2 *>(row_acc);
    double2 *cf2 = reinterpret_cast<double2 *>(cashflow);
    const double2 zero = make_double2(0.0, 0.0);
    // The coefficients live in registers for the whole sweep (the cp.async statements are
    // memory barriers to the compiler: read through the pointer they were re-loaded from shared
    // memory every trip).  "No usable fit" is folded into them: a continuation value of +inf
    // is never exceeded, so the loop carries no use_fit predicate.
    double b[DEG + 1];
#pragma unroll
    for (int k = 0; k <= DEG; ++k) b[k] = use_fit ? beta[k] : (k == 0 ? __longlong_as_double(0x7FF0000000000000ll) : 0.0);
    // slot of (stage, trip, array) for this thread: consecutive threads, consecutive 16 bytes
    const uint32_t mine = ring + threadIdx.x * 16u;
    auto slot = [&](int stage, int u, int arr) {
        return mine + (uint32_t)(((stage * U + u) * 3 + arr) * T) * 16u;
    };
    // 32-bit byte offsets into the block's slice (the host bounds a slice at 2^31 bytes): one
    // offset register serves the three streams
    const char *dec_b = reinterpret_cast<const char *>(dec2 + q0);
    const char *acc_b = FINAL ? nullptr : reinterpret_cast<const char *>(acc2 + q0);   // (no row at the last date)
    char *cf_b = reinterpret_cast<char *>(cf2 + q0);
    const uint32_t end = (uint32_t)(q1 - q0) * 16u;
    constexpr uint32_t TRIP = (uint32_t)T * 16u;
    auto request = [&](int stage, uint32_t oa) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t o = oa + (uint32_t)u * TRIP;
            if (o < end) {
                cp_async16_hint(slot(stage, u, 0), dec_b + o, pol_first);   // (also when no fit is usable)
                if (!INIT) cp_async16_hint(slot(stage, u, 1), cf_b + o, pol_last);
                if (!FINAL) cp_async16(slot(stage, u, 2), acc_b + o);
            }
        }
        cp_async_commit();                   // one group per stage, requested or not
    };
    if (!prologue_done)
        lsmc_ring_prologue<DEG, INIT, FINAL, S>(row_dec, row_acc, cashflow, q0, q1, ring, pol_first,
                                                pol_last);
    int stage = 0;
#pragma unroll 1
    for (uint32_t oa = threadIdx.x * 16u; oa < end; oa += (uint32_t)U * TRIP) {
        request((stage + S - 1) & (S - 1), oa + (uint32_t)(S - 1) * U * TRIP);
        cp_async_wait<S - 1>();              // everything but the S - 1 newest groups has landed
        double2 sd[U], y[U], sa[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            sd[u] = lds_f64x2(slot(stage, u, 0));
            y[u] = !INIT ? lds_f64x2(slot(stage, u, 1)) : zero;
            sa[u] = !FINAL ? lds_f64x2(slot(stage, u, 2)) : zero;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t o = oa + (uint32_t)u * TRIP;
            if (u == 0 || o < end) {
                y[u].x = lsmc_path<DEG, INIT, FINAL>(y[u].x, sd[u].x, sa[u].x, true, b, K, inv_K, call,
                                                     disc, acc);
                y[u].y = lsmc_path<DEG, INIT, FINAL>(y[u].y, sd[u].y, sa[u].y, true, b, K, inv_K, call,
                                                     disc, acc);
                if (!FINAL) *reinterpret_cast<double2 *>(cf_b + o) = y[u];
            }
        }
        stage = (stage + 1) & (S - 1);
    }
    cp_async_wait<0>();
}

// Sum P (16 or 32) values per lane over the warp with P - 1 (+1) shuffles instead
// of 5 P: at every halving step a lane keeps one half of its values and trades
// the other half with its partner.  Afterwards v[0] of lane l is the warp total of
// entry l (P = 32) or l / 2 (P = 16).  Fixed order, hence reproducible.
template <int P>
__device__ __forceinline__ void warp_transpose_sum(double (&v)[P], int lane) {
    static_assert(P == 16 || P == 32, "P must be 16 or 32");
    int m = 16;
#pragma unroll
    for (int c = P / 2; c >= 1; c >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < c; ++i) {
            const double send = up ? v[i] : v[i + c];
            const double keep = up ? v[i + c] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
    if constexpr (P == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int DEG, bool USE_RING = false>
__global__ void __launch_bounds__(lsmc_p_threads(DEG), 1)
k_lsmc_persistent(const LsmcPersistArgs A) {
    constexpr int NV = LsmcAcc<DEG>::NV, NP = LsmcAcc<DEG>::NP;
    constexpr int T = lsmc_p_threads(DEG), NW = T / 32;
    constexpr int GPAD = 161;      // >= gridDim.x (checked on the host), odd: no bank conflicts
    constexpr int SLOTS = 4;       // polls in flight per thread
    constexpr int P = NV <= 16 ? 16 : 32;
    __shared__ double scratch[NV * 32];
    __shared__ double vals[NV * GPAD];
    __shared__ double tot[OPTMC_LSMC_GRAM];
    __shared__ double peer_vals[OPTMC_MAX_PEERS * OPTMC_LSMC_GRAM];
    __shared__ double beta[8];
    __shared__ bool ok;
    constexpr int RING = USE_RING ? OPTMC_LSMC_ASYNC_STAGES : 0;
    extern __shared__ __align__(16) unsigned char ring_smem[];
    uint32_t ring = 0;
    uint64_t pol_first = 0, pol_last = 0;
    if constexpr (RING > 0) {
        ring = (uint32_t)__cvta_generic_to_shared(ring_smem);
        pol_first = l2_policy_evict_first();
        pol_last = l2_policy_evict_last();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int G = gridDim.x;
    const bool call = A.is_call != 0;
    // this block's slice of path pairs
    const int64_t n2 = A.n >> 1;
    // Slices start on 128-byte lines (8 pairs): with slices of 7085 pairs (2^21 paths) every other
    // block started 16 bytes into a 32-byte sector and its cp.async sweep took 7-9 us instead
    // of 5.8 (tools/lsmc_block_phases.py).
    const int64_t chunk = (((n2 + G - 1) / G) + OPTMC_LSMC_SLICE_ALIGN - 1) / OPTMC_LSMC_SLICE_ALIGN *
                          OPTMC_LSMC_SLICE_ALIGN;
    const int64_t q0 = min((int64_t)blockIdx.x * chunk, n2);
    const int64_t q1 = min(q0 + chunk, n2);
    const bool odd_owner = (A.n & 1) && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned int epoch = 0;
    bool timed_out = false;
    bool ring_ready = false;       // the ring already holds the first stages of the coming sweep
    const bool stamp = A.timing && (blockIdx.x == 0 || A.timing_all) && threadIdx.x == 0;
    unsigned long long *const stamps =
        A.timing + (A.timing_all ? (size_t)blockIdx.x * (A.n_steps + 1) * 4 : 0);
    for (int date = A.n_steps - 1; date >= -1; --date) {
        const bool first = date == A.n_steps - 1;
        if (stamp) stamps[(size_t)(date + 1) * 4 + 0] = globaltimer_ns();
        if (!first) {
            // totals of the previous sweep.  All threads poll: slot i = b * NV + e goes to
            // thread i mod T, every thread has its (<= SLOTS) loads in flight at once, and
            // the values land in shared memory as vals[e][b]; then warp e adds entry e
            // over the blocks in a fixed order (lane-strided partial sums, shuffle tree).
            const uint32_t tag = (A.nonce << 12) | ((epoch - 1) & 0xFFFu);
            const uint4 *src = A.slots + (size_t)((epoch - 1) & 1u) * G * LSMC_P_STRIDE;
            const unsigned int n_slots = G * NV;
            if (threadIdx.x < OPTMC_LSMC_GRAM) tot[threadIdx.x] = 0.0;   // unused slots read as zero
            for (unsigned int i0 = 0; i0 < n_slots; i0 += T * SLOTS) {
                uint4 w[SLOTS];
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) {
                    const unsigned int i = i0 + j * T + threadIdx.x;
                    if (i < n_slots) w[j] = ll_peek(src + (size_t)(i / NV) * LSMC_P_STRIDE + i % NV);
                }
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) {
                    const unsigned int i = i0 + j * T + threadIdx.x;
                    if (i < n_slots) {
                        const unsigned int b = i / NV, e = i % NV;
                        double v = 0.0;
                        if (!ll_wait(src + (size_t)b * LSMC_P_STRIDE + e, tag, w[j], v, timed_out))
                            timed_out = true;
                        vals[e * GPAD + b] = v;
                    }
                }
            }
            timed_out = __syncthreads_or(timed_out);     // block-wide: one lost slot poisons the block
            if (stamp) stamps[(size_t)(date + 1) * 4 + 1] = globaltimer_ns();
            for (int e = warp; e < NV; e += NW) {
                double t = 0.0;
                for (unsigned int b = lane; b < G; b += 32) t += vals[e * GPAD + b];
                t = warp_sum(t);
                if (lane == 0) tot[e < NP ? e : 16 + (e - NP)] = t;
            }
            __syncthreads();
            if (A.world > 1) {
                // Second level, across GPUs: block 0 stores this rank's totals into every
                // rank's buffer (peer stores over NVLink), every block of every rank polls
                // its LOCAL copy and adds the W rows in rank order -- the all-reduce of the
                // Gram matrix fused into the induction kernel, one NVLink hop per date.
                const uint32_t ptag = (A.peer_nonce << 12) | ((epoch - 1) & 0xFFFu);
                const size_t base = (((size_t)(A.peer_nonce & 1u) * 2 + ((epoch - 1) & 1u)) * A.world) *
                                    LSMC_P_STRIDE;
                if (blockIdx.x == 0 && threadIdx.x < OPTMC_LSMC_GRAM * A.world) {
                    const int r = threadIdx.x / OPTMC_LSMC_GRAM, e = threadIdx.x % OPTMC_LSMC_GRAM;
                    ll_store_sys(A.peer_rank_slots[r] + base + (size_t)A.rank * LSMC_P_STRIDE + e,
                                 timed_out ? ll_poison() : tot[e], ptag);
                }
                // one thread per (rank, entry): the W polls are in flight together (one L2 round
                // trip, not W), then entry e is added over the ranks in rank order
                if (threadIdx.x < OPTMC_LSMC_GRAM * A.world) {
                    const int r = threadIdx.x / OPTMC_LSMC_GRAM, e = threadIdx.x % OPTMC_LSMC_GRAM;
                    const uint4 *slot = A.local_rank_slots + base + (size_t)r * LSMC_P_STRIDE + e;
                    double x = 0.0;
                    if (!ll_wait(slot, ptag, ll_peek(slot), x, timed_out)) timed_out = true;
                    peer_vals[threadIdx.x] = x;
                }
                timed_out = __syncthreads_or(timed_out);   // (block 0 has also read tot[] for its stores)
                if (threadIdx.x < OPTMC_LSMC_GRAM) {
                    double v = 0.0;
                    for (int r = 0; r < A.world; ++r) v += peer_vals[r * OPTMC_LSMC_GRAM + threadIdx.x];
                    tot[threadIdx.x] = v;
                }
                __syncwarp();            // OPTMC_LSMC_GRAM <= 32: writers and the solving lane share warp 0
            }
            if (warp == 0) {
                if (blockIdx.x == 0 && lane < OPTMC_LSMC_GRAM)
                    A.gram[(size_t)(date + 1) * OPTMC_LSMC_GRAM + lane] = tot[lane];
                if (lane == 0) ok = date >= 0 ? lsmc_solve<DEG>(tot, beta) : false;
            }
            __syncthreads();
            if (date < 0) break;
        }
        if (stamp) stamps[(size_t)(date + 1) * 4 + 2] = globaltimer_ns();
        const bool use_fit = first ? false : ok;
        const bool final_ = date == 0;
        const double *row_dec = A.store + (size_t)date * A.ld;
        const double *row_acc = final_ ? nullptr : A.store + (size_t)(date - 1) * A.ld;
        LsmcAcc<DEG> acc;
        acc.clear();
#define OPTMC_SWEEP(I, F)                                                                       \
    do {                                                                                        \
        if constexpr (RING > 0)                                                                 \
            lsmc_sweep_async<DEG, I, F, (RING > 0 ? RING : 2)>(row_dec, row_acc, A.cashflow, q0, q1, use_fit, \
                                              beta, A.K, A.inv_K, call, A.disc, acc, ring,      \
                                              pol_first, pol_last, ring_ready);                 \
        else                                                                                    \
            lsmc_sweep<DEG, I, F>(row_dec, row_acc, A.cashflow, q0, q1, use_fit, beta, A.K,     \
                                  A.inv_K, call, A.disc, acc);                                  \
        if (odd_owner) {                                                                        \
            const int64_t p = A.n - 1;                                                          \
            double y = lsmc_path<DEG, I, F>(I ? 0.0 : A.cashflow[p], row_dec[p],                \
                                            F ? 0.0 : row_acc[p], use_fit, beta, A.K, A.inv_K,  \
                                            call, A.disc, acc);                                 \
            if (!F) A.cashflow[p] = y;                                                          \
        }                                                                                       \
    } while (0)
        if (first && final_) OPTMC_SWEEP(true, true);
        else if (first) OPTMC_SWEEP(true, false);
        else if (final_) OPTMC_SWEEP(false, true);
        else OPTMC_SWEEP(false, false);
#undef OPTMC_SWEEP
        if constexpr (RING > 0) {
            // the next date's first stages, requested before this date's sums are exchanged
            ring_ready = false;
            if (date >= 1) {
                const double *nd = A.store + (size_t)(date - 1) * A.ld;
                if (date - 1 == 0)
                    lsmc_ring_prologue<DEG, false, true, (RING > 0 ? RING : 2)>(nd, nullptr, A.cashflow, q0, q1, ring,
                                                               pol_first, pol_last);
                else
                    lsmc_ring_prologue<DEG, false, false, (RING > 0 ? RING : 2)>(nd, A.store + (size_t)(date - 2) * A.ld,
                                                                A.cashflow, q0, q1, ring, pol_first,
                                                                pol_last);
                ring_ready = true;
            }
        }
        if (stamp) stamps[(size_t)(date + 1) * 4 + 3] = globaltimer_ns();
        // The next sweep streams row date-2 of the store from HBM.  Ask L2 for this block's
        // slice of it now: the fetch overlaps the block sums, the exchange and the solve
        // (~5 us, about one row at HBM speed), and the sweep then starts on L2 hits.
        // Only while cashflows + three rows fit L2 (the host decides): at 2^22 paths the
        // extra row evicts the cashflows and the launch gets 13 % slower, at 2^20 9 % faster.
        if (A.prefetch_next_row && date >= 2) {
            const char *next = reinterpret_cast<const char *>(
                reinterpret_cast<const double2 *>(A.store + (size_t)(date - 2) * A.ld) + q0);
            const int64_t bytes = (q1 - q0) * (int64_t)sizeof(double2);
            for (int64_t off = (int64_t)threadIdx.x * 128; off < bytes; off += (int64_t)T * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(next + off));
        }
        // block totals: transposed warp reduction, then warp e adds entry e over the warps
        // and its lane 0 publishes it
        {
            double v[P];
#pragma unroll
            for (int k = 0; k < P; ++k) v[k] = k < NV ? acc.a[k] : 0.0;
            warp_transpose_sum<P>(v, lane);
            const int e = P == 32 ? lane : lane >> 1;
            if (e < NV && (P == 32 || (lane & 1) == 0)) scratch[e * 32 + warp] = v[0];
        }
        timed_out = __syncthreads_or(timed_out);
        for (int e = warp; e < NV; e += NW) {
            double t = lane < NW ? scratch[e * 32 + lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0)
                ll_store(A.slots + ((size_t)(epoch & 1u) * G + blockIdx.x) * LSMC_P_STRIDE + e,
                         timed_out ? ll_poison() : t, (A.nonce << 12) | (epoch & 0xFFFu));
        }
        ++epoch;
    }
    if (timed_out) atomicExch(A.error, 1u);
}

// ---- small all-reduce over NVLink peer memory (European moments) ------------------------
// The sum of n (<= PEER_AR_MAX) doubles over the W attached ranks, in place, by the same
// flag-carrying 16-byte slots the induction kernel exchanges its power sums with: thread i
// stores its value into slot [parity][rank][i] of EVERY rank's buffer (peer stores over
// NVLink), polls the W slots [parity][r][i] of the LOCAL buffer and adds them in rank order
// -- every rank ends with bit-identical sums one NVLink hop after the last rank has stored.
// This is the all-reduce that follows every partitioned pricing (48 bytes per
// contract); as a kernel of ours on the pricing's stream it costs one launch instead of a
// library collective (measured, 8 GPUs: ~5 us against 26 us for a 6-double NCCL all-reduce).
constexpr int PEER_AR_MAX = 1536;             // 256 contracts x 6 moments
struct PeerReduceArgs {
    double *vals;
    int n, rank, world;
    uint32_t tag;                             // per call, identical on every rank, never 0
    size_t euro_off;                          // offset (uint4) of the European region in a buffer
    uint4 *peer_rank_slots[OPTMC_MAX_PEERS];
    uint4 *local_slots;                       // = peer_rank_slots[rank] (a dynamic index into the
                                              // parameter array would cost a local-memory copy of it)
    unsigned int *error;
};

__global__ void __launch_bounds__(256) k_peer_allreduce(const PeerReduceArgs A) {
    bool bad = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
        const size_t base = A.euro_off + ((size_t)(A.tag & 1u) * A.world) * PEER_AR_MAX + i;
        const double v = A.vals[i];
        // (loops over the peers unrolled with static indices: the pointer array stays in the
        // constant bank instead of a local-memory copy)
        const uint4 *mine = A.local_slots + base;
#pragma unroll
        for (int r = 0; r < OPTMC_MAX_PEERS; ++r)
            if (r < A.world)
                ll_store_sys(A.peer_rank_slots[r] + base + (size_t)A.rank * PEER_AR_MAX, v, A.tag);
        // first looks at the W slots in flight together: one L2 round trip when the values are
        // there, not W; then the sum in rank order
        uint4 w[OPTMC_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < OPTMC_MAX_PEERS; ++r)
            if (r < A.world) w[r] = ll_peek(mine + (size_t)r * PEER_AR_MAX);
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < OPTMC_MAX_PEERS; ++r) {
            if (r < A.world) {
                double x = 0.0;
                if (!ll_wait(mine + (size_t)r * PEER_AR_MAX, A.tag, w[r], x, bad)) bad = true;
                sum += x;
            }
        }
        A.vals[i] = bad ? ll_poison() : sum;
    }
    if (bad) atomicExch(A.error, 1u);
}

}  // namespace optmc
