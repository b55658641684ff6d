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
struct StoreSink {
    double *store;
    int64_t ld, g, n_local;
    int kind;
    bool anti;
    RngView tab;
    __device__ __forceinline__ void operator()(int i, const State &sa, const State &sb) const {
        double *row = store + (size_t)i * ld;
        row[g] = step_spot(kind, sa, tab);                      // path_kernels.py:34
        if (anti) row[n_local + g] = step_spot(kind, sb, tab);
    }
};

template <int KIND, bool ANTI, int NIMPL>
__global__ void __launch_bounds__(EURO_THREADS) k_lsmc_paths(const EuroArgs A, double *store,
                                                             int64_t ld) {
    __shared__ RngShared<kind_two_factor(KIND)> tab;
    RngView tv{0u};
    tv = fill_rng_shared(&tab, kind_two_factor(KIND) ? A.c.m11 : A.c.sqrt_dt, A.c.m12, A.c.m21,
                         A.c.m22);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        State sa, sb;
        StoreSink sink{store, ld, g, A.n_local, KIND, ANTI, tv};
        simulate_draw<KIND, ANTI, NIMPL, true>(A, (uint64_t)(A.draw_offset + g), sa, sb, tv,
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
__device__ inline bool lsmc_solve(const double *gram, int deg, double *beta) {
    const int m = deg + 1;
    const double *P = gram;        // P[k] = sum x^k (P[0] = n_itm)
    const double *Q = gram + 16;   // Q[k] = sum y x^k
    if (!(P[0] >= (double)m)) return false;
    double L[8][8], d[8], rhs[8];
    for (int a = 0; a < m; ++a) {
        double diag = P[2 * a];
        if (!(diag > 0.0)) return false;
        d[a] = 1.0 / sqrt(diag);
    }
    for (int a = 0; a < m; ++a) {
        rhs[a] = Q[a] * d[a];
        for (int b = 0; b <= a; ++b) L[a][b] = P[a + b] * d[a] * d[b];
    }
    for (int j = 0; j < m; ++j) {
        double s = L[j][j];
        for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
        if (!(s > 1e-13)) return false;
        double lj = sqrt(s);
        L[j][j] = lj;
        for (int i = j + 1; i < m; ++i) {
            double t = L[i][j];
            for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
            L[i][j] = t / lj;
        }
    }
    for (int i = 0; i < m; ++i) {
        double t = rhs[i];
        for (int k = 0; k < i; ++k) t -= L[i][k] * rhs[k];
        rhs[i] = t / L[i][i];
    }
    for (int i = m - 1; i >= 0; --i) {
        double t = rhs[i];
        for (int k = i + 1; k < m; ++k) t -= L[k][i] * rhs[k];
        rhs[i] = t / L[i][i];
    }
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

// One path of the fused sweep (see k_lsmc_step).  Returns the discounted
// cashflow to store.
template <int DEG, bool INIT, bool FINAL>
__device__ __forceinline__ double lsmc_path(double y_in, double s_dec, double s_acc, bool use_fit,
                                            const double *beta, double K, double inv_K, bool call,
                                            double disc, LsmcAcc<DEG> &acc) {
    double y;
    if constexpr (INIT) {
        y = fmax(call ? s_dec - K : K - s_dec, 0.0);                 // american_mc_kernels.py:48-51
    } else {
        y = y_in;
        if (use_fit) {
            double iv = call ? s_dec - K : K - s_dec;
            if (iv > 0.0) {                                           // :62
                double x = fma(s_dec, inv_K, -1.0);
                double cont = beta[DEG];
#pragma unroll
                for (int k = DEG - 1; k >= 0; --k) cont = fma(cont, x, beta[k]);   // X @ beta, :73
                if (iv > cont) y = iv;                                // :75-80
            }
        }
    }
    y *= disc;                                                        // :54 / :84
    if constexpr (FINAL) {
        acc.a[0] += 1.0;
        acc.a[1] += y;
        acc.a[2] = fma(y, y, acc.a[2]);
    } else {
        double iv = call ? s_acc - K : K - s_acc;                     // :57-60
        if (iv > 0.0) {
            double x = fma(s_acc, inv_K, -1.0);
            double pw = 1.0;
#pragma unroll
            for (int k = 0; k < LsmcAcc<DEG>::NP; ++k) {
                acc.a[k] += pw;
                if (k < LsmcAcc<DEG>::NQ)
                    acc.a[LsmcAcc<DEG>::NP + k] = fma(y, pw, acc.a[LsmcAcc<DEG>::NP + k]);
                pw *= x;
            }
        }
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
    if (threadIdx.x == 0) ok = INIT ? false : lsmc_solve(gram_in, DEG, beta);
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

}  // namespace optmc
