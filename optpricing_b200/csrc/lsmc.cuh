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
    __device__ __forceinline__ void operator()(int i, const State &sa, const State &sb) const {
        double *row = store + (size_t)i * ld;
        row[g] = terminal_spot(kind, sa);                       // path_kernels.py:34
        if (anti) row[n_local + g] = terminal_spot(kind, sb);
    }
};

template <int KIND, bool ANTI, int NIMPL>
__global__ void __launch_bounds__(EURO_THREADS) k_lsmc_paths(const EuroArgs A, double *store,
                                                             int64_t ld) {
    __shared__ RngTables tab;
    RngView tv{0u};
    if constexpr (NIMPL != 0) {
        tv = load_rng_tables(&tab);
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        State sa, sb;
        StoreSink sink{store, ld, g, A.n_local, KIND, ANTI};
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

// ---- cashflow = payoff at T (american_mc_kernels.py:48-51) -----------------
__global__ void __launch_bounds__(LSMC_THREADS) k_lsmc_begin(const double *last_row, int64_t n,
                                                             double K, int is_call,
                                                             double *cashflow) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double s = last_row[p];
        cashflow[p] = fmax(is_call ? s - K : K - s, 0.0);
    }
}

// ---- per-date Gram accumulation (american_mc_kernels.py:54-69) -------------
// cashflow *= disc for every path; in-the-money paths add to the power sums.
// Per-block partials; the last block to finish sums them in fixed order
// (deterministic) into gram[OPTMC_LSMC_GRAM].
template <int DEG>
__global__ void __launch_bounds__(LSMC_THREADS)
k_lsmc_gram(const double *row, int64_t n, double K, double inv_K, int is_call, double disc,
            double *cashflow, double *partials, unsigned int *ticket, double *gram) {
    constexpr int NP = 2 * DEG + 1, NQ = DEG + 1, NV = NP + NQ;
    __shared__ double scratch[NV * 32];
    __shared__ bool last;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double y = cashflow[p] * disc;                          // :54
        cashflow[p] = y;
        double s = row[p];
        double iv = is_call ? s - K : K - s;                    // :57-60
        if (iv > 0.0) {                                         // :62
            double x = fma(s, inv_K, -1.0);
            double pw = 1.0;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                acc[k] += pw;
                if (k < NQ) acc[NP + k] = fma(y, pw, acc[NP + k]);
                pw *= x;
            }
        }
    }
    block_sum<NV>(acc, scratch);
    double *mine = partials + (size_t)blockIdx.x * OPTMC_LSMC_GRAM;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) mine[k] = acc[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        // fixed-order sum over blocks: thread k owns entry k
        if (threadIdx.x < NV) {
            double s = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b)
                s += __ldcg(partials + (size_t)b * OPTMC_LSMC_GRAM + threadIdx.x);
            int k = threadIdx.x;
            gram[k < NP ? k : 16 + (k - NP)] = s;
        }
        if (threadIdx.x == 0) *ticket = 0u;
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

// ---- exercise decision (american_mc_kernels.py:72-80) ----------------------
__global__ void __launch_bounds__(LSMC_THREADS)
k_lsmc_decide(const double *row, int64_t n, double K, double inv_K, int is_call, int deg,
              const double *gram, double *cashflow) {
    __shared__ double beta[8];
    __shared__ bool ok;
    if (threadIdx.x == 0) ok = lsmc_solve(gram, deg, beta);
    __syncthreads();
    if (!ok) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double s = row[p];
        double iv = is_call ? s - K : K - s;
        if (iv > 0.0) {
            double x = fma(s, inv_K, -1.0);
            double cont = beta[deg];
            for (int k = deg - 1; k >= 0; --k) cont = fma(cont, x, beta[k]);   // X @ beta, :73
            if (iv > cont) cashflow[p] = iv;                                   // :75-80
        }
    }
}

// ---- final mean (american_mc_kernels.py:84) --------------------------------
__global__ void __launch_bounds__(LSMC_THREADS)
k_lsmc_finish(const double *cashflow, int64_t n, double disc, double *partials) {
    __shared__ double scratch[2 * 32];
    double acc[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double y = cashflow[p] * disc;
        acc[0] += y;
        acc[1] = fma(y, y, acc[1]);
    }
    block_sum<2>(acc, scratch);
    if (threadIdx.x == 0) {
        partials[(size_t)blockIdx.x * NPART + 0] = acc[0];
        partials[(size_t)blockIdx.x * NPART + 1] = acc[1];
    }
}

__global__ void __launch_bounds__(256) k_lsmc_finish2(const double *partials, int n_blocks,
                                                      double n, double *out) {
    __shared__ double scratch[2 * 32];
    double acc[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
        acc[0] += partials[(size_t)b * NPART + 0];
        acc[1] += partials[(size_t)b * NPART + 1];
    }
    block_sum<2>(acc, scratch);
    if (threadIdx.x == 0) {
        out[0] = n;
        out[1] = acc[0];
        out[2] = acc[1];
    }
}

}  // namespace optmc
