// levy.cuh -- terminal-value samplers of the pure-Levy and exact-sampler models.
//
// Replaces the host samplers behind MonteCarloTechnique.price's single-step
// branches (monte_carlo.py:103-106, 259-280):
//   VG          models/vg.py:136-168         theta g + sigma sqrt(g) z,  g ~ Gamma(T/nu, scale nu)
//   NIG         models/nig.py:136-169        beta w + sqrt(w) z,         w ~ invgauss(mu_ig)  (lambda = 1)
//   CGMY (Y=1)  models/cgmy.py:134-178       Gamma(CT)/M - Gamma(CT)/G
//   Hyperbolic  models/hyperbolic.py:123-152 loc + scale (b v + sqrt(v) z), v ~ GIG(p=1, sqrt(a^2-b^2)) / sqrt(a^2-b^2)
//   CEV         models/cev.py:49-96          (ncx2(df, nc) / k)^(1 / (2 (1 - gamma)))
// followed by S_T = S0 exp(drift + x) (monte_carlo.py:276-280) for the Levy
// kinds.  One thread = one sample; every draw is a pure function of
// (seed, sample id, attempt, stream tag), so results do not depend on the launch
// geometry or on how samples are split across ranks.  Single-step kernels: their
// cost is a few hundred fp64 instructions per sample, microseconds per million
// samples, so they use the CUDA math library rather than the table-driven
// helpers of the step loops.
#pragma once
#include "european.cuh"

namespace optmc {

struct LevyArgs {
    int kind;
    double p[OPTMC_LEVY_NPARAM];
    PhiloxKey key;
    int64_t offset, n_local;
    double s0, drift;
    int is_call;
    double strike;
    double *partials;   // [gridDim.x][NPART]
    double *terminal;   // optional [n_local]
};

// Philox stream tags (counter word c3) of the samplers; c2 carries the attempt.
enum : uint32_t {
    LEVY_TAG_NORMAL = 0x20u, LEVY_TAG_GAMMA_A = 0x21u, LEVY_TAG_GAMMA_B = 0x22u,
    LEVY_TAG_BOOST_A = 0x23u, LEVY_TAG_BOOST_B = 0x24u, LEVY_TAG_IG = 0x25u, LEVY_TAG_GIG = 0x26u
};

struct LevyRng {
    const PhiloxKey &key;
    uint32_t id_lo, id_hi;
    __device__ __forceinline__ uint4 draw(uint32_t attempt, uint32_t tag) const {
        return philox4x32_10(make_uint4(id_lo, id_hi, attempt, tag), key);
    }
};

__device__ __forceinline__ double u01_32(uint32_t w) { return ((double)w + 0.5) * 0x1.0p-32; }

// Gamma(shape, 1): Marsaglia & Tsang (2000) squeeze-free form; shape < 1 by the
// boost Gamma(a) = Gamma(a + 1) U^(1/a).  d, c belong to max(shape, shape + 1).
__device__ inline double sample_gamma(const LevyRng &r, double shape, uint32_t tag,
                                      uint32_t boost_tag) {
    const bool boost = shape < 1.0;
    const double a = boost ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (uint32_t attempt = 0; attempt < 4096u; ++attempt) {
        uint4 w = r.draw(attempt, tag);
        double z, z_unused;
        normal_pair_lib(w.x, w.y, w.z, z, z_unused);
        double t = fma(c, z, 1.0);
        if (t <= 0.0) continue;
        double v = t * t * t;
        double u = u01_32(w.w);
        if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) {
            g = d * v;
            break;
        }
    }
    if (boost) {
        uint4 w = r.draw(0u, boost_tag);
        g *= pow(u01_open(w.x, w.y), 1.0 / shape);
    }
    return g;
}

// invgauss(mu) of scipy: inverse Gaussian with mean mu and shape 1
// (Michael, Schucany & Haas 1976).
__device__ inline double sample_invgauss(const LevyRng &r, double mu) {
    uint4 w = r.draw(0u, LEVY_TAG_IG);
    double z, z_unused;
    normal_pair_lib(w.x, w.y, w.z, z, z_unused);
    double y = z * z;
    double x = mu + 0.5 * mu * mu * y - 0.5 * mu * sqrt(4.0 * mu * y + mu * mu * y * y);
    double u = u01_32(w.w);
    return (u * (mu + x) <= mu) ? x : mu * mu / x;
}

// geninvgauss(p, b): density ~ x^(p-1) exp(-b (x + 1/x) / 2).  Ratio of uniforms
// with mode shift (Hoermann & Leydold 2014, the p >= 1 or b > 1 branch scipy
// takes): the bounding rectangle [um, up] x [0, 1] around the mode m comes from
// the host (levy_gig_setup in optmc.cu).
__device__ inline double sample_gig(const LevyRng &r, double pm1, double b, double m, double um,
                                    double up, double log_gm) {
    for (uint32_t attempt = 0; attempt < 4096u; ++attempt) {
        uint4 w = r.draw(attempt, LEVY_TAG_GIG);
        double v = u01_open(w.x, w.y);
        double u = fma(up - um, u01_32(w.z), um);
        double x = u / v + m;
        if (x <= 0.0) continue;
        double logg = pm1 * log(x) - 0.5 * b * (x + 1.0 / x) - log_gm;
        if (2.0 * log(v) <= logg) return x;
    }
    return m;
}

__device__ __forceinline__ double sample_normal(const LevyRng &r) {
    uint4 w = r.draw(0u, LEVY_TAG_NORMAL);
    double z, z_unused;
    normal_pair_lib(w.x, w.y, w.z, z, z_unused);
    return z;
}

// Terminal spot of sample `id`.
__device__ inline double levy_terminal_spot(const LevyArgs &A, uint64_t id) {
    LevyRng r{A.key, (uint32_t)id, (uint32_t)(id >> 32)};
    const double *p = A.p;
    switch (A.kind) {
        case OPTMC_LEVY_VG: {           // p: shape T/nu, scale nu, theta, sigma
            double g = p[1] * sample_gamma(r, p[0], LEVY_TAG_GAMMA_A, LEVY_TAG_BOOST_A);
            double x = fma(p[2], g, p[3] * sqrt(g) * sample_normal(r));
            return A.s0 * exp(A.drift + x);
        }
        case OPTMC_LEVY_NIG: {          // p: mu_ig, beta
            double w = sample_invgauss(r, p[0]);
            double x = fma(p[1], w, sqrt(w) * sample_normal(r));
            return A.s0 * exp(A.drift + x);
        }
        case OPTMC_LEVY_CGMY1: {        // p: shape C T, 1/M, 1/G
            double up = sample_gamma(r, p[0], LEVY_TAG_GAMMA_A, LEVY_TAG_BOOST_A) * p[1];
            double dn = sample_gamma(r, p[0], LEVY_TAG_GAMMA_B, LEVY_TAG_BOOST_B) * p[2];
            return A.s0 * exp(A.drift + (up - dn));
        }
        case OPTMC_LEVY_HYPERBOLIC: {   // p: b_gig, gig scale, beta, loc, scale, p-1, m, um, up, log g(m)
            double v = p[1] * sample_gig(r, p[5], p[0], p[6], p[7], p[8], p[9]);
            double z = fma(p[2], v, sqrt(v) * sample_normal(r));
            return A.s0 * exp(A.drift + fma(p[4], z, p[3]));
        }
        case OPTMC_LEVY_CEV: {          // p: df, nc, k, 1 / (2 (1 - gamma))
            // ncx2(df, nc) = (z + sqrt(nc))^2 + chi2(df - 1),  df > 1
            double z = sample_normal(r) + sqrt(p[1]);
            double chi = 2.0 * sample_gamma(r, 0.5 * (p[0] - 1.0), LEVY_TAG_GAMMA_A, LEVY_TAG_BOOST_A);
            return pow(fma(z, z, chi) / p[2], p[3]);
        }
        default: {                      // OPTMC_LEVY_LOGNORMAL (CEV with gamma -> 1): p: drift, sigma sqrt(T)
            return A.s0 * exp(fma(p[1], sample_normal(r), p[0]));
        }
    }
}

__global__ void __launch_bounds__(256) k_levy_terminal(const LevyArgs A) {
    __shared__ double scratch[5 * 32];
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_local; g += stride) {
        double ST = levy_terminal_spot(A, (uint64_t)(A.offset + g));
        if (A.terminal) A.terminal[g] = ST;
        double pay = fmax(A.is_call ? ST - A.strike : A.strike - ST, 0.0);   // monte_carlo.py:110-114
        acc[0] += pay;
        acc[1] = fma(pay, pay, acc[1]);
        acc[2] += ST;
        acc[3] = fma(ST, ST, acc[3]);
        acc[4] = fma(pay, ST, acc[4]);
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        double *q = A.partials + (size_t)blockIdx.x * NPART;
#pragma unroll
        for (int k = 0; k < 5; ++k) q[k] = acc[k];
    }
}

}  // namespace optmc
