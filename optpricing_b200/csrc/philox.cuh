// philox.cuh -- counter-based RNG for the path kernels.
//
// Replaces the reference's host-side draws: rng.standard_normal / rng.poisson in
// MonteCarloTechnique._simulate_sde_path (monte_carlo.py:222-240) and the numba
// global-RNG jump sizes (mc_kernels.py:111,170,269,314-316).  Nothing is
// materialised: every draw is a pure function of (seed, draw id, step, stream).
//
// Philox4x32-10 (Salmon et al., SC'11).  Counter layout used by the engine:
//   c0, c1 = draw id (64 bit)        c2 = step index       c3 = stream tag
// Stream tags (c3): 0 = Brownian pair + Poisson word of the step;
//   (1 + partner) | (jump index << 8) = jump-size draws of that partner at the step.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace optmc {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
        uint32_t hi1 = __umulhi(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
#else
        uint64_t p0 = (uint64_t)PHILOX_M0 * c.x, p1 = (uint64_t)PHILOX_M1 * c.z;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    return c;
}

__device__ __forceinline__ uint4 philox_draw(uint64_t seed, uint64_t id, uint32_t c2,
                                             uint32_t c3) {
    return philox4x32_10(make_uint4((uint32_t)id, (uint32_t)(id >> 32), c2, c3),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 64 random bits -> uniform double in (0, 1), 53 significant bits, never 0 or 1.
__device__ __forceinline__ double u01_open(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * 0x1.0p-53;
}

// sqrt for the step loops: MUFU.RSQ64H seed (2^-22) + two FMA-form Newton
// corrections = 6 fp64-pipe instructions, no branches, exact 0 for a == 0.
// a must be >= 0 (denormals are treated as 1e-290-ish; results stay < 1e-145).
__device__ __forceinline__ double sqrt_nonneg(double a) {
    int hi = __double2hiint(a);
    hi = max(hi, 0x03000000);  // keep the seed finite for a == 0 / tiny a
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(hi, 0)));
    double g = a * y;
    double h = 0.5 * y;
    double e = fma(-g, g, a);
    g = fma(e, h, g);
    e = fma(-g, g, a);
    return fma(e, h, g);
}

// ---------------------------------------------------------------------------
// Two N(0,1) from three Philox words (Box-Muller): radius from 64 bits, angle
// from 32.  IMPL 0 = CUDA math library calls (baseline, kept for A/B checks);
// IMPL 1 = hand-written for the fp64 pipe.
// ---------------------------------------------------------------------------
struct LogTable {        // IMPL 1: ln on [1,2) by 128 buckets, 2 KB of shared memory
    double inv_c[128];   // 1 / bucket midpoint
    double ln_c[128];    // ln(bucket midpoint)
};

__device__ __forceinline__ void fill_log_table(LogTable *t) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        double c = 1.0 + (i + 0.5) * (1.0 / 128.0);
        t->inv_c[i] = 1.0 / c;
        t->ln_c[i] = log(c);
    }
}

// -2 ln u, u = (x + .5) / 2^64 in (0,1) for the 64-bit word x = hi:lo.
//   u = 2^-(k+1) * m, m in [1,2): k = leading zeros (geometric -- exactly the
//   exponent law of a uniform), m = the next 52 bits.  ln m = ln c_i + ln(1+r),
//   r = m / c_i - 1, |r| <= 2^-8, degree-7 series.  No special cases reachable.
__device__ __forceinline__ double neg2_log_u(uint32_t lo, uint32_t hi, const LogTable *t) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    int k = min(__clzll((long long)x), 63);
    uint64_t frac = (x << k) << 1;  // drop the leading one (k <= 63, so no UB)
    uint64_t mbits = 0x3FF0000000000000ull | (frac >> 12);
    double m = __longlong_as_double((long long)mbits);
    int idx = (int)((mbits >> 45) & 127);
    double r = fma(m, t->inv_c[idx], -1.0);
    double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(p, r, 1.0 / 5.0);
    p = fma(p, r, -1.0 / 4.0);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    double ln_m = fma(p, r, t->ln_c[idx]);
    return fma((double)(k + 1), 1.3862943611198906188, -2.0 * ln_m);  // 2 ln2 (k+1) - 2 ln m
}

template <int IMPL>
__device__ __forceinline__ void normal_pair(uint32_t w0, uint32_t w1, uint32_t w2, double &z1,
                                            double &z2, const LogTable *t) {
    if constexpr (IMPL == 0) {
        double rad = sqrt(-2.0 * log(u01_open(w0, w1)));
        double s, c;
        sincospi(((double)w2 + 0.5) * 0x1.0p-31, &s, &c);  // angle = 2 pi (w2 + .5) / 2^32
        z1 = rad * c;
        z2 = rad * s;
    } else {
        double rad = sqrt_nonneg(neg2_log_u(w0, w1, t));
        // angle: 29 bits place x in (0, pi/4); the top 3 bits choose among the
        // 8 images of that octant under swap / sign symmetries -- together a
        // uniform angle on the circle.
        double x = (double)((w2 & 0x1FFFFFFFu) * 2u + 1u) * (0.78539816339744830962 * 0x1.0p-30);
        double x2 = x * x;
        double sp = fma(x2, -1.0 / 1307674368000.0, 1.0 / 6227020800.0);
        sp = fma(sp, x2, -1.0 / 39916800.0);
        sp = fma(sp, x2, 1.0 / 362880.0);
        sp = fma(sp, x2, -1.0 / 5040.0);
        sp = fma(sp, x2, 1.0 / 120.0);
        sp = fma(sp, x2, -1.0 / 6.0);
        double sx = fma(sp * x2, x, x);
        double cp = fma(x2, 1.0 / 20922789888000.0, -1.0 / 87178291200.0);
        cp = fma(cp, x2, 1.0 / 479001600.0);
        cp = fma(cp, x2, -1.0 / 3628800.0);
        cp = fma(cp, x2, 1.0 / 40320.0);
        cp = fma(cp, x2, -1.0 / 720.0);
        cp = fma(cp, x2, 1.0 / 24.0);
        cp = fma(cp, x2, -0.5);
        double cx = fma(cp, x2, 1.0);
        bool swap = (w2 >> 29) & 1u;
        double a = swap ? sx : cx;
        double b = swap ? cx : sx;
        // signs: flip the top bit of the high word (integer pipe)
        int ahi = __double2hiint(a) ^ (int)((w2 << 1) & 0x80000000u);
        int bhi = __double2hiint(b) ^ (int)(w2 & 0x80000000u);
        z1 = rad * __hiloint2double(ahi, __double2loint(a));
        z2 = rad * __hiloint2double(bhi, __double2loint(b));
    }
}

// Poisson(mu) by inversion, mu small (lambda * dt).  `w` is the 32-bit word of
// the step's Philox call; p0_bits = floor(exp(-mu) * 2^32).  The common case
// (no jump) is one integer compare; otherwise 32 more bits from a second call
// refine the uniform to 64 bits and the CDF is walked in fp64.
__device__ __forceinline__ bool poisson_maybe(uint32_t w, uint32_t p0_bits) { return w >= p0_bits; }

__device__ __forceinline__ int poisson_slow(uint32_t w, uint32_t extra, double mu, double p0) {
    double u = ((double)(((uint64_t)w << 32 | extra) >> 11) + 0.5) * 0x1.0p-53;
    double p = p0, cdf = p0;
    int n = 0;
    while (u > cdf && n < 64) {
        ++n;
        p *= mu / (double)n;
        cdf += p;
    }
    return n;
}

}  // namespace optmc
