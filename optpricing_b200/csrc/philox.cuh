// philox.cuh -- counter-based RNG for the path kernels.
//
// Replaces the reference's host-side draws: rng.standard_normal / rng.poisson in
// MonteCarloTechnique._simulate_sde_path (monte_carlo.py:222-240) and the numba
// global-RNG jump sizes (mc_kernels.py:111,170,269,314-316).  Nothing is
// materialised: every draw is a pure function of (seed, draw id, step, stream).
//
// Philox4x32-10 (Salmon et al., SC'11).  Counter layout used by the engine:
//   c0, c1 = draw id (64 bit)        c2 = step index       c3 = stream tag
// Stream tags (c3): 0 = Brownian pair + Poisson word of the step; 0xFF = extra
//   Poisson bits; (1 + partner) | (jump index << 8) = jump-size draws.
//
// Instruction budget (B200: the fp64 pipe AND the integer ALU pipe are both
// half-rate, 16 lanes per SM sub-partition; ncu on the first version of the
// Heston kernel showed fp64 42 % / ALU 41 % busy with issue slots the binding
// resource).  Hence: round keys are expanded once on the host (no per-round key
// adds), polynomial coefficients live in the constant bank (no UMOV pairs), the
// uniform's exponent comes from 12 dedicated bits (no 64-bit shifts), and the
// transcendental parts use two 2 KB shared-memory tables so that each needs
// only a short polynomial.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace optmc {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

// The ten round keys of one seed, expanded on the host.
struct PhiloxKey {
    uint32_t k0[10], k1[10];
};

inline void philox_expand_key(uint64_t seed, PhiloxKey *k) {
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
        k->k0[i] = a;
        k->k1[i] = b;
        a += PHILOX_W0;
        b += PHILOX_W1;
    }
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKey &k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
        uint32_t hi1 = __umulhi(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.k0[i], lo1, hi0 ^ c.w ^ k.k1[i], lo0);
    }
    return c;
}

__device__ __forceinline__ uint4 philox_draw(const PhiloxKey &k, uint64_t id, uint32_t c2,
                                             uint32_t c3) {
    return philox4x32_10(make_uint4((uint32_t)id, (uint32_t)(id >> 32), c2, c3), k);
}

// 64 random bits -> uniform double in (0, 1), 53 significant bits, never 0 or 1.
__device__ __forceinline__ double u01_open(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * 0x1.0p-53;
}

// MUFU.RSQ64H seed for a >= 0: y ~ 1/sqrt(a) to ~2^-22 relative, finite (huge)
// for a == 0.  Also returns h = y/2, formed by decrementing the exponent field
// on the integer pipe (y's low word is zero, so this is exact and costs no
// fp64-pipe slot).
__device__ __forceinline__ void rsqrt_seed(double a, double &y, double &h) {
    int hi = max(__double2hiint(a), 0x03000000);
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(hi, 0)));
    h = __hiloint2double(__double2hiint(y) - 0x00100000, 0);
}

// sqrt for the step loops: seed + two FMA-form Newton corrections = 5 fp64-pipe
// instructions, no branches, exact 0 for a == 0, < 1 ulp otherwise.
__device__ __forceinline__ double sqrt_nonneg(double a) {
    double y, h;
    rsqrt_seed(a, y, h);
    double g = a * y;
    double e = fma(-g, g, a);
    g = fma(e, h, g);
    e = fma(-g, g, a);
    return fma(e, h, g);
}

// One correction only (relative error <= 1.5 * 2^-44 = 9e-14): used for the
// Box-Muller radius, where that is far below anything a price can see.
__device__ __forceinline__ double sqrt_nonneg_fast(double a) {
    double y, h;
    rsqrt_seed(a, y, h);
    double g = a * y;
    double e = fma(-g, g, a);
    return fma(e, h, g);
}

// index of the highest set bit (x != 0): one FLO instruction
__device__ __forceinline__ uint32_t top_bit(uint32_t x) {
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

// Launder a value through an opaque move so the compiler keeps it in a register
// instead of re-deriving it inside the step loop.
__device__ __forceinline__ uint32_t keep_u32(uint32_t x) {
    asm volatile("mov.u32 %0, %0;" : "+r"(x));
    return x;
}

// ---------------------------------------------------------------------------
// Two N(0,1) from three Philox words (Box-Muller): radius from 64 bits, angle
// from 32.  IMPL 0 = CUDA math library calls (baseline, kept for A/B checks);
// IMPL 1 = hand-written for the fp64 pipe.
// ---------------------------------------------------------------------------
constexpr int TRIG_N = 1024;   // angle buckets over the full circle
struct RngTables {
    // ln on [1,2): bucket i has midpoint c_i = 1 + (i + .5)/128;
    //   x = -2 / c_i,  y = -2 ln c_i + 24 ln 2   (the 24 ln 2 pre-pays the exponent offset)
    double2 logt[128];
    // angle bucket i of [0, 2 pi): a_i = (i + .5) 2 pi / 1024;  x = cos a_i, y = sin a_i
    double2 trig[TRIG_N];
};
constexpr int RNG_TABLE_VEC = (int)(sizeof(RngTables) / sizeof(double2));

__device__ RngTables g_rng_tables;   // filled once per context from the host (optmc.cu)

// Shared-memory copy of the tables, addressed by a 32-bit shared-window offset so
// that lookups are a single LDS.128 with no generic-address arithmetic.
struct RngView {
    uint32_t base;   // shared address of RngTables
};

__device__ __forceinline__ RngView load_rng_tables(RngTables *t) {
    const double2 *src = reinterpret_cast<const double2 *>(&g_rng_tables);
    double2 *dst = reinterpret_cast<double2 *>(t);
    for (int i = threadIdx.x; i < RNG_TABLE_VEC; i += blockDim.x) dst[i] = src[i];
    RngView v;
    v.base = keep_u32((uint32_t)__cvta_generic_to_shared(t));
    return v;
}

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
    return r;
}

// Polynomial coefficients, read straight from the constant bank.
//   -2 ln(1 + r) with r = -r'/2:  r' + r'^2/4 + r'^3/12 + r'^4/32 + r'^5/80 + r'^6/192
__constant__ double C_LOG[5] = {1.0 / 192.0, 1.0 / 80.0, 1.0 / 32.0, 1.0 / 12.0, 0.25};
__constant__ double C_N2LN2 = -1.3862943611198906188;        // -2 ln 2
__constant__ double C_TRIG[4] = {1.0 / 24.0, -0.5, 1.0 / 120.0, -1.0 / 6.0};
// offset inside an angle bucket: d = (j + .5) h / 2^22 - h/2, j = 22 random bits, h = 2 pi / 1024
__constant__ double C_FINE[2] = {6.283185307179586477 / 1024.0 / 4194304.0,
                                 (0.5 / 4194304.0 - 0.5) * (6.283185307179586477 / 1024.0)};

// -2 ln u for a uniform u in (0,1) built from the words (lo, hi):
//   hi[31:20] exponent bits: k = their leading zeros, u in [2^-(k+1), 2^-k)
//   hi[19:0] : lo            the 52 mantissa bits of m in [1,2);  u = m 2^-(k+1)
// If all 12 exponent bits are zero (p = 2^-12) the exponent continues into the
// mantissa bits (slow path), so u stays exactly uniform on its 2^-64 grid.
__device__ __forceinline__ double neg2_log_poly(double m, double2 tb, double kfac) {
    double r = fma(m, tb.x, 2.0);                  // r' = -2 (m / c - 1), |r'| <= 2^-7
    double p = fma(r, C_LOG[0], C_LOG[1]);
    p = fma(p, r, C_LOG[2]);
    p = fma(p, r, C_LOG[3]);
    p = fma(p, r, C_LOG[4]);
    p = fma(p, r, 1.0);
    // -2 ln u = 2 (k+1) ln2 - 2 ln m;  tb.y carries 24 ln2, kfac = 11 - k (top set bit index)
    return fma(p, r, fma(kfac, C_N2LN2, tb.y));
}

__device__ __noinline__ double neg2_log_u_tail(uint32_t lo, uint32_t hi, RngView t) {
    // hi < 2^20: the remaining 52 bits x = hi:lo carry exponent and mantissa
    uint64_t x = ((uint64_t)hi << 32) | lo;        // < 2^52
    int k = x ? __clzll((long long)x) - 12 : 51;   // leading zeros within the 52 bits
    k = k > 51 ? 51 : k;
    uint64_t frac = ((x << k) << 1) & 0xFFFFFFFFFFFFFull;
    uint64_t mbits = 0x3FF0000000000000ull | frac;
    double m = __longlong_as_double((long long)mbits);
    double2 tb = lds_f64x2(t.base + (uint32_t)((mbits >> 45) & 127) * 16u);
    return neg2_log_poly(m, tb, (double)(11 - (12 + k)));
}

__device__ __forceinline__ double neg2_log_u(uint32_t lo, uint32_t hi, RngView t) {
    uint32_t e = hi >> 20;
    if (__builtin_expect(e == 0u, 0)) return neg2_log_u_tail(lo, hi, t);
    double m = __hiloint2double(0x3FF00000 | (hi & 0xFFFFF), lo);
    double2 tb = lds_f64x2(t.base + ((hi >> 9) & 0x7F0u));
    return neg2_log_poly(m, tb, (double)top_bit(e));
}

template <int IMPL>
__device__ __forceinline__ void normal_pair(uint32_t w0, uint32_t w1, uint32_t w2, double &z1,
                                            double &z2, RngView t) {
    if constexpr (IMPL == 0) {
        double rad = sqrt(-2.0 * log(u01_open(w0, w1)));
        double s, c;
        sincospi(((double)w2 + 0.5) * 0x1.0p-31, &s, &c);  // angle = 2 pi (w2 + .5) / 2^32
        z1 = rad * c;
        z2 = rad * s;
    } else {
        double rad = sqrt_nonneg_fast(neg2_log_u(w0, w1, t));
        // angle = a_i + d: the top 10 bits pick one of 1024 buckets of the circle
        // (cos, sin of its midpoint from the table), the low 22 bits the offset d
        // in (-h/2, h/2); rotate the table entry by d.
        double2 cs = lds_f64x2(t.base + (uint32_t)sizeof(double2) * 128u + ((w2 >> 18) & 0x3FF0u));
        double d = fma((double)(w2 & 0x3FFFFFu), C_FINE[0], C_FINE[1]);
        double d2 = d * d;
        double cd = fma(fma(d2, C_TRIG[0], C_TRIG[1]), d2, 1.0);       // cos d
        double sd = fma(fma(d2, C_TRIG[2], C_TRIG[3]), d2, 1.0) * d;   // sin d
        double cx = fma(cs.x, cd, -(cs.y * sd));
        double sx = fma(cs.y, cd, cs.x * sd);
        z1 = rad * cx;
        z2 = rad * sx;
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
