// philox.cuh -- counter-based RNG for the path kernels.
//
// Replaces the reference's host-side draws: rng.standard_normal / rng.poisson in
// MonteCarloTechnique._simulate_sde_path (monte_carlo.py:222-240) and the numba
// global-RNG jump sizes (mc_kernels.py:111,170,269,314-316).  Nothing is
// materialised: every draw is a pure function of (seed, draw id, step, stream).
//
// Philox4x32-10 (Salmon et al., SC'11).  Counter layout used by the engine:
//   c0, c1 = draw id (64 bit)        c2 = call index       c3 = stream tag
// Stream tags (c3): 0 = Brownian draws -- packed (philox_pair_words: calls 3g..3g+2
//   feed pairs 4g..4g+3) where the step loop needs no Poisson word, else one call
//   per step (three words for the pair, the fourth for the step's Poisson draw);
//   0xFE = whole-path Poisson draw; 0xFF = extra Poisson bits;
//   (1 + partner) | (jump index << 8) = jump-size draws; 0x40 = fp32-mode pairs;
//   0x20.. = single-step samplers (levy.cuh).
//
// Instruction budget (B200: the fp64 pipe AND the integer ALU pipe are both
// half-rate, 16 lanes per SM sub-partition; ncu on the first version of the
// Heston kernel showed fp64 42 % / ALU 41 % busy with issue slots the binding
// resource).  Hence: round keys are expanded once on the host (no per-round key
// adds), polynomial coefficients live in the constant bank (no UMOV pairs), the
// uniform's exponent comes from 12 dedicated bits (no 64-bit shifts), and the
// transcendental parts use shared-memory tables (512 x 16 B for the logarithm,
// 1024 x 16 or 32 B for the angle) so that each needs only a short polynomial.
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

#ifndef OPTMC_PHILOX_ROUNDS
#define OPTMC_PHILOX_ROUNDS 10      // experiments only (tools/): the engine ships Philox4x32-10
#endif
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKey &k) {
#pragma unroll
    for (int i = 0; i < OPTMC_PHILOX_ROUNDS; ++i) {
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

// The three words of Box-Muller pair `pair` of one draw under the packed layout
// the step loops use (european.cuh): calls c2 = 3g, 3g+1, 3g+2 of stream 0 yield
// twelve words = four pairs,
//   pair 4g: a.x a.y a.z   4g+1: a.w b.x b.y   4g+2: b.z b.w d.x   4g+3: d.y d.z d.w
// (first two words: the 64-bit radius uniform, third: the angle).
__device__ __forceinline__ void philox_pair_words(const PhiloxKey &k, uint64_t id, uint32_t pair,
                                                  uint32_t &w0, uint32_t &w1, uint32_t &w2) {
    const uint32_t g = pair >> 2, slot = pair & 3u;
    const uint4 a = philox_draw(k, id, 3u * g, 0u), b = philox_draw(k, id, 3u * g + 1u, 0u),
                d = philox_draw(k, id, 3u * g + 2u, 0u);
    switch (slot) {
        case 0: w0 = a.x; w1 = a.y; w2 = a.z; break;
        case 1: w0 = a.w; w1 = b.x; w2 = b.y; break;
        case 2: w0 = b.z; w1 = b.w; w2 = d.x; break;
        default: w0 = d.y; w1 = d.z; w2 = d.w;
    }
}

// 64 random bits -> uniform double in (0, 1), 53 significant bits, never 0 or 1.
__device__ __forceinline__ double u01_open(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * 0x1.0p-53;
}

// max(a, 0) as ONE integer-pipe instruction on the high word (fp64 max is DSETP + two
// selects here).  A negative a becomes a non-negative number below 2^-1022 instead of an
// exact 0 -- a difference no product or sum in a step loop can turn into anything
// representable next to the quantities it meets.
__device__ __forceinline__ double relu64(double a) {
    return __hiloint2double(max(__double2hiint(a), 0), __double2loint(a));
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

// One correction only: the seed is good to 2^-20 (measured, profiles/r01_sqrt_probe.log),
// so the result to 1.5 * 2^-40 = 1.4e-12 relative (measured 1.3e-12).  Used where the root
// scales a N(0,1) draw -- the Box-Muller radius, Heston's sqrt(v+ * (-2 ln u)) -- where
// that is far below anything a price can see.
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
// from 32.  IMPL 0 = CUDA math library calls (baseline, kept for A/B checks and
// used on the rare jump-size path); IMPL 1 = hand-written for the fp64 pipe.
// ---------------------------------------------------------------------------
constexpr int LOG_N = 512;     // buckets of [1,2) for ln
constexpr int TRIG_N = 1024;   // angle buckets over the full circle
constexpr int EXP_N = 512;     // 2^(j/512) for exp

// Host-filled base tables (one copy per context, in global memory).
struct RngBase {
    // bucket i has midpoint c_i = 1 + (i + .5)/512;
    //   x = -2 / c_i,  y = -2 ln c_i + 24 ln 2   (the 24 ln 2 pre-pays the exponent offset)
    double2 logt[LOG_N];
    // 2^(j/512), j = 0..511
    double exp2t[EXP_N];
    // angle bucket i of [0, 2 pi): a_i = (i + .5) 2 pi / 1024;  x = cos a_i, y = sin a_i
    double2 trig[TRIG_N];
};
__device__ RngBase g_rng_base;

// Shared-memory layout common to every kernel that uses the math helpers
// (fast_log / fast_exp) -- the RNG tables extend it.
struct MathShared {
    double2 logt[LOG_N];
    double exp2t[EXP_N];
};
constexpr uint32_t EXP_OFF = (uint32_t)sizeof(double2) * LOG_N;
constexpr uint32_t TRIG_OFF = EXP_OFF + (uint32_t)EXP_N * (uint32_t)sizeof(double);

// Per-block shared-memory tables.  The trig rows are specialised per launch:
//   MIXED = false: (s cos a_i, s sin a_i) with a scale s (sqrt(dt) for the
//                  one-factor models, 1 for plain normals);
//   MIXED = true : the 2x2 mixing matrix M of the two-factor models applied to
//                  the rotation basis, (A1, B1, A2, B2) with
//                  A_r = M_r1 cos a_i + M_r2 sin a_i,  B_r = M_r2 cos a_i - M_r1 sin a_i,
//                  so that  M (cos(a_i + d), sin(a_i + d))^T = (A_r cos d + B_r sin d)_r
//                  and the correlated, scaled increments need no further arithmetic.
template <bool MIXED>
struct RngShared {
    MathShared math;
    double2 trig[TRIG_N * (MIXED ? 2 : 1)];
};

// 32-bit shared-window address of the tables: lookups are one LDS.128 each.
struct RngView {
    uint32_t base;
};

__device__ __forceinline__ RngView fill_math_shared(MathShared *t) {
    for (int i = threadIdx.x; i < LOG_N; i += blockDim.x) t->logt[i] = g_rng_base.logt[i];
    for (int i = threadIdx.x; i < EXP_N; i += blockDim.x) t->exp2t[i] = g_rng_base.exp2t[i];
    RngView v;
    v.base = keep_u32((uint32_t)__cvta_generic_to_shared(t));
    return v;
}

template <bool MIXED>
__device__ __forceinline__ RngView fill_rng_shared(RngShared<MIXED> *t, double m11, double m12,
                                                   double m21, double m22) {
    RngView v = fill_math_shared(&t->math);
    for (int i = threadIdx.x; i < TRIG_N; i += blockDim.x) {
        double2 cs = g_rng_base.trig[i];
        if constexpr (MIXED) {
            t->trig[2 * i] = make_double2(fma(m12, cs.y, m11 * cs.x), fma(m12, cs.x, -(m11 * cs.y)));
            t->trig[2 * i + 1] = make_double2(fma(m22, cs.y, m21 * cs.x), fma(m22, cs.x, -(m21 * cs.y)));
        } else {
            t->trig[i] = make_double2(m11 * cs.x, m11 * cs.y);
        }
    }
    return v;
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(addr));
    return r;
}

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
    return r;
}

// Polynomial coefficients, read straight from the constant bank.
//   -2 ln(1 + r) with r = -r'/2, |r'| <= 2^-9:  r' + r'^2/4 + r'^3/12 + r'^4/32  (+ 3.6e-16)
// A DFMA can take one operand from the constant bank / an immediate; an immediate
// must have its low 32 bits zero.  Each two-constant FMA below therefore pairs one
// exactly-representable "short" literal with one constant-bank value, so that no
// coefficient has to live in (and be read from) a vector register.
__constant__ double C_1_12 = 1.0 / 12.0;
__constant__ double C_N2LN2 = -1.3862943611198906188;        // -2 ln 2
__constant__ double C_1_24 = 1.0 / 24.0;
__constant__ double C_N1_6 = -1.0 / 6.0;
// 1/120 rounded to 20 mantissa bits (rel. 6e-8): it multiplies d^5 <= 3e-13
#define OPTMC_SHORT_1_120 0x1.11111p-7
// offset inside an angle bucket: d = j' h / 2^23 with j' odd in (-2^22, 2^22), h = 2 pi / 1024
__constant__ double C_FINE = 6.283185307179586477 / 1024.0 / 8388608.0;

// -2 ln u for a uniform u in (0,1) built from the words (lo, hi):
//   hi[31:20] exponent bits: k = their leading zeros, u in [2^-(k+1), 2^-k)
//   hi[19:0] : lo            the 52 mantissa bits of m in [1,2);  u = m 2^-(k+1)
// If all 12 exponent bits are zero (p = 2^-12) the exponent continues into the
// mantissa bits (slow path), so u stays exactly uniform on its 2^-64 grid.
__device__ __forceinline__ double neg2_log_poly(double m, double2 tb, double kfac) {
    double r = fma(m, tb.x, 2.0);                  // r' = -2 (m / c - 1)
    double p = fma(r, 0.03125, C_1_12);
    p = fma(p, r, 0.25);
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
    double2 tb = lds_f64x2(t.base + (uint32_t)((mbits >> 43) & (LOG_N - 1)) * 16u);
    return neg2_log_poly(m, tb, (double)(11 - (12 + k)));
}

__device__ __forceinline__ double neg2_log_u(uint32_t lo, uint32_t hi, RngView t) {
    uint32_t e = hi >> 20;
    if (__builtin_expect(e == 0u, 0)) return neg2_log_u_tail(lo, hi, t);
    double m = __hiloint2double(0x3FF00000 | (hi & 0xFFFFF), lo);
    double2 tb = lds_f64x2(t.base + ((hi >> 7) & ((LOG_N - 1) * 16u)));
    return neg2_log_poly(m, tb, (double)top_bit(e));
}

// ln s for a positive normal double s: exponent from the bit pattern, ln of the
// mantissa from the same table/polynomial as above.  7 fp64-pipe instructions
// (CUDA's log: ~30 plus special-case branches); abs. error <= 4e-15.
__constant__ double C_LN2 = 0.69314718055994530942;
__device__ __forceinline__ double fast_log(double s, RngView t) {
    int hi = __double2hiint(s);
    double m = __hiloint2double((hi & 0xFFFFF) | 0x3FF00000, __double2loint(s));
    double2 tb = lds_f64x2(t.base + (((uint32_t)hi >> 7) & ((LOG_N - 1) * 16u)));
    double r = fma(m, tb.x, 2.0);
    double p = fma(r, 0.03125, C_1_12);
    p = fma(p, r, 0.25);
    p = fma(p, r, 1.0);
    double q = fma(p, r, tb.y);                       // -2 ln m + 24 ln 2
    return fma((double)((hi >> 20) - 1023 + 12), C_LN2, -0.5 * q);
}

// e^x for |x| < 700: x = (512 k + j) ln2/512 + r, |r| <= ln2/1024 = 6.8e-4;
// e^x = 2^k * 2^(j/512) * e^r with a degree-4 polynomial (r^5/120 < 1.2e-18).  9 fp64-pipe
// instructions (CUDA's exp: ~18 plus range branches; a 32-entry table needs degree 6 and
// 11); rel. error <= 3.2e-16 (exact emulation over |x| <= 700).
__constant__ double C_EXP[3] = {738.65986093514929,              // 512 / ln 2
                                1.0 / 24.0, 1.0 / 6.0};
#define OPTMC_LN2_N_HI 0x1.62e42fcp-10     /* ln2/512 with the low 26 mantissa bits cleared */
#define OPTMC_LN2_N_LO 1.0831887298761837e-11
__device__ __forceinline__ double fast_exp(double x, RngView t) {
    const double MAGIC = 6755399441055744.0;          // 1.5 * 2^52: low word = rint(value)
    double tt = fma(x, C_EXP[0], MAGIC);
    int n = __double2loint(tt);
    double nd = tt - MAGIC;
    double r = fma(nd, -OPTMC_LN2_N_HI, x);
    r = fma(nd, -OPTMC_LN2_N_LO, r);
    double p = fma(r, C_EXP[1], C_EXP[2]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double y = p * lds_f64(t.base + EXP_OFF + ((uint32_t)(n & (EXP_N - 1)) << 3));
    return __hiloint2double(__double2hiint(y) + ((n >> 9) << 20), __double2loint(y));
}

// Baseline pair by the CUDA math library (also the jump-size path).
__device__ __forceinline__ void normal_pair_lib(uint32_t w0, uint32_t w1, uint32_t w2, double &z1,
                                                double &z2) {
    double rad = sqrt(-2.0 * log(u01_open(w0, w1)));
    double s, c;
    sincospi(((double)w2 + 0.5) * 0x1.0p-31, &s, &c);  // angle = 2 pi (w2 + .5) / 2^32
    z1 = rad * c;
    z2 = rad * s;
}

// Hand-written pair.  angle = a_i + d: the top 10 bits of w2 pick the bucket, the
// low 22 bits the offset d in (-h/2, h/2).  normal_parts_fast returns the squared radius
// L = -2 ln u >= 0 and the direction,
//   MIXED = false: (t1, t2) = s (cos, sin)(angle)
//   MIXED = true : (t1, t2) = M (cos, sin)(angle)^T
// and the pair is (z1, z2) = sqrt(L) (t1, t2).  Heston's step needs the normals only as
// sqrt(v) z, so it takes the parts and draws ONE root, sqrt(v L) (models.cuh).
// L is floored at 0: for the two largest u of the top binade the table + polynomial
// value of -2 ln u (abs. error 5e-16 there) comes out at -2e-16 instead of +2e-16.
template <bool MIXED>
__device__ __forceinline__ void normal_parts_fast(uint32_t w0, uint32_t w1, uint32_t w2, double &L,
                                                  double &t1, double &t2, RngView t) {
    L = relu64(neg2_log_u(w0, w1, t));
    int fine = (int)((w2 & 0x3FFFFFu) * 2u) - 0x3FFFFF;            // odd, in (-2^22, 2^22)
    double d = (double)fine * C_FINE;
    double d2 = d * d;
    double cd = fma(fma(d2, C_1_24, -0.5), d2, 1.0);                         // cos d
    double sd = fma(fma(d2, OPTMC_SHORT_1_120, C_N1_6), d2, 1.0) * d;        // sin d
    if constexpr (MIXED) {
        uint32_t row = t.base + TRIG_OFF + ((w2 >> 17) & ((TRIG_N - 1) * 32u));
        double2 r1 = lds_f64x2(row), r2 = lds_f64x2(row + 16u);
        t1 = fma(r1.x, cd, r1.y * sd);
        t2 = fma(r2.x, cd, r2.y * sd);
    } else {
        double2 cs = lds_f64x2(t.base + TRIG_OFF + ((w2 >> 18) & ((TRIG_N - 1) * 16u)));
        t1 = fma(cs.x, cd, -(cs.y * sd));
        t2 = fma(cs.y, cd, cs.x * sd);
    }
}

template <bool MIXED>
__device__ __forceinline__ void normal_pair_fast(uint32_t w0, uint32_t w1, uint32_t w2, double &z1,
                                                 double &z2, RngView t) {
    double L, t1, t2;
    normal_parts_fast<MIXED>(w0, w1, w2, L, t1, t2, t);
    double rad = sqrt_nonneg_fast(L);
    z1 = rad * t1;
    z2 = rad * t2;
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
    while (u > cdf && n < 1024) {
        ++n;
        p *= mu / (double)n;
        cdf += p;
    }
    return n;
}

}  // namespace optmc
