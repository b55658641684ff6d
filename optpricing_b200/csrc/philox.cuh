// philox.cuh -- counter-based RNG for the path kernels.
//
// Replaces the reference's host-side draws: rng.standard_normal / rng.poisson in
// MonteCarloTechnique._simulate_sde_path (monte_carlo.py:222-240) and the numba
// global-RNG jump sizes (mc_kernels.py:111,170,269,314-316).  Nothing is
// materialised: every draw is a pure function of (seed, draw id, step, stream).
//
// Philox4x32-10 (Salmon et al., SC'11).  Counter layout used by the engine:
//   c0, c1 = draw id (64 bit)        c2 = call index       c3 = stream tag
// Stream tags (c3): 0 = Brownian draws -- packed (philox_pair_words: call g feeds
//   Box-Muller pairs 2g and 2g+1) where the step loop needs no Poisson word, else one
//   call per step (two words for the pair, the fourth for the step's Poisson draw);
//   0xFE = whole-path Poisson draw; 0xFF = extra Poisson bits;
//   (1 + partner) | (jump index << 8) = jump-size draws; 0x40 = fp32-mode pairs;
//   0x20.. = single-step samplers (levy.cuh).
//
// Instruction budget (B200: the fp64 pipe AND the integer ALU pipe are both
// half-rate, 16 lanes per SM sub-partition; ncu on the first version of the
// Heston kernel showed fp64 42 % / ALU 41 % busy with issue slots the binding
// resource).  Hence: round keys are expanded once on the host (no per-round key
// adds), polynomial coefficients live in the constant bank (no UMOV pairs), the
// uniform's exponent and mantissa come out of one integer-to-double conversion, and the
// transcendental parts use shared-memory tables: 512 x 16 B and a degree-4 polynomial
// for the logarithm, two 512-row tables and one complex product for the angle.
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

// rounds [R0, R1) of the same function, for step loops that spread a call over their body
template <int R0, int R1>
__device__ __forceinline__ uint4 philox_rounds(uint4 c, const PhiloxKey &k) {
#pragma unroll
    for (int i = R0; i < R1; ++i) {
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

// The two words of Box-Muller pair `pair` of one draw under the packed layout the step
// loops use (european.cuh): call c2 = g of stream 0 yields pairs 2g = (x, y) and
// 2g + 1 = (z, w).
__device__ __forceinline__ void philox_pair_words(const PhiloxKey &k, uint64_t id, uint32_t pair,
                                                  uint32_t &wa, uint32_t &wb) {
    const uint4 a = philox_draw(k, id, pair >> 1, 0u);
    wa = (pair & 1u) ? a.z : a.x;
    wb = (pair & 1u) ? a.w : a.y;
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

// The same for a normal a > 0 (no clamp of the seed's argument).
__device__ __forceinline__ double sqrt_pos_fast(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(__double2hiint(a), 0)));
    const double h = __hiloint2double(__double2hiint(y) - 0x00100000, 0);
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
// Two N(0,1) from TWO Philox words (Box-Muller), the mapping every fp64 step loop uses:
//
//   wa : wb[13:0]   x: 46 random bits; the radius uniform is u = (x + 1/2) 2^-46
//   wb[31:23]  i: coarse angle bucket, a_i = (i + 1/2) 2 pi / 512
//   wb[22:14]  j: fine offset,          b_j = ((j + 1/2) / 512 - 1/2) 2 pi / 512
//
//   (z1, z2) = sqrt(-2 ln u) (cos, sin)(a_i + b_j)
//
// The angle runs over 2^18 equally spaced directions: every trigonometric moment of order
// below 2^18 is exact, and the direction costs two table rows and one complex product --
// (cos a_i, sin a_i)(cos b_j, sin b_j) -- instead of a table row, two polynomials in the
// offset and the product (round 1: 96 bits and 11 fp64 instructions per pair; now 64 bits
// and 4).  |z| reaches sqrt(2 * 47 ln 2) = 8.07.  One Philox4x32 call feeds two pairs.
// IMPL 0 = the same mapping through the CUDA math library (A/B checks);
// IMPL 1 = hand-written for the fp64 pipe.
// ---------------------------------------------------------------------------
constexpr int LOG_N = 512;     // buckets of [1,2) for ln
constexpr int ANG_N = 512;     // coarse angle buckets, and fine offsets per bucket
constexpr int EXP_N = 512;     // 2^(j/512) for exp

// Host-filled base tables (one copy per context, in global memory).
struct RngBase {
    // bucket i has midpoint c_i = 1 + (i + .5)/512;
    //   x = -2 / c_i,  y = -2 ln c_i
    double2 logt[LOG_N];
    // 2^(j/512), j = 0..511
    double exp2t[EXP_N];
    // (cos a_i, sin a_i) and (cos b_j, sin b_j)
    double2 coarse[ANG_N];
    double2 fine[ANG_N];
};
__device__ RngBase g_rng_base;

// ---------------------------------------------------------------------------
// Shared-memory tables.  One layout template serves every kernel; what varies is how many
// COPIES of each table a block keeps:
//
//   copies = 1: compact (28 / 36 KB per block, several blocks per SM).  A random-address
//     LDS.128 then costs ~12.8 shared-memory wavefronts instead of 4 (ncu,
//     profiles/r01_ncu_heston_summary.md: l1tex 94 %): the eight lanes of a quarter-warp pick
//     eight random 16-byte bank groups.
//   copies = C > 1: every entry is stored C times, entry i of copy g at (C i + g) entries,
//     and lane l reads copy l mod C.  With 8 copies of a 16-byte table the eight lanes of a
//     quarter-warp hit eight different bank groups whatever their entries are: conflict-free
//     by construction (measured: 1.75e9 -> 6.6e3 conflict wavefronts per C2 launch).  The
//     tables then fill the shared memory of an SM, so ONE block of 768 / 1024 threads per SM
//     shares them (k_european_wide).
//
// Order in memory: log | exp | coarse T1 | coarse T2 (two-factor kinds) | fine.
//   log    (x, y) = (-2 / c_i, -2 ln c_i)                       16 B entries
//   exp    2^(j/512)                                             8 B entries
//   T1/T2  coarse rows, specialised per launch:
//            one-factor: T1 = (s cos a_i, s sin a_i), s = sqrt(dt) (1 for plain normals);
//            two-factor: the 2x2 mixing matrix M applied to the rotation basis,
//              T_r = (A_r, B_r),  A_r = M_r1 cos a_i + M_r2 sin a_i,  B_r = M_r2 cos a_i - M_r1 sin a_i,
//            so that M (cos(a_i + b), sin(a_i + b))^T = (A_r cos b + B_r sin b)_r and the
//            correlated, scaled increments need no further arithmetic.
//   fine   (cos b_j, sin b_j)
// ---------------------------------------------------------------------------
__host__ __device__ constexpr int ilog2(uint32_t x) { return x <= 1u ? 0 : 1 + ilog2(x >> 1); }

template <int LOGC, int EXPC, int ANGC, int FINEC>
struct TableLayout {
    static constexpr int LOG_COPIES = LOGC, EXP_COPIES = EXPC, ANG_COPIES = ANGC, FINE_COPIES = FINEC;
    static constexpr uint32_t LOG_STRIDE = 16u * LOGC, EXP_STRIDE = 8u * EXPC,
                              ANG_STRIDE = 16u * ANGC, FINE_STRIDE = 16u * FINEC;
    static constexpr uint32_t LOG_OFF = 0u;
    static constexpr uint32_t EXP_OFF = LOG_OFF + (uint32_t)LOG_N * LOG_STRIDE;
    static constexpr uint32_t T1_OFF = EXP_OFF + (uint32_t)EXP_N * EXP_STRIDE;
    static constexpr uint32_t T2_OFF = T1_OFF + (uint32_t)ANG_N * ANG_STRIDE;
    __host__ __device__ static constexpr uint32_t fine_off(bool mixed) {
        return T2_OFF + (mixed ? (uint32_t)ANG_N * ANG_STRIDE : 0u);
    }
    __host__ __device__ static constexpr uint32_t bytes(bool mixed) {
        return fine_off(mixed) + (uint32_t)ANG_N * FINE_STRIDE;
    }
};

// 32-bit shared-window address of a block's tables + the lane's copy offsets.
template <class LAYOUT>
struct RngViewT {
    using Lay = LAYOUT;
    uint32_t base;          // warp-uniform: address arithmetic stays in uniform registers
    uint32_t l8, l4, l2;    // 16 (l mod 8), 16 (l mod 4), 16 (l mod 2): OR-ed into entry offsets
    uint32_t le;            // 8 (l mod EXP_COPIES)
    uint32_t k3ff;          // 0x3FF00000 held in a register (mantissa_hi)
    // beta-specific power tables of the SABR scenario kernel (models.cuh: pow_beta), else unused:
    // per-lane base of c_i^beta (two copies, 8-byte entries) and base of 2^(e beta)
    uint32_t pw_c = 0, pw_e = 0;
    template <int COPIES>
    __device__ __forceinline__ uint32_t lane16() const {
        static_assert(COPIES == 1 || COPIES == 2 || COPIES == 4 || COPIES == 8, "copies");
        return COPIES == 8 ? l8 : COPIES == 4 ? l4 : COPIES == 2 ? l2 : 0u;
    }
    // addresses of the entries selected by the bit fields the generators use
    __device__ __forceinline__ uint32_t log_at(uint32_t hi) const {      // mantissa bits 19..11
        constexpr int sh = 11 - ilog2(Lay::LOG_STRIDE);
        return base + Lay::LOG_OFF + (((hi >> sh) & ((LOG_N - 1) * Lay::LOG_STRIDE)) |
                                      lane16<Lay::LOG_COPIES>());
    }
    __device__ __forceinline__ uint32_t exp_at(uint32_t n) const {       // n mod 512
        static_assert(Lay::EXP_COPIES >= 1, "this layout carries no exp table");
        return base + Lay::EXP_OFF + (((n & (EXP_N - 1)) * Lay::EXP_STRIDE) |
                                      (Lay::EXP_COPIES > 1 ? le : 0u));
    }
    __device__ __forceinline__ uint32_t coarse_at(uint32_t wb) const {   // wb[31:23]
        constexpr int sh = 23 - ilog2(Lay::ANG_STRIDE);
        return base + Lay::T1_OFF + (((wb >> sh) & ((ANG_N - 1) * Lay::ANG_STRIDE)) |
                                     lane16<Lay::ANG_COPIES>());
    }
    template <bool MIXED>
    __device__ __forceinline__ uint32_t fine_at(uint32_t wb) const {     // wb[22:14]
        constexpr int sh = 14 - ilog2(Lay::FINE_STRIDE);
        return base + Lay::fine_off(MIXED) + (((wb >> sh) & ((ANG_N - 1) * Lay::FINE_STRIDE)) |
                                              lane16<Lay::FINE_COPIES>());
    }
};
using CompactLayout = TableLayout<1, 1, 1, 1>;
using WideLayout = TableLayout<8, 0, 8, 4>;      // Heston, Bates, one-factor models: 160 / 224 KB
using WideExpLayout = TableLayout<8, 4, 8, 2>;   // SABR kinds (exp table in the step): 224 KB
using ScenLayout = TableLayout<8, 1, 8, 2>;      // SABR spot scenarios: one exp per vol state and step
                                                 // only (212 KB) + 9 KB of power tables behind it
constexpr int POW_E_N = 128, POW_E_BIAS = 64;    // 2^(e beta) for the binary exponents -64 .. 63
constexpr uint32_t POW_C_BYTES = 512u * 2u * 8u, POW_E_BYTES = (uint32_t)POW_E_N * 8u;
using RngView = RngViewT<CompactLayout>;
using RngViewW = RngViewT<WideLayout>;
using RngViewS = RngViewT<WideExpLayout>;

// Static shared-memory block of the compact layout.
template <bool MIXED>
struct RngShared {
    alignas(16) unsigned char bytes[CompactLayout::bytes(MIXED)];
};
struct MathShared {      // log + exp only (fast_log / fast_exp in kernels without normals)
    alignas(16) unsigned char bytes[CompactLayout::T1_OFF];
};

// one coarse row, mixed (see above)
__device__ __forceinline__ void mix_row(double2 cs, double m11, double m12, double m21, double m22,
                                        double2 &r1, double2 &r2) {
    r1 = make_double2(fma(m12, cs.y, m11 * cs.x), fma(m12, cs.x, -(m11 * cs.y)));
    r2 = make_double2(fma(m22, cs.y, m21 * cs.x), fma(m22, cs.x, -(m21 * cs.y)));
}

template <class LAYOUT>
__device__ __forceinline__ RngViewT<LAYOUT> make_view(const void *smem) {
    RngViewT<LAYOUT> v;
    uint32_t laneid;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(laneid));
    v.base = keep_u32((uint32_t)__cvta_generic_to_shared(smem));
    // (separately laundered, so that each stays its own register and an entry address is
    //  ONE three-input logic instruction: (bits & mask) | lane offset)
    v.l8 = keep_u32((laneid & 7u) * 16u);
    v.l4 = keep_u32((laneid & 3u) * 16u);
    v.l2 = keep_u32((laneid & 1u) * 16u);
    v.le = keep_u32((laneid & (uint32_t)(LAYOUT::EXP_COPIES > 1 ? LAYOUT::EXP_COPIES - 1 : 0)) * 8u);
    v.k3ff = keep_u32(0x3FF00000u);
    return v;
}

// log (+ exp) tables
template <class LAYOUT>
__device__ __forceinline__ void fill_math_tables(unsigned char *smem) {
    double2 *logt = reinterpret_cast<double2 *>(smem + LAYOUT::LOG_OFF);
    for (int i = threadIdx.x; i < LOG_N * LAYOUT::LOG_COPIES; i += blockDim.x)
        logt[i] = g_rng_base.logt[i / LAYOUT::LOG_COPIES];
    if constexpr (LAYOUT::EXP_COPIES >= 1) {
        double *expt = reinterpret_cast<double *>(smem + LAYOUT::EXP_OFF);
        for (int i = threadIdx.x; i < EXP_N * LAYOUT::EXP_COPIES; i += blockDim.x)
            expt[i] = g_rng_base.exp2t[i / LAYOUT::EXP_COPIES];
    }
}

// all tables of a kernel that draws normals
template <class LAYOUT, bool MIXED>
__device__ __forceinline__ RngViewT<LAYOUT> fill_rng_tables(unsigned char *smem, double m11,
                                                            double m12, double m21, double m22) {
    fill_math_tables<LAYOUT>(smem);
    double2 *t1 = reinterpret_cast<double2 *>(smem + LAYOUT::T1_OFF);
    double2 *t2 = reinterpret_cast<double2 *>(smem + LAYOUT::T2_OFF);
    double2 *fine = reinterpret_cast<double2 *>(smem + LAYOUT::fine_off(MIXED));
    for (int i = threadIdx.x; i < ANG_N * LAYOUT::ANG_COPIES; i += blockDim.x) {
        const double2 cs = g_rng_base.coarse[i / LAYOUT::ANG_COPIES];
        if constexpr (MIXED)
            mix_row(cs, m11, m12, m21, m22, t1[i], t2[i]);
        else
            t1[i] = make_double2(m11 * cs.x, m11 * cs.y);
    }
    for (int i = threadIdx.x; i < ANG_N * LAYOUT::FINE_COPIES; i += blockDim.x)
        fine[i] = g_rng_base.fine[i / LAYOUT::FINE_COPIES];
    return make_view<LAYOUT>(smem);
}

__device__ __forceinline__ RngView fill_math_shared(MathShared *t) {
    fill_math_tables<CompactLayout>(t->bytes);
    return make_view<CompactLayout>(t->bytes);
}

template <bool MIXED>
__device__ __forceinline__ RngView fill_rng_shared(RngShared<MIXED> *t, double m11, double m12,
                                                   double m21, double m22) {
    return fill_rng_tables<CompactLayout, MIXED>(t->bytes, m11, m12, m21, m22);
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
__constant__ double C_TINY = 1e-290;                         // see step_heston_parts
__constant__ double C_N2LN2 = -1.3862943611198906188;        // -2 ln 2

// -2 ln u for the uniform u = x 2^-64, x = hi:lo != 0 (callers set a low bit).  The
// integer goes through ONE conversion to double -- exponent and mantissa of u in one
// instruction, the deep tail (u down to 2^-64) included: no leading-zero count, no
// rare branch to split the step loop's basic block.
__device__ __forceinline__ double neg2_log_poly(double m, double2 tb, double kfac) {
    double r = fma(m, tb.x, 2.0);                  // r' = -2 (m / c - 1)
    double p = fma(r, 0.03125, C_1_12);
    p = fma(p, r, 0.25);
    p = fma(p, r, 1.0);
    // u = m 2^kfac, -64 <= kfac <= 0:  -2 ln u = -2 ln m - 2 kfac ln 2;  tb.y = -2 ln c
    return fma(p, r, fma(kfac, C_N2LN2, tb.y));
}

// high word of the mantissa m in [1, 2) of a positive double with high word hi:
// (hi & 0xFFFFF) | 0x3FF00000 as ONE three-input logic instruction
__device__ __forceinline__ int mantissa_hi(int hi, uint32_t k3ff) {
    int r;
    asm("lop3.b32 %0, %1, 0xFFFFF, %2, 0xEA;" : "=r"(r) : "r"(hi), "r"(k3ff));
    return r;
}

template <class V>
__device__ __forceinline__ double neg2_log_u(uint32_t lo, uint32_t hi, V t) {
    const double d = __ull2double_rn(((unsigned long long)hi << 32) | lo);      // u 2^64
    const int dh = __double2hiint(d);
    const double m = __hiloint2double(mantissa_hi(dh, t.k3ff), __double2loint(d));
    const double2 tb = lds_f64x2(t.log_at((uint32_t)dh));
    return neg2_log_poly(m, tb, (double)((dh - 0x43F00000) >> 20));       // exponent - 64
}

// ln s for a positive normal double s: exponent from the bit pattern, ln of the
// mantissa from the same table/polynomial as above.  7 fp64-pipe instructions
// (CUDA's log: ~30 plus special-case branches); abs. error <= 4e-15.
__constant__ double C_LN2 = 0.69314718055994530942;
template <class V>
__device__ __forceinline__ double fast_log(double s, V t) {
    int hi = __double2hiint(s);
    double m = __hiloint2double((hi & 0xFFFFF) | 0x3FF00000, __double2loint(s));
    double2 tb = lds_f64x2(t.log_at((uint32_t)hi));
    double r = fma(m, tb.x, 2.0);
    double p = fma(r, 0.03125, C_1_12);
    p = fma(p, r, 0.25);
    p = fma(p, r, 1.0);
    double q = fma(p, r, tb.y);                       // -2 ln m
    return fma((double)((hi >> 20) - 1023), C_LN2, -0.5 * q);
}

// e^x for |x| < 700: x = (512 k + j) ln2/512 + r, |r| <= ln2/1024 = 6.8e-4;
// e^x = 2^k * 2^(j/512) * e^r with a degree-4 polynomial (r^5/120 < 1.2e-18).  9 fp64-pipe
// instructions (CUDA's exp: ~18 plus range branches; a 32-entry table needs degree 6 and
// 11); rel. error <= 3.2e-16 (exact emulation over |x| <= 700).
__constant__ double C_EXP[3] = {738.65986093514929,              // 512 / ln 2
                                1.0 / 24.0, 1.0 / 6.0};
#define OPTMC_LN2_N_HI 0x1.62e42fcp-10     /* ln2/512 with the low 26 mantissa bits cleared */
#define OPTMC_LN2_N_LO 1.0831887298761837e-11
template <class V>
__device__ __forceinline__ double fast_exp(double x, V t) {
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
    double y = p * lds_f64(t.exp_at((uint32_t)n));
    return __hiloint2double(__double2hiint(y) + ((n >> 9) << 20), __double2loint(y));
}

// Baseline pair by the CUDA math library from three words (the single-step samplers of
// levy.cuh keep it: 64-bit radius uniform, 32-bit angle).
__device__ __forceinline__ void normal_pair_lib(uint32_t w0, uint32_t w1, uint32_t w2, double &z1,
                                                double &z2) {
    double rad = sqrt(-2.0 * log(u01_open(w0, w1)));
    double s, c;
    sincospi(((double)w2 + 0.5) * 0x1.0p-31, &s, &c);  // angle = 2 pi (w2 + .5) / 2^32
    z1 = rad * c;
    z2 = rad * s;
}

// low word of the radius uniform (x + 1/2) 2^18: wb[13:0], the midpoint bit, zeros
__device__ __forceinline__ uint32_t radius_lo(uint32_t wb) { return wb * 0x40000u + 0x20000u; }

// The step loops' mapping (top of this section) through the CUDA math library: the same
// reals as normal_pair_fast, for A/B checks (normals_impl = 0).
__device__ __forceinline__ void normal_pair_lib2(uint32_t wa, uint32_t wb, double &z1, double &z2) {
    const double u = ldexp(__ull2double_rn(((unsigned long long)wa << 32) | radius_lo(wb)), -64);
    const double rad = sqrt(fmax(-2.0 * log(u), 0.0));
    double s, c;
    sincospi(((double)(wb >> 14) + 0.5) * 0x1.0p-17, &s, &c);   // 2 pi ((wb >> 14) + .5) / 2^18
    z1 = rad * c;
    z2 = rad * s;
}

// Hand-written pair.  normal_parts_fast returns the squared radius L = -2 ln u >= 0 and the
// direction,
//   MIXED = false: (t1, t2) = s (cos, sin)(angle)
//   MIXED = true : (t1, t2) = M (cos, sin)(angle)^T
// and the pair is (z1, z2) = sqrt(L) (t1, t2).  Heston's step needs the normals only as
// sqrt(v) z, so it takes the parts and draws ONE root, sqrt(v L) (models.cuh).
template <bool MIXED, class V>
__device__ __forceinline__ void normal_parts_fast(uint32_t wa, uint32_t wb, double &L, double &t1,
                                                  double &t2, V t) {
    // (no floor at 0: the largest u of the 46-bit grid is 1 - 2^-47, where -2 ln u = 1.4e-14,
    //  thirty times the table + polynomial's absolute error -- tests/test_host_api.py emulates it)
    L = neg2_log_u(radius_lo(wb), wa, t);
    const uint32_t row = t.coarse_at(wb), fin = t.template fine_at<MIXED>(wb);
    const double2 f = lds_f64x2(fin);                              // (cos b_j, sin b_j)
    if constexpr (MIXED) {
        const double2 r1 = lds_f64x2(row);
        const double2 r2 = lds_f64x2(row + (V::Lay::T2_OFF - V::Lay::T1_OFF));
        t1 = fma(r1.x, f.x, r1.y * f.y);
        t2 = fma(r2.x, f.x, r2.y * f.y);
    } else {
        const double2 cs = lds_f64x2(row);
        t1 = fma(cs.x, f.x, -(cs.y * f.y));
        t2 = fma(cs.y, f.x, cs.x * f.y);
    }
}

template <bool MIXED, class V>
__device__ __forceinline__ void normal_pair_fast(uint32_t wa, uint32_t wb, double &z1, double &z2,
                                                 V t) {
    double L, t1, t2;
    normal_parts_fast<MIXED>(wa, wb, L, t1, t2, t);
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
