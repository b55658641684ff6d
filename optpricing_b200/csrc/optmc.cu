// optmc.cu -- C ABI of liboptmc.so (see include/optmc.h for the contract and
// the reference lines each entry point replaces).  Host side only: argument
// checks, constant folding in the reference's expression order, launch
// geometry.  All arithmetic on paths happens in the kernels included below.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/optmc.h"
#include "european.cuh"
#include "european_f32.cuh"
#include "fed.cuh"
#include "levy.cuh"
#include "lsmc.cuh"

using namespace optmc;

struct optmc_ctx {
    int device = 0;
    int sm_count = 0;
    int normals_impl = 1;       // 0: CUDA math library, 1: hand-written (philox.cuh)
    int euro_blocks_per_sm = 0; // 0: occupancy API decides
    int euro_threads = EURO_THREADS;
    int euro_wide = 1;          // 1: lane-replicated tables, one block per SM (k_european_wide)
    int euro_threads_auto = 1;  // 1: Heston / Bates take the 1024-thread variant where it idles less
    int lsmc_mode = 1;          // 1: one persistent launch for all dates, 0: one launch per date
    int lsmc_ring = 2;          // cp.async ring in the persistent sweep: 0 never, 1 always, 2 from 2^19 paths
    int jump_schedule = 1;      // 1: pre-drawn jump schedule where it applies, 0: Poisson word per step
    int fed_stream = 1;         // 1: k_fed_stream for the models without jumps, 0: tiled k_fed for all
    int64_t launches = 0;
    unsigned char *scratch = nullptr;  // device
    size_t scratch_bytes = 0;
    double *moments_dev = nullptr;     // SWEEP_MAX * OPTMC_NMOM
    double *moments_host = nullptr;    // pinned
    unsigned int *ticket = nullptr;    // device, zero between launches
    uint4 *lsmc_slots = nullptr;       // device: flag-carrying partial sums of k_lsmc_persistent
    uint32_t lsmc_nonce = 0;
    unsigned long long *lsmc_timing = nullptr;   // diagnostics: see optmc_set_ptr("lsmc_timing")
    int lsmc_timing_all = 0;                     // 1: every block stamps (buffer gridDim.x times as large)
    // multi-GPU exchange of k_lsmc_persistent (optmc_lsmc_peer_*)
    uint4 *rank_slots = nullptr;                 // local buffer, exported over CUDA IPC
    uint4 *peer_rank_slots[OPTMC_MAX_PEERS] = {};
    int peer_rank = 0, peer_world = 1;
    int use_peers = 0;                           // "lsmc_use_peers": next optmc_lsmc is a joint call
    uint32_t peer_nonce = 0;
    uint32_t euro_peer_nonce = 0;                // k_peer_allreduce: one per call, same on all ranks
    std::string err;
};

static std::string g_create_err;

#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            char b_[512];                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                               \
            ctx->err = b_;                                                              \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

#define FAIL(msg)          \
    do {                   \
        ctx->err = (msg);  \
        return -2;         \
    } while (0)

static constexpr size_t SCRATCH_BYTES = 8u << 20;
// The buffer every rank exports over CUDA IPC: the LSMC exchange rows, then the slots of the
// small all-reduce (k_peer_allreduce): [2 call parity][OPTMC_MAX_PEERS][PEER_AR_MAX].
static constexpr size_t RANK_SLOTS_LSMC = (size_t)4 * OPTMC_MAX_PEERS * LSMC_P_STRIDE;
static constexpr size_t RANK_SLOTS_TOTAL = RANK_SLOTS_LSMC + (size_t)2 * OPTMC_MAX_PEERS * PEER_AR_MAX;
static constexpr int MAX_GRID = 148 * 16 * 2;   // partial rows the scratch is carved for

extern "C" int optmc_version(void) { return 100; }

extern "C" const char *optmc_last_error(const optmc_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int64_t optmc_launch_count(const optmc_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int optmc_ctx_create(int device, optmc_ctx **out) {
    if (!out) return -2;
    *out = nullptr;
    optmc_ctx *ctx = new optmc_ctx();
    auto fail = [&](cudaError_t e, const char *what) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(e);
        delete ctx;
        return -1;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(e, "cudaGetDeviceProperties");
    if (prop.major != 10) {
        g_create_err = "liboptmc is built for sm_100a (B200) only; device is sm_" +
                       std::to_string(prop.major) + std::to_string(prop.minor);
        delete ctx;
        return -3;
    }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->scratch_bytes = SCRATCH_BYTES;
    if ((e = cudaMalloc(&ctx->scratch, SCRATCH_BYTES)) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&ctx->moments_dev, sizeof(double) * SWEEP_MAX * OPTMC_NMOM)) != cudaSuccess)
        return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&ctx->ticket, 64)) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMemset(ctx->ticket, 0, 64)) != cudaSuccess) return fail(e, "cudaMemset");
    {
        const size_t bytes = sizeof(uint4) * 2 * (size_t)ctx->sm_count * LSMC_P_STRIDE;
        if ((e = cudaMalloc(&ctx->lsmc_slots, bytes)) != cudaSuccess) return fail(e, "cudaMalloc");
        if ((e = cudaMemset(ctx->lsmc_slots, 0, bytes)) != cudaSuccess) return fail(e, "cudaMemset");
    }
    if ((e = cudaHostAlloc(&ctx->moments_host, sizeof(double) * SWEEP_MAX * OPTMC_NMOM,
                           cudaHostAllocDefault)) != cudaSuccess)
        return fail(e, "cudaHostAlloc");
    {
        const size_t bytes = sizeof(uint4) * RANK_SLOTS_TOTAL;
        if ((e = cudaMalloc(&ctx->rank_slots, bytes)) != cudaSuccess) return fail(e, "cudaMalloc");
        if ((e = cudaMemset(ctx->rank_slots, 0, bytes)) != cudaSuccess) return fail(e, "cudaMemset");
    }
    {
        RngBase *h = new RngBase();
        const double pi = 3.14159265358979323846;
        for (int i = 0; i < LOG_N; ++i) {
            double c = 1.0 + (i + 0.5) / LOG_N;
            h->logt[i] = make_double2(-2.0 / c, -2.0 * std::log(c));
        }
        for (int j = 0; j < EXP_N; ++j) h->exp2t[j] = std::exp2(j / (double)EXP_N);
        for (int i = 0; i < ANG_N; ++i) {
            const double a = (i + 0.5) * (2.0 * pi / ANG_N);
            const double b = ((i + 0.5) / ANG_N - 0.5) * (2.0 * pi / ANG_N);
            h->coarse[i] = make_double2(std::cos(a), std::sin(a));
            h->fine[i] = make_double2(std::cos(b), std::sin(b));
        }
        e = cudaMemcpyToSymbol(g_rng_base, h, sizeof(RngBase));
        delete h;
        if (e != cudaSuccess) return fail(e, "cudaMemcpyToSymbol");
    }
    if ((e = cudaFuncSetAttribute(k_fed<OPTMC_BSM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fed_smem_bytes())) != cudaSuccess)
        return fail(e, "cudaFuncSetAttribute");
#define FED_ATTR(K)                                                                      \
    cudaFuncSetAttribute(k_fed<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                         (int)fed_smem_bytes())
    FED_ATTR(OPTMC_HESTON); FED_ATTR(OPTMC_MERTON); FED_ATTR(OPTMC_BATES);
    FED_ATTR(OPTMC_SABR); FED_ATTR(OPTMC_SABR_JUMP); FED_ATTR(OPTMC_KOU); FED_ATTR(OPTMC_DUPIRE);
#undef FED_ATTR
    *out = ctx;
    return 0;
}

extern "C" void optmc_ctx_destroy(optmc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->scratch);
    cudaFree(ctx->moments_dev);
    cudaFree(ctx->ticket);
    cudaFree(ctx->lsmc_slots);
    for (int r = 0; r < ctx->peer_world; ++r)
        if (ctx->peer_world > 1 && r != ctx->peer_rank && ctx->peer_rank_slots[r])
            cudaIpcCloseMemHandle(ctx->peer_rank_slots[r]);
    cudaFree(ctx->rank_slots);
    cudaFreeHost(ctx->moments_host);
    delete ctx;
}

extern "C" int optmc_set_ptr(optmc_ctx *ctx, const char *key, void *value) {
    if (!ctx || !key) return -2;
    if (!strcmp(key, "lsmc_timing")) {         // device buffer of (n_steps + 1) * 4 uint64, or NULL
        ctx->lsmc_timing = static_cast<unsigned long long *>(value);
        return 0;
    }
    FAIL(std::string("unknown pointer option ") + key);
}

extern "C" int optmc_set_int(optmc_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return -2;
    if (!strcmp(key, "normals_impl")) {
        if (value != 0 && value != 1) FAIL("normals_impl must be 0 or 1");
        ctx->normals_impl = (int)value;
    } else if (!strcmp(key, "euro_blocks_per_sm")) {
        if (value < 0 || value > 16) FAIL("euro_blocks_per_sm out of range");
        ctx->euro_blocks_per_sm = (int)value;
    } else if (!strcmp(key, "euro_threads")) {
        if (value != 64 && value != 128 && value != 256) FAIL("euro_threads must be 64, 128 or 256");
        ctx->euro_threads = (int)value;
    } else if (!strcmp(key, "euro_threads_auto")) {
        if (value != 0 && value != 1) FAIL("euro_threads_auto must be 0 or 1");
        ctx->euro_threads_auto = (int)value;
    } else if (!strcmp(key, "euro_wide")) {
        if (value != 0 && value != 1) FAIL("euro_wide must be 0 or 1");
        ctx->euro_wide = (int)value;
    } else if (!strcmp(key, "fed_stream")) {
        if (value != 0 && value != 1) FAIL("fed_stream must be 0 or 1");
        ctx->fed_stream = (int)value;
    } else if (!strcmp(key, "jump_schedule")) {
        if (value != 0 && value != 1) FAIL("jump_schedule must be 0 or 1");
        ctx->jump_schedule = (int)value;
    } else if (!strcmp(key, "lsmc_use_peers")) {
        if (value != 0 && value != 1) FAIL("lsmc_use_peers must be 0 or 1");
        if (value == 1 && ctx->peer_world < 2) FAIL("no peers attached (optmc_lsmc_peer_attach)");
        ctx->use_peers = (int)value;
    } else if (!strcmp(key, "lsmc_ring")) {
        if (value < 0 || value > 2) FAIL("lsmc_ring must be 0 (never), 1 (always) or 2 (by size)");
        ctx->lsmc_ring = (int)value;
    } else if (!strcmp(key, "lsmc_timing_all")) {
        ctx->lsmc_timing_all = value != 0;
    } else if (!strcmp(key, "lsmc_mode")) {
        if (value != 0 && value != 1) FAIL("lsmc_mode must be 0 (launch per date) or 1 (persistent)");
        ctx->lsmc_mode = (int)value;
    } else {
        FAIL(std::string("unknown option ") + key);
    }
    return 0;
}

// ---------------------------------------------------------------------------
// constants, in the reference's expression order
// ---------------------------------------------------------------------------
static int make_consts(optmc_ctx *ctx, const optmc_model *m, double x0, double dt, int n_steps,
                       int correlation, Consts *out) {
    Consts c;
    memset(&c, 0, sizeof c);
    if (m->kind < OPTMC_BSM || m->kind > OPTMC_DUPIRE) FAIL("unknown model kind");
    if (m->kind == OPTMC_DUPIRE) {
        if (!m->lv_table_dev) FAIL("Dupire needs lv_table_dev");
        if (m->lv_n_s < 2 || !(m->lv_inv_h > 0.0)) FAIL("Dupire needs lv_n_s >= 2 and lv_inv_h > 0");
        if (m->lv_n_t < n_steps) FAIL("Dupire table has fewer time rows than n_steps");
    }
    if (!(dt > 0.0)) FAIL("dt must be positive");
    c.kind = m->kind;
    c.x0 = x0;
    c.v0 = m->v0;
    c.dt = dt;
    c.sqrt_dt = std::sqrt(dt);
    c.inv_sqrt_dt = 1.0 / c.sqrt_dt;
    c.r = m->r;
    c.q = m->q;
    double comp = 0.0;
    if (m->kind == OPTMC_MERTON || m->kind == OPTMC_BATES || m->kind == OPTMC_SABR_JUMP)
        comp = m->lambda * (std::exp(m->mu_j + 0.5 * m->sigma_j * m->sigma_j) - 1);  // mc_kernels.py:103
    if (m->kind == OPTMC_KOU)
        comp = m->lambda * ((m->p_up * m->eta1 / (m->eta1 - 1)) +
                            ((1 - m->p_up) * m->eta2 / (m->eta2 + 1)) - 1);          // :301-303
    c.sigma = m->sigma;
    c.drift = (m->r - m->q - 0.5 * m->sigma * m->sigma - comp) * dt;                 // :29,104,304
    c.rqc = m->r - m->q - comp;
    c.rqc_dt = c.rqc * dt;
    c.n_rqc_dt = n_steps * c.rqc_dt;
    c.n_drift = n_steps * c.drift;
    c.neg_half_dt = -0.5 * dt;
    c.kappa_theta_dt = m->kappa * m->theta * dt;
    c.one_minus_kappa_dt = 1.0 - m->kappa * dt;
    if ((m->kind == OPTMC_HESTON || m->kind == OPTMC_BATES) && !(m->v0 >= 0.0))
        FAIL("v0 must be non-negative for Heston/Bates");
    c.xi = m->vol_of_vol;
    c.alpha = m->alpha;
    c.beta = m->beta;
    c.neg_half_alpha2_dt = -0.5 * m->alpha * m->alpha * dt;
    c.log_v0 = m->v0 > 0.0 ? std::max(std::log(m->v0), -600.0) : -600.0;
    c.beta_mode = m->beta == 1.0 ? 1 : m->beta == 0.0 ? 2 : m->beta == 0.5 ? 3 : 0;
    {   // binomial series of (1 + rho)^beta in r' = -2 rho (models.cuh: pow_beta)
        const double b = m->beta;
        c.pw[0] = -b / 2.0;
        c.pw[1] = b * (b - 1.0) / 8.0;
        c.pw[2] = -b * (b - 1.0) * (b - 2.0) / 48.0;
        c.pw[3] = b * (b - 1.0) * (b - 2.0) * (b - 3.0) / 384.0;
    }
    c.rho = m->rho;
    c.rho_bar = std::sqrt(1 - m->rho * m->rho);                                      // :59
    const double sd = c.sqrt_dt, rho = c.rho, rb = c.rho_bar;
    if (correlation == OPTMC_CORR_REFERENCE_EUROPEAN) {
        // dw1 = rho a + rho_bar b, dw2 = a (monte_carlo.py:230-231); kernel: rho dw1 + rho_bar dw2
        c.m11 = sd * rho;
        c.m12 = sd * rb;
        c.m21 = sd * (rho * rho + rb);
        c.m22 = sd * rho * rb;
    } else if (correlation == OPTMC_CORR_INDEPENDENT_DRAWS) {
        c.m11 = sd;
        c.m12 = 0.0;
        c.m21 = sd * rho;
        c.m22 = sd * rb;
    } else {
        FAIL("unknown correlation mode");
    }
    if (m->kind == OPTMC_HESTON || m->kind == OPTMC_BATES) {
        c.m21 *= c.xi;
        c.m22 *= c.xi;
    }
    c.lv_table = m->lv_table_dev;
    c.lv_n_s = m->lv_n_s;
    c.lv_lo = m->lv_log_s_lo;
    c.lv_inv_h = m->lv_inv_h;
    c.lv_max_u = std::nextafter((double)(m->lv_n_s - 1), 0.0);
    c.rq = m->r - m->q;
    c.lam_dt = m->lambda * dt;
    c.mu_j = m->mu_j;
    c.sigma_j = m->sigma_j;
    c.p_up = m->p_up;
    if (m->kind == OPTMC_KOU) {
        c.inv_eta1 = 1.0 / m->eta1;
        c.inv_eta2 = 1.0 / m->eta2;
        double pb = std::floor(m->p_up * 4294967296.0);
        c.p_up_bits = (uint32_t)std::min(std::max(pb, 0.0), 4294967295.0);
    }
    *out = c;
    return 0;
}

static void poisson_consts(EuroArgs &A) {
    const double mu_T = A.c.lam_dt * A.n_steps;
    A.pois_chunks = mu_T > 16.0 ? (int)std::min(std::ceil(mu_T / 16.0), 65536.0) : 1;
    A.pois_mu_path = mu_T / A.pois_chunks;
    A.pois_p0_path = std::exp(-A.pois_mu_path);
    A.pois_p0_path_bits =
        (uint32_t)std::min(std::floor(A.pois_p0_path * 4294967296.0), 4294967295.0);
    for (int two = 0; two < 2; ++two) {
        double mu = A.c.lam_dt * (two + 1);
        double p0 = std::exp(-mu);
        A.pois_mu[two] = mu;
        A.pois_p0[two] = p0;
        A.pois_p0_bits[two] = (uint32_t)std::min(std::floor(p0 * 4294967296.0), 4294967295.0);
    }
}

static int check_sim(optmc_ctx *ctx, const optmc_sim *s) {
    if (!(s->s0 > 0.0)) FAIL("s0 must be positive");
    if (!(s->maturity > 0.0)) FAIL("maturity must be positive");
    if (s->n_steps < 1) FAIL("n_steps must be >= 1");
    if (s->n_local < 0 || s->draw_offset < 0 || s->n_draws < 0) FAIL("negative draw counts");
    if (s->draw_offset + s->n_local > s->n_draws) FAIL("draw range exceeds n_draws");
    return 0;
}

template <class K>
static int euro_grid(optmc_ctx *ctx, K kernel, int64_t n_local) {
    int per_sm = ctx->euro_blocks_per_sm;
    const int threads = ctx->euro_threads;
    if (per_sm <= 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) !=
                cudaSuccess || per_sm < 1)
            per_sm = 1;
    }
    int64_t need = (n_local + threads - 1) / threads;
    // twice the resident blocks: the second half back-fills SMs as the first drains
    // (measured 3.5 % faster than exactly one resident wave, profiles/r01_sweep*.log)
    int64_t cap = (int64_t)ctx->sm_count * per_sm * (ctx->euro_blocks_per_sm > 0 ? 1 : 2);
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min(need, cap), MAX_GRID));
}

// KIND x ANTI x NIMPL dispatch
#define DISPATCH_KIND(KIND_VAR, MACRO)                         \
    switch (KIND_VAR) {                                        \
        case OPTMC_BSM: MACRO(OPTMC_BSM); break;               \
        case OPTMC_HESTON: MACRO(OPTMC_HESTON); break;         \
        case OPTMC_MERTON: MACRO(OPTMC_MERTON); break;         \
        case OPTMC_BATES: MACRO(OPTMC_BATES); break;           \
        case OPTMC_SABR: MACRO(OPTMC_SABR); break;             \
        case OPTMC_SABR_JUMP: MACRO(OPTMC_SABR_JUMP); break;   \
        case OPTMC_KOU: MACRO(OPTMC_KOU); break;               \
        case OPTMC_DUPIRE: MACRO(OPTMC_DUPIRE); break;         \
        default: FAIL("unknown model kind");                   \
    }

// Path kernels of the jump models: jumps from a pre-drawn schedule (european.cuh) while lambda T
// is small enough for eight register slots to hold nearly every path's jumps, else a Poisson
// word per step.
static bool use_jump_schedule(const EuroArgs &A) {
    return A.pois_chunks == 1 && A.pois_mu_path <= 2.0 && A.n_steps < 65535;
}

template <int KIND, bool ANTI, int NIMPL, bool SCHED = false, bool BGEN = false>
static int launch_euro_t(optmc_ctx *ctx, EuroArgs &A, cudaStream_t st, int *grid_out) {
    auto kern = k_european<KIND, ANTI, NIMPL, SCHED, BGEN>;
    int grid = euro_grid(ctx, kern, A.n_local);
    kern<<<grid, ctx->euro_threads, 0, st>>>(A);
    ctx->launches++;
    CK(cudaGetLastError());
    *grid_out = grid;
    return 0;
}

// One block per SM on the lane-replicated tables (european.cuh: k_european_wide).
template <int KIND, bool ANTI, bool SCHED, bool BGEN, int THREADS>
static int launch_euro_wide_t(optmc_ctx *ctx, EuroArgs &A, cudaStream_t st, int *grid_out) {
    auto kern = k_european_wide<KIND, ANTI, SCHED, BGEN, THREADS>;
    const int smem = (int)wide_smem_bytes<KIND>();
    static bool attr_set[64] = {};                 // per device
    if (!attr_set[ctx->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[ctx->device & 63] = true;
    }
    const int64_t need = (A.n_local + THREADS - 1) / THREADS;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, ctx->sm_count));
    kern<<<grid, THREADS, smem, st>>>(A);
    ctx->launches++;
    CK(cudaGetLastError());
    *grid_out = grid;
    return 0;
}

template <int KIND, bool ANTI, bool SCHED = false, bool BGEN = false>
static int launch_euro_wide(optmc_ctx *ctx, EuroArgs &A, cudaStream_t st, int *grid_out) {
    constexpr int T = wide_threads(KIND, SCHED);
    if constexpr (KIND == OPTMC_HESTON || KIND == OPTMC_BATES) {
        // rows of sm_count * threads draws, one draw per thread per row; the last row costs a
        // whole row's time however few draws it holds.  1024 threads run ~2 % slower per draw.
        auto rows = [&](int t) { return (A.n_local + (int64_t)ctx->sm_count * t - 1) / ((int64_t)ctx->sm_count * t); };
        if (ctx->euro_threads_auto && rows(1024) * 1024 * 102 < rows(T) * T * 100)
            return launch_euro_wide_t<KIND, ANTI, SCHED, BGEN, 1024>(ctx, A, st, grid_out);
    }
    return launch_euro_wide_t<KIND, ANTI, SCHED, BGEN, T>(ctx, A, st, grid_out);
}

// Which kernel prices a job: a property of the whole job (n_draws), not of this rank's shard,
// so that every rank of a partitioned pricing runs the same code.
static bool use_wide(const optmc_ctx *ctx, int kind, int64_t n_draws) {
    return ctx->euro_wide && ctx->normals_impl == 1 && kind != OPTMC_DUPIRE &&
           n_draws >= (int64_t)ctx->sm_count * wide_threads(kind);
}

template <int KIND>
static int launch_euro_k(optmc_ctx *ctx, EuroArgs &A, bool anti, bool wide, cudaStream_t st,
                         int *grid) {
    if (ctx->normals_impl == 0)
        return anti ? launch_euro_t<KIND, true, 0>(ctx, A, st, grid)
                    : launch_euro_t<KIND, false, 0>(ctx, A, st, grid);
#define EURO_GO(SCHED_, BGEN_)                                                               \
    do {                                                                                     \
        if constexpr (KIND != OPTMC_DUPIRE) {                                                \
            if (wide)                                                                        \
                return anti ? launch_euro_wide<KIND, true, SCHED_, BGEN_>(ctx, A, st, grid)  \
                            : launch_euro_wide<KIND, false, SCHED_, BGEN_>(ctx, A, st, grid);\
        }                                                                                    \
        return anti ? launch_euro_t<KIND, true, 1, SCHED_, BGEN_>(ctx, A, st, grid)          \
                    : launch_euro_t<KIND, false, 1, SCHED_, BGEN_>(ctx, A, st, grid);        \
    } while (0)
    if constexpr (kind_sabr(KIND)) {
        // beta neither 0 nor 1: the variant whose step has the general exp(ln v + beta ln s) only
        const bool bgen = A.c.beta_mode != 1 && A.c.beta_mode != 2;
        // SABR-jump (jumps act on S inside the step loop): the jump schedule while lambda T is small
        if constexpr (KIND == OPTMC_SABR_JUMP) {
            if (ctx->jump_schedule && use_jump_schedule(A)) {
                if (bgen) EURO_GO(true, true);
                EURO_GO(true, false);
            }
        }
        if (bgen) EURO_GO(false, true);
    }
    EURO_GO(false, false);
#undef EURO_GO
}

template <int KIND>
static int launch_euro_f32(optmc_ctx *ctx, EuroArgs &A, bool anti, cudaStream_t st, int *grid_out) {
    const Consts &c = A.c;
    ConstsF F;
    F.sqrt_dt = (float)c.sqrt_dt;
    F.m11 = (float)c.m11; F.m12 = (float)c.m12; F.m21 = (float)c.m21; F.m22 = (float)c.m22;
    F.one_minus_kappa_dt = (float)c.one_minus_kappa_dt;
    F.kappa_theta_dt = (float)c.kappa_theta_dt;
    F.alpha = (float)c.alpha; F.beta = (float)c.beta;
    F.neg_half_alpha2_dt = (float)c.neg_half_alpha2_dt;
    F.log_v0 = (float)std::max(c.log_v0, -80.0);      // exp(-80) is 0 to every float that meets it
    F.rqc_dt = (float)c.rqc_dt;
    F.beta_mode = c.beta_mode == 3 ? 0 : c.beta_mode;  // beta = 1/2 takes the general exp/log form
    int per_sm = 1;
    const int threads = EURO_THREADS;
    int64_t need = (A.n_local + threads - 1) / threads;
    int grid;
    if (anti) {
        auto kern = k_european_f32<KIND, true>;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(need, (int64_t)ctx->sm_count * per_sm * 2), MAX_GRID));
        kern<<<grid, threads, 0, st>>>(A, F);
    } else {
        auto kern = k_european_f32<KIND, false>;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(need, (int64_t)ctx->sm_count * per_sm * 2), MAX_GRID));
        kern<<<grid, threads, 0, st>>>(A, F);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    *grid_out = grid;
    return 0;
}

static int fill_euro_args(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                          EuroArgs *A) {
    if (!model || !sim) FAIL("null model or sim");
    int rc = check_sim(ctx, sim);
    if (rc) return rc;
    double dt = sim->maturity / sim->n_steps;                                 // monte_carlo.py:209
    double x0 = kind_sabr(model->kind) ? sim->s0 : std::log(sim->s0);          // :217-220
    memset(A, 0, sizeof *A);
    rc = make_consts(ctx, model, x0, dt, sim->n_steps, sim->correlation, &A->c);
    if (rc) return rc;
    philox_expand_key(sim->seed, &A->key);
    A->draw_offset = sim->draw_offset;
    A->n_local = sim->n_local;
    A->n_steps = sim->n_steps;
    poisson_consts(*A);
    return 0;
}

static int sweep_launch(optmc_ctx *ctx, const double *terminal_dev, int64_t n_local, int antithetic,
                        int n_contracts, const double *strikes, const int32_t *is_call,
                        double *moments_dev, cudaStream_t st) {
    if (n_contracts < 1) FAIL("n_contracts must be >= 1");
    if (!terminal_dev || !strikes || !is_call || !moments_dev) FAIL("null pointer");
    double *partials = reinterpret_cast<double *>(ctx->scratch);
    const int gx = (int)std::max<int64_t>(
        1, std::min<int64_t>((n_local + 255) / 256, (int64_t)ctx->sm_count * 2));
    for (int c0 = 0; c0 < n_contracts; c0 += SWEEP_MAX) {
        SweepArgs S;
        memset(&S, 0, sizeof S);
        S.terminal = terminal_dev;
        S.n_local = n_local;
        S.antithetic = antithetic;
        S.n_contracts = std::min(SWEEP_MAX, n_contracts - c0);
        for (int k = 0; k < S.n_contracts; ++k) {
            S.strikes[k] = strikes[c0 + k];
            S.is_call[k] = is_call[c0 + k] ? 1 : 0;
        }
        S.partials = partials;
        k_payoff_sweep<<<dim3(gx, (S.n_contracts + SWEEP_G - 1) / SWEEP_G), 256, 0, st>>>(S);
        ctx->launches++;
        CK(cudaGetLastError());
        k_finish_sweep<<<S.n_contracts, 256, 0, st>>>(partials, gx, (double)n_local,
                                                       moments_dev + (size_t)c0 * OPTMC_NMOM);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int optmc_european_aux_async(optmc_ctx *ctx, const optmc_model *model,
                                        const optmc_sim *sim, int32_t n_contracts,
                                        const double *strikes, const int32_t *is_call,
                                        double *moments_dev, double *terminal_dev, double *aux_dev,
                                        void *stream);

extern "C" int optmc_european_async(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                                    int32_t n_contracts, const double *strikes,
                                    const int32_t *is_call, double *moments_dev,
                                    double *terminal_dev, void *stream) {
    return optmc_european_aux_async(ctx, model, sim, n_contracts, strikes, is_call, moments_dev,
                                    terminal_dev, nullptr, stream);
}

extern "C" int optmc_european_aux_async(optmc_ctx *ctx, const optmc_model *model,
                                        const optmc_sim *sim, int32_t n_contracts,
                                        const double *strikes, const int32_t *is_call,
                                        double *moments_dev, double *terminal_dev, double *aux_dev,
                                        void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (aux_dev && model && !(model->kind == OPTMC_BSM || model->kind == OPTMC_MERTON ||
                              model->kind == OPTMC_KOU || model->kind == OPTMC_BATES))
        FAIL("aux_dev exists for BSM, Merton, Kou (W, J) and Bates (J) only");
    if (aux_dev && sim && sim->precision != OPTMC_FP64) FAIL("aux_dev needs the fp64 mode");
    if (n_contracts < 1 || !strikes || !is_call || !moments_dev) FAIL("bad contract arguments");
    if (n_contracts > 1 && !terminal_dev)
        FAIL("terminal_dev is required when n_contracts > 1 (strike sweep)");
    EuroArgs A;
    int rc = fill_euro_args(ctx, model, sim, &A);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    A.strike = strikes[0];
    A.is_call = is_call[0] ? 1 : 0;
    A.partials = reinterpret_cast<double *>(ctx->scratch);
    A.terminal = terminal_dev;
    A.aux = aux_dev;
    int grid = 0;
    const bool anti = sim->antithetic != 0;
    if (sim->precision != OPTMC_FP64 && sim->precision != OPTMC_FP32) FAIL("unknown precision");
    if (sim->precision == OPTMC_FP32 && model->kind == OPTMC_DUPIRE)
        FAIL("the fp32 mode is not available for Dupire local volatility");
    if (sim->n_local > 0 && sim->precision == OPTMC_FP32) {
        switch (model->kind) {
            case OPTMC_BSM: rc = launch_euro_f32<OPTMC_BSM>(ctx, A, anti, st, &grid); break;
            case OPTMC_HESTON: rc = launch_euro_f32<OPTMC_HESTON>(ctx, A, anti, st, &grid); break;
            case OPTMC_MERTON: rc = launch_euro_f32<OPTMC_MERTON>(ctx, A, anti, st, &grid); break;
            case OPTMC_BATES: rc = launch_euro_f32<OPTMC_BATES>(ctx, A, anti, st, &grid); break;
            case OPTMC_SABR: rc = launch_euro_f32<OPTMC_SABR>(ctx, A, anti, st, &grid); break;
            case OPTMC_SABR_JUMP: rc = launch_euro_f32<OPTMC_SABR_JUMP>(ctx, A, anti, st, &grid); break;
            case OPTMC_KOU: rc = launch_euro_f32<OPTMC_KOU>(ctx, A, anti, st, &grid); break;
            default: FAIL("unknown model kind");
        }
        if (rc) return rc;
    } else if (sim->n_local > 0) {
        const bool wide = use_wide(ctx, model->kind, sim->n_draws);
#define GO(K) rc = launch_euro_k<K>(ctx, A, anti, wide, st, &grid)
        DISPATCH_KIND(model->kind, GO)
#undef GO
        if (rc) return rc;
    }
    if (n_contracts == 1) {
        k_finish_moments<<<1, 256, 0, st>>>(A.partials, grid, (double)sim->n_local, moments_dev);
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    return sweep_launch(ctx, terminal_dev, sim->n_local, anti ? 1 : 0, n_contracts, strikes,
                        is_call, moments_dev, st);
}

// ---------------------------------------------------------------------------
// spot scenarios on common draws (SABR kinds)
// ---------------------------------------------------------------------------
template <int KIND, bool ANTI, int NS, bool SCHED, bool BGEN>
static int launch_scen_t(optmc_ctx *ctx, EuroArgs &A, cudaStream_t st, int *grid_out) {
    auto kern = k_european_scen<KIND, ANTI, NS, SCHED, BGEN>;
    const int smem = (int)(ScenLayout::bytes(true) + POW_C_BYTES + POW_E_BYTES);
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[ctx->device & 63] = true;
    }
    constexpr int threads = scen_threads(SCHED);
    const int64_t need = (A.n_local + threads - 1) / threads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, ctx->sm_count));
    kern<<<grid, threads, smem, st>>>(A);
    ctx->launches++;
    CK(cudaGetLastError());
    *grid_out = grid;
    return 0;
}

template <int KIND, int NS>
static int launch_scen_k(optmc_ctx *ctx, EuroArgs &A, bool anti, cudaStream_t st, int *grid) {
    const bool bgen = A.c.beta_mode != 1 && A.c.beta_mode != 2;
    bool sched = false;
    if constexpr (KIND == OPTMC_SABR_JUMP) sched = ctx->jump_schedule && use_jump_schedule(A);
#define SCEN_GO(SCHED_, BGEN_)                                                            \
    return anti ? launch_scen_t<KIND, true, NS, SCHED_, BGEN_>(ctx, A, st, grid)         \
                : launch_scen_t<KIND, false, NS, SCHED_, BGEN_>(ctx, A, st, grid)
    if constexpr (KIND == OPTMC_SABR_JUMP) {
        if (sched && bgen) SCEN_GO(true, true);
        if (sched) SCEN_GO(true, false);
    }
    if (bgen) SCEN_GO(false, true);
    SCEN_GO(false, false);
#undef SCEN_GO
}

extern "C" int optmc_european_scen_async(optmc_ctx *ctx, const optmc_model *model,
                                         const optmc_sim *sim, int32_t n_scen,
                                         const double *spot_factors, double strike,
                                         int32_t is_call, double *moments_dev, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!model || !kind_sabr(model->kind))
        FAIL("spot scenarios on common draws exist for the SABR kinds (the log-Euler models "
             "get bumped prices by strike rescaling: optmc_payoff_sweep_async)");
    if (n_scen < 2 || n_scen > 3 || !spot_factors || !moments_dev) FAIL("bad scenario arguments");
    if (sim && sim->precision != OPTMC_FP64) FAIL("spot scenarios need the fp64 mode");
    if (ctx->normals_impl != 1) FAIL("spot scenarios need normals_impl = 1");
    EuroArgs A;
    int rc = fill_euro_args(ctx, model, sim, &A);
    if (rc) return rc;
    for (int k = 0; k < n_scen; ++k) {
        if (!(spot_factors[k] > 0.0)) FAIL("spot factors must be positive");
        A.c.scen[k] = spot_factors[k];
    }
    cudaStream_t st = (cudaStream_t)stream;
    A.strike = strike;
    A.is_call = is_call ? 1 : 0;
    A.partials = reinterpret_cast<double *>(ctx->scratch);
    int grid = 1;
    const bool anti = sim->antithetic != 0;
    if (sim->n_local > 0) {
        if (model->kind == OPTMC_SABR)
            rc = n_scen == 2 ? launch_scen_k<OPTMC_SABR, 2>(ctx, A, anti, st, &grid)
                             : launch_scen_k<OPTMC_SABR, 3>(ctx, A, anti, st, &grid);
        else
            rc = n_scen == 2 ? launch_scen_k<OPTMC_SABR_JUMP, 2>(ctx, A, anti, st, &grid)
                             : launch_scen_k<OPTMC_SABR_JUMP, 3>(ctx, A, anti, st, &grid);
        if (rc) return rc;
    } else {
        CK(cudaMemsetAsync(A.partials, 0, sizeof(double) * NPART * n_scen, st));
    }
    k_finish_sweep<<<n_scen, 256, 0, st>>>(A.partials, grid, (double)sim->n_local, moments_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int optmc_european(optmc_ctx *ctx, const optmc_model *model, const optmc_sim *sim,
                              int32_t n_contracts, const double *strikes, const int32_t *is_call,
                              double *moments_host, double *terminal_dev, void *stream) {
    if (!ctx) return -2;
    if (!moments_host) FAIL("moments_host is null");
    if (n_contracts > SWEEP_MAX) FAIL("optmc_european handles at most 64 contracts per call");
    int rc = optmc_european_async(ctx, model, sim, n_contracts, strikes, is_call, ctx->moments_dev,
                                  terminal_dev, stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = sizeof(double) * OPTMC_NMOM * (size_t)n_contracts;
    CK(cudaMemcpyAsync(ctx->moments_host, ctx->moments_dev, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(moments_host, ctx->moments_host, bytes);
    return 0;
}

extern "C" int optmc_sigma_sweep_async(optmc_ctx *ctx, const double *aux_dev, int64_t n_local,
                                       int32_t antithetic, int32_t n_scenarios, const double *a0,
                                       const double *sigma, const double *strikes,
                                       const int32_t *is_call, double *moments_dev, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!aux_dev || !a0 || !sigma || !strikes || !is_call || !moments_dev) FAIL("null pointer");
    if (n_scenarios < 1 || n_scenarios > SCEN_MAX) FAIL("n_scenarios must be in 1..16");
    if (n_local < 0) FAIL("negative n_local");
    SigmaSweepArgs S;
    memset(&S, 0, sizeof S);
    S.aux = aux_dev;
    S.n_local = n_local;
    S.antithetic = antithetic;
    S.n_scen = n_scenarios;
    for (int k = 0; k < n_scenarios; ++k) {
        S.a0[k] = a0[k];
        S.sigma[k] = sigma[k];
        S.strikes[k] = strikes[k];
        S.is_call[k] = is_call[k] ? 1 : 0;
    }
    S.partials = reinterpret_cast<double *>(ctx->scratch);
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (int)std::max<int64_t>(
        1, std::min<int64_t>((n_local + 255) / 256, (int64_t)ctx->sm_count * 2));
    k_sigma_sweep<<<dim3(gx, n_scenarios), 256, 0, st>>>(S);
    ctx->launches++;
    CK(cudaGetLastError());
    k_finish_sweep<<<n_scenarios, 256, 0, st>>>(S.partials, gx, (double)n_local, moments_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int optmc_payoff_sweep_async(optmc_ctx *ctx, const double *terminal_dev, int64_t n_local,
                                        int32_t antithetic, int32_t n_contracts,
                                        const double *strikes, const int32_t *is_call,
                                        double *moments_dev, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    return sweep_launch(ctx, terminal_dev, n_local, antithetic, n_contracts, strikes, is_call,
                        moments_dev, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// single-step terminal samplers (Levy / exact-sampler branches)
// ---------------------------------------------------------------------------
extern "C" int optmc_levy_gig_setup(optmc_levy *L) {
    if (!L) return -2;
    const double b = L->p[0], lm1 = L->p[5], lam = lm1 + 1.0;
    if (!(b > 0.0) || !(lam >= 1.0 || b > 1.0)) return -2;
    // Hoermann & Leydold (2014), Algorithm "ratio-of-uniforms with mode shift"
    const double m = (std::sqrt(lm1 * lm1 + b * b) + lm1) / b;
    auto logg = [&](double x) { return lm1 * std::log(x) - 0.5 * b * (x + 1.0 / x); };
    const double a2 = -(2.0 * (lam + 1.0) / b + m);
    const double a1 = 2.0 * lm1 * m / b - 1.0;
    const double a0 = m;
    const double pp = a1 - a2 * a2 / 3.0;
    const double qq = 2.0 * a2 * a2 * a2 / 27.0 - a2 * a1 / 3.0 + a0;
    double arg = -0.5 * qq * std::sqrt(-27.0 / (pp * pp * pp));
    arg = std::min(1.0, std::max(-1.0, arg));
    const double phi = std::acos(arg);
    const double pi = 3.14159265358979323846;
    const double amp = std::sqrt(-4.0 * pp / 3.0);
    const double xm = amp * std::cos(phi / 3.0 + 4.0 * pi / 3.0) - a2 / 3.0;
    const double xp = amp * std::cos(phi / 3.0) - a2 / 3.0;
    const double lgm = logg(m);
    L->p[6] = m;
    L->p[7] = (xm - m) * std::exp(0.5 * (logg(xm) - lgm));
    L->p[8] = (xp - m) * std::exp(0.5 * (logg(xp) - lgm));
    L->p[9] = lgm;
    return (std::isfinite(L->p[7]) && std::isfinite(L->p[8]) && L->p[7] < 0.0 && L->p[8] > 0.0) ? 0 : -2;
}

extern "C" int optmc_levy_terminal_async(optmc_ctx *ctx, const optmc_levy *levy, double s0,
                                         double drift, uint64_t seed, int64_t offset,
                                         int64_t n_local, int32_t n_contracts,
                                         const double *strikes, const int32_t *is_call,
                                         double *moments_dev, double *terminal_dev, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!levy) FAIL("null levy");
    if (levy->kind < OPTMC_LEVY_VG || levy->kind > OPTMC_LEVY_LOGNORMAL) FAIL("unknown levy kind");
    if (n_contracts < 1 || !strikes || !is_call || !moments_dev) FAIL("bad contract arguments");
    if (n_contracts > 1 && !terminal_dev)
        FAIL("terminal_dev is required when n_contracts > 1 (strike sweep)");
    if (!(s0 > 0.0)) FAIL("s0 must be positive");
    if (offset < 0 || n_local < 0) FAIL("negative sample range");
    const double *p = levy->p;
    switch (levy->kind) {
        case OPTMC_LEVY_VG:
            if (!(p[0] > 0.0 && p[1] > 0.0 && p[3] > 0.0)) FAIL("VG needs T/nu, nu, sigma > 0");
            break;
        case OPTMC_LEVY_NIG:
            if (!(p[0] > 0.0)) FAIL("NIG needs mu_ig > 0");
            break;
        case OPTMC_LEVY_CGMY1:
            if (!(p[0] > 0.0 && p[1] > 0.0 && p[2] > 0.0)) FAIL("CGMY needs C T, M, G > 0");
            break;
        case OPTMC_LEVY_HYPERBOLIC:
            if (!(p[0] > 0.0 && p[1] > 0.0 && p[4] > 0.0 && p[7] < 0.0 && p[8] > 0.0))
                FAIL("hyperbolic: bad GIG constants (call optmc_levy_gig_setup)");
            break;
        case OPTMC_LEVY_CEV:
            if (!(p[0] > 1.0 && p[1] >= 0.0 && p[2] > 0.0))
                FAIL("CEV needs df > 1, nc >= 0, k > 0 (gamma < 1 and r != 0)");
            break;
        default:
            if (!(p[1] > 0.0)) FAIL("lognormal needs sigma sqrt(T) > 0");
    }
    LevyArgs A;
    memset(&A, 0, sizeof A);
    A.kind = levy->kind;
    memcpy(A.p, levy->p, sizeof A.p);
    philox_expand_key(seed, &A.key);
    A.offset = offset;
    A.n_local = n_local;
    A.s0 = s0;
    A.drift = drift;
    A.is_call = is_call[0] ? 1 : 0;
    A.strike = strikes[0];
    A.partials = reinterpret_cast<double *>(ctx->scratch);
    A.terminal = terminal_dev;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)std::max<int64_t>(
        1, std::min<int64_t>((n_local + 255) / 256, (int64_t)ctx->sm_count * 8));
    k_levy_terminal<<<grid, 256, 0, st>>>(A);
    ctx->launches++;
    CK(cudaGetLastError());
    if (n_contracts == 1) {
        k_finish_moments<<<1, 256, 0, st>>>(A.partials, grid, (double)n_local, moments_dev);
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    return sweep_launch(ctx, terminal_dev, n_local, 0, n_contracts, strikes, is_call, moments_dev, st);
}

// ---------------------------------------------------------------------------
// fed-draw kernels
// ---------------------------------------------------------------------------
extern "C" int64_t optmc_fed_workspace_bytes(int32_t kind, int64_t n_paths, int32_t n_steps) {
    if (!kind_jumps(kind) || n_paths <= 0 || n_steps <= 0) return 8;
    int64_t n_blocks = (n_paths + 31) / 32;        // the finer of the two kernels' jump blocks
    return n_blocks * (int64_t)n_steps * (int64_t)sizeof(int64_t);
}

extern "C" int optmc_fed_async(optmc_ctx *ctx, const optmc_model *model, double x0, double dt,
                               int64_t n_paths, int32_t n_steps, const double *dw1_dev,
                               const double *dw2_dev, const int64_t *counts_dev,
                               const double *jumps_dev, int64_t n_jumps, double sign,
                               int32_t path_sum_mode, double *terminal_dev, double *paths_dev,
                               void *workspace_dev, int64_t workspace_bytes, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!model) FAIL("null model");
    if (n_paths < 0 || n_steps < 1) FAIL("bad n_paths / n_steps");
    if (!dw1_dev) FAIL("dw1_dev is null");
    if (kind_two_factor(model->kind) && !dw2_dev) FAIL("dw2_dev is required for two-factor models");
    if (kind_jumps(model->kind) && !counts_dev) FAIL("counts_dev is required for jump models");
    if (kind_jumps(model->kind) && n_jumps > 0 && !jumps_dev) FAIL("jumps_dev is null");
    if (!terminal_dev && !paths_dev) FAIL("no output requested");
    if (n_paths == 0) return 0;
    FedArgs A;
    memset(&A, 0, sizeof A);
    int rc = make_consts(ctx, model, x0, dt, n_steps, OPTMC_CORR_INDEPENDENT_DRAWS, &A.c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    A.n_paths = n_paths;
    A.n_steps = n_steps;
    A.dw1 = dw1_dev;
    A.dw2 = dw2_dev;
    A.counts = counts_dev;
    A.jumps = jumps_dev;
    A.n_jumps = n_jumps;
    A.sign = sign;
    A.sum_mode = path_sum_mode;
    A.terminal = terminal_dev;
    A.paths = paths_dev;
    const int64_t tiled_blocks = (n_paths + FED_P - 1) / FED_P;
    A.n_blocks = ctx->fed_stream ? (n_paths + 31) / 32 : tiled_blocks;     // jump blocks
    if (tiled_blocks > 0x7fffffffLL) FAIL("too many paths for one fed call");
    if (kind_jumps(model->kind)) {
        int64_t need = optmc_fed_workspace_bytes(model->kind, n_paths, n_steps);
        if (!workspace_dev || workspace_bytes < need) FAIL("fed workspace too small");
        A.base = reinterpret_cast<int64_t *>(workspace_dev);
        if (ctx->fed_stream) {
            CK(cudaMemsetAsync(A.base, 0, (size_t)(A.n_blocks * n_steps) * sizeof(int64_t), st));
            const unsigned g = (unsigned)std::max<int64_t>(
                1, std::min<int64_t>((n_paths + 255) / 256, (int64_t)ctx->sm_count * 32));
            k_fed_block_totals_stream<<<g, 256, 0, st>>>(counts_dev, n_paths, n_steps, A.n_blocks,
                                                         A.base);
        } else {
            k_fed_block_totals<<<(unsigned)A.n_blocks, FED_P, 0, st>>>(counts_dev, n_paths, n_steps,
                                                                       A.n_blocks, A.base);
        }
        ctx->launches++;
        CK(cudaGetLastError());
        {
            // exclusive scan over (step, block): a library device scan (temporary storage carved
            // from the context's scratch), else the single-block kernel
            const int64_t items = A.n_blocks * (int64_t)n_steps;
            size_t tmp_bytes = 0;
            cudaError_t qe = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, A.base, A.base, items, st);
            if (qe == cudaSuccess && tmp_bytes <= ctx->scratch_bytes && items < 0x7fffffffLL) {
                CK(cub::DeviceScan::ExclusiveSum(ctx->scratch, tmp_bytes, A.base, A.base, items, st));
            } else {
                k_scan_exclusive<<<1, 1024, 0, st>>>(A.base, items);
            }
        }
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (ctx->fed_stream) {
        // one thread streams one row (k_fed_stream)
        const int64_t nb = (n_paths + FEDS_THREADS - 1) / FEDS_THREADS;
        const unsigned grid = (unsigned)std::max<int64_t>(
            1, std::min<int64_t>(nb, (int64_t)ctx->sm_count * 4 * 8));
#define GOS(K)                                                          \
    if (A.paths) k_fed_stream<K, true><<<grid, FEDS_THREADS, 0, st>>>(A);  \
    else k_fed_stream<K, false><<<grid, FEDS_THREADS, 0, st>>>(A)
        DISPATCH_KIND(model->kind, GOS)
#undef GOS
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
    }
#define GO(K) k_fed<K><<<(unsigned)A.n_blocks, FED_P, fed_smem_bytes(), st>>>(A)
    DISPATCH_KIND(model->kind, GO)
#undef GO
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// Longstaff-Schwartz
// ---------------------------------------------------------------------------
template <int KIND, bool ANTI, int NIMPL, bool SCHED = false, bool BGEN = false>
static int launch_paths_t(optmc_ctx *ctx, EuroArgs &A, double *store, int64_t ld, cudaStream_t st) {
    auto kern = k_lsmc_paths<KIND, ANTI, NIMPL, SCHED, BGEN>;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EURO_THREADS, 0) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    int64_t need = (A.n_local + EURO_THREADS - 1) / EURO_THREADS;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)ctx->sm_count * per_sm));
    kern<<<grid, EURO_THREADS, 0, st>>>(A, store, ld);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int KIND, bool ANTI, bool SCHED = false, bool BGEN = false>
static int launch_paths_wide(optmc_ctx *ctx, EuroArgs &A, double *store, int64_t ld, cudaStream_t st) {
    auto kern = k_lsmc_paths_wide<KIND, ANTI, SCHED, BGEN>;
    constexpr int threads = paths_wide_threads(KIND, SCHED);
    const int smem = (int)paths_wide_smem_bytes<KIND>();
    static bool attr_set[64] = {};                 // per device
    if (!attr_set[ctx->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[ctx->device & 63] = true;
    }
    const int64_t need = (A.n_local + threads - 1) / threads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, ctx->sm_count));
    kern<<<grid, threads, smem, st>>>(A, store, ld);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int KIND>
static int launch_paths_k(optmc_ctx *ctx, EuroArgs &A, bool anti, bool wide, double *store,
                          int64_t ld, cudaStream_t st) {
    if (ctx->normals_impl == 0) {
        if constexpr (kind_jumps(KIND)) {
            if (use_jump_schedule(A) && ctx->jump_schedule)
                return anti ? launch_paths_t<KIND, true, 0, true>(ctx, A, store, ld, st)
                            : launch_paths_t<KIND, false, 0, true>(ctx, A, store, ld, st);
        }
        return anti ? launch_paths_t<KIND, true, 0>(ctx, A, store, ld, st)
                    : launch_paths_t<KIND, false, 0>(ctx, A, store, ld, st);
    }
#define PATHS_GO(SCHED_, BGEN_)                                                                   \
    do {                                                                                          \
        if constexpr (KIND != OPTMC_DUPIRE) {                                                     \
            if (wide)                                                                             \
                return anti ? launch_paths_wide<KIND, true, SCHED_, BGEN_>(ctx, A, store, ld, st) \
                            : launch_paths_wide<KIND, false, SCHED_, BGEN_>(ctx, A, store, ld, st);\
        }                                                                                         \
        return anti ? launch_paths_t<KIND, true, 1, SCHED_, BGEN_>(ctx, A, store, ld, st)         \
                    : launch_paths_t<KIND, false, 1, SCHED_, BGEN_>(ctx, A, store, ld, st);       \
    } while (0)
    // SABR kinds, beta neither 0 nor 1: the step without the switch
    bool bgen = false;
    if constexpr (kind_sabr(KIND)) bgen = A.c.beta_mode != 1 && A.c.beta_mode != 2;
    if constexpr (kind_jumps(KIND)) {
        if (use_jump_schedule(A) && ctx->jump_schedule) {
            if constexpr (kind_sabr(KIND)) {
                if (bgen) PATHS_GO(true, true);
            }
            PATHS_GO(true, false);
        }
    }
    if constexpr (kind_sabr(KIND)) {
        if (bgen) PATHS_GO(false, true);
    }
    PATHS_GO(false, false);
#undef PATHS_GO
}

extern "C" int optmc_lsmc_paths_async(optmc_ctx *ctx, const optmc_model *model,
                                      const optmc_sim *sim, double *store_dev, int64_t ld,
                                      void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!store_dev) FAIL("store_dev is null");
    if (sim && sim->precision != OPTMC_FP64) FAIL("the path store is fp64 only");
    if (model && model->kind == OPTMC_DUPIRE)
        FAIL("no path kernel for Dupire: American MC pricing requires an SDE model "
             "(american_monte_carlo.py:75-150)");
    EuroArgs A;
    int rc = fill_euro_args(ctx, model, sim, &A);
    if (rc) return rc;
    const bool anti = sim->antithetic != 0;
    if (ld < sim->n_local * (anti ? 2 : 1)) FAIL("ld smaller than the local path count");
    if (sim->n_local == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide = use_wide(ctx, model->kind, sim->n_draws);
#define GO(K) rc = launch_paths_k<K>(ctx, A, anti, wide, store_dev, ld, st)
    DISPATCH_KIND(model->kind, GO)
#undef GO
    return rc;
}

extern "C" int optmc_lsmc_load_async(optmc_ctx *ctx, const double *paths_dev, int64_t n_paths,
                                     int32_t n_steps, double *store_dev, int64_t ld, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!paths_dev || !store_dev) FAIL("null pointer");
    if (n_paths < 1 || n_steps < 1 || ld < n_paths) FAIL("bad shape");
    dim3 grid((unsigned)((n_paths + 31) / 32), (unsigned)((n_steps + 31) / 32));
    k_lsmc_load<<<grid, 256, 0, (cudaStream_t)stream>>>(paths_dev, n_paths, n_steps, store_dev, ld);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int lsmc_grid(const optmc_ctx *ctx, int64_t n) {
    int64_t need = (n / 2 + LSMC_THREADS - 1) / LSMC_THREADS;   // two paths per thread per trip
    return (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)ctx->sm_count * 4));
}

template <int D, bool INIT>
static void launch_step(bool final_, int grid, cudaStream_t st, const double *row_dec,
                        const double *row_acc, int64_t n, double K, int is_call, double disc,
                        const double *gram_in, double *cash, double *partials,
                        unsigned int *ticket, double *gram_out) {
    if (final_)
        k_lsmc_step<D, INIT, true><<<grid, LSMC_THREADS, 0, st>>>(
            row_dec, row_acc, n, K, 1.0 / K, is_call, disc, gram_in, cash, partials, ticket, gram_out);
    else
        k_lsmc_step<D, INIT, false><<<grid, LSMC_THREADS, 0, st>>>(
            row_dec, row_acc, n, K, 1.0 / K, is_call, disc, gram_in, cash, partials, ticket, gram_out);
}

extern "C" int optmc_lsmc_step_async(optmc_ctx *ctx, const double *store_dev, int64_t ld,
                                     int64_t n_paths, int32_t n_steps, int32_t date,
                                     double strike, int32_t is_call, double discount,
                                     int32_t degree, double *cashflow_dev, double *gram_dev,
                                     void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!store_dev || !cashflow_dev || !gram_dev) FAIL("null pointer");
    if (degree < 1 || degree > OPTMC_LSMC_MAX_DEGREE) FAIL("degree must be in 1..7");
    if (n_paths < 1 || n_steps < 1 || ld < n_paths || (ld & 1)) FAIL("bad shape (ld must be even)");
    if (date < 0 || date > n_steps - 1) FAIL("date must be in 0..n_steps-1");
    if (!(strike > 0.0)) FAIL("strike must be positive");
    const bool first = date == n_steps - 1;
    const double *row_dec = store_dev + (size_t)date * ld;                     // date + 1
    const double *row_acc = date >= 1 ? store_dev + (size_t)(date - 1) * ld : nullptr;
    const double *gram_in = first ? nullptr : gram_dev + (size_t)(date + 1) * OPTMC_LSMC_GRAM;
    double *gram_out = gram_dev + (size_t)date * OPTMC_LSMC_GRAM;
    double *partials = reinterpret_cast<double *>(ctx->scratch);
    int grid = lsmc_grid(ctx, n_paths);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(gram_out, 0, sizeof(double) * OPTMC_LSMC_GRAM, st));
#define STEP(D)                                                                                  \
    case D:                                                                                      \
        if (first)                                                                               \
            launch_step<D, true>(final_, grid, st, row_dec, row_acc, n_paths, strike, is_call,   \
                                 discount, gram_in, cashflow_dev, partials, ctx->ticket,         \
                                 gram_out);                                                      \
        else                                                                                     \
            launch_step<D, false>(final_, grid, st, row_dec, row_acc, n_paths, strike, is_call,  \
                                  discount, gram_in, cashflow_dev, partials, ctx->ticket,        \
                                  gram_out);                                                     \
        break;
    const bool final_ = date == 0;
    switch (degree) { STEP(1) STEP(2) STEP(3) STEP(4) STEP(5) STEP(6) STEP(7) }
#undef STEP
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int D>
static int launch_persistent(optmc_ctx *ctx, LsmcPersistArgs &A, cudaStream_t st) {
    if (ctx->sm_count > 160) FAIL("k_lsmc_persistent is sized for at most 160 SMs");
    void *args[] = {&A};
    if constexpr (lsmc_has_ring(D)) {
        // large shards: the sweep fed by a cp.async ring in shared memory (lsmc.cuh)
        // (measured per date, tools/lsmc_phases.py: 2^19 paths 5.4 us against 6.1 without, 2^18 the
        // same, 2^17 4.35 against 4.25.)  The ring addresses a block's slice by 32-bit byte offsets.
        const bool slice_fits = (A.n / 2 / ctx->sm_count + 16) * 16 < ((int64_t)1 << 31);
        if (ctx->lsmc_ring != 0 && slice_fits && (ctx->lsmc_ring == 1 || A.n >= (int64_t)1 << 19)) {
            const size_t ring = lsmc_ring_bytes(D);
            static bool attr_set[64] = {};
            if (!attr_set[ctx->device & 63]) {
                CK(cudaFuncSetAttribute(k_lsmc_persistent<D, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring));
                attr_set[ctx->device & 63] = true;
            }
            CK(cudaLaunchCooperativeKernel((const void *)k_lsmc_persistent<D, true>,
                                           dim3(ctx->sm_count), dim3(lsmc_p_threads(D)), args, ring,
                                           st));
            ctx->launches++;
            return 0;
        }
    }
    CK(cudaLaunchCooperativeKernel((const void *)k_lsmc_persistent<D, false>, dim3(ctx->sm_count),
                                   dim3(lsmc_p_threads(D)), args, 0, st));
    ctx->launches++;
    return 0;
}

// All exercise dates in one cooperative launch (one block per SM, grid barrier per date).
static int lsmc_persistent(optmc_ctx *ctx, const double *store_dev, int64_t ld, int64_t n_paths,
                           int32_t n_steps, double strike, int32_t is_call, double disc,
                           int32_t degree, double *cashflow_dev, double *gram_dev,
                           cudaStream_t st) {
    CK(cudaSetDevice(ctx->device));
    if (!store_dev || !cashflow_dev || !gram_dev) FAIL("null pointer");
    if (degree < 1 || degree > OPTMC_LSMC_MAX_DEGREE) FAIL("degree must be in 1..7");
    if (n_paths < 1 || n_steps < 1 || ld < n_paths || (ld & 1)) FAIL("bad shape (ld must be even)");
    if (!(strike > 0.0)) FAIL("strike must be positive");
    LsmcPersistArgs A;
    A.store = store_dev;
    A.ld = ld;
    A.n = n_paths;
    A.n_steps = n_steps;
    A.K = strike;
    A.inv_K = 1.0 / strike;
    A.disc = disc;
    A.is_call = is_call;
    A.cashflow = cashflow_dev;
    A.prefetch_next_row = (32.0 * (double)n_paths <= 96.0 * 1024 * 1024) ? 1 : 0;
    A.slots = ctx->lsmc_slots;
    A.nonce = (++ctx->lsmc_nonce) & 0xFFFFFu;      // tag 0 is never used: nonce >= 1
    if (A.nonce == 0) A.nonce = ctx->lsmc_nonce = 1;
    A.error = ctx->ticket + 8;
    A.gram = gram_dev;
    A.timing = ctx->lsmc_timing;
    A.timing_all = ctx->lsmc_timing_all;
    const bool joint = ctx->use_peers && ctx->peer_world > 1;
    A.rank = joint ? ctx->peer_rank : 0;
    A.world = joint ? ctx->peer_world : 1;
    for (int r = 0; r < OPTMC_MAX_PEERS; ++r) A.peer_rank_slots[r] = ctx->peer_rank_slots[r];
    A.local_rank_slots = ctx->peer_rank_slots[ctx->peer_rank];
    if (joint) {
        ctx->peer_nonce = (ctx->peer_nonce + 1) & 0xFFFFFu;
        if (ctx->peer_nonce == 0) ctx->peer_nonce = 2;       // keeps the call parity alternating
    }
    A.peer_nonce = ctx->peer_nonce;
    CK(cudaMemsetAsync(ctx->ticket + 8, 0, 4, st));
    int rc = 0;
    switch (degree) {
        case 1: rc = launch_persistent<1>(ctx, A, st); break;
        case 2: rc = launch_persistent<2>(ctx, A, st); break;
        case 3: rc = launch_persistent<3>(ctx, A, st); break;
        case 4: rc = launch_persistent<4>(ctx, A, st); break;
        case 5: rc = launch_persistent<5>(ctx, A, st); break;
        case 6: rc = launch_persistent<6>(ctx, A, st); break;
        case 7: rc = launch_persistent<7>(ctx, A, st); break;
    }
    return rc;
}

extern "C" int optmc_lsmc_peer_export(optmc_ctx *ctx, void *handle_out) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!handle_out) FAIL("handle_out is null");
    static_assert(sizeof(cudaIpcMemHandle_t) == OPTMC_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->rank_slots));
    memcpy(handle_out, &h, sizeof h);
    return 0;
}

extern "C" int optmc_lsmc_peer_attach(optmc_ctx *ctx, int32_t rank, int32_t world,
                                      const void *handles) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (world < 1 || world > OPTMC_MAX_PEERS || rank < 0 || rank >= world) FAIL("bad rank / world");
    for (int r = 0; r < ctx->peer_world; ++r)          // drop an earlier attachment
        if (ctx->peer_world > 1 && r != ctx->peer_rank && ctx->peer_rank_slots[r])
            cudaIpcCloseMemHandle(ctx->peer_rank_slots[r]);
    for (int r = 0; r < OPTMC_MAX_PEERS; ++r) ctx->peer_rank_slots[r] = nullptr;
    ctx->peer_rank = 0;
    ctx->peer_world = 1;
    ctx->peer_nonce = 0;
    ctx->use_peers = 0;
    CK(cudaMemset(ctx->rank_slots, 0, sizeof(uint4) * RANK_SLOTS_TOTAL));
    ctx->euro_peer_nonce = 0;
    if (world == 1) return 0;
    if (!handles) FAIL("handles is null");
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            ctx->peer_rank_slots[r] = ctx->rank_slots;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * OPTMC_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_rank_slots[r] = static_cast<uint4 *>(p);
    }
    ctx->peer_rank = rank;
    ctx->peer_world = world;
    return 0;
}

static int peer_allreduce_launch(optmc_ctx *ctx, double *vals_dev, int32_t n, cudaStream_t st);

extern "C" int optmc_european_allreduce(optmc_ctx *ctx, const optmc_model *model,
                                        const optmc_sim *sim, int32_t n_contracts,
                                        const double *strikes, const int32_t *is_call,
                                        double *moments_host, double *terminal_dev, void *stream) {
    if (!ctx) return -2;
    if (!moments_host) FAIL("moments_host is null");
    if (n_contracts > SWEEP_MAX) FAIL("optmc_european_allreduce handles at most 64 contracts per call");
    if (ctx->peer_world < 2) FAIL("no peers attached (optmc_lsmc_peer_attach)");
    int rc = optmc_european_async(ctx, model, sim, n_contracts, strikes, is_call, ctx->moments_dev,
                                  terminal_dev, stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = peer_allreduce_launch(ctx, ctx->moments_dev, OPTMC_NMOM * n_contracts, st);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * OPTMC_NMOM * (size_t)n_contracts;
    CK(cudaMemcpyAsync(ctx->moments_host, ctx->moments_dev, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(moments_host, ctx->moments_host, bytes);
    return 0;
}

extern "C" int optmc_read_doubles(optmc_ctx *ctx, const double *src_dev, int32_t n,
                                  double *out_host, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (!src_dev || !out_host || n < 1 || n > SWEEP_MAX * OPTMC_NMOM)
        FAIL("bad arguments (1 <= n <= 384)");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(ctx->moments_host, src_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out_host, ctx->moments_host, sizeof(double) * (size_t)n);
    return 0;
}

static int peer_allreduce_launch(optmc_ctx *ctx, double *vals_dev, int32_t n, cudaStream_t st) {
    if (!vals_dev || n < 1 || n > PEER_AR_MAX) FAIL("bad arguments (1 <= n <= 1536)");
    if (ctx->peer_world < 2) FAIL("no peers attached (optmc_lsmc_peer_attach)");
    PeerReduceArgs A;
    memset(&A, 0, sizeof A);
    A.vals = vals_dev;
    A.n = n;
    A.rank = ctx->peer_rank;
    A.world = ctx->peer_world;
    ctx->euro_peer_nonce = ctx->euro_peer_nonce == 0xFFFFFFFFu ? 1u : ctx->euro_peer_nonce + 1u;
    A.tag = ctx->euro_peer_nonce;
    A.euro_off = RANK_SLOTS_LSMC;
    for (int r = 0; r < OPTMC_MAX_PEERS; ++r) A.peer_rank_slots[r] = ctx->peer_rank_slots[r];
    A.local_slots = ctx->peer_rank_slots[ctx->peer_rank];
    A.error = ctx->ticket + 9;
    k_peer_allreduce<<<(n + 255) / 256, 256, 0, st>>>(A);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int optmc_peer_allreduce_async(optmc_ctx *ctx, double *vals_dev, int32_t n,
                                          void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    return peer_allreduce_launch(ctx, vals_dev, n, (cudaStream_t)stream);
}

extern "C" int optmc_lsmc(optmc_ctx *ctx, const double *store_dev, int64_t ld, int64_t n_paths,
                          int32_t n_steps, double strike, int32_t is_call, double r, double dt,
                          int32_t degree, double *cashflow_dev, double *gram_dev, double *out_host,
                          void *stream) {
    if (!ctx) return -2;
    if (!out_host) FAIL("out_host is null");
    const double disc = std::exp(-r * dt);                       // american_mc_kernels.py:54
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->lsmc_mode == 1 || (ctx->use_peers && ctx->peer_world > 1)) {
        int rc = lsmc_persistent(ctx, store_dev, ld, n_paths, n_steps, strike, is_call, disc, degree,
                                 cashflow_dev, gram_dev, st);
        if (rc) return rc;
        unsigned int *flag = reinterpret_cast<unsigned int *>(ctx->moments_host + 8);
        CK(cudaMemcpyAsync(ctx->moments_host, gram_dev, 3 * sizeof(double), cudaMemcpyDeviceToHost,
                           st));
        CK(cudaMemcpyAsync(flag, ctx->ticket + 8, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (*flag) FAIL("k_lsmc_persistent: grid barrier timed out");
        memcpy(out_host, ctx->moments_host, 3 * sizeof(double));
        return 0;
    }
    for (int i = n_steps - 1; i >= 0; --i) {                     // :53 plus the final mean :84
        int rc = optmc_lsmc_step_async(ctx, store_dev, ld, n_paths, n_steps, i, strike, is_call,
                                       disc, degree, cashflow_dev, gram_dev, stream);
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(ctx->moments_host, gram_dev, 3 * sizeof(double), cudaMemcpyDeviceToHost,
                       st));
    CK(cudaStreamSynchronize(st));
    memcpy(out_host, ctx->moments_host, 3 * sizeof(double));
    return 0;
}

// ---------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------
__global__ void k_philox_raw(PhiloxKey key, int64_t n, uint32_t c2, uint32_t c3, uint32_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 w = philox_draw(key, (uint64_t)i, c2, c3);
    reinterpret_cast<uint4 *>(out)[i] = w;
}

template <int NIMPL>
__global__ void k_normals(PhiloxKey key, int64_t n, int step, double *out) {
    __shared__ RngShared<false> tab;
    const RngView tv = fill_rng_shared(&tab, 1.0, 0.0, 0.0, 0.0);
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t wa, wb;
    philox_pair_words(key, (uint64_t)i, (uint32_t)step, wa, wb);
    double a, b;
    if constexpr (NIMPL == 0)
        normal_pair_lib2(wa, wb, a, b);
    else
        normal_pair_fast<false>(wa, wb, a, b, tv);
    out[2 * i] = a;
    out[2 * i + 1] = b;
}

__global__ void __launch_bounds__(256) k_fp64_peak(int64_t iters, double seed, double *sink) {
    // eight independent chains a_k = b_k * b_k + a_k: two distinct 64-bit register
    // operands per DFMA, the form that sustains one warp-DFMA per 2 cycles per
    // sub-partition (three distinct register operands cost 3 cycles on B200;
    // tools/pipe_probe2.cu).
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b0 = 1e-9 * a0, b1 = 1e-9 * a1, b2 = 1e-9 * a2, b3 = 1e-9 * a3;
    double b4 = 1e-9 * a4, b5 = 1e-9 * a5, b6 = 1e-9 * a6, b7 = 1e-9 * a7;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = fma(b0, b0, a0); a1 = fma(b1, b1, a1); a2 = fma(b2, b2, a2); a3 = fma(b3, b3, a3);
        a4 = fma(b4, b4, a4); a5 = fma(b5, b5, a5); a6 = fma(b6, b6, a6); a7 = fma(b7, b7, a7);
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) sink[0] = s;   // never true; keeps the loop alive
}

extern "C" int optmc_philox_raw(optmc_ctx *ctx, uint64_t seed, int64_t n, uint32_t c2, uint32_t c3,
                                uint32_t *out_dev, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (n < 1 || !out_dev) FAIL("bad arguments");
    PhiloxKey key;
    philox_expand_key(seed, &key);
    k_philox_raw<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key, n, c2, c3,
                                                                                out_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int optmc_normals(optmc_ctx *ctx, uint64_t seed, int64_t n, int32_t step,
                             double *out_dev, void *stream) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (n < 1 || !out_dev) FAIL("bad arguments");
    unsigned grid = (unsigned)((n + 255) / 256);
    PhiloxKey key;
    philox_expand_key(seed, &key);
    if (ctx->normals_impl == 0)
        k_normals<0><<<grid, 256, 0, (cudaStream_t)stream>>>(key, n, step, out_dev);
    else
        k_normals<1><<<grid, 256, 0, (cudaStream_t)stream>>>(key, n, step, out_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int optmc_fp64_peak(optmc_ctx *ctx, int64_t iters, double *fma_per_sec,
                               double *elapsed_ms) {
    if (!ctx) return -2;
    CK(cudaSetDevice(ctx->device));
    if (iters < 1 || !fma_per_sec) FAIL("bad arguments");
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = ctx->sm_count * 8;
    double *sink = reinterpret_cast<double *>(ctx->scratch);
    k_fp64_peak<<<grid, 256>>>(std::max<int64_t>(iters / 16, 1), 1.0, sink);   // warm-up
    CK(cudaEventRecord(e0));
    k_fp64_peak<<<grid, 256>>>(iters, 1.0, sink);
    CK(cudaEventRecord(e1));
    ctx->launches += 2;
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *fma_per_sec = (double)grid * 256.0 * 8.0 * (double)iters / (ms * 1e-3);
    if (elapsed_ms) *elapsed_ms = ms;
    return 0;
}
