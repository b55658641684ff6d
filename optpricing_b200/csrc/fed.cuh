// fed.cuh -- fed-draw kernels: the drop-in for the reference's 7 terminal
// kernels (mc_kernels.py) and 7 path kernels (path_kernels.py), consuming the
// reference's own C-order buffers.  Used for bit-level parity (<= 1e-12) and
// for the technique's rng_mode="numpy" compatibility mode.
//
// Layout problem: the reference's dw[n_paths][n_steps] is path-major, so a
// thread-per-path walk is a stride-n_steps access.  Each block stages a tile of
// FED_P paths x FED_T steps through shared memory: global reads and writes are
// coalesced row segments, per-thread reads come from padded shared rows.
//
// Jump sizes arrive as one flat buffer in the order the reference pulls them
// from its RNG: by step, then path, then jump.  The offset of (path p, step i)
// is base[i][block] + (exclusive prefix of counts over the block's paths at
// step i); base comes from k_fed_block_totals + k_scan_exclusive.
#pragma once
#include "models.cuh"

namespace optmc {

constexpr int FED_P = 64;   // paths per block (= threads)
constexpr int FED_T = 32;   // steps per tile
constexpr int FED_LD = FED_T + 1;

struct FedArgs {
    Consts c;
    int64_t n_paths;
    int n_steps;
    const double *dw1, *dw2;
    const int64_t *counts;
    const double *jumps;
    int64_t n_jumps;
    double sign;
    int sum_mode;
    double *terminal;
    double *paths;          // [n_paths][n_steps+1] or null
    int64_t *base;          // [n_steps][n_blocks] exclusive offsets (jump models)
    int64_t n_blocks;
};

constexpr size_t fed_smem_bytes() {
    return sizeof(double) * 2 * FED_P * FED_LD + sizeof(int) * 2 * FED_P * FED_LD;
}

// Coalesced load of a [FED_P][tile_len] tile of a row-major [n][n_steps] array.
template <class T, class U>
__device__ __forceinline__ void load_tile(U (*dst)[FED_LD], const T *src, int64_t row0,
                                          int64_t n_rows, int n_steps, int t0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int rr = warp; rr < FED_P; rr += FED_P / 32) {
        int64_t row = row0 + rr;
        int col = t0 + lane;
        U val = U(0);
        if (row < n_rows && col < n_steps) val = (U)src[(size_t)row * n_steps + col];
        dst[rr][lane] = val;
    }
}

// Per-block, per-step jump totals: tot[i * n_blocks + b].
__global__ void __launch_bounds__(FED_P) k_fed_block_totals(const int64_t *counts, int64_t n_paths,
                                                            int n_steps, int64_t n_blocks,
                                                            int64_t *tot) {
    __shared__ int cnt[FED_P][FED_LD];
    const int64_t row0 = (int64_t)blockIdx.x * FED_P;
    for (int t0 = 0; t0 < n_steps; t0 += FED_T) {
        load_tile<int64_t, int>(cnt, counts, row0, n_paths, n_steps, t0);
        __syncthreads();
        if (threadIdx.x < FED_T && t0 + threadIdx.x < n_steps) {
            int64_t s = 0;
            for (int rr = 0; rr < FED_P; ++rr) s += cnt[rr][threadIdx.x];
            tot[(size_t)(t0 + threadIdx.x) * n_blocks + blockIdx.x] = s;
        }
        __syncthreads();
    }
}

// In-place exclusive scan of n int64 values by one block (n is small: one
// entry per 64 paths per step).
__global__ void __launch_bounds__(1024) k_scan_exclusive(int64_t *a, int64_t n) {
    __shared__ int64_t sums[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t lo = min(n, (int64_t)t * chunk), hi = min(n, lo + chunk);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += a[i];
    sums[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            int64_t v = sums[i];
            sums[i] = run;
            run += v;
        }
    }
    __syncthreads();
    int64_t run = sums[t];
    for (int64_t i = lo; i < hi; ++i) {
        int64_t v = a[i];
        a[i] = run;
        run += v;
    }
}

template <int KIND>
__global__ void __launch_bounds__(FED_P) k_fed(const FedArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double(*t1)[FED_LD] = reinterpret_cast<double(*)[FED_LD]>(smem_raw);
    double(*t2)[FED_LD] = t1 + FED_P;
    int(*cnt)[FED_LD] = reinterpret_cast<int(*)[FED_LD]>(t2 + FED_P);
    int(*pre)[FED_LD] = cnt + FED_P;

    constexpr bool TWO = kind_two_factor(KIND);
    constexpr bool JUMPS = kind_jumps(KIND);
    __shared__ MathShared math;                    // fast_log / fast_exp tables
    const RngView tv = fill_math_shared(&math);
    __syncthreads();
    const Consts &c = A.c;
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * FED_P;
    const int64_t p = row0 + tid;
    const bool live = p < A.n_paths;
    const size_t pld = (size_t)A.n_steps + 1;

    State s;
    s.x = c.x0;
    s.v = c.v0;
    if (A.paths && live)
        A.paths[(size_t)p * pld] = kind_sabr(KIND) ? c.x0 : exp(c.x0);   // path_kernels.py:28

    for (int t0 = 0; t0 < A.n_steps; t0 += FED_T) {
        const int tile_len = min(FED_T, A.n_steps - t0);
        load_tile<double, double>(t1, A.dw1, row0, A.n_paths, A.n_steps, t0);
        if constexpr (TWO) load_tile<double, double>(t2, A.dw2, row0, A.n_paths, A.n_steps, t0);
        if constexpr (JUMPS) load_tile<int64_t, int>(cnt, A.counts, row0, A.n_paths, A.n_steps, t0);
        __syncthreads();
        if constexpr (JUMPS) {
            if (tid < FED_T) {
                int run = 0;
                for (int rr = 0; rr < FED_P; ++rr) {
                    pre[rr][tid] = run;
                    run += cnt[rr][tid];
                }
            }
            __syncthreads();
        }
        for (int j = 0; j < tile_len; ++j) {
            double z1, cz = 0.0;
            if constexpr (TWO)
                mix_fed(KIND, A.sign * t1[tid][j], A.sign * t2[tid][j], c, z1, cz);
            else
                z1 = A.sign * t1[tid][j];
            step_diffusion<KIND>(s, z1, cz, c, tv, t0 + j);
            if constexpr (JUMPS) {
                int n = cnt[tid][j];
                if (n > 0) {
                    const int64_t j0 = A.base[(size_t)(t0 + j) * A.n_blocks + blockIdx.x] + pre[tid][j];
                    // counts that ask for more sizes than the caller's buffer holds read none
                    // beyond it (a NaN marks the path instead of an out-of-bounds access)
                    if (j0 + n > A.n_jumps) { s.x = __longlong_as_double(0x7ff8000000000000LL); n = 0; }
                    const double *jb = A.jumps + j0;
                    if (KIND != OPTMC_SABR_JUMP && A.sum_mode) {
                        double acc = 0.0;                       // np.sum(slice), path_kernels.py:116
                        for (int k = 0; k < n; ++k) acc += jb[k];
                        s.x += acc;
                    } else {
                        for (int k = 0; k < n; ++k) apply_jump<KIND>(s, jb[k]);
                    }
                }
            }
            if (A.paths) t1[tid][j] = step_spot(KIND, s, tv);          // path_kernels.py:34
        }
        if (A.paths) {
            __syncthreads();
            const int lane = tid & 31, warp = tid >> 5;
            for (int rr = warp; rr < FED_P; rr += FED_P / 32) {
                int64_t row = row0 + rr;
                if (row < A.n_paths && lane < tile_len)
                    A.paths[(size_t)row * pld + 1 + t0 + lane] = t1[rr][lane];
            }
        }
        __syncthreads();
    }
    if (A.terminal && live) A.terminal[p] = s.x;
}

// Streaming version of k_fed_block_totals: a thread walks its own row of counts (32-byte
// chunks) and adds the rare non-zero ones to tot[step][block of FED_P paths] atomically
// (integer adds: the result does not depend on their order).  tot must be zero on entry.
// Blocks are the 32-path groups one warp of k_fed_stream owns.
__global__ void __launch_bounds__(256, 4) k_fed_block_totals_stream(const int64_t *counts,
                                                                     int64_t n_paths, int n_steps,
                                                                     int64_t n_blocks, int64_t *tot) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_paths; p += stride) {
        const int64_t *row = counts + (size_t)p * n_steps;
        const int64_t blk = p / 32;                // jump blocks of k_fed_stream: one warp
        int i = 0;
        auto add = [&](int64_t cnt, int step) {
            if (cnt != 0)
                atomicAdd(reinterpret_cast<unsigned long long *>(tot + (size_t)step * n_blocks + blk),
                          (unsigned long long)cnt);
        };
        while (i < n_steps && ((uintptr_t)(row + i) & 31u)) {
            add(__ldg(row + i), i);
            ++i;
        }
        for (; i + 4 <= n_steps; i += 4) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(row + i));
            const longlong2 b = __ldg(reinterpret_cast<const longlong2 *>(row + i) + 1);
            add(a.x, i); add(a.y, i + 1); add(b.x, i + 2); add(b.y, i + 3);
        }
        for (; i < n_steps; ++i) add(__ldg(row + i), i);
    }
}

// ---- streaming variant ---------------------------------------------------------------
// k_fed stages 64 x 32 tiles through shared memory with 64-thread blocks: 8 warps per SM,
// so both HBM latency and the serial fp64 recurrence stay exposed (0.6-1.3 TB/s measured,
// tools/bench_fed.py).  Here a thread simply walks its own row: 32-byte chunks (two 16-byte loads per array: every
// sector fetched is fully used, the other half of the 64-byte DRAM atom is the same
// thread's next chunk and waits in L2), the next chunk in flight while the current one is
// stepped, 1024 threads per SM.  Rows whose start is not 32-byte aligned (n_steps not a
// multiple of 4) take up to three scalar steps first.  Same step functions, same order of
// operations per path as k_fed.
// Jump models: the jump buffer is ordered by step, then path, then jump.  The offset of
// (path p, step i) is base[i][p / 32] (totals of the 32-path group one warp owns + a scan,
// as for k_fed) plus the counts of the lower lanes at step i: a warp ballot per step, and
// a shuffle scan only when some path of the warp jumps (rare: lambda dt).
constexpr int FEDS_THREADS = 256;

__device__ __forceinline__ void ld_chunk(const double *p, double (&v)[4]) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ void ld_chunk(const int64_t *p, int64_t (&v)[4]) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(p));
    const longlong2 b = __ldg(reinterpret_cast<const longlong2 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// resident 256-thread blocks per SM: 4 (64 registers); 3 for the two-factor jump models and
// for the variants that also write the path matrix
__host__ __device__ constexpr int feds_min_blocks(int kind, bool paths) {
    return (paths || (kind_jumps(kind) && kind_two_factor(kind))) ? 3 : 4;
}

template <int KIND, bool PATHS>
__global__ void __launch_bounds__(FEDS_THREADS, feds_min_blocks(KIND, PATHS))
k_fed_stream(const FedArgs A) {
    constexpr bool TWO = kind_two_factor(KIND);
    constexpr bool JUMPS = kind_jumps(KIND);
    __shared__ MathShared math;                    // fast_log / fast_exp tables
    const RngView tv = fill_math_shared(&math);
    __syncthreads();
    const Consts &c = A.c;
    const int m = A.n_steps;
    const size_t pld = (size_t)m + 1;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // Jump models keep the 32 lanes of a warp on the same step (chunks by step index, no
    // per-row alignment prologue) because the jump offsets need a warp scan; rows that
    // are not 32-byte aligned are then read with scalar loads.
    const bool vec = (m % 4 == 0) && !(((uintptr_t)A.dw1 | (TWO ? (uintptr_t)A.dw2 : 0) |
                                         (JUMPS ? (uintptr_t)A.counts : 0)) & 31u);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p - lane < A.n_paths;
         p += stride) {
        const bool live = p < A.n_paths;
        const int64_t pr = live ? p : 0;           // dead lanes of the last warp replay path 0
        State s;
        s.x = c.x0;
        s.v = c.v0;
        const double *r1 = A.dw1 + (size_t)pr * m;
        const double *r2 = TWO ? A.dw2 + (size_t)pr * m : nullptr;
        const int64_t *rc = JUMPS ? A.counts + (size_t)pr * m : nullptr;
        double *out = (PATHS && live) ? A.paths + (size_t)p * pld : nullptr;
        if (out) out[0] = kind_sabr(KIND) ? c.x0 : exp(c.x0);             // path_kernels.py:28
        const int64_t wblk = (p - lane) / 32;      // jump block of this warp: 32 consecutive paths
        // one step; with a path matrix requested returns the spot after the step and, when
        // `put`, stores it (8 bytes) -- the chunked loop below collects four and stores
        // aligned 32-byte quads instead
        auto one = [&](double d1, double d2, int64_t cnt64, int i, bool put) -> double {
            double z1, cz = 0.0;
            if constexpr (TWO)
                mix_fed(KIND, A.sign * d1, A.sign * d2, c, z1, cz);
            else
                z1 = A.sign * d1;
            step_diffusion<KIND>(s, z1, cz, c, tv, i);
            if constexpr (JUMPS) {
                const int cnt = live ? (int)cnt64 : 0;
                if (__ballot_sync(0xffffffffu, cnt > 0)) {        // rare: some path of the warp jumps
                    int incl = cnt;                               // inclusive scan over the lanes
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    if (cnt > 0) {
                        const int64_t j0 = A.base[(size_t)i * A.n_blocks + wblk] + (incl - cnt);
                        int nj = cnt;
                        // (see k_fed: no read beyond the caller's n_jumps; the path becomes NaN)
                        if (j0 + nj > A.n_jumps) { s.x = __longlong_as_double(0x7ff8000000000000LL); nj = 0; }
                        const double *jb = A.jumps + j0;
                        if (KIND != OPTMC_SABR_JUMP && A.sum_mode) {
                            double acc = 0.0;                     // np.sum(slice), path_kernels.py:116
                            for (int k = 0; k < nj; ++k) acc += jb[k];
                            s.x += acc;
                        } else {
                            for (int k = 0; k < nj; ++k) apply_jump<KIND>(s, jb[k]);
                        }
                    }
                }
            }
            double spot = 0.0;
            if (out) {
                spot = step_spot(KIND, s, tv);                             // path_kernels.py:34
                if (put) out[i + 1] = spot;
            }
            return spot;
        };
        int i = 0;
        if constexpr (!JUMPS) {
            // scalar steps up to a 32-byte boundary of the rows (none when n_steps % 4 == 0)
            while (i < m && (((uintptr_t)(r1 + i) | (TWO ? (uintptr_t)(r2 + i) : 0)) & 31u)) {
                one(__ldg(r1 + i), TWO ? __ldg(r2 + i) : 0.0, 0, i, true);
                ++i;
            }
        }
        if (!JUMPS || vec) {
            double a[4], b[4] = {0.0, 0.0, 0.0, 0.0}, an[4], bn[4] = {0.0, 0.0, 0.0, 0.0};
            int64_t n4[4] = {0, 0, 0, 0}, nn[4] = {0, 0, 0, 0};
            if (i + 4 <= m) {
                ld_chunk(r1 + i, a);
                if constexpr (TWO) ld_chunk(r2 + i, b);
                if constexpr (JUMPS) ld_chunk(rc + i, n4);
            }
            // Path matrix: rows have stride n_steps + 1, so a row's 32-byte quads sit at a
            // per-row phase r = (position of out[i + 1] in its quad).  Every chunk of four new
            // spots completes exactly one aligned quad -- the last r spots of the previous
            // chunk followed by the first 4 - r new ones -- written as two 16-byte stores
            // (a full sector: no partial-sector write-backs); the first chunk's head and the
            // last r spots go out as 8-byte stores.
            const int r = out ? (int)(((uintptr_t)(out + i + 1) >> 3) & 3u) : 0;
            double p0 = 0.0, p1 = 0.0, p2 = 0.0;      // the previous chunk's spots 1, 2, 3
            bool head = true;
            for (; i + 4 <= m; i += 4) {
                const bool more = i + 8 <= m;
                if (more) {                         // next chunk in flight during these 4 steps
                    ld_chunk(r1 + i + 4, an);
                    if constexpr (TWO) ld_chunk(r2 + i + 4, bn);
                    if constexpr (JUMPS) ld_chunk(rc + i + 4, nn);
                }
                double o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = one(a[k], b[k], n4[k], i + k, false);
                if (out) {
                    if (head && r) {
                        for (int t = 0; t < 4 - r; ++t) out[i + 1 + t] = o[t];
                    } else {
                        const double e0 = r == 0 ? o[0] : r == 1 ? p2 : r == 2 ? p1 : p0;
                        const double e1 = r == 0 ? o[1] : r == 1 ? o[0] : r == 2 ? p2 : p1;
                        const double e2 = r == 0 ? o[2] : r == 1 ? o[1] : r == 2 ? o[0] : p2;
                        const double e3 = r == 0 ? o[3] : r == 1 ? o[2] : r == 2 ? o[1] : o[0];
                        double2 *q = reinterpret_cast<double2 *>(out + i + 1 - r);
                        q[0] = make_double2(e0, e1);
                        q[1] = make_double2(e2, e3);
                    }
                    head = false;
                    p0 = o[1]; p1 = o[2]; p2 = o[3];
                }
                if (more) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        a[k] = an[k];
                        b[k] = bn[k];
                        n4[k] = nn[k];
                    }
                }
            }
            if (out && !head) {                     // the last r spots of the last chunk
                if (r >= 3) out[i - 2] = p0;
                if (r >= 2) out[i - 1] = p1;
                if (r >= 1) out[i] = p2;
            }
        }
        for (; i < m; ++i)
            one(__ldg(r1 + i), TWO ? __ldg(r2 + i) : 0.0, JUMPS ? __ldg(rc + i) : 0, i, true);
        if (A.terminal && live) A.terminal[p] = s.x;
    }
}

}  // namespace optmc
