// Batched QPHB solver: one persistent CTA per spectrum in flight, FP64 tensor cores (DMMA) for the dense algebra.
//
// Replaces, for a whole batch of spectra at once, the loop of DRT._qphb_fit_core
// (reference hybdrt/models/drt1d.py:556-1008): qphb.initialize_weights (qphb.py:1609),
// qphb.iterate_qphb (:606) = L2 assembly (:53) + weighted Gram + bound-constrained QP (:426, cvxopt
// coneqp) + closed-form s / rho updates (:320, :385) + error-structure weights (:1545), the
// convergence test (:597), xmx normalisation (drt1d.py:946) and the hybrid vz_offset column rewrite
// (drt1d.py:972).
//
// Data layout.  Every n x n symmetric / triangular matrix of the QP is cut into 8 x 8 tiles; tile (j, i), j >= i,
// of the (negated) QP matrix P belongs to warp (j % W, i % W) of a W x W warp grid and lives in that warp's registers
// ("slot" (a, b) = (j / W, i / W), b <= a) in the accumulator layout of mma.sync.m8n8k4.f64: lane (g, q) =
// (lane / 4, lane % 4) holds [g][2q] and [g][2q + 1].  In that layout a tile is at the same time a valid A operand
// and a valid B^T operand of the instruction (the k index is summed over, so the k permutation {2q} / {2q + 1} of the
// two issues is immaterial):  D += X Z^T costs two DMMAs and no data movement.  Everything is phrased in that form:
//   Gram      P_ji = sum over 8-row chunks of rm^T_j diag(w^2) rm^T_i^T   (rm chunks staged transposed by cp.async)
//   Cholesky  H = P + diag(z / s) = L L^T per interior-point iteration, left-looking over tile columns, with the
//             inverse U = -L^-T built row by row behind the critical path; factor and inverse go to shared memory in
//             the same lane order.  See factor_chol.
//   Solves    u = U (U^T b): two passes over the tiles with DFMA and shuffle reductions.  See solve_kkt.
//   P         the (negated) QP matrix itself lives in tensor memory (TMEM) between the Gram pass and the end of the QP.
// Shared memory per CTA: a dozen length-NV vectors, reduction scratch, the tiles of L / L^-T (lower triangle, 512 B per
// tile in lane order), per-warp matvec partials, w[N], r2[N].  The Gram staging ring lives in the tile area as well.
// The design matrix rm, the variance-estimation matrix vmm and the penalty matrices are shared by the batch and
// stay in global memory (L2 resident, read-only path).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace hdrt {

constexpr int kMaxCols = 160;
constexpr int kRedSlots = 8;
constexpr int kNumVec = 11;

// cvxopt coneqp defaults (cvxopt 1.3 coneprog.py; the reference only sets show_progress, qphb.py:25)
constexpr double kAbsTol = 1e-7;
constexpr double kRelTol = 1e-6;
constexpr double kFeasTol = 1e-7;
constexpr int kMaxIpm = 100;
constexpr double kStep = 0.99;

extern __shared__ __align__(16) double g_smem[];

// Development aid (-DHDRT_PROFILE): per-phase clock64 totals of warp 0 of block 0, read back with hdrt_debug_profile.
#ifdef HDRT_PROFILE
__device__ unsigned long long g_prof[32];
__shared__ unsigned long long s_prof[32];
#define PROF_DECL long long _pt = clock64()
#define PROF_ADD(slot)                                                        \
    do {                                                                       \
        if (threadIdx.x == 0) s_prof[slot] += (unsigned long long)(clock64() - _pt); \
        _pt = clock64();                                                       \
    } while (0)
#define PROF_COUNT(slot) do { if (threadIdx.x == 0) s_prof[slot] += 1; } while (0)
#else
#define PROF_DECL
#define PROF_ADD(slot)
#define PROF_COUNT(slot)
#endif

// ------------------------------------------------------------------------------------------------
// compile-time configuration: tile grid, register slots, shared-memory offsets (in doubles)
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// EXT_ = false compiles the optional paths out (outlier error structure, solve_rp / update_scale, PFRT continuation):
// the kernel is sensitive to its instruction footprint, and the default fit does not pay for what it does not use.
template <int TMAX_, int W_, int MINB_, bool EXT_ = false>
struct Cfg {
    static constexpr int TMAX = TMAX_, W = W_, MINB = MINB_;
    static constexpr bool EXT = EXT_;
    static constexpr int kWarps = W * W, kThreads = 32 * kWarps;
    static constexpr int NV = 8 * TMAX;               // padded vector length
    static constexpr int A = (TMAX + W - 1) / W;      // slot rows per warp; slot (a, b), b <= a
    static constexpr int NSLOT = A * (A + 1) / 2;
    static constexpr int NTILE = TMAX * (TMAX + 1) / 2;
    static constexpr int NPART = 2 * W;
    static constexpr int kChunk = 8;                  // rows of rm per Gram step (two DMMA k-steps)
    static constexpr int kStages = 4;                 // cp.async ring depth of the Gram (lives in the tile area)
    static constexpr int RU = 4;    // rows per warp and pass of the row-wise matrix-vector loops: RU x 4
                                                      // L2 loads in flight per lane
    // offsets
    static constexpr int oRed = kNumVec * NV;
    static constexpr int oRbuf = oRed + 2 * kRedSlots * kWarps;
    static constexpr int oUnion = oRbuf + 16;
    static constexpr int oPart = oUnion;                       // QP view: NPART x NV matvec partials
    static constexpr int oTiles = oPart + NPART * NV;          // NTILE x 64: L, then U = -L^-T; Gram staging ring
    static_assert(kStages * NV * kChunk <= NTILE * 64, "staging ring must fit in the tile area");
    static constexpr int oRows = oTiles + NTILE * 64;          // w[N], r2[N], aux[N]
    enum { XS = 0, BS, DSQ, QS, SV0, SV1, SV2, US0, US1, US2, XH };
    static __device__ __forceinline__ double* vec(int k) { return g_smem + k * NV; }
    static __device__ __forceinline__ double* red() { return g_smem + oRed; }
    static __device__ __forceinline__ double* rbuf() { return g_smem + oRbuf; }
    static __device__ __forceinline__ double* part(int p) { return g_smem + oPart + p * NV; }
    static __device__ __forceinline__ double* tiles() { return g_smem + oTiles; }
    // per-row vectors w[N], r2[N] (each padded to a multiple of 8): derived from g_smem so that the
    // compiler emits shared-memory loads, not generic ones
    static __device__ __forceinline__ double* roww() { return g_smem + oRows; }
    static __device__ __forceinline__ double* rowr2(int N) { return g_smem + oRows + ((N + 7) & ~7); }
    static __device__ __forceinline__ double* rowu(int N) { return g_smem + oRows + 2 * ((N + 7) & ~7); }   // outlier fits only
    __host__ __device__ static constexpr int sidx(int a, int b) { return a * (a + 1) / 2 + b; }
};

using CfgS = Cfg<13, 2, 3>;  // n <= 104:  4 warps, 28 register tiles per warp, three CTAs per SM
using CfgL = Cfg<20, 3, 1>;  // n <= 160: 3 x 3 warps, 28 register tiles per warp, one CTA per SM (4 x 4 warps: 128 registers, spills, -12 %)
using CfgSX = Cfg<13, 2, 3, true>;   // the same with the optional paths compiled in
using CfgLX = Cfg<20, 3, 1, true>;
// Long problems (hybrid fits: thousands of chrono rows) leave room for two CTAs per SM at most: the same kernel with the
// register budget of two (255 instead of 168: the Gram accumulators and fragments stay out of local memory)
using CfgS2 = Cfg<13, 2, 2>;
using CfgS2X = Cfg<13, 2, 2, true>;

__host__ __device__ inline int rows_pad(int N) { return (N + 7) & ~7; }
template <class C>
__host__ __device__ inline long long smem_doubles_cfg(int N, int row_vectors) {
    return (long long)C::oRows + (long long)row_vectors * rows_pad(N) + 2;
}
__host__ __device__ inline bool small_cfg(int n) { return n <= CfgS::NV; }
// row_vectors: 2 (w, r^2) or 3 (+ T^1/2 r^2 of the outlier error structure)
__host__ __device__ inline long long smem_doubles(int N, int n, int row_vectors = 2) {
    return small_cfg(n) ? smem_doubles_cfg<CfgS>(N, row_vectors) : smem_doubles_cfg<CfgL>(N, row_vectors);
}

struct Ctx {
    int N, n, T, ns, nc, dop_a, dop_b, vz, vb_a, vb_b;
    int wr, wc, role, lane, g, q;  // warp grid position (role = W wr + wc), lane, accumulator-layout coordinates
    bool dv;                 // the diagonal slots (a, a) of this warp are inside the lower triangle (wc <= wr)
    const double* __restrict__ rm;
    const double* __restrict__ rv;
    const double* __restrict__ vmm_eis;
    const double* __restrict__ vmm_chrono;
    const double* __restrict__ pen;
    const double* __restrict__ hvec;
    const double* __restrict__ l1;
    const double* __restrict__ vz_strength;
    double* vzcol;  // global, per spectrum
    double* vz0;    // global, per spectrum: vz_offset column frozen at the start of a continuation step, else NULL
    double* t_out;  // global, per spectrum: outlier_t (only with outlier_p)
    unsigned tm;    // TMEM address (lane quadrant | first column) of this warp's tiles of -P
    double outlier_p;  // < 0: no outlier error structure
    double rv_scale;   // running data scale (solve_rp / update_scale): the data vector is rv * rv_scale
    double dop_cs;     // column scale of the DOP block of rm (solve_rp's DOP rescale), 1 otherwise
    int red_phase;
};

// D += X Z^T for 8 x 8 tiles in accumulator layout (see the header comment)
__device__ __forceinline__ void tile_mma(double2& d, const double2& x, const double2& z) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d.x), "+d"(d.y) : "d"(x.x), "d"(z.x));
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d.x), "+d"(d.y) : "d"(x.y), "d"(z.y));
}

// the two halves of tile_mma, for loops that issue the first halves of several independent tiles before the second ones
__device__ __forceinline__ void mma_lo(double2& d, const double2& x, const double2& z) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d.x), "+d"(d.y) : "d"(x.x), "d"(z.x));
}
__device__ __forceinline__ void mma_hi(double2& d, const double2& x, const double2& z) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d.x), "+d"(d.y) : "d"(x.y), "d"(z.y));
}

// 1 / sqrt(x) for a positive finite x: hardware approximation (2^-22) + two Newton steps
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
}

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, const double2 v) { *reinterpret_cast<double2*>(p) = v; }
// The same with 32-bit shared-window byte addresses: address arithmetic stays in one register and the loads are
// issued where they are written (the factorisation and the sweeps order them by hand)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds2a(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds1a(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
// a tile read transposed: lane (g, q) gets [2q][g] and [2q + 1][g]; `a` is the tile's address + (16 q + g) * 8
__device__ __forceinline__ double2 lds2t(unsigned a) { return make_double2(lds1a(a), lds1a(a + 64)); }
__device__ __forceinline__ void sts2a(unsigned a, const double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int bytes = valid ? 8 : 0;  // zero-fill when !valid
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double reduce_q(double v) {
    v += shfl_xor_d(v, 1);
    v += shfl_xor_d(v, 2);
    return v;
}
__device__ __forceinline__ double reduce_g(double v) {
    v += shfl_xor_d(v, 4);
    v += shfl_xor_d(v, 8);
    v += shfl_xor_d(v, 16);
    return v;
}

// Reduce K per-thread values over the block; bit k of MAXMASK selects max instead of sum.  The result is
// broadcast to every thread.  One barrier: the scratch buffer alternates between two halves.
template <class C, int K, unsigned MAXMASK>
__device__ __forceinline__ void block_reduce(double (&v)[K], Ctx& c) {
    static_assert(K <= kRedSlots, "too many reduction slots");
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = ((MAXMASK >> k) & 1u) ? warp_max(v[k]) : warp_sum(v[k]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* red = C::red() + (c.red_phase & 1) * (kRedSlots * C::kWarps);
    c.red_phase ^= 1;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[k * C::kWarps + w] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = red[k * C::kWarps];
#pragma unroll
        for (int ww = 1; ww < C::kWarps; ++ww) {
            const double t = red[k * C::kWarps + ww];
            a = ((MAXMASK >> k) & 1u) ? fmax(a, t) : a + t;
        }
        v[k] = a;
    }
}

// ------------------------------------------------------------------------------------------------
// The QP matrix P (negated, lower-triangle tiles) lives in tensor memory (TMEM) for the duration of a QP.  It is loop
// invariant (only the diagonal of the KKT matrix changes between interior-point iterations) and it is read once per
// iteration, a tile column at a time, by the warp that owns the tiles (tile (j, i) belongs to warp (j % W, i % W),
// as in the Gram pass) -- which is exactly what TMEM allows: a warp reaches the 32 TMEM lanes of its own quadrant
// only, and with the 32x32b shape lane l of the warp reads / writes N consecutive 32-bit columns of TMEM lane l.
// One 8 x 8 FP64 tile in accumulator layout (two doubles per lane) is four columns; slot (a, b) of a warp sits at
// column 4 sidx(a, b) of the warp's column range.  tcgen05.mma has no FP64 kind, so nothing else uses the 256 KB;
// holding P there frees the registers (the tiles used to be parked in them) and the shared memory (which holds the
// factor), and a tile comes back in a few tens of cycles where an L2 round trip takes several hundred.
// ------------------------------------------------------------------------------------------------
template <class C>
struct Tm {
    static constexpr int kGroups = (C::kWarps + 3) / 4;        // warps per TMEM lane quadrant
    static constexpr int kColsWarp = 4 * C::NSLOT;             // 32-bit columns per warp
    static constexpr int kColsNeed = kGroups * kColsWarp;
    static constexpr int kCols = kColsNeed <= 32 ? 32 : kColsNeed <= 64 ? 64 : kColsNeed <= 128 ? 128 : kColsNeed <= 256 ? 256 : 512;
    static_assert(kColsNeed <= 512, "QP matrix does not fit in tensor memory");
};

__device__ __forceinline__ void tmem_st2(unsigned taddr, const double2 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(__double2loint(v.x)), "r"(__double2hiint(v.x)), "r"(__double2loint(v.y)), "r"(__double2hiint(v.y))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// N tiles in one go: the loads are asynchronous, the wait sits in the same asm block so that nothing reads the
// destination registers early
template <int N>
__device__ __forceinline__ void tmem_ld_tiles(const unsigned (&ta)[N], double2 (&out)[N]) {
    static_assert(N >= 1 && N <= 4, "one to four tiles per block");
    unsigned r[4 * N];
    if constexpr (N == 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta[0]) : "memory");
    } else if constexpr (N == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%8];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [%9];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(ta[0]), "r"(ta[1]) : "memory");
    } else if constexpr (N == 3) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%12];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [%13];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [%14];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11])
                     : "r"(ta[0]), "r"(ta[1]), "r"(ta[2]) : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%16];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [%17];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [%18];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%12, %13, %14, %15}, [%19];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(ta[0]), "r"(ta[1]), "r"(ta[2]), "r"(ta[3]) : "memory");
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
        out[i] = make_double2(__hiloint2double((int)r[4 * i + 1], (int)r[4 * i]), __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]));
}

// C::A tiles (one tile row / column of a warp) in blocks of at most four
template <class C>
__device__ __forceinline__ void tmem_ld_row(const unsigned (&ta)[C::A], double2 (&dst)[C::A]) {
    constexpr int A = C::A, N0 = A < 4 ? A : 4;
    {
        unsigned t0[N0];
        double2 o0[N0];
#pragma unroll
        for (int i = 0; i < N0; ++i) t0[i] = ta[i];
        tmem_ld_tiles<N0>(t0, o0);
#pragma unroll
        for (int i = 0; i < N0; ++i) dst[i] = o0[i];
    }
    if constexpr (A > 4) {
        constexpr int N1 = A - 4;
        static_assert(N1 <= 4, "at most eight tile rows per warp");
        unsigned t1[N1];
        double2 o1[N1];
#pragma unroll
        for (int i = 0; i < N1; ++i) t1[i] = ta[4 + i];
        tmem_ld_tiles<N1>(t1, o1);
#pragma unroll
        for (int i = 0; i < N1; ++i) dst[4 + i] = o1[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Gram: P = rm^T diag(w^2) rm + L2,  q = -rm^T (w^2 rv) + l1.  The negated lower tiles of P go to tensor
// memory.  L2 = sum_k S_k^1/2 M~_k S_k^1/2 as in qphb.calculate_qp_l2_matrix (qphb.py:53-120).
// ------------------------------------------------------------------------------------------------
struct L2Factors {
    double drt[3];  // l2_lambda_0 * dw_k * rho_k
    double dop[3];  // dop_l2_lambda_0 * dop_dw_k * dop_rho_k
    bool use[3];    // derivative_weights[k] > 0
};

// One chunk = 8 rows of rm, staged transposed: stage[col][rr] = rm[r0 + rr][col] (zero beyond N / n).  One warp
// instruction copies a 4-column x 8-row block (lane = 4 rr + cl): 32-byte global segments, conflict-free stores.
// The vz_offset column of a hybrid fit is read from its per-spectrum buffer.
template <class C>
__device__ __forceinline__ void stage_chunk(const Ctx& c, int r0, int buf) {
    const int warp = threadIdx.x >> 5, cl = c.lane & 3, rr = c.lane >> 2;
    const bool rok = r0 + rr < c.N;
    const double* srow = c.rm + (size_t)(r0 + rr) * c.n;
    double* drow = C::tiles() + buf * C::NV * C::kChunk + rr;
#pragma unroll
    for (int u = 0; u < (C::NV / 4 + C::kWarps - 1) / C::kWarps; ++u) {
        const int col = 4 * (warp + C::kWarps * u) + cl;
        if (col < C::NV) {
            const bool ok = rok && col < c.n;
            const double* src = (col == c.vz) ? (c.vzcol + r0 + rr) : (srow + col);
            cp_async8(drow + col * C::kChunk, ok ? src : c.rm, ok);
        }
    }
    cp_async_commit();
}

template <class C>
__device__ __noinline__ void gram_phase(Ctx& cref, const L2Factors& fref, bool l1_scalar, double l1_value, double* p_out,
                                        double* q_out) {
    Ctx c = cref;            // by value: the fields stay in registers instead of the caller's stack frame
    const L2Factors f = fref;
    const int tid = threadIdx.x;
    const int n = c.n, N = c.N, T = c.T;
    const int g = c.g, q = c.q;
    PROF_DECL;
    double* w2 = C::rowr2(c.N);      // w^2, zero padded to a multiple of 8
    __syncthreads();        // previous users of the tile area and of r2 are done
    // rm chunks travel through a ring of kStages buffers inside the (still unused) tile area: the copies are L2
    // hits with a latency of a few chunks' worth of DMMA work
    constexpr int NS = C::kStages;
    const int nchunks = (N + C::kChunk - 1) / C::kChunk;
#pragma unroll
    for (int st = 0; st < NS - 1; ++st) {
        if (st < nchunks) stage_chunk<C>(c, st * C::kChunk, st); else cp_async_commit();
    }
    for (int r = tid; r < rows_pad(N); r += C::kThreads) {
        w2[r] = (r < N) ? C::roww()[r] * C::roww()[r] : 0.0;
    }
    constexpr int QU = (4 * C::NV + C::kThreads - 1) / C::kThreads;  // q: lane group of 4 per column
    double qacc[QU];
#pragma unroll
    for (int u = 0; u < QU; ++u) qacc[u] = 0.0;
    double2 S[C::NSLOT];
#pragma unroll
    for (int e = 0; e < C::NSLOT; ++e) S[e] = make_double2(0.0, 0.0);
    for (int ci = 0; ci < nchunks; ++ci) {
        const int r0 = ci * C::kChunk, buf = ci % NS;
        cp_async_wait<NS - 2>();   // chunk ci has landed (this thread's copies); the barrier publishes it
        __syncthreads();
        if (ci + NS - 1 < nchunks) stage_chunk<C>(c, (ci + NS - 1) * C::kChunk, (ci + NS - 1) % NS);   // refill the
        else cp_async_commit();                                            // buffer everyone left last iteration
        const double* base = C::tiles() + buf * C::NV * C::kChunk;
        // fragment of tile column X: .x = rm[r0 + 2q][8X + g], .y = rm[r0 + 2q + 1][8X + g]
        const double* frow = base + (8 * c.wr + g) * C::kChunk + 2 * q;
        const double* fcol = base + (8 * c.wc + g) * C::kChunk + 2 * q;
        const double2 wq = lds2(w2 + r0 + 2 * q);
#pragma unroll
        for (int a = 0; a < C::A; ++a) {
            if (C::W * a + c.wr < T) {
                double2 Fa = lds2(frow + a * (C::W * 64));
                Fa.x *= wq.x;
                Fa.y *= wq.y;
#pragma unroll
                for (int b = 0; b < a; ++b) tile_mma(S[C::sidx(a, b)], Fa, lds2(fcol + b * (C::W * 64)));
                if (c.dv) tile_mma(S[C::sidx(a, a)], Fa, lds2(fcol + a * (C::W * 64)));
            }
        }
        {
            const int rq = r0 + 2 * (tid & 3);       // q = -rm^T (w^2 rv): rv straight from global (issued early)
            const double2 w2q = lds2(w2 + rq);
            const double rsc = C::EXT ? c.rv_scale : 1.0;
            const double2 wv = make_double2(rq < N ? w2q.x * (c.rv[rq] * rsc) : 0.0,
                                            rq + 1 < N ? w2q.y * (c.rv[rq + 1] * rsc) : 0.0);
#pragma unroll
            for (int u = 0; u < QU; ++u) {
                const int e = tid + C::kThreads * u;   // column e / 4, rows 2 (e % 4), + 1 of the chunk
                if (e < 4 * C::NV) {
                    const double2 v = lds2(base + 2 * e);
                    qacc[u] = fma(v.x, wv.x, qacc[u]);
                    qacc[u] = fma(v.y, wv.y, qacc[u]);
                }
            }
        }
    }
    PROF_ADD(5);
    // -Gram to tensor memory (that frees the accumulators), then one pass per tile row of this warp adds the penalty,
    // sets the padding rows / columns to the identity and writes the optional dense copy: the tiles of the row come
    // back from TMEM while all penalty entries of the row (L2 hits, several hundred cycles each) are in flight
#pragma unroll
    for (int a = 0; a < C::A; ++a) {
        if (C::W * a + c.wr < T) {
#pragma unroll
            for (int b = 0; b <= a; ++b)
                if (b < a || c.dv) tmem_st2(c.tm + 4 * C::sidx(a, b), make_double2(-S[C::sidx(a, b)].x, -S[C::sidx(a, b)].y));
        }
    }
    tmem_wait_st();
    {
        const int nn = n * n;
        const bool dopb = c.dop_a >= 0;
        constexpr int NB4 = 4;      // slots per batch: 24 penalty loads in flight per lane
#pragma unroll 1
        for (int ab = 0; ab < C::A * ((C::A + NB4 - 1) / NB4); ++ab) {
            const int a = ab / ((C::A + NB4 - 1) / NB4), b0 = NB4 * (ab % ((C::A + NB4 - 1) / NB4));
            const int j = C::W * a + c.wr;
            const int nb = a + (c.dv ? 1 : 0);           // slots (a, 0 .. nb - 1)
            if (j >= T) break;
            if (b0 >= nb) continue;
            const int r = 8 * j + g, rl = min(r, n - 1);
            const unsigned trow = c.tm + 4 * (a * (a + 1) / 2);
            double pm[NB4][3][2];
#pragma unroll
            for (int bb = 0; bb < NB4; ++bb) {
                const int cc0 = 8 * (C::W * min(b0 + bb, nb - 1) + c.wc) + 2 * q;
                const int c0 = min(cc0, n - 1), c1 = min(cc0 + 1, n - 1);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    pm[bb][k][0] = f.use[k] ? c.pen[k * nn + rl * n + c0] : 0.0;
                    pm[bb][k][1] = f.use[k] ? c.pen[k * nn + rl * n + c1] : 0.0;
                }
            }
            double2 t[NB4];
            {
                unsigned ta[NB4];
#pragma unroll
                for (int bb = 0; bb < NB4; ++bb) ta[bb] = trow + 4 * min(b0 + bb, nb - 1);
                tmem_ld_tiles<NB4>(ta, t);
            }
            double usr[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) usr[k] = C::vec(C::US0 + k)[rl];
#pragma unroll
            for (int bb = 0; bb < NB4; ++bb) {
                const int b = b0 + bb;
                if (b < nb) {
                    const int cc0 = 8 * (C::W * b + c.wc) + 2 * q;
                    double o[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = cc0 + e;
                        double gram = e ? t[bb].y : t[bb].x;     // negated Gram entry
                        double v;
                        if (r < n && cc < n) {
                            // qphb.calculate_qp_l2_matrix, qphb.py:53-120
                            const bool drt = (r >= c.ns) && (cc >= c.ns);
                            const bool dop = dopb && (r >= c.dop_a) && (r < c.dop_b) && (cc >= c.dop_a) && (cc < c.dop_b);
                            double acc = 0.0;
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                if (!f.use[k]) continue;
                                double m = pm[bb][k][e];
                                if (drt) m *= f.drt[k];
                                if (dop) m *= f.dop[k];
                                acc += (usr[k] * m) * C::vec(C::US0 + k)[cc];
                            }
                            if (C::EXT && dopb) {  // DOP columns of rm carry the per-spectrum rescale (drt1d.py:589-596)
                                if (r >= c.dop_a && r < c.dop_b) gram *= c.dop_cs;
                                if (cc >= c.dop_a && cc < c.dop_b) gram *= c.dop_cs;
                            }
                            v = acc - gram;
                            if (p_out && cc <= r) {
                                p_out[(size_t)r * n + cc] = v;
                                p_out[(size_t)cc * n + r] = v;
                            }
                        } else {
                            v = (r == cc) ? 1.0 : 0.0;
                        }
                        o[e] = -v;
                    }
                    tmem_st2(trow + 4 * b, make_double2(o[0], o[1]));
                }
            }
        }
        tmem_wait_st();
    }
    PROF_ADD(6);
#pragma unroll
    for (int u = 0; u < QU; ++u) {
        const int col = (tid + C::kThreads * u) >> 2;
        const double s = reduce_q(qacc[u]);
        if ((tid & 3) == 0 && col < n) {
            const double sc = (C::EXT && col >= c.dop_a && col < c.dop_b) ? s * c.dop_cs : s;
            const double qv = -sc + (l1_scalar ? l1_value : c.l1[col]);
            C::vec(C::QS)[col] = qv;
            if (q_out) q_out[col] = qv;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Diagonal tile: s = -C_kk in accumulator layout (one warp).  Gaussian elimination on M = [C_kk | I] (8 x 16)
// without row scaling: row g ends as [d_g Lt_g | Lt^-1_g] of C_kk = Lt D Lt^T (Lt unit lower); scaling row g by
// d_g^-1/2 afterwards gives [L^T | L^-1] of the Cholesky factor.  Lane (g, q) keeps M[g][2q], M[g][2q + 1],
// M[g][8 + 2q], M[g][9 + 2q] in registers -- the accumulator layout of both halves; row cc, its pivot and the
// column-cc element of the own row travel by shuffles.  This runs on one warp while its block waits, and a lone
// warp issues an instruction only every few cycles: the loop is rolled (it must stay in the instruction cache)
// and carries as few instructions as possible (one reciprocal, one multiply, four FMAs, six shuffles).
// Publishes -L^-1 (row-major = accumulator layout) to `binv` (this lane's shared-window address of the tile).  false
// on breakdown (non-positive or non-finite pivot), uniformly over the warp.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x) {   // 1 / x, x positive and finite: 2^-22 seed + 2 Newton
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

__device__ __forceinline__ bool diag_factor(const double2 s, unsigned binv, int lane) {
    const int g = lane >> 2, q = lane & 3;
    double m0 = -s.x, m1 = -s.y;
    double m2 = (g == 2 * q) ? 1.0 : 0.0, m3 = (g == 2 * q + 1) ? 1.0 : 0.0;
#pragma unroll 1
    for (int cc = 0; cc < 7; ++cc) {
        const int h = cc >> 1;
        const double colv = (cc & 1) ? m1 : m0;                      // column cc of the own row, where q == h
        const double agc = shfl_idx_d(colv, 4 * g + h);      // M[g][cc]
        const double piv = shfl_idx_d(colv, 4 * cc + h);     // M[cc][cc]
        const int src = 4 * cc + q;                                  // row cc
        const double r0 = shfl_idx_d(m0, src), r1 = shfl_idx_d(m1, src);
        const double r2 = shfl_idx_d(m2, src), r3 = shfl_idx_d(m3, src);
        // f = -agc / piv with the reciprocal folded in: y0 = rcp seed, e = 1 - piv y0, 1 / piv = y0 (1 + e + e^2 + O(e^3))
        double y0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(piv));
        const double e = fma(-piv, y0, 1.0);
        const double t = fma(e, e, e);
        const double f0 = -agc * y0;
        const double f = (g > cc) ? fma(f0, t, f0) : 0.0;
        m0 = fma(f, r0, m0);
        m1 = fma(f, r1, m1);
        m2 = fma(f, r2, m2);
        m3 = fma(f, r3, m3);
    }
    const double dg = shfl_idx_d((g & 1) ? m1 : m0, 4 * g + (g >> 1));   // d_g = M[g][g]
    const bool ok = __all_sync(kFull, (dg > 0.0) && (dg < INFINITY));
    const double rinv = fast_rsqrt(dg);
    m2 *= rinv;
    m3 *= rinv;
    sts2a(binv, make_double2(-m2, -m3));
    return ok;
}

// ------------------------------------------------------------------------------------------------
// H = P + diag(dsq) = L L^T and U = -L^-T, left-looking over tile columns.  -P is read from tensor memory; the
// results go to the shared-memory tile area:  while the factorisation runs, tile (j, k) holds L_jk (j > k) and the
// diagonal slot (k, k) holds -L_kk^-1;  at the end tile (j, k) holds the block U_kj (block row k, block column j); the
// diagonal blocks U_kk = (-L_kk^-1)^T are the diagonal slots read transposed.
//
// The warps with the same wc form a team that owns the tile columns k = wc (mod W), one at a time: its
// accumulators start from -P_jk (from tensor memory; dsq joins on the diagonal tile) and collect  sum_m L_jm L_km^T  over
// the columns m that are already published.  Step s finishes column s: the team adds the last term (m = s - 1), the warp
// holding the diagonal tile factorises and inverts it (diag_factor), a team barrier publishes -L_ss^-1, every team
// warp scales its tiles (L_js = C_js L_ss^-T, two DMMAs) and stores them; one block barrier per step publishes the
// column.  Only one term per column sits on the critical path, which is
// T x (one tile update + diag_factor + one tile scaling + two barriers).
// The other teams spend the step catching up on the columns published so far, and every warp but the one busy with the
// diagonal tile takes a share of turning row s - 1 of L, which no later column needs, into row s - 1 of the inverse
// in place:
//     U_i,s-1 = (U_ii L_s-1,i^T + sum_{i < m < s-1} U_i,m L_s-1,m^T) (-L_s-1,s-1^-1)^T,
// the same accumulate-and-scale pattern as a column of L.  A predicated-off DMMA still occupies the FP64 pipe for its
// 16 cycles, so the unrolled tile loops are instantiated per active-tile count and entered through a branch.
// Returns false on breakdown (uniform across the block).
// ------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void team_barrier(int wc) {   // named barrier 1 + wc over the W warps of a team
    static_assert(W <= 4, "one named barrier per team");
    if (wc == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * W) : "memory");
    else if (wc == 1) asm volatile("bar.sync 2, %0;" ::"n"(32 * W) : "memory");
    else if (wc == 2) asm volatile("bar.sync 3, %0;" ::"n"(32 * W) : "memory");
    else asm volatile("bar.sync 4, %0;" ::"n"(32 * W) : "memory");
}

// The tile rows of a warp are indexed from the bottom of the matrix upwards: r = 0 is its last tile row, so that the
// rows of a column that are still active are always r = 0 .. nv - 1.
template <class C>
struct RowMap {
    int a_hi;                     // tile rows of this warp: j = W a + wr < T  <=>  a < a_hi
    unsigned addr[C::A];          // this lane's byte address of tile (j_r, 0)
    __device__ __forceinline__ int row(const Ctx& c, int r) const { return C::W * (a_hi - 1 - r) + c.wr; }
};

// tiles (j_r, kc), r < nv, of -P for this warp, from TMEM (slot (a, bc) with a = a_hi - 1 - r; the rows r >= nv read
// a valid slot nobody uses)
template <class C>
__device__ __forceinline__ void load_column(const Ctx& c, const RowMap<C>& rm, double2 (&dst)[C::A], int bc) {
    unsigned ta[C::A];
#pragma unroll
    for (int r = 0; r < C::A; ++r) {
        const int a = max(rm.a_hi - 1 - r, bc);
        ta[r] = c.tm + 4 * (a * (a + 1) / 2 + bc);
    }
    tmem_ld_row<C>(ta, dst);
}

template <class C, int N>
__device__ __forceinline__ void catchup_body(double2 (&acc)[C::A], const unsigned (&rowaddr)[C::A], unsigned zaddr, int m0, int m1) {
    // the operands of term m + 1 are requested before the DMMAs of term m are issued
    double2 Z = lds2a(zaddr), X[N];
#pragma unroll
    for (int r = 0; r < N; ++r) X[r] = lds2a(rowaddr[r] + m0 * 512);
#pragma unroll 1
    for (int m = m0; m < m1; ++m) {
        const int mn = min(m + 1, m1 - 1);
        const double2 Zn = lds2a(zaddr + (mn - m0) * 512);
        double2 Xn[N];
#pragma unroll
        for (int r = 0; r < N; ++r) Xn[r] = lds2a(rowaddr[r] + mn * 512);
#pragma unroll
        for (int r = 0; r < N; ++r) mma_lo(acc[r], X[r], Z);
#pragma unroll
        for (int r = 0; r < N; ++r) mma_hi(acc[r], X[r], Z);
        Z = Zn;
#pragma unroll
        for (int r = 0; r < N; ++r) X[r] = Xn[r];
    }
}
template <class C, int N>
__device__ __forceinline__ void catchup(int nv, double2 (&acc)[C::A], const unsigned (&rowaddr)[C::A], unsigned zaddr, int m0, int m1) {
    if (nv == N) catchup_body<C, N>(acc, rowaddr, zaddr, m0, m1);
    else if constexpr (N > 1) catchup<C, N - 1>(nv, acc, rowaddr, zaddr, m0, m1);
}

// L_js = C_js L_ss^-T for the first ns accumulators (D = acc bn^T), stored to tile (j_r, s)
template <class C, int N>
__device__ __forceinline__ void scale_store(int ns, const double2 (&acc)[C::A], const unsigned (&rowaddr)[C::A], const double2 bn, int s) {
    if (ns == N) {
        double2 r2[N];
#pragma unroll
        for (int r = 0; r < N; ++r) { r2[r] = make_double2(0.0, 0.0); mma_lo(r2[r], acc[r], bn); }
#pragma unroll
        for (int r = 0; r < N; ++r) mma_hi(r2[r], acc[r], bn);
#pragma unroll
        for (int r = 0; r < N; ++r) sts2a(rowaddr[r] + s * 512, r2[r]);
    } else if constexpr (N > 1) {
        scale_store<C, N - 1>(ns, acc, rowaddr, bn, s);
    }
}

// The inverse row is shared by all warps but the one that is busy with the diagonal tile of the step ("workers",
// NWK = kWarps - 1 of them): worker wk takes the block columns i_e = wk + NWK e.
template <class C>
struct Inv {
    static constexpr int NWK = C::kWarps - 1;
    static constexpr int EI = (C::TMAX - 1 + NWK - 1) / NWK;      // block columns per worker
};

// one term of the inverse row: inv[e] += Xop_e Z^T for the n active columns i_e <= m of this worker
template <class C, int N>
__device__ __forceinline__ void invert_term(int n, double2 (&inv)[Inv<C>::EI], unsigned zaddr, unsigned mrow, unsigned lastaddr, bool diag) {
    if (n == N) {
        const double2 Z = lds2a(zaddr);
        double2 X[N];
#pragma unroll
        for (int e = 0; e < N - 1; ++e) X[e] = lds2a(mrow + e * (Inv<C>::NWK * 512));
        X[N - 1] = diag ? lds2t(lastaddr) : lds2a(lastaddr);     // U_mm = (-L_mm^-1)^T: the diagonal slot read transposed
#pragma unroll
        for (int e = 0; e < N; ++e) mma_lo(inv[e], X[e], Z);
#pragma unroll
        for (int e = 0; e < N; ++e) mma_hi(inv[e], X[e], Z);
    } else if constexpr (N > 1) {
        invert_term<C, N - 1>(n, inv, zaddr, mrow, lastaddr, diag);
    }
}
template <class C, int N>
__device__ __forceinline__ void invert_store(int ne, const double2 (&inv)[Inv<C>::EI], const double2 bn, unsigned dst) {
    if (ne == N) {
        double2 r2[N];
#pragma unroll
        for (int e = 0; e < N; ++e) { r2[e] = make_double2(0.0, 0.0); mma_lo(r2[e], inv[e], bn); }
#pragma unroll
        for (int e = 0; e < N; ++e) mma_hi(r2[e], inv[e], bn);
#pragma unroll
        for (int e = 0; e < N; ++e) sts2a(dst + e * (Inv<C>::NWK * 512), r2[e]);
    } else if constexpr (N > 1) {
        invert_store<C, N - 1>(ne, inv, bn, dst);
    }
}

// Row sr of L -> row sr of the inverse (blocks U_i,sr, i < sr, into the tiles (sr, i)), worker wk's share.
template <class C>
__device__ __forceinline__ void invert_row(int sr, int wk, unsigned tl, unsigned tt) {
    constexpr int NWK = Inv<C>::NWK, EI = Inv<C>::EI;
    const int ne = sr > wk ? (sr - wk + NWK - 1) / NWK : 0;      // block columns i_e = wk + NWK e < sr
    const unsigned srow = tl + (sr * (sr + 1) / 2) * 512;         // tiles (sr, m)
    double2 inv[EI];
#pragma unroll
    for (int e = 0; e < EI; ++e) inv[e] = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int m = wk; m < sr; ++m) {
        const int n = (m - wk) / NWK + 1;                         // active columns: i_e <= m
        const unsigned mrow = tl + (m * (m + 1) / 2 + wk) * 512;  // tiles (m, i_e) = U_(i_e, m), e = 0 ..
        const bool diag = (m - wk) % NWK == 0;                    // i_(n-1) == m: the block is U_mm
        const unsigned lastaddr = diag ? tt + (m * (m + 1) / 2 + m) * 512 : mrow + (n - 1) * (NWK * 512);
        invert_term<C, EI>(n, inv, srow + m * 512, mrow, lastaddr, diag);
    }
    asm volatile("bar.sync 5, %0;" ::"n"(32 * NWK) : "memory");   // every worker is done reading row sr of L
    if (ne > 0) invert_store<C, EI>(ne, inv, lds2a(srow + sr * 512), srow + wk * 512);
}

template <class C>
__device__ __noinline__ bool factor_chol(const Ctx& cref) {   // a function of its own: the caller's live values are
    const Ctx c = cref;                                       // parked on its stack once per call, not in these loops
    const int T = c.T, lane = c.lane;
    constexpr int W = C::W;
    PROF_DECL;
    const unsigned tl = smem_u32(C::tiles()) + 16 * lane;   // this lane's element pair of tile 0 (byte address)
    const unsigned tt = smem_u32(C::tiles()) + (16 * c.q + c.g) * 8;   // ... for a transposed tile read
    RowMap<C> rm;
    rm.a_hi = T > c.wr ? (T - c.wr + W - 1) / W : 0;
#pragma unroll
    for (int r = 0; r < C::A; ++r) {
        const int j = max(rm.row(c, r), 0);
        rm.addr[r] = tl + (j * (j + 1) / 2) * 512;
    }
    const bool owner = c.wr == c.wc;             // this warp holds the diagonal tiles of its team's columns
    double2 acc[C::A];
    int kc = c.wc, bc = 0, mdone = 0;            // the team's current column, its slot column, terms collected
    // active tile rows of column kc: slots a >= bc (+ 1 if the slot (bc, bc) lies above the diagonal), a < a_hi
    int nv = max(rm.a_hi - (bc + (c.dv ? 0 : 1)), 0);
    bool my_ok = true;
    if (kc < T) load_column<C>(c, rm, acc, bc);
    const double* dsq = C::vec(C::DSQ);
    PROF_ADD(16);
#pragma unroll 1
    for (int s = 0; s <= T; ++s) {
        const int ow = (s % W) * (W + 1);          // role of the warp that factorises the diagonal tile of this step
        const bool crit = s < T && kc == s;        // this warp's team finishes its column in this step
        if (crit && mdone < s) {                   // the critical path first: the last term(s) of column s
            catchup<C, C::A>(nv, acc, rm.addr, tl + (kc * (kc + 1) / 2 + mdone) * 512, mdone, s);
            mdone = s;
        }
        PROF_ADD(17);
        if (s >= 1 && c.role != ow) {
            invert_row<C>(s - 1, c.role - (c.role > ow ? 1 : 0), tl, tt);
            PROF_ADD(21);
        }
        if (!crit && s < T && kc < T && mdone < s) {
            catchup<C, C::A>(nv, acc, rm.addr, tl + (kc * (kc + 1) / 2 + mdone) * 512, mdone, s);
            mdone = s;
            PROF_ADD(17);
        }
        if (crit) {
            const unsigned dslot = tl + (s * (s + 1) / 2 + s) * 512;
            if (owner) {
                double2 sk = make_double2(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < C::A; ++r) if (r == nv - 1) sk = acc[r];
                const double d = dsq[8 * s + c.g];     // -(C_ss + diag(dsq))
                if (c.g == 2 * c.q) sk.x -= d;
                if (c.g == 2 * c.q + 1) sk.y -= d;
                my_ok = diag_factor(sk, dslot, lane);
                PROF_COUNT(24);
                PROF_ADD(18);
            }
            team_barrier<W>(c.wc);
            const int ns = nv - (owner ? 1 : 0);
            if (ns > 0) scale_store<C, C::A>(ns, acc, rm.addr, lds2a(dslot), s);
            kc += W;
            ++bc;
            mdone = 0;
            nv = max(nv - 1, 0);
            if (kc < T) load_column<C>(c, rm, acc, bc);
            PROF_ADD(19);
        }
        const int all_ok = __syncthreads_and(my_ok);   // column s and inverse row s - 1 published
        PROF_ADD(20);
        if (!all_ok) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Solves with U = -L^-T in the shared-memory tiles:  H^-1 b = U (U^T b).  Tile (j, i), j > i, holds the block U_ij
// (block row i, block column j), the diagonal slot (j, j) the transpose of U_jj; warp (wr, wc) of the grid takes the tiles
// (j, i) = (W a + wr, W b + wc) as for P.  Lane (g, q) of a block holds u[g][2q], u[g][2q + 1]:
//   (U^T b)_j [2q (+1)] += u[g][2q (+1)] b_i[g]          (reduce over g), partial per wc:  part(wc)
//   (U t)_i [g]         += u[g][2q] t_j[2q] + u[g][2q + 1] t_j[2q + 1]   (reduce over q), partial per wr:  part(W + wr)
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void load_u_tiles(const Ctx& c, double2 (&S)[C::NSLOT]) {
    const unsigned tl = smem_u32(C::tiles()) + 16 * c.lane, tt = smem_u32(C::tiles()) + (16 * c.q + c.g) * 8;
#pragma unroll
    for (int a = 0; a < C::A; ++a) {
        const int j = C::W * a + c.wr;
#pragma unroll
        for (int b = 0; b <= a; ++b) S[C::sidx(a, b)] = make_double2(0.0, 0.0);
        if (j < c.T) {
            const unsigned trow = tl + (j * (j + 1) / 2 + c.wc) * 512;
#pragma unroll
            for (int b = 0; b < a; ++b) S[C::sidx(a, b)] = lds2a(trow + b * (C::W * 512));
            if (c.dv) S[C::sidx(a, a)] = (c.wr == c.wc) ? lds2t(tt + (j * (j + 1) / 2 + j) * 512) : lds2a(trow + a * (C::W * 512));
        }
    }
}

// part(wc)[8j + ..] <- this warp's contribution to U^T bs; sum over wc to finish
template <class C>
__device__ __forceinline__ void matvec_ut(const Ctx& c, const double2 (&S)[C::NSLOT], const double* bs) {
    const int T = c.T, g = c.g, q = c.q;
    const double* bc = bs + 8 * c.wc + g;
    double bg[C::A];
#pragma unroll
    for (int b = 0; b < C::A; ++b) bg[b] = bc[b * (8 * C::W)];   // NV-padded: always in range
#pragma unroll
    for (int a = 0; a < C::A; ++a) {
        const int j = C::W * a + c.wr;
        if (j < T) {
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int b = 0; b <= a; ++b) {   // out-of-triangle diagonal slots hold zeros
                acc.x = fma(S[C::sidx(a, b)].x, bg[b], acc.x);
                acc.y = fma(S[C::sidx(a, b)].y, bg[b], acc.y);
            }
            const double v0 = reduce_g(acc.x), v1 = reduce_g(acc.y);
            if (g == 0) sts2(C::part(c.wc) + 8 * j + 2 * q, make_double2(v0, v1));
        }
    }
}

// part(W + wr)[8i + ..] <- this warp's contribution to U t, t = sum_{p < W} part(p); sum over wr to finish
template <class C>
__device__ __forceinline__ void matvec_u(const Ctx& c, const double2 (&S)[C::NSLOT]) {
    const int T = c.T, g = c.g, q = c.q;
    double acc[C::A];
#pragma unroll
    for (int b = 0; b < C::A; ++b) acc[b] = 0.0;
    const double* tr = C::part(0) + 8 * c.wr + 2 * q;
#pragma unroll
    for (int a = 0; a < C::A; ++a) {
        const int j = C::W * a + c.wr;
        if (j < T) {
            double2 tv = lds2(tr + a * (8 * C::W));
#pragma unroll
            for (int p = 1; p < C::W; ++p) {
                const double2 t = lds2(tr + p * C::NV + a * (8 * C::W));
                tv.x += t.x;
                tv.y += t.y;
            }
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                acc[b] = fma(S[C::sidx(a, b)].x, tv.x, acc[b]);
                acc[b] = fma(S[C::sidx(a, b)].y, tv.y, acc[b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < C::A; ++b) {
        const int i = C::W * b + c.wc;
        const double v = reduce_q(acc[b]);
        if (q == 0 && i < T) C::part(C::W + c.wr)[8 * i + g] = v;
    }
}

// H u = bs; on return u[t] = sum_{p < W} part(W + p)[t].  The tiles are streamed from shared memory a tile row at a
// time (the same row feeds both products of its pass), so that the solve needs few registers and stays inline in
// the QP loop: a call would park the loop's state in local memory, which at this occupancy means L2.
template <class C>
__device__ __forceinline__ void solve_kkt(const Ctx& c, const double* bs) {
    const int T = c.T, g = c.g, q = c.q;
    const unsigned tl = smem_u32(C::tiles()) + 16 * c.lane, tt = smem_u32(C::tiles()) + (16 * c.q + c.g) * 8;
    const bool dg = c.wr == c.wc;
    {   // part(wc) <- U^T bs (this warp's tiles)
        const double* bc = bs + 8 * c.wc + g;
        double bg[C::A];
#pragma unroll
        for (int b = 0; b < C::A; ++b) bg[b] = bc[b * (8 * C::W)];   // NV-padded: always in range
#pragma unroll
        for (int a = 0; a < C::A; ++a) {
            const int j = C::W * a + c.wr;
            if (j < T) {
                const unsigned trow = tl + (j * (j + 1) / 2 + c.wc) * 512;
                double2 t[C::A];
#pragma unroll
                for (int b = 0; b < a; ++b) t[b] = lds2a(trow + b * (C::W * 512));
                t[a] = c.dv ? (dg ? lds2t(tt + (j * (j + 1) / 2 + j) * 512) : lds2a(trow + a * (C::W * 512))) : make_double2(0.0, 0.0);
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int b = 0; b <= a; ++b) {
                    acc.x = fma(t[b].x, bg[b], acc.x);
                    acc.y = fma(t[b].y, bg[b], acc.y);
                }
                const double v0 = reduce_g(acc.x), v1 = reduce_g(acc.y);
                if (g == 0) sts2(C::part(c.wc) + 8 * j + 2 * q, make_double2(v0, v1));
            }
        }
    }
    __syncthreads();
    {   // part(W + wr) <- U t, t = sum_p part(p)
        double acc[C::A];
#pragma unroll
        for (int b = 0; b < C::A; ++b) acc[b] = 0.0;
        const double* tr = C::part(0) + 8 * c.wr + 2 * q;
#pragma unroll
        for (int a = 0; a < C::A; ++a) {
            const int j = C::W * a + c.wr;
            if (j < T) {
                const unsigned trow = tl + (j * (j + 1) / 2 + c.wc) * 512;
                double2 t[C::A];
#pragma unroll
                for (int b = 0; b < a; ++b) t[b] = lds2a(trow + b * (C::W * 512));
                t[a] = c.dv ? (dg ? lds2t(tt + (j * (j + 1) / 2 + j) * 512) : lds2a(trow + a * (C::W * 512))) : make_double2(0.0, 0.0);
                double2 tv = lds2(tr + a * (8 * C::W));
#pragma unroll
                for (int p = 1; p < C::W; ++p) {
                    const double2 tp = lds2(tr + p * C::NV + a * (8 * C::W));
                    tv.x += tp.x;
                    tv.y += tp.y;
                }
#pragma unroll
                for (int b = 0; b <= a; ++b) {
                    acc[b] = fma(t[b].x, tv.x, acc[b]);
                    acc[b] = fma(t[b].y, tv.y, acc[b]);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < C::A; ++b) {
            const int i = C::W * b + c.wc;
            const double v = reduce_q(acc[b]);
            if (q == 0 && i < T) C::part(C::W + c.wr)[8 * i + g] = v;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// QP: cvxopt coneqp for the orthant cone with G = -I (see oracle/coneqp.py for the restatement and
// its provenance).  Thread i < n owns element i of every vector.
// ------------------------------------------------------------------------------------------------
struct QpOut {
    double xi;
    double pcost;
    int iters;
    int status;  // HDRT_ST_QP_MAXITERS / HDRT_ST_KKT_FAIL bits
    bool fatal;  // Cholesky failed before the first iterate existed (cvxopt raises ValueError)
};

template <class C>
__device__ __forceinline__ double sum_parts(int first, int count, int t) {
    double v = 0.0;
#pragma unroll
    for (int p = 0; p < C::NPART; ++p)
        if (p >= first && p < first + count) v += C::part(p)[t];
    return v;
}

template <class C>
__device__ __noinline__ QpOut qp_phase(Ctx& cref) {
    Ctx c = cref;
    const int tid = threadIdx.x;
    const int n = c.n;
    const bool act = tid < n;
    double* bs = C::vec(C::BS);
    const double qi = act ? C::vec(C::QS)[tid] : 0.0;
    const double hi = act ? c.hvec[tid] : 0.0;
    QpOut out;
    out.xi = 0.0; out.pcost = 0.0; out.iters = 0; out.status = 0; out.fatal = false;

    double resx0, resz0;
    {
        double t2[2] = {qi * qi, hi * hi};
        block_reduce<C, 2, 0u>(t2, c);
        resx0 = fmax(1.0, sqrt(t2[0]));
        resz0 = fmax(1.0, sqrt(t2[1]));
    }
    // -P sits in tensor memory (gram_phase); the shared-memory tile area belongs to the factor
    PROF_DECL;
    double xi = 0.0, si = 1.0, zi = 1.0, di = 1.0, dinv = 1.0, lam = 1.0;
    double rxi = 0.0, rzi = 0.0, gap = 0.0, pcost = 0.0;
    double pxi = 0.0;         // (P x)_i
    int iters;
    // iters == -1 is the initial point (W = I); 0.. are the interior-point iterations
#pragma unroll 1
    for (iters = -1; iters <= kMaxIpm; ++iters) {
        if (iters >= 0) {
            // P x is carried along instead of recomputed: every solve H u = b of this loop has H = P + diag(dinv^2), so
            // (P u)_i = b_i - dinv_i^2 u_i costs the owner of element i two operations and no communication
            const double px = pxi;
            rxi = px + qi;
            const double f0p = act ? (xi * rxi + xi * qi) : 0.0;
            rxi -= zi;
            rzi = si - hi - xi;
            double t5[5] = {f0p, act ? rxi * rxi : 0.0, act ? rzi * rzi : 0.0, act ? zi * rzi : 0.0,
                            act ? (iters == 0 ? si * zi : lam * lam) : 0.0};
            block_reduce<C, 5, 0u>(t5, c);
            const double f0 = 0.5 * t5[0];
            const double resx = sqrt(t5[1]), resz = sqrt(t5[2]);
            gap = t5[4];
            pcost = f0;
            const double dcost = f0 + t5[3] - gap;
            double relgap = 0.0;
            bool have_rel = true;
            if (pcost < 0.0) relgap = gap / -pcost;
            else if (dcost > 0.0) relgap = gap / dcost;
            else have_rel = false;
            const double pres = resz / resz0, dres = resx / resx0;
            const bool done = (pres <= kFeasTol) && (dres <= kFeasTol) &&
                              ((gap <= kAbsTol) || (have_rel && relgap <= kRelTol));
            if (done) break;
            if (iters == kMaxIpm) { out.status |= HDRT_ST_QP_MAXITERS; break; }
            if (iters == 0) {
                di = sqrt(si / zi);
                dinv = 1.0 / di;
                lam = sqrt(si * zi);
            }
        }
        if (act) C::vec(C::DSQ)[tid] = dinv * dinv;
        __syncthreads();
        PROF_ADD(8);
        const bool fact_ok = factor_chol<C>(c);
        PROF_ADD(9);
        if (!fact_ok) {
            out.status |= HDRT_ST_KKT_FAIL;
            if (iters <= 0) { out.fatal = true; xi = nan(""); }
            break;
        }
        // One instance of the solve serves the initial point (a single pass with rhs -q - h) and the two Mehrotra
        // passes of an iteration.
        const bool start = iters < 0;
        const double lamsq = lam * lam;
        const double mu = gap / (double)n;
        double sigma = 0.0, step = 1.0;
        double ws3 = 0.0, dxi = 0.0, dsi = 0.0, dzi = 0.0, zs = 0.0, rhs = 0.0;
#pragma unroll 1
        for (int pass = start ? 1 : 0; pass < 2; ++pass) {
            if (start) {
                // solve [P+I] x = -q - h ; z = -x - h ; s = -z, shifted into the cone
                rhs = -qi - hi;
                if (act) bs[tid] = rhs;
            } else {
                dsi = 0.0;
                if (pass == 1) dsi -= ws3;
                dsi -= lamsq;
                dsi += sigma * mu;
                dxi = -rxi;
                dzi = -rzi;
                dsi = dsi / lam;
                dzi = dzi - di * dsi;
                zs = dinv * dzi;
                rhs = dxi - dinv * zs;
                if (act) bs[tid] = rhs;
            }
            __syncthreads();
            solve_kkt<C>(c, bs);
            dxi = act ? sum_parts<C>(C::W, C::W, tid) : 0.0;
            if (start) {
                xi = dxi;
                pxi = act ? rhs - xi : 0.0;          // (P + I) x = rhs
                zi = -xi - hi;
                si = -zi;
                double t4[4] = {act ? si * si : 0.0, act ? -si : -INFINITY, act ? zi * zi : 0.0, act ? -zi : -INFINITY};
                block_reduce<C, 4, 0xAu>(t4, c);
                const double nrms = sqrt(t4[0]), ts = t4[1], nrmz = sqrt(t4[2]), tz = t4[3];
                if (ts >= -1e-8 * fmax(nrms, 1.0)) si += 1.0 + ts;
                if (tz >= -1e-8 * fmax(nrmz, 1.0)) zi += 1.0 + tz;
                break;
            }
            dzi = -dinv * dxi - zs;
            dsi = dsi - dzi;
            const double prod = dsi * dzi;
            if (pass == 0) ws3 = prod;
            dsi = dsi / lam;
            dzi = dzi / lam;
            double t3[3] = {act ? prod : 0.0, act ? -dsi : -INFINITY, act ? -dzi : -INFINITY};
            block_reduce<C, 3, 0x6u>(t3, c);
            const double t = fmax(0.0, fmax(t3[1], t3[2]));
            if (t == 0.0) step = 1.0;
            else if (pass == 0) step = fmin(1.0, 1.0 / t);
            else step = fmin(1.0, kStep / t);
            if (pass == 0) {
                const double sg = fmin(1.0, fmax(0.0, 1.0 - step + t3[0] / gap * (step * step)));
                sigma = sg * sg * sg;
            }
        }
        if (start) continue;
        PROF_ADD(10);
        pxi = act ? fma(step, rhs - (dinv * dinv) * dxi, pxi) : 0.0;
        xi += step * dxi;
        dsi = step * dsi + 1.0;
        dzi = step * dzi + 1.0;
        dsi *= lam;
        dzi *= lam;
        const double sqs = sqrt(dsi), sqz = sqrt(dzi);
        di = di * sqs / sqz;
        dinv = 1.0 / di;
        lam = sqs * sqz;
        si = lam * di;
        zi = lam * dinv;
    }
    out.xi = xi;
    out.pcost = pcost;
    out.iters = iters < 0 ? 0 : iters;
    cref.red_phase = c.red_phase;
    return out;
}
// ------------------------------------------------------------------------------------------------
// Hyper-parameter updates for one coefficient block (DRT or DOP): qphb.solve_s / solve_rho
// ------------------------------------------------------------------------------------------------
struct BlockHyp {
    double dw[3], sigma[3], s_alpha[3], s_0[3], rho_alpha[3], rho_0[3];
    bool use_gmat;  // DRT block: k = 0 gets G = Xh M1 Xh (qphb.py:769-772); DOP block: 0 (drt1d.py quirk)
};

template <class C>
__device__ __noinline__ void hyper_block(Ctx& cref, const BlockHyp& hpref, int start, int len, double* rho, double* xmx,
                                         bool first_iter) {
    Ctx c = cref;
    const BlockHyp hp = hpref;
    const int tid = threadIdx.x;
    const int n = c.n, nn = c.n * c.n;
    const bool act = tid < len;
    const int gi = start + tid;
    const double* xs = C::vec(C::XS);
    double* xh = C::vec(C::XH);
    PROF_DECL;
    const double xi = act ? xs[gi] : 0.0;
    if (act) {
        const double ax = fabs(xi);
        xh[gi] = (xi > 0.0 ? 1.0 : (xi < 0.0 ? -1.0 : 0.0)) * sqrt(ax);
    }
    __syncthreads();
    const double xhi = act ? xh[gi] : 0.0;
    double bsum[3] = {0, 0, 0}, gd[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    if (act) {
        const double inv2s0 = 1.0 / (2.0 * hp.sigma[0] * hp.sigma[0]);
        const double* __restrict__ pcol = c.pen + (start * n + gi);  // symmetric: read column-wise (coalesced)
        // rows in batches of eight: the 24 loads (L2 hits, several hundred cycles each) are issued together
#pragma unroll 1
        for (int j0 = 0; j0 < len; j0 += 4) {
            double mm[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jj = min(j0 + u, len - 1);
#pragma unroll
                for (int k = 0; k < 3; ++k) mm[u][k] = pcol[k * nn + jj * n];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j < len) {
                    const int gj = start + j;
                    const double xj = xs[gj];
                    double gam[3] = {(xi * mm[u][0]) * xj, (xi * mm[u][1]) * xj, (xi * mm[u][2]) * xj};
                    if (hp.use_gmat) gam[0] += ((xhi * mm[u][1]) * xh[gj]) * inv2s0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (j == tid) {
                            gd[k] = gam[k] + (hp.s_alpha[k] - 1.0) / hp.s_0[k];
                        } else {
                            const double g = gam[k] * C::vec(C::US0 + k)[gj];
                            bsum[k] += g;
                            mx[k] = fmax(mx[k], fabs(g));
                        }
                    }
                }
            }
        }
    }
    PROF_ADD(13);
    block_reduce<C, 3, 0x7u>(mx, c);
    if (act) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!(hp.dw[k] > 0.0)) continue;
            const double am1 = hp.s_alpha[k] - 1.0;
            double s_hat;
            if (mx[k] > 1e-10) {
                const double b = bsum[k];
                const double sg = (b > 0.0 ? 1.0 : (b < 0.0 ? -1.0 : 0.0));
                const double u = (-b + sg * sqrt(b * b + 4.0 * gd[k] * am1)) / (2.0 * gd[k]);
                s_hat = u * u;
            } else {
                s_hat = am1 / gd[k];
            }
            if (isnan(s_hat)) s_hat = 1.0;
            if (s_hat <= 0.0) s_hat = 1e-15;
            C::vec(C::SV0 + k)[gi] = s_hat;
        }
    }
    __syncthreads();
    if (act) {
#pragma unroll
        for (int k = 0; k < 3; ++k) C::vec(C::US0 + k)[gi] = sqrt(C::vec(C::SV0 + k)[gi]);
    }
    __syncthreads();
    // rho: alpha / (x' S^1/2 M S^1/2 x / xmx + beta)
    double tr[3] = {0, 0, 0}, tx[3] = {0, 0, 0};
    if (act) {
        const double* __restrict__ pcol = c.pen + (start * n + gi);
#pragma unroll 1
        for (int j0 = 0; j0 < len; j0 += 4) {
            double mm[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jj = min(j0 + u, len - 1);
#pragma unroll
                for (int k = 0; k < 3; ++k) mm[u][k] = pcol[k * nn + jj * n];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j < len) {
                    const int gj = start + j;
                    const double xj = xs[gj];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        tr[k] += (xj * C::vec(C::US0 + k)[gj]) * mm[u][k];
                        tx[k] += xj * mm[u][k];
                    }
                }
            }
        }
    }
    PROF_ADD(14);
    double t6[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t6[k] = act ? (tr[k] * C::vec(C::US0 + k)[gi]) * xi : 0.0;
        t6[3 + k] = act ? tx[k] * xi : 0.0;
    }
    block_reduce<C, 6, 0u>(t6, c);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (hp.dw[k] > 0.0) {
            const double beta = hp.rho_alpha[k] / hp.rho_0[k];
            rho[k] = hp.rho_alpha[k] / (t6[k] / xmx[k] + beta);
        }
    }
    if (first_iter) {
#pragma unroll
        for (int k = 0; k < 3; ++k) xmx[k] = t6[3 + k];
    }
    cref.red_phase = c.red_phase;
}

// ------------------------------------------------------------------------------------------------
// Rows rb + u kWarps (u < RU) of vmm @ uin for one warp; uin is a per-row vector in shared memory.  vmm is block
// diagonal: chrono rows x chrono columns (dense, or NULL = uniform: the caller substitutes the mean), EIS rows x
// EIS columns.  The sums are warp-reduced (valid in every lane).
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void vmm_rows(const Ctx& c, const double* uin, int rb, double (&sh)[C::RU]) {
    constexpr int RU = C::RU, CU = (C::NV + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int N = c.N, nc = c.nc;
#pragma unroll
    for (int u = 0; u < RU; ++u) sh[u] = 0.0;
    const bool all_eis = rb >= nc;   // rows of a pass are ascending: every row of it is an EIS row
    if (all_eis) {
        const int ne = N - nc;
        for (int c0 = 0; c0 < ne; c0 += 32 * CU) {
            double v[RU][CU];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const double* __restrict__ vr = c.vmm_eis + (size_t)(min(rb + u * C::kWarps, N - 1) - nc) * ne;
#pragma unroll
                for (int w = 0; w < CU; ++w) v[u][w] = vr[min(c0 + lane + 32 * w, ne - 1)];
            }
#pragma unroll
            for (int w = 0; w < CU; ++w) {
                const int col = c0 + lane + 32 * w;
                const double rr = (col < ne) ? uin[nc + col] : 0.0;
#pragma unroll
                for (int u = 0; u < RU; ++u) sh[u] = fma(v[u][w], rr, sh[u]);
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = rb + u * C::kWarps;
            if (r < N) {
                if (r < nc) {
                    if (c.vmm_chrono != nullptr) {
                        // dense chrono block (flexible error structure): rows are long (n_chrono columns); four
                        // independent loads in flight per lane
                        const double* __restrict__ vr = c.vmm_chrono + (size_t)r * nc;
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                        int col = lane;
                        for (; col + 96 < nc; col += 128) {
                            const double v0 = vr[col], v1 = vr[col + 32], v2 = vr[col + 64], v3 = vr[col + 96];
                            a0 = fma(v0, uin[col], a0);
                            a1 = fma(v1, uin[col + 32], a1);
                            a2 = fma(v2, uin[col + 64], a2);
                            a3 = fma(v3, uin[col + 96], a3);
                        }
                        for (; col < nc; col += 32) a0 = fma(vr[col], uin[col], a0);
                        sh[u] = (a0 + a1) + (a2 + a3);
                    }
                } else {
                    const int ne = N - nc;
                    const double* __restrict__ vr = c.vmm_eis + (size_t)(r - nc) * ne;
                    for (int col = lane; col < ne; col += 32) sh[u] = fma(vr[col], uin[nc + col], sh[u]);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) sh[u] = warp_sum(sh[u]);
}

// diag(vmm)[r] (qphb.py:1646: the first initialisation pass with outliers excludes the point itself)
__device__ __forceinline__ double vmm_diag_at(const Ctx& c, int r) {
    if (r < c.nc) return c.vmm_chrono ? c.vmm_chrono[(size_t)r * c.nc + r] : 1.0 / (double)c.nc;
    const int ne = c.N - c.nc;
    return c.vmm_eis[(size_t)(r - c.nc) * ne + (r - c.nc)];
}

// Squared residuals of all rows into r2[]; with VZ the vz_offset column of a hybrid fit is rewritten from the same pass
// (drt1d.py:972-979: the prediction without the vz_offset and v_baseline columns, sign by domain, times vz_strength).
// The residual itself uses the column as it stood when the pass started.
template <class C, bool VZ>
__device__ __forceinline__ void residual_rows(const Ctx& c, const double (&xw)[(C::NV + 31) / 32], const double (&xv)[(C::NV + 31) / 32],
                                              double x_vz) {
    constexpr int CU = (C::NV + 31) / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = c.N, n = c.n;
    double* r2 = C::rowr2(N);
#pragma unroll 1
    for (int rb = 8 * warp; rb < N; rb += 8 * C::kWarps) {
        double v[8][CU];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double* __restrict__ src = c.rm + (size_t)min(rb + u, N - 1) * n;
#pragma unroll
            for (int w = 0; w < CU; ++w) v[u][w] = src[min(lane + 32 * w, n - 1)];
        }
        // the lanes with lane % 4 == 0 finish row rb + lane / 4
        const int r = min(rb + (lane >> 2), N - 1);
        const double rvr = C::EXT ? c.rv[r] * c.rv_scale : c.rv[r];
        double vzr = 0.0, vzs = 0.0, vz0r = 0.0;
        if (c.vz >= 0) vzr = c.vzcol[r];
        if (VZ) {
            vzs = c.vz_strength[r];
            if (C::EXT && c.vz0 != nullptr) vz0r = c.vz0[r];
        }
        double acc[8], accv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc[u] = 0.0;
            accv[u] = 0.0;
#pragma unroll
            for (int w = 0; w < CU; ++w) {
                acc[u] = fma(v[u][w], xw[w], acc[u]);         // xw is zero beyond n and at column vz
                if (VZ) accv[u] = fma(v[u][w], xv[w], accv[u]);
            }
        }
        const double tot = warp_reduce8(acc, lane);
        double totv = 0.0;
        if (VZ) totv = warp_reduce8(accv, lane);
        if ((lane & 3) == 0 && rb + (lane >> 2) < N) {
            const double resid = fma(vzr, x_vz, tot) - rvr;
            r2[r] = resid * resid;
            if (VZ) {
                // a continuation predicts with the vz_offset column it started from (drt1d.py:1296-1302 copies
                // the matrix once, with that column in place); the plain fit copies it while it is still zero
                const double pred = fma(vz0r, x_vz, totv);
                c.vzcol[r] = ((r < c.nc) ? pred : -pred) * vzs;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Error-structure weights (qphb.estimate_weights, qphb.py:1545-1594) + vz_offset column rewrite
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __noinline__ void weights_phase(Ctx& cref, const double* est, double var_floor, bool update_vz, bool base) {
    Ctx c = cref;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = c.N, n = c.n, nc = c.nc;
    const double* xs = C::vec(C::XS);
    PROF_DECL;
    // residuals: a warp takes eight consecutive rows per pass.  All loads of a pass (L2 hits, several hundred cycles
    // each) are issued before the first use -- row / column indices are clamped instead of branched around -- and the
    // eight (sixteen, with the vz_offset prediction) row sums cross the warp in one transposed reduction.
    constexpr int CU = (C::NV + 31) / 32;
    double xw[CU];     // this lane's coefficients (DOP ones with the column rescale of solve_rp folded in); not column vz
    double xv[CU];     // ... and without the v_baseline columns: the prediction that sets the vz_offset column
#pragma unroll
    for (int w = 0; w < CU; ++w) {
        const int col = lane + 32 * w;
        xw[w] = (col < n && col != c.vz) ? xs[col] : 0.0;
        if (C::EXT && col >= c.dop_a && col < c.dop_b) xw[w] *= c.dop_cs;
        xv[w] = (col < c.vb_a || col >= c.vb_b) ? xw[w] : 0.0;
    }
    const double x_vz = (c.vz >= 0) ? xs[c.vz] : 0.0;
    if (update_vz) residual_rows<C, true>(c, xw, xv, x_vz);
    else residual_rows<C, false>(c, xw, xv, x_vz);
    __syncthreads();
    PROF_ADD(11);
    constexpr int RU = C::RU;
    const bool uniform_chrono = nc > 0 && c.vmm_chrono == nullptr;
    const bool outl = C::EXT && c.outlier_p >= 0.0;
    base = C::EXT && base;
    const double* uin = C::rowr2(c.N);
    // with the 'uniform' chrono error structure the chrono rows of vmm are never touched: their estimate is the mean
    const int row0 = uniform_chrono ? nc : 0;
    if (outl) {
        // qphb.solve_outlier_t (qphb.py:1497-1519): s_bar = vmm r^2, t = 1 - P(outlier | r); the averaging matrix
        // then becomes T^1/2 vmm T^1/2 + (I - T) (qphb.outlier_tvt :1522-1538), applied below without forming it
        const double p = c.outlier_p;
        double mean1 = 0.0;
        if (uniform_chrono) {
            double t1[1] = {0.0};
            for (int r = tid; r < nc; r += C::kThreads) t1[0] += uin[r];
            block_reduce<C, 1, 0u>(t1, c);
            mean1 = t1[0] / (double)nc;
        }
        auto outlier_row = [&](int r, double s_bar) {
            const double r2 = uin[r];
            if (base) { const double d = vmm_diag_at(c, r); s_bar = (s_bar - d * r2) / (1.0 - d); }
            const double sd = sqrt(s_bar), ar = sqrt(r2);
            const double k2pi = 2.5066282746310002;   // sqrt(2 pi)
            const double pdf_in = 1.0 / (sd * k2pi) * exp(-0.5 * r2 / (sd * sd));
            const double pdf_out = 1.0 / (ar * k2pi) * exp(-0.5 * r2 / (ar * ar));
            double t = 1.0 - p * pdf_out / ((1.0 - p) * pdf_in + p * pdf_out);
            if (sd > ar) t = 1.0;
            c.t_out[r] = t;
            C::rowu(c.N)[r] = sqrt(t) * r2;
        };
        for (int r = tid; r < row0; r += C::kThreads) outlier_row(r, mean1);
        for (int rb = row0 + warp; rb < N; rb += RU * C::kWarps) {
            double sh[RU];
            vmm_rows<C>(c, uin, rb, sh);
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const int r = rb + u * C::kWarps;
                if (lane == 0 && r < N) outlier_row(r, sh[u]);
            }
        }
        __syncthreads();
        uin = C::rowu(c.N);
    }
    double chrono_mean = 0.0;
    if (uniform_chrono) {
        double t1[1] = {0.0};
        for (int r = tid; r < nc; r += C::kThreads) t1[0] += uin[r];
        block_reduce<C, 1, 0u>(t1, c);
        chrono_mean = t1[0] / (double)nc;
    }
    // variance estimate s_hat = vmm_eff r2 (block diagonal: chrono rows x chrono columns, EIS rows x EIS columns)
    auto weight_row = [&](int r, double s_hat) {
        if (base) { const double d = vmm_diag_at(c, r); s_hat = (s_hat - d * uin[r]) / (1.0 - d); }
        if (outl) {
            const double t = c.t_out[r];
            s_hat = sqrt(t) * s_hat + (1.0 - t) * C::rowr2(c.N)[r];
        }
        if (s_hat < var_floor) s_hat = var_floor;
        double w = 1.0 / sqrt(s_hat);
        if (est != nullptr) {
            const double e = est[r];
            const double frac = w / (w + e);
            w = frac * w + (1.0 - frac) * e;
        }
        C::roww()[r] = fmax(w, 1e-10);
    };
    for (int r = tid; r < row0; r += C::kThreads) weight_row(r, chrono_mean);
    for (int rb = row0 + warp; rb < N; rb += RU * C::kWarps) {
        double sh[RU];
        vmm_rows<C>(c, uin, rb, sh);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = rb + u * C::kWarps;
            if (lane == 0 && r < N) weight_row(r, sh[u]);
        }
    }
    __syncthreads();
    PROF_ADD(12);
    cref.red_phase = c.red_phase;
}

// ------------------------------------------------------------------------------------------------
// Post-fit diagnostics for the mapping path (DRT.estimate_distribution_cov, drt1d.py:3063-3151, with
// estimate_param_cov :4116-4138): diag(B P^-1 B^T) for the rows b_t of the evaluation matrix B, in the scaled
// space (the host multiplies by coefficient_scale^2).  P = L L^T is factorised by the same tile sweep as the QP's
// KKT matrices (its negated tiles are in shared memory after the calculate_pq Gram pass); then
// diag_t = |U^T b_t|^2 with U = -L^-T.
// Returns false if P is not positive definite (the reference warns 'Singular P matrix' and reports no covariance).
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __noinline__ bool postfit_variance(Ctx& cref, const double* __restrict__ eval_mat, int n_eval, double* out) {
    Ctx c = cref;
    const int tid = threadIdx.x, n = c.n;
    double* bs = C::vec(C::BS);
    if (tid < C::NV) C::vec(C::DSQ)[tid] = 0.0;
    __syncthreads();
    const bool ok = factor_chol<C>(c);
    if (ok) {
        double2 S[C::NSLOT];
        load_u_tiles<C>(c, S);
#pragma unroll 1
        for (int t = 0; t < n_eval; ++t) {
            if (tid < n) bs[tid] = eval_mat[(size_t)t * n + tid];
            __syncthreads();
            matvec_ut<C>(c, S, bs);
            __syncthreads();
            double v[1] = {0.0};
            if (tid < n) {
                const double y = sum_parts<C>(0, C::W, tid);
                v[0] = y * y;
            }
            block_reduce<C, 1, 0u>(v, c);
            if (tid == 0) out[t] = v[0];
        }
    }
    if (tid < C::NV) bs[tid] = 0.0;   // the padding entries of the solve right-hand side must read as zero
    __syncthreads();
    cref.red_phase = c.red_phase;
    return ok;
}

// ------------------------------------------------------------------------------------------------
// One spectrum.  The outer loop runs phase -1 (initialize_weights), 0..max_iter-1 (iterate_qphb) and,
// when P/q are requested, one final Gram-only phase (calculate_pq) through the same code.
template <class C>
__device__ __forceinline__ void fit_one(const hdrt_qphb_problem& p, int b, Ctx& c) {
    const int tid = threadIdx.x;
    const int N = p.n_rows, n = p.n_cols;
    const hdrt_hypers& hy = p.hyp;
    c.rm = p.rm + (size_t)b * p.rm_stride;
    c.rv = p.rv + (size_t)b * N;
    c.vmm_eis = p.vmm_eis ? p.vmm_eis + (size_t)b * p.vmm_eis_stride : nullptr;
    c.vmm_chrono = p.vmm_chrono ? p.vmm_chrono + (size_t)b * p.vmm_chrono_stride : nullptr;
    c.pen = p.pen + (size_t)b * p.pen_stride;
    c.vzcol = p.vz_col ? p.vz_col + (size_t)b * N : nullptr;
    const bool outl = C::EXT && hy.has_outlier_p != 0;
    c.outlier_p = outl ? hy.outlier_p : -1.0;
    c.t_out = outl ? p.outlier_t + (size_t)b * N : nullptr;
    c.vz0 = nullptr;
    c.rv_scale = 1.0;
    c.dop_cs = 1.0;
    double us_factor = 1.0;   // product of the update_scale factors (drt1d.py:914-936)
    double* est_g = p.est_weights + (size_t)b * N;

    // var floor = var(y) * 1e-7 (qphb.py:1560-1561)
    double var_floor;
    {
        double t1[1] = {0.0};
        for (int r = tid; r < N; r += C::kThreads) t1[0] += c.rv[r];
        block_reduce<C, 1, 0u>(t1, c);
        const double mean = t1[0] / (double)N;
        double t2[1] = {0.0};
        for (int r = tid; r < N; r += C::kThreads) { const double d = c.rv[r] - mean; t2[0] += d * d; }
        block_reduce<C, 1, 0u>(t2, c);
        var_floor = (t2[0] / (double)N) * 1e-7;
    }

    double rho[3], dop_rho[3], xmx[3] = {1, 1, 1}, dop_xmx[3] = {1, 1, 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) { rho[k] = hy.rho_0[k]; dop_rho[k] = hy.dop_rho_0[k]; }
    // PFRT (DRT._pfrt_fit_core, drt1d.py:2558-2698): step 0 is a plain fit at s_0 * f_0, l2_lambda_0 / f_0; every
    // further factor continues from the previous state (_continue_from_init, :1270-1365)
    const bool pfrt = C::EXT && p.n_pfrt > 0;
    const int n_steps = pfrt ? p.n_pfrt : 1;
    const double fac0 = pfrt ? p.pfrt_factors[0] : 1.0;
    const double lam_init = hy.l2_lambda_0 / fac0;     // qphb_params['hypers']['l2_lambda_0'] of the reference
    if (tid < C::NV) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double s0 = hy.s_0[k] * fac0;        // s_0 of the step hypers fills the whole vector (drt1d.py:565)
            C::vec(C::SV0 + k)[tid] = s0;
            C::vec(C::US0 + k)[tid] = sqrt(s0);
        }
    }
    for (int r = tid; r < N; r += C::kThreads) {
        C::roww()[r] = 1.0;
        if (c.vz >= 0) c.vzcol[r] = 0.0;
    }
    __syncthreads();

    BlockHyp hd, hp;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        hd.dw[k] = hy.derivative_weights[k]; hd.sigma[k] = hy.sigma_ds[k]; hd.s_alpha[k] = hy.s_alpha[k];
        hd.s_0[k] = hy.s_0[k]; hd.rho_alpha[k] = hy.rho_alpha[k]; hd.rho_0[k] = hy.rho_0[k];
        hp.dw[k] = hy.dop_derivative_weights[k]; hp.sigma[k] = hy.dop_sigma_ds[k]; hp.s_alpha[k] = hy.dop_s_alpha[k];
        hp.s_0[k] = hy.dop_s_0[k]; hp.rho_alpha[k] = hy.dop_rho_alpha[k]; hp.rho_0[k] = hy.dop_rho_0[k];
    }
    hd.use_gmat = true;
    hp.use_gmat = false;

    // chrono / EIS weight factors of a hybrid fit (drt1d.py:745-800, applied at :882-884): the scalars of hyp, or one
    // pair per spectrum from the host (hybrid_weight_factor_method='rp'), or computed from the initial weight scales
    // after the initialisation ('weight')
    double cwf = hy.chrono_weight_factor, ewf = hy.eis_weight_factor;
    if (C::EXT && p.hybrid_wf_in) { cwf = p.hybrid_wf_in[2 * (size_t)b]; ewf = p.hybrid_wf_in[2 * (size_t)b + 1]; }
    const bool separate = C::EXT && hy.init_weights_separately != 0 && c.nc > 0 && c.nc < N;
    double var_floor_c = var_floor, var_floor_e = var_floor;
    if (separate) {   // qphb.estimate_weights floors the variance at var(y) 1e-7 of the vector it is given (drt1d.py:647-669)
        double t2[2] = {0.0, 0.0};
        for (int r = tid; r < N; r += C::kThreads) { if (r < c.nc) t2[0] += c.rv[r]; else t2[1] += c.rv[r]; }
        block_reduce<C, 2, 0u>(t2, c);
        const double mc = t2[0] / (double)c.nc, me = t2[1] / (double)(N - c.nc);
        double v2[2] = {0.0, 0.0};
        for (int r = tid; r < N; r += C::kThreads) {
            const double d = c.rv[r] - (r < c.nc ? mc : me);
            if (r < c.nc) v2[0] += d * d; else v2[1] += d * d;
        }
        block_reduce<C, 2, 0u>(v2, c);
        var_floor_c = v2[0] / (double)c.nc * 1e-7;
        var_floor_e = v2[1] / (double)(N - c.nc) * 1e-7;
    }
    int status = 0, n_ipm = 0;
    PROF_DECL;
    L2Factors f;
#pragma unroll
    for (int k = 0; k < 3; ++k) f.use[k] = hy.derivative_weights[k] > 0.0;
    double xi = 1e-6;  // drt1d.py:612
    double fun = 0.0;
    int it = 0, n_outer0 = 0;
    bool conv = false, fatal = false, final_pq = false;
#pragma unroll 1
    for (int step = 0; step < n_steps && !fatal; ++step) {
    const bool cont = C::EXT && step > 0;
    const double fac = pfrt ? p.pfrt_factors[step] : 1.0;
    const double lam_step = hy.l2_lambda_0 / fac;
#pragma unroll
    for (int k = 0; k < 3; ++k) hd.s_0[k] = hy.s_0[k] * fac;
    const int max_it = cont ? p.pfrt_max_iter : hy.max_iter;
    const bool solve_rp = C::EXT && hy.solve_rp != 0;
    it = cont ? 0 : (solve_rp ? -3 : (outl ? -2 : (separate ? -5 : -1)));   // -3: estimate_x_rp (qphb.py:1684-1717);
                                        // -5 / -4: initialize_weights on the chrono rows, then on the EIS rows; -1: initialize_weights (drt1d.py:640-675, qphb.py:1609-1681); with outlier_p
                                        // the initialisation runs twice (-2, -1), the second time weighted by the
                                        // first estimate
    conv = false;
    final_pq = false;
    if (cont && c.vz >= 0) {
        c.vz0 = p.vz_scratch + (size_t)b * N;
        for (int r = tid; r < N; r += C::kThreads) c.vz0[r] = c.vzcol[r];
        __syncthreads();
    }
#pragma unroll 1
    while (true) {
        const bool init = it < 0;
        if (C::EXT && hy.update_scale && !init && !cont && !final_pq && it > 1) {
            // keep the data at the requested rp_scale as the Rp estimate improves (drt1d.py:914-936)
            const bool drt = tid >= c.ns && tid < n;
            double t1[1] = {drt ? fabs(xi) : 0.0};
            block_reduce<C, 1, 0u>(t1, c);
            const double sf = sqrt(hy.rp_scale / (t1[0] * hy.basis_area));   // damped (drt1d.py:919)
            xi *= sf;
            c.rv_scale *= sf;
            us_factor *= sf;
            var_floor *= sf * sf;
            const double sq = sqrt(sf);
#pragma unroll
            for (int k = 0; k < 3; ++k) { xmx[k] *= sq; dop_xmx[k] *= sq; }
            for (int r = tid; r < N; r += C::kThreads) {
                est_g[r] /= sf;
                C::roww()[r] /= sf;
            }
            __syncthreads();
        }
        const double x_in = xi;
        if (C::EXT && (it == -5 || it == -4)) {   // one domain at a time: the other rows carry no weight
            for (int r = tid; r < N; r += C::kThreads) C::roww()[r] = ((r < c.nc) == (it == -5)) ? 1.0 : 0.0;
            __syncthreads();
        }
        // weights entering the Gram: 1 (init) / weight factors (drt1d.py:881-892) / scaled weights (:991-1008)
        if (!init) {
            for (int r = tid; r < N; r += C::kThreads) {
                double w = C::roww()[r];
                const double wf = (C::EXT && p.weight_factor_vec) ? p.weight_factor_vec[r] : hy.weight_factor;
                if (final_pq) {
                    w *= wf;
                    if (p.hybrid) w *= (r < c.nc) ? cwf : ewf;
                } else {
                    if (p.hybrid) w *= (r < c.nc) ? cwf : ewf;
                    if (it > 0 || cont) w = w * wf;   // a continuation scales on every pass (drt1d.py:1320)
                }
                C::roww()[r] = w;
            }
        }
        {
            const double lam_iw = (C::EXT && it == -3) ? 1e-4 : hy.iw_l2_lambda_0;   // estimate_x_rp: l2_lambda_0 = 1e-4 (drt1d.py:5425)
            const double lam0 = init ? lam_iw : lam_step;
            const double dlam0 = init ? hy.dop_l2_lambda_0 * (lam_iw / lam_step) : hy.dop_l2_lambda_0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                f.drt[k] = lam0 * hy.derivative_weights[k] * rho[k];
                f.dop[k] = dlam0 * hy.dop_derivative_weights[k] * dop_rho[k];
            }
        }
        __syncthreads();
        PROF_ADD(0);
        gram_phase<C>(c, f, init, (C::EXT && it == -3) ? 1e-3 : hy.iw_l1_lambda_0, (final_pq && p.p_matrix) ? p.p_matrix + (size_t)b * n * n : nullptr,
                        (final_pq && p.q_vector) ? p.q_vector + (size_t)b * n : nullptr);
        if (final_pq) {
            if (p.dist_var && !postfit_variance<C>(c, p.eval_mat, p.n_eval, p.dist_var + (size_t)b * p.n_eval))
                status |= HDRT_ST_COV_FAIL;
            break;
        }

        PROF_ADD(1);
        QpOut qo = qp_phase<C>(c);
        PROF_ADD(2);
        status |= qo.status;
        n_ipm += qo.iters;
        if (qo.fatal) { fatal = true; xi = qo.xi; break; }
        if (tid < n) C::vec(C::XS)[tid] = qo.xi;
        __syncthreads();
        if (C::EXT && it == -3) {
            // DRT._solve_data_scale (drt1d.py:5421-5437, applied at :573-607): Rp from a lightly regularised
            // elastic-net solution; the data are rescaled to rp_scale, the DOP columns to the DRT magnitude
            const bool drt = tid >= c.ns && tid < n;
            const bool dopc = tid >= c.dop_a && tid < c.dop_b;
            double t3[3] = {drt ? fabs(qo.xi) : 0.0, drt ? fabs(qo.xi) : 0.0, dopc ? fabs(qo.xi) : 0.0};
            block_reduce<C, 3, 0x6u>(t3, c);
            const double sf = hy.rp_scale / (t3[0] * hy.basis_area);
            c.rv_scale *= sf;
            var_floor *= sf * sf;
            if (c.dop_a >= 0 && hy.normalize_dop) c.dop_cs = 1.0 / (t3[1] / t3[2]);
            if (tid == 0 && p.scale_factors) p.scale_factors[3 * (size_t)b] = sf;
            it = outl ? -2 : -1;
            continue;
        }
        if (C::EXT && it == -5) {                 // chrono rows done: keep their estimate, go on to the EIS rows
            if (tid < n && p.x_overfit) p.x_overfit[(size_t)b * n + tid] = qo.xi;
            weights_phase<C>(c, nullptr, var_floor_c, false, false);
            for (int r = tid; r < c.nc; r += C::kThreads) est_g[r] = C::roww()[r];
            __syncthreads();
            it = -4;
            continue;
        }
        if (init) {
            if (C::EXT && it == -4) {
                if (tid < n && p.x_overfit_eis) p.x_overfit_eis[(size_t)b * n + tid] = qo.xi;
                weights_phase<C>(c, nullptr, var_floor_e, false, false);
                for (int r = tid; r < c.nc; r += C::kThreads) C::roww()[r] = est_g[r];
                __syncthreads();
            } else {
                if (tid < n && p.x_overfit) p.x_overfit[(size_t)b * n + tid] = qo.xi;
                weights_phase<C>(c, nullptr, var_floor, false, outl);
            }
            PROF_ADD(4);
            if (it == -2) { it = -1; continue; }   // qphb.py:1634-1655: second pass with w = est_weights
            for (int r = tid; r < N; r += C::kThreads) {
                const double e = C::roww()[r];
                est_g[r] = e;
                double wi = e;
                if (hy.has_iw_prior) {  // qphb.solve_init_weight_scale, qphb.py:1471-1479
                    const double bq = 0.5 - hy.iw_alpha + 1.0;
                    const double s_hat = (-bq + sqrt(bq * bq + 2.0 * hy.iw_beta / (e * e))) / (2.0 * hy.iw_beta);
                    wi = 1.0 / sqrt(s_hat);
                }
                if (p.init_weights) p.init_weights[(size_t)b * N + r] = wi;
                C::roww()[r] = wi;
            }
            __syncthreads();
            if (C::EXT && p.hybrid && hy.hybrid_wf_method == 1 && !p.hybrid_wf_in) {
                // hybrid_weight_factor_method='weight' (drt1d.py:749-759): balance the two domains by the fourth root of
                // the ratio of their weight scales, mean(est_weights^-2)^-1/2
                double t2[2] = {0.0, 0.0};
                for (int r = tid; r < N; r += C::kThreads) {
                    const double e = est_g[r];
                    if (r < c.nc) t2[0] += 1.0 / (e * e); else t2[1] += 1.0 / (e * e);
                }
                block_reduce<C, 2, 0u>(t2, c);
                const double cws = 1.0 / sqrt(t2[0] / (double)c.nc), ews = 1.0 / sqrt(t2[1] / (double)(N - c.nc));
                const double ratio = sqrt(sqrt(ews / cws));
                ewf = 1.0 / ratio;
                cwf = ratio;
            }
            it = 0;
            if (max_it <= 0) break;
            continue;
        }
        xi = qo.xi;
        fun = qo.pcost;
        hyper_block<C>(c, hd, c.ns, n - c.ns, rho, xmx, it == 0 && !cont);
        if (c.dop_a >= 0) hyper_block<C>(c, hp, c.dop_a, c.dop_b - c.dop_a, dop_rho, dop_xmx, it == 0 && !cont);
        PROF_ADD(3);
        weights_phase<C>(c, est_g, var_floor, c.vz >= 0, false);
        PROF_ADD(4);
        {   // convergence, qphb.py:597-603,969-970
            const bool act = tid < n;
            const double dx = xi - x_in;
            double t3[3] = {act ? fabs(dx / (x_in + 1e-15)) : 0.0, act ? fabs(dx) : 0.0, act ? x_in : 0.0};
            block_reduce<C, 3, 0x3u>(t3, c);
            const double atol = (t3[2] / (double)n) * 1e-3;
            conv = (t3[0] <= hy.xtol) || (t3[1] <= atol);
        }
        ++it;
        // a continuation ignores convergence before pass min_iter (drt1d.py:1355)
        if ((conv && (!cont || it >= p.pfrt_min_iter)) || it >= max_it) {
            // ---- outputs of the fit proper (before the optional calculate_pq pass rescales c.w)
            if (p.weights) for (int r = tid; r < N; r += C::kThreads) p.weights[(size_t)b * N + r] = C::roww()[r];
            if (p.resid_ss && !cont) {   // sum of squared residuals of the final x per domain (evaluate_rss / evaluate_llh)
                if (c.vz >= 0) {
                    // the reference evaluates them with the design matrix whose vz_offset column has already been
                    // rewritten from this x (drt1d.py:972-979, 4433-4496); the last weights pass used the previous column
                    const int lane = tid & 31, warp = tid >> 5;
                    const double* xs = C::vec(C::XS);
                    __syncthreads();
                    for (int r = warp; r < N; r += C::kWarps) {
                        const double* __restrict__ src = c.rm + (size_t)r * n;
                        double acc = 0.0;
                        for (int col = lane; col < n; col += 32) {
                            double xv = xs[col];
                            if (C::EXT && col >= c.dop_a && col < c.dop_b) xv *= c.dop_cs;
                            acc += ((col == c.vz) ? c.vzcol[r] : src[col]) * xv;
                        }
                        acc = warp_sum(acc);
                        if (lane == 0) {
                            const double resid = acc - (C::EXT ? c.rv[r] * c.rv_scale : c.rv[r]);
                            C::rowr2(c.N)[r] = resid * resid;
                        }
                    }
                    __syncthreads();
                }
                double t2[2] = {0.0, 0.0};
                for (int r = tid; r < N; r += C::kThreads) {
                    if (r < c.nc) t2[0] += C::rowr2(c.N)[r]; else t2[1] += C::rowr2(c.N)[r];
                }
                block_reduce<C, 2, 0u>(t2, c);
                if (tid == 0) { p.resid_ss[2 * (size_t)b] = t2[0]; p.resid_ss[2 * (size_t)b + 1] = t2[1]; }
            }
            // calculate_pq / the post-fit diagnostics belong to the plain fit (with PFRT: to its first step, whose
            // attributes the reference keeps, drt1d.py:2588-2603)
            if (cont || (p.p_matrix == nullptr && p.dist_var == nullptr)) break;
            final_pq = true;
        }
    }
    if (!cont) n_outer0 = it < 0 ? 0 : it;
    if (pfrt && !fatal) {
        // step_update of the reference (drt1d.py:2617-2650): weights re-estimated from x alone (no blend with
        // est_weights, no outlier structure), the two data terms of the marginal llh (qphb.py:1359-1373), and
        // calculate_pq under those weights with the hypers of the *first* step
        const size_t so = (size_t)b * p.n_pfrt + step;
        if (tid < n) p.pfrt_x[so * n + tid] = xi;
        if (tid == 0 && p.pfrt_iters) p.pfrt_iters[so] = it;
        const double keep_p = c.outlier_p;
        c.outlier_p = -1.0;
        weights_phase<C>(c, nullptr, var_floor, false, false);
        c.outlier_p = keep_p;
        double t2[2] = {0.0, 0.0};
        for (int r = tid; r < N; r += C::kThreads) {
            const double w = C::roww()[r];
            t2[0] += (w * w) * C::rowr2(c.N)[r];
            t2[1] += log(w);
        }
        block_reduce<C, 2, 0u>(t2, c);
        if (tid == 0) { p.pfrt_llh[2 * so] = t2[0]; p.pfrt_llh[2 * so + 1] = t2[1]; }
        if (p.pfrt_p) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                f.drt[k] = lam_init * hy.derivative_weights[k] * rho[k];
                f.dop[k] = hy.dop_l2_lambda_0 * hy.dop_derivative_weights[k] * dop_rho[k];
            }
            gram_phase<C>(c, f, false, 0.0, p.pfrt_p + so * n * n, nullptr);
        }
        __syncthreads();
        for (int r = tid; r < N; r += C::kThreads) C::roww()[r] = p.weights[(size_t)b * N + r];   // back to the fit's weights
        __syncthreads();
    }
    }   // steps

    if (tid < n) {
        p.x[(size_t)b * n + tid] = xi;
        if (p.s_vectors) {
#pragma unroll
            for (int k = 0; k < 3; ++k) p.s_vectors[((size_t)b * 3 + k) * n + tid] = C::vec(C::SV0 + k)[tid];
        }
    }
    {
        const bool act = tid < n;
        double t1[1] = {act && !isfinite(xi) ? 1.0 : 0.0};
        block_reduce<C, 1, 0x1u>(t1, c);
        if (t1[0] > 0.0 || fatal) status |= HDRT_ST_NAN;
    }
    if (conv) status |= HDRT_ST_CONVERGED;
    else if (!fatal) status |= HDRT_ST_MAXITER;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (p.rho) p.rho[(size_t)b * 3 + k] = rho[k];
            if (p.xmx_norms) p.xmx_norms[(size_t)b * 3 + k] = xmx[k];
            if (p.dop_rho) p.dop_rho[(size_t)b * 3 + k] = dop_rho[k];
            if (p.dop_xmx_norms) p.dop_xmx_norms[(size_t)b * 3 + k] = dop_xmx[k];
        }
        if (C::EXT && p.hybrid_wf_out) { p.hybrid_wf_out[2 * (size_t)b] = cwf; p.hybrid_wf_out[2 * (size_t)b + 1] = ewf; }
        if (C::EXT && p.scale_factors) {
            if (!hy.solve_rp) p.scale_factors[3 * (size_t)b] = 1.0;
            p.scale_factors[3 * (size_t)b + 1] = us_factor;
            p.scale_factors[3 * (size_t)b + 2] = c.dop_cs;
        }
        if (p.fun) p.fun[b] = fun;
        if (p.n_outer) p.n_outer[b] = pfrt ? n_outer0 : (it < 0 ? 0 : it);
        if (p.n_ipm) p.n_ipm[b] = n_ipm;
        if (p.status) p.status[b] = status;
    }
    if (fatal && p.weights) for (int r = tid; r < N; r += C::kThreads) p.weights[(size_t)b * N + r] = C::roww()[r];
    __syncthreads();
}

template <class C>
__global__ void __launch_bounds__(C::kThreads, C::MINB)
qphb_kernel(const hdrt_qphb_problem p, int* work_counter) {
    __shared__ int s_work;
    __shared__ unsigned s_tmem;
    Ctx c;
    c.N = p.n_rows; c.n = p.n_cols; c.ns = p.n_special; c.nc = p.n_chrono;
    c.dop_a = p.dop_start; c.dop_b = p.dop_end; c.vz = p.vz_index; c.vb_a = p.vb_start; c.vb_b = p.vb_end;
    c.hvec = p.h; c.l1 = p.l1; c.vz_strength = p.vz_strength;
    c.T = (p.n_cols + 7) >> 3;
    c.lane = threadIdx.x & 31;
    c.g = c.lane >> 2; c.q = c.lane & 3;
    // tensor memory for the QP matrix: one allocation per (persistent) CTA, by warp 0, released at the end
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(Tm<C>::kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int wphys = threadIdx.x >> 5;     // the TMEM lane quadrant is a property of the physical warp
        c.tm = s_tmem + ((unsigned)(32 * (wphys & 3)) << 16) + (unsigned)((wphys >> 2) * Tm<C>::kColsWarp);
    }
    // tile-grid role of this warp, rotated per block so that the heavy roles (diagonal owners) of the blocks
    // sharing an SM land on different warp schedulers
    const int role = ((threadIdx.x >> 5) + blockIdx.x % 3) % C::kWarps;
    c.role = role; c.wr = role / C::W; c.wc = role % C::W;
    c.dv = c.wc <= c.wr;
    c.red_phase = 0;
#ifdef HDRT_PROFILE
    if (threadIdx.x < 32) s_prof[threadIdx.x] = 0;
#endif
    // zero the whole vector area once: padding entries (index >= n) of the column buffers must read as zero
    for (int i = threadIdx.x; i < kNumVec * C::NV; i += C::kThreads) g_smem[i] = 0.0;
    __syncthreads();

    while (true) {
        if (threadIdx.x == 0) s_work = atomicAdd(work_counter, 1);
        __syncthreads();
        const int b = s_work;
        __syncthreads();
        if (b >= p.batch) break;
        fit_one<C>(p, b, c);
        PROF_COUNT(25);
    }
#ifdef HDRT_PROFILE
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < 32) g_prof[threadIdx.x] += s_prof[threadIdx.x];
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(Tm<C>::kCols) : "memory");
}

}  // namespace hdrt

#include "qphb_warp.cuh"
#include "resolve_kernel.cuh"

namespace hdrt {

__global__ void fp64_probe_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace hdrt

using namespace hdrt;

static long long qphb_smem_bytes(int n_rows, int n_cols, int row_vectors) {
    if (n_rows <= 0 || n_cols <= 0 || n_cols > kMaxCols) return -1;
    const long long bytes = smem_doubles(n_rows, n_cols, row_vectors) * 8;
    if (bytes > 227 * 1024) return -1;
    return bytes;
}

extern "C" long long hdrt_qphb_smem_bytes(int n_rows, int n_cols) {
    if (n_rows <= 0 || n_cols <= 0 || n_cols > kMaxCols) return -1;
    const long long bytes = smem_doubles(n_rows, n_cols) * 8;
    if (bytes > 227 * 1024) return -1;
    return bytes;
}

template <class C>
static int launch_qphb(hdrt_handle* h, const hdrt_qphb_problem& p, size_t smem, cudaStream_t st) {
    HDRT_CUDA_CHECK(cudaFuncSetAttribute(qphb_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Resident CTAs per SM from the kernel's own resources: the occupancy API answers 1 for any kernel that allocates
    // tensor memory, while the hardware co-schedules CTAs as long as their allocations fit in the 512 columns
    // (tools/probes/tmem_probe.cu)
    cudaFuncAttributes fa;
    HDRT_CUDA_CHECK(cudaFuncGetAttributes(&fa, qphb_kernel<C>));
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * C::kThreads;
    const size_t smem_per_cta = smem + fa.sharedSizeBytes + h->smem_reserved_per_cta;
    int occ = (int)(h->smem_per_sm / smem_per_cta);
    occ = occ < h->regs_per_sm / regs_per_cta ? occ : h->regs_per_sm / regs_per_cta;
    occ = occ < 512 / Tm<C>::kCols ? occ : 512 / Tm<C>::kCols;
    occ = occ < h->threads_per_sm / C::kThreads ? occ : h->threads_per_sm / C::kThreads;
    if (occ < 1) { set_error("kernel cannot be resident (smem %zu)", smem); return HDRT_ERR_UNSUPPORTED; }
    if (const char* e = getenv("HDRT_DEBUG_OCC")) { const int cap = atoi(e); if (cap > 0 && cap < occ) occ = cap; }  // dev knob
    {   // no more shared memory than the resident CTAs need: the rest of the unified array is L1, where the stack lives
        // (the supported carve-outs of sm_100 are 0, 8, 16, 32, 64, 100, 132, 164, 196 and 228 KB; the hint is a
        // percentage of the largest and is rounded to the nearest one, so aim at the smallest that fits)
        static const int kCarve[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
        const size_t need = (size_t)occ * smem_per_cta;
        int kb = 228;
        for (int i = 8; i >= 0; --i) if ((size_t)kCarve[i] * 1024 >= need) kb = kCarve[i];
        const int pct = (100 * kb + 227) / 228;
        if (!getenv("HDRT_NO_CARVE"))
        HDRT_CUDA_CHECK(cudaFuncSetAttribute(qphb_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    int grid = h->sm_count * occ;
    if (grid > p.batch) grid = p.batch;
    // per-launch work counter: fits in flight on different streams of one handle do not share it
    hdrt_launch_slot* sl;
    {
        std::lock_guard<std::mutex> lock(*static_cast<std::mutex*>(h->mu));
        sl = &h->slots[h->launches++ % kLaunchSlots];
        if (sl->used) HDRT_CUDA_CHECK(cudaStreamWaitEvent(st, sl->done, 0));
        sl->used = true;
        HDRT_CUDA_CHECK(cudaMemsetAsync(sl->work_counter, 0, sizeof(int), st));
        qphb_kernel<C><<<grid, C::kThreads, smem, st>>>(p, sl->work_counter);
        HDRT_CUDA_CHECK(cudaGetLastError());
        HDRT_CUDA_CHECK(cudaEventRecord(sl->done, st));
    }
    return HDRT_OK;
}

// Warp-per-spectrum form (qphb_warp.cuh): one CTA of four independent warps per SM.
static constexpr size_t kMaxDynSmem = 227 * 1024;
static bool warp_path_eligible(const hdrt_qphb_problem& p, bool ext) {
    if (ext || p.n_cols > wk::NV || p.vz_index >= 0) return false;
    if (const char* e = getenv("HDRT_QPHB_PATH")) { if (e[0] == 'c') return false; }      // dev knob: force the CTA kernel
    return (size_t)wk::warp_doubles(p.n_rows) * 8 * 4 + 512 <= kMaxDynSmem;
}

static int launch_qphb_warp(hdrt_handle* h, const hdrt_qphb_problem& p, cudaStream_t st) {
    const int stride = wk::warp_doubles(p.n_rows);
    const size_t smem = (size_t)stride * 8 * 4;
    HDRT_CUDA_CHECK(cudaFuncSetAttribute(wk::qphb_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HDRT_CUDA_CHECK(cudaFuncSetAttribute(wk::qphb_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int grid = h->sm_count;
    if (grid > (p.batch + 3) / 4) grid = (p.batch + 3) / 4;
    hdrt_launch_slot* sl;
    {
        std::lock_guard<std::mutex> lock(*static_cast<std::mutex*>(h->mu));
        sl = &h->slots[h->launches++ % kLaunchSlots];
        if (sl->used) HDRT_CUDA_CHECK(cudaStreamWaitEvent(st, sl->done, 0));
        sl->used = true;
        HDRT_CUDA_CHECK(cudaMemsetAsync(sl->work_counter, 0, sizeof(int), st));
        wk::qphb_warp_kernel<<<grid, 128, smem, st>>>(p, sl->work_counter, stride);
        HDRT_CUDA_CHECK(cudaGetLastError());
        HDRT_CUDA_CHECK(cudaEventRecord(sl->done, st));
    }
    return HDRT_OK;
}

extern "C" int hdrt_qphb_fit_batch(hdrt_handle* h, const hdrt_qphb_problem* prob, void* stream) {
    if (!h || !prob) { set_error("null handle or problem"); return HDRT_ERR_ARG; }
    const hdrt_qphb_problem& p = *prob;
    if (p.batch < 0 || p.n_rows <= 0 || p.n_cols <= 0 || p.n_special < 0 || p.n_special >= p.n_cols ||
        p.n_chrono < 0 || p.n_chrono > p.n_rows) {
        set_error("invalid sizes");
        return HDRT_ERR_ARG;
    }
    if (p.batch == 0) return HDRT_OK;
    if (!p.rm || !p.rv || !p.pen || !p.h || !p.l1 || !p.x || !p.est_weights) {
        set_error("rm, rv, pen, h, l1, x and est_weights are required");
        return HDRT_ERR_ARG;
    }
    if (p.n_chrono < p.n_rows && !p.vmm_eis) { set_error("vmm_eis required when EIS rows exist"); return HDRT_ERR_ARG; }
    if (p.vz_index >= 0 && (!p.vz_col || !p.vz_strength)) { set_error("vz_col and vz_strength required with vz_index"); return HDRT_ERR_ARG; }
    if (p.dop_start >= 0 && (p.dop_end <= p.dop_start || p.dop_end > p.n_special)) { set_error("invalid DOP range"); return HDRT_ERR_ARG; }
    if (p.dist_var && (!p.eval_mat || p.n_eval <= 0)) { set_error("dist_var needs eval_mat and n_eval > 0"); return HDRT_ERR_ARG; }
    if (p.n_cols > kMaxCols) { set_error("n_cols %d > %d unsupported", p.n_cols, kMaxCols); return HDRT_ERR_UNSUPPORTED; }
    if (p.hyp.has_outlier_p && (!p.outlier_t || !(p.hyp.outlier_p > 0.0) || !(p.hyp.outlier_p < 1.0))) {
        set_error("outlier_p must lie in (0, 1) and needs the outlier_t buffer");
        return HDRT_ERR_ARG;
    }
    if (p.n_pfrt > 0 && (!p.pfrt_factors || !p.pfrt_x || !p.pfrt_llh || !p.weights || p.pfrt_max_iter <= 0)) {
        set_error("PFRT needs pfrt_factors, pfrt_x, pfrt_llh, weights and pfrt_max_iter > 0");
        return HDRT_ERR_ARG;
    }
    if ((p.hyp.solve_rp || p.hyp.update_scale) && (!p.scale_factors || !(p.hyp.rp_scale > 0.0) || !(p.hyp.basis_area > 0.0))) {
        set_error("solve_rp / update_scale need scale_factors, rp_scale > 0 and basis_area > 0");
        return HDRT_ERR_ARG;
    }
    if (p.n_pfrt > 1 && p.vz_index >= 0 && !p.vz_scratch) { set_error("PFRT with a vz_offset column needs vz_scratch"); return HDRT_ERR_ARG; }
    const long long smem = qphb_smem_bytes(p.n_rows, p.n_cols, p.hyp.has_outlier_p ? 3 : 2);
    if (smem < 0) { set_error("problem %d x %d does not fit in shared memory", p.n_rows, p.n_cols); return HDRT_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    HDRT_CUDA_CHECK(cudaSetDevice(h->device));
    const bool ext = p.hyp.has_outlier_p || p.hyp.solve_rp || p.hyp.update_scale || p.n_pfrt > 0 || p.weight_factor_vec ||
                     p.hyp.init_weights_separately || p.hyp.hybrid_wf_method || p.hybrid_wf_in;
    if (p.hyp.init_weights_separately && (p.hyp.has_outlier_p || p.hyp.solve_rp)) {
        set_error("init_weights_separately cannot be combined with outlier_p / solve_rp");
        return HDRT_ERR_UNSUPPORTED;
    }
    if (warp_path_eligible(p, ext)) return launch_qphb_warp(h, p, st);
    if (small_cfg(p.n_cols)) {
        const bool three = 3 * ((size_t)smem + 64 + h->smem_reserved_per_cta) <= h->smem_per_sm;   // three CTAs per SM fit
        if (three) return ext ? launch_qphb<CfgSX>(h, p, (size_t)smem, st) : launch_qphb<CfgS>(h, p, (size_t)smem, st);
        return ext ? launch_qphb<CfgS2X>(h, p, (size_t)smem, st) : launch_qphb<CfgS2>(h, p, (size_t)smem, st);
    }
    return ext ? launch_qphb<CfgLX>(h, p, (size_t)smem, st) : launch_qphb<CfgL>(h, p, (size_t)smem, st);
}

#ifdef HDRT_PROFILE
extern "C" int hdrt_debug_profile(unsigned long long* out32, int reset) {
    if (out32) cudaMemcpyFromSymbol(out32, g_prof, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_prof, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int hdrt_probe_fp64(hdrt_handle* h, double* tflops_host, void* stream) {
    if (!h || !tflops_host) { set_error("null argument"); return HDRT_ERR_ARG; }
    HDRT_CUDA_CHECK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = h->sm_count * 8, threads = 256, iters = 20000;
    double* out = nullptr;
    HDRT_CUDA_CHECK(cudaMallocAsync(&out, sizeof(double) * blocks * threads, st));
    cudaEvent_t e0, e1;
    HDRT_CUDA_CHECK(cudaEventCreate(&e0));
    HDRT_CUDA_CHECK(cudaEventCreate(&e1));
    fp64_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
    HDRT_CUDA_CHECK(cudaEventRecord(e0, st));
    fp64_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
    HDRT_CUDA_CHECK(cudaEventRecord(e1, st));
    HDRT_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    HDRT_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *tflops_host = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    HDRT_CUDA_CHECK(cudaFreeAsync(out, st));
    return HDRT_OK;
}

// ------------------------------------------------------------------------------------------------
// cross-observation resolve (resolve_kernel.cuh)
// ------------------------------------------------------------------------------------------------
static int resolve_grid(const hdrt_handle* h, int n_windows) { return n_windows < h->sm_count ? n_windows : h->sm_count; }

extern "C" long long hdrt_resolve_work_bytes(const hdrt_handle* h, int n_windows, int nr, int nc) {
    if (!h || n_windows < 0 || nr <= 0 || nc <= 0 || nr * nc > rs::kMaxN) return -1;
    return (long long)resolve_grid(h, n_windows) * rs::work_doubles(nr * nc) * 8;
}

extern "C" int hdrt_resolve_qp_batch(hdrt_handle* h, const hdrt_resolve_problem* prob, void* work, void* stream) {
    if (!h || !prob) { set_error("null handle or problem"); return HDRT_ERR_ARG; }
    const hdrt_resolve_problem& p = *prob;
    if (p.n_windows == 0) return HDRT_OK;
    if (p.n_windows < 0 || p.nr <= 0 || p.nc <= 0 || !p.p || !p.q || !p.first_obs || !p.my || !p.param_scale || !p.h || !p.x || !work) {
        set_error("hdrt_resolve_qp_batch: invalid argument");
        return HDRT_ERR_ARG;
    }
    const int n = p.nr * p.nc;
    if (n > rs::kMaxN) { set_error("resolve window of %d unknowns > %d unsupported", n, rs::kMaxN); return HDRT_ERR_UNSUPPORTED; }
    const size_t smem = rs::smem_bytes(n);
    if (smem > 227 * 1024) { set_error("resolve window does not fit in shared memory"); return HDRT_ERR_UNSUPPORTED; }
    HDRT_CUDA_CHECK(cudaSetDevice(h->device));
    HDRT_CUDA_CHECK(cudaFuncSetAttribute(rs::resolve_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = resolve_grid(h, p.n_windows);
    rs::resolve_qp_kernel<<<grid, rs::kThreads, smem, (cudaStream_t)stream>>>(p, (double*)work, rs::work_doubles(n));
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}
