// Batched QPHB solver: one persistent CTA per spectrum in flight.
//
// Replaces, for a whole batch of spectra at once, the loop of DRT._qphb_fit_core
// (reference hybdrt/models/drt1d.py:556-1008): qphb.initialize_weights (qphb.py:1609),
// qphb.iterate_qphb (:606) = L2 assembly (:53) + weighted Gram + bound-constrained QP (:426, cvxopt
// coneqp) + closed-form s / rho updates (:320, :385) + error-structure weights (:1545), the
// convergence test (:597), xmx normalisation (drt1d.py:946) and the hybrid vz_offset column rewrite
// (drt1d.py:972).
//
// Shared memory per CTA (all FP64; offsets of everything but w / r2 / PL are compile-time constants of the
// instantiation so that loads and stores carry immediate offsets):
//   16 vectors of length NV = 16 * NBK (pdiag, x broadcast, solve rhs, ...), reduction scratch,
//   a double-buffered 4-row staging tile for the Gram pass, w[N], r2[N], then
//   PL   n x ld (ld odd): strict upper triangle = P of the current QP; lower triangle + diagonal =
//        Cholesky factor L of H = P + diag(1/d^2) of the current interior-point iteration, with each
//        32 x 32 diagonal block of L replaced by its inverse (the triangular solves are then blocked
//        matrix-vector products instead of n-step substitution chains).
// Register layout: thread (ty, tx) = (tid / 16, tid % 16) owns the entries (i, j), i >= j, i = 16a + ty,
// j = 16b + tx of the symmetric matrix being built (Gram) or factorised (Cholesky) -- NBK (NBK + 1) / 2
// doubles per thread; the rank-1 updates of both phases run out of registers and shared memory only carries
// the broadcast row / column.
// The design matrix rm, the variance-estimation matrix vmm and the penalty matrices are shared by the
// batch and stay in global memory (L2 resident, read-only path).
#include "common.cuh"

namespace hdrt {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStageRows = 4;  // per buffer; two buffers
constexpr int kMaxCols = 160;
constexpr int kRedSlots = 8;
constexpr int kNumVec = 16;

// cvxopt coneqp defaults (cvxopt 1.3 coneprog.py; the reference only sets show_progress, qphb.py:25)
constexpr double kAbsTol = 1e-7;
constexpr double kRelTol = 1e-6;
constexpr double kFeasTol = 1e-7;
constexpr int kMaxIpm = 100;
constexpr double kStep = 0.99;

extern __shared__ __align__(16) double g_smem[];

// compile-time part of the shared-memory layout (offsets in doubles)
template <int NBK>
struct SM {
    static constexpr int NV = 16 * NBK;
    static constexpr int LDA = NV + 2;  // staged row: NV columns of W*rm (zero padded) + W*rv at column NV
    static constexpr int kRed = kNumVec * NV;
    static constexpr int kRbuf = kRed + 2 * kRedSlots * kWarps;
    static constexpr int kStage = kRbuf + 16;
    static constexpr int kRows = kStage + 2 * kStageRows * LDA;
    enum { PDIAG = 0, XS, BS, DSQ, QS, RDIAG, PIV, SV0, SV1, SV2, US0, US1, US2, XH, COLA, COLB };
    static __device__ __forceinline__ double* vec(int k) { return g_smem + k * NV; }
    static __device__ __forceinline__ double* red() { return g_smem + kRed; }
    static __device__ __forceinline__ double* rbuf() { return g_smem + kRbuf; }
    static __device__ __forceinline__ double* stage() { return g_smem + kStage; }
};

__host__ __device__ inline int nbk_for(int n) { return (n + 15) / 16 <= 7 ? 7 : 10; }

__host__ __device__ inline long long smem_doubles(int N, int n) {
    const int nv = 16 * nbk_for(n);
    const long long fixed = (long long)kNumVec * nv + 2 * kRedSlots * kWarps + 16 + 2 * kStageRows * (nv + 2);
    const int ld = n | 1;
    return fixed + 2LL * ((N + 1) & ~1) + (long long)n * ld + 2;
}

struct Ctx {
    int N, n, ns, nc, dop_a, dop_b, vz, vb_a, vb_b, ld;
    const double* __restrict__ rm;
    const double* __restrict__ rv;
    const double* __restrict__ vmm_eis;
    const double* __restrict__ vmm_chrono;
    const double* __restrict__ pen;
    const double* __restrict__ hvec;
    const double* __restrict__ l1;
    const double* __restrict__ vz_strength;
    double* vzcol;  // global, per spectrum
    double *PL, *w, *r2;
    int red_phase;
};

// Reduce K per-thread values over the block; bit k of MAXMASK selects max instead of sum.  The result is
// broadcast to every thread.  One barrier: the scratch buffer alternates between two halves.
template <int NBK, int K, unsigned MAXMASK>
__device__ __forceinline__ void block_reduce(double (&v)[K], Ctx& c) {
    static_assert(K <= kRedSlots, "too many reduction slots");
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = ((MAXMASK >> k) & 1u) ? warp_max(v[k]) : warp_sum(v[k]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* red = SM<NBK>::red() + (c.red_phase & 1) * (kRedSlots * kWarps);
    c.red_phase ^= 1;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[k * kWarps + w] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = red[k * kWarps];
#pragma unroll
        for (int ww = 1; ww < kWarps; ++ww) {
            const double t = red[k * kWarps + ww];
            a = ((MAXMASK >> k) & 1u) ? fmax(a, t) : a + t;
        }
        v[k] = a;
    }
}

__host__ __device__ constexpr int tri(int a, int b) { return a * (a + 1) / 2 + b; }

// ------------------------------------------------------------------------------------------------
// Gram: P = (W rm)^T (W rm) + L2  (upper triangle + diagonal),  q = -(W rm)^T (W rv) + l1
// L2 = sum_k S_k^1/2 M~_k S_k^1/2 as in qphb.calculate_qp_l2_matrix (qphb.py:53-120)
// ------------------------------------------------------------------------------------------------
struct L2Factors {
    double drt[3];  // l2_lambda_0 * dw_k * rho_k
    double dop[3];  // dop_l2_lambda_0 * dop_dw_k * dop_rho_k
    bool use[3];    // derivative_weights[k] > 0
};

template <int NBK>
__device__ __forceinline__ double l2_entry(const Ctx& c, const L2Factors& f, int i, int j) {
    double acc = 0.0;
    const bool drt = (i >= c.ns) && (j >= c.ns);
    const bool dop = (c.dop_a >= 0) && (i >= c.dop_a) && (i < c.dop_b) && (j >= c.dop_a) && (j < c.dop_b);
    const int nn = c.n * c.n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!f.use[k]) continue;
        double m = c.pen[k * nn + i * c.n + j];
        if (drt) m *= f.drt[k];
        if (dop) m *= f.dop[k];
        const double* us = SM<NBK>::vec(SM<NBK>::US0 + k);
        acc += (us[i] * m) * us[j];
    }
    return acc;
}

// Staging: chunk = kStageRows rows; warp w loads row (w & 3), columns (w >> 2) * 32 + lane + 64 u.
template <int NBK>
struct StageRegs {
    static constexpr int U = (SM<NBK>::LDA + 63) / 64;
    double v[U];
};

template <int NBK>
__device__ __forceinline__ void stage_load(const Ctx& c, int r0, StageRegs<NBK>& s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = r0 + (warp & 3);
    const int cbase = (warp >> 2) * 32 + lane;
#pragma unroll
    for (int u = 0; u < StageRegs<NBK>::U; ++u) s.v[u] = 0.0;
    if (r < c.N) {
        const double wr = c.w[r];
        const double* __restrict__ src = c.rm + (size_t)r * c.n;
#pragma unroll
        for (int u = 0; u < StageRegs<NBK>::U; ++u) {
            const int col = cbase + 64 * u;
            double v = 0.0;
            if (col < c.n) v = ((col == c.vz) ? c.vzcol[r] : src[col]) * wr;
            else if (col == SM<NBK>::NV) v = wr * c.rv[r];
            s.v[u] = v;
        }
    }
}

template <int NBK>
__device__ __forceinline__ void stage_store(int buf, const StageRegs<NBK>& s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* dst = SM<NBK>::stage() + (buf * kStageRows + (warp & 3)) * SM<NBK>::LDA;
    const int cbase = (warp >> 2) * 32 + lane;
#pragma unroll
    for (int u = 0; u < StageRegs<NBK>::U; ++u) {
        const int col = cbase + 64 * u;
        if (col < SM<NBK>::LDA) dst[col] = s.v[u];
    }
}

template <int NBK>
__device__ __forceinline__ void gram_phase(Ctx& c, const L2Factors& f, bool l1_scalar, double l1_value, double* p_out,
                                           double* q_out) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = c.n, N = c.N;
    constexpr int LDA = SM<NBK>::LDA;
    double qacc = 0.0;
    double C[NBK * (NBK + 1) / 2];
#pragma unroll
    for (int e = 0; e < NBK * (NBK + 1) / 2; ++e) C[e] = 0.0;
    StageRegs<NBK> sr;
    stage_load<NBK>(c, 0, sr);
    __syncthreads();  // previous users of the staging area are done
    stage_store<NBK>(0, sr);
    __syncthreads();
    int buf = 0;
    for (int r0 = 0; r0 < N; r0 += kStageRows) {
        const bool more = r0 + kStageRows < N;
        if (more) stage_load<NBK>(c, r0 + kStageRows, sr);  // global loads in flight during the FMAs below
        const double* base = SM<NBK>::stage() + buf * kStageRows * LDA;
#pragma unroll
        for (int rr = 0; rr < kStageRows; ++rr) {   // rows beyond N are staged as zeros
            const double* row = base + rr * LDA;
            double ri[NBK], cj[NBK];
#pragma unroll
            for (int a = 0; a < NBK; ++a) { ri[a] = row[16 * a + ty]; cj[a] = row[16 * a + tx]; }
#pragma unroll
            for (int a = 0; a < NBK; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) C[tri(a, b)] += ri[a] * cj[b];
            if (tid < SM<NBK>::NV) qacc += row[tid] * row[SM<NBK>::NV];
        }
        if (more) stage_store<NBK>(buf ^ 1, sr);
        __syncthreads();
        buf ^= 1;
    }
    double* pdiag = SM<NBK>::vec(SM<NBK>::PDIAG);
#pragma unroll
    for (int a = 0; a < NBK; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            const int i = 16 * a + ty, j = 16 * b + tx;
            if (i < n && j <= i) {
                const double val = C[tri(a, b)] + l2_entry<NBK>(c, f, j, i);
                if (i == j) pdiag[i] = val; else c.PL[j * c.ld + i] = val;
                if (p_out) {
                    p_out[(size_t)i * n + j] = val;
                    p_out[(size_t)j * n + i] = val;
                }
            }
        }
    if (tid < n) {
        const double qv = -qacc + (l1_scalar ? l1_value : c.l1[tid]);
        SM<NBK>::vec(SM<NBK>::QS)[tid] = qv;
        if (q_out) q_out[tid] = qv;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Cholesky of H = P + diag(dsq) into the lower triangle of PL, then in-place inversion of the 32 x 32
// diagonal blocks.  rdiag[j] = 1 / L_jj.  Returns false on breakdown (uniform across the block).
// ------------------------------------------------------------------------------------------------
template <int NBK>
__device__ __forceinline__ bool factor_phase(Ctx& c) {
    using S = SM<NBK>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int n = c.n, ld = c.ld;
    double* PL = c.PL;
    double* rdiag = S::vec(S::RDIAG);
    bool ok = true;
    {
        const double* pdiag = S::vec(S::PDIAG);
        const double* dsq = S::vec(S::DSQ);
        double C[NBK * (NBK + 1) / 2];
#pragma unroll
        for (int a = 0; a < NBK; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                const int i = 16 * a + ty, j = 16 * b + tx;
                double v = 0.0;
                if (i < n && j <= i) v = (i == j) ? (pdiag[i] + dsq[i]) : PL[j * ld + i];
                C[tri(a, b)] = v;
            }
        // Right-looking, one barrier per column.  The owners of column k (tx == k % 16) publish it unscaled;
        // the owner of the pivot also publishes 1 / a_kk (NaN on breakdown).  Every thread then applies
        // C_ij -= a_ik (a_jk / a_kk) to the entries it holds.  Padding rows (i >= n) hold zeros throughout.
        double* piv = S::vec(S::PIV);
        double* rb = S::rbuf();
#pragma unroll
        for (int kb = 0; kb < NBK; ++kb) {
#pragma unroll 1
            for (int kx = 0; kx < 16; ++kx) {
                const int k = 16 * kb + kx;
                if (!ok || k >= n) break;
                double* col = S::vec(S::COLA + (k & 1));
                if (tx == kx) {
#pragma unroll
                    for (int a = kb; a < NBK; ++a) col[16 * a + ty] = C[tri(a, kb)];
                    if (ty == kx) {
                        const double pv = C[tri(kb, kb)];
                        piv[k] = pv;
                        rb[k & 1] = (pv > 0.0 && isfinite(pv)) ? 1.0 / pv : nan("");
                    }
                }
                __syncthreads();
                const double r = rb[k & 1];
                if (isnan(r)) { ok = false; break; }
                double ri[NBK], cj[NBK];
#pragma unroll
                for (int a = kb; a < NBK; ++a) { ri[a] = col[16 * a + ty]; cj[a] = col[16 * a + tx] * r; }
                if (tx <= kx) cj[kb] = 0.0;  // columns <= k of this block are final
#pragma unroll
                for (int a = kb; a < NBK; ++a)
#pragma unroll
                    for (int b = kb; b <= a; ++b) C[tri(a, b)] -= ri[a] * cj[b];
            }
        }
        __syncthreads();
        if (!ok) return false;
        if (tid < n) rdiag[tid] = 1.0 / sqrt(piv[tid]);
        __syncthreads();
#pragma unroll
        for (int a = 0; a < NBK; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                const int i = 16 * a + ty, j = 16 * b + tx;
                if (i < n && j <= i) PL[i * ld + j] = C[tri(a, b)] * rdiag[j];
            }
    }
    __syncthreads();
    // Invert the diagonal blocks in place: warp w takes block w (n <= 160 -> at most 5 blocks); lane l builds
    // column l of the inverse row by row: x_i = -(1 / L_ii) sum_{k < i} L_ik x_k, x_l = 1 / L_ll, x_{k<l} = 0.
    const int nblk = (n + 31) >> 5;
    if (warp < nblk) {
        const int r0 = 32 * warp, m = min(32, n - r0);
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            x[i] = 0.0;
            if (i < m) {
                const double* lr = PL + (r0 + i) * ld + r0;
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < i; ++k) acc[k & 3] += lr[k] * x[k];
                const double rd = rdiag[r0 + i];
                const double dot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
                x[i] = (i == lane) ? rd : ((i > lane) ? -rd * dot : 0.0);
            }
        }
        __syncwarp();
        if (lane < m) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i >= lane && i < m) PL[(r0 + i) * ld + r0 + lane] = x[i];
        }
    }
    __syncthreads();
    return true;
}

// Solve L L^T u = bs in place with the block-inverted factor.
template <int NBK>
__device__ __forceinline__ void solve_phase(Ctx& c) {
    using S = SM<NBK>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = c.n, ld = c.ld;
    const double* PL = c.PL;
    double* bs = S::vec(S::BS);
    const int nblk = (n + 31) >> 5;
    __syncthreads();
    // forward: y = L^-1 b
    for (int blk = 0; blk < nblk; ++blk) {
        const int r0 = 32 * blk, m = min(32, n - r0);
        if (warp == 0) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (lane < m) {
                const double* xr = PL + (r0 + lane) * ld + r0;
                const double* bb = bs + r0;
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    if (k <= lane) a0 += xr[k] * bb[k];
                    if (k + 1 <= lane) a1 += xr[k + 1] * bb[k + 1];
                    if (k + 2 <= lane) a2 += xr[k + 2] * bb[k + 2];
                    if (k + 3 <= lane) a3 += xr[k + 3] * bb[k + 3];
                }
            }
            __syncwarp();
            if (lane < m) bs[r0 + lane] = (a0 + a1) + (a2 + a3);
        }
        if (r0 + m >= n) break;
        __syncthreads();
        const int i = r0 + 32 + tid;
        if (i < n) {
            const double* lr = PL + i * ld + r0;
            const double* bb = bs + r0;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                a0 += lr[k] * bb[k];
                a1 += lr[k + 1] * bb[k + 1];
                a2 += lr[k + 2] * bb[k + 2];
                a3 += lr[k + 3] * bb[k + 3];
            }
            bs[i] -= (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
    __syncthreads();
    // backward: x = L^-T y
    for (int blk = nblk - 1; blk >= 0; --blk) {
        const int r0 = 32 * blk, m = min(32, n - r0);
        if (warp == 0) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (lane < m) {
                const double* xc = PL + r0 * ld + r0 + lane;
                const double* bb = bs + r0;
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    if (k >= lane && k < m) a0 += xc[k * ld] * bb[k];
                    if (k + 1 >= lane && k + 1 < m) a1 += xc[(k + 1) * ld] * bb[k + 1];
                    if (k + 2 >= lane && k + 2 < m) a2 += xc[(k + 2) * ld] * bb[k + 2];
                    if (k + 3 >= lane && k + 3 < m) a3 += xc[(k + 3) * ld] * bb[k + 3];
                }
            }
            __syncwarp();
            if (lane < m) bs[r0 + lane] = (a0 + a1) + (a2 + a3);
        }
        if (blk == 0) break;
        __syncthreads();
        if (tid < r0) {
            const double* lc = PL + r0 * ld + tid;
            const double* bb = bs + r0;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (m == 32) {
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    a0 += lc[k * ld] * bb[k];
                    a1 += lc[(k + 1) * ld] * bb[k + 1];
                    a2 += lc[(k + 2) * ld] * bb[k + 2];
                    a3 += lc[(k + 3) * ld] * bb[k + 3];
                }
            } else {
                for (int k = 0; k < m; ++k) a0 += lc[k * ld] * bb[k];
            }
            bs[tid] -= (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
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

template <int NBK>
__device__ __forceinline__ QpOut qp_phase(Ctx& c) {
    using S = SM<NBK>;
    const int tid = threadIdx.x;
    const int n = c.n, ld = c.ld;
    const bool act = tid < n;
    double* xs = S::vec(S::XS);
    double* bs = S::vec(S::BS);
    const double qi = act ? S::vec(S::QS)[tid] : 0.0;
    const double hi = act ? c.hvec[tid] : 0.0;
    QpOut out;
    out.xi = 0.0; out.pcost = 0.0; out.iters = 0; out.status = 0; out.fatal = false;

    double resx0, resz0;
    {
        double t2[2] = {qi * qi, hi * hi};
        block_reduce<NBK, 2, 0u>(t2, c);
        resx0 = fmax(1.0, sqrt(t2[0]));
        resz0 = fmax(1.0, sqrt(t2[1]));
    }
    double xi = 0.0, si = 1.0, zi = 1.0, di = 1.0, dinv = 1.0, lam = 1.0;
    double rxi = 0.0, rzi = 0.0, gap = 0.0, pcost = 0.0;
    int iters;
    // iters == -1 is the initial point (W = I); 0.. are the interior-point iterations
#pragma unroll 1
    for (iters = -1; iters <= kMaxIpm; ++iters) {
        if (iters >= 0) {
            if (act) xs[tid] = xi;
            __syncthreads();
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (act) {
                const double* colp = c.PL + tid;       // P[j][tid], j < tid  (upper triangle, column tid)
                int j = 0;
                for (; j + 3 < tid; j += 4) {
                    a0 += colp[j * ld] * xs[j];
                    a1 += colp[(j + 1) * ld] * xs[j + 1];
                    a2 += colp[(j + 2) * ld] * xs[j + 2];
                    a3 += colp[(j + 3) * ld] * xs[j + 3];
                }
                for (; j < tid; ++j) a0 += colp[j * ld] * xs[j];
                a1 += S::vec(S::PDIAG)[tid] * xi;
                const double* rowp = c.PL + tid * ld;  // P[tid][j], j > tid
                j = tid + 1;
                for (; j + 3 < n; j += 4) {
                    a0 += rowp[j] * xs[j];
                    a1 += rowp[j + 1] * xs[j + 1];
                    a2 += rowp[j + 2] * xs[j + 2];
                    a3 += rowp[j + 3] * xs[j + 3];
                }
                for (; j < n; ++j) a0 += rowp[j] * xs[j];
            }
            rxi = ((a0 + a1) + (a2 + a3)) + qi;
            const double f0p = act ? (xi * rxi + xi * qi) : 0.0;
            rxi -= zi;
            rzi = si - hi - xi;
            double t5[5] = {f0p, act ? rxi * rxi : 0.0, act ? rzi * rzi : 0.0, act ? zi * rzi : 0.0,
                            act ? (iters == 0 ? si * zi : lam * lam) : 0.0};
            block_reduce<NBK, 5, 0u>(t5, c);
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
        if (act) S::vec(S::DSQ)[tid] = dinv * dinv;
        __syncthreads();
        if (!factor_phase<NBK>(c)) {
            out.status |= HDRT_ST_KKT_FAIL;
            if (iters <= 0) { out.fatal = true; xi = nan(""); }
            break;
        }
        if (iters < 0) {
            // solve [P+I] x = -q - h ; z = -x - h ; s = -z, shifted into the cone
            if (act) bs[tid] = -qi - hi;
            solve_phase<NBK>(c);
            xi = act ? bs[tid] : 0.0;
            zi = -xi - hi;
            si = -zi;
            double t4[4] = {act ? si * si : 0.0, act ? -si : -INFINITY, act ? zi * zi : 0.0, act ? -zi : -INFINITY};
            block_reduce<NBK, 4, 0xAu>(t4, c);
            const double nrms = sqrt(t4[0]), ts = t4[1], nrmz = sqrt(t4[2]), tz = t4[3];
            if (ts >= -1e-8 * fmax(nrms, 1.0)) si += 1.0 + ts;
            if (tz >= -1e-8 * fmax(nrmz, 1.0)) zi += 1.0 + tz;
            continue;
        }
        const double lamsq = lam * lam;
        const double mu = gap / (double)n;
        double sigma = 0.0, step = 1.0;
        double ws3 = 0.0, dxi = 0.0, dsi = 0.0, dzi = 0.0;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            dsi = 0.0;
            if (pass == 1) dsi -= ws3;
            dsi -= lamsq;
            dsi += sigma * mu;
            dxi = -rxi;
            dzi = -rzi;
            dsi = dsi / lam;
            dzi = dzi - di * dsi;
            const double zs = dinv * dzi;
            if (act) bs[tid] = dxi - dinv * zs;
            solve_phase<NBK>(c);
            dxi = act ? bs[tid] : 0.0;
            dzi = -dinv * dxi - zs;
            dsi = dsi - dzi;
            const double prod = dsi * dzi;
            if (pass == 0) ws3 = prod;
            dsi = dsi / lam;
            dzi = dzi / lam;
            double t3[3] = {act ? prod : 0.0, act ? -dsi : -INFINITY, act ? -dzi : -INFINITY};
            block_reduce<NBK, 3, 0x6u>(t3, c);
            const double t = fmax(0.0, fmax(t3[1], t3[2]));
            if (t == 0.0) step = 1.0;
            else if (pass == 0) step = fmin(1.0, 1.0 / t);
            else step = fmin(1.0, kStep / t);
            if (pass == 0) {
                const double sg = fmin(1.0, fmax(0.0, 1.0 - step + t3[0] / gap * (step * step)));
                sigma = sg * sg * sg;
            }
        }
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
    return out;
}

// ------------------------------------------------------------------------------------------------
// Hyper-parameter updates for one coefficient block (DRT or DOP): qphb.solve_s / solve_rho
// ------------------------------------------------------------------------------------------------
struct BlockHyp {
    double dw[3], sigma[3], s_alpha[3], s_0[3], rho_alpha[3], rho_0[3];
    bool use_gmat;  // DRT block: k = 0 gets G = Xh M1 Xh (qphb.py:769-772); DOP block: 0 (drt1d.py quirk)
};

template <int NBK>
__device__ __forceinline__ void hyper_block(Ctx& c, const BlockHyp& hp, int start, int len, double* rho, double* xmx,
                                            bool first_iter) {
    using S = SM<NBK>;
    const int tid = threadIdx.x;
    const int n = c.n, nn = c.n * c.n;
    const bool act = tid < len;
    const int gi = start + tid;
    const double* xs = S::vec(S::XS);
    double* xh = S::vec(S::XH);
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
        for (int j = 0; j < len; ++j) {
            const int gj = start + j;
            const double xj = xs[gj];
            const double m0 = pcol[j * n], m1 = pcol[nn + j * n], m2 = pcol[2 * nn + j * n];
            double gam[3] = {(xi * m0) * xj, (xi * m1) * xj, (xi * m2) * xj};
            if (hp.use_gmat) gam[0] += ((xhi * m1) * xh[gj]) * inv2s0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (j == tid) {
                    gd[k] = gam[k] + (hp.s_alpha[k] - 1.0) / hp.s_0[k];
                } else {
                    const double g = gam[k] * S::vec(S::US0 + k)[gj];
                    bsum[k] += g;
                    mx[k] = fmax(mx[k], fabs(g));
                }
            }
        }
    }
    block_reduce<NBK, 3, 0x7u>(mx, c);
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
            S::vec(S::SV0 + k)[gi] = s_hat;
        }
    }
    __syncthreads();
    if (act) {
#pragma unroll
        for (int k = 0; k < 3; ++k) S::vec(S::US0 + k)[gi] = sqrt(S::vec(S::SV0 + k)[gi]);
    }
    __syncthreads();
    // rho: alpha / (x' S^1/2 M S^1/2 x / xmx + beta)
    double tr[3] = {0, 0, 0}, tx[3] = {0, 0, 0};
    if (act) {
        const double* __restrict__ pcol = c.pen + (start * n + gi);
        for (int j = 0; j < len; ++j) {
            const int gj = start + j;
            const double xj = xs[gj];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double m = pcol[k * nn + j * n];
                tr[k] += (xj * S::vec(S::US0 + k)[gj]) * m;
                tx[k] += xj * m;
            }
        }
    }
    double t6[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t6[k] = act ? (tr[k] * S::vec(S::US0 + k)[gi]) * xi : 0.0;
        t6[3 + k] = act ? tx[k] * xi : 0.0;
    }
    block_reduce<NBK, 6, 0u>(t6, c);
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
}

// ------------------------------------------------------------------------------------------------
// Error-structure weights (qphb.estimate_weights, qphb.py:1545-1594) + vz_offset column rewrite
// ------------------------------------------------------------------------------------------------
template <int NBK>
__device__ __forceinline__ void weights_phase(Ctx& c, const double* est, double var_floor, bool update_vz) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = c.N, n = c.n, nc = c.nc;
    const double* xs = SM<NBK>::vec(SM<NBK>::XS);
    for (int r = warp; r < N; r += kWarps) {
        double acc = 0.0, accv = 0.0;
        const double* __restrict__ src = c.rm + (size_t)r * n;
        for (int col = lane; col < n; col += 32) {
            const double xv = xs[col];
            if (col == c.vz) {
                acc += c.vzcol[r] * xv;
            } else {
                const double t = src[col] * xv;
                acc += t;
                if (col < c.vb_a || col >= c.vb_b) accv += t;
            }
        }
        acc = warp_sum(acc);
        accv = warp_sum(accv);
        if (lane == 0) {
            const double resid = acc - c.rv[r];
            c.r2[r] = resid * resid;
            if (update_vz) {
                const double sep = (r < nc) ? accv : -accv;
                c.vzcol[r] = sep * c.vz_strength[r];
            }
        }
    }
    __syncthreads();
    double chrono_mean = 0.0;
    if (nc > 0 && c.vmm_chrono == nullptr) {
        double t1[1] = {0.0};
        for (int r = tid; r < nc; r += kThreads) t1[0] += c.r2[r];
        block_reduce<NBK, 1, 0u>(t1, c);
        chrono_mean = t1[0] / (double)nc;
    }
    for (int r = warp; r < N; r += kWarps) {
        double s_hat;
        if (r < nc) {
            if (c.vmm_chrono == nullptr) {
                s_hat = chrono_mean;
            } else {
                double acc = 0.0;
                for (int col = lane; col < nc; col += 32) acc += c.vmm_chrono[(size_t)r * nc + col] * c.r2[col];
                s_hat = warp_sum(acc);
            }
        } else {
            const int ne = N - nc;
            double acc = 0.0;
            const double* __restrict__ vr = c.vmm_eis + (size_t)(r - nc) * ne;
            for (int col = lane; col < ne; col += 32) acc += vr[col] * c.r2[nc + col];
            s_hat = warp_sum(acc);
        }
        if (lane == 0) {
            if (s_hat < var_floor) s_hat = var_floor;
            double w = 1.0 / sqrt(s_hat);
            if (est != nullptr) {
                const double e = est[r];
                const double frac = w / (w + e);
                w = frac * w + (1.0 - frac) * e;
            }
            c.w[r] = fmax(w, 1e-10);
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// One spectrum.  The outer loop runs phase -1 (initialize_weights), 0..max_iter-1 (iterate_qphb) and,
// when P/q are requested, one final Gram-only phase (calculate_pq) through the same code.
// ------------------------------------------------------------------------------------------------
template <int NBK>
__device__ __forceinline__ void fit_one(const hdrt_qphb_problem& p, int b, Ctx& c) {
    using S = SM<NBK>;
    const int tid = threadIdx.x;
    const int N = p.n_rows, n = p.n_cols;
    const hdrt_hypers& hy = p.hyp;
    c.rm = p.rm + (size_t)b * p.rm_stride;
    c.rv = p.rv + (size_t)b * N;
    c.vmm_eis = p.vmm_eis ? p.vmm_eis + (size_t)b * p.vmm_eis_stride : nullptr;
    c.vmm_chrono = p.vmm_chrono ? p.vmm_chrono + (size_t)b * p.vmm_chrono_stride : nullptr;
    c.pen = p.pen + (size_t)b * p.pen_stride;
    c.vzcol = p.vz_col ? p.vz_col + (size_t)b * N : nullptr;
    double* est_g = p.est_weights + (size_t)b * N;

    // var floor = var(y) * 1e-7 (qphb.py:1560-1561)
    double var_floor;
    {
        double t1[1] = {0.0};
        for (int r = tid; r < N; r += kThreads) t1[0] += c.rv[r];
        block_reduce<NBK, 1, 0u>(t1, c);
        const double mean = t1[0] / (double)N;
        double t2[1] = {0.0};
        for (int r = tid; r < N; r += kThreads) { const double d = c.rv[r] - mean; t2[0] += d * d; }
        block_reduce<NBK, 1, 0u>(t2, c);
        var_floor = (t2[0] / (double)N) * 1e-7;
    }

    double rho[3], dop_rho[3], xmx[3] = {1, 1, 1}, dop_xmx[3] = {1, 1, 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) { rho[k] = hy.rho_0[k]; dop_rho[k] = hy.dop_rho_0[k]; }
    if (tid < S::NV) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { S::vec(S::SV0 + k)[tid] = hy.s_0[k]; S::vec(S::US0 + k)[tid] = sqrt(hy.s_0[k]); }
    }
    for (int r = tid; r < N; r += kThreads) {
        c.w[r] = 1.0;
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

    int status = 0, n_ipm = 0;
    L2Factors f;
#pragma unroll
    for (int k = 0; k < 3; ++k) f.use[k] = hy.derivative_weights[k] > 0.0;
    double xi = 1e-6;  // drt1d.py:612
    double fun = 0.0;
    int it = -1;       // -1: initialize_weights (drt1d.py:640-675, qphb.py:1609-1681)
    bool conv = false, fatal = false, final_pq = false;
#pragma unroll 1
    while (true) {
        const bool init = it < 0;
        const double x_in = xi;
        // weights entering the Gram: 1 (init) / weight factors (drt1d.py:881-892) / scaled weights (:991-1008)
        if (!init) {
            for (int r = tid; r < N; r += kThreads) {
                double w = c.w[r];
                if (final_pq) {
                    w *= hy.weight_factor;
                    if (p.hybrid) w *= (r < c.nc) ? hy.chrono_weight_factor : hy.eis_weight_factor;
                } else {
                    if (p.hybrid) w *= (r < c.nc) ? hy.chrono_weight_factor : hy.eis_weight_factor;
                    if (it > 0) w = w * hy.weight_factor;
                }
                c.w[r] = w;
            }
        }
        {
            const double lam0 = init ? hy.iw_l2_lambda_0 : hy.l2_lambda_0;
            const double dlam0 = init ? hy.dop_l2_lambda_0 * (hy.iw_l2_lambda_0 / hy.l2_lambda_0) : hy.dop_l2_lambda_0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                f.drt[k] = lam0 * hy.derivative_weights[k] * rho[k];
                f.dop[k] = dlam0 * hy.dop_derivative_weights[k] * dop_rho[k];
            }
        }
        __syncthreads();
        gram_phase<NBK>(c, f, init, hy.iw_l1_lambda_0, final_pq ? p.p_matrix + (size_t)b * n * n : nullptr,
                        (final_pq && p.q_vector) ? p.q_vector + (size_t)b * n : nullptr);
        if (final_pq) break;

        QpOut qo = qp_phase<NBK>(c);
        status |= qo.status;
        n_ipm += qo.iters;
        if (qo.fatal) { fatal = true; xi = qo.xi; break; }
        if (tid < n) S::vec(S::XS)[tid] = qo.xi;
        __syncthreads();
        if (init) {
            if (tid < n && p.x_overfit) p.x_overfit[(size_t)b * n + tid] = qo.xi;
            weights_phase<NBK>(c, nullptr, var_floor, false);
            for (int r = tid; r < N; r += kThreads) {
                const double e = c.w[r];
                est_g[r] = e;
                double wi = e;
                if (hy.has_iw_prior) {  // qphb.solve_init_weight_scale, qphb.py:1471-1479
                    const double bq = 0.5 - hy.iw_alpha + 1.0;
                    const double s_hat = (-bq + sqrt(bq * bq + 2.0 * hy.iw_beta / (e * e))) / (2.0 * hy.iw_beta);
                    wi = 1.0 / sqrt(s_hat);
                }
                if (p.init_weights) p.init_weights[(size_t)b * N + r] = wi;
                c.w[r] = wi;
            }
            __syncthreads();
            it = 0;
            if (hy.max_iter <= 0) break;
            continue;
        }
        xi = qo.xi;
        fun = qo.pcost;
        hyper_block<NBK>(c, hd, c.ns, n - c.ns, rho, xmx, it == 0);
        if (c.dop_a >= 0) hyper_block<NBK>(c, hp, c.dop_a, c.dop_b - c.dop_a, dop_rho, dop_xmx, it == 0);
        weights_phase<NBK>(c, est_g, var_floor, c.vz >= 0);
        {   // convergence, qphb.py:597-603,969-970
            const bool act = tid < n;
            const double dx = xi - x_in;
            double t3[3] = {act ? fabs(dx / (x_in + 1e-15)) : 0.0, act ? fabs(dx) : 0.0, act ? x_in : 0.0};
            block_reduce<NBK, 3, 0x3u>(t3, c);
            const double atol = (t3[2] / (double)n) * 1e-3;
            conv = (t3[0] <= hy.xtol) || (t3[1] <= atol);
        }
        ++it;
        if (conv || it >= hy.max_iter) {
            // ---- outputs of the fit proper (before the optional calculate_pq pass rescales c.w)
            if (p.weights) for (int r = tid; r < N; r += kThreads) p.weights[(size_t)b * N + r] = c.w[r];
            if (p.p_matrix == nullptr) break;
            final_pq = true;
        }
    }

    if (tid < n) {
        p.x[(size_t)b * n + tid] = xi;
        if (p.s_vectors) {
#pragma unroll
            for (int k = 0; k < 3; ++k) p.s_vectors[((size_t)b * 3 + k) * n + tid] = S::vec(S::SV0 + k)[tid];
        }
    }
    {
        const bool act = tid < n;
        double t1[1] = {act && !isfinite(xi) ? 1.0 : 0.0};
        block_reduce<NBK, 1, 0x1u>(t1, c);
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
        if (p.fun) p.fun[b] = fun;
        if (p.n_outer) p.n_outer[b] = it < 0 ? 0 : it;
        if (p.n_ipm) p.n_ipm[b] = n_ipm;
        if (p.status) p.status[b] = status;
    }
    if (fatal && p.weights) for (int r = tid; r < N; r += kThreads) p.weights[(size_t)b * N + r] = c.w[r];
    __syncthreads();
}

template <int NBK>
__global__ void __launch_bounds__(kThreads, NBK <= 7 ? 2 : 1)
qphb_kernel(const hdrt_qphb_problem p, int* work_counter) {
    __shared__ int s_work;
    Ctx c;
    c.N = p.n_rows; c.n = p.n_cols; c.ns = p.n_special; c.nc = p.n_chrono;
    c.dop_a = p.dop_start; c.dop_b = p.dop_end; c.vz = p.vz_index; c.vb_a = p.vb_start; c.vb_b = p.vb_end;
    c.hvec = p.h; c.l1 = p.l1; c.vz_strength = p.vz_strength;
    c.ld = p.n_cols | 1;
    const int npad = (p.n_rows + 1) & ~1;
    c.w = g_smem + SM<NBK>::kRows;
    c.r2 = c.w + npad;
    c.PL = c.r2 + npad;
    c.red_phase = 0;
    // zero the whole vector area once: padding entries (index >= n) of the column buffers must read as zero
    for (int i = threadIdx.x; i < kNumVec * SM<NBK>::NV; i += kThreads) g_smem[i] = 0.0;
    __syncthreads();

    while (true) {
        if (threadIdx.x == 0) s_work = atomicAdd(work_counter, 1);
        __syncthreads();
        const int b = s_work;
        __syncthreads();
        if (b >= p.batch) break;
        fit_one<NBK>(p, b, c);
    }
}

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

extern "C" long long hdrt_qphb_smem_bytes(int n_rows, int n_cols) {
    if (n_rows <= 0 || n_cols <= 0 || n_cols > kMaxCols) return -1;
    const long long bytes = smem_doubles(n_rows, n_cols) * 8;
    if (bytes > 227 * 1024) return -1;
    return bytes;
}

template <int NBK>
static int launch_qphb(hdrt_handle* h, const hdrt_qphb_problem& p, size_t smem, cudaStream_t st) {
    HDRT_CUDA_CHECK(cudaFuncSetAttribute(qphb_kernel<NBK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    HDRT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qphb_kernel<NBK>, kThreads, smem));
    if (occ < 1) { set_error("kernel cannot be resident (smem %zu)", smem); return HDRT_ERR_UNSUPPORTED; }
    int grid = h->sm_count * occ;
    if (grid > p.batch) grid = p.batch;
    HDRT_CUDA_CHECK(cudaMemsetAsync(h->work_counter, 0, sizeof(int), st));
    qphb_kernel<NBK><<<grid, kThreads, smem, st>>>(p, h->work_counter);
    HDRT_CUDA_CHECK(cudaGetLastError());
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
    if (p.n_cols > kMaxCols) { set_error("n_cols %d > %d unsupported", p.n_cols, kMaxCols); return HDRT_ERR_UNSUPPORTED; }
    const long long smem = hdrt_qphb_smem_bytes(p.n_rows, p.n_cols);
    if (smem < 0) { set_error("problem %d x %d does not fit in shared memory", p.n_rows, p.n_cols); return HDRT_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    HDRT_CUDA_CHECK(cudaSetDevice(h->device));
    if (nbk_for(p.n_cols) == 7) return launch_qphb<7>(h, p, (size_t)smem, st);
    return launch_qphb<10>(h, p, (size_t)smem, st);
}

extern "C" int hdrt_probe_fp64(hdrt_handle* h, double* tflops_host) {
    if (!h || !tflops_host) { set_error("null argument"); return HDRT_ERR_ARG; }
    HDRT_CUDA_CHECK(cudaSetDevice(h->device));
    const int blocks = h->sm_count * 8, threads = 256, iters = 20000;
    double* out = nullptr;
    HDRT_CUDA_CHECK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    HDRT_CUDA_CHECK(cudaEventCreate(&e0));
    HDRT_CUDA_CHECK(cudaEventCreate(&e1));
    fp64_probe_kernel<<<blocks, threads>>>(out, iters);
    HDRT_CUDA_CHECK(cudaEventRecord(e0));
    fp64_probe_kernel<<<blocks, threads>>>(out, iters);
    HDRT_CUDA_CHECK(cudaEventRecord(e1));
    HDRT_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    HDRT_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *tflops_host = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return HDRT_OK;
}
