// Batched QPHB solver: one persistent CTA per spectrum in flight.
//
// Replaces, for a whole batch of spectra at once, the loop of DRT._qphb_fit_core
// (reference hybdrt/models/drt1d.py:556-1008): qphb.initialize_weights (qphb.py:1609),
// qphb.iterate_qphb (:606) = L2 assembly (:53) + weighted Gram + bound-constrained QP (:426, cvxopt
// coneqp) + closed-form s / rho updates (:320, :385) + error-structure weights (:1545), the
// convergence test (:597), xmx normalisation (drt1d.py:946) and the hybrid vz_offset column rewrite
// (drt1d.py:972).
//
// Data layout per CTA (dynamic shared memory, all FP64):
//   PL   n x ld (ld odd): strict upper triangle = P of the current QP; lower triangle + diagonal =
//        Cholesky factor L of H = P + diag(1/d^2) of the current interior-point iteration, with each
//        32 x 32 diagonal block of L replaced by its inverse (the triangular solves are then blocked
//        matrix-vector products instead of n-step substitution chains)
//   18 vectors of length nv (pdiag, x broadcast, solve rhs, ...), 2 vectors of length N (w, r^2),
//   a kStageRows x ldA staging tile for the Gram pass.
// Register layout (kernels instantiated with NBK = ceil(n/16) <= 10): thread (ty, tx) = (tid/16, tid%16)
// owns the entries (i, j), i >= j, with i = 16a + ty, j = 16b + tx of the symmetric matrix being built
// (Gram) or factorised (Cholesky) -- NBK(NBK+1)/2 doubles per thread, so the rank-1 updates of both
// phases run out of registers and shared memory only carries the broadcast row / column.
// The design matrix rm, the variance-estimation matrix vmm and the penalty matrices are shared by the
// batch and stay in global memory (L2 resident, read-only path).
#include "common.cuh"

namespace hdrt {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStageRows = 8;
constexpr int kMaxCols = 256;
constexpr int kRedSlots = 8;
constexpr int kNumVec = 18;

// cvxopt coneqp defaults (cvxopt 1.3 coneprog.py; the reference only sets show_progress, qphb.py:25)
constexpr double kAbsTol = 1e-7;
constexpr double kRelTol = 1e-6;
constexpr double kFeasTol = 1e-7;
constexpr int kMaxIpm = 100;
constexpr double kStep = 0.99;

struct SmemLayout {
    int ld;   // leading dimension of PL (odd)
    int nv;   // padded vector length (16 * NBK of the kernel that runs this shape)
    int ldA;  // staging leading dimension
    size_t pl, vec, rows, stage, red, total;  // offsets in doubles
};

__host__ __device__ inline SmemLayout make_layout(int N, int n) {
    SmemLayout L;
    L.ld = n | 1;
    const int nbk = (n + 15) / 16;
    L.nv = nbk <= 7 ? 112 : (nbk <= 10 ? 160 : 16 * nbk);  // = 16 * NBK of the kernel instantiation
    L.ldA = L.nv + 2;  // column nv holds w*y
    L.pl = 0;
    L.vec = L.pl + (size_t)n * L.ld + (((size_t)n * L.ld) & 1);
    L.rows = L.vec + (size_t)kNumVec * L.nv;
    L.stage = L.rows + (size_t)2 * ((N + 1) & ~1);
    L.red = L.stage + (size_t)kStageRows * L.ldA;
    L.total = L.red + (size_t)2 * kRedSlots * kWarps;
    return L;
}

struct Ctx {
    // problem
    int N, n, ns, nc, dop_a, dop_b, vz, vb_a, vb_b;
    const double* __restrict__ rm;
    const double* __restrict__ rv;
    const double* __restrict__ vmm_eis;
    const double* __restrict__ vmm_chrono;
    const double* __restrict__ pen;
    const double* __restrict__ hvec;
    const double* __restrict__ l1;
    const double* __restrict__ vz_strength;
    double* vzcol;  // global, per spectrum
    // shared memory
    int ld, nv, ldA;
    double *PL, *pdiag, *xs, *bs, *dsq, *qs, *rdiag, *piv, *sv[3], *us[3], *xh, *colA, *colB, *w, *r2, *stage, *red;
    int red_phase;
};

// Reduce K per-thread values over the block; bit k of MAXMASK selects max instead of sum.  The result is
// broadcast to every thread.  One barrier: the scratch buffer alternates between two halves.
template <int K, unsigned MAXMASK>
__device__ __forceinline__ void block_reduce(double (&v)[K], Ctx& c) {
    static_assert(K <= kRedSlots, "too many reduction slots");
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = ((MAXMASK >> k) & 1u) ? warp_max(v[k]) : warp_sum(v[k]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* red = c.red + (c.red_phase & 1) * (kRedSlots * kWarps);
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
        acc += (c.us[k][i] * m) * c.us[k][j];
    }
    return acc;
}

// stage rows [r0, r0+rows) of W*rm (and W*rv in column nv) into shared memory: one warp per row
__device__ __forceinline__ void stage_rows(const Ctx& c, int r0, int rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < rows) {
        const int r = r0 + warp;
        const double wr = c.w[r];
        const double* __restrict__ src = c.rm + (size_t)r * c.n;
        double* dst = c.stage + warp * c.ldA;
        for (int col = lane; col < c.ldA; col += 32) {
            double v = 0.0;
            if (col < c.n) v = ((col == c.vz) ? c.vzcol[r] : src[col]) * wr;
            else if (col == c.nv) v = wr * c.rv[r];
            dst[col] = v;
        }
    }
}

__device__ __forceinline__ void store_p_entry(const Ctx& c, const L2Factors& f, int i, int j, double acc, double* p_out) {
    // i >= j
    const double val = acc + l2_entry(c, f, j, i);
    if (i == j) c.pdiag[i] = val; else c.PL[j * c.ld + i] = val;
    if (p_out) {
        p_out[(size_t)i * c.n + j] = val;
        p_out[(size_t)j * c.n + i] = val;
    }
}

template <int NBK>
__device__ __forceinline__ void gram_phase(Ctx& c, const L2Factors& f, bool l1_scalar, double l1_value, double* p_out,
                                           double* q_out) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = c.n, N = c.N, ldA = c.ldA;
    double qacc = 0.0;
    if constexpr (NBK > 0) {
        double C[NBK * (NBK + 1) / 2];
#pragma unroll
        for (int e = 0; e < NBK * (NBK + 1) / 2; ++e) C[e] = 0.0;
        for (int r0 = 0; r0 < N; r0 += kStageRows) {
            const int rows = min(kStageRows, N - r0);
            __syncthreads();
            stage_rows(c, r0, rows);
            __syncthreads();
            for (int rr = 0; rr < rows; ++rr) {
                const double* row = c.stage + rr * ldA;
                double ri[NBK], cj[NBK];
#pragma unroll
                for (int a = 0; a < NBK; ++a) { ri[a] = row[16 * a + ty]; cj[a] = row[16 * a + tx]; }
#pragma unroll
                for (int a = 0; a < NBK; ++a)
#pragma unroll
                    for (int b = 0; b <= a; ++b) C[tri(a, b)] += ri[a] * cj[b];
                if (tid < n) qacc += row[tid] * row[c.nv];
            }
        }
#pragma unroll
        for (int a = 0; a < NBK; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                const int i = 16 * a + ty, j = 16 * b + tx;
                if (i < n && j <= i) store_p_entry(c, f, i, j, C[tri(a, b)], p_out);
            }
    } else {
        // generic path (n > 160): one output entry at a time per thread, rows staged the same way
        const int total = n * (n + 1) / 2;
        for (int e0 = 0; e0 < total; e0 += kThreads) {
            const int e = e0 + tid;
            int i = 0, j = 0;
            if (e < total) {
                i = (int)floor((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while (i * (i + 1) / 2 > e) --i;
                while ((i + 1) * (i + 2) / 2 <= e) ++i;
                j = e - i * (i + 1) / 2;
            }
            double acc = 0.0;
            for (int r0 = 0; r0 < N; r0 += kStageRows) {
                const int rows = min(kStageRows, N - r0);
                __syncthreads();
                stage_rows(c, r0, rows);
                __syncthreads();
                for (int rr = 0; rr < rows; ++rr) {
                    const double* row = c.stage + rr * ldA;
                    acc += row[i] * row[j];
                    if (e0 == 0 && tid < n) qacc += row[tid] * row[c.nv];
                }
            }
            if (e < total) store_p_entry(c, f, i, j, acc, p_out);
        }
    }
    if (tid < n) {
        const double qv = -qacc + (l1_scalar ? l1_value : c.l1[tid]);
        c.qs[tid] = qv;
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
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = c.n, ld = c.ld;
    double* PL = c.PL;
    bool ok = true;
    if constexpr (NBK > 0) {
        const int tx = tid & 15, ty = tid >> 4;
        double C[NBK * (NBK + 1) / 2];
#pragma unroll
        for (int a = 0; a < NBK; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                const int i = 16 * a + ty, j = 16 * b + tx;
                double v = 0.0;
                if (i < n && j <= i) v = (i == j) ? (c.pdiag[i] + c.dsq[i]) : PL[j * ld + i];
                C[tri(a, b)] = v;
            }
        // right-looking, one barrier per column: the owners of column k publish it (unscaled) in shared
        // memory; every thread then applies C_ij -= a_ik a_jk / a_kk to the entries it holds.
#pragma unroll
        for (int kb = 0; kb < NBK; ++kb) {
#pragma unroll 1
            for (int kx = 0; kx < 16; ++kx) {
                const int k = 16 * kb + kx;
                if (!ok || k >= n) break;
                double* col = (k & 1) ? c.colB : c.colA;
                if (tx == kx) {
#pragma unroll
                    for (int a = kb; a < NBK; ++a) {
                        const int i = 16 * a + ty;
                        if (i < n && i >= k) col[i] = C[tri(a, kb)];
                    }
                }
                __syncthreads();
                const double pv = col[k];
                if (!(pv > 0.0) || !isfinite(pv)) { ok = false; break; }
                if (tid == 0) c.piv[k] = pv;
                const double r = 1.0 / pv;
                double ri[NBK], cj[NBK];
#pragma unroll
                for (int a = kb; a < NBK; ++a) { ri[a] = col[16 * a + ty]; cj[a] = col[16 * a + tx] * r; }
#pragma unroll
                for (int a = kb; a < NBK; ++a)
#pragma unroll
                    for (int b = kb; b <= a; ++b)
                        if (b > kb || tx > kx) C[tri(a, b)] -= ri[a] * cj[b];
            }
        }
        __syncthreads();
        if (!ok) return false;
        if (tid < n) c.rdiag[tid] = 1.0 / sqrt(c.piv[tid]);
        __syncthreads();
#pragma unroll
        for (int a = 0; a < NBK; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                const int i = 16 * a + ty, j = 16 * b + tx;
                if (i < n && j <= i) PL[i * ld + j] = C[tri(a, b)] * c.rdiag[j];
            }
    } else {
        for (int i = warp; i < n; i += kWarps)
            for (int j = lane; j < i; j += 32) PL[i * ld + j] = PL[j * ld + i];
        if (tid < n) PL[tid * ld + tid] = c.pdiag[tid] + c.dsq[tid];
        __syncthreads();
        const int tx = tid & 15, ty = tid >> 4;
        for (int k = 0; k < n; ++k) {
            const double akk = PL[k * ld + k];
            if (!(akk > 0.0) || !isfinite(akk)) { ok = false; break; }
            const double r = 1.0 / akk;
            for (int i = k + 1 + ty; i < n; i += 16) {
                const double ci = PL[i * ld + k] * r;
                for (int j = k + 1 + tx; j <= i; j += 16) PL[i * ld + j] -= ci * PL[j * ld + k];
            }
            __syncthreads();
        }
        if (!ok) { __syncthreads(); return false; }
        if (tid < n) c.rdiag[tid] = 1.0 / sqrt(PL[tid * ld + tid]);
        __syncthreads();
        for (int i = warp; i < n; i += kWarps)
            for (int j = lane; j <= i; j += 32) PL[i * ld + j] *= c.rdiag[j];
    }
    __syncthreads();
    // invert the diagonal blocks in place: warp w takes block(s) w, w + 8, ...; lane l builds column l
    const int nblk = (n + 31) >> 5;
    for (int blk = warp; blk < nblk; blk += kWarps) {
        const int r0 = 32 * blk, m = min(32, n - r0);
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k < m) {
                double xk = 0.0;
                if (k == lane) xk = c.rdiag[r0 + k];
                else if (k > lane) xk = -c.rdiag[r0 + k] * x[k];
                x[k] = xk;
#pragma unroll
                for (int i = k + 1; i < 32; ++i)
                    if (i < m) x[i] += PL[(r0 + i) * ld + r0 + k] * xk;
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
__device__ __forceinline__ void solve_phase(Ctx& c) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = c.n, ld = c.ld;
    const double* PL = c.PL;
    double* bs = c.bs;
    const int nblk = (n + 31) >> 5;
    __syncthreads();
    // forward: y = L^-1 b
    for (int blk = 0; blk < nblk; ++blk) {
        const int r0 = 32 * blk, m = min(32, n - r0);
        if (warp == 0) {
            double a0 = 0.0, a1 = 0.0;
            if (lane < m) {
                const double* xr = PL + (r0 + lane) * ld + r0;
                int k = 0;
                for (; k + 1 <= lane; k += 2) { a0 += xr[k] * bs[r0 + k]; a1 += xr[k + 1] * bs[r0 + k + 1]; }
                if (k <= lane) a0 += xr[k] * bs[r0 + k];
            }
            __syncwarp();
            if (lane < m) bs[r0 + lane] = a0 + a1;
        }
        if (r0 + m >= n) break;
        __syncthreads();
        const int i = r0 + m + tid;
        if (i < n) {
            const double* lr = PL + i * ld + r0;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
            for (int k = 0; k < 32; k += 4) {
                a0 += lr[k] * bs[r0 + k];
                a1 += lr[k + 1] * bs[r0 + k + 1];
                a2 += lr[k + 2] * bs[r0 + k + 2];
                a3 += lr[k + 3] * bs[r0 + k + 3];
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
            double a0 = 0.0, a1 = 0.0;
            if (lane < m) {
                const double* xc = PL + r0 * ld + r0 + lane;
                int k = lane;
                for (; k + 1 < m; k += 2) { a0 += xc[k * ld] * bs[r0 + k]; a1 += xc[(k + 1) * ld] * bs[r0 + k + 1]; }
                if (k < m) a0 += xc[k * ld] * bs[r0 + k];
            }
            __syncwarp();
            if (lane < m) bs[r0 + lane] = a0 + a1;
        }
        if (blk == 0) break;
        __syncthreads();
        if (tid < r0) {
            const double* lc = PL + r0 * ld + tid;
            double a0 = 0.0, a1 = 0.0;
            int k = 0;
            for (; k + 1 < m; k += 2) { a0 += lc[k * ld] * bs[r0 + k]; a1 += lc[(k + 1) * ld] * bs[r0 + k + 1]; }
            if (k < m) a0 += lc[k * ld] * bs[r0 + k];
            bs[tid] -= a0 + a1;
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
    const int tid = threadIdx.x;
    const int n = c.n, ld = c.ld;
    const bool act = tid < n;
    const double qi = act ? c.qs[tid] : 0.0;
    const double hi = act ? c.hvec[tid] : 0.0;
    QpOut out;
    out.xi = 0.0; out.pcost = 0.0; out.iters = 0; out.status = 0; out.fatal = false;

    double resx0, resz0;
    {
        double t2[2] = {qi * qi, hi * hi};
        block_reduce<2, 0u>(t2, c);
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
            if (act) c.xs[tid] = xi;
            __syncthreads();
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (act) {
                const double* colp = c.PL + tid;       // P[j][tid], j < tid  (upper triangle, column tid)
                int j = 0;
                for (; j + 3 < tid; j += 4) {
                    a0 += colp[j * ld] * c.xs[j];
                    a1 += colp[(j + 1) * ld] * c.xs[j + 1];
                    a2 += colp[(j + 2) * ld] * c.xs[j + 2];
                    a3 += colp[(j + 3) * ld] * c.xs[j + 3];
                }
                for (; j < tid; ++j) a0 += colp[j * ld] * c.xs[j];
                a1 += c.pdiag[tid] * xi;
                const double* rowp = c.PL + tid * ld;  // P[tid][j], j > tid
                j = tid + 1;
                for (; j + 3 < n; j += 4) {
                    a0 += rowp[j] * c.xs[j];
                    a1 += rowp[j + 1] * c.xs[j + 1];
                    a2 += rowp[j + 2] * c.xs[j + 2];
                    a3 += rowp[j + 3] * c.xs[j + 3];
                }
                for (; j < n; ++j) a0 += rowp[j] * c.xs[j];
            }
            rxi = ((a0 + a1) + (a2 + a3)) + qi;
            const double f0p = act ? (xi * rxi + xi * qi) : 0.0;
            rxi -= zi;
            rzi = si - hi - xi;
            double t5[5] = {f0p, act ? rxi * rxi : 0.0, act ? rzi * rzi : 0.0, act ? zi * rzi : 0.0,
                            act ? (iters == 0 ? si * zi : lam * lam) : 0.0};
            block_reduce<5, 0u>(t5, c);
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
        if (act) c.dsq[tid] = dinv * dinv;
        __syncthreads();
        if (!factor_phase<NBK>(c)) {
            out.status |= HDRT_ST_KKT_FAIL;
            if (iters <= 0) { out.fatal = true; xi = nan(""); }
            break;
        }
        if (iters < 0) {
            // solve [P+I] x = -q - h ; z = -x - h ; s = -z, shifted into the cone
            if (act) c.bs[tid] = -qi - hi;
            solve_phase(c);
            xi = act ? c.bs[tid] : 0.0;
            zi = -xi - hi;
            si = -zi;
            double t4[4] = {act ? si * si : 0.0, act ? -si : -INFINITY, act ? zi * zi : 0.0, act ? -zi : -INFINITY};
            block_reduce<4, 0xAu>(t4, c);
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
            if (act) c.bs[tid] = dxi - dinv * zs;
            solve_phase(c);
            dxi = act ? c.bs[tid] : 0.0;
            dzi = -dinv * dxi - zs;
            dsi = dsi - dzi;
            const double prod = dsi * dzi;
            if (pass == 0) ws3 = prod;
            dsi = dsi / lam;
            dzi = dzi / lam;
            double t3[3] = {act ? prod : 0.0, act ? -dsi : -INFINITY, act ? -dzi : -INFINITY};
            block_reduce<3, 0x6u>(t3, c);
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

__device__ __forceinline__ void hyper_block(Ctx& c, const BlockHyp& hp, int start, int len, double* rho, double* xmx,
                                            bool first_iter) {
    const int tid = threadIdx.x;
    const int n = c.n, nn = c.n * c.n;
    const bool act = tid < len;
    const int gi = start + tid;
    const double xi = act ? c.xs[gi] : 0.0;
    if (act) {
        const double ax = fabs(xi);
        c.xh[gi] = (xi > 0.0 ? 1.0 : (xi < 0.0 ? -1.0 : 0.0)) * sqrt(ax);
    }
    __syncthreads();
    const double xhi = act ? c.xh[gi] : 0.0;
    double bsum[3] = {0, 0, 0}, gd[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    if (act) {
        const double inv2s0 = 1.0 / (2.0 * hp.sigma[0] * hp.sigma[0]);
        const double* __restrict__ pcol = c.pen + (start * n + gi);  // symmetric: read column-wise (coalesced)
        for (int j = 0; j < len; ++j) {
            const int gj = start + j;
            const double xj = c.xs[gj];
            const double m0 = pcol[j * n], m1 = pcol[nn + j * n], m2 = pcol[2 * nn + j * n];
            double gam[3] = {(xi * m0) * xj, (xi * m1) * xj, (xi * m2) * xj};
            if (hp.use_gmat) gam[0] += ((xhi * m1) * c.xh[gj]) * inv2s0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (j == tid) {
                    gd[k] = gam[k] + (hp.s_alpha[k] - 1.0) / hp.s_0[k];
                } else {
                    const double g = gam[k] * c.us[k][gj];
                    bsum[k] += g;
                    mx[k] = fmax(mx[k], fabs(g));
                }
            }
        }
    }
    block_reduce<3, 0x7u>(mx, c);
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
            c.sv[k][gi] = s_hat;
        }
    }
    __syncthreads();
    if (act) {
#pragma unroll
        for (int k = 0; k < 3; ++k) c.us[k][gi] = sqrt(c.sv[k][gi]);
    }
    __syncthreads();
    // rho: alpha / (x' S^1/2 M S^1/2 x / xmx + beta)
    double tr[3] = {0, 0, 0}, tx[3] = {0, 0, 0};
    if (act) {
        const double* __restrict__ pcol = c.pen + (start * n + gi);
        for (int j = 0; j < len; ++j) {
            const int gj = start + j;
            const double xj = c.xs[gj];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double m = pcol[k * nn + j * n];
                tr[k] += (xj * c.us[k][gj]) * m;
                tx[k] += xj * m;
            }
        }
    }
    double t6[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t6[k] = act ? (tr[k] * c.us[k][gi]) * xi : 0.0;
        t6[3 + k] = act ? tx[k] * xi : 0.0;
    }
    block_reduce<6, 0u>(t6, c);
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
__device__ __forceinline__ void weights_phase(Ctx& c, const double* est, double var_floor, bool update_vz) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = c.N, n = c.n, nc = c.nc;
    for (int r = warp; r < N; r += kWarps) {
        double acc = 0.0, accv = 0.0;
        const double* __restrict__ src = c.rm + (size_t)r * n;
        for (int col = lane; col < n; col += 32) {
            const double xv = c.xs[col];
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
        block_reduce<1, 0u>(t1, c);
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
        block_reduce<1, 0u>(t1, c);
        const double mean = t1[0] / (double)N;
        double t2[1] = {0.0};
        for (int r = tid; r < N; r += kThreads) { const double d = c.rv[r] - mean; t2[0] += d * d; }
        block_reduce<1, 0u>(t2, c);
        var_floor = (t2[0] / (double)N) * 1e-7;
    }

    double rho[3], dop_rho[3], xmx[3] = {1, 1, 1}, dop_xmx[3] = {1, 1, 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) { rho[k] = hy.rho_0[k]; dop_rho[k] = hy.dop_rho_0[k]; }
    if (tid < c.nv) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { c.sv[k][tid] = hy.s_0[k]; c.us[k][tid] = sqrt(hy.s_0[k]); }
        c.colA[tid] = 0.0;
        c.colB[tid] = 0.0;
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
        if (tid < n) c.xs[tid] = qo.xi;
        __syncthreads();
        if (init) {
            if (tid < n && p.x_overfit) p.x_overfit[(size_t)b * n + tid] = qo.xi;
            weights_phase(c, nullptr, var_floor, false);
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
        hyper_block(c, hd, c.ns, n - c.ns, rho, xmx, it == 0);
        if (c.dop_a >= 0) hyper_block(c, hp, c.dop_a, c.dop_b - c.dop_a, dop_rho, dop_xmx, it == 0);
        weights_phase(c, est_g, var_floor, c.vz >= 0);
        {   // convergence, qphb.py:597-603,969-970
            const bool act = tid < n;
            const double dx = xi - x_in;
            double t3[3] = {act ? fabs(dx / (x_in + 1e-15)) : 0.0, act ? fabs(dx) : 0.0, act ? x_in : 0.0};
            block_reduce<3, 0x3u>(t3, c);
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
            for (int k = 0; k < 3; ++k) p.s_vectors[((size_t)b * 3 + k) * n + tid] = c.sv[k][tid];
        }
    }
    {
        const bool act = tid < n;
        double t1[1] = {act && !isfinite(xi) ? 1.0 : 0.0};
        block_reduce<1, 0x1u>(t1, c);
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
__global__ void __launch_bounds__(kThreads, (NBK > 0 && NBK <= 7) ? 2 : 1)
qphb_kernel(const hdrt_qphb_problem p, int* work_counter) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_work;
    Ctx c;
    c.N = p.n_rows; c.n = p.n_cols; c.ns = p.n_special; c.nc = p.n_chrono;
    c.dop_a = p.dop_start; c.dop_b = p.dop_end; c.vz = p.vz_index; c.vb_a = p.vb_start; c.vb_b = p.vb_end;
    c.hvec = p.h; c.l1 = p.l1; c.vz_strength = p.vz_strength;
    const SmemLayout L = make_layout(p.n_rows, p.n_cols);
    c.ld = L.ld; c.nv = L.nv; c.ldA = L.ldA;
    c.PL = smem + L.pl;
    double* v = smem + L.vec;
    const int nv = L.nv;
    c.pdiag = v; c.xs = v + nv; c.bs = v + 2 * nv; c.dsq = v + 3 * nv; c.qs = v + 4 * nv; c.rdiag = v + 5 * nv;
    c.piv = v + 6 * nv;
    for (int k = 0; k < 3; ++k) { c.sv[k] = v + (7 + k) * nv; c.us[k] = v + (10 + k) * nv; }
    c.xh = v + 13 * nv; c.colA = v + 14 * nv; c.colB = v + 15 * nv;
    c.w = smem + L.rows;
    c.r2 = c.w + ((p.n_rows + 1) & ~1);
    c.stage = smem + L.stage;
    c.red = smem + L.red;
    c.red_phase = 0;

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
    const SmemLayout L = make_layout(n_rows, n_cols);
    const long long bytes = (long long)L.total * 8;
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
    const int nbk = (p.n_cols + 15) / 16;
    if (nbk <= 7) return launch_qphb<7>(h, p, (size_t)smem, st);
    if (nbk <= 10) return launch_qphb<10>(h, p, (size_t)smem, st);
    return launch_qphb<0>(h, p, (size_t)smem, st);
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
