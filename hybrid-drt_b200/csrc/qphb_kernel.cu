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
//   PL   n x ld (ld odd): strict upper triangle = P of the current QP, lower triangle + diagonal =
//        Cholesky factor of H = P + diag(1/d^2) of the current interior-point iteration
//   17 vectors of length n (pdiag, x broadcast, solve rhs, ...), 2 vectors of length N (w, r^2),
//   a kStageRows x ldA staging tile for the Gram pass.
// The design matrix rm, the variance-estimation matrix vmm and the penalty matrices are shared by the
// batch and stay in global memory (L2 / L1 resident, read-only path).
#include "common.cuh"

namespace hdrt {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStageRows = 8;
constexpr int kMaxCols = 256;
constexpr int kRedSlots = 8;

// cvxopt coneqp defaults (cvxopt 1.3 coneprog.py; the reference only sets show_progress, qphb.py:25)
constexpr double kAbsTol = 1e-7;
constexpr double kRelTol = 1e-6;
constexpr double kFeasTol = 1e-7;
constexpr int kMaxIpm = 100;
constexpr double kStep = 0.99;

struct SmemLayout {
    int ld;     // leading dimension of PL (odd)
    int nv;     // padded vector length
    int ldA;    // staging leading dimension
    int nb4;    // 4x4 tile count per side
    size_t pl, vec, rows, stage, red, total;  // offsets in doubles
};

__host__ __device__ inline SmemLayout make_layout(int N, int n) {
    SmemLayout L;
    L.ld = n | 1;
    L.nv = (n + 1) & ~1;
    L.nb4 = (n + 3) / 4;
    L.ldA = 4 * L.nb4 + 4;
    L.pl = 0;
    L.vec = L.pl + (size_t)n * L.ld + (((size_t)n * L.ld) & 1);
    L.rows = L.vec + (size_t)17 * L.nv;
    L.stage = L.rows + (size_t)2 * ((N + 1) & ~1);
    L.red = L.stage + (size_t)kStageRows * L.ldA;
    L.total = L.red + (size_t)kRedSlots * kWarps;
    return L;
}

// Reduce K per-thread values over the block; bit k of MAXMASK selects max instead of sum.
// Result is broadcast to every thread.  Two barriers.
template <int K, unsigned MAXMASK>
__device__ __forceinline__ void block_reduce(double (&v)[K], double* red) {
    static_assert(K <= kRedSlots, "too many reduction slots");
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = ((MAXMASK >> k) & 1u) ? warp_max(v[k]) : warp_sum(v[k]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
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

struct Ctx {
    // problem
    int N, n, ns, nc, nd, dop_a, dop_b, vz, vb_a, vb_b;
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
    SmemLayout L;
    double *PL, *pdiag, *xs, *bs, *dsq, *qs, *rdiag, *sv[3], *us[3], *xh, *tv[3], *w, *r2, *stage, *red;
};

// ------------------------------------------------------------------------------------------------
// Gram: P = (W rm)^T (W rm) + L2  (upper triangle + diagonal),  q = -(W rm)^T (W rv) + l1
// L2 = sum_k S_k^1/2 M~_k S_k^1/2 as in qphb.calculate_qp_l2_matrix (qphb.py:53-120)
// ------------------------------------------------------------------------------------------------
struct L2Factors {
    double drt[3];  // l2_lambda_0 * dw_k * rho_k        (0 => derivative order unused)
    double dop[3];  // dop_l2_lambda_0 * dop_dw_k * dop_rho_k
    bool use[3];
};

__device__ __forceinline__ void tile_coords(int t, int nb4, int& bi, int& bj) {
    // row-major enumeration of the upper-triangular tile set: row bi holds nb4 - bi tiles
    const double bb = 2.0 * nb4 + 1.0;
    int r = (int)floor((bb - sqrt(bb * bb - 8.0 * (double)t)) * 0.5);
    if (r < 0) r = 0;
    while (r > 0 && (r * (2 * nb4 - r + 1)) / 2 > t) --r;
    while (((r + 1) * (2 * nb4 - r)) / 2 <= t) ++r;
    bi = r;
    bj = r + (t - (r * (2 * nb4 - r + 1)) / 2);
}

__device__ __forceinline__ double l2_entry(const Ctx& c, const L2Factors& f, int i, int j) {
    double acc = 0.0;
    const bool drt = (i >= c.ns) && (j >= c.ns);
    const bool dop = (c.dop_a >= 0) && (i >= c.dop_a) && (i < c.dop_b) && (j >= c.dop_a) && (j < c.dop_b);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!f.use[k]) continue;
        double m = c.pen[(size_t)k * c.n * c.n + (size_t)i * c.n + j];
        if (drt) m *= f.drt[k];
        if (dop) m *= f.dop[k];
        acc += (c.us[k][i] * m) * c.us[k][j];
    }
    return acc;
}

// p_out/q_out != nullptr: also write the full symmetric matrix and vector to global (calculate_pq).
__device__ void gram_phase(const Ctx& c, const L2Factors& f, bool l1_scalar, double l1_value, double* p_out,
                           double* q_out) {
    const int tid = threadIdx.x;
    const int n = c.n, N = c.N, ldA = c.L.ldA, nb4 = c.L.nb4, ld = c.L.ld;
    const int ntiles = nb4 * (nb4 + 1) / 2;
    const int ycol = 4 * nb4;
    const int npass = (ntiles + 2 * kThreads - 1) / (2 * kThreads);

    for (int pass = 0; pass < npass; ++pass) {
        int t0 = pass * 2 * kThreads + tid, t1 = t0 + kThreads;
        const bool has0 = t0 < ntiles, has1 = t1 < ntiles;
        int bi0 = 0, bj0 = 0, bi1 = 0, bj1 = 0;
        if (has0) tile_coords(t0, nb4, bi0, bj0);
        if (has1) tile_coords(t1, nb4, bi1, bj1);
        double acc0[16], acc1[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) acc0[e] = acc1[e] = 0.0;
        double qacc = 0.0;

        for (int r0 = 0; r0 < N; r0 += kStageRows) {
            const int rows = min(kStageRows, N - r0);
            __syncthreads();
            for (int idx = tid; idx < rows * ldA; idx += kThreads) {
                const int rr = idx / ldA, col = idx - rr * ldA;
                const int r = r0 + rr;
                double v = 0.0;
                if (col < n) {
                    v = (col == c.vz) ? c.vzcol[r] : c.rm[(size_t)r * n + col];
                    v *= c.w[r];
                } else if (col == ycol) {
                    v = c.w[r] * c.rv[r];
                }
                c.stage[idx] = v;
            }
            __syncthreads();
            for (int rr = 0; rr < rows; ++rr) {
                const double* row = c.stage + rr * ldA;
                if (has0) {
                    const double2 a01 = *reinterpret_cast<const double2*>(row + 4 * bi0);
                    const double2 a23 = *reinterpret_cast<const double2*>(row + 4 * bi0 + 2);
                    const double2 b01 = *reinterpret_cast<const double2*>(row + 4 * bj0);
                    const double2 b23 = *reinterpret_cast<const double2*>(row + 4 * bj0 + 2);
                    const double a[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) acc0[ii * 4 + jj] += a[ii] * b[jj];
                }
                if (has1) {
                    const double2 a01 = *reinterpret_cast<const double2*>(row + 4 * bi1);
                    const double2 a23 = *reinterpret_cast<const double2*>(row + 4 * bi1 + 2);
                    const double2 b01 = *reinterpret_cast<const double2*>(row + 4 * bj1);
                    const double2 b23 = *reinterpret_cast<const double2*>(row + 4 * bj1 + 2);
                    const double a[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) acc1[ii * 4 + jj] += a[ii] * b[jj];
                }
                if (pass == 0 && tid < n) qacc += row[tid] * row[ycol];
            }
        }
        // write back with the L2 term
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const bool has = which ? has1 : has0;
            if (!has) continue;
            const int bi = which ? bi1 : bi0, bj = which ? bj1 : bj0;
            const double* acc = which ? acc1 : acc0;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int i = 4 * bi + ii, j = 4 * bj + jj;
                    if (i < n && j < n && i <= j) {
                        const double val = acc[ii * 4 + jj] + l2_entry(c, f, i, j);
                        if (i == j) c.pdiag[i] = val; else c.PL[(size_t)i * ld + j] = val;
                        if (p_out) {
                            p_out[(size_t)i * n + j] = val;
                            p_out[(size_t)j * n + i] = val;
                        }
                    }
                }
            }
        }
        if (pass == 0 && tid < n) {
            const double qv = -qacc + (l1_scalar ? l1_value : c.l1[tid]);
            c.qs[tid] = qv;
            if (q_out) q_out[tid] = qv;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Cholesky of H = P + diag(dsq) into the lower triangle of PL.  Returns false on breakdown.
// ------------------------------------------------------------------------------------------------
__device__ bool factor_phase(const Ctx& c) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = c.n, ld = c.L.ld;
    double* PL = c.PL;
    for (int i = warp; i < n; i += kWarps)
        for (int j = lane; j < i; j += 32) PL[(size_t)i * ld + j] = PL[(size_t)j * ld + i];
    if (tid < n) PL[(size_t)tid * ld + tid] = c.pdiag[tid] + c.dsq[tid];
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        const double akk = PL[(size_t)k * ld + k];
        if (!(akk > 0.0) || !isfinite(akk)) { ok = false; break; }  // uniform: every thread reads the same akk
        const double r = 1.0 / akk;
        for (int i = k + 1 + ty; i < n; i += 16) {
            const double ci = PL[(size_t)i * ld + k] * r;
            for (int j = k + 1 + tx; j <= i; j += 16) PL[(size_t)i * ld + j] -= ci * PL[(size_t)j * ld + k];
        }
        __syncthreads();
    }
    if (!ok) { __syncthreads(); return false; }
    if (tid < n) c.rdiag[tid] = 1.0 / sqrt(PL[(size_t)tid * ld + tid]);
    __syncthreads();
    for (int i = warp; i < n; i += kWarps)
        for (int j = lane; j < i; j += 32) PL[(size_t)i * ld + j] *= c.rdiag[j];
    __syncthreads();
    return true;
}

// Solve L L^T u = bs in place (warp 0 only; the rest of the block waits at the caller's barrier).
__device__ void solve_warp0(const Ctx& c) {
    const int lane = threadIdx.x & 31;
    const int n = c.n, ld = c.L.ld;
    const double* PL = c.PL;
    const double* rdiag = c.rdiag;
    constexpr int S = kMaxCols / 32;
    const int ns = (n + 31) >> 5;
    double v[S];
#pragma unroll
    for (int m = 0; m < S; ++m) {
        const int i = lane + 32 * m;
        v[m] = (i < n) ? c.bs[i] : 0.0;
    }
    // forward: L y = b (column sweeps)
#pragma unroll
    for (int m0 = 0; m0 < S; ++m0) {
        if (m0 < ns) {
            const int kend = min(32, n - 32 * m0);
            for (int kk = 0; kk < kend; ++kk) {
                const int k = 32 * m0 + kk;
                const double yk = __shfl_sync(kFull, v[m0] * rdiag[k], kk);
                if (lane == kk) v[m0] = yk;
#pragma unroll
                for (int m = m0; m < S; ++m) {
                    const int i = lane + 32 * m;
                    if (m < ns && i > k && i < n) v[m] -= PL[(size_t)i * ld + k] * yk;
                }
            }
        }
    }
    // backward: L^T x = y (row sweeps)
#pragma unroll
    for (int m0 = S - 1; m0 >= 0; --m0) {
        if (m0 < ns) {
            const int kend = min(32, n - 32 * m0);
            for (int kk = kend - 1; kk >= 0; --kk) {
                const int k = 32 * m0 + kk;
                const double xk = __shfl_sync(kFull, v[m0] * rdiag[k], kk);
                if (lane == kk) v[m0] = xk;
#pragma unroll
                for (int m = 0; m <= m0; ++m) {
                    const int j = lane + 32 * m;
                    if (j < k) v[m] -= PL[(size_t)k * ld + j] * xk;
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < S; ++m) {
        const int i = lane + 32 * m;
        if (i < n) c.bs[i] = v[m];
    }
}

__device__ __forceinline__ void solve_phase(const Ctx& c) {
    __syncthreads();
    if (threadIdx.x < 32) solve_warp0(c);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// QP: cvxopt coneqp for the orthant cone with G = -I (see oracle/coneqp.py for the restatement and
// its provenance).  Thread i < n owns element i of every vector.  Returns x_i; fills status bits.
// ------------------------------------------------------------------------------------------------
struct QpOut {
    double xi;
    double pcost;
    int iters;
    int status;  // HDRT_ST_QP_MAXITERS / HDRT_ST_KKT_FAIL bits
    bool fatal;  // Cholesky failed before the first iterate existed (cvxopt raises ValueError)
};

__device__ QpOut qp_phase(const Ctx& c) {
    const int tid = threadIdx.x;
    const int n = c.n, ld = c.L.ld;
    const bool act = tid < n;
    const double qi = act ? c.qs[tid] : 0.0;
    const double hi = act ? c.hvec[tid] : 0.0;
    QpOut out;
    out.xi = 0.0; out.pcost = 0.0; out.iters = 0; out.status = 0; out.fatal = false;

    double r4[4];
    r4[0] = qi * qi; r4[1] = hi * hi; r4[2] = 0.0; r4[3] = 0.0;
    {
        double t2[2] = {r4[0], r4[1]};
        block_reduce<2, 0u>(t2, c.red);
        r4[0] = t2[0]; r4[1] = t2[1];
    }
    const double resx0 = fmax(1.0, sqrt(r4[0]));
    const double resz0 = fmax(1.0, sqrt(r4[1]));

    // initial point: W = I
    if (act) c.dsq[tid] = 1.0;
    __syncthreads();
    if (!factor_phase(c)) { out.fatal = true; out.status = HDRT_ST_KKT_FAIL; out.xi = nan(""); return out; }
    if (act) c.bs[tid] = -qi - hi;
    solve_phase(c);
    double xi = act ? c.bs[tid] : 0.0;
    double zi = -xi - hi;
    double si = -zi;
    {
        double t4[4] = {act ? si * si : 0.0, act ? -si : -INFINITY, act ? zi * zi : 0.0, act ? -zi : -INFINITY};
        block_reduce<4, 0xAu>(t4, c.red);
        const double nrms = sqrt(t4[0]), ts = t4[1], nrmz = sqrt(t4[2]), tz = t4[3];
        if (ts >= -1e-8 * fmax(nrms, 1.0)) si += 1.0 + ts;
        if (tz >= -1e-8 * fmax(nrmz, 1.0)) zi += 1.0 + tz;
    }
    double di = 1.0, dinv = 1.0, lam = 1.0;
    double gap = 0.0;
    double pcost = 0.0;
    int iters = 0;
    for (iters = 0; iters <= kMaxIpm; ++iters) {
        if (act) c.xs[tid] = xi;
        __syncthreads();
        double a0 = 0.0, a1 = 0.0;
        if (act) {
            int j = 0;
            for (; j + 1 < tid; j += 2) {
                a0 += c.PL[(size_t)j * ld + tid] * c.xs[j];
                a1 += c.PL[(size_t)(j + 1) * ld + tid] * c.xs[j + 1];
            }
            for (; j < tid; ++j) a0 += c.PL[(size_t)j * ld + tid] * c.xs[j];
            a1 += c.pdiag[tid] * xi;
            j = tid + 1;
            for (; j + 1 < n; j += 2) {
                a0 += c.PL[(size_t)tid * ld + j] * c.xs[j];
                a1 += c.PL[(size_t)tid * ld + j + 1] * c.xs[j + 1];
            }
            for (; j < n; ++j) a0 += c.PL[(size_t)tid * ld + j] * c.xs[j];
        }
        double rxi = (a0 + a1) + qi;
        const double f0p = act ? (xi * rxi + xi * qi) : 0.0;
        rxi -= zi;
        const double rzi = si - hi - xi;
        double t5[5] = {f0p, act ? rxi * rxi : 0.0, act ? rzi * rzi : 0.0, act ? zi * rzi : 0.0,
                        act ? (iters == 0 ? si * zi : lam * lam) : 0.0};
        block_reduce<5, 0u>(t5, c.red);
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
        const bool done = (pres <= kFeasTol) && (dres <= kFeasTol) && ((gap <= kAbsTol) || (have_rel && relgap <= kRelTol));
        if (done) break;
        if (iters == kMaxIpm) { out.status |= HDRT_ST_QP_MAXITERS; break; }

        if (iters == 0) {
            di = sqrt(si / zi);
            dinv = 1.0 / di;
            lam = sqrt(si * zi);
        }
        const double lamsq = lam * lam;
        if (act) c.dsq[tid] = dinv * dinv;
        __syncthreads();
        if (!factor_phase(c)) {
            out.status |= HDRT_ST_KKT_FAIL;
            if (iters == 0) out.fatal = true;
            break;
        }
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
            block_reduce<3, 0x6u>(t3, c.red);
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
    out.iters = iters;
    return out;
}

// ------------------------------------------------------------------------------------------------
// Hyper-parameter updates for one coefficient block (DRT or DOP): qphb.solve_s / solve_rho
// ------------------------------------------------------------------------------------------------
struct BlockHyp {
    double dw[3], sigma[3], s_alpha[3], s_0[3], rho_alpha[3], rho_0[3];
    bool use_gmat;  // DRT block: k = 0 gets G = Xh M1 Xh (qphb.py:769-772); DOP block: 0 (drt1d.py quirk)
};

__device__ void hyper_block(const Ctx& c, const BlockHyp& hp, int start, int len, double* rho, double* xmx,
                            bool first_iter) {
    const int tid = threadIdx.x;
    const int n = c.n;
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
        for (int j = 0; j < len; ++j) {
            const int gj = start + j;
            const size_t off = (size_t)gj * n + gi;  // symmetric: read column-wise for coalescing
            const double xj = c.xs[gj];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (!(hp.dw[k] > 0.0)) continue;
                const double m = c.pen[(size_t)k * n * n + off];
                double gam = (xi * m) * xj;
                if (k == 0 && hp.use_gmat) {
                    const double m1 = c.pen[(size_t)n * n + off];
                    gam += ((xhi * m1) * c.xh[gj]) * inv2s0;
                }
                if (j == tid) {
                    gd[k] = gam + (hp.s_alpha[k] - 1.0) / hp.s_0[k];
                } else {
                    const double g = gam * c.us[k][gj];
                    bsum[k] += g;
                    mx[k] = fmax(mx[k], fabs(g));
                }
            }
        }
    }
    block_reduce<3, 0x7u>(mx, c.red);
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
        for (int j = 0; j < len; ++j) {
            const int gj = start + j;
            const size_t off = (size_t)gj * n + gi;
            const double xj = c.xs[gj];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double m = c.pen[(size_t)k * n * n + off];
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
    block_reduce<6, 0u>(t6, c.red);
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
__device__ void weights_phase(const Ctx& c, const double* est, double var_floor, bool update_vz) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = c.N, n = c.n, nc = c.nc;
    for (int r = warp; r < N; r += kWarps) {
        double acc = 0.0, accv = 0.0;
        for (int col = lane; col < n; col += 32) {
            const double xv = c.xs[col];
            if (col == c.vz) {
                acc += c.vzcol[r] * xv;
            } else {
                const double t = c.rm[(size_t)r * n + col] * xv;
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
        block_reduce<1, 0u>(t1, c.red);
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
            for (int col = lane; col < ne; col += 32) acc += c.vmm_eis[(size_t)(r - nc) * ne + col] * c.r2[nc + col];
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
// One spectrum
// ------------------------------------------------------------------------------------------------
__device__ void fit_one(const hdrt_qphb_problem& p, int b, Ctx& c) {
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
        block_reduce<1, 0u>(t1, c.red);
        const double mean = t1[0] / (double)N;
        double t2[1] = {0.0};
        for (int r = tid; r < N; r += kThreads) { const double d = c.rv[r] - mean; t2[0] += d * d; }
        block_reduce<1, 0u>(t2, c.red);
        var_floor = (t2[0] / (double)N) * 1e-7;
    }

    double rho[3], dop_rho[3], xmx[3] = {1, 1, 1}, dop_xmx[3] = {1, 1, 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) { rho[k] = hy.rho_0[k]; dop_rho[k] = hy.dop_rho_0[k]; }
    if (tid < n) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { c.sv[k][tid] = hy.s_0[k]; c.us[k][tid] = sqrt(hy.s_0[k]); }
    }
    for (int r = tid; r < N; r += kThreads) {
        c.w[r] = 1.0;
        if (c.vz >= 0) c.vzcol[r] = 0.0;
    }
    __syncthreads();

    int status = 0, n_ipm = 0;
    L2Factors f;
    // ---- initialize_weights: overfit QP with iw lambdas (drt1d.py:640-645, qphb.py:1609-1681)
    {
        const double dop_l2 = hy.dop_l2_lambda_0 * (hy.iw_l2_lambda_0 / hy.l2_lambda_0);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            f.use[k] = hy.derivative_weights[k] > 0.0;
            f.drt[k] = hy.iw_l2_lambda_0 * hy.derivative_weights[k] * rho[k];
            f.dop[k] = dop_l2 * hy.dop_derivative_weights[k] * dop_rho[k];
        }
    }
    gram_phase(c, f, true, hy.iw_l1_lambda_0, nullptr, nullptr);
    QpOut qo = qp_phase(c);
    status |= qo.status;
    n_ipm += qo.iters;
    bool fatal = qo.fatal;
    if (tid < n) {
        c.xs[tid] = qo.xi;
        if (p.x_overfit) p.x_overfit[(size_t)b * n + tid] = qo.xi;
    }
    __syncthreads();
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

    double xi = 1e-6;  // drt1d.py:612
    double fun = 0.0;
    int it = 0;
    bool conv = false;
    while (!fatal && it < hy.max_iter) {
        const double x_in = xi;
        // weight factors, drt1d.py:881-892
        for (int r = tid; r < N; r += kThreads) {
            double w = c.w[r];
            if (p.hybrid) w *= (r < c.nc) ? hy.chrono_weight_factor : hy.eis_weight_factor;
            if (it > 0) w = w * hy.weight_factor;
            c.w[r] = w;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            f.drt[k] = hy.l2_lambda_0 * hy.derivative_weights[k] * rho[k];
            f.dop[k] = hy.dop_l2_lambda_0 * hy.dop_derivative_weights[k] * dop_rho[k];
        }
        __syncthreads();
        gram_phase(c, f, false, 0.0, nullptr, nullptr);
        qo = qp_phase(c);
        status |= qo.status;
        n_ipm += qo.iters;
        if (qo.fatal) { fatal = true; xi = qo.xi; break; }
        xi = qo.xi;
        fun = qo.pcost;
        if (tid < n) c.xs[tid] = xi;
        __syncthreads();
        hyper_block(c, hd, c.ns, n - c.ns, rho, xmx, it == 0);
        if (c.dop_a >= 0) hyper_block(c, hp, c.dop_a, c.dop_b - c.dop_a, dop_rho, dop_xmx, it == 0);
        weights_phase(c, est_g, var_floor, c.vz >= 0);
        // convergence, qphb.py:597-603,969-970
        {
            const bool act = tid < n;
            const double dx = xi - x_in;
            double t3[3] = {act ? fabs(dx / (x_in + 1e-15)) : 0.0, act ? fabs(dx) : 0.0, act ? x_in : 0.0};
            block_reduce<3, 0x3u>(t3, c.red);
            const double atol = (t3[2] / (double)n) * 1e-3;
            conv = (t3[0] <= hy.xtol) || (t3[1] <= atol);
        }
        ++it;
        if (conv) break;
    }

    // ---- outputs
    if (tid < n) {
        p.x[(size_t)b * n + tid] = xi;
        if (p.s_vectors) {
#pragma unroll
            for (int k = 0; k < 3; ++k) p.s_vectors[((size_t)b * 3 + k) * n + tid] = c.sv[k][tid];
        }
    }
    if (p.weights) for (int r = tid; r < N; r += kThreads) p.weights[(size_t)b * N + r] = c.w[r];
    {
        const bool act = tid < n;
        double t1[1] = {act && !isfinite(xi) ? 1.0 : 0.0};
        block_reduce<1, 0x1u>(t1, c.red);
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
        if (p.n_outer) p.n_outer[b] = it;
        if (p.n_ipm) p.n_ipm[b] = n_ipm;
        if (p.status) p.status[b] = status;
    }
    // ---- qphb.calculate_pq with the final state and the scaled weights (drt1d.py:991-1008)
    if (p.p_matrix != nullptr && !fatal) {
        for (int r = tid; r < N; r += kThreads) {
            double w = c.w[r] * hy.weight_factor;
            if (p.hybrid) w *= (r < c.nc) ? hy.chrono_weight_factor : hy.eis_weight_factor;
            c.w[r] = w;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            f.drt[k] = hy.l2_lambda_0 * hy.derivative_weights[k] * rho[k];
            f.dop[k] = hy.dop_l2_lambda_0 * hy.dop_derivative_weights[k] * dop_rho[k];
        }
        __syncthreads();
        gram_phase(c, f, false, 0.0, p.p_matrix + (size_t)b * n * n, p.q_vector ? p.q_vector + (size_t)b * n : nullptr);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 2) qphb_kernel(const hdrt_qphb_problem p, int* work_counter) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_work;
    Ctx c;
    c.N = p.n_rows; c.n = p.n_cols; c.ns = p.n_special; c.nc = p.n_chrono;
    c.dop_a = p.dop_start; c.dop_b = p.dop_end; c.vz = p.vz_index; c.vb_a = p.vb_start; c.vb_b = p.vb_end;
    c.hvec = p.h; c.l1 = p.l1; c.vz_strength = p.vz_strength;
    c.L = make_layout(p.n_rows, p.n_cols);
    c.PL = smem + c.L.pl;
    double* v = smem + c.L.vec;
    const int nv = c.L.nv;
    c.pdiag = v; c.xs = v + nv; c.bs = v + 2 * nv; c.dsq = v + 3 * nv; c.qs = v + 4 * nv; c.rdiag = v + 5 * nv;
    for (int k = 0; k < 3; ++k) { c.sv[k] = v + (6 + k) * nv; c.us[k] = v + (9 + k) * nv; c.tv[k] = v + (13 + k) * nv; }
    c.xh = v + 12 * nv;
    c.w = smem + c.L.rows;
    c.r2 = c.w + ((p.n_rows + 1) & ~1);
    c.stage = smem + c.L.stage;
    c.red = smem + c.L.red;

    while (true) {
        if (threadIdx.x == 0) s_work = atomicAdd(work_counter, 1);
        __syncthreads();
        const int b = s_work;
        __syncthreads();
        if (b >= p.batch) break;
        fit_one(p, b, c);
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
    HDRT_CUDA_CHECK(cudaFuncSetAttribute(qphb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    HDRT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qphb_kernel, kThreads, (size_t)smem));
    if (occ < 1) { set_error("kernel cannot be resident (smem %lld)", smem); return HDRT_ERR_UNSUPPORTED; }
    int grid = h->sm_count * occ;
    if (grid > p.batch) grid = p.batch;
    HDRT_CUDA_CHECK(cudaMemsetAsync(h->work_counter, 0, sizeof(int), st));
    qphb_kernel<<<grid, kThreads, (size_t)smem, st>>>(p, h->work_counter);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
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
