// Cross-observation resolve QP (reference: hybdrt/mapping/resolve.py:176-341, the cvxopt call at :334), batched over
// the windows of a map: one CTA of 16 warps per window.  Included by qphb_kernel.cu, which holds the tile helpers.
//
// A window couples nr observations (7 by default) of nc parameters each:
//     P = blockdiag(P_1 .. P_nr) + My (x) diag(param_scale),      q = [q_1 .. q_nr],      -x <= h,
// n = nr nc unknowns (about 650).  The reference hands the dense matrix to cvxopt's coneqp; this kernel runs the same
// interior-point iteration (same start point, step rule and stopping test as qp_phase / oracle/coneqp.py).  P is never
// formed: its tiles are assembled from the per-observation matrices when a factorisation starts from them.  The
// Cholesky factor of H = P + diag(z / s) (T (T + 1) / 2 tiles of 8 x 8, about 1.8 MB) does not fit in shared memory, so it
// lives in a per-CTA scratch area in global memory (L2): left-looking over tile columns, every warp owning tile rows
// k + w, k + w + 16, ... of column k; the tiles of row k, which every warp needs, are staged in shared memory once per
// column, the own rows stream from L2 with the next term's tiles requested before the DMMAs of the current one.  The
// negated inverses of the diagonal tiles stay in shared memory for the substitutions.
#pragma once

namespace hdrt {
namespace rs {

constexpr int kWarps = 16, kThreads = 32 * kWarps;
constexpr int kMaxN = 1024;                 // unknowns per window (nr * nc)
constexpr int kMaxT = kMaxN / 8;
constexpr int kRows = kMaxT / kWarps;       // tile rows of one column per warp
constexpr int EU = kMaxN / kThreads;        // elements per thread

__host__ __device__ inline long long work_doubles(int n) { const long long T = (n + 7) / 8; return T * (T + 1) / 2 * 64; }
__host__ __device__ inline size_t smem_bytes(int n) {
    const size_t T = (n + 7) / 8;
    return sizeof(double) * (3 * 8 * T          // bs, ys, dsq
                             + 64 * T            // tiles of row k
                             + 64 * T            // -L_kk^-1 for every k
                             + 16 * kWarps       // per-warp partial sums of a substitution step
                             + 2 * 8 * kWarps);  // block reductions
}

struct RCtx {
    int n, T, nr, nc, lane, g, q, warp;
    const double* __restrict__ p;        // [nr][nc][nc] of this window
    const double* __restrict__ my;       // [nr][nr]
    const double* __restrict__ pscale;   // [nc]
    double* tiles;                       // global scratch: tile (j, i) at tidx(j, i) * 64
    double* bs; double* ys; double* dsq; double* zrow; double* ninv; double* part; double* red;
    int red_phase;
};

__device__ __forceinline__ int tidx(int j, int i) { return j * (j + 1) / 2 + i; }

// tile (j, k) of -P for this lane
__device__ __forceinline__ double2 neg_p_tile(const RCtx& c, int j, int k) {
    const int r = 8 * j + c.g, c0 = 8 * k + 2 * c.q;
    double o[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int cc = c0 + e;
        double v;
        if (r < c.n && cc < c.n) {
            const int orow = r / c.nc, prow = r - orow * c.nc, ocol = cc / c.nc, pcol = cc - ocol * c.nc;
            v = 0.0;
            if (orow == ocol) v = c.p[((size_t)orow * c.nc + prow) * c.nc + pcol];
            if (prow == pcol) v += c.pscale[prow] * c.my[orow * c.nr + ocol];
        } else {
            v = (r == cc) ? 1.0 : 0.0;
        }
        o[e] = -v;
    }
    return make_double2(o[0], o[1]);
}

__device__ __forceinline__ double2 ldg2(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int K, unsigned MAXMASK>
__device__ __forceinline__ void rblock_reduce(double (&v)[K], RCtx& c) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = ((MAXMASK >> k) & 1u) ? warp_max(v[k]) : warp_sum(v[k]);
    double* red = c.red + (c.red_phase & 1) * (8 * kWarps);
    c.red_phase ^= 1;
    if (c.lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[k * kWarps + c.warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = red[k * kWarps];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
            const double t = red[k * kWarps + w];
            a = ((MAXMASK >> k) & 1u) ? fmax(a, t) : a + t;
        }
        v[k] = a;
    }
}

// H = P + diag(dsq) = L L^T; tiles of L to the global scratch, -L_kk^-1 to shared memory.  false on breakdown.
__device__ __noinline__ bool rfactor(const RCtx& cref) {
    const RCtx c = cref;
    const int T = c.T, lane = c.lane, w = c.warp;
    const unsigned zs = smem_u32(c.zrow) + 16 * lane, ns = smem_u32(c.ninv) + 16 * lane;
    int ok = 1;
#pragma unroll 1
    for (int k = 0; k < T; ++k) {
        // row k of L (the Z operand of every term of this column) to shared memory
        for (int m = w; m < k; m += kWarps)
            sts2a(zs + m * 512, ldg2(c.tiles + (size_t)tidx(k, m) * 64 + 2 * lane));
        __syncthreads();
        double2 acc[kRows];
        int jrow[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            jrow[i] = k + w + kWarps * i;
            acc[i] = jrow[i] < T ? neg_p_tile(c, jrow[i], k) : make_double2(0.0, 0.0);
        }
        if (k > 0) {
            double2 X[kRows];
#pragma unroll
            for (int i = 0; i < kRows; ++i) X[i] = ldg2(c.tiles + (size_t)tidx(min(jrow[i], T - 1), 0) * 64 + 2 * lane);
#pragma unroll 1
            for (int m = 0; m < k; ++m) {
                const double2 Z = lds2a(zs + m * 512);
                const int mn = min(m + 1, k - 1);
                double2 Xn[kRows];
#pragma unroll
                for (int i = 0; i < kRows; ++i) Xn[i] = ldg2(c.tiles + (size_t)tidx(min(jrow[i], T - 1), mn) * 64 + 2 * lane);
#pragma unroll
                for (int i = 0; i < kRows; ++i) if (jrow[i] < T) mma_lo(acc[i], X[i], Z);
#pragma unroll
                for (int i = 0; i < kRows; ++i) if (jrow[i] < T) mma_hi(acc[i], X[i], Z);
#pragma unroll
                for (int i = 0; i < kRows; ++i) X[i] = Xn[i];
            }
        }
        if (w == 0) {                       // the diagonal tile is this warp's first row
            double2 sk = acc[0];
            const double d = c.dsq[8 * k + c.g];
            if (c.g == 2 * c.q) sk.x -= d;
            if (c.g == 2 * c.q + 1) sk.y -= d;
            if (!diag_factor(sk, ns + k * 512, lane)) ok = 0;
        }
        ok = __syncthreads_and(ok);
        if (!ok) return false;
        const double2 bn = lds2a(ns + k * 512);
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            if (jrow[i] < T && jrow[i] > k) {
                double2 r2 = make_double2(0.0, 0.0);
                tile_mma(r2, acc[i], bn);
                *reinterpret_cast<double2*>(c.tiles + (size_t)tidx(jrow[i], k) * 64 + 2 * lane) = r2;
            }
        }
        __syncthreads();                    // column k is visible to the whole CTA
    }
    return true;
}

// H u = b by substitution (b in bs, u in ys), the terms of a step spread over the warps
__device__ __noinline__ void rsolve(const RCtx& cref) {
    const RCtx c = cref;
    const int T = c.T, lane = c.lane, w = c.warp, g = c.g, q = c.q;
    const unsigned nt = smem_u32(c.ninv) + (16 * q + g) * 8;       // a diagonal slot read transposed
#pragma unroll 1
    for (int k = 0; k < T; ++k) {             // forward: y_k = N_k (sum_{m<k} L_km y_m - b_k)
        double p = 0.0;
        for (int m = w; m < k; m += kWarps) {
            const double2 A = ldg2(c.tiles + (size_t)tidx(k, m) * 64 + 2 * lane);
            const double2 y = lds2(c.ys + 8 * m + 2 * q);
            p = fma(A.x, y.x, p);
            p = fma(A.y, y.y, p);
        }
        p = reduce_q(p);
        if (q == 0) c.part[w * 16 + g] = p;
        __syncthreads();
        if (w == 0) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < kWarps; ++ww) s += c.part[ww * 16 + g];
            const double t = s - c.bs[8 * k + g];
            const double2 Nt = lds2t(nt + k * 512);
            const double y0 = reduce_g(Nt.x * t), y1 = reduce_g(Nt.y * t);
            if (g == 0) sts2(c.ys + 8 * k + 2 * q, make_double2(y0, y1));
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int k = T - 1; k >= 0; --k) {        // backward: u_k = N_k^T (sum_{j>k} L_jk^T u_j - y_k)
        double2 p = make_double2(0.0, 0.0);
        for (int j = k + 1 + w; j < T; j += kWarps) {
            const double2 A = ldg2(c.tiles + (size_t)tidx(j, k) * 64 + 2 * lane);
            const double u = c.ys[8 * j + g];
            p.x = fma(A.x, u, p.x);
            p.y = fma(A.y, u, p.y);
        }
        p.x = reduce_g(p.x);
        p.y = reduce_g(p.y);
        if (g == 0) sts2(c.part + w * 16 + 2 * q, p);
        __syncthreads();
        if (w == 0) {
            double2 s = make_double2(0.0, 0.0);
#pragma unroll
            for (int ww = 0; ww < kWarps; ++ww) { const double2 t = lds2(c.part + ww * 16 + 2 * q); s.x += t.x; s.y += t.y; }
            const double2 yk = lds2(c.ys + 8 * k + 2 * q);
            const double2 Nt = lds2t(nt + k * 512);
            const double u = reduce_q(fma(Nt.x, s.x - yk.x, Nt.y * (s.y - yk.y)));
            __syncwarp();
            if (q == 0) c.ys[8 * k + g] = u;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads, 1)
resolve_qp_kernel(const hdrt_resolve_problem p, double* work, long long work_stride) {
    RCtx c;
    c.nr = p.nr; c.nc = p.nc; c.n = p.nr * p.nc; c.T = (c.n + 7) >> 3;
    c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5; c.g = c.lane >> 2; c.q = c.lane & 3;
    c.tiles = work + (size_t)blockIdx.x * work_stride;
    const int T = c.T, n = c.n, tid = threadIdx.x;
    c.bs = g_smem; c.ys = c.bs + 8 * T; c.dsq = c.ys + 8 * T; c.zrow = c.dsq + 8 * T; c.ninv = c.zrow + 64 * T;
    c.part = c.ninv + 64 * T; c.red = c.part + 16 * kWarps;
    c.red_phase = 0;
    for (int win = blockIdx.x; win < p.n_windows; win += gridDim.x) {
        c.p = p.p + (size_t)p.first_obs[win] * p.nc * p.nc;
        c.my = p.my + (size_t)win * p.nr * p.nr;
        c.pscale = p.param_scale + (size_t)win * p.nc;
        const double* qv = p.q + (size_t)p.first_obs[win] * p.nc;
        for (int i = tid; i < 8 * T; i += kThreads) { c.bs[i] = 0.0; c.ys[i] = 0.0; c.dsq[i] = 0.0; }
        __syncthreads();
        bool act[EU];
        double qi[EU], hi[EU];
#pragma unroll
        for (int u = 0; u < EU; ++u) {
            const int e = tid + kThreads * u;
            act[u] = e < n;
            qi[u] = act[u] ? qv[e] : 0.0;              // the observations of a window are consecutive: q is contiguous
            hi[u] = act[u] ? p.h[e % p.nc] : 0.0;
        }
        double resx0, resz0;
        {
            double t2[2] = {0.0, 0.0};
#pragma unroll
            for (int u = 0; u < EU; ++u) { t2[0] += qi[u] * qi[u]; t2[1] += hi[u] * hi[u]; }
            rblock_reduce<2, 0u>(t2, c);
            resx0 = fmax(1.0, sqrt(t2[0]));
            resz0 = fmax(1.0, sqrt(t2[1]));
        }
        double xi[EU], si[EU], zi[EU], di[EU], dinv[EU], lam[EU], rxi[EU], rzi[EU], pxi[EU];
#pragma unroll
        for (int u = 0; u < EU; ++u) {
            xi[u] = 0.0; si[u] = 1.0; zi[u] = 1.0; di[u] = 1.0; dinv[u] = 1.0; lam[u] = 1.0; rxi[u] = 0.0; rzi[u] = 0.0; pxi[u] = 0.0;
        }
        double gap = 0.0;
        int status = 0, iters;
#pragma unroll 1
        for (iters = -1; iters <= kMaxIpm; ++iters) {
            if (iters >= 0) {
                double t5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int u = 0; u < EU; ++u) {
                    rxi[u] = pxi[u] + qi[u];
                    const double f0p = act[u] ? (xi[u] * rxi[u] + xi[u] * qi[u]) : 0.0;
                    rxi[u] -= zi[u];
                    rzi[u] = si[u] - hi[u] - xi[u];
                    if (act[u]) {
                        t5[0] += f0p; t5[1] += rxi[u] * rxi[u]; t5[2] += rzi[u] * rzi[u]; t5[3] += zi[u] * rzi[u];
                        t5[4] += (iters == 0) ? si[u] * zi[u] : lam[u] * lam[u];
                    }
                }
                rblock_reduce<5, 0u>(t5, c);
                const double f0 = 0.5 * t5[0], resx = sqrt(t5[1]), resz = sqrt(t5[2]);
                gap = t5[4];
                const double pcost = f0, dcost = f0 + t5[3] - gap;
                double relgap = 0.0;
                bool have_rel = true;
                if (pcost < 0.0) relgap = gap / -pcost;
                else if (dcost > 0.0) relgap = gap / dcost;
                else have_rel = false;
                const bool done = (resz / resz0 <= kFeasTol) && (resx / resx0 <= kFeasTol) &&
                                  ((gap <= kAbsTol) || (have_rel && relgap <= kRelTol));
                if (done) break;
                if (iters == kMaxIpm) { status |= HDRT_ST_QP_MAXITERS; break; }
                if (iters == 0) {
#pragma unroll
                    for (int u = 0; u < EU; ++u) { di[u] = sqrt(si[u] / zi[u]); dinv[u] = 1.0 / di[u]; lam[u] = sqrt(si[u] * zi[u]); }
                }
            }
#pragma unroll
            for (int u = 0; u < EU; ++u) if (act[u]) c.dsq[tid + kThreads * u] = dinv[u] * dinv[u];
            __syncthreads();
            if (!rfactor(c)) {
                status |= HDRT_ST_KKT_FAIL;
                if (iters <= 0) {
#pragma unroll
                    for (int u = 0; u < EU; ++u) xi[u] = nan("");
                }
                break;
            }
            const bool start = iters < 0;
            const double mu = gap / (double)n;
            double sigma = 0.0, step = 1.0;
            double ws3[EU], dxi[EU], dsi[EU], dzi[EU], zs[EU], rhs[EU];
#pragma unroll
            for (int u = 0; u < EU; ++u) { ws3[u] = 0.0; dxi[u] = 0.0; dsi[u] = 0.0; dzi[u] = 0.0; zs[u] = 0.0; rhs[u] = 0.0; }
#pragma unroll 1
            for (int pass = start ? 1 : 0; pass < 2; ++pass) {
#pragma unroll
                for (int u = 0; u < EU; ++u) {
                    if (start) {
                        rhs[u] = -qi[u] - hi[u];
                    } else {
                        double ds = 0.0;
                        if (pass == 1) ds -= ws3[u];
                        ds -= lam[u] * lam[u];
                        ds += sigma * mu;
                        dxi[u] = -rxi[u];
                        double dz = -rzi[u];
                        ds = ds / lam[u];
                        dz = dz - di[u] * ds;
                        zs[u] = dinv[u] * dz;
                        rhs[u] = dxi[u] - dinv[u] * zs[u];
                        dsi[u] = ds;
                    }
                    if (act[u]) c.bs[tid + kThreads * u] = rhs[u];
                }
                __syncthreads();
                rsolve(c);
#pragma unroll
                for (int u = 0; u < EU; ++u) dxi[u] = act[u] ? c.ys[tid + kThreads * u] : 0.0;
                __syncthreads();
                if (start) {
                    double t4[4] = {0.0, -INFINITY, 0.0, -INFINITY};
#pragma unroll
                    for (int u = 0; u < EU; ++u) {
                        xi[u] = dxi[u];
                        pxi[u] = act[u] ? rhs[u] - xi[u] : 0.0;
                        zi[u] = -xi[u] - hi[u];
                        si[u] = -zi[u];
                        if (act[u]) {
                            t4[0] += si[u] * si[u]; t4[1] = fmax(t4[1], -si[u]);
                            t4[2] += zi[u] * zi[u]; t4[3] = fmax(t4[3], -zi[u]);
                        }
                    }
                    rblock_reduce<4, 0xAu>(t4, c);
                    const double nrms = sqrt(t4[0]), ts = t4[1], nrmz = sqrt(t4[2]), tz = t4[3];
#pragma unroll
                    for (int u = 0; u < EU; ++u) {
                        if (ts >= -1e-8 * fmax(nrms, 1.0)) si[u] += 1.0 + ts;
                        if (tz >= -1e-8 * fmax(nrmz, 1.0)) zi[u] += 1.0 + tz;
                    }
                    break;
                }
                double t3[3] = {0.0, -INFINITY, -INFINITY};
#pragma unroll
                for (int u = 0; u < EU; ++u) {
                    double dz = -dinv[u] * dxi[u] - zs[u];
                    double ds = dsi[u] - dz;
                    const double prod = ds * dz;
                    if (pass == 0) ws3[u] = prod;
                    ds = ds / lam[u];
                    dz = dz / lam[u];
                    dsi[u] = ds;
                    dzi[u] = dz;
                    if (act[u]) { t3[0] += prod; t3[1] = fmax(t3[1], -ds); t3[2] = fmax(t3[2], -dz); }
                }
                rblock_reduce<3, 0x6u>(t3, c);
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
#pragma unroll
            for (int u = 0; u < EU; ++u) {
                pxi[u] = act[u] ? fma(step, rhs[u] - (dinv[u] * dinv[u]) * dxi[u], pxi[u]) : 0.0;
                xi[u] += step * dxi[u];
                double ds = step * dsi[u] + 1.0, dz = step * dzi[u] + 1.0;
                ds *= lam[u];
                dz *= lam[u];
                const double sqs = sqrt(ds), sqz = sqrt(dz);
                di[u] = di[u] * sqs / sqz;
                dinv[u] = 1.0 / di[u];
                lam[u] = sqs * sqz;
                si[u] = lam[u] * di[u];
                zi[u] = lam[u] * dinv[u];
            }
        }
#pragma unroll
        for (int u = 0; u < EU; ++u) if (act[u]) p.x[(size_t)win * n + tid + kThreads * u] = xi[u];
        if (tid == 0) {
            if (p.iters) p.iters[win] = iters < 0 ? 0 : iters;
            if (p.status) p.status[win] = status;
        }
        __syncthreads();
    }
}

}  // namespace rs
}  // namespace hdrt
