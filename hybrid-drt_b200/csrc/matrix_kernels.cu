// Response-matrix builders (reference: hybdrt/matrices/{basis,mat1d,phasance}.py).
//
// All outputs are written with coalesced stores, one CTA per (grid, row-tile); the frequency / time
// and tau vectors of the grid are staged in shared memory once per CTA.  Interp-mode kernels are
// HBM-write bound (2 table loads + ~10 flops per 16 bytes stored); trapz-mode kernels are
// exp-throughput bound (1000 integrand evaluations per entry, one warp per entry).
#include "common.cuh"

namespace hdrt {

constexpr int kMThreads = 256;

// numpy.linspace(start, stop, num)[i]
__device__ __forceinline__ double linspace_at(double start, double stop, int num, int i) {
    if (i == num - 1) return stop;
    const double step = (stop - start) / (double)(num - 1);
    return __dadd_rn(__dmul_rn((double)i, step), start);
}

// Gaussian RBF, basis.py:93-95
__device__ __forceinline__ double rbf(double y, double eps) {
    const double t = eps * y;
    return exp(-(t * t));
}

// integrands, basis.py:562-570 and :616-618.  kind 0 = Re z, 1 = Im z, 2 = step response
__device__ __forceinline__ double integrand(int kind, double y, double arg, double eps) {
    if (kind == 0) return rbf(y, eps) / (1.0 + exp(2.0 * (y + arg)));
    if (kind == 1) return -rbf(y, eps) * exp(y) * exp(arg) / (1.0 + exp(2.0 * (y + arg)));
    return rbf(y, eps) * (1.0 - exp(-arg / exp(y)));
}

// np.trapezoid(f(y), x=y) over y = linspace(-20, 20, quad_points); one warp cooperates.
// For kind 0/1 `arg` is ln(omega tau); for kind 2 it is dt/tau.
// sum_i (y_{i+1} - y_i) (f_{i+1} + f_i) / 2 regrouped by point -- f_i (y_{i+1} - y_{i-1}) / 2 with one-sided weights
// at the two ends -- so that every integrand value (three exponentials) is computed once, not twice.
__device__ double warp_trapz(int kind, double arg, double eps, int quad_points) {
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int i = lane; i < quad_points; i += 32) {
        const double y = linspace_at(-20.0, 20.0, quad_points, i);
        const double ym = linspace_at(-20.0, 20.0, quad_points, max(i - 1, 0));
        const double yp = linspace_at(-20.0, 20.0, quad_points, min(i + 1, quad_points - 1));
        acc = fma(integrand(kind, y, arg, eps), (yp - ym) / 2.0, acc);
    }
    return warp_sum(acc);
}

// numpy.interp(x, gx, gv) with edge clamping (numpy/_core/src/multiarray/compiled_base.c)
// The index guess (the tables are uniform up to rounding) uses a precomputed inverse spacing; the two loops then
// move it to exactly where numpy's binary search lands.
struct InterpGrid {
    double x0, xn, inv_dx;
};
__device__ __forceinline__ InterpGrid interp_grid(const double* __restrict__ gx, int npts) {
    InterpGrid g;
    g.x0 = __ldg(gx);
    g.xn = __ldg(gx + npts - 1);
    g.inv_dx = (double)(npts - 1) / (g.xn - g.x0);
    return g;
}
__device__ __forceinline__ double interp_clamped(double x, const double* __restrict__ gx,
                                                 const double* __restrict__ gv, int npts, const InterpGrid& ig) {
    if (isnan(x)) return x;
    const double x0 = ig.x0, xn = ig.xn;
    if (x > xn) return __ldg(gv + npts - 1);
    if (x < x0) return __ldg(gv);
    int j = (int)((x - x0) * ig.inv_dx);
    j = max(0, min(npts - 1, j));
    while (j > 0 && __ldg(gx + j) > x) --j;
    while (j < npts - 1 && __ldg(gx + j + 1) <= x) ++j;
    if (j == npts - 1) return __ldg(gv + j);
    const double xj = __ldg(gx + j), vj = __ldg(gv + j);
    if (xj == x) return vj;
    const double slope = (__ldg(gv + j + 1) - vj) / (__ldg(gx + j + 1) - xj);
    return __dadd_rn(__dmul_rn(slope, x - xj), vj);
}

// ---- lookup tables (basis.py:648-689): one warp per table entry ---------------------------------
__global__ void lookup_kernel(double eps, int grid_points, int quad_points, double* re_x, double* re_v,
                              double* im_x, double* im_v, double* td_x, double* td_v) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= 3 * grid_points) return;
    const int table = warp / grid_points, i = warp - table * grid_points;
    double lo, hi;
    if (table == 0) { lo = -2.7; hi = 2.7; }
    else if (table == 1) { lo = -5.4; hi = 5.4; }   // im_lim = re_lim * 2
    else { lo = -6.0; hi = 2.0; }
    const double g = pow(10.0, linspace_at(lo, hi, grid_points, i));  // np.logspace
    const double lg = log(g);
    const double val = warp_trapz(table, table == 2 ? g : lg, eps, quad_points);
    if (lane == 0) {
        double* gx = table == 0 ? re_x : (table == 1 ? im_x : td_x);
        double* gv = table == 0 ? re_v : (table == 1 ? im_v : td_v);
        gx[i] = lg;
        gv[i] = val;
    }
}

// ---- impedance matrix (mat1d.py:212-374) ---------------------------------------------------------
// Generic path (lookup tables too large for shared memory): grid.x = n_grids, grid.y = row tiles, tables read
// through the read-only path.
__global__ void impedance_interp_global_kernel(const double* __restrict__ freq, const double* __restrict__ tau, int nf,
                                               int nb, const double* __restrict__ re_x, const double* __restrict__ re_v,
                                               const double* __restrict__ im_x, const double* __restrict__ im_v,
                                               int npts, double* __restrict__ a_re, double* __restrict__ a_im,
                                               int rows_per_cta) {
    extern __shared__ double sm[];
    double* s_tau = sm;
    double* s_om = sm + nb;
    const int g = blockIdx.x;
    const int r0 = blockIdx.y * rows_per_cta;
    const int rows = min(rows_per_cta, nf - r0);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_tau[i] = log(tau[(size_t)g * nb + i]);
    for (int i = threadIdx.x; i < rows; i += blockDim.x)
        s_om[i] = log(freq[(size_t)g * nf + r0 + i] * 2.0 * 3.141592653589793);  // frequencies * 2 * np.pi
    __syncthreads();
    const size_t base = ((size_t)g * nf + r0) * nb;
    const InterpGrid gre = interp_grid(re_x, npts), gim = interp_grid(im_x, npts);
    for (int idx = threadIdx.x; idx < rows * nb; idx += blockDim.x) {
        const int rr = idx / nb, m = idx - rr * nb;
        const double x = s_om[rr] + s_tau[m];
        a_re[base + idx] = interp_clamped(x, re_x, re_v, npts, gre);
        a_im[base + idx] = interp_clamped(x, im_x, im_v, npts, gim);
    }
}

// Production path: persistent CTAs (two per SM) keep both lookup tables in shared memory together with the
// interval slopes (computed once per CTA with the same division numpy.interp does per evaluation), so an entry
// costs ln(omega_n) + ln(tau_m) (logs once per row / column of the work item), one index guess and a handful of
// shared-memory loads -- no log, no division: the kernel is bound by the HBM writes.  ln(omega tau) is formed as
// ln(omega) + ln(tau); the reference itself evaluates entries either way (full evaluation, or first row / column
// + Toeplitz fill, mat1d.py:341-372), the two differ at the 1e-16 level.
// A work item is `rows_per_item` rows of one grid; stores are 8-byte, streaming (the output is not re-read): consecutive
// threads write consecutive columns, a warp instruction covers 256 contiguous bytes.
struct SmemTable {
    const double* x;      // knots
    const double2* sc;    // per interval: slope, intercept (v_j - slope x_j): the interpolant is fma(slope, x, intercept)
    double x0, xn, inv_dx, pos0, v0, vn;
    bool uniform;   // every knot within 1e-9 of a spacing of the uniform grid: the index guess can be trusted
};
// numpy.interp semantics: piecewise linear between the knots, the edge values outside the span.
// Uniform table (the lookup tables are linspace grids): x is clamped to the span, the interval index is the floor of
// (x - x0) / dx taken by adding 2^52 + 2^51 (the low word of the sum is the rounded integer: one DADD instead of a
// double -> int conversion on the quarter-rate pipe), and the value is one fma on a 16-byte shared load.  At a knot
// the guess may fall on either neighbouring interval: the interpolant is continuous, both give the knot value to the
// last bit or two.  Against slope * (x - x_j) + v_j the fma form differs at the 1e-16 level of the table scale.
// A non-uniform table takes the search loop.
template <bool UNIFORM, bool CHECK = true>
__device__ __forceinline__ double interp_smem(double x, const SmemTable& t, int npts) {
    if (UNIFORM && !CHECK) {      // the caller has bounded |x| for the whole work item: no branch at all
        const double big = fma(x, t.inv_dx, t.pos0) + 6755399441055744.0;
        const int j = max(-1, min(__double2loint(big), npts - 1));
        const double2 sc = t.sc[j + 1];
        return fma(sc.x, x, sc.y);
    }
    if (UNIFORM) {
        // sc[0] and sc[npts] are flat sentinel intervals (slope 0, the edge value): an x outside the span needs no
        // comparison, its index clamps onto them.  fmin / fmax on doubles cost eight instructions each on this target.
        const double pos = fma(x, t.inv_dx, t.pos0);                      // (x - x0) / dx - 0.5
        const double big = pos + 6755399441055744.0;                      // 2^52 + 2^51: low word = round(pos)
        if ((unsigned)(__double2hiint(pos) & 0x7fffffff) < 0x41d00000u) { // |pos| < 2^30 (always, for finite frequencies)
            const int j = max(-1, min(__double2loint(big), npts - 1));
            const double2 sc = t.sc[j + 1];
            return fma(sc.x, x, sc.y);
        }
        return (x != x) ? x : (x < t.x0 ? t.v0 : t.vn);                   // +-inf (f = 0), NaN
    }
    int j = 0;
    {
        int lo = 0, hi = npts - 1;                            // binary search: x[j] <= x < x[j + 1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (t.x[mid] <= x) lo = mid; else hi = mid; }
        j = lo;
    }
    double r = fma(t.sc[j + 1].x, x, t.sc[j + 1].y);
    r = (x >= t.xn) ? t.vn : r;
    r = (x < t.x0) ? t.v0 : r;
    return r;
}

constexpr int kIThreads = 512;

// The entries of one work item that belong to this thread: fixed column, every `groups`-th row; four rows in flight.
template <bool UNIFORM, bool CHECK>
__device__ __forceinline__ void interp_tile(const double* s_lt, const double* s_lw, double* __restrict__ o_re,
                                            double* __restrict__ o_im, const SmemTable& tr, const SmemTable& ti, int npts,
                                            int nb, int rows, int col0, int rg, int groups) {
    const size_t step = (size_t)groups * nb;
    for (int col = col0; col < nb; col += kIThreads) {
        const double lt = s_lt[col];
        double* __restrict__ pr = o_re + (size_t)rg * nb + col;
        double* __restrict__ pi = o_im + (size_t)rg * nb + col;
        int rr = rg;
        for (; rr + 3 * groups < rows; rr += 4 * groups, pr += 4 * step, pi += 4 * step) {
            double vr[4], vi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double x = s_lw[rr + u * groups] + lt;
                vr[u] = interp_smem<UNIFORM, CHECK>(x, tr, npts);
                vi[u] = interp_smem<UNIFORM, CHECK>(x, ti, npts);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                __stcs(pr + u * step, vr[u]);
                __stcs(pi + u * step, vi[u]);
            }
        }
        for (; rr < rows; rr += groups, pr += step, pi += step) {
            const double x = s_lw[rr] + lt;
            __stcs(pr, interp_smem<UNIFORM, CHECK>(x, tr, npts));
            __stcs(pi, interp_smem<UNIFORM, CHECK>(x, ti, npts));
        }
    }
}

__global__ void __launch_bounds__(kIThreads, 2)
impedance_interp_kernel(const double* __restrict__ freq, const double* __restrict__ tau, int n_grids, int nf, int nb,
                        const double* __restrict__ re_x, const double* __restrict__ re_v,
                        const double* __restrict__ im_x, const double* __restrict__ im_v, int npts,
                        double* __restrict__ a_re, double* __restrict__ a_im, int rows_per_item, int tiles_per_grid) {
    extern __shared__ __align__(16) double sm[];
    double2* t_rsc = reinterpret_cast<double2*>(sm);          // [npts + 1] (slope, intercept), real table; see interp_smem
    double2* t_isc = t_rsc + npts + 1;                        // imaginary table
    double* t_rx = reinterpret_cast<double*>(t_isc + npts + 1);
    double* t_ix = t_rx + npts;
    double* s_vec = t_ix + npts;                              // two buffers of [nb ln tau | rows_per_item ln omega]
    const int vlen = nb + rows_per_item;
    const int tid = threadIdx.x;
    for (int i = tid; i < npts; i += kIThreads) { t_rx[i] = __ldg(re_x + i); t_ix[i] = __ldg(im_x + i); }
    for (int i = tid; i < npts - 1; i += kIThreads) {         // interval i = [x_i, x_i+1) sits at index i + 1
        const double sr = (__ldg(re_v + i + 1) - __ldg(re_v + i)) / (__ldg(re_x + i + 1) - __ldg(re_x + i));
        const double si = (__ldg(im_v + i + 1) - __ldg(im_v + i)) / (__ldg(im_x + i + 1) - __ldg(im_x + i));
        t_rsc[i + 1] = make_double2(sr, fma(-sr, __ldg(re_x + i), __ldg(re_v + i)));
        t_isc[i + 1] = make_double2(si, fma(-si, __ldg(im_x + i), __ldg(im_v + i)));
    }
    if (tid == 0) {
        t_rsc[0] = make_double2(0.0, __ldg(re_v)); t_rsc[npts] = make_double2(0.0, __ldg(re_v + npts - 1));
        t_isc[0] = make_double2(0.0, __ldg(im_v)); t_isc[npts] = make_double2(0.0, __ldg(im_v + npts - 1));
    }
    __shared__ int s_nonuniform[2];
    if (tid < 2) s_nonuniform[tid] = 0;
    SmemTable tr, ti;
    tr.x = t_rx; tr.sc = t_rsc; tr.x0 = __ldg(re_x); tr.xn = __ldg(re_x + npts - 1); tr.v0 = __ldg(re_v); tr.vn = __ldg(re_v + npts - 1);
    ti.x = t_ix; ti.sc = t_isc; ti.x0 = __ldg(im_x); ti.xn = __ldg(im_x + npts - 1); ti.v0 = __ldg(im_v); ti.vn = __ldg(im_v + npts - 1);
    tr.inv_dx = (double)(npts - 1) / (tr.xn - tr.x0);
    ti.inv_dx = (double)(npts - 1) / (ti.xn - ti.x0);
    tr.pos0 = -tr.x0 * tr.inv_dx - 0.5;
    ti.pos0 = -ti.x0 * ti.inv_dx - 0.5;
    __syncthreads();
    for (int i = tid; i < npts; i += kIThreads) {
        if (fabs((t_rx[i] - tr.x0) * tr.inv_dx - (double)i) > 1e-9) s_nonuniform[0] = 1;
        if (fabs((t_ix[i] - ti.x0) * ti.inv_dx - (double)i) > 1e-9) s_nonuniform[1] = 1;
    }
    __syncthreads();
    tr.uniform = s_nonuniform[0] == 0;
    ti.uniform = s_nonuniform[1] == 0;
    const bool fast = tr.uniform && ti.uniform;
    // |x| = |ln tau + ln omega| <= 2 half_xmax keeps |pos| below 2^30 for both tables
    const double half_xmax = 0.5 * fmin((1073741824.0 - fabs(tr.pos0)) / tr.inv_dx, (1073741824.0 - fabs(ti.pos0)) / ti.inv_dx);
    // thread -> (column, row group): the column of a thread is fixed, so ln(tau) sits in a register and the row loop
    // carries no index arithmetic; consecutive threads write consecutive columns (coalesced 8-byte stores)
    const int groups = nb <= kIThreads ? kIThreads / nb : 1;
    const int col0 = nb <= kIThreads ? tid % nb : tid;
    const int rg = nb <= kIThreads ? tid / nb : 0;
    const bool active = rg < groups;
    const long long items = (long long)n_grids * tiles_per_grid;
    // ln(tau) / ln(omega) of the NEXT work item are fetched while the current one is computed (one element per thread,
    // nb + rows_per_item <= kIThreads for every shape the host sends here), so that an item starts without a global
    // round trip in front of it
    auto fetch = [&](long long item) -> double {
        if (item >= items || tid >= vlen) return 1.0;
        const int g = (int)(item / tiles_per_grid);
        const int r0 = (int)(item - (long long)g * tiles_per_grid) * rows_per_item;
        if (tid < nb) return tau[(size_t)g * nb + tid];
        const int rr = r0 + tid - nb;
        return rr < nf ? freq[(size_t)g * nf + rr] * (2.0 * 3.141592653589793) : 1.0;     // 2 pi f
    };
    const bool prefetch = vlen <= kIThreads;
    int buf = 0;
    if (prefetch) {
        const double v0 = fetch(blockIdx.x);
        if (tid < vlen) s_vec[tid] = log(v0);
    }
    for (long long item = blockIdx.x; item < items; item += gridDim.x, buf ^= 1) {
        const int g = (int)(item / tiles_per_grid);
        const int r0 = (int)(item - (long long)g * tiles_per_grid) * rows_per_item;
        const int rows = min(rows_per_item, nf - r0);
        double* s_lt = s_vec + buf * vlen;
        double* s_lw = s_lt + nb;
        double nextv = 1.0;
        if (prefetch) {
            nextv = fetch(item + gridDim.x);
        } else {
            for (int i = tid; i < nb + rows; i += kIThreads) {
                if (i < nb) s_lt[i] = log(tau[(size_t)g * nb + i]);
                else s_lw[i - nb] = log(freq[(size_t)g * nf + r0 + (i - nb)] * 2.0 * 3.141592653589793);
            }
        }
        // the barrier also orders the tables on the first pass (the other buffer is free by now), and tells whether every
        // ln(tau) / ln(omega) of this item is small enough for the index arithmetic of the branch-free path
        bool mine_ok = true;
        if (prefetch && tid < vlen) mine_ok = fabs(s_vec[buf * vlen + tid]) < half_xmax;
        const bool bounded = __syncthreads_and(mine_ok) && prefetch;
        double* __restrict__ o_re = a_re + ((size_t)g * nf + r0) * nb;
        double* __restrict__ o_im = a_im + ((size_t)g * nf + r0) * nb;
        if (active) {
            if (fast && bounded) interp_tile<true, false>(s_lt, s_lw, o_re, o_im, tr, ti, npts, nb, rows, col0, rg, groups);
            else if (fast) interp_tile<true, true>(s_lt, s_lw, o_re, o_im, tr, ti, npts, nb, rows, col0, rg, groups);
            else interp_tile<false, true>(s_lt, s_lw, o_re, o_im, tr, ti, npts, nb, rows, col0, rg, groups);
        }
        if (prefetch && tid < vlen) s_vec[(buf ^ 1) * vlen + tid] = log(nextv);    // published by the next barrier
    }
}

// Step-response matrix, interp mode, production path (mat1d.py:16-122): the same persistent scheme as the impedance
// builder.  ln((t - t_k) / tau_m) is formed as ln(t - t_k) - ln(tau_m) (logs once per row and step / per column, against
// one division and one log per entry and step), the lookup is the (slope, intercept) table in shared memory.  A work
// item is `rows_per_item` rows of one grid.  Needs a uniform lookup table and bounded arguments (else the generic
// kernel below runs).
__global__ void __launch_bounds__(kIThreads, 2)
response_interp_fast_kernel(const double* __restrict__ times, const double* __restrict__ tau,
                            const double* __restrict__ step_times, const double* __restrict__ step_sizes, int n_grids,
                            int nt, int nb, int n_steps, const double* __restrict__ td_x, const double* __restrict__ td_v,
                            int npts, double* __restrict__ rm, int rows_per_item, int tiles_per_grid, int* __restrict__ fallback) {
    extern __shared__ __align__(16) double sm[];
    double2* t_sc = reinterpret_cast<double2*>(sm);                  // [npts + 1], see interp_smem
    double* s_ltau = reinterpret_cast<double*>(t_sc + npts + 1);     // [nb] ln tau
    double* s_lt = s_ltau + nb;                                      // [rows_per_item][n_steps] ln(t - t_k), NaN where t <= t_k
    double* s_sa = s_lt + (size_t)rows_per_item * n_steps;           // [n_steps]
    const int tid = threadIdx.x;
    for (int i = tid; i < npts - 1; i += kIThreads) {
        const double sl = (__ldg(td_v + i + 1) - __ldg(td_v + i)) / (__ldg(td_x + i + 1) - __ldg(td_x + i));
        t_sc[i + 1] = make_double2(sl, fma(-sl, __ldg(td_x + i), __ldg(td_v + i)));
    }
    if (tid == 0) { t_sc[0] = make_double2(0.0, __ldg(td_v)); t_sc[npts] = make_double2(0.0, __ldg(td_v + npts - 1)); }
    SmemTable tt;
    tt.x = nullptr; tt.sc = t_sc; tt.x0 = __ldg(td_x); tt.xn = __ldg(td_x + npts - 1); tt.v0 = __ldg(td_v); tt.vn = __ldg(td_v + npts - 1);
    tt.inv_dx = (double)(npts - 1) / (tt.xn - tt.x0);
    tt.pos0 = -tt.x0 * tt.inv_dx - 0.5;
    tt.uniform = true;
    int bad = 0;
    for (int i = tid; i < npts; i += kIThreads)
        if (fabs((__ldg(td_x + i) - tt.x0) * tt.inv_dx - (double)i) > 1e-9) bad = 1;
    const double half_xmax = 0.5 * (1073741824.0 - fabs(tt.pos0)) / tt.inv_dx;
    const int groups = nb <= kIThreads ? kIThreads / nb : 1;
    const int col0 = nb <= kIThreads ? tid % nb : tid;
    const int rg = nb <= kIThreads ? tid / nb : 0;
    const long long items = (long long)n_grids * tiles_per_grid;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int g = (int)(item / tiles_per_grid);
        const int r0 = (int)(item - (long long)g * tiles_per_grid) * rows_per_item;
        const int rows = min(rows_per_item, nt - r0);
        __syncthreads();                                              // the previous item is done with the staging vectors
        for (int i = tid; i < nb; i += kIThreads) {
            const double v = log(tau[(size_t)g * nb + i]);
            s_ltau[i] = v;
            if (!(fabs(v) < half_xmax)) bad = 1;
        }
        for (int i = tid; i < n_steps; i += kIThreads) s_sa[i] = step_sizes[(size_t)g * n_steps + i];
        for (int i = tid; i < rows * n_steps; i += kIThreads) {
            const int rr = i / n_steps, k = i - rr * n_steps;
            const double dt = times[(size_t)g * nt + r0 + rr] - step_times[(size_t)g * n_steps + k];
            const double v = dt > 0.0 ? log(dt) : nan("");
            s_lt[i] = v;
            if (dt > 0.0 && !(fabs(v) < half_xmax)) bad = 1;
        }
        if (__syncthreads_or(bad)) {                                  // odd table or unbounded arguments: the generic kernel redoes it all
            if (tid == 0) *fallback = 1;
            return;
        }
        if (rg < groups) {
            for (int col = col0; col < nb; col += kIThreads) {
                const double ltau = s_ltau[col];
                double* __restrict__ po = rm + ((size_t)g * nt + r0 + rg) * nb + col;
                const size_t step = (size_t)groups * nb;
                for (int rr = rg; rr < rows; rr += groups, po += step) {
                    double acc = 0.0;
                    for (int k = 0; k < n_steps; ++k) {
                        const double lt = s_lt[rr * n_steps + k];
                        if (lt == lt) acc += __dmul_rn(interp_smem<true, false>(lt - ltau, tt, npts), s_sa[k]);
                    }
                    __stcs(po, acc);
                }
            }
        }
    }
}

// trapz mode: one warp per entry, both parts
__global__ void impedance_trapz_kernel(const double* __restrict__ freq, const double* __restrict__ tau, int n_grids,
                                       int nf, int nb, double eps, int quad_points, double* __restrict__ a_re,
                                       double* __restrict__ a_im) {
    const long long total = (long long)n_grids * nf * nb;
    const int lane = threadIdx.x & 31;
    for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total;
         e += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int m = (int)(e % nb);
        const long long gf = e / nb;
        const int g = (int)(gf / nf);
        const double om = freq[gf] * 2.0 * 3.141592653589793;
        const double x = log(om * tau[(size_t)g * nb + m]);
        const double vr = warp_trapz(0, x, eps, quad_points);
        const double vi = warp_trapz(1, x, eps, quad_points);
        if (lane == 0) { a_re[e] = vr; a_im[e] = vi; }
    }
}

// ---- step-response matrix (mat1d.py:16-122) ------------------------------------------------------
__global__ void response_interp_kernel(const double* __restrict__ times, const double* __restrict__ tau,
                                       const double* __restrict__ step_times, const double* __restrict__ step_sizes,
                                       int nt, int nb, int n_steps, const double* __restrict__ td_x,
                                       const double* __restrict__ td_v, int npts, double* __restrict__ rm,
                                       int rows_per_cta, const int* __restrict__ run_flag) {
    if (run_flag != nullptr && *run_flag == 0) return;     // the fast kernel took the job
    extern __shared__ double sm[];
    double* s_tau = sm;
    double* s_t = sm + nb;
    double* s_st = s_t + rows_per_cta;
    double* s_sa = s_st + n_steps;
    const int g = blockIdx.x;
    const int r0 = blockIdx.y * rows_per_cta;
    const int rows = min(rows_per_cta, nt - r0);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_tau[i] = tau[(size_t)g * nb + i];
    for (int i = threadIdx.x; i < rows; i += blockDim.x) s_t[i] = times[(size_t)g * nt + r0 + i];
    for (int i = threadIdx.x; i < n_steps; i += blockDim.x) {
        s_st[i] = step_times[(size_t)g * n_steps + i];
        s_sa[i] = step_sizes[(size_t)g * n_steps + i];
    }
    __syncthreads();
    const size_t base = ((size_t)g * nt + r0) * nb;
    const InterpGrid gtd = interp_grid(td_x, npts);
    for (int idx = threadIdx.x; idx < rows * nb; idx += blockDim.x) {
        const int rr = idx / nb, m = idx - rr * nb;
        const double t = s_t[rr];
        double acc = 0.0;
        for (int k = 0; k < n_steps; ++k) {
            if (t > s_st[k]) {
                const double x = log((t - s_st[k]) / s_tau[m]);
                acc += __dmul_rn(interp_clamped(x, td_x, td_v, npts, gtd), s_sa[k]);
            }
        }
        rm[base + idx] = acc;
    }
}

__global__ void response_trapz_kernel(const double* __restrict__ times, const double* __restrict__ tau,
                                      const double* __restrict__ step_times, const double* __restrict__ step_sizes,
                                      int n_grids, int nt, int nb, int n_steps, double eps, int quad_points,
                                      double* __restrict__ rm) {
    const long long total = (long long)n_grids * nt * nb;
    const int lane = threadIdx.x & 31;
    for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total;
         e += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int m = (int)(e % nb);
        const long long gt = e / nb;
        const int g = (int)(gt / nt);
        const double t = times[gt];
        const double tm = tau[(size_t)g * nb + m];
        double acc = 0.0;
        for (int k = 0; k < n_steps; ++k) {
            const double st = step_times[(size_t)g * n_steps + k];
            if (t > st) acc += __dmul_rn(warp_trapz(2, (t - st) / tm, eps, quad_points), step_sizes[(size_t)g * n_steps + k]);
        }
        if (lane == 0) rm[e] = acc;
    }
}

// ---- derivative penalty matrices (mat1d.py:125-209, basis.py:382-395) ----------------------------
__global__ void penalty_kernel(const double* __restrict__ grid, int n_grids, int nb, double eps, int toeplitz,
                               double* __restrict__ m) {
    // one thread per (grid, i, j): the three orders share the Gaussian factor; 32-bit index arithmetic (the round-1 form
    // spent most of its instructions in 64-bit divisions and evaluated the exponential once per order)
    const unsigned per = (unsigned)nb * (unsigned)nb;
    const double c = sqrt(3.141592653589793 / 2.0);
    const double c0 = c * (1.0 / eps), c1 = -c * eps, c2 = c * (eps * eps * eps);
    for (int g = blockIdx.y; g < n_grids; g += gridDim.y) {
        const double* x = grid + (size_t)g * nb;
        double* out = m + (size_t)g * 3 * per;
        for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < per; e += gridDim.x * blockDim.x) {
            const unsigned i = e / (unsigned)nb, j = e - i * (unsigned)nb;
            double a;
            if (toeplitz) {
                const unsigned d = i > j ? i - j : j - i;
                a = eps * (x[0] - x[d]);
            } else {
                a = eps * (x[j] - x[i]);
            }
            const double a2 = a * a;
            const double ex = exp(-(a2 / 2.0));
            __stcs(out + e, c0 * ex);
            __stcs(out + per + e, c1 * (-1.0 + a2) * ex);
            __stcs(out + 2 * (size_t)per + e, c2 * (3.0 - 6.0 * a2 + a2 * a2) * ex);
        }
    }
}

// ---- EIS variance-estimation matrix (mat1d.py:493-515): one warp per row --------------------------
__global__ void eis_vmm_kernel(const double* __restrict__ freq, int n_grids, int nf, double vmm_eps, double reim_cor,
                               int uniform, double* __restrict__ vmm) {
    const int n2 = 2 * nf;
    const long long rows = (long long)n_grids * n2;
    const int lane = threadIdx.x & 31;
    for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows;
         row += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int g = (int)(row / n2), i = (int)(row % n2);
        const double* f = freq + (size_t)g * nf;
        const double lfi = log(f[i % nf]);
        double sum = 0.0;
        for (int j = lane; j < n2; j += 32) {
            double v = uniform ? 1.0 : rbf(lfi - log(f[j % nf]), vmm_eps);
            if ((i < nf) != (j < nf)) v *= reim_cor;
            sum += v;
        }
        sum = warp_sum(sum);
        for (int j = lane; j < n2; j += 32) {
            double v = uniform ? 1.0 : rbf(lfi - log(f[j % nf]), vmm_eps);
            if ((i < nf) != (j < nf)) v *= reim_cor;
            vmm[row * n2 + j] = v / sum;
        }
    }
}

// ---- chrono variance-estimation matrix (mat1d.py:455-490, utils/chrono.py:5-39) ----------------------
// error_structure=None: Gaussian RBF in transformed time (log time since the last step, segments laid end to
// end), no correlation between steps, rows normalised.  One CTA per (grid, block of rows); every CTA rebuilds
// the transformed-time vector of its grid in shared memory (n_t logs against rows x n_t exps of real work).
constexpr int kVRows = 32;   // rows per CTA, four per warp

__global__ void chrono_vmm_kernel(const double* __restrict__ times, const double* __restrict__ step_times, int nt,
                                  int n_steps, double vmm_eps, int uniform, double* __restrict__ vmm, int staged) {
    extern __shared__ __align__(16) double sm[];
    double* s_tt = sm;                                       // [nt] transformed times
    double* s_off = sm + nt;                                 // [n_steps] segment offsets
    int* s_seg = reinterpret_cast<int*>(s_off + n_steps);    // [nt] segment index (0 = before the first step)
    __shared__ double s_red[kMThreads / 32];
    const int g = blockIdx.x, r0 = blockIdx.y * kVRows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* t = times + (size_t)g * nt;
    const double* st = step_times + (size_t)g * n_steps;
    double* out = vmm + (size_t)g * nt * nt;
    if (uniform) {
        for (int rr = warp; rr < kVRows && r0 + rr < nt; rr += kMThreads / 32)
            for (int j = lane; j < nt; j += 32) out[(size_t)(r0 + rr) * nt + j] = 1.0 / (double)nt;
        return;
    }
    // t_sample = min(diff(times))
    double mn = INFINITY;
    for (int i = tid; i + 1 < nt; i += kMThreads) mn = fmin(mn, t[i + 1] - t[i]);
    mn = -warp_max(-mn);
    if (lane == 0) s_red[warp] = mn;
    __syncthreads();
    double t_sample = s_red[0];
    for (int w = 1; w < kMThreads / 32; ++w) t_sample = fmin(t_sample, s_red[w]);
    const double trans_base = log(t_sample / 4.0);
    if (tid == 0) {                                          // trans_offsets = [0, cumsum(log(diff(start_times)) - base)]
        double acc = 0.0;
        s_off[0] = 0.0;
        for (int k = 1; k < n_steps; ++k) {
            acc += log(st[k] - st[k - 1]) - trans_base;
            s_off[k] = acc;
        }
    }
    __syncthreads();
    for (int i = tid; i < nt; i += kMThreads) {
        const double ti = t[i];
        int k = -1;
        for (int m = 0; m < n_steps; ++m) if (ti >= st[m]) k = m;   // start_times are ascending
        double tt;
        if (k < 0) {
            tt = ti - st[0];                                 // before the first step: linear
        } else {
            const double td = fmax(ti - st[k], t_sample / 2.0);
            tt = s_off[k] + log(td) - trans_base;
        }
        s_tt[i] = tt;
        s_seg[i] = k + 1;
    }
    __syncthreads();
    for (int rr = warp; rr < kVRows; rr += kMThreads / 32) {
        const int r = r0 + rr;
        if (r >= nt) break;
        const double tr = s_tt[r];
        const int sr = s_seg[r];
        double sum = 0.0;
        if (staged) {
            // the row is evaluated once (one exponential per entry) into this warp's staging row, then normalised on
            // its way out; without the staging space (long traces) it is evaluated twice
            double* s_row = reinterpret_cast<double*>(s_seg + nt + (nt & 1)) + (size_t)warp * nt;
            for (int j = lane; j < nt; j += 32) {
                const double v = (s_seg[j] == sr) ? rbf(tr - s_tt[j], vmm_eps) : 0.0;
                s_row[j] = v;
                sum += v;
            }
            sum = warp_sum(sum);
            for (int j = lane; j < nt; j += 32) __stcs(out + (size_t)r * nt + j, s_row[j] / sum);
            continue;
        }
        for (int j = lane; j < nt; j += 32) sum += (s_seg[j] == sr) ? rbf(tr - s_tt[j], vmm_eps) : 0.0;
        sum = warp_sum(sum);
        for (int j = lane; j < nt; j += 32)
            out[(size_t)r * nt + j] = ((s_seg[j] == sr) ? rbf(tr - s_tt[j], vmm_eps) : 0.0) / sum;
    }
}

// ---- DOP impedance columns (phasance.py:19-37,61-80,108-118) --------------------------------------
struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cexp_(cplx a) { const double e = exp(a.re); double s, c; sincos(a.im, &s, &c); return {e * c, e * s}; }

// erf(x + i y) = erf(x) + (2 i / sqrt(pi)) exp(-x^2) int_0^y exp(t^2 - 2 i x t) dt, 24-point Gauss-Legendre.
// Valid for the small |y| = pi / (4 nu_eps) this path produces (host rejects |y| > 0.25).
__constant__ double kGL24x[12] = {0.06405689286260563, 0.19111886747361631, 0.3150426796961634, 0.43379350762604513, 0.54542147138883956, 0.64809365193697555, 0.74012419157855436, 0.82000198597390295, 0.88641552700440107, 0.9382745520027328, 0.97472855597130947, 0.99518721999702131};
__constant__ double kGL24w[12] = {0.12793819534675202, 0.12583745634682825, 0.12167047292780329, 0.11550566805372552, 0.10744427011596556, 0.097618652104113926, 0.086190161531953205, 0.073346481411080161, 0.05929858491543636, 0.044277438817419412, 0.028531388628933559, 0.01234122979998869};

__device__ cplx cerf_small_imag(double x, double y) {
    const double base = erf(x);
    const double ex2 = exp(-x * x);
    if (ex2 == 0.0 || y == 0.0) return {base, 0.0};
    // I = int_0^y exp(t^2) (cos(2xt) - i sin(2xt)) dt
    const double h = 0.5 * y;
    double ir = 0.0, ii = 0.0;
#pragma unroll 1
    for (int q = 0; q < 12; ++q) {
#pragma unroll 1
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            const double t = h + sgn * h * kGL24x[q];
            const double e = exp(t * t) * kGL24w[q];
            double s, c;
            sincos(2.0 * x * t, &s, &c);
            ir += e * c;
            ii -= e * s;
        }
    }
    ir *= h; ii *= h;
    const double k = 1.1283791670955125738961589031215451716881 * ex2;  // 2/sqrt(pi)
    // erf(z) = base + i*k*(ir + i ii) = base - k*ii + i*k*ir
    return {base - k * ii, k * ir};
}

__global__ void dop_z_kernel(const double* __restrict__ freq, const double* __restrict__ nu, int n_grids, int nf,
                             int n_nu, double nu_eps, double* __restrict__ zm) {
    const long long total = (long long)n_grids * nf * n_nu;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(e % n_nu);
        const long long gf = e / n_nu;
        const double om = 2.0 * 3.141592653589793 * freq[gf];
        const double num = nu[m];
        const cplx lj = {log(om), 1.5707963267948966};  // ln(j omega)
        // prefactor: 0.5 sqrt(pi) (jw)^nu_m / eps * (jw)^(ln(jw) / (4 eps^2))
        const double q4 = 4.0 * nu_eps * nu_eps;
        const cplx ex1 = {num * lj.re, num * lj.im};
        const cplx l2 = cmul(lj, {lj.re / q4, lj.im / q4});
        cplx pre = cmul(cexp_(ex1), cexp_(l2));
        const double s = 0.5 * sqrt(3.141592653589793) / nu_eps;
        pre.re *= s; pre.im *= s;
        const double sg = (num > 0.0) ? 1.0 : ((num < 0.0) ? -1.0 : 0.0);
        const double la = fmin(0.0, sg), lb = fmax(0.0, sg);
        const double sh_re = lj.re / (2.0 * nu_eps), sh_im = lj.im / (2.0 * nu_eps);
        const cplx eb = cerf_small_imag(nu_eps * (lb - num) - sh_re, -sh_im);
        const cplx ea = cerf_small_imag(nu_eps * (la - num) - sh_re, -sh_im);
        const cplx fb = cmul(pre, eb), fa = cmul(pre, ea);
        zm[2 * e] = fb.re - fa.re;
        zm[2 * e + 1] = fb.im - fa.im;
    }
}

// ---- DOP voltage-response columns (phasance.py:8-9,40-57,83-99,121-144): real arithmetic ---------------
// column m at time t after a step: F(b) - F(a), F(nu) = 1/2 sqrt(pi) t^-nu_m / Gamma(1 - nu_m) / eps
//                                   * t^(ln t / 4 eps^2) * erf(eps (nu - nu_m) + ln t / 2 eps); summed over steps.
__global__ void dop_v_kernel(const double* __restrict__ times, const double* __restrict__ nu,
                             const double* __restrict__ step_times, const double* __restrict__ step_sizes, int n_grids,
                             int nt, int n_nu, int n_steps, double nu_eps, double* __restrict__ rm) {
    const long long total = (long long)n_grids * nt * n_nu;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(e % n_nu);
        const long long gt = e / n_nu;
        const int g = (int)(gt / nt);
        const double t = times[gt];
        const double num = nu[m];
        const double sg = (num > 0.0) ? 1.0 : ((num < 0.0) ? -1.0 : 0.0);
        const double la = fmin(0.0, sg), lb = fmax(0.0, sg);
        const double inv_gamma = 1.0 / tgamma(1.0 - num);
        double acc = 0.0;
        for (int k = 0; k < n_steps; ++k) {
            const double st = step_times[(size_t)g * n_steps + k];
            if (t > st) {
                const double lt = log(t - st);
                double pre = 0.5 * sqrt(3.141592653589793) * (exp(-num * lt) * inv_gamma) / nu_eps;
                pre *= exp(lt * (lt / (4.0 * nu_eps * nu_eps)));
                const double sh = lt / (2.0 * nu_eps);
                const double f = pre * erf(nu_eps * (lb - num) + sh) - pre * erf(nu_eps * (la - num) + sh);
                acc += step_sizes[(size_t)g * n_steps + k] * f;
            }
        }
        rm[e] = acc;
    }
}

}  // namespace hdrt

using namespace hdrt;

static int grid_for(long long work_items, int per_block) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > 148LL * 64) g = 148LL * 64;
    return (int)g;
}

extern "C" int hdrt_build_lookup(double eps, int grid_points, int quad_points, double* re_x, double* re_v,
                                 double* im_x, double* im_v, double* td_x, double* td_v, void* stream) {
    if (grid_points < 2 || quad_points < 2 || !re_x || !re_v || !im_x || !im_v || !td_x || !td_v) {
        set_error("hdrt_build_lookup: invalid argument");
        return HDRT_ERR_ARG;
    }
    const long long warps = 3LL * grid_points;
    const int blocks = (int)((warps * 32 + kMThreads - 1) / kMThreads);
    lookup_kernel<<<blocks, kMThreads, 0, (cudaStream_t)stream>>>(eps, grid_points, quad_points, re_x, re_v, im_x,
                                                                  im_v, td_x, td_v);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}

extern "C" int hdrt_build_impedance(int mode, const double* freq, const double* tau, int n_grids, int nf, int nb,
                                    double eps, const double* re_x, const double* re_v, const double* im_x,
                                    const double* im_v, int grid_points, int quad_points, double* a_re, double* a_im,
                                    void* stream) {
    if (n_grids == 0) return HDRT_OK;      // an empty batch is a no-op (its buffers may be null)
    if (!freq || !tau || !a_re || !a_im || n_grids < 0 || nf <= 0 || nb <= 0) {
        set_error("hdrt_build_impedance: invalid argument");
        return HDRT_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == HDRT_MODE_INTERP) {
        if (!re_x || !re_v || !im_x || !im_v || grid_points < 2) { set_error("interp mode needs lookup tables"); return HDRT_ERR_ARG; }
        int dev = 0, sms = 148;
        HDRT_CUDA_CHECK(cudaGetDevice(&dev));
        HDRT_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int ctas = 2 * sms;
        // work items: whole grids when there are enough of them, otherwise row tiles (>= 4 rows) so that every
        // CTA of the persistent grid has work
        int rows_per_item = nf;
        while (rows_per_item > 4 && (long long)n_grids * ((nf + rows_per_item - 1) / rows_per_item) < ctas) rows_per_item = (rows_per_item + 1) / 2;
        const int tiles = (nf + rows_per_item - 1) / rows_per_item;
        const size_t smem = sizeof(double) * (6 * (size_t)grid_points + 4 + 2 * (size_t)(nb + rows_per_item));
        if (smem <= 110 * 1024) {
            HDRT_CUDA_CHECK(cudaFuncSetAttribute(impedance_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const long long items = (long long)n_grids * tiles;
            const int grid = (int)(items < ctas ? items : ctas);
            impedance_interp_kernel<<<grid, kIThreads, smem, st>>>(freq, tau, n_grids, nf, nb, re_x, re_v, im_x, im_v,
                                                                   grid_points, a_re, a_im, rows_per_item, tiles);
        } else {
            int rows_per_cta = nf;
            while (rows_per_cta > 8 && (long long)n_grids * ((nf + rows_per_cta - 1) / rows_per_cta) < 148 * 8) rows_per_cta = (rows_per_cta + 1) / 2;
            dim3 grid(n_grids, (nf + rows_per_cta - 1) / rows_per_cta);
            const size_t smem_g = sizeof(double) * (nb + rows_per_cta);
            if (smem_g > 48 * 1024) { set_error("grid too large for the staging buffer"); return HDRT_ERR_UNSUPPORTED; }
            impedance_interp_global_kernel<<<grid, kMThreads, smem_g, st>>>(freq, tau, nf, nb, re_x, re_v, im_x, im_v,
                                                                            grid_points, a_re, a_im, rows_per_cta);
        }
    } else if (mode == HDRT_MODE_TRAPZ) {
        if (quad_points < 2) { set_error("quad_points < 2"); return HDRT_ERR_ARG; }
        const long long total = (long long)n_grids * nf * nb;
        impedance_trapz_kernel<<<grid_for(total * 32, kMThreads), kMThreads, 0, st>>>(freq, tau, n_grids, nf, nb, eps,
                                                                                     quad_points, a_re, a_im);
    } else {
        set_error("unknown mode %d", mode);
        return HDRT_ERR_UNSUPPORTED;
    }
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}

extern "C" int hdrt_build_response(int mode, const double* times, const double* tau, const double* step_times,
                                   const double* step_sizes, int n_grids, int nt, int nb, int n_steps, double eps,
                                   const double* td_x, const double* td_v, int grid_points, int quad_points,
                                   double* rm, void* stream) {
    if (n_grids == 0) return HDRT_OK;
    if (!times || !tau || !step_times || !step_sizes || !rm || n_grids < 0 || nt <= 0 || nb <= 0 || n_steps <= 0) {
        set_error("hdrt_build_response: invalid argument");
        return HDRT_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == HDRT_MODE_INTERP) {
        if (!td_x || !td_v || grid_points < 2) { set_error("interp mode needs the response lookup"); return HDRT_ERR_ARG; }
        const int* run_flag = nullptr;
        {   // production path: persistent CTAs with the lookup table in shared memory (falls through if it declines)
            int dev = 0, sms = 148;
            HDRT_CUDA_CHECK(cudaGetDevice(&dev));
            HDRT_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            const int ctas = 2 * sms;
            int rows_per_item = 128;
            while (rows_per_item > 8 && (long long)n_grids * ((nt + rows_per_item - 1) / rows_per_item) < ctas) rows_per_item /= 2;
            const int tiles = (nt + rows_per_item - 1) / rows_per_item;
            const size_t fsmem = sizeof(double) * (2 * ((size_t)grid_points + 1) + nb + (size_t)rows_per_item * n_steps + n_steps);
            static int* d_flag = nullptr;           // one-word device flag: the fast kernel sets it when it declines
            if (fsmem <= 110 * 1024 && nb <= 4096) {
                if (!d_flag) HDRT_CUDA_CHECK(cudaMalloc(&d_flag, sizeof(int)));
                HDRT_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
                HDRT_CUDA_CHECK(cudaFuncSetAttribute(response_interp_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                const long long items = (long long)n_grids * tiles;
                response_interp_fast_kernel<<<(int)(items < ctas ? items : ctas), kIThreads, fsmem, st>>>(
                    times, tau, step_times, step_sizes, n_grids, nt, nb, n_steps, td_x, td_v, grid_points, rm, rows_per_item, tiles, d_flag);
                HDRT_CUDA_CHECK(cudaGetLastError());
                run_flag = d_flag;          // the generic kernel below runs only if the fast one declined (no host round trip)
            }
        }
        int rows_per_cta = 64;
        while (rows_per_cta > 8 && (long long)n_grids * ((nt + rows_per_cta - 1) / rows_per_cta) < 148 * 8) rows_per_cta /= 2;
        dim3 grid(n_grids, (nt + rows_per_cta - 1) / rows_per_cta);
        const size_t smem = sizeof(double) * (nb + rows_per_cta + 2 * n_steps);
        if (smem > 48 * 1024) { set_error("too many steps / basis points for the staging buffer"); return HDRT_ERR_UNSUPPORTED; }
        response_interp_kernel<<<grid, kMThreads, smem, st>>>(times, tau, step_times, step_sizes, nt, nb, n_steps, td_x,
                                                              td_v, grid_points, rm, rows_per_cta, run_flag);
    } else if (mode == HDRT_MODE_TRAPZ) {
        const long long total = (long long)n_grids * nt * nb;
        response_trapz_kernel<<<grid_for(total * 32, kMThreads), kMThreads, 0, st>>>(times, tau, step_times, step_sizes,
                                                                                   n_grids, nt, nb, n_steps, eps,
                                                                                   quad_points, rm);
    } else {
        set_error("unknown mode %d", mode);
        return HDRT_ERR_UNSUPPORTED;
    }
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}

extern "C" int hdrt_build_penalty(const double* grid, int n_grids, int nb, double eps, int toeplitz, double* m,
                                  void* stream) {
    if (n_grids == 0) return HDRT_OK;
    if (!grid || !m || n_grids < 0 || nb <= 0) { set_error("hdrt_build_penalty: invalid argument"); return HDRT_ERR_ARG; }
    const unsigned per_blocks = ((unsigned)nb * (unsigned)nb + kMThreads - 1) / kMThreads;
    dim3 pgrid(per_blocks < 64 ? per_blocks : 64, n_grids < 32768 ? n_grids : 32768);
    penalty_kernel<<<pgrid, kMThreads, 0, (cudaStream_t)stream>>>(grid, n_grids, nb, eps, toeplitz, m);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}

extern "C" int hdrt_build_eis_vmm(const double* freq, int n_grids, int nf, double vmm_eps, double reim_cor, int uniform,
                                  double* vmm, void* stream) {
    if (n_grids == 0) return HDRT_OK;
    if (!freq || !vmm || n_grids < 0 || nf <= 0) { set_error("hdrt_build_eis_vmm: invalid argument"); return HDRT_ERR_ARG; }
    const long long rows = (long long)n_grids * 2 * nf;
    eis_vmm_kernel<<<grid_for(rows * 32, kMThreads), kMThreads, 0, (cudaStream_t)stream>>>(freq, n_grids, nf, vmm_eps,
                                                                                         reim_cor, uniform, vmm);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}

extern "C" int hdrt_build_chrono_vmm(const double* times, const double* step_times, int n_grids, int nt, int n_steps,
                                     double vmm_eps, int uniform, double* vmm, void* stream) {
    if (!times || !step_times || !vmm || n_grids < 0 || nt < 2 || n_steps <= 0) {
        set_error("hdrt_build_chrono_vmm: invalid argument");
        return HDRT_ERR_ARG;
    }
    if (n_grids == 0) return HDRT_OK;
    size_t smem = sizeof(double) * ((size_t)nt + n_steps) + sizeof(int) * (size_t)nt;
    if (smem > 200 * 1024) { set_error("hdrt_build_chrono_vmm: %d samples do not fit the staging buffer", nt); return HDRT_ERR_UNSUPPORTED; }
    const size_t staged_smem = sizeof(double) * ((size_t)nt + n_steps) + sizeof(int) * ((size_t)nt + (nt & 1)) +
                               sizeof(double) * (size_t)nt * (kMThreads / 32);
    // measured: staging the row costs the occupancy more than the second exponential costs the pipes (1.21 vs 1.50 TB/s
    // at nt = 2000); it pays only while several CTAs still fit on an SM
    const int staged = (!uniform && staged_smem <= 48 * 1024) ? 1 : 0;
    if (staged) smem = staged_smem;
    HDRT_CUDA_CHECK(cudaFuncSetAttribute(chrono_vmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_grids, (nt + kVRows - 1) / kVRows);
    chrono_vmm_kernel<<<grid, kMThreads, smem, (cudaStream_t)stream>>>(times, step_times, nt, n_steps, vmm_eps, uniform, vmm, staged);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}

extern "C" int hdrt_build_dop_v(const double* times, const double* nu, const double* step_times, const double* step_sizes,
                                int n_grids, int nt, int n_nu, int n_steps, double nu_eps, double* rm, void* stream) {
    if (!times || !nu || !step_times || !step_sizes || !rm || n_grids < 0 || nt <= 0 || n_nu <= 0 || n_steps <= 0 ||
        !(nu_eps > 0.0)) {
        set_error("hdrt_build_dop_v: invalid argument");
        return HDRT_ERR_ARG;
    }
    if (n_grids == 0) return HDRT_OK;
    const long long total = (long long)n_grids * nt * n_nu;
    dop_v_kernel<<<grid_for(total, kMThreads), kMThreads, 0, (cudaStream_t)stream>>>(times, nu, step_times, step_sizes,
                                                                                     n_grids, nt, n_nu, n_steps, nu_eps, rm);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}

extern "C" int hdrt_build_dop_z(const double* freq, const double* nu, int n_grids, int nf, int n_nu, double nu_eps,
                                double* zm, void* stream) {
    if (!freq || !nu || !zm || n_grids < 0 || nf <= 0 || n_nu <= 0 || !(nu_eps > 0.0)) {
        set_error("hdrt_build_dop_z: invalid argument");
        return HDRT_ERR_ARG;
    }
    if (3.141592653589793 / (4.0 * nu_eps) > 0.25) {
        set_error("nu_epsilon %.3g too small for the small-imaginary-part complex erf (need >= pi)", nu_eps);
        return HDRT_ERR_UNSUPPORTED;
    }
    if (n_grids == 0) return HDRT_OK;
    const long long total = (long long)n_grids * nf * n_nu;
    dop_z_kernel<<<grid_for(total, kMThreads), kMThreads, 0, (cudaStream_t)stream>>>(freq, nu, n_grids, nf, n_nu, nu_eps, zm);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}
