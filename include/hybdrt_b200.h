/*
 * hybdrt_b200 -- C ABI of the B200-native DRT/DOP inversion engine.
 *
 * The reference (jdhuang-csm/hybrid-drt) is pure Python and has no FFI; the boundary this library
 * replaces is the set of Python functions listed beside each entry point (file:line into the
 * reference tree).  A maintainer binds these symbols with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; all matrices are C-contiguous
 *     (row-major) float64, exactly numpy's default layout;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and stream-ordered;
 *   - the caller owns every buffer; the library owns only the per-device handle;
 *   - return value: 0 = launched, <0 = error (hdrt_last_error() gives the text); there is no CPU
 *     fallback anywhere behind this interface.
 */
#ifndef HYBDRT_B200_H
#define HYBDRT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define HDRT_OK 0
#define HDRT_ERR_ARG (-1)         /* invalid argument                                   */
#define HDRT_ERR_UNSUPPORTED (-2) /* shape/mode outside what the kernels cover          */
#define HDRT_ERR_CUDA (-3)        /* CUDA runtime error (no device, launch failure ...) */

#define HDRT_MODE_INTERP 0 /* np.interp into the 2000-point lookup tables (reference default) */
#define HDRT_MODE_TRAPZ 1  /* per-entry 1000-point trapezoid quadrature                       */

/* per-spectrum status bits written by hdrt_qphb_fit_batch */
#define HDRT_ST_CONVERGED 1   /* outer loop met qphb.is_converged                                 */
#define HDRT_ST_MAXITER 2     /* outer loop ran to max_iter (reference: warning only)             */
#define HDRT_ST_QP_MAXITERS 4 /* some QP hit the interior-point iteration cap (cvxopt 'unknown')  */
#define HDRT_ST_KKT_FAIL 8    /* Cholesky breakdown inside a QP (cvxopt: singular KKT matrix)     */
#define HDRT_ST_NAN 16        /* non-finite coefficients                                          */
#define HDRT_ST_COV_FAIL 32   /* final P not positive definite: no covariance (reference: 'Singular P matrix') */

typedef struct hdrt_handle hdrt_handle;

int hdrt_version(void);
const char* hdrt_last_error(void);

/* Per-device workspace (work-queue counter, SM count).  One handle per device; thread-safe per handle. */
int hdrt_create(hdrt_handle** out, int device);
int hdrt_destroy(hdrt_handle* h);
int hdrt_sm_count(const hdrt_handle* h);

/* ---------------------------------------------------------------------------------------------
 * L1: response-matrix builders (reference: hybdrt/matrices/{basis,mat1d,phasance}.py)
 * ------------------------------------------------------------------------------------------- */

/* basis.generate_impedance_lookup (basis.py:648-669) + basis.generate_response_lookup (:672-689).
 * Six arrays of `grid_points` doubles: ln(omega*tau) grids and values for the real and imaginary
 * impedance integrals, ln(dt/tau) grid and values for the step response. */
int hdrt_build_lookup(double eps, int grid_points, int quad_points, double* re_x, double* re_v, double* im_x,
                      double* im_v, double* td_x, double* td_v, void* stream);

/* mat1d.construct_impedance_matrix (mat1d.py:212-374), both parts at once, for `n_grids` independent
 * (frequency, tau) grids.  freq [n_grids][nf], tau [n_grids][nb] -> a_re, a_im [n_grids][nf][nb].
 * Tables are only read in HDRT_MODE_INTERP. */
int hdrt_build_impedance(int mode, const double* freq, const double* tau, int n_grids, int nf, int nb, double eps,
                         const double* re_x, const double* re_v, const double* im_x, const double* im_v,
                         int grid_points, int quad_points, double* a_re, double* a_im, void* stream);

/* mat1d.construct_response_matrix (mat1d.py:16-122), galvanostatic ideal steps, summed over steps.
 * times [n_grids][nt], tau [n_grids][nb], step_times/step_sizes [n_grids][n_steps] -> rm [n_grids][nt][nb]. */
int hdrt_build_response(int mode, const double* times, const double* tau, const double* step_times,
                        const double* step_sizes, int n_grids, int nt, int nb, int n_steps, double eps,
                        const double* td_x, const double* td_v, int grid_points, int quad_points, double* rm,
                        void* stream);

/* mat1d.construct_integrated_derivative_matrix orders 0,1,2 (mat1d.py:125-209, basis.py:382-395).
 * grid [n_grids][nb] (ln tau, or nu for the DOP block) -> m [n_grids][3][nb][nb].
 * toeplitz != 0 reproduces the reference's uniform-grid shortcut (entry (i,j) from grid[|i-j|]-grid[0]). */
int hdrt_build_penalty(const double* grid, int n_grids, int nb, double eps, int toeplitz, double* m, void* stream);

/* mat1d.construct_eis_var_matrix (mat1d.py:493-515).  freq [n_grids][nf] -> vmm [n_grids][2nf][2nf]. */
int hdrt_build_eis_vmm(const double* freq, int n_grids, int nf, double vmm_eps, double reim_cor, int uniform,
                       double* vmm, void* stream);

/* mat1d.construct_chrono_var_matrix (mat1d.py:455-490) with utils.chrono.get_time_transforms (chrono.py:5-39).
 * times [n_grids][nt] (ascending), step_times [n_grids][n_steps] -> vmm [n_grids][nt][nt].  uniform != 0 is
 * error_structure='uniform' (every entry 1/nt; the solver never needs it materialised: pass vmm_chrono = NULL). */
int hdrt_build_chrono_vmm(const double* times, const double* step_times, int n_grids, int nt, int n_steps,
                          double vmm_eps, int uniform, double* vmm, void* stream);

/* phasance.construct_phasor_z_matrix, gaussian basis, normalize=False (phasance.py:19-37,61-80,108-118).
 * freq [n_grids][nf], nu [n_nu] -> interleaved complex128 zm [n_grids][nf][n_nu][2]. */
int hdrt_build_dop_z(const double* freq, const double* nu, int n_grids, int nf, int n_nu, double nu_eps, double* zm,
                     void* stream);

/* phasance.construct_phasor_v_matrix, gaussian basis, galvanostatic ideal steps, normalize=False
 * (phasance.py:8-9,40-57,83-99,121-144).  times [n_grids][nt], nu [n_nu], step_times/step_sizes [n_grids][n_steps]
 * -> rm [n_grids][nt][n_nu] (summed over steps). */
int hdrt_build_dop_v(const double* times, const double* nu, const double* step_times, const double* step_sizes,
                     int n_grids, int nt, int n_nu, int n_steps, double nu_eps, double* rm, void* stream);

/* ---------------------------------------------------------------------------------------------
 * L0: chrono conditioning (reference: preprocessing.downsample_data preprocessing.py:335-468 with
 * filter_chrono_signal :507-574 and filters.nonuniform_gaussian_filter1d filters/_filters.py:261-341)
 * ------------------------------------------------------------------------------------------- */
/* Antialiasing filter evaluated at the kept samples only, for a batch of traces on one time grid:
 *   out[s][j] = sum_{k=-lw[j]..lw[j]} taps[woff[j] + lw[j] + k] * y[s][mirror(idx[j] + k)]
 * mirrored (scipy.ndimage 'reflect') inside the step segment [seg_lo[j], seg_lo[j] + seg_len[j]).
 * y [n_sig][nt]; idx, seg_lo, seg_len, lw [m] int32; woff [m] int64; taps: the blended, normalised Gaussian
 * weights of every kept sample (laid out by the host once per grid); out [n_sig][m].  All pointers are device. */
int hdrt_filter_gather(const double* y, int n_sig, int nt, const int* idx, const int* seg_lo, const int* seg_len,
                       const long long* woff, const int* lw, const double* taps, int m, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * L2: batched QPHB fit (reference: the loop of DRT._qphb_fit_core, drt1d.py:556-1008, which calls
 * qphb.initialize_weights qphb.py:1609, qphb.iterate_qphb :606, qphb.calculate_pq :1154; the QP is
 * cvxopt.solvers.qp reached at qphb.py:519)
 * ------------------------------------------------------------------------------------------- */
typedef struct hdrt_hypers {
    double derivative_weights[3]; /* qphb.get_default_hypers, qphb.py:208-255 */
    double sigma_ds[3];
    double s_alpha[3];
    double s_0[3];
    double rho_alpha[3];
    double rho_0[3];
    double l2_lambda_0;
    double dop_derivative_weights[3];
    double dop_sigma_ds[3];
    double dop_s_alpha[3];
    double dop_s_0[3];
    double dop_rho_alpha[3];
    double dop_rho_0[3];
    double dop_l2_lambda_0;
    double iw_l1_lambda_0; /* drt1d.py:123; added to EVERY entry of q in the init QP (qphb.py:1660) */
    double iw_l2_lambda_0;
    double iw_alpha; /* only read when has_iw_prior != 0 (qphb.solve_init_weight_scale :1471) */
    double iw_beta;
    double xtol;                 /* drt1d.py:135 */
    double weight_factor;        /* drt1d.py:129 */
    double chrono_weight_factor; /* drt1d.py:126, hybrid only */
    double eis_weight_factor;
    int has_iw_prior;
    int max_iter; /* drt1d.py:135 */
    double outlier_p;  /* prior probability of a point being an outlier (qphb.py:232, :1497-1538); read iff has_outlier_p */
    int has_outlier_p; /* 0 = hypers['outlier_p'] is None */
    int solve_rp;      /* drt1d.py:573-607 + qphb.estimate_x_rp :1684-1717: rescale the data (and the DOP columns)
                          from a lightly regularised solution before the weight initialisation                  */
    int update_scale;  /* drt1d.py:914-936: damped rescaling of the data from iteration 2 on                       */
    int normalize_dop; /* DOP column rescale of solve_rp (drt1d.py:586)                                            */
    int init_weights_separately; /* drt1d.py:647-669: initialise the chrono and the EIS weights from separate fits  */
    int hybrid_wf_method;        /* 1 = hybrid_weight_factor_method='weight' (drt1d.py:749-759); 0 = factors as given */
    double rp_scale;   /* hypers['rp_scale'] (qphb.py:213); read iff solve_rp or update_scale                      */
    double basis_area; /* area of one basis function, sqrt(pi) / epsilon (predict_r_p, drt1d.py:3552-3571)         */
} hdrt_hypers;

typedef struct hdrt_qphb_problem {
    int batch;     /* number of spectra                                               */
    int n_rows;    /* N: data rows (chrono samples first, then Re z, then Im z)       */
    int n_cols;    /* n: special parameters first, then the DRT coefficients          */
    int n_special; /* DRT block = columns [n_special, n_cols)                         */
    int n_chrono;  /* leading rows that are chrono samples (0 for fit_eis)            */
    int dop_start; /* DOP block [dop_start, dop_end) inside the specials; -1 if none  */
    int dop_end;
    int vz_index; /* column rewritten every iteration (drt1d.py:972-979); -1 if none */
    int vb_start; /* v_baseline columns excluded from the vz prediction (:510-511)   */
    int vb_end;
    int hybrid; /* apply chrono/eis weight factors (drt1d.py:882-884)               */

    const double* rm;           /* [N][n] design matrix; + b*rm_stride for spectrum b (0 = shared)        */
    long long rm_stride;
    const double* rv;           /* [batch][N] scaled data vectors                                         */
    const double* vmm_eis;      /* [(N-n_chrono)]^2 variance-estimation block (mat1d.py:493); may be NULL */
    long long vmm_eis_stride;   /*   when N == n_chrono                                                   */
    const double* vmm_chrono;   /* [n_chrono]^2, or NULL = 'uniform' error structure (mat1d.py:482-488)   */
    long long vmm_chrono_stride;
    const double* pen;          /* [3][n][n] embedded penalty matrices m0..m2 (drt1d.py:5863-5910)        */
    long long pen_stride;
    const double* h;            /* [n] rhs of -x <= h (qphb.make_h_constraint :521-557), shared           */
    const double* l1;           /* [n] l1 lambda vector (drt1d.py:557-561), shared                        */
    const double* vz_strength;  /* [N] (drt1d.py:514-519); NULL if vz_index < 0                           */
    hdrt_hypers hyp;

    /* outputs, all [batch][...]; optional ones may be NULL */
    double* x;             /* [n]  cvx_result['x'] of the last QP                          */
    double* weights;       /* [N]  weights returned by the last iterate_qphb (unscaled)    */
    double* est_weights;   /* [N]                                                          */
    double* init_weights;  /* [N]                                                          */
    double* x_overfit;     /* [n]                                                          */
    double* s_vectors;     /* [3][n]                                                       */
    double* rho;           /* [3]                                                          */
    double* dop_rho;       /* [3]   (NULL if no DOP block)                                 */
    double* xmx_norms;     /* [3]                                                          */
    double* dop_xmx_norms; /* [3]   (NULL if no DOP block)                                 */
    double* fun;           /* [1]   'primal objective' of the last QP                      */
    double* vz_col;        /* [N]   final vz_offset column (NULL if vz_index < 0)          */
    double* p_matrix;      /* [n][n] optional: qphb.calculate_pq with the final state      */
    double* q_vector;      /* [n]    optional                                              */
    int* n_outer;          /* [1]   outer iterations executed                              */
    int* n_ipm;            /* [1]   interior-point iterations summed over all QPs          */
    int* status;           /* [1]   HDRT_ST_* bits                                         */

    /* optional post-fit diagnostics of the mapping path (drtmd.py:256-279) */
    const double* eval_mat; /* [n_eval][n] rows of the distribution-evaluation matrix (basis.construct_func_eval_matrix,
                               zero in the special columns), shared by the batch; needed iff dist_var != NULL       */
    int n_eval;
    double* dist_var;      /* [n_eval] diag(B P^-1 B^T) in the scaled space (drt1d.py:3063-3151, :4116-4138)         */
    double* resid_ss;      /* [2]   sum of squared residuals of the final x: chrono rows, EIS rows (qphb.py:1347)   */
    double* outlier_t;     /* [N]   1 - outlier probability of the last weight update (qphb.py:1497-1519); required
                                    iff hyp.has_outlier_p                                                            */

    double* scale_factors; /* [3]   solve_rp data factor, product of the update_scale factors, DOP column factor
                                    (the data vector the fit ended on is rv * [0] * [1]); required iff solve_rp or
                                    update_scale                                                                    */

    /* PFRT (DRT._pfrt_fit_core, drt1d.py:2558-2698): n_pfrt > 0 turns the call into a fit at s_0 * f_0 and
     * l2_lambda_0 / f_0 (hyp holds the base values; hyp.max_iter = max_init_iter) followed by one warm-started
     * continuation (DRT._continue_from_init, :1270-1365) per remaining factor, all inside the kernel.            */
    int n_pfrt;                 /* number of factors; 0 = plain fit                                                */
    int pfrt_max_iter;          /* max_iter_per_step (drt1d.py:2558)                                               */
    int pfrt_min_iter;          /* min_iter of a continuation (drt1d.py:1275): convergence ends it only from here  */
    const double* pfrt_factors; /* [n_pfrt] device                                                                */
    double* pfrt_x;             /* [batch][n_pfrt][n]     step_x: final x of every step                            */
    double* pfrt_llh;           /* [batch][n_pfrt][2]     weighted rss and sum(log w) under weights re-estimated
                                                          from x alone: the data terms of step_llh (qphb.py:1359)  */
    double* pfrt_p;             /* [batch][n_pfrt][n][n]  step_p_mat (optional)                                    */
    int* pfrt_iters;            /* [batch][n_pfrt]        outer iterations per step (optional)                     */
    const double* hybrid_wf_in;  /* [batch][2] chrono / EIS weight factors per spectrum (hybrid_weight_factor_method='rp',
                                    computed by the caller, drt1d.py:761-791); NULL = the scalars of hyp             */
    double* hybrid_wf_out;       /* [batch][2] the factors the fit used (optional)                                  */
    double* x_overfit_eis;       /* [batch][n] x of the EIS-only initial fit when hyp.init_weights_separately (then
                                    x_overfit holds the chrono-only one)                                            */
    const double* weight_factor_vec; /* [N] per-row weight factor, shared by the batch (kk_fit down-weights outliers this
                                        way, drt1d.py:1394-1405); NULL = the scalar hyp.weight_factor                  */
    double* vz_scratch;         /* [batch][N] work buffer: the vz_offset column a continuation step starts from
                                                          (drt1d.py:1296-1302); required iff n_pfrt > 1 and
                                                          vz_index >= 0                                            */
    const double* pen_toeplitz; /* optional hint, [3][n_cols - n_special], shared by the batch: when the DRT block of every
                                   m_k is an exactly symmetric Toeplitz matrix (a uniform ln(tau) grid of Gaussians builds
                                   them that way, mat1d.py:125-209, basis.py:382-395) these are its first rows, and
                                   pen_band is the distance |i - j| beyond which every entry is below 1e-45 of the
                                   largest one.  The kernel may then take the entries of the block from these rows
                                   instead of from pen (which must still be complete).  NULL = no hint.            */
    int pen_band;
} hdrt_qphb_problem;

/* Fills `hyp` with the reference defaults (qphb.py:208-255 eff_hp=True, drt1d.py:102-137). */
void hdrt_default_hypers(hdrt_hypers* hyp);

/* Bytes of dynamic shared memory one spectrum's CTA needs; <0 if the shape does not fit on sm_100a. */
long long hdrt_qphb_smem_bytes(int n_rows, int n_cols);

/* One persistent CTA per spectrum in flight; spectra are pulled from a device-side work counter. */
int hdrt_qphb_fit_batch(hdrt_handle* h, const hdrt_qphb_problem* prob, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cross-observation resolve (reference: hybdrt/mapping/resolve.py resolve_observations :176-341, reached from
 * DRTMD.resolve_observations / resolve_group, mapping/drtmd.py:432-560): one QP per window of `nr` consecutive
 * observations, n = nr * nc unknowns,
 *     minimise 1/2 x'Px + q'x,  -x <= h,    P = blockdiag(P_1 .. P_nr) + My (x) diag(param_scale),
 * solved by the same interior-point iteration as cvxopt.solvers.qp (resolve.py:334).  The caller supplies, per
 * observation, the fit's P / q with the data-dependent parameters folded away and resized to the window's common tau
 * range (resolve.get_offset_pq :11-62, resize_pq :65-135).
 * ------------------------------------------------------------------------------------------- */
typedef struct hdrt_resolve_problem {
    int n_windows;
    int nr;                    /* observations per window                                                   */
    int nc;                    /* parameters per observation; nr * nc <= 1024                               */
    const double* p;           /* [n_obs][nc][nc] per-observation P (symmetric, row-major)                  */
    const double* q;           /* [n_obs][nc]                                                               */
    const int* first_obs;      /* [n_windows] first observation of the window (its nr observations are
                                  consecutive in p / q)                                                     */
    const double* my;          /* [n_windows][nr][nr]  lambda_psi * My (resolve.py:223-273)                 */
    const double* param_scale; /* [n_windows][nc]      (resolve.py:236-262)                                 */
    const double* h;           /* [nc] rhs of -x <= h for one observation (resolve.py:318-329)              */
    double* x;                 /* [n_windows][nr][nc]  res['x'] of every window                             */
    int* iters;                /* [n_windows] interior-point iterations (optional)                          */
    int* status;               /* [n_windows] HDRT_ST_QP_MAXITERS / HDRT_ST_KKT_FAIL bits (optional)        */
} hdrt_resolve_problem;

/* Bytes of device scratch hdrt_resolve_qp_batch needs for this shape (the Cholesky factors of the windows in flight). */
long long hdrt_resolve_work_bytes(const hdrt_handle* h, int n_windows, int nr, int nc);

/* One CTA per window; `work` is a device buffer of hdrt_resolve_work_bytes bytes. */
int hdrt_resolve_qp_batch(hdrt_handle* h, const hdrt_resolve_problem* prob, void* work, void* stream);

/* FP64 FMA probe: runs a register-resident DFMA loop on every SM, on `stream`, and returns achieved TFLOP/s
 * (host-synchronous on that stream; used by bench.py for the FP64 roofline denominator). */
int hdrt_probe_fp64(hdrt_handle* h, double* tflops_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HYBDRT_B200_H */
