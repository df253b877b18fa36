/*
 * nkb200.h — C ABI of the B200-native hot path of Newton-Krylov_OOC.
 *
 * The reference (klindsay28/Newton-Krylov_OOC) has no FFI: its plug-in boundary is the
 * Python operator surface ModelStateBase / TracerModuleStateBase.  The
 * entry points below are what a native replacement of the hot functions behind that
 * surface binds (ctypes stub in INTEGRATION.md).  Each one cites the reference code it
 * replaces (file:line under the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; nkb_last_error() gives the
 *     message of the last failure on the calling thread.  No exceptions cross the boundary.
 *   - pointers named d_* are DEVICE pointers (float64 unless stated), h_* are HOST pointers.
 *   - a batch of B independent states ("members") is stored member-fastest:
 *         x[((t*nz + k)*ny + j)*ldb + b],  t tracer, k depth, j ypos, b member, ldb >= B,
 *     ldb a multiple of 2 (16-byte vector access).  B == 1 is the reference's own layout
 *     [tracer, depth, ypos] with ypos fastest (py_driver_2d/tracer_module_state.py:103-108).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Nothing
 *     synchronises the host unless stated.
 *   - all arithmetic is IEEE float64.
 */
#ifndef NKB200_H
#define NKB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NKB_MAX_TRACERS 6
#define NKB_MAX_CLASSES 3

/* tracer-module kinds (py_driver_2d: iage.py, forced.py, phosphorus.py;
 * test_problem: iage.py, dye_decay.py, phosphorus.py) */
enum {
    NKB_MOD_LINEAR = 0,      /* constant explicit source per tracer (iage; forced const/decay/none) */
    NKB_MOD_FORCED_FILE = 1, /* forced with sms from a forcing record (+ optional sink_thres) */
    NKB_MOD_PHOSPHORUS = 2,  /* py_driver_2d phosphorus: po4/dop/pop */
    NKB_MOD_PHOSPHORUS_1D = 3 /* test_problem phosphorus: po4,dop,pop + shadows */
};

/* Parameters of one tracer module on one grid.  Host arrays are copied by nkb_model_create. */
typedef struct nkb_model_desc {
    int32_t nz, ny;        /* ny == 1: test_problem column model */
    int32_t n_tracers;     /* T */
    int32_t kind;          /* NKB_MOD_* */
    int32_t n_classes;     /* number of distinct implicit (vertical) operators */
    int32_t class_of[NKB_MAX_TRACERS];
    int32_t column_model;  /* 0: py_driver_2d vertical mixing (vert_mix.py:43-101);
                              1: test_problem vertical mixing (test_problem/vert_mix.py:27-57) */
    double t0, t1;         /* time_range (py_driver_2d/model_state.py:49) */

    /* axis metrics (spatial_axis.py:35-39) */
    const double *h_depth_edges; /* [nz+1] */
    const double *h_ypos_mid;    /* [ny]   */
    /* time-invariant transport fields */
    const double *h_wvel;        /* [nz+1][ny]   advection.py:45 (NULL: zero) */
    const double *h_estencil;    /* [3][nz][ny]  eL,eC,eR: horizontal advection+mixing in
                                    coefficient form (advection.py:58-65, horiz_mix.py:59-65);
                                    NULL when ny == 1 */
    const double *h_bld_max;     /* [ny] vert_mix.py:93-97 (column_model 0) */

    /* implicit extras per class */
    double surf_diag[NKB_MAX_CLASSES];   /* added to diag at k=0 (surface restoring / piston velocity*dz_r) */
    double surf_aff[NKB_MAX_CLASSES];    /* constant source at k=0 (rate*restore_to) */
    double decay[NKB_MAX_CLASSES];       /* added to diag at every k (e.g. -lambda) */
    double sink_vel[NKB_MAX_CLASSES];    /* upwind sinking velocity (m/s), phosphorus pop */
    /* optional time-dependent surface flux (test_problem dye_decay.py:17-47): piecewise linear */
    int32_t n_flux_pts;
    double flux_t[8], flux_v[8];         /* added as flux_v(t)*dz_r[0] to the k=0 tendency, all classes */

    /* explicit sources */
    double src_const[NKB_MAX_TRACERS];   /* e.g. 1/T_yr for iage (iage.py:39) */
    double sink_thres;                   /* <=0: off (forced.py:141-151) */
    int32_t n_frc;                       /* forcing records (forced sms file), 0: none */
    const double *h_frc_time;            /* [n_frc] */
    const double *h_frc_data;            /* [n_frc][nz][ny], scalef applied, on the model grid */
    const double *h_light;               /* [nz][ny] phosphorus light limitation (phosphorus.py:23-26) */
    double po4_halfsat, max_uptake_rate, sigma, dop_remin_rate, pop_remin_rate; /* phosphorus.py:41-47 */
    int32_t po4_s_restoring_opt;         /* test_problem phosphorus.py:58-70 */

    /* optional surface restoring to a time-dependent record (forced_surf_restore_opt = file,
     * py_driver_2d/forced.py:46-51,124-130): restore_to(t, ypos) linearly interpolated in time (with
     * extrapolation, utils.py:533-535); srf_rate[c] * restore_to is added to the k=0 tendency of class
     * c (the matching -rate goes into surf_diag).  n_srf == 0: off */
    int32_t n_srf;
    const double *h_srf_time;            /* [n_srf] */
    const double *h_srf_data;            /* [n_srf][ny], on the model's ypos axis */
    double srf_rate[NKB_MAX_CLASSES];
} nkb_model_desc;

typedef struct nkb_model nkb_model; /* opaque */

const char *nkb_last_error(void);
int nkb_version(void);
/* number of kernel launches issued by this library since load (bench.py: gpu_launches) */
uint64_t nkb_launch_count(void);

/* ---- model set-up ------------------------------------------------------------------ */
int nkb_model_create(nkb_model **out, const nkb_model_desc *desc);
void nkb_model_destroy(nkb_model *m);

/* Fixed integration schedule: n_steps steps, step n covers [t_start[n], t_start[n]+h[n]].
 * Builds on the device, for every implicit stage of every step, the member-independent
 * tables of the vertical operator (boundary-layer depth, conservative remap of the log
 * mixing ramp, Peclet limiter, tridiagonal assembly, LU factors) — replaces
 * VertMix.mixing_coeff/bldepth (py_driver_2d/vert_mix.py:43-101),
 * SpatialAxis.remap_linear_interpolant (spatial_axis.py:136-187), VertMix.comp_jacobian
 * (vert_mix.py:140-188) and SciPy Radau's LU of (mu/h I - J).  Synchronises. */
int nkb_model_set_schedule(nkb_model *m, int n_steps, const double *h_t_start, const double *h_h);

/* kappa/dz_mid at interior edges at one time, [nz-1][ny] on the device (vert_mix.py:43-87) */
int nkb_model_mixing_coeff(nkb_model *m, double time, double *d_out, void *stream);

/* full tendency dc/dt(time, x) for B members — replaces TracerModuleState.comp_tend
 * (py_driver_2d/tracer_module_state.py:98-108 + iage.py:22-41 / forced.py:114-154 /
 * phosphorus.py:58-95; test_problem/iage.py:20-29, dye_decay.py:26-47, phosphorus.py:28-120) */
int nkb_model_tend(nkb_model *m, double time, const double *d_x, double *d_tend, int B, int ldb,
                   void *stream);

/* F(x) = x(T) - x(0): one model year for B members — replaces ModelState.comp_fcn's
 * solve_ivp loop (py_driver_2d/model_state.py:94-121; test_problem/model_state.py:79-103).
 * d_work: scratch of nkb_model_work_doubles(m, B, ldb) doubles.
 * hist: if n_hist > 0, d_hist[n_hist][T][nz][ny] receives member 0's state after the steps
 * listed in h_hist_steps (step index 0 = initial state, n_steps = final state). */
size_t nkb_model_work_doubles(const nkb_model *m, int B, int ldb);
int nkb_model_eval(nkb_model *m, const double *d_x0, double *d_f, double *d_work, int B, int ldb,
                   int n_hist, const int *h_hist_steps, double *d_hist, void *stream);

/* Non-blocking health check of the persistent step kernel: returns 1 (and clears the flag) when a
 * CTA of an earlier nkb_model_eval gave up waiting for a neighbour tile (its results are then
 * invalid), else 0.  Meaningful after the stream of that evaluation has been synchronised;
 * nkb_model_eval and nkb_model_eval_host check it themselves. */
int nkb_model_poll_error(nkb_model *m);

/* Same through HOST buffers: h_x0/h_f are [B][T][nz][ny] (member-major, the reference's
 * per-state layout); does H2D, pack, eval, unpack, D2H and synchronises.  Host buffers
 * should be pinned for full PCIe rate. */
int nkb_model_eval_host(nkb_model *m, const double *h_x0, double *h_f, int B);

/* ---- preconditioner (K4) ----------------------------------------------------------- */
/* Member-shared banded LU:  factor once (host-assembled band, LAPACK gbtrf-style storage
 * without pivoting is NOT assumed: partial pivoting is done on the device), then batched
 * solves with member-fastest right-hand sides.  Replaces scipy.linalg.solve_banded
 * (test_problem/iage.py:50, dye_decay.py:71) and scipy.sparse.linalg.spsolve
 * (py_driver_2d/iage.py:91, forced.py:239). */
typedef struct nkb_banded nkb_banded;
int nkb_banded_create(nkb_banded **out, int n, int kl, int ku, const double *h_ab /* [kl+ku+1][n] */);
void nkb_banded_destroy(nkb_banded *f);
/* number of independent diagonal blocks found in the matrix (rows that no band entry couples, e.g.
 * the per-column systems of a grid without lateral processes): they are factored and solved in
 * parallel */
int nkb_banded_blocks(const nkb_banded *f);
/* which substitution the factor was prepared for: 3 = panel kernel (one wide block, no row interchanges: 16 rows per
 * barrier pair), 2 = batched Thomas kernel (bandwidth <= 4, no interchanges, used for B >= 16), 1 = row-by-row window
 * kernel (everything else, incl. every matrix that needed interchanges) */
int nkb_banded_path(const nkb_banded *f);
/* d_y, d_x: [n][ldb]; x = A^-1 y (in place allowed); if subtract_rhs, x = A^-1 (scale*y) - y */
int nkb_banded_solve(nkb_banded *f, const double *d_y, double *d_x, int B, int ldb, double scale,
                     int subtract_rhs, void *stream);

/* ---- Krylov vector kernels (K5/K6) --------------------------------------------------- */
/* layout conversion between member-major [B][n] and member-fastest [n][ldb] */
int nkb_pack_members(const double *d_src_major, double *d_dst_fast, int n, int B, int ldb, void *stream);
int nkb_unpack_members(const double *d_src_fast, double *d_dst_major, int n, int B, int ldb, void *stream);

/* region-weighted dot products: out[r][b] = sum_t sum_{cell in r} w[cell]*a*b — replaces
 * TracerModuleStateBase.dot_prod/mean (tracer_module_state_base.py:371-388).  The weights
 * are the CSR region-mean matrix of model_config.py:292-315 (rows = regions): d_indptr
 * [R+1], d_indices [nnz] (cell ids), d_wdata [nnz] (grid_weight / region sum).
 * a/b: [T][ncell][ldb]; d_b == NULL computes the mean of a.  Deterministic two-pass
 * reduction: d_partial is scratch of n_chunks*R*B doubles (n_chunks from nkb_wdot_chunks). */
int nkb_wdot_chunks(int ncell_max);
int nkb_wdot(const int32_t *d_indptr, const int32_t *d_indices, const double *d_wdata, int R, int T,
             int ncell, const double *d_a, const double *d_b, int B, int ldb, double *d_partial,
             int n_chunks, double *d_out /* [R][B] */, void *stream);

/* y = alpha[r(cell)][b]*x + beta[r(cell)][b]*y with per-(region, member) scalars — replaces the
 * operators of tracer_module_state_base.py:255-369 via broadcast_region_vals (:502-515).
 * d_region: int32 [ncell], 1-based, 0 = outside every region (scalars there are `fill`). */
int nkb_axpby(const int32_t *d_region, int R, int T, int ncell, const double *d_alpha,
              const double *d_x, const double *d_beta, double *d_y, double fill_alpha,
              double fill_beta, int B, int ldb, void *stream);

/* modified Gram-Schmidt of w against k basis vectors — replaces ModelStateBase.mod_gram_schmidt
 * (model_state_base.py:365-377: for i < k: h_i = dot(w, v_i); w -= h_i v_i, each step re-reading a basis
 * file).  One cooperative launch with w resident in registers and every basis vector read once
 * (8 N (k + 2) bytes) when w fits on the chip and R*B <= 1024, k <= 64; otherwise one dot and one update
 * per vector.  Either way the k [R][B] scalars are written to d_h and nothing is synchronised with the
 * host.  Region weights as for nkb_wdot plus their dense form: d_region int32 [ncell] (1-based, 0 = no
 * region), d_cellw [ncell] (grid_weight / region sum, 0 outside regions).  h_basis: host array of k device
 * pointers, each [T][ncell][ldb].  d_scratch: nkb_mgs_scratch_doubles(R, B, ncell_max_row) doubles. */
size_t nkb_mgs_scratch_doubles(int R, int B, int ncell_max_row);
int nkb_mgs(const int32_t *d_indptr, const int32_t *d_indices, const double *d_wdata, const int32_t *d_region,
            const double *d_cellw, int R, int T, int ncell, int ncell_max_row, double *d_w,
            const double *const *h_basis, int k, int B, int ldb, double *d_scratch, size_t scratch_doubles,
            double *d_h /* [k][R][B] */, void *stream);

/* out = sum_i coeff[i][r(cell)][b] * basis_i (+ add) in one pass — replaces model_state_base.lin_comb
 * (model_state_base.py:619-624; krylov_solver.py:141-154).  d_coeff [k][R][B]; cells outside every
 * region use `fill` (broadcast_region_vals, tracer_module_state_base.py:502-515); d_add may be NULL or
 * equal to d_out. */
int nkb_lin_comb(const int32_t *d_region, int R, int T, int ncell, const double *d_coeff,
                 const double *const *h_basis, int k, const double *d_add, double *d_out, double fill, int B,
                 int ldb, void *stream);

/* result columns gathered from G ranks ([G][n][W], the output of an all-gather of member-fastest blocks of W
 * members) -> one member-fastest batch [n][ldo] holding the first B of the G*W members: the gather of
 * F / JVP columns to the owner of the Krylov basis (krylov_solver.py:127). */
int nkb_interleave_blocks(const double *d_gathered, double *d_out, size_t n, int G, int W, int ldo, int B,
                          void *stream);

/* finite-difference JVP pieces (model_state_base.py:492-527):
 *   sigma[r][b] = 1e-4*norm (1 where 0);  perturb = x + sigma*v;  jvp = (fp - f0)/sigma */
int nkb_fd_sigma(const double *d_norm, double *d_sigma, int n, void *stream);

/* limiter of the Newton increment (tracer_module_state_base.py:115-151 apply_limiter ->
 * utils.py:544-600 comp_scalef_lob / comp_scalef_upb / min_by_region): per (region, member) the
 * largest scale factor in [0, 1] such that base + scalef*inc stays inside [lob, upb] for every
 * tracer and cell of the region.  d_out [R][B] must be pre-filled by the caller with the value
 * for regions that hold no cell (+inf in the reference); it is lowered by an exact atomic min.
 * d_flag [B] (zeroed by the caller) collects per member the bits 1: base < lob somewhere, 2: base + inc < lob
 * somewhere, 4: base > upb, 8: base + inc > upb.  The reference raises ValueError("base < lob") only when 1 and
 * 2 are both set (a bound that needs enforcing is already violated by base), likewise 4 and 8 for upb. */
int nkb_limiter_scalef(const int32_t *d_region, int R, int T, int ncell, const double *d_base,
                       const double *d_inc, double lob, int has_lob, double upb, int has_upb, int B,
                       int ldb, double *d_out /* [R][B] */, int32_t *d_flag, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NKB200_H */
