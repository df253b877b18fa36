// K3 — member-independent coefficient generation on the device.
//
// For every implicit stage time of the fixed schedule: boundary-layer depth, conservative
// remap of the log-mixing ramp onto the interior depth edges, Peclet limiter, assembly of
// the vertical (tridiagonal) operator per tracer class and its LU (Thomas) factors.
// Replaces, per reference RHS/Jacobian evaluation:
//   py_driver_2d/vert_mix.py:43-101 (mixing_coeff, bldepth), spatial_axis.py:136-187
//   (remap_linear_interpolant), vert_mix.py:140-188 + advection.py:111-179 (vertical part of
//   comp_jacobian), test_problem/vert_mix.py:27-57, and SciPy Radau's sparse LU.
// One thread per (stage, column); writes are coalesced over ypos.

#include "nkb_common.cuh"

namespace nkb {

__device__ __forceinline__ double interp_clamped2(double x, double x0, double x1, double y0, double y1) {
    // np.interp with two points: constant outside, slope*(x-x0)+y0 inside
    if (x <= x0) return y0;
    if (x >= x1) return y1;
    return (y1 - y0) / (x1 - x0) * (x - x0) + y0;
}

// average over [a, b] of the clamped linear ramp (x0,y0)-(x1,y1)  == remap_linear_interpolant
// for a two-point interpolant (spatial_axis.py:136-187): exact integral by trapezoids
__device__ __forceinline__ double ramp_layer_mean(double a, double b, double x0, double x1, double y0,
                                                  double y1) {
    const double fa = interp_clamped2(a, x0, x1, y0, y1);
    const double fb = interp_clamped2(b, x0, x1, y0, y1);
    const bool x0_in = (x0 >= a) && (x0 < b);
    const bool x1_in = (x1 >= a) && (x1 < b);
    if (!x0_in && !x1_in) return 0.5 * (fa + fb);
    double acc;
    if (x0_in && x1_in) {
        acc = (x0 - a) * (0.5 * (fa + y0)) + (x1 - x0) * (0.5 * (y0 + y1)) + (b - x1) * (0.5 * (y1 + fb));
    } else if (x0_in) {
        acc = (x0 - a) * (0.5 * (fa + y0)) + (b - x0) * (0.5 * (y0 + fb));
    } else {
        acc = (x1 - a) * (0.5 * (fa + y1)) + (b - x1) * (0.5 * (y1 + fb));
    }
    return acc / (b - a);
}

__device__ __forceinline__ double interp_pw(double x, const double *xp, const double *fp, int n) {
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    int i = 0;
    while (i < n - 2 && x >= xp[i + 1]) ++i;
    return (fp[i + 1] - fp[i]) / (xp[i + 1] - xp[i]) * (x - xp[i]) + fp[i];
}

// boundary layer depth: vert_mix.py:89-101 (column_model 0), test_problem/vert_mix.py:50-57 (1)
__device__ __forceinline__ double bldepth(const ModelDev &m, double time, int j) {
    if (m.column_model == 0) {
        const double T = 365.0 * 86400.0;
        const double tv[4] = {T * 0.25, T * 0.35, T * 0.65, T * 0.75};
        const double fv[4] = {0.0, 1.0, 1.0, 0.0};
        const double frac = interp_pw(time, tv, fv, 4);
        return 35.0 + (m.bld_max[j] - 35.0) * frac;
    }
    const double year_per_sec = 1.0 / (86400.0 * 365.0);
    const double frac = 0.5 + 0.5 * cos((2.0 * 3.141592653589793) * (year_per_sec * time - 0.25));
    return 50.0 + (150.0 - 50.0) * frac;
}

// mixing coefficient / dz_mid at interior edge ke (1..nz-1)
__device__ __forceinline__ double mixing_coeff_edge(const ModelDev &m, double bld, int ke, int j) {
    const int i = ke - 1;
    if (m.column_model == 0) {
        const double lg = ramp_layer_mean(m.depth_mid[i], m.depth_mid[i + 1], bld - 20.0, bld + 20.0,
                                          2.302585092994046 /* ln 10 */, -7.600902459542082 /* ln 5e-4 */);
        double kap = exp(lg);
        const double pe = 0.5 * m.dz_mid[i] * fabs(m.wvel[ke * m.ny + j]) / kap;
        kap *= (pe > 1.0 ? pe : 1.0);
        return kap * m.dz_mid_r[i];
    }
    const double lg = interp_clamped2(m.depth_edges[ke], bld - 20.0, bld + 20.0, 0.0, -5.0);
    return pow(10.0, lg) * m.dz_mid_r[i];
}

__device__ __forceinline__ double surf_flux(const ModelDev &m, double time) {
    if (m.n_flux_pts <= 0) return 0.0;
    return interp_pw(time, m.flux_t, m.flux_v, m.n_flux_pts);
}

// raw coefficients (sub, diag, sup) of the vertical operator L of class c at level k, column j
// (vert_mix.py:140-188 + vertical part of advection.py:111-179 + module extras)
__device__ __forceinline__ void vert_coeffs(const ModelDev &m, int k, int j, int c, double mc_up, double mc_dn,
                                            double &sub, double &diag, double &sup) {
    const int nz = m.nz, ny = m.ny;
    const double w_up = (k > 0 && m.wvel) ? m.wvel[k * ny + j] : 0.0;
    const double w_dn = (k < nz - 1 && m.wvel) ? m.wvel[(k + 1) * ny + j] : 0.0;
    const double dzr = m.dz_r[k];
    sub = (k > 0) ? dzr * (-0.5 * w_up + mc_up) : 0.0;
    sup = (k < nz - 1) ? dzr * (0.5 * w_dn + mc_dn) : 0.0;
    diag = dzr * (0.5 * w_dn - 0.5 * w_up - mc_dn - mc_up);
    if (k == 0) diag += m.surf_diag[c];
    diag += m.decay[c];
    const double sv = m.sink_vel[c];
    if (sv != 0.0) {
        if (k > 0) sub += sv * dzr;
        if (k < nz - 1) diag -= sv * dzr;
    }
}

// mode 0: raw (sub, diag, sup, aff);  mode 1: factored (m, ib, g, hg*aff) of I - hg*L
// stage times/hg: t_stage[s], hg_stage[s]
__global__ void stage_tables_kernel(ModelDev m, int n_stages, const double *__restrict__ t_stage,
                                    const double *__restrict__ hg_stage, int mode,
                                    double *__restrict__ tri, double *__restrict__ aff) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (j >= m.ny || s >= n_stages) return;
    const int nz = m.nz, ny = m.ny, nc = m.n_classes;
    const double time = t_stage[s];
    const double hg = hg_stage[s];
    const double bld = bldepth(m, time, j);
    const size_t plane = (size_t)nz * ny;

    double mc_up = 0.0;                       // mixing coeff at edge k (above cell k)
    double prev_ib[NKB_MAX_CLASSES], prev_c[NKB_MAX_CLASSES];
    for (int k = 0; k < nz; ++k) {
        const double mc_dn = (k < nz - 1) ? mixing_coeff_edge(m, bld, k + 1, j) : 0.0;
        for (int c = 0; c < nc; ++c) {
            double sub, diag, sup;
            vert_coeffs(m, k, j, c, mc_up, mc_dn, sub, diag, sup);
            double *base = tri + ((((size_t)s * nc + c) * plane) + (size_t)k * ny + j) * 4;
            base[3] = 0.0;
            if (mode == 0) {
                base[0] = sub;
                base[1] = diag;
                base[2] = sup;
            } else {
                const double a = -hg * sub, b = 1.0 - hg * diag, cc = -hg * sup;
                double mk = 0.0, beta = b;
                if (k > 0) {
                    mk = a * prev_ib[c];
                    beta = b - mk * prev_c[c];
                }
                const double ib = 1.0 / beta;
                base[0] = ib;
                base[1] = (k < nz - 1) ? cc * ib : 0.0;
                base[2] = mk;
                prev_ib[c] = ib;
                prev_c[c] = cc;
            }
        }
        mc_up = mc_dn;
    }
    const double flux = surf_flux(m, time) * m.dz_r[0];
    // surface restoring to a record (forced.py:124-130): linear in time, extrapolated at the ends
    double restore_to = 0.0;
    if (m.n_srf > 0) {
        int i = 0;
        while (i < m.n_srf - 2 && time >= m.srf_time[i + 1]) ++i;
        const double w = (time - m.srf_time[i]) / (m.srf_time[i + 1] - m.srf_time[i]);
        const double v0 = m.srf_data[(size_t)i * ny + j], v1 = m.srf_data[(size_t)(i + 1) * ny + j];
        restore_to = v0 + w * (v1 - v0);
    }
    for (int c = 0; c < nc; ++c) {
        const double a = m.surf_aff[c] + flux + m.srf_rate[c] * restore_to;
        aff[((size_t)s * nc + c) * ny + j] = (mode == 0) ? a : hg * a;
    }
}

// forcing record interpolated to time t at one cell (utils.py:533-535: interp1d, linear,
// fill_value="extrapolate")
__device__ __forceinline__ double forcing_at(const ModelDev &m, double t, size_t cell) {
    const size_t plane = (size_t)m.nz * m.ny;
    int i = 0;
    while (i < m.n_frc - 2 && t >= m.frc_time[i + 1]) ++i;
    const double lo = m.frc_data[(size_t)i * plane + cell];
    const double hi = m.frc_data[(size_t)(i + 1) * plane + cell];
    const double slope = (hi - lo) / (m.frc_time[i + 1] - m.frc_time[i]);
    return slope * (t - m.frc_time[i]) + lo;
}

// Tables of the fused step kernel (nkb_step_fused.cu).  Everything member independent that a
// time step needs at (level k, column j) is folded here, once per schedule, into 16-byte pairs
// so that a thread of the step kernel gets two coefficients per shared-memory load:
//   pair plane   contents                                   used by
//   0  {aL, aC}  hg*eL, 1 + hg*eC                            sweep A: rhs1 = aL c_{j-1} + aC c_j + aR c_{j+1} + hg*s
//   1  {aR, m1}  hg*eR, LU multiplier of stage 1
//   2  {fA, 0}   hg*frc(t_n)                       (forcing record, FORCED_FILE)
//   3  {bL, bC}  he1*eL, a1 + he1*eC                         sweep B: rhs2 = P + bL u1_{j-1} + bC u1_j + bR u1_{j+1} + he1*s
//   4  {bR, ib1} he1*eR, 1/beta of stage 1
//   5  {g1, m1}  U factor and L multiplier of stage 1
//   6  {m2, fB}  UL multiplier of stage 2, he1*frc(t_n + gamma h)
//   7  {ib2, g2} UL factors of stage 2                       sweep C
// with hg = gamma h, he1 = h (1 - delta), a1 = (1 - gamma)/gamma (ARS(2,2,2)).  Stage 1 is factored
// top-down (LU, same values as the tri tables), stage 2 bottom-up (UL):
//   elimination  y_k = r_k - m2_k y_{k+1} (k = nz-2..0),  substitution x_k = ib2_k y_k - g2_k x_{k-1}
// so that it can consume the stage-1 solution level by level while that is being back-substituted
// upwards.  Layout ctab[step][class][8][nz][np], np = 2*(ny + 1) doubles: pair of column j at
// [2*(j+1)], one zero pair on the left (the box of a column tile starts at column j0 - 1).
__global__ void step_ctab_kernel(ModelDev m, int n_steps, const double *__restrict__ t_stage,
                                 const double *__restrict__ hg_stage, const double *__restrict__ t_exp,
                                 double *__restrict__ ctab) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (j >= m.ny || s >= n_steps) return;
    const int nz = m.nz, ny = m.ny, nc = m.n_classes;
    const int np = 2 * (ny + 1);
    const size_t pl = (size_t)nz * np;
    const double hg = hg_stage[2 * s];
    const double he1 = hg / kGamma * (1.0 - kDelta);
    const double a1 = (1.0 - kGamma) / kGamma;
    for (int stage = 0; stage < 2; ++stage) {
        const double time = t_stage[2 * s + stage];
        const double bld = bldepth(m, time, j);
        double mc_up = 0.0;
        double prev_ib[NKB_MAX_CLASSES], prev_c[NKB_MAX_CLASSES];
        for (int k = 0; k < nz; ++k) {
            const double mc_dn = (k < nz - 1) ? mixing_coeff_edge(m, bld, k + 1, j) : 0.0;
            const size_t cell = (size_t)k * ny + j;
            double eL = 0.0, eC = 0.0, eR = 0.0;
            if (m.estencil) {
                eL = m.estencil[cell * 4 + 0];
                eC = m.estencil[cell * 4 + 1];
                eR = m.estencil[cell * 4 + 2];
            }
            for (int c = 0; c < nc; ++c) {
                double sub, diag, sup;
                vert_coeffs(m, k, j, c, mc_up, mc_dn, sub, diag, sup);
                const double a = -hg * sub, b = 1.0 - hg * diag, cc = -hg * sup;
                double *base = ctab + ((size_t)s * nc + c) * 8 * pl + (size_t)k * np + 2 * (j + 1);
                if (stage == 0) {
                    double mk = 0.0, beta = b;
                    if (k > 0) {
                        mk = a * prev_ib[c];
                        beta = b - mk * prev_c[c];
                    }
                    const double ib = 1.0 / beta;
                    base[0 * pl + 0] = hg * eL;
                    base[0 * pl + 1] = 1.0 + hg * eC;
                    base[1 * pl + 0] = hg * eR;
                    base[1 * pl + 1] = mk;
                    // explicit source table times the stage weight: forcing record (forced.py:141-151) or
                    // max_uptake_rate * light (phosphorus.py:75-78)
                    base[2 * pl + 0] = (m.kind == NKB_MOD_FORCED_FILE) ? hg * forcing_at(m, t_exp[2 * s], cell)
                                       : (m.kind == NKB_MOD_PHOSPHORUS) ? hg * m.max_uptake_rate * m.light[cell]
                                                                        : 0.0;
                    base[2 * pl + 1] = 0.0;
                    base[3 * pl + 0] = he1 * eL;
                    base[3 * pl + 1] = a1 + he1 * eC;
                    // plane 5 alone serves the stage-1 back substitution (u1_k = ib y_k - g u1_{k+1})
                    base[4 * pl + 0] = he1 * eR;
                    base[4 * pl + 1] = mk;
                    base[5 * pl + 0] = (k < nz - 1) ? cc * ib : 0.0;
                    base[5 * pl + 1] = ib;
                    prev_ib[c] = ib;
                    prev_c[c] = cc;
                } else {  // raw rows first (parked in planes 6 and 7), factored bottom-up below
                    base[6 * pl + 0] = a;
                    base[7 * pl + 0] = b;
                    base[7 * pl + 1] = cc;
                    base[6 * pl + 1] = (m.kind == NKB_MOD_FORCED_FILE) ? he1 * forcing_at(m, t_exp[2 * s + 1], cell)
                                       : (m.kind == NKB_MOD_PHOSPHORUS) ? he1 * m.max_uptake_rate * m.light[cell]
                                                                        : 0.0;
                }
            }
            mc_up = mc_dn;
        }
    }
    for (int c = 0; c < nc; ++c) {
        double ib_next = 0.0, a_next = 0.0;
        for (int k = nz - 1; k >= 0; --k) {
            double *base = ctab + ((size_t)s * nc + c) * 8 * pl + (size_t)k * np + 2 * (j + 1);
            const double a = base[6 * pl + 0], b = base[7 * pl + 0], cc = base[7 * pl + 1];
            double mk = 0.0, beta = b;
            if (k < nz - 1) {
                mk = cc * ib_next;
                beta = b - mk * a_next;
            }
            const double ib = 1.0 / beta;
            base[6 * pl + 0] = mk;
            base[7 * pl + 0] = ib;
            base[7 * pl + 1] = (k > 0) ? a * ib : 0.0;
            ib_next = ib;
            a_next = a;
        }
    }
}

__global__ void mixing_coeff_kernel(ModelDev m, double time, double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m.ny) return;
    const double bld = bldepth(m, time, j);
    for (int ke = 1; ke < m.nz; ++ke) out[(size_t)(ke - 1) * m.ny + j] = mixing_coeff_edge(m, bld, ke, j);
}

// forcing record interpolated to the explicit stage times (utils.py:533-535: interp1d,
// linear, fill_value="extrapolate").  t_eval holds 2 times per step; output [step][cell][2].
__global__ void forcing_tables_kernel(ModelDev m, int n_times, const double *__restrict__ t_eval,
                                      double *__restrict__ src) {
    const size_t plane = (size_t)m.nz * m.ny;
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (cell >= plane || s >= n_times) return;
    const double t = t_eval[s];
    int i = 0;
    while (i < m.n_frc - 2 && t >= m.frc_time[i + 1]) ++i;
    const double lo = m.frc_data[(size_t)i * plane + cell];
    const double hi = m.frc_data[(size_t)(i + 1) * plane + cell];
    const double slope = (hi - lo) / (m.frc_time[i + 1] - m.frc_time[i]);
    src[((size_t)(s >> 1) * plane + cell) * 2 + (s & 1)] = slope * (t - m.frc_time[i]) + lo;
}

int launch_stage_tables(const ModelDev &m, int n_stages, const double *d_t, const double *d_hg, int mode,
                        double *tri, double *aff, cudaStream_t st) {
    dim3 block(64), grid((m.ny + 63) / 64, n_stages);
    stage_tables_kernel<<<grid, block, 0, st>>>(m, n_stages, d_t, d_hg, mode, tri, aff);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int launch_step_ctab(const ModelDev &m, int n_steps, const double *d_t, const double *d_hg, const double *d_texp,
                     double *ctab, cudaStream_t st) {
    dim3 block(64), grid((m.ny + 63) / 64, n_steps);
    step_ctab_kernel<<<grid, block, 0, st>>>(m, n_steps, d_t, d_hg, d_texp, ctab);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int launch_mixing_coeff(const ModelDev &m, double time, double *out, cudaStream_t st) {
    mixing_coeff_kernel<<<(m.ny + 63) / 64, 64, 0, st>>>(m, time, out);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int launch_forcing_tables(const ModelDev &m, int n_times, const double *d_t, double *src, cudaStream_t st) {
    const size_t plane = (size_t)m.nz * m.ny;
    dim3 block(128), grid((unsigned)((plane + 127) / 128), n_times);
    forcing_tables_kernel<<<grid, block, 0, st>>>(m, n_times, d_t, src);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nkb
