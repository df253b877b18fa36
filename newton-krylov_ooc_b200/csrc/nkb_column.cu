// Column-model year kernel (test_problem, ny == 1): the state of one member is nz*T doubles, so
// the whole model year runs inside ONE persistent kernel with the state resident in shared memory
// (thread-private slices, conflict-free [k][thread] layout).  HBM traffic is the initial read and
// the final write; the member-independent K3 tables (LU factors per stage) stream through L2/L1.
//
// Replaces test_problem.ModelState.comp_fcn's solve_ivp loop (test_problem/model_state.py:79-103)
// with vert_mix.py:19-25 (mixing tendency), iage.py:20-29, dye_decay.py:26-47 and
// phosphorus.py:28-120 (sources).  Scheme: ARS(2,2,2) IMEX; with no horizontal transport the
// implicit part is the L-stable SDIRK2, sources of phosphorus are explicit.

#include "nkb_common.cuh"

namespace nkb {

// explicit sources at one level for all tracers of the member
template <int KIND, int T>
__device__ __forceinline__ void column_sources(const ColumnArgs &p, int k, const double (&c)[T], double (&s)[T]) {
    if constexpr (KIND == NKB_MOD_PHOSPHORUS_1D) {
        // test_problem/phosphorus.py:28-120; sinking of pop/pop_s is in the implicit operator
        const double day_r = 1.0 / 86400.0;
        const double light = __ldg(p.light + k);
        const double po4 = c[0];
        const double u = day_r * light * (po4 / (po4 + 0.5));
        const double rem = 0.01 * day_r;
#pragma unroll
        for (int o3 = 0; o3 < 6; o3 += 3) {
            s[o3 + 0] = -u + rem * c[o3 + 1] + rem * c[o3 + 2];
            s[o3 + 1] = 0.67 * u - rem * c[o3 + 1];
            s[o3 + 2] = (1.0 - 0.67) * u - rem * c[o3 + 2];
        }
        double tau;
        if (p.restoring_opt == 0) {
            tau = (k == 0) ? day_r : 0.0;
        } else {
            double delta = 1.0e-3 * fabs(po4);
            if (delta < 1.0e-8) delta = 1.0e-8;
            const double pd = po4 + delta;
            tau = (day_r * light * (pd / (pd + 0.5)) - u) / delta;
        }
        const double rest = tau * (c[0] - c[3]);
        s[3] += rest;
        s[4] -= 0.67 * rest;
        s[5] -= 0.33 * rest;
    } else {
#pragma unroll
        for (int t = 0; t < T; ++t) s[t] = p.src_const[t];
    }
}

template <int KIND, int T>
__global__ void __launch_bounds__(128) column_year_kernel(const ColumnArgs p) {
    extern __shared__ double sm[];
    const int nthr = blockDim.x, tid = threadIdx.x;
    const int b = blockIdx.x * nthr + tid;
    const int nz = p.nz;
    const bool active = b < p.B;
    // thread-private slices: un (state at step start), u1 (stage 1 / result), yy (forward sweep)
    double *un = sm, *u1 = sm + (size_t)T * nz * nthr, *yy = sm + (size_t)2 * T * nz * nthr;
    auto idx = [&](int t, int k) { return ((size_t)t * nz + k) * nthr + tid; };
    const size_t ldb = p.ldb;
    if (active)
        for (int t = 0; t < T; ++t)
            for (int k = 0; k < nz; ++k) un[idx(t, k)] = p.x0[((size_t)t * nz + k) * ldb + b];
    const double a1 = (1.0 - kGamma) / kGamma, a0 = 1.0 - a1;
    const size_t tri_stage = (size_t)p.ncls * nz * 4;
    if (active && b == 0 && p.hist_slot && p.hist_slot[0] >= 0)
        for (int t = 0; t < T; ++t)
            for (int k = 0; k < nz; ++k) p.hist[((size_t)p.hist_slot[0] * T + t) * nz + k] = un[idx(t, k)];

    for (int n = 0; n < p.n_steps && active; ++n) {
        const double h = __ldg(p.h + n);
        const double hg = kGamma * h;
#pragma unroll 1
        for (int stage = 0; stage < 2; ++stage) {
            const double *tri = p.tri + (size_t)(2 * n + stage) * tri_stage;
            const double *aff = p.aff + (size_t)(2 * n + stage) * p.ncls;
            const double c0 = stage == 0 ? 1.0 : a0, c1 = stage == 0 ? 0.0 : a1;
            const double e0 = stage == 0 ? hg : h * (kDelta - 1.0 + kGamma), e1 = stage == 0 ? 0.0 : h * (1.0 - kDelta);
            double yprev[T];
#pragma unroll
            for (int t = 0; t < T; ++t) yprev[t] = 0.0;
            for (int k = 0; k < nz; ++k) {
                double cn[T], sn[T], rhs[T];
#pragma unroll
                for (int t = 0; t < T; ++t) cn[t] = un[idx(t, k)];
                column_sources<KIND, T>(p, k, cn, sn);
#pragma unroll
                for (int t = 0; t < T; ++t) rhs[t] = c0 * cn[t] + e0 * sn[t];
                if (stage == 1) {
                    double c1v[T], s1[T];
#pragma unroll
                    for (int t = 0; t < T; ++t) c1v[t] = u1[idx(t, k)];
                    column_sources<KIND, T>(p, k, c1v, s1);
#pragma unroll
                    for (int t = 0; t < T; ++t) rhs[t] += c1 * c1v[t] + e1 * s1[t];
                }
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const int cls = p.class_of[t];
                    if (k == 0) rhs[t] += __ldg(aff + cls);
                    const double mk = __ldg(tri + ((size_t)cls * nz + k) * 4 + 2);
                    yprev[t] = fma(-mk, yprev[t], rhs[t]);
                    yy[idx(t, k)] = yprev[t];
                }
            }
            // stage 0 -> u1; stage 1 -> un (all reads of un/u1 at this level are done before)
            double *dst = stage == 0 ? u1 : un;
            double xnext[T];
#pragma unroll
            for (int t = 0; t < T; ++t) xnext[t] = 0.0;
            for (int k = nz - 1; k >= 0; --k) {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const int cls = p.class_of[t];
                    const double2 ig = __ldg(reinterpret_cast<const double2 *>(tri + ((size_t)cls * nz + k) * 4));
                    xnext[t] = fma(-ig.y, xnext[t], ig.x * yy[idx(t, k)]);
                    dst[idx(t, k)] = xnext[t];
                }
            }
        }
        if (b == 0 && p.hist_slot && p.hist_slot[n + 1] >= 0)
            for (int t = 0; t < T; ++t)
                for (int k = 0; k < nz; ++k) p.hist[((size_t)p.hist_slot[n + 1] * T + t) * nz + k] = un[idx(t, k)];
    }
    if (active)
        for (int t = 0; t < T; ++t)
            for (int k = 0; k < nz; ++k) {
                const size_t off = ((size_t)t * nz + k) * ldb + b;
                p.out[off] = un[idx(t, k)] - p.x0[off];
            }
}

int launch_column_year(int kind, const ColumnArgs &a, cudaStream_t st) {
    const int T = a.T;
    int nthr = 128;
    while ((size_t)3 * T * a.nz * nthr * sizeof(double) > 200 * 1024 && nthr > 32) nthr >>= 1;
    const size_t smem = (size_t)3 * T * a.nz * nthr * sizeof(double);
    if (smem > 220 * 1024) {
        set_error("column_year_kernel: column too deep for shared memory");
        return 2;
    }
    const dim3 grid((a.B + nthr - 1) / nthr), block(nthr);
#define NKB_COL(K, TT)                                                                                   \
    {                                                                                                    \
        auto kern = column_year_kernel<K, TT>;                                                           \
        NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));   \
        kern<<<grid, block, smem, st>>>(a);                                                              \
    }
    if (kind == NKB_MOD_PHOSPHORUS_1D && T == 6) NKB_COL(NKB_MOD_PHOSPHORUS_1D, 6)
    else if (kind == NKB_MOD_LINEAR && T == 1) NKB_COL(NKB_MOD_LINEAR, 1)
    else if (kind == NKB_MOD_LINEAR && T == 2) NKB_COL(NKB_MOD_LINEAR, 2)
    else {
        set_error("column_year_kernel: unsupported module kind / tracer count");
        return 2;
    }
#undef NKB_COL
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nkb
