// Column-model year kernel (test_problem, ny == 1): the state of one member is nz*T doubles, so
// the whole model year runs inside ONE persistent kernel with the state resident in shared memory.
// HBM traffic is the initial read and the final write; the member-independent K3 tables (LU
// factors per stage) stream through L2 into a double-buffered shared-memory copy per CTA.
//
// Replaces test_problem.ModelState.comp_fcn's solve_ivp loop (test_problem/model_state.py:79-103)
// with vert_mix.py:19-25 (mixing tendency), iage.py:20-29, dye_decay.py:26-47 and
// phosphorus.py:28-120 (sources).  Scheme: ARS(2,2,2) IMEX; with no horizontal transport the
// implicit part is the L-stable SDIRK2, sources of phosphorus are explicit.

#include <cstdlib>

#include "nkb_common.cuh"

namespace nkb {

// explicit sources at one level for all tracers of the member
template <int KIND, int T>
__device__ __forceinline__ void column_sources(const ColumnArgs &p, double light, int k, const double (&c)[T],
                                               double (&s)[T]) {
    if constexpr (KIND == NKB_MOD_PHOSPHORUS_1D) {
        // test_problem/phosphorus.py:28-120; sinking of pop/pop_s is in the implicit operator
        const double day_r = 1.0 / 86400.0;
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

__device__ __forceinline__ void col_cp8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src)
                 : "memory");
}

// G lanes per member (a power of two >= T: 1, 2, or 16 / 32 for the six-tracer module): the explicit sources of a stage are
// evaluated with the lanes spread over the LEVELS (pointwise work: divisions of the uptake term), the
// two substitution sweeps with the lanes spread over the TRACERS (T independent recurrences); the
// passes are separated by warp barriers only.  The member's state, stage-1 solution, sweep
// intermediates and S(u_n) live in shared memory as [level][tracer]; the member-independent LU
// factors of the next steps arrive through a four-deep cp.async ring shared by the CTA, so no
// global load sits on a recurrence.  S(u_n) is evaluated once per step and used by both stages.
template <int KIND, int T, int G>
__global__ void __launch_bounds__(128) column_year_kernel(const ColumnArgs p) {
    extern __shared__ __align__(16) double sm[];
    constexpr int MPB = 128 / G;  // members per CTA
    constexpr int CNB = 4;        // ring of staged factor tables: three steps of prefetch distance
    const int tid = threadIdx.x;
    const int mi = tid / G, g = tid % G;
    const int nz = p.nz, ncls = p.ncls;
    const int b_raw = blockIdx.x * MPB + mi;
    const bool active = b_raw < p.B;
    const int b = active ? b_raw : p.B - 1;  // idle lane groups shadow the last member (uniform control flow)
    // member stride: a half-warp (16 lanes = 16/G members) must hit 16 distinct 8-byte banks
    int mstride = nz * T;
    while ((mstride & 15) != (G & 15)) ++mstride;
    double *un = sm + (size_t)mi * 4 * mstride, *u1 = un + mstride, *yy = u1 + mstride, *sn = yy + mstride;
    // CTA-shared: light[nz], then a ring of CNB buffers of {tri[2][ncls][nz][4], aff[2][ncls]}
    double *light = sm + (size_t)MPB * 4 * mstride;
    const int cbuf = 2 * ncls * nz * 4 + 2 * ncls;
    double *coef = light + nz;
    const size_t ldb = p.ldb;
    for (int i = g; i < T * nz; i += G) {
        const int t = i / nz, k = i % nz;
        un[k * T + t] = p.x0[((size_t)t * nz + k) * ldb + b];
    }
    for (int k = tid; k < nz; k += blockDim.x) light[k] = p.light ? p.light[k] : 0.0;
    auto issue_coef = [&](int n) {
        if (n < p.n_steps) {
            double *dst = coef + (size_t)(n % CNB) * cbuf;
            const double *tri = p.tri + (size_t)(2 * n) * ncls * nz * 4;
            const double *aff = p.aff + (size_t)(2 * n) * ncls;
            for (int i = tid; i < 2 * ncls * nz * 4; i += blockDim.x) col_cp8(dst + i, tri + i);
            for (int i = tid; i < 2 * ncls; i += blockDim.x) col_cp8(dst + 2 * ncls * nz * 4 + i, aff + i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int n = 0; n < CNB - 1; ++n) issue_coef(n);
    __syncwarp();
    const bool hist0 = (b_raw == 0) && p.hist_slot;
    if (hist0 && p.hist_slot[0] >= 0)
        for (int i = g; i < T * nz; i += G)
            p.hist[(size_t)p.hist_slot[0] * T * nz + i] = un[(i % nz) * T + i / nz];
    const double a1 = (1.0 - kGamma) / kGamma, a0 = 1.0 - a1;
    const int cls = p.class_of[g < T ? g : 0];

    // substitution sweeps of tracer g with the factors {ib, g, m, 0} of one stage: yy (rhs) -> dst.
    // Four levels per round: all shared-memory loads first, then the dependent fma chain, then the stores.
    auto sweeps = [&](const double *tri, double aff, double *dst, auto rhs) {
        if (g < T) {
            const double *tr = tri + (size_t)cls * nz * 4;
            double yprev = 0.0;
            int k = 0;
            for (; k + 4 <= nz; k += 4) {
                double r[4], m[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    r[q] = rhs(k + q);
                    m[q] = tr[(k + q) * 4 + 2];
                }
                if (k == 0) r[0] += aff;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    yprev = fma(-m[q], yprev, r[q]);
                    r[q] = yprev;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) yy[(k + q) * T + g] = r[q];
            }
            for (; k < nz; ++k) {
                double r = rhs(k);
                if (k == 0) r += aff;
                yprev = fma(-tr[k * 4 + 2], yprev, r);
                yy[k * T + g] = yprev;
            }
            double xn = 0.0;
            k = nz - 1;
            for (; k >= 3; k -= 4) {
                double r[4];
                double2 ig[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    r[q] = yy[(k - q) * T + g];
                    ig[q] = *reinterpret_cast<const double2 *>(tr + (k - q) * 4);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    xn = fma(-ig[q].y, xn, ig[q].x * r[q]);
                    r[q] = xn;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[(k - q) * T + g] = r[q];
            }
            for (; k >= 0; --k) {
                const double2 ig = *reinterpret_cast<const double2 *>(tr + k * 4);
                xn = fma(-ig.y, xn, ig.x * yy[k * T + g]);
                dst[k * T + g] = xn;
            }
        }
    };

    double h_next = __ldg(p.h);
    for (int n = 0; n < p.n_steps; ++n) {
        asm volatile("cp.async.wait_group %0;" ::"n"(CNB - 2) : "memory");
        __syncthreads();  // factors of step n staged; every warp is done with the buffer of step n - 1
        issue_coef(n + CNB - 1);
        const double *cb = coef + (size_t)(n % CNB) * cbuf;
        const double *tri0 = cb, *tri1 = cb + (size_t)ncls * nz * 4;
        const double *aff = cb + 2 * ncls * nz * 4;
        const double h = h_next;
        if (n + 1 < p.n_steps) h_next = __ldg(p.h + n + 1);  // off the critical path of the next step
        const double hg = kGamma * h;
        const double e0 = h * (kDelta - 1.0 + kGamma), e1 = h * (1.0 - kDelta);
        if constexpr (KIND == NKB_MOD_LINEAR) {
            // constant sources: the right-hand sides are formed inside the sweeps, no separate pass
            const double src = p.src_const[g < T ? g : 0];
            sweeps(tri0, aff[cls], u1, [&](int k) { return fma(hg, src, un[k * T + g]); });
            __syncwarp();
            const double es = (e0 + e1) * src;
            sweeps(tri1, aff[ncls + cls], un, [&](int k) { return fma(a0, un[k * T + g], fma(a1, u1[k * T + g], es)); });
            __syncwarp();
        } else {
            // stage 1 right-hand side, lanes over levels: S(u_n) kept for stage 2
            for (int k = g; k < nz; k += G) {
                double c[T], sv[T];
#pragma unroll
                for (int t = 0; t < T; ++t) c[t] = un[k * T + t];
                column_sources<KIND, T>(p, light[k], k, c, sv);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    sn[k * T + t] = sv[t];
                    yy[k * T + t] = fma(hg, sv[t], c[t]);
                }
            }
            __syncwarp();
            sweeps(tri0, aff[cls], u1, [&](int k) { return yy[k * T + g]; });
            __syncwarp();
            // stage 2 right-hand side
            for (int k = g; k < nz; k += G) {
                double c1[T], s1[T];
#pragma unroll
                for (int t = 0; t < T; ++t) c1[t] = u1[k * T + t];
                column_sources<KIND, T>(p, light[k], k, c1, s1);
#pragma unroll
                for (int t = 0; t < T; ++t)
                    yy[k * T + t] = a0 * un[k * T + t] + e0 * sn[k * T + t] + (a1 * c1[t] + e1 * s1[t]);
            }
            __syncwarp();
            sweeps(tri1, aff[ncls + cls], un, [&](int k) { return yy[k * T + g]; });
            __syncwarp();
        }
        if (hist0 && p.hist_slot[n + 1] >= 0)
            for (int i = g; i < T * nz; i += G)
                p.hist[(size_t)p.hist_slot[n + 1] * T * nz + i] = un[(i % nz) * T + i / nz];
    }
    if (active)
        for (int i = g; i < T * nz; i += G) {
            const size_t off = (size_t)i * ldb + b;
            p.out[off] = un[(i % nz) * T + i / nz] - p.x0[off];
        }
}

int launch_column_year(int kind, const ColumnArgs &a, cudaStream_t st) {
    const int T = a.T;
    // six tracers: a whole warp per member while the batch is small (shortest step), half a warp beyond
    // (better lane utilisation once there are enough members to fill the SMs)
    const int G = T == 1 ? 1 : (T == 2 ? 2 : (a.B <= 1024 ? 32 : 16));
    const int mpb = 128 / G;
    const size_t smem = ((size_t)mpb * 4 * (a.nz * T + 16) + a.nz + 4 * (2 * a.ncls * a.nz * 4 + 2 * a.ncls)) *
                        sizeof(double);
    if (smem > 220 * 1024) {
        set_error("column_year_kernel: column too deep for shared memory");
        return 2;
    }
    const dim3 grid((a.B + mpb - 1) / mpb), block(128);
#define NKB_COL(K, TT, GG)                                                                               \
    {                                                                                                    \
        auto kern = column_year_kernel<K, TT, GG>;                                                       \
        NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));   \
        kern<<<grid, block, smem, st>>>(a);                                                              \
    }
    if (kind == NKB_MOD_PHOSPHORUS_1D && T == 6 && G == 32) NKB_COL(NKB_MOD_PHOSPHORUS_1D, 6, 32)
    else if (kind == NKB_MOD_PHOSPHORUS_1D && T == 6) NKB_COL(NKB_MOD_PHOSPHORUS_1D, 6, 16)
    else if (kind == NKB_MOD_LINEAR && T == 1) NKB_COL(NKB_MOD_LINEAR, 1, 1)
    else if (kind == NKB_MOD_LINEAR && T == 2) NKB_COL(NKB_MOD_LINEAR, 2, 2)
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
