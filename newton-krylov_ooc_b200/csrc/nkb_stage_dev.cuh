// Device helpers shared by the stage kernels (nkb_stage.cu, nkb_stage_tma.cu).
#pragma once

#include "nkb_common.cuh"

namespace nkb {

template <int MPT>
struct Vec;
template <>
struct Vec<1> {
    double v;
    __device__ __forceinline__ static Vec ld(const double *p) { return {__ldg(p)}; }
    __device__ __forceinline__ void st(double *p) const { *p = v; }
    __device__ __forceinline__ static Vec splat(double s) { return {s}; }
};
template <>
struct Vec<2> {
    double2 v;
    __device__ __forceinline__ static Vec ld(const double *p) {
        return {__ldg(reinterpret_cast<const double2 *>(p))};
    }
    __device__ __forceinline__ void st(double *p) const { *reinterpret_cast<double2 *>(p) = v; }
    __device__ __forceinline__ static Vec splat(double s) { return {make_double2(s, s)}; }
};

// L2 eviction-priority policies (createpolicy + .L2::cache_hint): the streamed inputs are marked
// evict_first and the forward-sweep intermediates evict_last so that the latter survive in L2
// until the back substitution reads them.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ Vec<1> ld_hint(const double *p, uint64_t pol, Vec<1> *) {
    Vec<1> r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r.v) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ Vec<2> ld_hint(const double *p, uint64_t pol, Vec<2> *) {
    Vec<2> r;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.v.x), "=d"(r.v.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_hint(double *p, Vec<1> v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v.v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(double *p, Vec<2> v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.v.x), "d"(v.v.y), "l"(pol) : "memory");
}

__device__ __forceinline__ Vec<1> fma_s(double a, Vec<1> x, Vec<1> y) { return {fma(a, x.v, y.v)}; }
__device__ __forceinline__ Vec<2> fma_s(double a, Vec<2> x, Vec<2> y) {
    return {make_double2(fma(a, x.v.x, y.v.x), fma(a, x.v.y, y.v.y))};
}
__device__ __forceinline__ Vec<1> mul_s(double a, Vec<1> x) { return {a * x.v}; }
__device__ __forceinline__ Vec<2> mul_s(double a, Vec<2> x) { return {make_double2(a * x.v.x, a * x.v.y)}; }
__device__ __forceinline__ Vec<1> add_v(Vec<1> x, Vec<1> y) { return {x.v + y.v}; }
__device__ __forceinline__ Vec<2> add_v(Vec<2> x, Vec<2> y) { return {make_double2(x.v.x + y.v.x, x.v.y + y.v.y)}; }
__device__ __forceinline__ Vec<1> sub_v(Vec<1> x, Vec<1> y) { return {x.v - y.v}; }
__device__ __forceinline__ Vec<2> sub_v(Vec<2> x, Vec<2> y) { return {make_double2(x.v.x - y.v.x, x.v.y - y.v.y)}; }

template <typename F>
__device__ __forceinline__ Vec<1> map_v(Vec<1> x, F f) { return {f(x.v)}; }
template <typename F>
__device__ __forceinline__ Vec<2> map_v(Vec<2> x, F f) { return {make_double2(f(x.v.x), f(x.v.y))}; }
template <typename F>
__device__ __forceinline__ Vec<1> map2_v(Vec<1> x, Vec<1> y, F f) { return {f(x.v, y.v)}; }
template <typename F>
__device__ __forceinline__ Vec<2> map2_v(Vec<2> x, Vec<2> y, F f) {
    return {make_double2(f(x.v.x, y.v.x), f(x.v.y, y.v.y))};
}


// explicit sources of the tracer group at one cell; c[tg] are the tracer values
template <int KIND, int TG, int MPT>
__device__ __forceinline__ void explicit_sources(const StageArgs &p, int tr0, double light, double frc,
                                                 const Vec<MPT> (&c)[TG], Vec<MPT> (&s)[TG]) {
    if constexpr (KIND == NKB_MOD_LINEAR) {
#pragma unroll
        for (int g = 0; g < TG; ++g) s[g] = Vec<MPT>::splat(p.src_const[tr0 + g]);
    } else if constexpr (KIND == NKB_MOD_FORCED_FILE) {
        const double thr_r = p.sink_thres_r;
        s[0] = map_v(c[0], [=](double cv) {
            const double q = thr_r * cv;
            return (thr_r > 0.0 && frc < 0.0 && q > 0.0 && q < 1.0) ? frc * q : frc;
        });
    } else if constexpr (KIND == NKB_MOD_PHOSPHORUS) {
        // phosphorus.py:58-103: uptake, remineralisation (sinking is implicit, class 1)
        const double ul = p.umax * light;
        const double hs = p.halfsat, sg = p.sigma, rd = p.rdop, rp = p.rpop;
        const Vec<MPT> u = map_v(c[0], [=](double po4) { return ul * (po4 / (po4 + hs)); });
        const Vec<MPT> d = mul_s(rd, c[1]);
        const Vec<MPT> q = mul_s(rp, c[2]);
        s[0] = sub_v(add_v(d, q), u);
        s[1] = sub_v(mul_s(sg, u), d);
        s[2] = sub_v(mul_s(1.0 - sg, u), q);
    }
}

// Per-thread pointers of one sweep.  All offsets are element counts.
template <int TG, int NIN>
struct ColPtrs {
    const double *uc[NIN][TG];  // centre value of input i, tracer g at the current level
    double *out[TG];            // output / global y slot at the current level
    const double *sub[TG];      // optional subtrahend (final F = x(T) - x(0))
    const double *est4;         // {eL, eC, eR, 0} at (k, j)
    const double *tri4[TG];     // {ib, g, m, 0} of the tracer's class at (k, j)
    const double *src2;         // {frc(t_exp0), frc(t_exp1)} at (k, j)
    const double *light;
    uint64_t pol_first, pol_last;  // L2 policies (0: no hints)
    ptrdiff_t dl, dr;           // offsets of the south / north neighbour column
    size_t stepk;               // one level down, state arrays
    size_t stepk4;              // one level down, packed [nz][ny][4] tables
};

// One chunk of KCH consecutive levels of the forward sweep: all loads of the chunk are issued
// before the first use (memory-level parallelism), then the recurrences run level by level.
template <int KIND, int TG, int NIN, int MPT, int KCH>
__device__ __forceinline__ void forward_chunk(const StageArgs &p, ColPtrs<TG, NIN> &cp, int k0, int ksm, int tr0,
                                              Vec<MPT> *ys, int nthr, int tid, Vec<MPT> (&yprev)[TG],
                                              const double (&aff)[TG]) {
    Vec<MPT> c[KCH][NIN][TG], cl[KCH][NIN][TG], cr[KCH][NIN][TG];
    double eL[KCH], eC[KCH], eR[KCH], frc[KCH][NIN], mk[KCH][TG], lgt[KCH];
    const bool has_e = (cp.est4 != nullptr);
#pragma unroll
    for (int q = 0; q < KCH; ++q) {
#pragma unroll
        for (int i = 0; i < NIN; ++i) {
#pragma unroll
            for (int g = 0; g < TG; ++g) {
                const double *a = cp.uc[i][g] + q * cp.stepk;
                c[q][i][g] = ld_hint(a, cp.pol_first, (Vec<MPT> *)nullptr);
                if (has_e) {
                    cl[q][i][g] = ld_hint(a + cp.dl, cp.pol_first, (Vec<MPT> *)nullptr);
                    cr[q][i][g] = ld_hint(a + cp.dr, cp.pol_first, (Vec<MPT> *)nullptr);
                }
            }
        }
        if constexpr (KIND == NKB_MOD_FORCED_FILE) {
            if constexpr (NIN == 2) {
                const double2 f = __ldg(reinterpret_cast<const double2 *>(cp.src2 + q * (cp.stepk4 >> 1)));
                frc[q][0] = f.x;
                frc[q][1] = f.y;
            } else {
                frc[q][0] = __ldg(cp.src2 + q * (cp.stepk4 >> 1));
            }
        } else {
#pragma unroll
            for (int i = 0; i < NIN; ++i) frc[q][i] = 0.0;
        }
        if constexpr (KIND == NKB_MOD_PHOSPHORUS) lgt[q] = __ldg(cp.light + q * (cp.stepk4 >> 2)); else lgt[q] = 0.0;
        if (has_e) {
            const double2 e01 = __ldg(reinterpret_cast<const double2 *>(cp.est4 + q * cp.stepk4));
            eL[q] = e01.x;
            eC[q] = e01.y;
            eR[q] = __ldg(cp.est4 + q * cp.stepk4 + 2);
        }
#pragma unroll
        for (int g = 0; g < TG; ++g) mk[q][g] = __ldg(cp.tri4[g] + q * cp.stepk4 + 2);
    }
#pragma unroll
    for (int q = 0; q < KCH; ++q) {
        const int k = k0 + q;
        Vec<MPT> rhs[TG];
#pragma unroll
        for (int g = 0; g < TG; ++g) rhs[g] = Vec<MPT>::splat(0.0);
#pragma unroll
        for (int i = 0; i < NIN; ++i) {
            Vec<MPT> e[TG], s[TG];
#pragma unroll
            for (int g = 0; g < TG; ++g) {
                if (has_e) e[g] = fma_s(eL[q], cl[q][i][g], fma_s(eR[q], cr[q][i][g], mul_s(eC[q], c[q][i][g])));
                else e[g] = Vec<MPT>::splat(0.0);
            }
            explicit_sources<KIND, TG, MPT>(p, tr0, lgt[q], frc[q][i], c[q][i], s);
#pragma unroll
            for (int g = 0; g < TG; ++g) {
                rhs[g] = fma_s(p.a[i], c[q][i][g], rhs[g]);
                rhs[g] = fma_s(p.he[i], add_v(e[g], s[g]), rhs[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < TG; ++g) {
            if (k == 0) rhs[g] = add_v(rhs[g], Vec<MPT>::splat(aff[g]));
            yprev[g] = fma_s(-mk[q][g], yprev[g], rhs[g]);
            if (k < ksm) ys[(size_t)(g * ksm + k) * nthr + tid] = yprev[g];
            else st_hint(cp.out[g] + q * cp.stepk, yprev[g], cp.pol_last);
        }
    }
    // advance to the next chunk
#pragma unroll
    for (int g = 0; g < TG; ++g) {
#pragma unroll
        for (int i = 0; i < NIN; ++i) cp.uc[i][g] += KCH * cp.stepk;
        cp.out[g] += KCH * cp.stepk;
        cp.tri4[g] += KCH * cp.stepk4;
    }
    if (has_e) cp.est4 += KCH * cp.stepk4;
    if constexpr (KIND == NKB_MOD_FORCED_FILE) cp.src2 += KCH * (cp.stepk4 >> 1);
    if constexpr (KIND == NKB_MOD_PHOSPHORUS) cp.light += KCH * (cp.stepk4 >> 2);
}

// levels khi, khi-1, ..., khi-KCH+1; pointers in cp address level khi on entry
template <int TG, int NIN, int MPT, int KCH>
__device__ __forceinline__ void backward_chunk(ColPtrs<TG, NIN> &cp, int khi, int ksm, const Vec<MPT> *ys,
                                               int nthr, int tid, Vec<MPT> (&xnext)[TG]) {
    Vec<MPT> y[KCH][TG], sb[KCH][TG];
    double ib[KCH][TG], gk[KCH][TG];
    const bool has_sub = (cp.sub[0] != nullptr);
#pragma unroll
    for (int q = 0; q < KCH; ++q) {
        const int k = khi - q;
#pragma unroll
        for (int g = 0; g < TG; ++g) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(cp.tri4[g] - q * cp.stepk4));
            ib[q][g] = t.x;
            gk[q][g] = t.y;
            if (k < ksm) {
                y[q][g] = ys[(size_t)(g * ksm + k) * nthr + tid];
            } else {
                // written by this very thread in the forward sweep: plain (coherent) load
                const double *yp = cp.out[g] - q * cp.stepk;
                if constexpr (MPT == 2) y[q][g].v = *reinterpret_cast<const double2 *>(yp);
                else y[q][g].v = *yp;
            }
            if (has_sub) sb[q][g] = Vec<MPT>::ld(cp.sub[g] - q * cp.stepk);
        }
    }
#pragma unroll
    for (int q = 0; q < KCH; ++q) {
#pragma unroll
        for (int g = 0; g < TG; ++g) {
            xnext[g] = fma_s(-gk[q][g], xnext[g], mul_s(ib[q][g], y[q][g]));
            Vec<MPT> o = xnext[g];
            if (has_sub) o = sub_v(o, sb[q][g]);
            st_hint(cp.out[g] - q * cp.stepk, o, cp.pol_first);
        }
    }
#pragma unroll
    for (int g = 0; g < TG; ++g) {
        cp.out[g] -= KCH * cp.stepk;
        cp.tri4[g] -= KCH * cp.stepk4;
        if (has_sub) cp.sub[g] -= KCH * cp.stepk;
    }
}

}  // namespace nkb
