// K1 + K2 — fused IMEX stage kernel: explicit horizontal advection/mixing stencil + tracer
// sources (K1) assembled straight into the right-hand side of the per-column tridiagonal
// solve of the implicit vertical operator (K2), for a batch of independent members stored
// member-fastest.
//
// Replaces, per reference RHS call and per Radau linear solve:
//   Advection.comp_tend (py_driver_2d/advection.py:51-76), HorizMix.comp_tend
//   (horiz_mix.py:48-67), VertMix.comp_tend (vert_mix.py:24-41), the tracer-module source
//   terms (iage.py:22-41, forced.py:114-154, phosphorus.py:58-95; test_problem/iage.py:20-29,
//   dye_decay.py:26-47) and scipy's splu/solve of the stage systems.
//
// Thread mapping: threadIdx.x -> member (pairs of members when MPT == 2, 16-byte accesses),
// threadIdx.y -> ypos column of the CTA's column tile.  Every thread sweeps its column
// top->bottom (forward elimination, intermediate y kept in shared memory, conflict-free) and
// bottom->top (back substitution, result streamed to HBM).  The LU factors are
// member-independent (K3 tables) so the per-member work is 1 FMA forward and 2 backward per
// cell; horizontal neighbours are the same member in the adjacent column, i.e. the same lane
// of a neighbouring row of the CTA — those loads hit L1.

#include <cstdlib>

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

// grid: (member blocks, column tiles, tracer groups); block: (BX, J).
// Forward-sweep intermediates y: levels [0, ksm) in shared memory (long reuse distance), levels
// [ksm, nz) in the output buffer itself (short reuse distance: stays in L2).
template <int KIND, int TG, int NIN, int MPT, int KC>
__global__ void __launch_bounds__(256) stage_kernel(const StageArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Vec<MPT> *ys = reinterpret_cast<Vec<MPT> *>(smem_raw);

    const int nz = p.nz, ny = p.ny, ksm = p.ksm;
    const int nthr = blockDim.x * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) * MPT;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int tr0 = blockIdx.z * TG;
    if (b >= p.B || j >= ny) return;  // no block-wide barriers below

    const size_t ldb = p.ldb;
    const size_t plane = (size_t)nz * ny;
    ColPtrs<TG, NIN> cp;
    cp.stepk = (size_t)ny * ldb;
    cp.stepk4 = (size_t)ny * 4;
    cp.pol_first = policy_evict_first();
    cp.pol_last = policy_evict_last();
    cp.dl = (j > 0) ? -(ptrdiff_t)ldb : 0;
    cp.dr = (j < ny - 1) ? (ptrdiff_t)ldb : 0;
    double aff[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) {
        const size_t off0 = ((size_t)(tr0 + g) * plane + j) * ldb + b;
#pragma unroll
        for (int i = 0; i < NIN; ++i) cp.uc[i][g] = p.u[i] + off0;
        cp.out[g] = p.out + off0;
        const int cls = p.class_of[tr0 + g];
        cp.tri4[g] = p.tri + ((size_t)cls * plane + j) * 4;
        aff[g] = __ldg(p.aff + (size_t)cls * ny + j);
    }
    cp.est4 = p.est ? p.est + (size_t)j * 4 : nullptr;
    cp.src2 = p.src2 ? p.src2 + (size_t)j * 2 : nullptr;
    cp.light = p.light ? p.light + j : nullptr;

    Vec<MPT> yprev[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) yprev[g] = Vec<MPT>::splat(0.0);

    // ---- forward elimination, top -> bottom ----
    int k = 0;
    for (; k + KC <= nz; k += KC)
        forward_chunk<KIND, TG, NIN, MPT, KC>(p, cp, k, ksm, tr0, ys, nthr, tid, yprev, aff);
    for (; k < nz; ++k) forward_chunk<KIND, TG, NIN, MPT, 1>(p, cp, k, ksm, tr0, ys, nthr, tid, yprev, aff);

    // ---- back substitution, bottom -> top (pointers now address level nz; step back one) ----
#pragma unroll
    for (int g = 0; g < TG; ++g) {
        cp.out[g] -= cp.stepk;
        cp.tri4[g] -= cp.stepk4;
        cp.sub[g] = p.sub ? p.sub + (cp.out[g] - p.out) : nullptr;
    }
    Vec<MPT> xnext[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) xnext[g] = Vec<MPT>::splat(0.0);
    k = nz - 1;
    for (; k - KC + 1 >= 0; k -= KC) backward_chunk<TG, NIN, MPT, KC>(cp, k, ksm, ys, nthr, tid, xnext);
    for (; k >= 0; --k) backward_chunk<TG, NIN, MPT, 1>(cp, k, ksm, ys, nthr, tid, xnext);
}

// full tendency (testing / parity with comp_tend): tend = E(t, c) + L(t) c + aff
// tri here holds RAW {sub, diag, sup, 0}; one thread per (member, column), sweeping k.
template <int KIND, int TG>
__global__ void tend_kernel(const StageArgs p) {
    const int nz = p.nz, ny = p.ny;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int tr0 = blockIdx.z * TG;
    if (b >= p.B || j >= ny) return;
    const size_t ldb = p.ldb;
    const size_t plane = (size_t)nz * ny;
    const int jm = (j > 0) ? j - 1 : j, jp = (j < ny - 1) ? j + 1 : j;
    for (int k = 0; k < nz; ++k) {
        const size_t cell = (size_t)k * ny + j;
        double eL = 0.0, eC = 0.0, eR = 0.0;
        if (p.est) {
            eL = p.est[4 * cell];
            eC = p.est[4 * cell + 1];
            eR = p.est[4 * cell + 2];
        }
        Vec<1> c[TG], e[TG], s[TG];
        for (int g = 0; g < TG; ++g) {
            const double *base = p.u[0] + ((size_t)(tr0 + g) * plane + (size_t)k * ny) * ldb + b;
            c[g] = Vec<1>::ld(base + (size_t)j * ldb);
            e[g] = Vec<1>::splat(0.0);
            if (p.est) {
                const Vec<1> cl = Vec<1>::ld(base + (size_t)jm * ldb);
                const Vec<1> cr = Vec<1>::ld(base + (size_t)jp * ldb);
                e[g] = fma_s(eL, cl, fma_s(eR, cr, mul_s(eC, c[g])));
            }
        }
        double frc = 0.0, lgt = 0.0;
        if constexpr (KIND == NKB_MOD_FORCED_FILE) frc = p.src2[2 * cell];
        if constexpr (KIND == NKB_MOD_PHOSPHORUS) lgt = p.light[cell];
        explicit_sources<KIND, TG, 1>(p, tr0, lgt, frc, c, s);
        for (int g = 0; g < TG; ++g) {
            const int cls = p.class_of[tr0 + g];
            const double *tri = p.tri + ((size_t)cls * plane + cell) * 4;
            const size_t off = ((size_t)(tr0 + g) * plane + cell) * ldb + b;
            double t = e[g].v + s[g].v + tri[1] * c[g].v;
            if (k > 0) t += tri[0] * p.u[0][off - (size_t)ny * ldb];
            if (k < nz - 1) t += tri[2] * p.u[0][off + (size_t)ny * ldb];
            if (k == 0) t += p.aff[(size_t)cls * ny + j];
            p.out[off] = t;
        }
    }
}

// copy member `b` of a member-fastest batch into a dense [n] vector
__global__ void gather_member_kernel(const double *__restrict__ src, double *__restrict__ dst, size_t n,
                                     size_t ldb, int b) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i * ldb + b];
}

struct StageGeom {
    dim3 grid, block;
    size_t smem;
    int mpt, ksm, kc;
};

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// choose (BX, J, MPT) and how many levels keep their forward-sweep intermediates in shared
// memory (the rest goes through the output buffer, i.e. L2).
// Tunables (for experiments): NKB_MPT, NKB_BX, NKB_JT, NKB_KSM (-1: as many as fit), NKB_KC.
static StageGeom pick_geometry(int nz, int ny, int T, int TG, int B, int ldb) {
    StageGeom g;
    const size_t smem_max = (size_t)env_int("NKB_SMEM_KB", 100) * 1024;  // per CTA; 2 CTAs/SM
    int mpt = (B >= 64 && (ldb % 2) == 0) ? 2 : 1;
    mpt = env_int("NKB_MPT", mpt);
    if (mpt == 2 && ((ldb % 2) != 0 || B < 2)) mpt = 1;
    const int lanes = (B + mpt - 1) / mpt;  // member lanes needed
    int bx = 1;
    while (bx < lanes && bx < 32) bx <<= 1;
    if (bx > 16 && mpt == 2) bx = 16;  // 16 lanes x 16 B = 256 B contiguous per row
    bx = env_int("NKB_BX", bx);
    int jt = env_int("NKB_JT", 128 / bx);
    if (jt < 1) jt = 1;
    if (jt > ny) jt = ny;
    const int ntile = (ny + jt - 1) / jt;  // balance column tiles
    jt = (ny + ntile - 1) / ntile;
    g.block = dim3(bx, jt, 1);
    g.grid = dim3((lanes + bx - 1) / bx, ntile, T / TG);
    const size_t per_level = (size_t)TG * sizeof(double) * mpt * bx * jt;
    int ksm = (int)(smem_max / per_level);
    const int ksm_env = env_int("NKB_KSM", 0);  // -1: as many levels as fit in NKB_SMEM_KB
    if (ksm_env >= 0) ksm = ksm_env;
    if (ksm > nz) ksm = nz;
    g.ksm = ksm;
    g.smem = per_level * ksm;
    g.mpt = mpt;
    g.kc = env_int("NKB_KC", (TG >= 3) ? 2 : 4);
    return g;
}

template <int KIND, int TG, int NIN, int MPT, int KC>
static int launch_stage_k(const StageArgs &a, const StageGeom &g, cudaStream_t st) {
    auto kern = stage_kernel<KIND, TG, NIN, MPT, KC>;
    static bool attr_set = false;
    if (!attr_set) {
        NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    kern<<<g.grid, g.block, g.smem, st>>>(a);
    count_launch();
    return 0;
}

template <int KIND, int TG, int NIN>
static int launch_stage_t(const StageArgs &a, const StageGeom &g, cudaStream_t st) {
    if (g.kc >= 4 && TG < 3) {
        if (g.mpt == 2) return launch_stage_k<KIND, TG, NIN, 2, 4>(a, g, st);
        return launch_stage_k<KIND, TG, NIN, 1, 4>(a, g, st);
    }
    if (g.mpt == 2) return launch_stage_k<KIND, TG, NIN, 2, 2>(a, g, st);
    return launch_stage_k<KIND, TG, NIN, 1, 2>(a, g, st);
}

static int tracer_group(int kind, int T) { return kind == NKB_MOD_PHOSPHORUS ? 3 : 1; }

int launch_stage(int kind, int nin, const StageArgs &a_in, cudaStream_t st) {
    const int TG = tracer_group(kind, a_in.T);
    const StageGeom g = pick_geometry(a_in.nz, a_in.ny, a_in.T, TG, a_in.B, a_in.ldb);
    StageArgs a = a_in;
    a.ksm = g.ksm;
    a.hints = env_int("NKB_HINTS", 3);
    int rc = 0;
#define NKB_DISPATCH(K, G)                                          \
    if (nin == 1) rc = launch_stage_t<K, G, 1>(a, g, st);           \
    else rc = launch_stage_t<K, G, 2>(a, g, st);
    switch (kind) {
        case NKB_MOD_LINEAR: NKB_DISPATCH(NKB_MOD_LINEAR, 1); break;
        case NKB_MOD_FORCED_FILE: NKB_DISPATCH(NKB_MOD_FORCED_FILE, 1); break;
        case NKB_MOD_PHOSPHORUS: NKB_DISPATCH(NKB_MOD_PHOSPHORUS, 3); break;
        default: set_error("launch_stage: unsupported module kind"); return 2;
    }
#undef NKB_DISPATCH
    return rc;
}

int launch_tend(int kind, const StageArgs &a, cudaStream_t st) {
    dim3 block(32, 4), grid((a.B + 31) / 32, (a.ny + 3) / 4, 1);
    switch (kind) {
        case NKB_MOD_LINEAR:
            grid.z = a.T;
            tend_kernel<NKB_MOD_LINEAR, 1><<<grid, block, 0, st>>>(a);
            break;
        case NKB_MOD_FORCED_FILE:
            grid.z = a.T;
            tend_kernel<NKB_MOD_FORCED_FILE, 1><<<grid, block, 0, st>>>(a);
            break;
        case NKB_MOD_PHOSPHORUS:
            grid.z = a.T / 3;
            tend_kernel<NKB_MOD_PHOSPHORUS, 3><<<grid, block, 0, st>>>(a);
            break;
        default: set_error("launch_tend: unsupported module kind"); return 2;
    }
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int launch_gather_member(const double *src, double *dst, size_t n, size_t ldb, int b, cudaStream_t st) {
    gather_member_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, n, ldb, b);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nkb
