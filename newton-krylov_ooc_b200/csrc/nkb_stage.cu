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
__device__ __forceinline__ void explicit_sources(const StageArgs &p, int tr0, size_t cell, double frc,
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
        const double ul = p.umax * __ldg(p.light + cell);
        const double hs = p.halfsat, sg = p.sigma, rd = p.rdop, rp = p.rpop;
        const Vec<MPT> u = map_v(c[0], [=](double po4) { return ul * (po4 / (po4 + hs)); });
        const Vec<MPT> d = mul_s(rd, c[1]);
        const Vec<MPT> q = mul_s(rp, c[2]);
        s[0] = sub_v(add_v(d, q), u);
        s[1] = sub_v(mul_s(sg, u), d);
        s[2] = sub_v(mul_s(1.0 - sg, u), q);
    }
}

// grid: (member blocks, column tiles, tracer groups); block: (BX, J)
template <int KIND, int TG, int NIN, int MPT>
__global__ void __launch_bounds__(512) stage_kernel(const StageArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Vec<MPT> *ys = reinterpret_cast<Vec<MPT> *>(smem_raw);

    const int nz = p.nz, ny = p.ny;
    const int nthr = blockDim.x * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) * MPT;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int tr0 = blockIdx.z * TG;
    if (b >= p.B || j >= ny) return;  // no block-wide barriers below

    const size_t ldb = p.ldb;
    const size_t plane = (size_t)nz * ny;
    const int jm = (j > 0) ? j - 1 : j, jp = (j < ny - 1) ? j + 1 : j;

    Vec<MPT> yprev[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) yprev[g] = Vec<MPT>::splat(0.0);

    // ---- forward elimination, top -> bottom ----
#pragma unroll 2
    for (int k = 0; k < nz; ++k) {
        const size_t cell = (size_t)k * ny + j;
        double eL = 0.0, eC = 0.0, eR = 0.0;
        if (p.est) {
            eL = __ldg(p.est + cell);
            eC = __ldg(p.est + plane + cell);
            eR = __ldg(p.est + 2 * plane + cell);
        }
        Vec<MPT> rhs[TG];
#pragma unroll
        for (int g = 0; g < TG; ++g) rhs[g] = Vec<MPT>::splat(0.0);
#pragma unroll
        for (int i = 0; i < NIN; ++i) {
            Vec<MPT> c[TG], e[TG];
#pragma unroll
            for (int g = 0; g < TG; ++g) {
                const double *base = p.u[i] + ((size_t)(tr0 + g) * plane + (size_t)k * ny) * ldb + b;
                c[g] = Vec<MPT>::ld(base + (size_t)j * ldb);
                if (p.est) {
                    const Vec<MPT> cl = Vec<MPT>::ld(base + (size_t)jm * ldb);
                    const Vec<MPT> cr = Vec<MPT>::ld(base + (size_t)jp * ldb);
                    e[g] = fma_s(eL, cl, fma_s(eR, cr, mul_s(eC, c[g])));
                } else {
                    e[g] = Vec<MPT>::splat(0.0);
                }
            }
            double frc = 0.0;
            if constexpr (KIND == NKB_MOD_FORCED_FILE) frc = __ldg(p.src[i] + cell);
            Vec<MPT> s[TG];
            explicit_sources<KIND, TG, MPT>(p, tr0, cell, frc, c, s);
#pragma unroll
            for (int g = 0; g < TG; ++g) {
                rhs[g] = fma_s(p.a[i], c[g], rhs[g]);
                rhs[g] = fma_s(p.he[i], add_v(e[g], s[g]), rhs[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < TG; ++g) {
            const int cls = p.class_of[tr0 + g];
            const double *tri = p.tri + (size_t)cls * 3 * plane;
            if (k == 0) rhs[g] = add_v(rhs[g], Vec<MPT>::splat(__ldg(p.aff + (size_t)cls * ny + j)));
            const double mk = __ldg(tri + cell);
            yprev[g] = fma_s(-mk, yprev[g], rhs[g]);
            ys[(size_t)(g * nz + k) * nthr + tid] = yprev[g];
        }
    }

    // ---- back substitution, bottom -> top ----
    Vec<MPT> xnext[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) xnext[g] = Vec<MPT>::splat(0.0);
#pragma unroll 2
    for (int k = nz - 1; k >= 0; --k) {
        const size_t cell = (size_t)k * ny + j;
#pragma unroll
        for (int g = 0; g < TG; ++g) {
            const int cls = p.class_of[tr0 + g];
            const double *tri = p.tri + (size_t)cls * 3 * plane;
            const double ib = __ldg(tri + plane + cell);
            const double gk = __ldg(tri + 2 * plane + cell);
            const Vec<MPT> y = ys[(size_t)(g * nz + k) * nthr + tid];
            xnext[g] = fma_s(-gk, xnext[g], mul_s(ib, y));
            const size_t off = ((size_t)(tr0 + g) * plane + cell) * ldb + b;
            Vec<MPT> o = xnext[g];
            if (p.sub) o = sub_v(o, Vec<MPT>::ld(p.sub + off));
            o.st(p.out + off);
        }
    }
}

// full tendency (testing / parity with comp_tend): tend = E(t, c) + L(t) c + aff
// tri here holds RAW (sub, diag, sup); one thread per (member, column), sweeping k.
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
            eL = p.est[cell];
            eC = p.est[plane + cell];
            eR = p.est[2 * plane + cell];
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
        double frc = 0.0;
        if constexpr (KIND == NKB_MOD_FORCED_FILE) frc = p.src[0][cell];
        explicit_sources<KIND, TG, 1>(p, tr0, cell, frc, c, s);
        for (int g = 0; g < TG; ++g) {
            const int cls = p.class_of[tr0 + g];
            const double *tri = p.tri + (size_t)cls * 3 * plane;
            const size_t off = ((size_t)(tr0 + g) * plane + cell) * ldb + b;
            double t = e[g].v + s[g].v + tri[plane + cell] * c[g].v;
            if (k > 0) t += tri[cell] * p.u[0][off - (size_t)ny * ldb];
            if (k < nz - 1) t += tri[2 * plane + cell] * p.u[0][off + (size_t)ny * ldb];
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
    int mpt;
};

// choose (BX, J, MPT) so that the forward-sweep intermediates fit in shared memory
static StageGeom pick_geometry(int nz, int ny, int T, int TG, int B, int ldb) {
    StageGeom g;
    const size_t smem_max = 200 * 1024;
    int mpt = (B >= 2 && (ldb % 2) == 0) ? 2 : 1;
    if (B < 64) mpt = 1;
    const int lanes = (B + mpt - 1) / mpt;   // member lanes needed
    int bx = 1;
    while (bx < lanes && bx < 32) bx <<= 1;
    if (bx > 16 && mpt == 2) bx = 16;         // 16 lanes x 16 B = 256 B contiguous per row
    const size_t per_thread = (size_t)TG * nz * sizeof(double) * mpt;
    int max_thr = (int)(smem_max / per_thread);
    if (max_thr > 512) max_thr = 512;
    int jt = max_thr / bx;
    if (jt < 1) {  // very deep columns: fall back to fewer lanes
        while (bx > 1 && (size_t)bx * per_thread > smem_max) bx >>= 1;
        jt = 1;
    }
    if (jt > ny) jt = ny;
    // balance column tiles
    const int ntile = (ny + jt - 1) / jt;
    jt = (ny + ntile - 1) / ntile;
    g.block = dim3(bx, jt, 1);
    g.grid = dim3((lanes + bx - 1) / bx, ntile, T / TG);
    g.smem = per_thread * bx * jt;
    g.mpt = mpt;
    return g;
}

template <int KIND, int TG, int NIN>
static int launch_stage_t(const StageArgs &a, const StageGeom &g, cudaStream_t st) {
    if (g.mpt == 2) {
        auto kern = stage_kernel<KIND, TG, NIN, 2>;
        static bool attr_set = false;
        if (!attr_set) {
            NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attr_set = true;
        }
        kern<<<g.grid, g.block, g.smem, st>>>(a);
    } else {
        auto kern = stage_kernel<KIND, TG, NIN, 1>;
        static bool attr_set = false;
        if (!attr_set) {
            NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attr_set = true;
        }
        kern<<<g.grid, g.block, g.smem, st>>>(a);
    }
    count_launch();
    return 0;
}

static int tracer_group(int kind, int T) { return kind == NKB_MOD_PHOSPHORUS ? 3 : 1; }

int launch_stage(int kind, int nin, const StageArgs &a, cudaStream_t st) {
    const int TG = tracer_group(kind, a.T);
    const StageGeom g = pick_geometry(a.nz, a.ny, a.T, TG, a.B, a.ldb);
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
