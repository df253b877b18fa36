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
#include "nkb_stage_dev.cuh"

namespace nkb {

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

// test_problem phosphorus (po4, dop, pop and their shadows; test_problem/phosphorus.py:28-120): tendency of the
// column model, one thread per member.  Sources as in column_sources (nkb_column.cu), vertical operator
// from the raw tridiagonal rows {sub, diag, sup, 0} of the tracer's class.
__global__ void tend_p1d_kernel(const StageArgs p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int nz = p.nz;
    const size_t ldb = p.ldb;
    const double day_r = 1.0 / 86400.0, rem = 0.01 * day_r;
    for (int k = 0; k < nz; ++k) {
        double c[6], s[6];
        for (int t = 0; t < 6; ++t) c[t] = p.u[0][((size_t)t * nz + k) * ldb + b];
        const double po4 = c[0];
        const double u = day_r * p.light[k] * (po4 / (po4 + 0.5));
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
            tau = (day_r * p.light[k] * (pd / (pd + 0.5)) - u) / delta;
        }
        const double rest = tau * (c[0] - c[3]);
        s[3] += rest;
        s[4] -= 0.67 * rest;
        s[5] -= 0.33 * rest;
        for (int t = 0; t < 6; ++t) {
            const int cls = p.class_of[t];
            const double *tri = p.tri + ((size_t)cls * nz + k) * 4;
            const size_t off = ((size_t)t * nz + k) * ldb + b;
            double v = s[t] + tri[1] * c[t];
            if (k > 0) v += tri[0] * p.u[0][off - ldb];
            if (k < nz - 1) v += tri[2] * p.u[0][off + ldb];
            if (k == 0) v += p.aff[cls];
            p.out[off] = v;
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
    static unsigned long long attr_mask = 0;
    if (nkb::first_use_on_device(attr_mask)) {
        NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
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
    if (tma_path_usable(a_in)) return launch_stage_tma(kind, nin, a_in, st);
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
        case NKB_MOD_PHOSPHORUS_1D:
            if (a.T != 6 || a.ny != 1) {
                set_error("launch_tend: the column phosphorus module has six tracers");
                return 2;
            }
            tend_p1d_kernel<<<(a.B + 63) / 64, 64, 0, st>>>(a);
            break;
        default: set_error("launch_tend: unsupported module kind"); return 2;
    }
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

// dense [n] vector -> member 0 of a member-fastest batch with row pitch ldb (the other lanes are zeroed)
__global__ void scatter_member_kernel(const double *__restrict__ src, double *__restrict__ dst, size_t n, size_t ldb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * ldb) return;
    dst[i] = (i % ldb == 0) ? src[i / ldb] : 0.0;
}

int launch_scatter_member(const double *src, double *dst, size_t n, size_t ldb, cudaStream_t st) {
    scatter_member_kernel<<<(unsigned)((n * ldb + 255) / 256), 256, 0, st>>>(src, dst, n, ldb);
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
