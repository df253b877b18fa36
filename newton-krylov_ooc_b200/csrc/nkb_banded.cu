// K4 — member-shared banded LU with partial pivoting + batched member-fastest solves.
//
// The preconditioner matrix is member-independent, so it is factored ONCE on the device
// (LAPACK dgbtf2-style right-looking elimination inside the band) and the factor is then applied
// to every member's right-hand side.  Replaces scipy.linalg.solve_banded((1,1), ...)
// (test_problem/iage.py:50, dye_decay.py:71) and scipy.sparse.linalg.spsolve
// (py_driver_2d/iage.py:91, forced.py:239).
//
// Parallelism of a solve (banded_solve_win_kernel):
//   * independent diagonal blocks (rows that no band entry couples: the per-column systems of a
//     grid without lateral processes, the column blocks of the probe preconditioner) are found at
//     set-up and solved by different CTAs (grid.y) — they are factored in parallel as well;
//   * members: MB <= 32 member lanes per CTA (coalesced rows of MB doubles), grid.x = B / MB;
//   * the band: RW row lanes share the kl (forward) / kl+ku (backward) updates of a step, so that a
//     single right-hand side (B = 1, the usual Krylov case) still uses a whole CTA.
//   The part of the right-hand side that a step can touch lives in a circular shared-memory window;
//   rows ahead of the sweep and the factor columns they need (stored transposed at set-up: the
//   multipliers of a step are contiguous) arrive through a 4-stage cp.async ring, so no global
//   load sits on the step-to-step dependency chain.  A^-1 (scale y) - y is fused into the backward
//   sweep.
//
// Band storage (row-major): A(i,j) lives at ab[(kv + i - j)*n + j], kv = kl + ku, rows
// 0..kl-1 are fill-in space, total 2*kl + ku + 1 rows.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "nkb_common.cuh"

struct nkb_banded {
    int n = 0, kl = 0, ku = 0;
    double *ab = nullptr;  // [(2kl+ku+1)][n]
    int *ipiv = nullptr;   // [n]
    int *info = nullptr;
    double *lt = nullptr;  // [n][kl]      multipliers of column j: L(j+1..j+kl, j)
    double *ut = nullptr;  // [n][kv+1]    {1/U(j,j), U(j-1,j), ..., U(j-kv,j)}
    int nblk = 1;          // independent diagonal blocks
    int *blk = nullptr;    // [nblk+1] first row of each block (device)
    // narrow bands factored without row interchanges (the per-column tridiagonal systems): compact
    // rows {L(j,j-K..j-1), 1/U(j,j), U(j,j+1..j+K)}, K = max(kl, ku) <= 4, for banded_thomas_kernel
    int nb_k = 0;
    double *nb = nullptr;  // [n][2K+1]
    // one wide block factored without row interchanges: panel (blocked) substitution, banded_solve_panel_kernel
    int klp = 0, kup = 0;      // row pitches of lp / up (kl, ku rounded up to even: 16-byte cp.async)
    double *lp = nullptr;      // [n][klp]  L(j+1..j+kl, j)
    double *up = nullptr;      // [n][kup]  U(j-1..j-ku, j)
    double *linv = nullptr;    // [ceil(n/PR)][PR][PR] inverse of the unit lower diagonal block of rows p*PR.. (top aligned)
    double *uinv = nullptr;    // [ceil(n/PR)][PR][PR] inverse of the upper diagonal block of rows ..n-1-q*PR (bottom aligned)
};

namespace nkb {

// one CTA per independent diagonal block [blk[b], blk[b+1]); *info (zeroed by the host) receives
// 1 + the first column with a zero pivot
__global__ void __launch_bounds__(256) banded_factor_kernel(double *__restrict__ ab, int *__restrict__ ipiv,
                                                            int n, int kl, int ku, const int *__restrict__ blk,
                                                            int *info) {
    __shared__ double s_val[256];
    __shared__ int s_idx[256];
    __shared__ int s_ju;
    const int kv = kl + ku;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int r0 = blk[blockIdx.x], r1 = blk[blockIdx.x + 1];
    if (tid == 0) s_ju = r0;
    __syncthreads();
    for (int j = r0; j < r1; ++j) {
        const int km = min(kl, r1 - 1 - j);
        // pivot search over the km+1 candidates of column j
        double best = -1.0;
        int besti = 0;
        for (int i = tid; i <= km; i += nt) {
            const double v = fabs(ab[(size_t)(kv + i) * n + j]);
            if (v > best) { best = v; besti = i; }
        }
        s_val[tid] = best;
        s_idx[tid] = besti;
        __syncthreads();
        for (int s = nt / 2; s > 0; s >>= 1) {
            if (tid < s) {
                const double o = s_val[tid + s];
                const int oi = s_idx[tid + s];
                if (o > s_val[tid] || (o == s_val[tid] && oi < s_idx[tid])) { s_val[tid] = o; s_idx[tid] = oi; }
            }
            __syncthreads();
        }
        const int jp = s_idx[0];
        const double pv = s_val[0];
        if (tid == 0) {
            ipiv[j] = j + jp;
            if (pv == 0.0) atomicCAS(info, 0, j + 1);
            if (pv != 0.0) s_ju = max(s_ju, min(j + ku + jp, r1 - 1));
        }
        __syncthreads();
        if (pv != 0.0) {
            const int ju = s_ju;
            if (jp != 0) {
                for (int c = j + tid; c <= ju; c += nt) {
                    double *p0 = ab + (size_t)(kv + j - c) * n + c;
                    double *p1 = ab + (size_t)(kv + j + jp - c) * n + c;
                    const double t = *p0; *p0 = *p1; *p1 = t;
                }
                __syncthreads();
            }
            const double inv = 1.0 / ab[(size_t)kv * n + j];
            __syncthreads();
            for (int i = 1 + tid; i <= km; i += nt) ab[(size_t)(kv + i) * n + j] *= inv;
            __syncthreads();
            const int ncol = ju - j;
            for (int w = tid; w < ncol * km; w += nt) {
                const int ci = w / km, i = 1 + w % km;
                const int c = j + 1 + ci;
                ab[(size_t)(kv + j + i - c) * n + c] -= ab[(size_t)(kv + i) * n + j] * ab[(size_t)(kv + j - c) * n + c];
            }
        }
        __syncthreads();
    }
}

// transposed copies of the factor for the solves: the coefficients of one step are contiguous
__global__ void banded_transpose_kernel(const double *__restrict__ ab, int n, int kl, int ku,
                                        double *__restrict__ lt, double *__restrict__ ut) {
    const int kv = kl + ku;
    const int j = blockIdx.x;
    for (int i = threadIdx.x; i < kl; i += blockDim.x) {
        const int row = j + 1 + i;
        lt[(size_t)j * kl + i] = (row < n) ? ab[(size_t)(kv + 1 + i) * n + j] : 0.0;
    }
    for (int i = threadIdx.x; i <= kv; i += blockDim.x) {
        double v = 0.0;
        if (i == 0) v = 1.0 / ab[(size_t)kv * n + j];
        else if (j - i >= 0) v = ab[(size_t)(kv - i) * n + j];
        ut[(size_t)j * (kv + 1) + i] = v;
    }
}

__device__ __forceinline__ void bs_cp8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src)
                 : "memory");
}
__device__ __forceinline__ void bs_cp4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src)
                 : "memory");
}
__device__ __forceinline__ void bs_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bs_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Batched Thomas / narrow-band substitution, members fastest: one warp per (32 members, diagonal
// block), a lane owns one member; many small CTAs per SM hide the latency of the recurrences.
// Forward: y streams from HBM in chunks of UR rows (all loads of a chunk in flight before the first
// use); z = L^-1 y goes through the output buffer (it comes back from L2: keeping it in shared
// memory instead costs occupancy and measured 25 % slower).  Backward: x = U^-1 z and the epilogue
// scale*x - y, where y_j is rebuilt from the z values (y_j = z_j + sum_i L(j,j-i) z_(j-i)) instead of
// being read again.
template <int K, int UR>
__global__ void __launch_bounds__(32) banded_thomas_kernel(const double *__restrict__ nb, const int *__restrict__ blk,
                                                           const double *__restrict__ y, double *__restrict__ x,
                                                           int B, size_t ldb, double scale, int subtract) {
    constexpr int NC = 2 * K + 1;
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= B) return;
    const int r0 = blk[blockIdx.y], r1 = blk[blockIdx.y + 1];
    double zp[K];
#pragma unroll
    for (int i = 0; i < K; ++i) zp[i] = 0.0;
    for (int j0 = r0; j0 < r1; j0 += UR) {
        double yv[UR], cl[UR][K];
#pragma unroll
        for (int q = 0; q < UR; ++q) {
            const int j = min(j0 + q, r1 - 1);
            yv[q] = __ldcs(y + (size_t)j * ldb + b);
#pragma unroll
            for (int i = 0; i < K; ++i) cl[q][i] = __ldg(nb + (size_t)j * NC + i);
        }
#pragma unroll
        for (int q = 0; q < UR; ++q) {
            if (j0 + q < r1) {
                double z = yv[q];
#pragma unroll
                for (int i = 1; i <= K; ++i) z = fma(-cl[q][K - i], zp[i - 1], z);
#pragma unroll
                for (int i = K - 1; i > 0; --i) zp[i] = zp[i - 1];
                zp[0] = z;
                x[(size_t)(j0 + q) * ldb + b] = z;
            }
        }
    }
    // backward; zc[i] = z_(j-i) (zero above the block), xp[i] = x_(j+1+i)
    auto zload = [&](int j) -> double { return (j < r0) ? 0.0 : x[(size_t)j * ldb + b]; };
    double xp[K], zc[K + 1];
#pragma unroll
    for (int i = 0; i < K; ++i) xp[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= K; ++i) zc[i] = zload(r1 - 1 - i);
    for (int j0 = r1 - 1; j0 >= r0; j0 -= UR) {
        double cu[UR][NC], zn[UR];
#pragma unroll
        for (int q = 0; q < UR; ++q) {
            const int j = max(j0 - q, r0);
#pragma unroll
            for (int i = 0; i < NC; ++i) cu[q][i] = __ldg(nb + (size_t)j * NC + i);
            zn[q] = zload(j0 - q - K - 1);  // enters the z window after row j0 - q (read before any store of
                                            // this chunk can reach it)
        }
#pragma unroll
        for (int q = 0; q < UR; ++q) {
            const int j = j0 - q;
            if (j >= r0) {
                double t = zc[0];
#pragma unroll
                for (int i = 1; i <= K; ++i) t = fma(-cu[q][K + i], xp[i - 1], t);
                const double xj = t * cu[q][K];
                double o = scale * xj;
                if (subtract) {
                    double yj = zc[0];
#pragma unroll
                    for (int i = 1; i <= K; ++i) yj = fma(cu[q][K - i], zc[i], yj);
                    o -= yj;
                }
                __stcs(x + (size_t)j * ldb + b, o);
#pragma unroll
                for (int i = K - 1; i > 0; --i) xp[i] = xp[i - 1];
                xp[0] = xj;
#pragma unroll
                for (int i = 0; i < K; ++i) zc[i] = zc[i + 1];
                zc[K] = zn[q];
            }
        }
    }
}

template <int K>
static int launch_thomas(const nkb_banded *f, const double *y, double *x, int B, size_t ldb, double scale,
                         int subtract, cudaStream_t st) {
    // in place (x == y, no subtraction) is safe: row j of x is written after row j of y has been read, and
    // every row is visited once per sweep
    dim3 grid((B + 31) / 32, f->nblk);
    banded_thomas_kernel<K, 8><<<grid, 32, 0, st>>>(f->nb, f->blk, y, x, B, ldb, scale, subtract);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

constexpr int BS_NSTG = 4;  // cp.async stages in flight

struct BandedSolveArgs {
    const double *lt, *ut;
    const int *ipiv, *blk;
    int kl, ku;
    const double *y;
    double *x;
    int B;
    size_t ldb;
    double scale;
    int subtract;
    int MB, RW, CH, Wn;  // member lanes, row lanes, steps per stage, window rows
};

// threads = RW row lanes x MB member lanes (member fastest); grid (member groups, diagonal blocks).
// Shared memory: W[Wn][MB] window | C[NSTG][CH*(kv+1)] factor columns | Y[NSTG][CH][MB] | P[NSTG][CH]
__global__ void __launch_bounds__(256) banded_solve_win_kernel(const BandedSolveArgs a) {
    extern __shared__ __align__(16) double bs_smem[];
    const int MB = a.MB, RW = a.RW, CH = a.CH, Wn = a.Wn, kl = a.kl, kv = a.kl + a.ku;
    double *W = bs_smem;
    double *C = W + (size_t)Wn * MB;
    const int cstride = CH * (kv + 1);
    double *Y = C + (size_t)BS_NSTG * cstride;
    int *P = reinterpret_cast<int *>(Y + (size_t)BS_NSTG * CH * MB);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int m = tid % MB, rl = tid / MB;
    const int b0 = blockIdx.x * MB;
    const bool live = (rl == 0) && (b0 + m < a.B);
    const int nlive = min(MB, a.B - b0);
    const int r0 = a.blk[blockIdx.y], r1 = a.blk[blockIdx.y + 1];
    const int len = r1 - r0;
    for (int i = tid; i < Wn * MB; i += nt) W[i] = 0.0;
    __syncthreads();

    // rows [ra, rb) of src -> window slots (row - r0) % Wn
    auto load_rows = [&](const double *src, int ra, int rb) {
        ra = max(ra, r0);
        rb = min(rb, r1);
        const int cnt = (rb - ra) * MB;
        for (int i = tid; i < cnt; i += nt) {
            const int r = ra + i / MB, mm = i % MB;
            if (mm < nlive) bs_cp8(W + (size_t)((r - r0) % Wn) * MB + mm, src + (size_t)r * a.ldb + b0 + mm);
        }
    };

    // ---------------- forward: L z = P y (scale is applied at the very end: the solve is linear) ----------------
    const int nchunk = (len + CH - 1) / CH;
    auto stage_fwd = [&](int q) {
        if (q < nchunk) {
            const int jq = r0 + q * CH, je = min(jq + CH, r1);
            const double *src = a.lt + (size_t)jq * kl;
            double *dst = C + (size_t)(q % BS_NSTG) * cstride;
            for (int i = tid; i < (je - jq) * kl; i += nt) bs_cp8(dst + i, src + i);
            for (int i = tid; i < je - jq; i += nt) bs_cp4(P + (q % BS_NSTG) * CH + i, a.ipiv + jq + i);
            load_rows(a.y, q == 0 ? r0 : jq + kl, jq + kl + CH);
        }
        bs_commit();
    };
    for (int q = 0; q < BS_NSTG - 1; ++q) stage_fwd(q);
    for (int q = 0; q < nchunk; ++q) {
        stage_fwd(q + BS_NSTG - 1);
        bs_wait<BS_NSTG - 1>();
        __syncthreads();
        const int jq = r0 + q * CH, je = min(jq + CH, r1);
        const double *cq = C + (size_t)(q % BS_NSTG) * cstride;
        const int *pq = P + (q % BS_NSTG) * CH;
        int sj = (jq - r0) % Wn;
        for (int j = jq; j < je; ++j) {
            const int p = pq[j - jq];
            if (p != j) {  // uniform over the CTA
                if (rl == 0) {
                    int sp = sj + (p - j);
                    if (sp >= Wn) sp -= Wn;
                    const double t = W[(size_t)sj * MB + m];
                    W[(size_t)sj * MB + m] = W[(size_t)sp * MB + m];
                    W[(size_t)sp * MB + m] = t;
                }
                __syncthreads();
            }
            const double xj = W[(size_t)sj * MB + m];
            const int lm = min(kl, r1 - 1 - j);
            const double *cj = cq + (size_t)(j - jq) * kl;
            for (int i = 1 + rl; i <= lm; i += RW) {
                int si = sj + i;
                if (si >= Wn) si -= Wn;
                W[(size_t)si * MB + m] = fma(-cj[i - 1], xj, W[(size_t)si * MB + m]);
            }
            if (live) a.x[(size_t)j * a.ldb + b0 + m] = xj;
            if (++sj == Wn) sj = 0;
            __syncthreads();
        }
    }
    bs_wait<0>();
    __threadfence_block();
    __syncthreads();

    // ---------------- backward: U x = z, out = scale * x [- y] ----------------
    auto stage_bwd = [&](int q) {
        if (q < nchunk) {
            const int jq = r1 - 1 - q * CH, jl = max(jq - CH + 1, r0);  // steps jq, jq-1, ..., jl
            const double *src = a.ut + (size_t)jl * (kv + 1);
            double *dst = C + (size_t)(q % BS_NSTG) * cstride;
            for (int i = tid; i < (jq - jl + 1) * (kv + 1); i += nt) bs_cp8(dst + i, src + i);
            if (a.subtract) {
                double *yd = Y + (size_t)(q % BS_NSTG) * CH * MB;
                for (int i = tid; i < (jq - jl + 1) * MB; i += nt) {
                    const int mm = i % MB;
                    if (mm < nlive) bs_cp8(yd + i, a.y + (size_t)(jl + i / MB) * a.ldb + b0 + mm);
                }
            }
            load_rows(a.x, jq - CH + 1 - kv, q == 0 ? r1 : jq - kv + 1);
        }
        bs_commit();
    };
    for (int q = 0; q < BS_NSTG - 1; ++q) stage_bwd(q);
    for (int q = 0; q < nchunk; ++q) {
        stage_bwd(q + BS_NSTG - 1);
        bs_wait<BS_NSTG - 1>();
        __syncthreads();
        const int jq = r1 - 1 - q * CH, jl = max(jq - CH + 1, r0);
        const double *cq = C + (size_t)(q % BS_NSTG) * cstride;
        const double *yq = Y + (size_t)(q % BS_NSTG) * CH * MB;
        int sj = (jq - r0) % Wn;
        for (int j = jq; j >= jl; --j) {
            const double *cj = cq + (size_t)(j - jl) * (kv + 1);
            const double xj = W[(size_t)sj * MB + m] * cj[0];
            const int um = min(kv, j - r0);
            for (int i = 1 + rl; i <= um; i += RW) {
                int si = sj - i;
                if (si < 0) si += Wn;
                W[(size_t)si * MB + m] = fma(-cj[i], xj, W[(size_t)si * MB + m]);
            }
            if (live) {
                double o = a.scale * xj;
                if (a.subtract) o -= yq[(size_t)(j - jl) * MB + m];
                a.x[(size_t)j * a.ldb + b0 + m] = o;
            }
            if (--sj < 0) sj = Wn - 1;
            __syncthreads();
        }
    }
    bs_wait<0>();
}

// ---- panel (blocked) substitution for ONE wide block factored without row interchanges -------------------
// The window kernel above advances one row per block barrier (0.43 us per row: 16 ms for the refined grid's 2-D
// preconditioner, n = 18 750, kl = ku = 450, every Krylov iteration).  Without interchanges L and U keep their
// bandwidths (kl, ku) and the sweeps can advance PR rows per barrier pair:
//   (1) the PR unknowns of a panel from the panel's right-hand side with the INVERSE of the PR x PR diagonal
//       block (formed once at set-up, banded_panel_inverse_kernel: the matrices here are diagonally dominant
//       I - dt J products, their triangular blocks well conditioned) — a small dense product, no recurrence;
//   (2) the kl (ku) rows below (above) the panel updated with the panel's PR columns at once, one
//       (row, member) pair per thread and pass, no reduction.
// The factor columns of a panel are ONE contiguous run of PR*klp doubles (lp / up are stored column by column)
// and arrive with the panel's inverse block and the rows entering the window through a cp.async ring.
constexpr int PR = 16;

struct PanelArgs {
    const double *lp, *up, *linv, *uinv;
    int n, kl, ku, klp, kup;
    const double *y;
    double *x;
    int B;
    size_t ldb;
    double scale;
    int subtract;
    int MB, Wn, NS, cw;  // member lanes, window rows, ring stages, doubles per factor column in the ring
};

__device__ __forceinline__ void bs_cp16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src)
                 : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion counted on an mbarrier: the PR factor columns
// of a panel are one contiguous run (57.6 KB for kl = 450) — as 16-byte cp.async requests of 256 threads a single
// SM fetched them at only 18 GB/s (3.3 us per panel, measured)
__device__ __forceinline__ void bs_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bs_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bs_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bs_wait_dyn(int pending) {  // pending = NS - 2 in {0, 1}
    if (pending >= 1) bs_wait<1>(); else bs_wait<0>();
}

// compact copies of the factor for the panel kernel (column j contiguous, zero padded)
__global__ void banded_panel_pack_kernel(const double *__restrict__ ab, int n, int kl, int ku, int klp, int kup,
                                         double *__restrict__ lp, double *__restrict__ up) {
    const int kv = kl + ku;
    const int j = blockIdx.x;
    for (int d = threadIdx.x; d < klp; d += blockDim.x)
        lp[(size_t)j * klp + d] = (d < kl && j + 1 + d < n) ? ab[(size_t)(kv + 1 + d) * n + j] : 0.0;
    for (int d = threadIdx.x; d < kup; d += blockDim.x)
        up[(size_t)j * kup + d] = (d < ku && j - 1 - d >= 0) ? ab[(size_t)(kv - 1 - d) * n + j] : 0.0;
}

// inverses of the PR x PR diagonal blocks: block p of L covers rows p*PR.. (top aligned), block q of U covers
// rows ..n-1-q*PR (bottom aligned); thread c solves for column c; short blocks are padded with the identity
__global__ void __launch_bounds__(2 * PR) banded_panel_inverse_kernel(const double *__restrict__ ab,
                                                                      const double *__restrict__ lp,
                                                                      const double *__restrict__ up, int n, int kl,
                                                                      int ku, int klp, int kup,
                                                                      double *__restrict__ linv,
                                                                      double *__restrict__ uinv) {
    const int blk = blockIdx.x, c = threadIdx.x % PR;
    const int kv = kl + ku;
    double x[PR];
    if (threadIdx.x < PR) {
        const int j0 = blk * PR, nr = min(PR, n - j0);
        double *out = linv + (size_t)blk * PR * PR;
        for (int ii = 0; ii < PR; ++ii) x[ii] = (ii == c) ? 1.0 : 0.0;
        if (c < nr) {
            for (int ii = c + 1; ii < nr; ++ii) {
                double acc = 0.0;
                for (int jj = c; jj < ii; ++jj) {
                    const int d = ii - jj - 1;
                    if (d < kl) acc = fma(lp[(size_t)(j0 + jj) * klp + d], x[jj], acc);
                }
                x[ii] = -acc;
            }
        }
        for (int ii = 0; ii < PR; ++ii) out[ii * PR + c] = x[ii];
    } else {
        const int j1 = n - 1 - blk * PR, jl = max(j1 - PR + 1, 0), nr = j1 - jl + 1;
        double *out = uinv + (size_t)blk * PR * PR;
        for (int ii = 0; ii < PR; ++ii) x[ii] = 0.0;
        if (c < nr) {
            x[c] = 1.0 / ab[(size_t)kv * n + jl + c];
            for (int ii = c - 1; ii >= 0; --ii) {
                double acc = 0.0;
                for (int jj = ii + 1; jj <= c; ++jj) {
                    const int d = jj - ii - 1;
                    if (d < ku) acc = fma(up[(size_t)(jl + jj) * kup + d], x[jj], acc);
                }
                x[ii] = -acc / ab[(size_t)kv * n + jl + ii];
            }
        } else {
            x[c] = 1.0;
        }
        for (int ii = 0; ii < PR; ++ii) out[ii * PR + c] = x[ii];
    }
}

// threads = 256 = RW row lanes x MB member lanes (member fastest, MB a power of two); grid = member groups.
// Shared memory: W[Wn][MB] window | X[PR][MB] panel unknowns | C[NS][PR*cw] factor columns | D[NS][PR*PR] inverse
// blocks | Y[NS][PR][MB] right-hand side rows for the epilogue.
// Ring discipline: wait for panel q's stage, block barrier (every thread has left panel q-1: its ring stage and
// the window rows behind the sweep may be overwritten), THEN issue the stage of panel q+NS-1.
__global__ void __launch_bounds__(256) banded_solve_panel_kernel(const PanelArgs a) {
    extern __shared__ __align__(16) double bs_smem[];
    const int MB = a.MB, Wn = a.Wn, NS = a.NS, n = a.n, kl = a.kl, ku = a.ku;
    double *W = bs_smem;
    double *X = W + (size_t)Wn * MB;
    double *C = X + PR * MB;
    double *D = C + (size_t)NS * PR * a.cw;
    double *Y = D + (size_t)NS * PR * PR;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(Y + (size_t)NS * PR * MB);  // NS mbarriers
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {
        for (int sidx = 0; sidx < NS; ++sidx) bs_mbar_init(bar0 + 8 * sidx, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int msh = 31 - __clz(MB);
    const int m = tid & (MB - 1), rl = tid >> msh, RW = nt >> msh;
    const int b0 = blockIdx.x * MB;
    const bool live = (b0 + m < a.B);
    const int nlive = min(MB, a.B - b0);
    for (int i = tid; i < Wn * MB; i += nt) W[i] = 0.0;
    __syncthreads();

    auto slot = [&](int r) {  // window slot of row r >= 0 (r - Wn*floor(r/Wn) without the division in the hot loops)
        return r % Wn;
    };
    auto load_rows = [&](const double *src, int ra, int rb) {  // rows [ra, rb) -> window
        ra = max(ra, 0);
        rb = min(rb, n);
        if (rb <= ra) return;
        int s0 = slot(ra);
        for (int i = tid; i < ((rb - ra) << msh); i += nt) {
            const int dr = i >> msh, mm = i & (MB - 1);
            int sl = s0 + dr;
            if (sl >= Wn) sl -= Wn;
            if (mm < nlive) bs_cp8(W + ((size_t)sl << msh) + mm, src + (size_t)(ra + dr) * a.ldb + b0 + mm);
        }
    };
    const int nP = (n + PR - 1) / PR;
    const int half = (PR * PR) / 2;
    // phase (1) roles: tid = ii*16 + part*MB + m — P = 16/MB lanes per (panel row, member) pair
    const int P = PR >> msh;
    const int p_ii = tid >> 4, p_part = (tid & 15) >> msh;
    // factor columns of ring use g (forward panels 0..nP-1, backward nP..2nP-1): ring stage g % NS keeps the
    // forward and the backward sweep on ONE phase sequence per mbarrier (backward panel q uses stage (nP + q) % NS)
    auto bulk = [&](double *dst, const double *src, uint32_t bytes, int g) {
        const uint32_t bar = bar0 + 8 * (g % NS);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was read through the generic proxy
        bs_mbar_expect_tx(bar, bytes);
        uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
        const char *sp = reinterpret_cast<const char *>(src);
        while (bytes > 0) {  // pieces of at most 32 KB
            const uint32_t piece = bytes > 32768u ? 32768u : bytes;
            bs_bulk_load(d, sp, piece, bar);
            d += piece; sp += piece; bytes -= piece;
        }
    };

    // ---------------- forward: L z = y ----------------
    auto stage_fwd = [&](int q) {
        if (q < nP) {
            const int j0 = q * PR, nr = min(PR, n - j0);
            const double *src = a.lp + (size_t)j0 * a.klp;
            double *dst = C + (size_t)(q % NS) * PR * a.cw;
            if (tid == 0) bulk(dst, src, (uint32_t)nr * a.klp * 8u, q);
            const double *dsrc = a.linv + (size_t)q * PR * PR;
            double *ddst = D + (size_t)(q % NS) * PR * PR;
            if (tid < half) bs_cp16(ddst + 2 * tid, dsrc + 2 * tid);
            load_rows(a.y, q == 0 ? 0 : j0 + kl, j0 + kl + PR);
        }
        bs_commit();
    };
    for (int q = 0; q < NS - 1; ++q) stage_fwd(q);
    for (int q = 0; q < nP; ++q) {
        bs_wait_dyn(NS - 2);
        bs_mbar_wait(bar0 + 8 * (q % NS), (q / NS) & 1);
        __syncthreads();
        stage_fwd(q + NS - 1);
        const int j0 = q * PR, nr = min(PR, n - j0);
        const double *cq = C + (size_t)(q % NS) * PR * a.cw;
        const double *dq = D + (size_t)(q % NS) * PR * PR;
        const int s0 = slot(j0);
        // (1) panel unknowns x = Linv w on all 256 threads: 16/MB lanes share the 16 products of a (row, member)
        //     pair and combine them by xor shuffles (the inverse block is stored with its zeros, no triangle tests)
        {
            double acc = 0.0;
            if (p_ii < nr) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int jj = p_part + t * P;
                    if (t < MB && jj < nr) {
                        int sl = s0 + jj;
                        if (sl >= Wn) sl -= Wn;
                        acc = fma(dq[p_ii * PR + jj], W[((size_t)sl << msh) + m], acc);
                    }
                }
            }
            for (int off = MB; off < PR; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (p_part == 0 && p_ii < nr) {
                X[(p_ii << msh) + m] = acc;
                if (live) a.x[(size_t)(j0 + p_ii) * a.ldb + b0 + m] = acc;
            }
        }
        __syncthreads();
        // (2) the kl rows below the panel; the panel's unknowns of this thread's member sit in registers, a row
        //     whose 16 coefficients all lie inside the band takes the unrolled path (all loads ahead of the FMAs)
        {
            double xr[PR];
#pragma unroll
            for (int jj = 0; jj < PR; ++jj) xr[jj] = (jj < nr) ? X[(jj << msh) + m] : 0.0;
            int sbase = s0 + nr;
            if (sbase >= Wn) sbase -= Wn;
            const int nrow = min(kl, n - (j0 + nr));
            for (int ir = rl; ir < nrow; ir += RW) {
                int sl = sbase + ir;
                if (sl >= Wn) sl -= Wn;
                double *w = W + ((size_t)sl << msh) + m;
                const int dmax = ir + nr - 1;  // d = dmax - jj
                if (nr == PR && dmax < kl) {
                    const double *c0 = cq + dmax;
                    double a0 = *w, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                    for (int jj = 0; jj < PR; jj += 4) {
                        a0 = fma(-c0[(size_t)jj * (a.klp - 1)], xr[jj], a0);
                        a1 = fma(-c0[(size_t)(jj + 1) * (a.klp - 1)], xr[jj + 1], a1);
                        a2 = fma(-c0[(size_t)(jj + 2) * (a.klp - 1)], xr[jj + 2], a2);
                        a3 = fma(-c0[(size_t)(jj + 3) * (a.klp - 1)], xr[jj + 3], a3);
                    }
                    *w = (a0 + a1) + (a2 + a3);
                } else {
                    double acc = *w;
                    for (int jj = max(0, dmax - kl + 1); jj < nr; ++jj)
                        acc = fma(-cq[(size_t)jj * a.klp + dmax - jj], X[(jj << msh) + m], acc);
                    *w = acc;
                }
            }
        }
    }
    bs_wait<0>();
    __threadfence_block();
    __syncthreads();

    // ---------------- backward: U x = z, out = scale * x [- y] ----------------
    auto stage_bwd = [&](int q) {
        if (q < nP) {
            const int j1 = n - 1 - q * PR, jl = max(j1 - PR + 1, 0), nr = j1 - jl + 1;
            const double *src = a.up + (size_t)jl * a.kup;
            double *dst = C + (size_t)((nP + q) % NS) * PR * a.cw;
            if (tid == 0) bulk(dst, src, (uint32_t)nr * a.kup * 8u, nP + q);
            const double *dsrc = a.uinv + (size_t)q * PR * PR;
            double *ddst = D + (size_t)(q % NS) * PR * PR;
            if (tid < half) bs_cp16(ddst + 2 * tid, dsrc + 2 * tid);
            if (a.subtract) {
                double *yd = Y + ((size_t)(q % NS) * PR << msh);
                for (int i = tid; i < (nr << msh); i += nt) {
                    const int mm = i & (MB - 1);
                    if (mm < nlive) bs_cp8(yd + i, a.y + (size_t)(jl + (i >> msh)) * a.ldb + b0 + mm);
                }
            }
            load_rows(a.x, j1 - PR + 1 - ku, q == 0 ? n : j1 + 1 - ku);
        }
        bs_commit();
    };
    for (int q = 0; q < NS - 1; ++q) stage_bwd(q);
    for (int q = 0; q < nP; ++q) {
        bs_wait_dyn(NS - 2);
        bs_mbar_wait(bar0 + 8 * ((nP + q) % NS), ((nP + q) / NS) & 1);
        __syncthreads();
        stage_bwd(q + NS - 1);
        const int j1 = n - 1 - q * PR, jl = max(j1 - PR + 1, 0), nr = j1 - jl + 1;
        const double *cq = C + (size_t)((nP + q) % NS) * PR * a.cw;
        const double *dq = D + (size_t)(q % NS) * PR * PR;
        const double *yq = Y + ((size_t)(q % NS) * PR << msh);
        const int s0 = slot(jl);
        // (1) panel unknowns x = Uinv w (as in the forward sweep), and the output rows of the panel
        {
            double acc = 0.0;
            if (p_ii < nr) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int jj = p_part + t * P;
                    if (t < MB && jj < nr) {
                        int sl = s0 + jj;
                        if (sl >= Wn) sl -= Wn;
                        acc = fma(dq[p_ii * PR + jj], W[((size_t)sl << msh) + m], acc);
                    }
                }
            }
            for (int off = MB; off < PR; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (p_part == 0 && p_ii < nr) {
                X[(p_ii << msh) + m] = acc;
                if (live) {
                    double o = a.scale * acc;
                    if (a.subtract) o -= yq[(p_ii << msh) + m];
                    a.x[(size_t)(jl + p_ii) * a.ldb + b0 + m] = o;
                }
            }
        }
        __syncthreads();
        // (2) the ku rows above the panel: U(i, jl + jj) = up[jl + jj][jl + jj - i - 1], i = jl - 1 - ir
        {
            double xr[PR];
#pragma unroll
            for (int jj = 0; jj < PR; ++jj) xr[jj] = (jj < nr) ? X[(jj << msh) + m] : 0.0;
            const int nrow = min(ku, jl);
            for (int ir = rl; ir < nrow; ir += RW) {
                int sl = s0 - 1 - ir;
                if (sl < 0) sl += Wn;
                double *w = W + ((size_t)sl << msh) + m;
                if (nr == PR && ir + PR <= ku) {  // d = ir + jj < ku for every column of the panel
                    const double *c0 = cq + ir;
                    double a0 = *w, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                    for (int jj = 0; jj < PR; jj += 4) {
                        a0 = fma(-c0[(size_t)jj * (a.kup + 1)], xr[jj], a0);
                        a1 = fma(-c0[(size_t)(jj + 1) * (a.kup + 1)], xr[jj + 1], a1);
                        a2 = fma(-c0[(size_t)(jj + 2) * (a.kup + 1)], xr[jj + 2], a2);
                        a3 = fma(-c0[(size_t)(jj + 3) * (a.kup + 1)], xr[jj + 3], a3);
                    }
                    *w = (a0 + a1) + (a2 + a3);
                } else {
                    double acc = *w;
                    const int jhi = min(nr, ku - ir);
                    for (int jj = 0; jj < jhi; ++jj)
                        acc = fma(-cq[(size_t)jj * a.kup + ir + jj], X[(jj << msh) + m], acc);
                    *w = acc;
                }
            }
        }
    }
    bs_wait<0>();
}

// Whole-device factorisation of ONE wide diagonal block (the 2-D preconditioners with lateral processes:
// n = nz*ny rows, kl = ku = 3*ny): the same unblocked right-looking elimination, but the rank-1 update of a
// column step (up to kl x (kl+ku) entries) is spread over all CTAs of a cooperative launch, walking the
// band storage along its rows (coalesced).  Two grid barriers per column: every CTA first copies the
// pivot row and the multiplier column it needs into shared memory (pivot search redundantly per CTA), then
// the swap / scale / update writes start.  refined 125 x 150 grid (n = 18 750, kl = ku = 450): seconds with
// one CTA, ~0.1 s here.
__global__ void __launch_bounds__(256) banded_factor_coop_kernel(double *__restrict__ ab, int *__restrict__ ipiv,
                                                                 int n, int kl, int ku, int *info) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double fc_sh[];  // lcol[kl + 1] | urow[kl + ku + 1] | rowj[kl + ku + 1]
    __shared__ double s_val[256];
    __shared__ int s_idx[256];
    const int kv = kl + ku;
    double *lcol = fc_sh, *urow = fc_sh + kl + 1, *rowj = urow + kv + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int gwarp = (blockIdx.x * nt + tid) >> 5, nwarp = (gridDim.x * nt) >> 5, lane = tid & 31;
    int ju = 0;
    for (int j = 0; j < n; ++j) {
        const int km = min(kl, n - 1 - j);
        // ---- read phase: pivot search over column j (every CTA does it: same result) ----
        double best = -1.0;
        int besti = 0;
        for (int i = tid; i <= km; i += nt) {
            const double v = fabs(ab[(size_t)(kv + i) * n + j]);
            if (v > best) { best = v; besti = i; }
        }
        s_val[tid] = best;
        s_idx[tid] = besti;
        __syncthreads();
        for (int sft = nt / 2; sft > 0; sft >>= 1) {
            if (tid < sft) {
                const double o = s_val[tid + sft];
                const int oi = s_idx[tid + sft];
                if (o > s_val[tid] || (o == s_val[tid] && oi < s_idx[tid])) { s_val[tid] = o; s_idx[tid] = oi; }
            }
            __syncthreads();
        }
        const int jp = s_idx[0];
        const double pv = s_val[0];
        if (pv != 0.0) ju = max(ju, min(j + ku + jp, n - 1));
        const int ncol = ju - j;  // columns j+1 .. ju are updated
        if (pv != 0.0) {
            // the pivot row (old row j + jp) and old row j over the columns j .. ju, the multiplier column in
            // its order AFTER the interchange, scaled
            const double inv = 1.0 / ab[(size_t)(kv + jp) * n + j];
            for (int e = tid; e <= ncol; e += nt) {
                const int c = j + e;
                urow[e] = (jp - e >= -kv) ? ab[(size_t)(kv + jp - e) * n + c] : 0.0;
                rowj[e] = ab[(size_t)(kv - e) * n + c];
            }
            for (int i = 1 + tid; i <= km; i += nt) {
                const double v = (i == jp) ? ab[(size_t)kv * n + j] : ab[(size_t)(kv + i) * n + j];
                lcol[i] = v * inv;
            }
        }
        __syncthreads();
        grid.sync();
        // ---- write phase ----
        if (blockIdx.x == 0 && tid == 0) {
            ipiv[j] = j + jp;
            if (pv == 0.0) atomicCAS(info, 0, j + 1);
        }
        if (pv != 0.0) {
            if (blockIdx.x == 0) {
                // column j below the diagonal (the entry of row j + jp is a multiplier too) and row j, which
                // receives the pivot row; the old row j moves to row j + jp through the update below
                for (int i = 1 + tid; i <= km; i += nt) ab[(size_t)(kv + i) * n + j] = lcol[i];
                if (jp != 0)
                    for (int e = tid; e <= ncol; e += nt) ab[(size_t)(kv - e) * n + j + e] = urow[e];
            }
            // A(j+i, j+e) -= l_i * u_e for i = 1..km, e = 1..ncol, one band row (dd = i - e) per warp.  Row j+jp
            // holds the old row j after the interchange: its new value is formed from rowj (this is the
            // only write to that row in this step).
            for (int dd = 1 - ncol + gwarp; dd <= km - 1; dd += nwarp) {
                const int e_lo = max(1, 1 - dd), e_hi = min(ncol, km - dd);
                double *row = ab + (size_t)(kv + dd) * n + j;
                for (int e = e_lo + lane; e <= e_hi; e += 32) {
                    const int i = dd + e;
                    const double old = (jp != 0 && i == jp) ? rowj[e] : row[e];
                    row[e] = fma(-lcol[i], urow[e], old);
                }
            }
        }
        grid.sync();
    }
}

// fallback for bands too wide for the shared-memory window: one thread per member; x [n][ldb] in place
__global__ void banded_solve_kernel(const double *__restrict__ ab, const int *__restrict__ ipiv, int n, int kl,
                                    int ku, const double *__restrict__ y, double *__restrict__ x, int B,
                                    size_t ldb, double scale, int subtract_rhs) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int kv = kl + ku;
    double *xb = x + b;
    const double *yb = y + b;
    if (x != y || scale != 1.0)
        for (int j = 0; j < n; ++j) xb[(size_t)j * ldb] = scale * yb[(size_t)j * ldb];
    // forward: L y = P b
    for (int j = 0; j < n; ++j) {
        const int lm = min(kl, n - 1 - j);
        const int p = ipiv[j];
        double xj = xb[(size_t)j * ldb];
        if (p != j) {
            const double t = xb[(size_t)p * ldb];
            xb[(size_t)p * ldb] = xj;
            xb[(size_t)j * ldb] = t;
            xj = t;
        }
        for (int i = 1; i <= lm; ++i) xb[(size_t)(j + i) * ldb] -= ab[(size_t)(kv + i) * n + j] * xj;
    }
    // backward: U x = y
    for (int j = n - 1; j >= 0; --j) {
        const double xj = xb[(size_t)j * ldb] / ab[(size_t)kv * n + j];
        xb[(size_t)j * ldb] = xj;
        const int um = min(kv, j);
        for (int i = 1; i <= um; ++i) xb[(size_t)(j - i) * ldb] -= ab[(size_t)(kv - i) * n + j] * xj;
    }
    if (subtract_rhs) {
        // res = A^-1 (scale*y) - y.  y may alias x only when subtract_rhs == 0
        for (int j = 0; j < n; ++j) xb[(size_t)j * ldb] -= yb[(size_t)j * ldb];
    }
}

}  // namespace nkb

extern "C" {

int nkb_banded_create(nkb_banded **out, int n, int kl, int ku, const double *h_ab) {
    NKB_REQUIRE(out && h_ab && n >= 1 && kl >= 0 && ku >= 0, "nkb_banded_create: bad argument");
    nkb_banded *f = new nkb_banded();
    struct Guard {  // the handle is released on every early return below
        nkb_banded *f;
        ~Guard() { if (f) nkb_banded_destroy(f); }
    } guard{f};
    f->n = n; f->kl = kl; f->ku = ku;
    const int rows = 2 * kl + ku + 1;
    std::vector<double> host((size_t)rows * n, 0.0);
    // scipy solve_banded layout in: ab_in[ku + i - j][j]  ->  rows kl.. of the working band
    for (int r = 0; r < kl + ku + 1; ++r)
        for (int j = 0; j < n; ++j) host[(size_t)(kl + r) * n + j] = h_ab[(size_t)r * n + j];
    // independent diagonal blocks: a boundary in front of row b is "cut" by every non-zero A(i,c) with
    // min(i,c) < b <= max(i,c); rows between two uncut boundaries form a system of their own
    std::vector<int> cut(n + 1, 0);
    for (int r = 0; r < kl + ku + 1; ++r) {
        const int d = r - ku;  // i - c
        if (d == 0) continue;
        for (int c = 0; c < n; ++c) {
            const int i = c + d;
            if (i < 0 || i >= n || h_ab[(size_t)r * n + c] == 0.0) continue;
            cut[std::min(i, c) + 1] += 1;
            cut[std::max(i, c) + 1] -= 1;
        }
    }
    std::vector<int> blk(1, 0);
    for (int b = 1, acc = 0; b < n; ++b) {
        acc += cut[b];
        if (acc == 0) blk.push_back(b);
    }
    blk.push_back(n);
    f->nblk = (int)blk.size() - 1;
    const int kv = kl + ku;
    if (cudaMalloc(&f->ab, host.size() * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&f->ipiv, n * sizeof(int)) != cudaSuccess || cudaMalloc(&f->info, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&f->lt, std::max<size_t>(1, (size_t)n * kl) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&f->ut, (size_t)n * (kv + 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&f->blk, blk.size() * sizeof(int)) != cudaSuccess) {
        nkb::set_error("nkb_banded_create: cudaMalloc failed (is a CUDA device present?)");
        return 1;
    }
    NKB_CUDA(cudaMemcpy(f->ab, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
    NKB_CUDA(cudaMemcpy(f->blk, blk.data(), blk.size() * sizeof(int), cudaMemcpyHostToDevice));
    NKB_CUDA(cudaMemset(f->info, 0, sizeof(int)));
    // one wide block: all SMs share the column updates (cooperative launch); many blocks: one CTA each
    bool coop_done = false;
    {
        const char *env = getenv("NKB_BANDED_COOP");
        const bool want = (env && *env) ? (env[0] != '0') : ((double)n * kl * (kv + 1) > 2.0e6);
        if (f->nblk == 1 && kl >= 1 && want) {
            int dev = 0, n_sm = 0, coop = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
            const size_t smem = (size_t)(kl + 1 + 2 * (kv + 1)) * sizeof(double);
            if (coop && smem <= 96 * 1024 &&
                cudaFuncSetAttribute(nkb::banded_factor_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nkb::banded_factor_coop_kernel, 256, smem) ==
                    cudaSuccess &&
                per_sm >= 1) {
                double *ab_d = f->ab;
                int *ipiv_d = f->ipiv, *info_d = f->info;
                int nn = n, kll = kl, kuu = ku;
                void *args[] = {&ab_d, &ipiv_d, &nn, &kll, &kuu, &info_d};
                if (cudaLaunchCooperativeKernel((void *)nkb::banded_factor_coop_kernel, dim3(n_sm), dim3(256), args,
                                                smem, 0) == cudaSuccess) {
                    coop_done = true;
                    nkb::count_launch();
                } else {
                    cudaGetLastError();
                }
            }
        }
    }
    if (!coop_done) {
        nkb::banded_factor_kernel<<<f->nblk, 256>>>(f->ab, f->ipiv, n, kl, ku, f->blk, f->info);
        nkb::count_launch();
    }
    int info = 0;
    NKB_CUDA(cudaMemcpy(&info, f->info, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) {
        nkb::set_error("nkb_banded_create: matrix is singular at column " + std::to_string(info));
        return 3;
    }
    nkb::banded_transpose_kernel<<<n, 128>>>(f->ab, n, kl, ku, f->lt, f->ut);
    nkb::count_launch();
    NKB_CUDA(cudaDeviceSynchronize());
    const int K = std::max(std::max(kl, ku), 1);
    if (K <= 4) {
        // narrow band: compact rows for the Thomas kernel if the factorisation did not interchange rows
        std::vector<int> piv(n);
        NKB_CUDA(cudaMemcpy(piv.data(), f->ipiv, n * sizeof(int), cudaMemcpyDeviceToHost));
        bool plain = true;
        for (int j = 0; j < n && plain; ++j) plain = (piv[j] == j);
        if (plain) {
            NKB_CUDA(cudaMemcpy(host.data(), f->ab, host.size() * sizeof(double), cudaMemcpyDeviceToHost));
            const int NC = 2 * K + 1;
            std::vector<double> nbh((size_t)n * NC, 0.0);
            for (int j = 0; j < n; ++j) {
                for (int i = 1; i <= kl; ++i)
                    if (j - i >= 0) nbh[(size_t)j * NC + K - i] = host[(size_t)(kv + i) * n + (j - i)];
                nbh[(size_t)j * NC + K] = 1.0 / host[(size_t)kv * n + j];
                for (int i = 1; i <= ku; ++i)
                    if (j + i < n) nbh[(size_t)j * NC + K + i] = host[(size_t)(kv - i) * n + (j + i)];
            }
            if (cudaMalloc(&f->nb, nbh.size() * sizeof(double)) != cudaSuccess) {
                nkb::set_error("nkb_banded_create: cudaMalloc failed");
                        return 1;
            }
            NKB_CUDA(cudaMemcpy(f->nb, nbh.data(), nbh.size() * sizeof(double), cudaMemcpyHostToDevice));
            f->nb_k = K;
        }
    }
    if (f->nblk == 1 && K >= 16) {
        // one wide block: panel substitution if the factorisation did not interchange rows
        std::vector<int> piv(n);
        NKB_CUDA(cudaMemcpy(piv.data(), f->ipiv, n * sizeof(int), cudaMemcpyDeviceToHost));
        bool plain = true;
        for (int j = 0; j < n && plain; ++j) plain = (piv[j] == j);
        const char *env = getenv("NKB_BANDED_PANEL");
        if (plain && !(env && env[0] == '0')) {
            f->klp = (kl + 1) & ~1;
            f->kup = std::max(2, (ku + 1) & ~1);
            if (f->klp < 2) f->klp = 2;
            const size_t np = (size_t)(n + nkb::PR - 1) / nkb::PR;
            if (cudaMalloc(&f->lp, (size_t)n * f->klp * sizeof(double)) != cudaSuccess ||
                cudaMalloc(&f->up, (size_t)n * f->kup * sizeof(double)) != cudaSuccess ||
                cudaMalloc(&f->linv, np * nkb::PR * nkb::PR * sizeof(double)) != cudaSuccess ||
                cudaMalloc(&f->uinv, np * nkb::PR * nkb::PR * sizeof(double)) != cudaSuccess) {
                nkb::set_error("nkb_banded_create: cudaMalloc failed");
                        return 1;
            }
            nkb::banded_panel_pack_kernel<<<n, 128>>>(f->ab, n, kl, ku, f->klp, f->kup, f->lp, f->up);
            nkb::count_launch();
            nkb::banded_panel_inverse_kernel<<<(unsigned)np, 2 * nkb::PR>>>(f->ab, f->lp, f->up, n, kl, ku, f->klp,
                                                                          f->kup, f->linv, f->uinv);
            nkb::count_launch();
            NKB_CUDA(cudaDeviceSynchronize());
        }
    }
    guard.f = nullptr;
    *out = f;
    return 0;
}

int nkb_banded_blocks(const nkb_banded *f) { return f ? f->nblk : 0; }

int nkb_banded_path(const nkb_banded *f) {
    if (!f) return 0;
    return f->lp ? 3 : (f->nb_k > 0 ? 2 : 1);
}

void nkb_banded_destroy(nkb_banded *f) {
    if (!f) return;
    cudaFree(f->ab); cudaFree(f->ipiv); cudaFree(f->info);
    cudaFree(f->lt); cudaFree(f->ut); cudaFree(f->blk); cudaFree(f->nb);
    cudaFree(f->lp); cudaFree(f->up); cudaFree(f->linv); cudaFree(f->uinv);
    delete f;
}

int nkb_banded_solve(nkb_banded *f, const double *d_y, double *d_x, int B, int ldb, double scale, int subtract_rhs,
                     void *stream) {
    NKB_REQUIRE(f && d_y && d_x && B >= 1 && ldb >= B, "nkb_banded_solve: bad argument");
    NKB_REQUIRE(!(subtract_rhs && d_y == d_x), "nkb_banded_solve: subtract_rhs needs distinct x and y");
    {
        const char *env = getenv("NKB_BANDED_THOMAS");
        if (f->nb_k > 0 && B >= 16 && !(env && env[0] == '0')) {
            cudaStream_t st = (cudaStream_t)stream;
            switch (f->nb_k) {
                case 1: return nkb::launch_thomas<1>(f, d_y, d_x, B, (size_t)ldb, scale, subtract_rhs, st);
                case 2: return nkb::launch_thomas<2>(f, d_y, d_x, B, (size_t)ldb, scale, subtract_rhs, st);
                case 3: return nkb::launch_thomas<3>(f, d_y, d_x, B, (size_t)ldb, scale, subtract_rhs, st);
                default: return nkb::launch_thomas<4>(f, d_y, d_x, B, (size_t)ldb, scale, subtract_rhs, st);
            }
        }
    }
    // panel kernel (one wide block, no row interchanges): PR rows per barrier pair
    if (f->lp != nullptr) {
        const size_t budget = 220 * 1024;
        const int cw = std::max(f->klp, f->kup);
        int MB = 1;
        while (MB < B && MB < 8) MB <<= 1;
        auto need = [&](int mb, int ns) {
            const size_t wn = (((size_t)std::max(f->kl, f->ku) + 1) & ~(size_t)1) + (size_t)(ns + 1) * nkb::PR;
            return (wn * mb + (size_t)nkb::PR * mb + (size_t)ns * nkb::PR * cw + (size_t)ns * nkb::PR * nkb::PR +
                    (size_t)ns * nkb::PR * mb) * 8 + 64;
        };
        int NS = 3;
        if (need(MB, NS) > budget) NS = 2;
        while (MB > 1 && need(MB, NS) > budget) MB >>= 1;
        if (need(MB, NS) <= budget) {
            nkb::PanelArgs a;
            a.lp = f->lp; a.up = f->up; a.linv = f->linv; a.uinv = f->uinv;
            a.n = f->n; a.kl = f->kl; a.ku = f->ku; a.klp = f->klp; a.kup = f->kup;
            a.y = d_y; a.x = d_x; a.B = B; a.ldb = (size_t)ldb; a.scale = scale; a.subtract = subtract_rhs;
            a.MB = MB; a.NS = NS; a.cw = cw; a.Wn = ((std::max(f->kl, f->ku) + 1) & ~1) + (NS + 1) * nkb::PR;  // even: the ring behind it is 16-byte aligned
            NKB_CUDA(cudaFuncSetAttribute(nkb::banded_solve_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)budget));
            nkb::banded_solve_panel_kernel<<<(B + MB - 1) / MB, 256, need(MB, NS), (cudaStream_t)stream>>>(a);
            nkb::count_launch();
            NKB_CUDA(cudaGetLastError());
            return 0;
        }
    }
    // window kernel: pick the stage depth and the member lanes so that the shared memory fits
    {
        const int kv = f->kl + f->ku;
        const size_t budget = 200 * 1024;
        int CH = 16, MB = 1;
        while (MB < B && MB < 32) MB <<= 1;
        while (CH > 1 && (size_t)nkb::BS_NSTG * CH * (kv + 1) * 8 > budget / 2) CH >>= 1;
        auto need = [&](int mb) {
            const size_t wn = (size_t)nkb::BS_NSTG * CH + kv + 1;
            return (wn * mb + (size_t)nkb::BS_NSTG * CH * (kv + 1) + (size_t)nkb::BS_NSTG * CH * mb) * 8 +
                   (size_t)nkb::BS_NSTG * CH * 4 + 16;
        };
        while (MB > 1 && need(MB) > budget) MB >>= 1;
        const char *env = getenv("NKB_BANDED_WINDOW");
        if (need(MB) <= budget && !(env && env[0] == '0')) {
            int RW = 1;
            while (RW < kv && RW * MB < 256) RW <<= 1;
            nkb::BandedSolveArgs a;
            a.lt = f->lt; a.ut = f->ut; a.ipiv = f->ipiv; a.blk = f->blk; a.kl = f->kl; a.ku = f->ku;
            a.y = d_y; a.x = d_x; a.B = B; a.ldb = (size_t)ldb; a.scale = scale; a.subtract = subtract_rhs;
            a.MB = MB; a.RW = RW; a.CH = CH; a.Wn = nkb::BS_NSTG * CH + kv + 1;
            static unsigned long long attr_mask = 0;
            if (nkb::first_use_on_device(attr_mask)) {
                NKB_CUDA(cudaFuncSetAttribute(nkb::banded_solve_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)budget));
            }
            dim3 grid((B + MB - 1) / MB, f->nblk);
            nkb::banded_solve_win_kernel<<<grid, RW * MB, need(MB), (cudaStream_t)stream>>>(a);
            nkb::count_launch();
            NKB_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const int bs = B >= 128 ? 128 : 32;
    nkb::banded_solve_kernel<<<(B + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(
        f->ab, f->ipiv, f->n, f->kl, f->ku, d_y, d_x, B, (size_t)ldb, scale, subtract_rhs);
    nkb::count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
