// K5 / K6 — Krylov vector kernels on member-fastest batches.
//   pack/unpack : layout conversion member-major <-> member-fastest (tiled transpose)
//   wdot        : region-weighted dot products / means (TracerModuleStateBase.dot_prod/mean,
//                 tracer_module_state_base.py:371-388, with the CSR region-mean matrix of
//                 model_config.py:292-315); deterministic two-pass reduction, warp shuffles
//                 across the cell lanes of a warp
//   axpby       : y = alpha[r][b]*x + beta[r][b]*y with per-(region, member) scalars
//                 (tracer_module_state_base.py:255-369, broadcast_region_vals :502-515)
//   fd_sigma    : sigma = 1e-4*norm, 1 where 0 (model_state_base.py:509-511)
#include "nkb_common.cuh"

namespace nkb {

// batches wide enough for the two-members-per-lane kernels (16-byte accesses need an even member
// stride and 16-byte aligned bases)
static inline bool wide_batch(int B, int ldb, const void *p, const void *q) {
    return B >= 64 && (ldb % 2) == 0 && (((uintptr_t)p | (uintptr_t)q) & 15) == 0;
}

// the streaming kernels loop over cell groups: about 16 CTAs of 256 threads per SM in total, so that a
// CTA lives long enough to amortise its launch
static inline unsigned cap_cell_ctas(unsigned gx, unsigned gy) {
    const unsigned want = (148u * 16u + gx - 1) / gx;
    return gy < want ? gy : (want < 1 ? 1 : want);
}

// ---- transpose ------------------------------------------------------------------------
// src [rows][src_ld] -> dst [cols][dst_ld] for the rows x cols logical matrix
__global__ void transpose_kernel(const double *__restrict__ src, double *__restrict__ dst, int rows, int cols,
                                 size_t src_ld, size_t dst_ld) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[(size_t)r * src_ld + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * dst_ld + r] = tile[threadIdx.x][i];
    }
}

static int launch_transpose(const double *src, double *dst, int rows, int cols, size_t src_ld, size_t dst_ld,
                            cudaStream_t st) {
    dim3 block(32, 8), grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grid, block, 0, st>>>(src, dst, rows, cols, src_ld, dst_ld);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

// member-major [B][n] -> member-fastest [n][ldb]
int launch_pack(const double *src, double *dst, int n, int B, int ldb, cudaStream_t st) {
    return launch_transpose(src, dst, B, n, (size_t)n, (size_t)ldb, st);
}
// member-fastest [n][ldb] -> member-major [B][n]
int launch_unpack(const double *src, double *dst, int n, int B, int ldb, cudaStream_t st) {
    return launch_transpose(src, dst, n, B, (size_t)ldb, (size_t)n, st);
}

// ---- weighted dot -----------------------------------------------------------------------
// block (BX member lanes, BY cell lanes), BX*BY == 256, BX power of two <= 32.
// grid (member blocks, regions, chunks).  partial[chunk][r][b].
__global__ void __launch_bounds__(256)
wdot_partial_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                    const double *__restrict__ wdata, int T, size_t ncell, const double *__restrict__ a,
                    const double *__restrict__ bb, int B, size_t ldb, int R, double *__restrict__ partial) {
    __shared__ double red[256];
    const int bx = blockDim.x, by = blockDim.y;
    const int b = blockIdx.x * bx + threadIdx.x;
    const int r = blockIdx.y;
    const int lo = indptr[r], hi = indptr[r + 1];
    const int nchunk = gridDim.z;
    const int per = (hi - lo + nchunk - 1) / nchunk;
    const int c_lo = lo + blockIdx.z * per;
    const int c_hi = min(hi, c_lo + per);
    double acc = 0.0;
    if (b < B) {
        for (int i = c_lo + threadIdx.y; i < c_hi; i += by) {
            const size_t cell = indices[i];
            const double w = wdata[i];
            double s = 0.0;
            for (int t = 0; t < T; ++t) {
                const size_t off = ((size_t)t * ncell + cell) * ldb + b;
                const double av = a[off];
                s += bb ? av * bb[off] : av;
            }
            acc = fma(w, s, acc);
        }
    }
    // reduce over the cell lanes: first inside each warp (lanes that share a member are
    // bx apart), then across warps through shared memory
    for (int off = 16; off >= bx; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    const int tid = threadIdx.y * bx + threadIdx.x;
    red[tid] = acc;
    __syncthreads();
    if (tid < bx) {
        double s = 0.0;
        for (int wrp = 0; wrp < 256 / 32; ++wrp) s += red[wrp * 32 + tid];
        if (b < B) partial[((size_t)blockIdx.z * R + r) * B + b] = s;
    }
}

// Wide batches (B >= 64, even ldb): a lane owns TWO members (16-byte loads, 512 contiguous bytes per
// warp and cell row), the 8 warps of a CTA stride over the cells of the chunk, four cells in flight
// per thread.  Same two-pass, order-fixed reduction as above (deterministic).
template <bool HAS_B>
__global__ void __launch_bounds__(256)
wdot_partial_vec_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                        const double *__restrict__ wdata, int T, size_t ncell, const double *__restrict__ a,
                        const double *__restrict__ bb, int B, size_t ldb, int R, double *__restrict__ partial) {
    __shared__ double2 red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = 2 * (blockIdx.x * 32 + lane);
    const int r = blockIdx.y;
    const int lo = indptr[r], hi = indptr[r + 1];
    const int nchunk = gridDim.z;
    const int per = (hi - lo + nchunk - 1) / nchunk;
    const int c_lo = lo + blockIdx.z * per;
    const int c_hi = min(hi, c_lo + per);
    double2 acc = make_double2(0.0, 0.0);
    if (b < B) {
        constexpr int U = 4;
        int i = c_lo + warp;
        for (; i + 8 * (U - 1) < c_hi; i += 8 * U) {
            size_t cell[U];
            double w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cell[u] = (size_t)__ldg(indices + i + 8 * u);
                w[u] = __ldg(wdata + i + 8 * u);
            }
            for (int t = 0; t < T; ++t) {
                double2 av[U], bv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t off = ((size_t)t * ncell + cell[u]) * ldb + b;
                    av[u] = __ldcs(reinterpret_cast<const double2 *>(a + off));
                    if (HAS_B) bv[u] = __ldcs(reinterpret_cast<const double2 *>(bb + off));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (HAS_B) {
                        acc.x = fma(w[u], av[u].x * bv[u].x, acc.x);
                        acc.y = fma(w[u], av[u].y * bv[u].y, acc.y);
                    } else {
                        acc.x = fma(w[u], av[u].x, acc.x);
                        acc.y = fma(w[u], av[u].y, acc.y);
                    }
                }
            }
        }
        for (; i < c_hi; i += 8) {
            const size_t cell = (size_t)__ldg(indices + i);
            const double w = __ldg(wdata + i);
            for (int t = 0; t < T; ++t) {
                const size_t off = ((size_t)t * ncell + cell) * ldb + b;
                const double2 av = __ldcs(reinterpret_cast<const double2 *>(a + off));
                if (HAS_B) {
                    const double2 bv = __ldcs(reinterpret_cast<const double2 *>(bb + off));
                    acc.x = fma(w, av.x * bv.x, acc.x);
                    acc.y = fma(w, av.y * bv.y, acc.y);
                } else {
                    acc.x = fma(w, av.x, acc.x);
                    acc.y = fma(w, av.y, acc.y);
                }
            }
        }
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && b < B) {
        double2 s = red[0][lane];
#pragma unroll
        for (int wv = 1; wv < 8; ++wv) {
            s.x += red[wv][lane].x;
            s.y += red[wv][lane].y;
        }
        double *dst = partial + ((size_t)blockIdx.z * R + r) * B + b;
        dst[0] = s.x;
        if (b + 1 < B) dst[1] = s.y;
    }
}

__global__ void wdot_final_kernel(const double *__restrict__ partial, int nchunk, size_t n, double *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int c = 0; c < nchunk; ++c) s += partial[(size_t)c * n + i];
    out[i] = s;
}

// ---- axpby ---------------------------------------------------------------------------------
__global__ void axpby_kernel(const int *__restrict__ region, int R, int T, size_t ncell,
                             const double *__restrict__ alpha, const double *__restrict__ x,
                             const double *__restrict__ beta, double *__restrict__ y, double fill_alpha,
                             double fill_beta, int B, size_t ldb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t cell = (size_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (b >= B || cell >= ncell) return;
    const int r = region ? region[cell] : 1;
    double al = fill_alpha, be = fill_beta;
    if (r > 0) {
        if (alpha) al = alpha[(size_t)(r - 1) * B + b];
        if (beta) be = beta[(size_t)(r - 1) * B + b];
    }
    for (int t = 0; t < T; ++t) {
        const size_t off = ((size_t)t * ncell + cell) * ldb + b;
        double v = 0.0;
        if (x) v = al * x[off];
        if (be != 0.0) v = fma(be, y[off], v);
        y[off] = v;
    }
}

// Wide batches: a lane owns two members, a warp four consecutive cells (all tracers of a cell by the
// same thread so that the region scalars are fetched once)
template <bool HAS_X>
__global__ void __launch_bounds__(256)
axpby_vec_kernel(const int *__restrict__ region, int T, size_t ncell, const double *__restrict__ alpha,
                 const double *__restrict__ x, const double *__restrict__ beta, double *__restrict__ y,
                 double fill_alpha, double fill_beta, int B, size_t ldb) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = 2 * (blockIdx.x * 32 + lane);
    if (b >= B) return;
    const bool two = (b + 1 < B);
    for (size_t cell0 = ((size_t)blockIdx.y * 8 + warp) * U; cell0 < ncell; cell0 += (size_t)gridDim.y * 8 * U) {
    double2 al[U], be[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        al[u] = make_double2(fill_alpha, fill_alpha);
        be[u] = make_double2(fill_beta, fill_beta);
        const size_t cell = cell0 + u;
        const int r = (cell < ncell) ? (region ? __ldg(region + cell) : 1) : 0;
        if (r > 0) {
            if (alpha) {
                al[u].x = __ldg(alpha + (size_t)(r - 1) * B + b);
                if (two) al[u].y = __ldg(alpha + (size_t)(r - 1) * B + b + 1);
            }
            if (beta) {
                be[u].x = __ldg(beta + (size_t)(r - 1) * B + b);
                if (two) be[u].y = __ldg(beta + (size_t)(r - 1) * B + b + 1);
            }
        }
    }
    for (int t = 0; t < T; ++t) {
        double2 xv[U], yv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t off = ((size_t)t * ncell + cell0 + u) * ldb + b;
            xv[u] = yv[u] = make_double2(0.0, 0.0);
            if (cell0 + u < ncell) {
                if (HAS_X) xv[u] = __ldcs(reinterpret_cast<const double2 *>(x + off));
                yv[u] = __ldcs(reinterpret_cast<const double2 *>(y + off));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (cell0 + u >= ncell) continue;
            const size_t off = ((size_t)t * ncell + cell0 + u) * ldb + b;
            double2 v = make_double2(0.0, 0.0);
            if (HAS_X) v = make_double2(al[u].x * xv[u].x, al[u].y * xv[u].y);
            // beta == 0 discards y (also its NaN / Inf), as in the scalar kernel
            if (be[u].x != 0.0) v.x = fma(be[u].x, yv[u].x, v.x);
            if (be[u].y != 0.0) v.y = fma(be[u].y, yv[u].y, v.y);
            __stcs(reinterpret_cast<double2 *>(y + off), v);
        }
    }
    }
}

__global__ void fd_sigma_kernel(const double *__restrict__ nrm, double *__restrict__ sigma, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = 1.0e-4 * nrm[i];
    sigma[i] = (s == 0.0) ? 1.0 : s;
}

// ---- limiter -------------------------------------------------------------------------------
// out[r][b] = min over tracers and cells of region r of the largest scale factor in [0, 1] that
// keeps base + scalef*inc inside [lob, upb] (utils.py:561-600 comp_scalef_lob/upb +
// min_by_region :544-558).  Scale factors are non-negative doubles, whose bit patterns order like
// unsigned integers: atomicMin on the bits is exact and order independent (deterministic).
// flag[b] collects, per member: 1 base < lob somewhere, 2 base + inc < lob somewhere, 4 base > upb, 8 base +
// inc > upb.  The reference raises ValueError only when a bound needs enforcing (2 / 8) AND base itself violates
// it (1 / 4); an iterate a hair outside a bound with a harmless increment passes with scale factor 1.
__global__ void limiter_scalef_kernel(const int *__restrict__ region, int T, size_t ncell,
                                      const double *__restrict__ base, const double *__restrict__ inc, double lob,
                                      int has_lob, double upb, int has_upb, int B, size_t ldb,
                                      unsigned long long *__restrict__ out_bits, int *__restrict__ flag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t cell = (size_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (b >= B || cell >= ncell) return;
    const int r = region[cell];
    if (r <= 0) return;
    double sc = 1.0;
    int bits = 0;
    for (int t = 0; t < T; ++t) {
        const size_t off = ((size_t)t * ncell + cell) * ldb + b;
        const double x = base[off], d = inc[off];
        if (has_lob) {
            if (x < lob) bits |= 1;
            if (x + d < lob) {
                bits |= 2;
                sc = fmin(sc, fabs((lob - x) / d));
            }
        }
        if (has_upb) {
            if (x > upb) bits |= 4;
            if (x + d > upb) {
                bits |= 8;
                sc = fmin(sc, fabs((upb - x) / d));
            }
        }
    }
    if (bits) atomicOr(flag + b, bits);
    atomicMin(out_bits + (size_t)(r - 1) * B + b, (unsigned long long)__double_as_longlong(sc));
}

// Wide batches: a lane owns two members, a warp LC consecutive cells; the running minimum of a
// thread is flushed (one atomic per member) only when the region changes: 1/LC of the atomics.
__global__ void __launch_bounds__(256)
limiter_scalef_vec_kernel(const int *__restrict__ region, int T, size_t ncell, const double *__restrict__ base,
                          const double *__restrict__ inc, double lob, int has_lob, double upb, int has_upb, int B,
                          size_t ldb, unsigned long long *__restrict__ out_bits, int *__restrict__ flag) {
    constexpr int LC = 16, U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = 2 * (blockIdx.x * 32 + lane);
    if (b >= B) return;
    const bool two = (b + 1 < B);
    auto flush = [&](int r, double2 sc) {
        if (r <= 0) return;
        atomicMin(out_bits + (size_t)(r - 1) * B + b, (unsigned long long)__double_as_longlong(sc.x));
        if (two) atomicMin(out_bits + (size_t)(r - 1) * B + b + 1, (unsigned long long)__double_as_longlong(sc.y));
    };
    int bits0 = 0, bits1 = 0;
    auto one = [&](double x, double d, double &sc, bool live, int &bits) {
        if (has_lob) {
            if (live && x < lob) bits |= 1;
            if (x + d < lob) {
                if (live) bits |= 2;
                sc = fmin(sc, fabs((lob - x) / d));
            }
        }
        if (has_upb) {
            if (live && x > upb) bits |= 4;
            if (x + d > upb) {
                if (live) bits |= 8;
                sc = fmin(sc, fabs((upb - x) / d));
            }
        }
    };
    int rcur = 0;
    double2 sc = make_double2(1.0, 1.0);
    for (size_t cell0 = ((size_t)blockIdx.y * 8 + warp) * LC; cell0 < ncell; cell0 += (size_t)gridDim.y * 8 * LC)
    for (int c = 0; c < LC; c += U) {
        int r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = (cell0 + c + u < ncell) ? __ldg(region + cell0 + c + u) : 0;
        for (int t = 0; t < T; ++t) {
            double2 xv[U], dv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                xv[u] = dv[u] = make_double2(0.0, 0.0);
                if (r[u] > 0) {
                    const size_t off = ((size_t)t * ncell + cell0 + c + u) * ldb + b;
                    xv[u] = __ldcs(reinterpret_cast<const double2 *>(base + off));
                    dv[u] = __ldcs(reinterpret_cast<const double2 *>(inc + off));
                }
            }
            // (with several tracers the cells of this group are revisited per tracer)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (r[u] <= 0) continue;
                double2 s1 = make_double2(1.0, 1.0);
                one(xv[u].x, dv[u].x, s1.x, true, bits0);
                one(xv[u].y, dv[u].y, s1.y, two, bits1);
                if (r[u] != rcur) {
                    flush(rcur, sc);
                    rcur = r[u];
                    sc = s1;
                } else {
                    sc.x = fmin(sc.x, s1.x);
                    sc.y = fmin(sc.y, s1.y);
                }
            }
        }
    }
    flush(rcur, sc);
    if (bits0) atomicOr(flag + b, bits0);
    if (bits1) atomicOr(flag + b + 1, bits1);
}

}  // namespace nkb

extern "C" {

int nkb_limiter_scalef(const int32_t *d_region, int R, int T, int ncell, const double *d_base, const double *d_inc,
                       double lob, int has_lob, double upb, int has_upb, int B, int ldb, double *d_out,
                       int32_t *d_flag, void *stream) {
    NKB_REQUIRE(d_region && d_base && d_inc && d_out && d_flag, "nkb_limiter_scalef: null argument");
    NKB_REQUIRE(R >= 1 && T >= 1 && ncell >= 1 && B >= 1 && ldb >= B, "nkb_limiter_scalef: bad size");
    if (nkb::wide_batch(B, ldb, d_base, d_inc)) {
        dim3 grid((B + 63) / 64, (ncell + 8 * 16 - 1) / (8 * 16));
        grid.y = nkb::cap_cell_ctas(grid.x, grid.y);
        nkb::limiter_scalef_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
            d_region, T, (size_t)ncell, d_base, d_inc, lob, has_lob, upb, has_upb, B, (size_t)ldb,
            reinterpret_cast<unsigned long long *>(d_out), d_flag);
        nkb::count_launch();
        NKB_CUDA(cudaGetLastError());
        return 0;
    }
    int bx = 1;
    while (bx < B && bx < 32) bx <<= 1;
    dim3 block(bx, 256 / bx), grid((B + bx - 1) / bx, (ncell + block.y - 1) / block.y);
    nkb::limiter_scalef_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        d_region, T, (size_t)ncell, d_base, d_inc, lob, has_lob, upb, has_upb, B, (size_t)ldb,
        reinterpret_cast<unsigned long long *>(d_out), d_flag);
    nkb::count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int nkb_pack_members(const double *d_src_major, double *d_dst_fast, int n, int B, int ldb, void *stream) {
    NKB_REQUIRE(d_src_major && d_dst_fast && n > 0 && B > 0 && ldb >= B, "nkb_pack_members: bad argument");
    return nkb::launch_pack(d_src_major, d_dst_fast, n, B, ldb, (cudaStream_t)stream);
}

int nkb_unpack_members(const double *d_src_fast, double *d_dst_major, int n, int B, int ldb, void *stream) {
    NKB_REQUIRE(d_src_fast && d_dst_major && n > 0 && B > 0 && ldb >= B, "nkb_unpack_members: bad argument");
    return nkb::launch_unpack(d_src_fast, d_dst_major, n, B, ldb, (cudaStream_t)stream);
}

int nkb_wdot_chunks(int ncell_max) {
    int c = (ncell_max + 1023) / 1024;
    return c < 1 ? 1 : (c > 64 ? 64 : c);
}

int nkb_wdot(const int32_t *d_indptr, const int32_t *d_indices, const double *d_wdata, int R, int T, int ncell,
             const double *d_a, const double *d_b, int B, int ldb, double *d_partial, int n_chunks,
             double *d_out, void *stream) {
    NKB_REQUIRE(d_indptr && d_indices && d_wdata && d_a && d_out && d_partial, "nkb_wdot: null argument");
    NKB_REQUIRE(R >= 1 && T >= 1 && ncell >= 1 && B >= 1 && ldb >= B && n_chunks >= 1, "nkb_wdot: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    if (nkb::wide_batch(B, ldb, d_a, d_b ? d_b : d_a)) {
        dim3 grid((B + 63) / 64, R, n_chunks);
        if (d_b)
            nkb::wdot_partial_vec_kernel<true><<<grid, 256, 0, st>>>(d_indptr, d_indices, d_wdata, T, (size_t)ncell, d_a,
                                                                     d_b, B, (size_t)ldb, R, d_partial);
        else
            nkb::wdot_partial_vec_kernel<false><<<grid, 256, 0, st>>>(d_indptr, d_indices, d_wdata, T, (size_t)ncell,
                                                                      d_a, d_b, B, (size_t)ldb, R, d_partial);
    } else {
        int bx = 1;
        while (bx < B && bx < 32) bx <<= 1;
        dim3 block(bx, 256 / bx), grid((B + bx - 1) / bx, R, n_chunks);
        nkb::wdot_partial_kernel<<<grid, block, 0, st>>>(d_indptr, d_indices, d_wdata, T, (size_t)ncell, d_a, d_b, B,
                                                         (size_t)ldb, R, d_partial);
    }
    nkb::count_launch();
    const size_t n = (size_t)R * B;
    nkb::wdot_final_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_partial, n_chunks, n, d_out);
    nkb::count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int nkb_axpby(const int32_t *d_region, int R, int T, int ncell, const double *d_alpha, const double *d_x,
              const double *d_beta, double *d_y, double fill_alpha, double fill_beta, int B, int ldb,
              void *stream) {
    NKB_REQUIRE(d_y && T >= 1 && ncell >= 1 && B >= 1 && ldb >= B, "nkb_axpby: bad argument");
    if (nkb::wide_batch(B, ldb, d_y, d_x ? d_x : d_y)) {
        dim3 grid((B + 63) / 64, (ncell + 8 * 4 - 1) / (8 * 4));
        grid.y = nkb::cap_cell_ctas(grid.x, grid.y);
        if (d_x)
            nkb::axpby_vec_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
                d_region, T, (size_t)ncell, d_alpha, d_x, d_beta, d_y, fill_alpha, fill_beta, B, (size_t)ldb);
        else
            nkb::axpby_vec_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
                d_region, T, (size_t)ncell, d_alpha, d_x, d_beta, d_y, fill_alpha, fill_beta, B, (size_t)ldb);
        nkb::count_launch();
        NKB_CUDA(cudaGetLastError());
        return 0;
    }
    int bx = 1;
    while (bx < B && bx < 32) bx <<= 1;
    dim3 block(bx, 256 / bx), grid((B + bx - 1) / bx, (ncell + block.y - 1) / block.y);
    nkb::axpby_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_region, R, T, (size_t)ncell, d_alpha, d_x, d_beta,
                                                               d_y, fill_alpha, fill_beta, B, (size_t)ldb);
    nkb::count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int nkb_fd_sigma(const double *d_norm, double *d_sigma, int n, void *stream) {
    NKB_REQUIRE(d_norm && d_sigma && n >= 1, "nkb_fd_sigma: bad argument");
    nkb::fd_sigma_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_norm, d_sigma, n);
    nkb::count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
