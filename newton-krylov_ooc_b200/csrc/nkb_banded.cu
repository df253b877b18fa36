// K4 — member-shared banded LU with partial pivoting + batched member-fastest solves.
//
// The preconditioner matrix is member-independent, so it is factored ONCE on the device
// (one CTA, LAPACK dgbtf2-style right-looking elimination inside the band) and the factor is
// then applied to every member's right-hand side, one thread per member, coalesced across
// members.  Replaces scipy.linalg.solve_banded((1,1), ...) (test_problem/iage.py:50,
// dye_decay.py:71) and scipy.sparse.linalg.spsolve (py_driver_2d/iage.py:91, forced.py:239).
//
// Band storage (row-major): A(i,j) lives at ab[(kv + i - j)*n + j], kv = kl + ku, rows
// 0..kl-1 are fill-in space, total 2*kl + ku + 1 rows.
#include <vector>

#include "nkb_common.cuh"

struct nkb_banded {
    int n = 0, kl = 0, ku = 0;
    double *ab = nullptr;  // [(2kl+ku+1)][n]
    int *ipiv = nullptr;   // [n]
    int *info = nullptr;
};

namespace nkb {

__global__ void __launch_bounds__(256) banded_factor_kernel(double *__restrict__ ab, int *__restrict__ ipiv,
                                                            int n, int kl, int ku, int *info) {
    __shared__ double s_val[256];
    __shared__ int s_idx[256];
    __shared__ int s_ju;
    const int kv = kl + ku;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { s_ju = 0; *info = 0; }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        const int km = min(kl, n - 1 - j);
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
            if (pv == 0.0 && *info == 0) *info = j + 1;
            if (pv != 0.0) s_ju = max(s_ju, min(j + ku + jp, n - 1));
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

// one thread per member; x [n][ldb] in place
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
    f->n = n; f->kl = kl; f->ku = ku;
    const int rows = 2 * kl + ku + 1;
    std::vector<double> host((size_t)rows * n, 0.0);
    // scipy solve_banded layout in: ab_in[ku + i - j][j]  ->  rows kl.. of the working band
    for (int r = 0; r < kl + ku + 1; ++r)
        for (int j = 0; j < n; ++j) host[(size_t)(kl + r) * n + j] = h_ab[(size_t)r * n + j];
    if (cudaMalloc(&f->ab, host.size() * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&f->ipiv, n * sizeof(int)) != cudaSuccess || cudaMalloc(&f->info, sizeof(int)) != cudaSuccess) {
        nkb::set_error("nkb_banded_create: cudaMalloc failed (is a CUDA device present?)");
        delete f;
        return 1;
    }
    NKB_CUDA(cudaMemcpy(f->ab, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
    nkb::banded_factor_kernel<<<1, 256>>>(f->ab, f->ipiv, n, kl, ku, f->info);
    nkb::count_launch();
    int info = 0;
    NKB_CUDA(cudaMemcpy(&info, f->info, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) {
        nkb::set_error("nkb_banded_create: matrix is singular at column " + std::to_string(info));
        nkb_banded_destroy(f);
        return 3;
    }
    *out = f;
    return 0;
}

void nkb_banded_destroy(nkb_banded *f) {
    if (!f) return;
    cudaFree(f->ab); cudaFree(f->ipiv); cudaFree(f->info);
    delete f;
}

int nkb_banded_solve(nkb_banded *f, const double *d_y, double *d_x, int B, int ldb, double scale, int subtract_rhs,
                     void *stream) {
    NKB_REQUIRE(f && d_y && d_x && B >= 1 && ldb >= B, "nkb_banded_solve: bad argument");
    NKB_REQUIRE(!(subtract_rhs && d_y == d_x), "nkb_banded_solve: subtract_rhs needs distinct x and y");
    const int bs = B >= 128 ? 128 : 32;
    nkb::banded_solve_kernel<<<(B + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(
        f->ab, f->ipiv, f->n, f->kl, f->ku, d_y, d_x, B, (size_t)ldb, scale, subtract_rhs);
    nkb::count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
