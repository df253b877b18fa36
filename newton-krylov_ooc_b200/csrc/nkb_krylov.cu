// K5b — fused Krylov building blocks on member-fastest batches.
//   mgs          : modified Gram-Schmidt of w against k basis vectors in ONE cooperative launch
//                  (ModelStateBase.mod_gram_schmidt, nk_ooc/model_state_base.py:365-377: for i < k:
//                  h_i = dot(w, v_i); w -= h_i v_i, sequentially).  w lives in registers for the
//                  whole launch, every basis vector is read exactly once: 8 N (k + 2) bytes
//                   instead of the 5 k N 8 bytes of k x (dot, axpy); the k
//                  [region, member] scalars stay on the device.  Deterministic: per-warp partial sums
//                  combined in a fixed order, no floating-point atomics.
//   lin_comb     : out = sum_i coeff_i[r][b] v_i (+ add) in one pass (model_state_base.lin_comb,
//                  nk_ooc/model_state_base.py:619-624; krylov_solver.py:141-154): 8 N (k + 1) bytes
//   interleave   : [G][n][W] blocks of an all-gather -> member-fastest [n][G*W] (the gather of result
//                  columns to the owner of the Krylov basis)
#include <cooperative_groups.h>

#include "nkb_common.cuh"

namespace cg = cooperative_groups;

namespace nkb {

constexpr int MGS_MAXK = 64;      // basis pointers travel by value in the kernel parameters
constexpr int MGS_THREADS = 256;  // 8 warps: 8 x R*B doubles of shared memory
constexpr int MGS_MAXKEYS = 1024; // R*B of the resident kernel (8 x 1024 x 8 B = 64 KB)

struct BasisPtrs {
    const double *p[MGS_MAXK];
};

// element e of the [T][ncell][B] index space: row = e / B (tracer, cell), member b = e % B
template <int S>
__global__ void __launch_bounds__(MGS_THREADS)
mgs_resident_kernel(const int *__restrict__ region, const double *__restrict__ cellw, int R, size_t ncell,
                    size_t n_elems, int B, size_t ldb, double *__restrict__ w, const BasisPtrs basis, int k,
                    double *__restrict__ partial, double *__restrict__ h) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double red[];  // [warps][R*B]
    const int RB = R * B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const size_t nthr = (size_t)gridDim.x * blockDim.x, gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double wv[S], vv[S], cw[S];
    size_t off[S];
    int key[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const size_t e = gtid + (size_t)s * nthr;
        key[s] = -1;
        wv[s] = vv[s] = cw[s] = 0.0;
        off[s] = 0;
        if (e < n_elems) {
            const size_t row = e / B;
            const int b = (int)(e - row * B);
            const size_t cell = row % ncell;
            const int r = region ? region[cell] : 1;
            off[s] = row * ldb + b;
            wv[s] = w[off[s]];
            if (r > 0) {
                key[s] = (r - 1) * B + b;
                cw[s] = cellw[cell];
            }
        }
    }
    double *mine = red + (size_t)warp * RB;
    for (int i = 0; i < k; ++i) {
        const double *__restrict__ v = basis.p[i];
        for (int t = lane; t < RB; t += 32) mine[t] = 0.0;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const bool valid = (gtid + (size_t)s * nthr) < n_elems;
            vv[s] = valid ? v[off[s]] : 0.0;
            const double pr = cw[s] * wv[s] * vv[s];
            // lanes with the same (region, member) key are summed in lane order; the lowest such lane adds the
            // sum to the warp's own accumulator (distinct keys -> distinct addresses: no race, fixed order)
            double sum = 0.0;
            int leader = 32;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const double pl = __shfl_sync(0xffffffffu, pr, l);
                const int kl = __shfl_sync(0xffffffffu, key[s], l);
                if (kl == key[s]) {
                    sum += pl;
                    if (l < leader) leader = l;
                }
            }
            if (key[s] >= 0 && leader == lane) mine[key[s]] += sum;
            __syncwarp();
        }
        __syncthreads();
        for (int t = threadIdx.x; t < RB; t += blockDim.x) {
            double tot = 0.0;
            for (int q = 0; q < nwarp; ++q) tot += red[(size_t)q * RB + t];
            partial[(size_t)blockIdx.x * RB + t] = tot;
        }
        grid.sync();
        // one thread per key sums the partial results of all CTAs in a fixed order
        for (size_t t = gtid; t < (size_t)RB; t += nthr) {
            double tot = 0.0;
            for (unsigned c = 0; c < gridDim.x; ++c) tot += partial[(size_t)c * RB + t];
            h[(size_t)i * RB + t] = tot;
        }
        grid.sync();
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (key[s] >= 0) wv[s] = fma(-__ldcg(h + (size_t)i * RB + key[s]), vv[s], wv[s]);
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
        if ((gtid + (size_t)s * nthr) < n_elems) w[off[s]] = wv[s];
}

// y -= h[r][b] * x  (general path of mgs: after nkb_wdot)
__global__ void sub_scaled_kernel(const int *__restrict__ region, size_t ncell, size_t n_elems, int B, size_t ldb,
                                  const double *__restrict__ hval, const double *__restrict__ x,
                                  double *__restrict__ y) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += (size_t)gridDim.x * blockDim.x) {
        const size_t row = e / B;
        const int b = (int)(e - row * B);
        const int r = region ? region[row % ncell] : 1;
        if (r > 0) {
            const size_t o = row * ldb + b;
            y[o] = fma(-hval[(size_t)(r - 1) * B + b], x[o], y[o]);
        }
    }
}

__global__ void __launch_bounds__(256)
lin_comb_kernel(const int *__restrict__ region, size_t ncell, size_t n_elems, int B, size_t ldb,
                const double *__restrict__ coeff, int RB, const BasisPtrs basis, int k, const double *add,
                double *out, double fill) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += (size_t)gridDim.x * blockDim.x) {
        const size_t row = e / B;
        const int b = (int)(e - row * B);
        const int r = region ? region[row % ncell] : 1;
        const size_t o = row * ldb + b;
        double acc = add ? add[o] : 0.0;
        if (r > 0) {
            const double *c = coeff + (size_t)(r - 1) * B + b;
            for (int i = 0; i < k; ++i) acc = fma(__ldg(c + (size_t)i * RB), __ldcs(basis.p[i] + o), acc);
        } else {
            for (int i = 0; i < k; ++i) acc = fma(fill, __ldcs(basis.p[i] + o), acc);
        }
        out[o] = acc;
    }
}

// gathered [G][n][W] -> out [n][ldo], member g*W + m of row i from block g
__global__ void __launch_bounds__(256)
interleave_blocks_kernel(const double *__restrict__ src, double *__restrict__ dst, size_t n, int G, int W, size_t ldo,
                         int B) {
    const size_t total = n * (size_t)G * W;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(e % W);
        const size_t t = e / W;
        const int g = (int)(t % G);
        const size_t i = t / G;
        const int member = g * W + m;
        if (member < B) dst[i * ldo + member] = __ldcs(src + ((size_t)g * n + i) * W + m);
    }
}

static int n_sm() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int S>
static int mgs_try_resident(const int *region, const double *cellw, int R, size_t ncell, size_t n_elems, int B, size_t ldb,
                            double *w, const BasisPtrs &bp, int k, double *partial, size_t partial_cap, double *h,
                            cudaStream_t st) {
    const size_t smem = (size_t)(MGS_THREADS / 32) * R * B * sizeof(double);
    auto kern = mgs_resident_kernel<S>;
    if (smem > 48 * 1024) NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MGS_THREADS, smem));
    if (per_sm < 1) return -1;
    if (per_sm > 2) per_sm = 2;
    size_t want = (n_elems + (size_t)MGS_THREADS * S - 1) / ((size_t)MGS_THREADS * S);
    const size_t cap = (size_t)n_sm() * per_sm;
    if (want > cap) return -1;  // does not fit with S elements per thread
    if (want < 1) want = 1;
    if (want * (size_t)R * B > partial_cap) return -1;
    int Ri = R, Bi = B, ki = k;
    void *args[] = {(void *)&region, (void *)&cellw, (void *)&Ri, (void *)&ncell, (void *)&n_elems, (void *)&Bi,
                    (void *)&ldb, (void *)&w, (void *)&bp, (void *)&ki, (void *)&partial, (void *)&h};
    const cudaError_t err = cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)want), dim3(MGS_THREADS), args,
                                                        smem, st);
    if (err == cudaErrorCooperativeLaunchTooLarge) {
        cudaGetLastError();
        return -1;
    }
    NKB_CUDA(err);
    count_launch();
    return 0;
}

}  // namespace nkb

extern "C" {

size_t nkb_mgs_scratch_doubles(int R, int B, int ncell_max) {
    // resident kernel: one [R*B] partial per CTA (<= 2 per SM); general path: the two-pass scratch of nkb_wdot
    const size_t a = (size_t)2 * nkb::n_sm() * R * B;
    const size_t b = (size_t)nkb_wdot_chunks(ncell_max) * R * B;
    return a > b ? a : b;
}

int nkb_mgs(const int32_t *d_indptr, const int32_t *d_indices, const double *d_wdata, const int32_t *d_region,
            const double *d_cellw, int R, int T, int ncell, int ncell_max_row, double *d_w,
            const double *const *h_basis, int k, int B, int ldb, double *d_scratch, size_t scratch_doubles,
            double *d_h, void *stream) {
    NKB_REQUIRE(d_indptr && d_indices && d_wdata && d_cellw && d_w && d_h && d_scratch, "nkb_mgs: null argument");
    NKB_REQUIRE(k >= 0 && R >= 1 && T >= 1 && ncell >= 1 && B >= 1 && ldb >= B, "nkb_mgs: bad argument");
    if (k == 0) return 0;
    NKB_REQUIRE(h_basis, "nkb_mgs: null basis");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_elems = (size_t)T * ncell * B;
    const char *env = getenv("NKB_MGS_RESIDENT");
    const bool allow = !(env && env[0] == '0');
    if (allow && k <= nkb::MGS_MAXK && R * B <= nkb::MGS_MAXKEYS) {
        nkb::BasisPtrs bp;
        for (int i = 0; i < k; ++i) bp.p[i] = h_basis[i];
        int rc = -1;
        const size_t nc = (size_t)ncell, ld = (size_t)ldb;
        if (rc == -1) rc = nkb::mgs_try_resident<1>(d_region, d_cellw, R, nc, n_elems, B, ld, d_w, bp, k, d_scratch, scratch_doubles, d_h, st);
        if (rc == -1) rc = nkb::mgs_try_resident<4>(d_region, d_cellw, R, nc, n_elems, B, ld, d_w, bp, k, d_scratch, scratch_doubles, d_h, st);
        if (rc == -1) rc = nkb::mgs_try_resident<16>(d_region, d_cellw, R, nc, n_elems, B, ld, d_w, bp, k, d_scratch, scratch_doubles, d_h, st);
        if (rc != -1) return rc;
    }
    // general path (w does not fit on the chip): per basis vector one dot and one update, scalars on the device
    const int nch = nkb_wdot_chunks(ncell_max_row);
    NKB_REQUIRE((size_t)nch * R * B <= scratch_doubles, "nkb_mgs: scratch too small");
    for (int i = 0; i < k; ++i) {
        double *hi = d_h + (size_t)i * R * B;
        if (nkb_wdot(d_indptr, d_indices, d_wdata, R, T, ncell, d_w, h_basis[i], B, ldb, d_scratch, nch, hi, stream)) return 1;
        const unsigned blocks = (unsigned)((n_elems + 255) / 256 < 148u * 16u ? (n_elems + 255) / 256 : 148u * 16u);
        nkb::sub_scaled_kernel<<<blocks, 256, 0, st>>>(d_region, (size_t)ncell, n_elems, B, (size_t)ldb, hi, h_basis[i], d_w);
        nkb::count_launch();
    }
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int nkb_lin_comb(const int32_t *d_region, int R, int T, int ncell, const double *d_coeff, const double *const *h_basis,
                 int k, const double *d_add, double *d_out, double fill, int B, int ldb, void *stream) {
    NKB_REQUIRE(d_coeff && h_basis && d_out && k >= 1 && R >= 1 && T >= 1 && ncell >= 1 && B >= 1 && ldb >= B,
                "nkb_lin_comb: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_elems = (size_t)T * ncell * B;
    const unsigned blocks = (unsigned)((n_elems + 255) / 256 < 148u * 16u ? (n_elems + 255) / 256 : 148u * 16u);
    const double *add = d_add;
    for (int i0 = 0; i0 < k; i0 += nkb::MGS_MAXK) {  // more than 64 vectors: accumulate in passes
        nkb::BasisPtrs bp;
        const int kk = (k - i0 < nkb::MGS_MAXK) ? k - i0 : nkb::MGS_MAXK;
        for (int i = 0; i < kk; ++i) bp.p[i] = h_basis[i0 + i];
        nkb::lin_comb_kernel<<<blocks, 256, 0, st>>>(d_region, (size_t)ncell, n_elems, B, (size_t)ldb,
                                                     d_coeff + (size_t)i0 * R * B, R * B, bp, kk, add, d_out, fill);
        nkb::count_launch();
        add = d_out;
    }
    NKB_CUDA(cudaGetLastError());
    return 0;
}

int nkb_interleave_blocks(const double *d_gathered, double *d_out, size_t n, int G, int W, int ldo, int B, void *stream) {
    NKB_REQUIRE(d_gathered && d_out && n >= 1 && G >= 1 && W >= 1 && B >= 1 && ldo >= B, "nkb_interleave_blocks: bad argument");
    const size_t total = n * (size_t)G * W;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148u * 32u ? (total + 255) / 256 : 148u * 32u);
    nkb::interleave_blocks_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_gathered, d_out, n, G, W, (size_t)ldo, B);
    nkb::count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
