// minimal probe: 3-D vs 4-D tiled TMA load of a [KC][16] box of doubles from a plane table
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, double *out) {
    __shared__ alignas(1024) double buf[64];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(s32(&bar)) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(buf)), "l"((uint64_t)&map), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(buf)), "l"((uint64_t)&map), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(0) : "memory");
    }
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(&bar)) : "memory");
    } while (!ok);
    if (threadIdx.x < 64) out[threadIdx.x] = buf[threadIdx.x];
}
int main(int argc, char **argv) {
    const int rank = argc > 1 ? atoi(argv[1]) : 3;
    const int ny = argc > 2 ? atoi(argv[2]) : 33, nz = 21, npl = 6;
    const int nyp = (ny + 1) & ~1; const int c0 = argc > 3 ? atoi(argv[3]) : -1;
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fp, 12000, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    std::vector<double> h((size_t)npl * nz * nyp);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *out;
    cudaMalloc(&d, h.size() * 8);
    cudaMalloc(&out, 64 * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)ny, (cuuint64_t)nz, (cuuint64_t)npl, 1};
    cuuint64_t strides[3] = {(cuuint64_t)nyp * 8, (cuuint64_t)nz * nyp * 8, (cuuint64_t)npl * nz * nyp * 8};
    cuuint32_t box[4] = {16, 4, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d ny %d encode rc=%d\n", rank, ny, (int)rc);
    if (rank == 3) probe<3><<<1, 64>>>(map, c0, 4, 2, out); else probe<4><<<1, 64>>>(map, c0, 4, 2, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        double ho[64];
        cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
        printf("c0=%d ", c0); printf("buf[0..3] = %g %g %g %g ; expect 0 %g %g; row1[1] %g expect %g\n", ho[0], ho[1], ho[2], ho[3],
               h[((size_t)2 * nz + 4) * nyp + 0], h[((size_t)2 * nz + 4) * nyp + 1], ho[17], h[((size_t)2 * nz + 5) * nyp + 0]);
    }
    return 0;
}
