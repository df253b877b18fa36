// K1 + K2, TMA variant — the fused IMEX stage kernel with the stage inputs streamed into shared
// memory by the Tensor Memory Accelerator (cp.async.bulk.tensor, SASS UTMALDG) through an
// mbarrier full/empty ring, warp-specialised: one producer warp issues box copies of
// [KC levels] x [JT+2 columns incl. halo] x [32 members], the consumer warps run the explicit
// stencil + sources + forward elimination out of shared memory and never wait on a global load
// of state data.  Out-of-range halo columns (j = -1, ny) and levels past nz are zero-filled by
// the TMA unit (their stencil coefficients are zero).  The back substitution reads the
// forward-sweep intermediates back from the output buffer (L2, evict_last) as in nkb_stage.cu.
//
// Same arithmetic, same order of operations as nkb_stage.cu: results are bit-identical.

#include <cuda.h>

#include <cstdlib>
#include <map>
#include <mutex>

#include "nkb_stage_dev.cuh"

namespace nkb {

constexpr int kBM = 32;  // members per CTA (16 lanes x 2 members)
constexpr int kBX = 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "l"(hint)
        : "memory");
}

struct TmaMaps {
    CUtensorMap u[2];
};

// grid: (member blocks of 32, column tiles of JT, tracer groups); block: (16, JT + 2) threads:
// threadIdx.y < JT are consumers (one column each), the last 32 threads are the producer warp.
template <int KIND, int TG, int NIN, int KC, int NS>
__global__ void __launch_bounds__(288) stage_tma_kernel(const StageArgs p, const __grid_constant__ TmaMaps maps) {
    constexpr int MPT = 2;
    extern __shared__ unsigned char smem_dyn[];
    const int jt = blockDim.y - 2;
    const int nz = p.nz, ny = p.ny;
    const int rows = jt + 2;
    // carve: barriers, then NS stages of NIN*TG boxes
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~uintptr_t(127));
    uint64_t *full = reinterpret_cast<uint64_t *>(base);
    uint64_t *empty = full + NS;
    const size_t box_doubles = (size_t)KC * rows * kBM;
    double *ring = reinterpret_cast<double *>(base + 128);
    const uint32_t stage_bytes = (uint32_t)(NIN * TG * box_doubles * sizeof(double));

    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int n_cons = jt * kBX;
    const int n_cons_warps = n_cons / 32;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], n_cons_warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int b0 = blockIdx.x * kBM;
    const int j0 = blockIdx.y * jt;
    const int tr0 = blockIdx.z * TG;
    const int nchunk = (nz + KC - 1) / KC;

    if (tid >= n_cons) {
        // ===== producer warp: one elected lane issues the box copies =====
        if (tid == n_cons) {
            const uint64_t hint = 0x12F0000000000000ull;  // L2 evict_first: the inputs are streamed
            for (int c = 0; c < nchunk; ++c) {
                const int s = c % NS;
                const uint32_t ph = (c / NS) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                double *dst = ring + (size_t)s * NIN * TG * box_doubles;
#pragma unroll
                for (int i = 0; i < NIN; ++i)
#pragma unroll
                    for (int g = 0; g < TG; ++g)
                        tma_load_4d(dst + (size_t)(i * TG + g) * box_doubles, &maps.u[i], &full[s], b0, j0 - 1, c * KC,
                                    tr0 + g, hint);
            }
        }
        return;
    }

    // ===== consumers =====
    const int lane16 = threadIdx.x;
    const int jj = threadIdx.y;
    const int j = j0 + jj;
    const int b = b0 + lane16 * MPT;
    const bool active = (j < ny) && (b < p.B);
    const int jc = (j < ny) ? j : ny - 1;  // inactive threads still follow the barrier protocol
    const size_t ldb = p.ldb;
    const size_t plane = (size_t)nz * ny;

    ColPtrs<TG, NIN> cp;
    cp.stepk = (size_t)ny * ldb;
    cp.stepk4 = (size_t)ny * 4;
    cp.pol_first = policy_evict_first();
    cp.pol_last = policy_evict_last();
    double aff[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) {
        const size_t off0 = ((size_t)(tr0 + g) * plane + jc) * ldb + (active ? b : 0);
        cp.out[g] = p.out + off0;
        const int cls = p.class_of[tr0 + g];
        cp.tri4[g] = p.tri + ((size_t)cls * plane + jc) * 4;
        aff[g] = __ldg(p.aff + (size_t)cls * ny + jc);
    }
    cp.est4 = p.est ? p.est + (size_t)jc * 4 : nullptr;
    cp.src2 = p.src2 ? p.src2 + (size_t)jc * 2 : nullptr;
    cp.light = p.light ? p.light + jc : nullptr;
    const bool has_e = (cp.est4 != nullptr);

    Vec<MPT> yprev[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) yprev[g] = Vec<MPT>::splat(0.0);

    // ---- forward elimination ----
    for (int c = 0; c < nchunk; ++c) {
        const int s = c % NS;
        const uint32_t ph = (c / NS) & 1;
        const int k0 = c * KC;
        // member-independent coefficients of the chunk: issued before waiting for the state data
        double eL[KC], eC[KC], eR[KC], frc[KC][NIN], mk[KC][TG], lgt[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            const bool lv = (k0 + q < nz);
            if constexpr (KIND == NKB_MOD_FORCED_FILE) {
                if constexpr (NIN == 2) {
                    const double2 f = lv ? __ldg(reinterpret_cast<const double2 *>(cp.src2 + q * (cp.stepk4 >> 1)))
                                         : make_double2(0.0, 0.0);
                    frc[q][0] = f.x;
                    frc[q][1] = f.y;
                } else {
                    frc[q][0] = lv ? __ldg(cp.src2 + q * (cp.stepk4 >> 1)) : 0.0;
                }
            } else {
#pragma unroll
                for (int i = 0; i < NIN; ++i) frc[q][i] = 0.0;
            }
            if constexpr (KIND == NKB_MOD_PHOSPHORUS) lgt[q] = lv ? __ldg(cp.light + q * (cp.stepk4 >> 2)) : 0.0;
            else lgt[q] = 0.0;
            eL[q] = eC[q] = eR[q] = 0.0;
            if (has_e && lv) {
                const double2 e01 = __ldg(reinterpret_cast<const double2 *>(cp.est4 + q * cp.stepk4));
                eL[q] = e01.x;
                eC[q] = e01.y;
                eR[q] = __ldg(cp.est4 + q * cp.stepk4 + 2);
            }
#pragma unroll
            for (int g = 0; g < TG; ++g) mk[q][g] = lv ? __ldg(cp.tri4[g] + q * cp.stepk4 + 2) : 0.0;
        }
        mbar_wait(&full[s], ph);
        const double *stage = ring + (size_t)s * NIN * TG * box_doubles;
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            const int k = k0 + q;
            if (k < nz) {
                Vec<MPT> rhs[TG];
#pragma unroll
                for (int g = 0; g < TG; ++g) rhs[g] = Vec<MPT>::splat(0.0);
#pragma unroll
                for (int i = 0; i < NIN; ++i) {
                    Vec<MPT> cv[TG], e[TG], sv[TG];
#pragma unroll
                    for (int g = 0; g < TG; ++g) {
                        const double *row = stage + (size_t)(i * TG + g) * box_doubles + ((size_t)q * rows + jj) * kBM +
                                            lane16 * MPT;
                        const Vec<MPT> cl = {*reinterpret_cast<const double2 *>(row)};
                        cv[g].v = *reinterpret_cast<const double2 *>(row + kBM);
                        const Vec<MPT> cr = {*reinterpret_cast<const double2 *>(row + 2 * kBM)};
                        if (has_e) e[g] = fma_s(eL[q], cl, fma_s(eR[q], cr, mul_s(eC[q], cv[g])));
                        else e[g] = Vec<MPT>::splat(0.0);
                    }
                    explicit_sources<KIND, TG, MPT>(p, tr0, lgt[q], frc[q][i], cv, sv);
#pragma unroll
                    for (int g = 0; g < TG; ++g) {
                        rhs[g] = fma_s(p.a[i], cv[g], rhs[g]);
                        rhs[g] = fma_s(p.he[i], add_v(e[g], sv[g]), rhs[g]);
                    }
                }
#pragma unroll
                for (int g = 0; g < TG; ++g) {
                    if (k == 0) rhs[g] = add_v(rhs[g], Vec<MPT>::splat(aff[g]));
                    yprev[g] = fma_s(-mk[q][g], yprev[g], rhs[g]);
                    if (active) st_hint(cp.out[g] + q * cp.stepk, yprev[g], cp.pol_last);
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        // advance the coefficient / output pointers by one chunk
#pragma unroll
        for (int g = 0; g < TG; ++g) {
            cp.out[g] += KC * cp.stepk;
            cp.tri4[g] += KC * cp.stepk4;
        }
        if (has_e) cp.est4 += KC * cp.stepk4;
        if constexpr (KIND == NKB_MOD_FORCED_FILE) cp.src2 += KC * (cp.stepk4 >> 1);
        if constexpr (KIND == NKB_MOD_PHOSPHORUS) cp.light += KC * (cp.stepk4 >> 2);
    }
    if (!active) return;

    // ---- back substitution (pointers address level nchunk*KC; step back to level nz-1) ----
    const ptrdiff_t over = (ptrdiff_t)nchunk * KC - (nz - 1);
#pragma unroll
    for (int g = 0; g < TG; ++g) {
        cp.out[g] -= over * (ptrdiff_t)cp.stepk;
        cp.tri4[g] -= over * (ptrdiff_t)cp.stepk4;
        cp.sub[g] = p.sub ? p.sub + (cp.out[g] - p.out) : nullptr;
    }
    Vec<MPT> xnext[TG];
#pragma unroll
    for (int g = 0; g < TG; ++g) xnext[g] = Vec<MPT>::splat(0.0);
    constexpr int KB = (TG >= 3) ? 2 : 8;
    int k = nz - 1;
    for (; k - KB + 1 >= 0; k -= KB) backward_chunk<TG, NIN, MPT, KB>(cp, k, 0, nullptr, 0, 0, xnext);
    for (; k >= 0; --k) backward_chunk<TG, NIN, MPT, 1>(cp, k, 0, nullptr, 0, 0, xnext);
}

// ---- host side: tensor maps -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &ptr, 12000, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static int make_map(CUtensorMap *map, const double *ptr, int T, int nz, int ny, int ldb, int rows, int kc) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return 1;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)ldb, (cuuint64_t)ny, (cuuint64_t)nz, (cuuint64_t)T};
    const cuuint64_t strides[3] = {(cuuint64_t)ldb * 8, (cuuint64_t)ny * ldb * 8, (cuuint64_t)nz * ny * ldb * 8};
    const cuuint32_t box[4] = {(cuuint32_t)kBM, (cuuint32_t)rows, (cuuint32_t)kc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
        return 1;
    }
    return 0;
}

static int env_int_tma(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

bool tma_path_usable(const StageArgs &a) {
    if (env_int_tma("NKB_TMA", 1) == 0) return false;
    return a.B >= 32 && (a.ldb % 32) == 0 && a.ny >= 2 && encode_fn() != nullptr;
}

template <int KIND, int TG, int NIN>
static int launch_tma_t(const StageArgs &a, cudaStream_t st) {
    constexpr int KC = (TG >= 3) ? 2 : 4;
    constexpr int NS = (TG >= 3) ? 3 : 4;
    int jt = env_int_tma("NKB_TMA_JT", 8);
    if (jt > a.ny) jt = a.ny;
    if (jt & 1) jt += 1;  // consumer warps are whole (16 lanes x 2 columns per warp)
    if (jt > 16) jt = 16;
    const int rows = jt + 2;
    TmaMaps maps;
    for (int i = 0; i < NIN; ++i)
        if (make_map(&maps.u[i], a.u[i], a.T, a.nz, a.ny, a.ldb, rows, KC)) return 1;
    if (NIN == 1) maps.u[1] = maps.u[0];
    size_t smem = 128 + 128 + (size_t)NS * NIN * TG * KC * rows * kBM * sizeof(double);
    // optional cap on resident CTAs per SM (bounds the in-flight forward-sweep intermediates)
    const size_t pad = (size_t)env_int_tma("NKB_TMA_MIN_SMEM_KB", 0) * 1024;
    if (smem < pad) smem = pad;
    auto kern = stage_tma_kernel<KIND, TG, NIN, KC, NS>;
    static unsigned long long attr_mask = 0;
    if (nkb::first_use_on_device(attr_mask)) {
        NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (smem > 227 * 1024) {
        set_error("stage_tma_kernel: tile does not fit in shared memory");
        return 2;
    }
    dim3 block(kBX, jt + 2), grid((a.B + kBM - 1) / kBM, (a.ny + jt - 1) / jt, a.T / TG);
    kern<<<grid, block, smem, st>>>(a, maps);
    count_launch();
    return 0;
}

int launch_stage_tma(int kind, int nin, const StageArgs &a, cudaStream_t st) {
    int rc = 0;
#define NKB_DISPATCH_TMA(K, G)                        \
    if (nin == 1) rc = launch_tma_t<K, G, 1>(a, st);  \
    else rc = launch_tma_t<K, G, 2>(a, st);
    switch (kind) {
        case NKB_MOD_LINEAR: NKB_DISPATCH_TMA(NKB_MOD_LINEAR, 1); break;
        case NKB_MOD_FORCED_FILE: NKB_DISPATCH_TMA(NKB_MOD_FORCED_FILE, 1); break;
        case NKB_MOD_PHOSPHORUS: NKB_DISPATCH_TMA(NKB_MOD_PHOSPHORUS, 3); break;
        default: set_error("launch_stage_tma: unsupported module kind"); return 2;
    }
#undef NKB_DISPATCH_TMA
    return rc;
}

}  // namespace nkb
