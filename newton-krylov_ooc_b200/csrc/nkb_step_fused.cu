// Fused IMEX step kernel — one launch per ARS(2,2,2) time step (both implicit stages), so that a
// model state is read from HBM once (+ halo) and written once per step instead of 3 reads and
// 2 writes with one launch per stage (nkb_stage_tma.cu).
//
// Replaces, per step and for all members at once, what the reference does inside
// solve_ivp(Radau) for py_driver_2d (nk_ooc/py_driver_2d/model_state.py:94-121):
// TracerModuleState.comp_tend (tracer_module_state.py:98-108: advection.py:51-76,
// horiz_mix.py:48-67, vert_mix.py:24-41, iage.py:22-41 / forced.py:114-154) and the implicit
// solves of its Jacobian (vert_mix.py:140-188).
//
// Work decomposition (B200: 148 SMs, one persistent CTA per SM, static tile round-robin)
//   tile  = 16 members x 14 interior columns (+1 halo column each side for the stage-1 solution,
//           +2 for the state) x all levels of one tracer;  rows of 16 members = 128 bytes.
//   warp  = 16 columns x 2 member pairs; thread = (column, 2 adjacent members), lane = column +
//           16*pair.  Warps are self-contained: the only cross-thread exchange (stage-1 solution
//           of the neighbour columns) is a warp shuffle.
//   three sweeps over depth per tile:
//     A (top->bottom)  stage-1 right-hand side (explicit horizontal stencil + sources) and LU
//                      forward elimination;  intermediates y1_k -> TMEM
//     B (bottom->top)  stage-1 back substitution u1_k, exchange with the neighbour columns,
//                      stage-2 right-hand side (the explicit term of stage 1 is recovered from
//                      y1_k + m1_k y1_{k-1} and a re-read of the state, which hits L2) and UL
//                      elimination upwards;  y2_k overwrites y1_k in TMEM
//     C (top->bottom)  stage-2 substitution u2_k -> shared-memory staging -> TMA store
//   Tensor memory is used as a per-lane scratchpad (tcgen05.st/ld 32x32b): 512 columns x 4 B =
//   2 members x 128 levels of float64 per thread — the forward-sweep intermediates never leave
//   the SM.  No tcgen05.mma is issued: nothing here is a contraction.
//   All state and coefficient traffic global->shared goes through TMA (cp.async.bulk.tensor,
//   128-byte swizzle for the state boxes so that lane=column accesses are bank-conflict free)
//   into one mbarrier full/empty ring shared by the three sweeps; a producer warp runs ahead of
//   the consumers across sweeps and tiles, a store warp drains the output staging ring.

#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "nkb_common.cuh"

namespace nkb {

constexpr int FS_KC = 4;      // levels per chunk (one tcgen05 x16 access = 4 levels x 2 members)
constexpr int FS_COLS = 16;   // stage-1 columns per tile (= lanes per member pair)
constexpr int FS_UCOLS = 18;  // state columns per tile (halo of 2)
constexpr int FS_JT = 14;     // max interior columns per tile
constexpr int FS_MEM = 16;    // members per tile
constexpr int FS_NCW = 4;     // consumer warps (one per TMEM lane quarter)
constexpr int FS_NS = 8;      // load ring slots
constexpr int FS_NO = 3;      // output staging slots
constexpr int FS_UBYTES = FS_KC * FS_UCOLS * FS_MEM * 8;  // 9216
constexpr int FS_PLANE = FS_KC * FS_COLS * 8;             // 512
constexpr int FS_NPLANES = 8;
constexpr int FS_SLOT = FS_UBYTES + FS_NPLANES * FS_PLANE;  // 13312 = 13 * 1024
constexpr int FS_OUT = FS_KC * FS_JT * FS_MEM * 8;          // 7168 = 7 * 1024
constexpr int FS_SMEM = 1024 + FS_NS * FS_SLOT + FS_NO * FS_OUT + 1024;
constexpr int FS_THREADS = (FS_NCW + 2) * 32;
static_assert(FS_SLOT % 1024 == 0 && FS_OUT % 1024 == 0, "swizzled boxes need 1024-byte aligned bases");

struct StepArgs {
    int nz, ny, B, T, ncls, n_steps;
    int nct, jt, nmb, ntiles;  // column tiles, interior columns per tile, member blocks, total
    int step;
    int dbg;  // debug switches (NKB_FUSED_DBG): 1 no TMEM, 2 no TMA store, 4 single cache hint, 8 no alloc
    int class_of[NKB_MAX_TRACERS];
    double src_const[NKB_MAX_TRACERS];
    double sink_thres_r;
    double hg, a0, a1, r, he1;  // gamma*h; stage-2 weights (see nkb_api.cu)
    const double *aff1, *aff2;  // [ncls][ny] of the two stages of this step
};

struct StepMaps {
    CUtensorMap uin, uout, est, ftab, src;
};

struct D2 {
    double x, y;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fs_smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void fs_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fs_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fs_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void fs_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fs_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fs_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = fs_smem_u32(bar);
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
__device__ __forceinline__ void fs_tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                               int c2, int c3, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        :
        : "r"(fs_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(fs_smem_u32(bar)), "r"(c0), "r"(c1),
          "r"(c2), "r"(c3), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void fs_tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                               int c2, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        :
        : "r"(fs_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(fs_smem_u32(bar)), "r"(c0), "r"(c1),
          "r"(c2), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void fs_tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2,
                                                int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(map)), "r"(fs_smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void fs_tmem_st16(uint32_t taddr, const D2 (&v)[FS_KC]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15, %16};"
        :
        : "r"(taddr), "r"(__double2loint(v[0].x)), "r"(__double2hiint(v[0].x)), "r"(__double2loint(v[0].y)),
          "r"(__double2hiint(v[0].y)), "r"(__double2loint(v[1].x)), "r"(__double2hiint(v[1].x)),
          "r"(__double2loint(v[1].y)), "r"(__double2hiint(v[1].y)), "r"(__double2loint(v[2].x)),
          "r"(__double2hiint(v[2].x)), "r"(__double2loint(v[2].y)), "r"(__double2hiint(v[2].y)),
          "r"(__double2loint(v[3].x)), "r"(__double2hiint(v[3].x)), "r"(__double2loint(v[3].y)),
          "r"(__double2hiint(v[3].y))
        : "memory");
}
// the load is asynchronous: the raw words are only valid after fs_tmem_wait_ld (which takes them
// as in/out operands so that no use can be scheduled ahead of the wait)
struct Raw16 {
    uint32_t w[16];
};
__device__ __forceinline__ void fs_tmem_ld16(uint32_t taddr, Raw16 &r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
          "=r"(r.w[7]), "=r"(r.w[8]), "=r"(r.w[9]), "=r"(r.w[10]), "=r"(r.w[11]), "=r"(r.w[12]), "=r"(r.w[13]),
          "=r"(r.w[14]), "=r"(r.w[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void fs_tmem_wait_ld(Raw16 &r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]),
                   "+r"(r.w[7]), "+r"(r.w[8]), "+r"(r.w[9]), "+r"(r.w[10]), "+r"(r.w[11]), "+r"(r.w[12]),
                   "+r"(r.w[13]), "+r"(r.w[14]), "+r"(r.w[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void fs_unpack(const Raw16 &r, D2 (&v)[FS_KC]) {
#pragma unroll
    for (int q = 0; q < FS_KC; ++q) {
        v[q].x = __hiloint2double((int)r.w[4 * q + 1], (int)r.w[4 * q]);
        v[q].y = __hiloint2double((int)r.w[4 * q + 3], (int)r.w[4 * q + 2]);
    }
}
__device__ __forceinline__ void fs_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ D2 fs_fma(double a, D2 x, D2 y) { return {fma(a, x.x, y.x), fma(a, x.y, y.y)}; }
__device__ __forceinline__ D2 fs_mul(double a, D2 x) { return {a * x.x, a * x.y}; }
__device__ __forceinline__ D2 fs_add(D2 x, D2 y) { return {x.x + y.x, x.y + y.y}; }
__device__ __forceinline__ D2 fs_sub(D2 x, D2 y) { return {x.x - y.x, x.y - y.y}; }
__device__ __forceinline__ D2 fs_ld(const unsigned char *p) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    return {v.x, v.y};
}
__device__ __forceinline__ D2 fs_shfl_up(D2 v) {
    return {__shfl_up_sync(0xffffffffu, v.x, 1, 16), __shfl_up_sync(0xffffffffu, v.y, 1, 16)};
}
__device__ __forceinline__ D2 fs_shfl_down(D2 v) {
    return {__shfl_down_sync(0xffffffffu, v.x, 1, 16), __shfl_down_sync(0xffffffffu, v.y, 1, 16)};
}

// explicit source of one tracer (iage.py:39 constant; forced.py:141-151 forcing record with the
// sink_thres limiter) — same expressions as explicit_sources() in nkb_stage_dev.cuh
template <int KIND>
__device__ __forceinline__ D2 fs_source(double srcc, double thr_r, double frc, D2 c) {
    if constexpr (KIND == NKB_MOD_LINEAR) {
        return {srcc, srcc};
    } else {
        const double qx = thr_r * c.x, qy = thr_r * c.y;
        const bool lim = (thr_r > 0.0 && frc < 0.0);
        return {(lim && qx > 0.0 && qx < 1.0) ? frc * qx : frc, (lim && qy > 0.0 && qy < 1.0) ? frc * qy : frc};
    }
}

constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// coefficient plane q of a ring slot: [FS_KC][FS_COLS] doubles
//   sweep A: 0 eL, 1 eC, 2 eR, 3 m1, 4 frc(t_n)
//   sweep B: 0 eL, 1 eC, 2 eR, 3 ib1, 4 g1, 5 m1, 6 m2, 7 frc(t_n + gamma h)
//   sweep C: 0 ib2, 1 g2
template <int KIND, bool HAS_E>
__global__ void __launch_bounds__(FS_THREADS, 1) step_fused_kernel(const StepArgs p,
                                                                    const __grid_constant__ StepMaps maps) {
    extern __shared__ unsigned char fs_smem[];
    unsigned char *base =
        reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(fs_smem) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(base);
    uint64_t *empty = full + FS_NS;
    uint64_t *ofull = empty + FS_NS;
    uint64_t *oempty = ofull + FS_NO;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + 960);
    unsigned char *ring = base + 1024;
    unsigned char *oring = ring + FS_NS * FS_SLOT;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FS_NS; ++s) {
            fs_mbar_init(&full[s], 1);
            fs_mbar_init(&empty[s], FS_NCW);
        }
        for (int s = 0; s < FS_NO; ++s) {
            fs_mbar_init(&ofull[s], FS_NCW);
            fs_mbar_init(&oempty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && !(p.dbg & 8)) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         fs_smem_u32(tmem_holder))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_holder);

    const uint64_t hN = (p.dbg & 4) ? kEvictFirst : kEvictNormal, hL = (p.dbg & 4) ? kEvictFirst : kEvictLast;
    const int nz = p.nz, ny = p.ny;
    const int nchunk = (nz + FS_KC - 1) / FS_KC;
    constexpr int NPA = (HAS_E ? 3 : 0) + 1 + (KIND == NKB_MOD_FORCED_FILE ? 1 : 0);
    constexpr int NPB = (HAS_E ? 3 : 0) + 4 + (KIND == NKB_MOD_FORCED_FILE ? 1 : 0);

    if (warp == FS_NCW) {
        // ===== producer: one lane issues every TMA load of this CTA, in consumption order =====
        if (lane == 0) {
            uint32_t g = 0;
            const bool no4 = (p.dbg & 16) != 0, no3 = (p.dbg & 32) != 0;
            auto fs_tma_load_4d = [&](void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3,
                                      uint64_t hint) {
                if (!no4) nkb::fs_tma_load_4d(dst, map, bar, c0, c1, c2, c3, hint);
            };
            auto fs_tma_load_3d = [&](void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                      uint64_t hint) {
                if (!no3) nkb::fs_tma_load_3d(dst, map, bar, c0, c1, c2, hint);
            };
            auto fs_mbar_expect_tx = [&](uint64_t *bar, uint32_t bytes) {
                uint32_t b = 0;
                if (!no4 && bytes > 2 * FS_PLANE) b += FS_UBYTES;
                if (!no3) b += (bytes > 2 * FS_PLANE) ? bytes - FS_UBYTES : bytes;
                nkb::fs_mbar_expect_tx(bar, b);
            };
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int mb = tile % p.nmb;
                const int ct = (tile / p.nmb) % p.nct;
                const int tr = tile / (p.nmb * p.nct);
                const int m0 = mb * FS_MEM, j0 = ct * p.jt;
                const int zt = (p.step * p.ncls + p.class_of[tr]) * 6;
                for (int sweep = 0; sweep < 3; ++sweep) {
                    for (int cc = 0; cc < nchunk; ++cc) {
                        const int c = (sweep == 1) ? nchunk - 1 - cc : cc;
                        const int k0 = c * FS_KC;
                        const uint32_t s = g % FS_NS, ph = (g / FS_NS) & 1;
                        fs_mbar_wait(&empty[s], ph ^ 1);
                        unsigned char *sb = ring + s * FS_SLOT;
                        unsigned char *pl = sb + FS_UBYTES;
                        if (sweep == 0) {
                            fs_mbar_expect_tx(&full[s], FS_UBYTES + NPA * FS_PLANE);
                            fs_tma_load_4d(sb, &maps.uin, &full[s], m0, j0 - 2, k0, tr, hN);
                            if constexpr (HAS_E) {
#pragma unroll
                                for (int q = 0; q < 3; ++q)
                                    fs_tma_load_3d(pl + q * FS_PLANE, &maps.est, &full[s], j0, k0, q, hL);
                            }
                            fs_tma_load_3d(pl + 3 * FS_PLANE, &maps.ftab, &full[s], j0, k0, zt + 0, hL);
                            if constexpr (KIND == NKB_MOD_FORCED_FILE)
                                fs_tma_load_3d(pl + 4 * FS_PLANE, &maps.src, &full[s], j0, k0, 2 * p.step,
                                               hL);
                        } else if (sweep == 1) {
                            fs_mbar_expect_tx(&full[s], FS_UBYTES + NPB * FS_PLANE);
                            fs_tma_load_4d(sb, &maps.uin, &full[s], m0, j0 - 2, k0, tr, kEvictFirst);
                            if constexpr (HAS_E) {
#pragma unroll
                                for (int q = 0; q < 3; ++q)
                                    fs_tma_load_3d(pl + q * FS_PLANE, &maps.est, &full[s], j0, k0, q, hL);
                            }
                            fs_tma_load_3d(pl + 3 * FS_PLANE, &maps.ftab, &full[s], j0, k0, zt + 1, hL);
                            fs_tma_load_3d(pl + 4 * FS_PLANE, &maps.ftab, &full[s], j0, k0, zt + 2, hL);
                            fs_tma_load_3d(pl + 5 * FS_PLANE, &maps.ftab, &full[s], j0, k0, zt + 0, hL);
                            fs_tma_load_3d(pl + 6 * FS_PLANE, &maps.ftab, &full[s], j0, k0, zt + 3, hL);
                            if constexpr (KIND == NKB_MOD_FORCED_FILE)
                                fs_tma_load_3d(pl + 7 * FS_PLANE, &maps.src, &full[s], j0, k0, 2 * p.step + 1,
                                               hL);
                        } else {
                            fs_mbar_expect_tx(&full[s], 2 * FS_PLANE);
                            fs_tma_load_3d(pl + 0 * FS_PLANE, &maps.ftab, &full[s], j0, k0, zt + 4, hL);
                            fs_tma_load_3d(pl + 1 * FS_PLANE, &maps.ftab, &full[s], j0, k0, zt + 5, hL);
                        }
                        ++g;
                    }
                }
            }
        }
    } else if (warp == FS_NCW + 1) {
        // ===== store warp: drains the output staging ring with TMA stores =====
        if (lane == 0) {
            uint32_t go = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int mb = tile % p.nmb;
                const int ct = (tile / p.nmb) % p.nct;
                const int tr = tile / (p.nmb * p.nct);
                for (int c = 0; c < nchunk; ++c) {
                    const uint32_t s = go % FS_NO, ph = (go / FS_NO) & 1;
                    fs_mbar_wait(&ofull[s], ph);
                    if (!(p.dbg & 2)) fs_tma_store_4d(&maps.uout, oring + s * FS_OUT, mb * FS_MEM, ct * p.jt, c * FS_KC, tr);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    fs_mbar_arrive(&oempty[s]);
                    ++go;
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // ===== consumers =====
        const int col = lane & 15;
        const int c16 = 2 * warp + (lane >> 4);  // 16-byte chunk (member pair) within the 128-byte row
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        int offU[FS_KC][3], offO[FS_KC];
#pragma unroll
        for (int q = 0; q < FS_KC; ++q) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int r = q * FS_UCOLS + col + d;
                offU[q][d] = 128 * r + 16 * (c16 ^ (r & 7));
            }
            const int ro = q * p.jt + col - 1;
            offO[q] = 128 * ro + 16 * (c16 ^ (ro & 7));
        }
        const bool interior = (col >= 1 && col <= p.jt);
        const double thr_r = p.sink_thres_r;
        uint32_t g = 0, go = 0;

        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int ct = (tile / p.nmb) % p.nct;
            const int tr = tile / (p.nmb * p.nct);
            const int j = ct * p.jt - 1 + col;
            const int cls = p.class_of[tr];
            double aff1 = 0.0, aff2 = 0.0;
            if (j >= 0 && j < ny) {
                aff1 = __ldg(p.aff1 + (size_t)cls * ny + j);
                aff2 = __ldg(p.aff2 + (size_t)cls * ny + j);
            }
            const double srcc = p.src_const[tr];

            // ---------------- sweep A: stage-1 rhs + LU forward elimination, top -> bottom ----------------
            D2 yprev = {0.0, 0.0};
            for (int c = 0; c < nchunk; ++c) {
                const uint32_t s = g % FS_NS, ph = (g / FS_NS) & 1;
                fs_mbar_wait(&full[s], ph);
                const unsigned char *sb = ring + s * FS_SLOT;
                const double *pl = reinterpret_cast<const double *>(sb + FS_UBYTES);
                D2 yb[FS_KC];
#pragma unroll
                for (int q = 0; q < FS_KC; ++q) {
                    const D2 cv = fs_ld(sb + offU[q][1]);
                    D2 es;  // e + s
                    double frc = 0.0;
                    if constexpr (KIND == NKB_MOD_FORCED_FILE) frc = pl[4 * 64 + q * 16 + col];
                    const D2 sv = fs_source<KIND>(srcc, thr_r, frc, cv);
                    if constexpr (HAS_E) {
                        const D2 cl = fs_ld(sb + offU[q][0]);
                        const D2 cr = fs_ld(sb + offU[q][2]);
                        const double eL = pl[q * 16 + col], eC = pl[64 + q * 16 + col], eR = pl[128 + q * 16 + col];
                        es = fs_add(fs_fma(eL, cl, fs_fma(eR, cr, fs_mul(eC, cv))), sv);
                    } else {
                        es = fs_add(D2{0.0, 0.0}, sv);
                    }
                    const double m1 = pl[3 * 64 + q * 16 + col];
                    D2 rhs = fs_fma(p.hg, es, cv);
                    if (c == 0 && q == 0) rhs = fs_add(rhs, D2{aff1, aff1});
                    yprev = fs_fma(-m1, yprev, rhs);
                    yb[q] = yprev;
                }
                __syncwarp();
                if (lane == 0) fs_mbar_arrive(&empty[s]);
                if (!(p.dbg & 1)) fs_tmem_st16(taddr + c * 16, yb);
                ++g;
            }
            if (!(p.dbg & 1)) fs_tmem_wait_st();

            // ------- sweep B: stage-1 back substitution + stage-2 rhs + UL elimination, bottom -> top -------
            {
                Raw16 rcur, rnx;
                D2 ycur[FS_KC], ynx[FS_KC];
                for (int i = 0; i < 16; ++i) rcur.w[i] = rnx.w[i] = 0;
                if (!(p.dbg & 1)) { fs_tmem_ld16(taddr + (nchunk - 1) * 16, rcur);
                fs_tmem_wait_ld(rcur); }
                fs_unpack(rcur, ycur);
                D2 u1n = {0.0, 0.0}, y2n = {0.0, 0.0};
                for (int c = nchunk - 1; c >= 0; --c) {
                    if (c > 0 && !(p.dbg & 1)) fs_tmem_ld16(taddr + (c - 1) * 16, rnx);
                    const uint32_t s = g % FS_NS, ph = (g / FS_NS) & 1;
                    fs_mbar_wait(&full[s], ph);
                    if (c > 0) {
                        if (!(p.dbg & 1)) fs_tmem_wait_ld(rnx);
                        fs_unpack(rnx, ynx);
                    } else {
#pragma unroll
                        for (int q = 0; q < FS_KC; ++q) ynx[q] = D2{0.0, 0.0};
                    }
                    const unsigned char *sb = ring + s * FS_SLOT;
                    const double *pl = reinterpret_cast<const double *>(sb + FS_UBYTES);
                    D2 yb[FS_KC];
#pragma unroll
                    for (int q = FS_KC - 1; q >= 0; --q) {
                        const D2 y1 = ycur[q];
                        const D2 y1m = (q > 0) ? ycur[q > 0 ? q - 1 : 0] : ynx[FS_KC - 1];
                        const double ib1 = pl[3 * 64 + q * 16 + col], g1 = pl[4 * 64 + q * 16 + col];
                        const double m1 = pl[5 * 64 + q * 16 + col], m2 = pl[6 * 64 + q * 16 + col];
                        const D2 u1 = fs_fma(-g1, u1n, fs_mul(ib1, y1));
                        u1n = u1;
                        D2 rhs1 = fs_fma(m1, y1m, y1);
                        if (c == 0 && q == 0) rhs1 = fs_sub(rhs1, D2{aff1, aff1});
                        const D2 un = fs_ld(sb + offU[q][1]);
                        const D2 pp = fs_fma(p.r, fs_sub(rhs1, un), fs_mul(p.a0, un));
                        double frc = 0.0;
                        if constexpr (KIND == NKB_MOD_FORCED_FILE) frc = pl[7 * 64 + q * 16 + col];
                        const D2 sv = fs_source<KIND>(srcc, thr_r, frc, u1);
                        D2 es;
                        if constexpr (HAS_E) {
                            const D2 ul = fs_shfl_up(u1), ur = fs_shfl_down(u1);
                            const double eL = pl[q * 16 + col], eC = pl[64 + q * 16 + col],
                                         eR = pl[128 + q * 16 + col];
                            es = fs_add(fs_fma(eL, ul, fs_fma(eR, ur, fs_mul(eC, u1))), sv);
                        } else {
                            es = fs_add(D2{0.0, 0.0}, sv);
                        }
                        D2 rhs2 = fs_fma(p.he1, es, fs_fma(p.a1, u1, pp));
                        if (c == 0 && q == 0) rhs2 = fs_add(rhs2, D2{aff2, aff2});
                        y2n = fs_fma(-m2, y2n, rhs2);
                        yb[q] = y2n;
                    }
                    __syncwarp();
                    if (lane == 0) fs_mbar_arrive(&empty[s]);
                    if (!(p.dbg & 1)) fs_tmem_st16(taddr + c * 16, yb);
#pragma unroll
                    for (int q = 0; q < FS_KC; ++q) ycur[q] = ynx[q];
                    ++g;
                }
                if (!(p.dbg & 1)) fs_tmem_wait_st();
            }

            // ---------------- sweep C: stage-2 substitution, top -> bottom, staged TMA store ----------------
            {
                Raw16 rcur, rnx;
                D2 ycur[FS_KC];
                for (int i = 0; i < 16; ++i) rcur.w[i] = rnx.w[i] = 0;
                if (!(p.dbg & 1)) { fs_tmem_ld16(taddr, rcur);
                fs_tmem_wait_ld(rcur); }
                fs_unpack(rcur, ycur);
                D2 u2p = {0.0, 0.0};
                for (int c = 0; c < nchunk; ++c) {
                    if (c + 1 < nchunk && !(p.dbg & 1)) fs_tmem_ld16(taddr + (c + 1) * 16, rnx);
                    const uint32_t s = g % FS_NS, ph = (g / FS_NS) & 1;
                    const uint32_t so = go % FS_NO, pho = (go / FS_NO) & 1;
                    fs_mbar_wait(&full[s], ph);
                    fs_mbar_wait(&oempty[so], pho ^ 1);
                    const double *pl = reinterpret_cast<const double *>(ring + s * FS_SLOT + FS_UBYTES);
                    unsigned char *ob = oring + so * FS_OUT;
#pragma unroll
                    for (int q = 0; q < FS_KC; ++q) {
                        const double ib2 = pl[q * 16 + col], g2 = pl[64 + q * 16 + col];
                        u2p = fs_fma(-g2, u2p, fs_mul(ib2, ycur[q]));
                        if (interior) *reinterpret_cast<double2 *>(ob + offO[q]) = make_double2(u2p.x, u2p.y);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        fs_mbar_arrive(&empty[s]);
                        fs_mbar_arrive(&ofull[so]);
                    }
                    if (c + 1 < nchunk) {
                        if (!(p.dbg & 1)) fs_tmem_wait_ld(rnx);
                        fs_unpack(rnx, ycur);
                    }
                    ++g;
                    ++go;
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && !(p.dbg & 8)) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// out = out - x0 (the final F = x(T) - x(0), once per model year)
__global__ void sub_inplace_kernel(double *__restrict__ out, const double *__restrict__ x0, size_t n2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    double2 a = reinterpret_cast<double2 *>(out)[i];
    const double2 b = reinterpret_cast<const double2 *>(x0)[i];
    a.x -= b.x;
    a.y -= b.y;
    reinterpret_cast<double2 *>(out)[i] = a;
}

static int fs_env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*FsEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static FsEncodeTiledFn fs_encode_fn() {
    static FsEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &ptr, 12000, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<FsEncodeTiledFn>(ptr);
    }
    return fn;
}

static int fs_encode(CUtensorMap *map, const void *ptr, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                     const cuuint32_t *box, CUtensorMapSwizzle swz) {
    FsEncodeTiledFn fn = fs_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return 1;
    }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, const_cast<void *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
        return 1;
    }
    return 0;
}

// geometry of the column tiling: nct tiles of jt <= 14 interior columns
static void fs_col_tiles(int ny, int &nct, int &jt) {
    int jmax = fs_env_int("NKB_FUSED_JT", FS_JT);
    if (jmax < 1) jmax = 1;
    if (jmax > FS_JT) jmax = FS_JT;
    nct = (ny + jmax - 1) / jmax;
    jt = (ny + nct - 1) / nct;
    jt = (jt + 1) & ~1;  // even: the plane boxes start at table column j0 = ct*jt, which TMA wants 16-byte aligned
}

bool fused_step_usable(const ModelDev &v, int B, int ldb, const double *x0, const double *f, const double *work) {
    if (fs_env_int("NKB_FUSED", 1) == 0) return false;
    if (v.column_model == 1 && v.ny == 1) return false;
    if (v.kind != NKB_MOD_LINEAR && v.kind != NKB_MOD_FORCED_FILE) return false;
    if ((v.nz + FS_KC - 1) / FS_KC * 16 > 512) return false;  // TMEM: 4 columns per level
    if (B < fs_env_int("NKB_FUSED_MIN_B", 8) || (ldb % 2) != 0) return false;
    if (((uintptr_t)x0 | (uintptr_t)f | (uintptr_t)work) & 15) return false;
    return fs_encode_fn() != nullptr;
}

int fused_encode_state_maps(const ModelDev &v, int B, int ldb, const double *buf, CUtensorMap *in, CUtensorMap *out) {
    int nct, jt;
    fs_col_tiles(v.ny, nct, jt);
    const cuuint64_t dims[4] = {(cuuint64_t)B, (cuuint64_t)v.ny, (cuuint64_t)v.nz, (cuuint64_t)v.T};
    const cuuint64_t strides[3] = {(cuuint64_t)ldb * 8, (cuuint64_t)v.ny * ldb * 8, (cuuint64_t)v.nz * v.ny * ldb * 8};
    const cuuint32_t box_in[4] = {FS_MEM, FS_UCOLS, FS_KC, 1};
    const cuuint32_t box_out[4] = {FS_MEM, (cuuint32_t)jt, FS_KC, 1};
    if (in && fs_encode(in, buf, 4, dims, strides, box_in, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (out && fs_encode(out, buf, 4, dims, strides, box_out, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    return 0;
}

int fused_encode_plane_map(int nz, int ny, int nyp, size_t nplanes, const double *buf, CUtensorMap *map) {
    // plane tables carry one zero column on the left (table column = j + 1): the box of a tile starts
    // at column j0 - 1, and TMA faults ("illegal instruction") on a box whose innermost start
    // address is not 16-byte aligned (odd float64 coordinate), measured on B200
    const cuuint64_t dims[3] = {(cuuint64_t)(ny + 1), (cuuint64_t)nz, (cuuint64_t)nplanes};
    const cuuint64_t strides[2] = {(cuuint64_t)nyp * 8, (cuuint64_t)nz * nyp * 8};
    const cuuint32_t box[3] = {FS_COLS, FS_KC, 1};
    return fs_encode(map, buf, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

struct FusedLaunch {
    const ModelDev *v;
    int B, n_steps;
    const CUtensorMap *uin, *uout, *est, *ftab, *src;
};

template <int KIND, bool HAS_E>
static int fs_launch_t(const StepArgs &a, const StepMaps &maps, int grid, cudaStream_t st) {
    auto kern = step_fused_kernel<KIND, HAS_E>;
    static bool attr_set = false;
    if (!attr_set) {
        NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
        attr_set = true;
    }
    kern<<<grid, FS_THREADS, FS_SMEM, st>>>(a, maps);
    count_launch();
    return 0;
}

int launch_step_fused(const ModelDev &v, int B, int n_steps, int step, double h, const double *aff1,
                      const double *aff2, const CUtensorMap &uin, const CUtensorMap &uout, const CUtensorMap *est,
                      const CUtensorMap &ftab, const CUtensorMap *src, cudaStream_t st) {
    StepArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nz = v.nz; a.ny = v.ny; a.B = B; a.T = v.T; a.ncls = v.n_classes; a.n_steps = n_steps;
    fs_col_tiles(v.ny, a.nct, a.jt);
    a.nmb = (B + FS_MEM - 1) / FS_MEM;
    a.ntiles = a.nmb * a.nct * v.T;
    a.step = step;
    a.dbg = fs_env_int("NKB_FUSED_DBG", 0);
    for (int t = 0; t < NKB_MAX_TRACERS; ++t) { a.class_of[t] = v.class_of[t]; a.src_const[t] = v.src_const[t]; }
    a.sink_thres_r = v.sink_thres > 0.0 ? 1.0 / v.sink_thres : 0.0;
    a.hg = kGamma * h;
    a.a1 = (1.0 - kGamma) / kGamma;
    a.a0 = 1.0 - a.a1;
    a.r = (kDelta - 1.0 + kGamma) / kGamma;
    a.he1 = h * (1.0 - kDelta);
    a.aff1 = aff1; a.aff2 = aff2;
    StepMaps maps;
    std::memset(&maps, 0, sizeof(maps));
    maps.uin = uin; maps.uout = uout; maps.ftab = ftab;
    if (est) maps.est = *est;
    if (src) maps.src = *src;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        NKB_CUDA(cudaGetDevice(&dev));
        NKB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = fs_env_int("NKB_FUSED_GRID", n_sm);
    if (grid > a.ntiles) grid = a.ntiles;
    const bool has_e = (est != nullptr);
    if (v.kind == NKB_MOD_LINEAR) {
        return has_e ? fs_launch_t<NKB_MOD_LINEAR, true>(a, maps, grid, st)
                     : fs_launch_t<NKB_MOD_LINEAR, false>(a, maps, grid, st);
    }
    if (v.kind == NKB_MOD_FORCED_FILE) {
        NKB_REQUIRE(src != nullptr, "launch_step_fused: forcing planes missing");
        return has_e ? fs_launch_t<NKB_MOD_FORCED_FILE, true>(a, maps, grid, st)
                     : fs_launch_t<NKB_MOD_FORCED_FILE, false>(a, maps, grid, st);
    }
    set_error("launch_step_fused: unsupported module kind");
    return 2;
}

int launch_sub_inplace(double *out, const double *x0, size_t n, cudaStream_t st) {
    const size_t n2 = n / 2;  // n is a multiple of ldb (even)
    sub_inplace_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(out, x0, n2);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nkb
