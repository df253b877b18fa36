// Fused IMEX step kernel — both implicit stages of an ARS(2,2,2) time step in one pass over the
// state, and ALL time steps of a model year in ONE persistent cooperative launch.  A state is read
// from HBM once (+ halo) and written once per step instead of 3 reads and 2 writes with one launch
// per stage (nkb_stage_tma.cu); the elimination intermediates never leave the SM.
//
// Replaces, per step and for all members at once, what the reference does inside
// solve_ivp(Radau) for py_driver_2d (nk_ooc/py_driver_2d/model_state.py:94-121):
// TracerModuleState.comp_tend (tracer_module_state.py:98-108: advection.py:51-76,
// horiz_mix.py:48-67, vert_mix.py:24-41, iage.py:22-41 / forced.py:114-154) and the implicit
// solves of its Jacobian (vert_mix.py:140-188).
//
// Work decomposition (B200: 148 SMs, one persistent CTA per SM)
//   tile  = 16 members x 14 interior columns (+1 halo column each side for the stage-1 solution,
//           +2 for the state) x all levels of one tracer;  rows of 16 members = 128 bytes.
//   warp  = 16 columns x 2 members, lane = 2*column + (member & 1), 8 consumer warps (default
//           layout; the other one is 4 warps x 2 members per thread).  Warps are self-contained: the
//           only cross-thread exchange (stage-1 solution of the neighbour columns) is a warp shuffle.
//   sweeps over depth per (time step, tile) item:
//     A (top->bottom)  stage-1 right-hand side (explicit horizontal stencil + sources) and LU
//                      forward elimination;  intermediates y1_k -> TMEM
//     B (bottom->top)  stage-1 back substitution u1_k, exchange with the neighbour columns,
//                      stage-2 right-hand side (the explicit term of stage 1 is recovered from
//                      y1_k + m1_k y1_{k-1} and a re-read of the state, which hits L2) and UL
//                      elimination upwards;  y2_k overwrites y1_k in TMEM
//     C (top->bottom)  stage-2 substitution u2_k -> shared-memory staging -> TMA store
//   C of an item and A of the next item run as ONE pass (chunk c of the TMEM scratch is read by C
//   and overwritten by A) unless the next item needs this one's result: two passes per tile.
//   Tensor memory is used as a per-lane scratchpad (tcgen05.st/ld 32x32b.x16): 256 32-bit columns
//   per thread = 128 levels of float64, two warps per lane quarter; 512 columns x 128 lanes hold the
//   256 (column, member) pairs of a tile.  No tcgen05.mma is issued: nothing here is a contraction.
//   All state and coefficient traffic global->shared goes through TMA (cp.async.bulk.tensor,
//   128-byte swizzle for the state boxes so that lane = column accesses are bank-conflict free)
//   into one mbarrier full/empty ring shared by all sweeps; a producer lane runs ahead of the
//   consumers across sweeps, tiles and time steps, a store lane drains the output staging ring.
//   Time steps are chained inside the launch: the store lane publishes a per-tile step counter
//   (release), the producer lane acquires the counters of the tile and of its two column neighbours
//   before the first load of the next step; the tile -> CTA assignment rotates from step to step so
//   that the left-over tiles of a step do not always land on the same CTAs.

#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "nkb_common.cuh"

namespace nkb {

constexpr int FS_COLS = 16;   // stage-1 columns per tile
constexpr int FS_UCOLS = 18;  // state columns per tile (halo of 2)
constexpr int FS_JT = 14;     // max interior columns per tile
constexpr int FS_MEM = 16;    // members per tile (128-byte rows)
constexpr int FS_NPP = 4;     // coefficient pair planes per ring slot
// consumer warps: MPT members per thread, KC levels per chunk; 256 (column, member) pairs per tile
//   MPT = 1, KC = 8: 8 warps, lane = 2*column + (member & 1)   (two warps per scheduler: default)
//   MPT = 2, KC = 4: 4 warps, lane = column + 16*pair
// either way one chunk of one thread is 16 32-bit TMEM columns (one tcgen05 x16 access)
#ifndef FS_NS1
#define FS_NS1 5  // input ring slots of the 8-warp layout
#endif
#ifndef FS_NO1
#define FS_NO1 2  // output staging slots
#endif
struct FsCfg {
    int kc, ncw, threads, ns, no, ubytes, ppbytes, slot, out, smem;
};
__host__ __device__ constexpr FsCfg fs_cfg(int mpt) {
    const int kc = (mpt == 1) ? 8 : 4;
    const int ncw = 8 / mpt;
    const int ns = (mpt == 1) ? FS_NS1 : 8, no = (mpt == 1) ? FS_NO1 : 3;
    const int ubytes = kc * FS_UCOLS * FS_MEM * 8;  // 1024-byte multiple
    const int ppbytes = kc * FS_COLS * 16;
    const int slot = ubytes + FS_NPP * ppbytes;
    const int out = kc * FS_JT * FS_MEM * 8;
    return {kc, ncw, (ncw + 2) * 32, ns, no, ubytes, ppbytes, slot, out, 1024 + ns * slot + no * out};
}
static_assert(fs_cfg(1).slot % 1024 == 0 && fs_cfg(1).out % 1024 == 0, "swizzled boxes need 1024-byte aligned bases");
static_assert(fs_cfg(2).slot % 1024 == 0 && fs_cfg(2).out % 1024 == 0, "swizzled boxes need 1024-byte aligned bases");

struct StepArgs {
    int nz, ny, B, T, ncls, n_steps;
    int nct, jt, nmb, ntiles;  // column tiles, interior columns per tile, member blocks, total
    int step0, step1;          // this launch integrates the time steps [step0, step1)
    int rot;                   // tile -> CTA assignment is rotated by rot CTAs per step (load balance)
    int cross_step_fuse;       // sweep C of a step's last tile may share a pass with sweep A of the next step
    int tgroup;                // phosphorus kernel: member-adjacent tiles taken back to back by one CTA
    int class_of[NKB_MAX_TRACERS];
    double src_const[NKB_MAX_TRACERS];
    double sink_thres_r;
    double r, a0r;             // P = r*rhs1 + a0r*u_n (see launch_steps_fused)
    double p3_hs, p3_sigma, p3_rd, p3_rp;  // phosphorus: half saturation, dop fraction, remineralisation rates
    const double *h;           // [n_steps] step sizes
    const double *aff;         // [2*n_steps][ncls][ny]: gamma*h*(affine surface source) of every stage
    int *done;                 // [ntiles] number of time steps completed for the tile (inter-CTA dependencies)
    int *err;                  // set when a dependency wait timed out (host-mapped)
};

// state buffers of a model-year evaluation: x0 (read by step 0 only) and the two buffers the steps
// alternate between (the last step writes f)
struct StepMaps {
    CUtensorMap in_x0, in_f, in_w, out_f, out_w, ctab;
};

// tile id -> (member block, column tile, tracer).  Column tile fastest: the tiles a round of CTAs works on at the
// same time are column NEIGHBOURS of one member block, so the two halo columns a tile shares with each neighbour come
// out of L2 instead of being read from HBM twice (member block fastest: neighbours are nmb ids = almost two rounds
// apart and the halo was re-read: 1.28 x the state per step, ncu) — and the tiles a tile waits for at a step boundary
// are the ones the neighbouring CTAs have just finished.
#ifndef FS_CT_FASTEST
#define FS_CT_FASTEST 1
#endif
__device__ __forceinline__ void fs_tile_split(int tile, int nmb, int nct, int &mb, int &ct, int &tr) {
#if FS_CT_FASTEST
    ct = tile % nct;
    const int r = tile / nct;
    mb = r % nmb;
    tr = r / nmb;
#else
    mb = tile % nmb;
    const int r = tile / nmb;
    ct = r % nct;
    tr = r / nct;
#endif
}
// distance in tile ids between column neighbours
__host__ __device__ __forceinline__ int fs_nbr(int nmb) { return FS_CT_FASTEST ? 1 : nmb; }

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fs_smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void fs_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fs_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the waiting warp sleeps in hardware instead of burning issue
// slots of the consumer warps that share its scheduler
template <int HINT_NS>
__device__ __forceinline__ void fs_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(HINT_NS)
            : "memory");
    } while (!ok);
}
// plain try_wait + a real sleep between polls (the hinted form above wakes on every barrier event of the CTA:
// in the phosphorus kernel the producer and store lanes were executing 20 % of all instructions; in the
// single-tracer kernel the two forms measure the same)
template <int SLEEP_NS>
__device__ __forceinline__ void fs_mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(SLEEP_NS);
    }
}
#ifndef FS_DEP_SLEEP
#define FS_DEP_SLEEP 200  // ns between polls of a neighbour tile's step counter
#endif
// spin until *flag >= want (acquire); gives up after 2e10 cycles (~10 s) and raises *err instead of hanging
__device__ __forceinline__ void fs_wait_done(const int *flag, int want, int *err) {
    int v;
    const long long t0 = clock64();
    while (true) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= want) return;
        __nanosleep(FS_DEP_SLEEP);
        if (clock64() - t0 > 20000000000ll) {
            *reinterpret_cast<volatile int *>(err) = 1;
            return;
        }
    }
}
__device__ __forceinline__ void fs_tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                               int c2, int c3, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void fs_tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                               int c2, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void fs_tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2,
                                                int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// ---- tensor memory as a per-thread scratchpad: W 32-bit columns of this thread's TMEM lane --------
template <int W>
struct Raw {
    uint32_t w[W];
};
__device__ __forceinline__ void fs_tmem_st(uint32_t taddr, const Raw<8> &r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :
                 : "r"(taddr), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]), "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]),
                   "r"(r.w[6]), "r"(r.w[7])
                 : "memory");
}
__device__ __forceinline__ void fs_tmem_st(uint32_t taddr, const Raw<16> &r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15, %16};"
        :
        : "r"(taddr), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]), "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]),
          "r"(r.w[7]), "r"(r.w[8]), "r"(r.w[9]), "r"(r.w[10]), "r"(r.w[11]), "r"(r.w[12]), "r"(r.w[13]),
          "r"(r.w[14]), "r"(r.w[15])
        : "memory");
}
// the load is asynchronous: the words are only valid after fs_tmem_wait_ld, which takes them as
// in/out operands so that no use can be scheduled ahead of the wait
__device__ __forceinline__ void fs_tmem_ld(uint32_t taddr, Raw<8> &r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void fs_tmem_ld(uint32_t taddr, Raw<16> &r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
          "=r"(r.w[7]), "=r"(r.w[8]), "=r"(r.w[9]), "=r"(r.w[10]), "=r"(r.w[11]), "=r"(r.w[12]), "=r"(r.w[13]),
          "=r"(r.w[14]), "=r"(r.w[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void fs_tmem_wait_ld(Raw<8> &r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]),
                   "+r"(r.w[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void fs_tmem_wait_ld(Raw<16> &r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]),
                   "+r"(r.w[7]), "+r"(r.w[8]), "+r"(r.w[9]), "+r"(r.w[10]), "+r"(r.w[11]), "+r"(r.w[12]),
                   "+r"(r.w[13]), "+r"(r.w[14]), "+r"(r.w[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void fs_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- MPT members of one (level, column) ---------------------------------------------------------
template <int M>
struct Vd {
    double v[M];
};
template <int M>
__device__ __forceinline__ Vd<M> fs_splat(double a) {
    Vd<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i) r.v[i] = a;
    return r;
}
template <int M>
__device__ __forceinline__ Vd<M> fs_fma(double a, Vd<M> x, Vd<M> y) {
    Vd<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i) r.v[i] = fma(a, x.v[i], y.v[i]);
    return r;
}
template <int M>
__device__ __forceinline__ Vd<M> fs_mul(double a, Vd<M> x) {
    Vd<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i) r.v[i] = a * x.v[i];
    return r;
}
template <int M>
__device__ __forceinline__ Vd<M> fs_add(Vd<M> x, Vd<M> y) {
    Vd<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i) r.v[i] = x.v[i] + y.v[i];
    return r;
}
template <int M>
__device__ __forceinline__ Vd<M> fs_sub(Vd<M> x, Vd<M> y) {
    Vd<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i) r.v[i] = x.v[i] - y.v[i];
    return r;
}
__device__ __forceinline__ Vd<2> fs_ld(const unsigned char *p, Vd<2> *) {
    const double2 t = *reinterpret_cast<const double2 *>(p);
    return {{t.x, t.y}};
}
__device__ __forceinline__ Vd<1> fs_ld(const unsigned char *p, Vd<1> *) {
    return {{*reinterpret_cast<const double *>(p)}};
}
__device__ __forceinline__ void fs_st(unsigned char *p, Vd<2> v) {
    *reinterpret_cast<double2 *>(p) = make_double2(v.v[0], v.v[1]);
}
__device__ __forceinline__ void fs_st(unsigned char *p, Vd<1> v) { *reinterpret_cast<double *>(p) = v.v[0]; }
// value of the neighbour column: lane - DELTA / lane + DELTA (lanes at the tile edge get their own)
template <int M, int DELTA, int WIDTH>
__device__ __forceinline__ Vd<M> fs_shfl_up(Vd<M> x) {
    Vd<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i) r.v[i] = __shfl_up_sync(0xffffffffu, x.v[i], DELTA, WIDTH);
    return r;
}
template <int M, int DELTA, int WIDTH>
__device__ __forceinline__ Vd<M> fs_shfl_down(Vd<M> x) {
    Vd<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i) r.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], DELTA, WIDTH);
    return r;
}
template <int M, int KC>
__device__ __forceinline__ void fs_pack(const Vd<M> (&v)[KC], Raw<KC * 2 * M> &r) {
#pragma unroll
    for (int q = 0; q < KC; ++q)
#pragma unroll
        for (int i = 0; i < M; ++i) {
            r.w[(q * M + i) * 2] = (uint32_t)__double2loint(v[q].v[i]);
            r.w[(q * M + i) * 2 + 1] = (uint32_t)__double2hiint(v[q].v[i]);
        }
}
template <int M, int KC>
__device__ __forceinline__ void fs_unpack(const Raw<KC * 2 * M> &r, Vd<M> (&v)[KC]) {
#pragma unroll
    for (int q = 0; q < KC; ++q)
#pragma unroll
        for (int i = 0; i < M; ++i)
            v[q].v[i] = __hiloint2double((int)r.w[(q * M + i) * 2 + 1], (int)r.w[(q * M + i) * 2]);
}

// explicit source of one tracer times the stage weight w (w = gamma h or h (1 - delta)):
// LINEAR: w * constant (iage.py:39), passed in ws.  FORCED_FILE: fw = w * forcing record with the
// sink_thres limiter of forced.py:141-151 (sms scaled by c/thres where sms < 0 and 0 < c < thres;
// thr_r = 1/thres, or 0 when the limiter is off so that q = 0 fails the test).  The three
// comparisons are chained into one predicate (DSETP .AND) and a single select.
template <int KIND, int M>
__device__ __forceinline__ Vd<M> fs_source(double ws, double thr_r, double fw, Vd<M> c) {
    Vd<M> r;
    if constexpr (KIND == NKB_MOD_LINEAR) {
#pragma unroll
        for (int i = 0; i < M; ++i) r.v[i] = ws;
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) {
#ifdef FS_FP_LIMITER
            asm("{\n\t.reg .pred pl, p1, p2;\n\t.reg .f64 q, fq;\n\t"
                "setp.lt.f64 pl, %2, 0d0000000000000000;\n\t"
                "mul.f64 q, %3, %1;\n\t"
                "setp.gt.and.f64 p1, q, 0d0000000000000000, pl;\n\t"
                "setp.lt.and.f64 p2, q, 0d3FF0000000000000, p1;\n\t"
                "mul.f64 fq, %2, q;\n\t"
                "selp.f64 %0, fq, %2, p2;\n\t}"
                : "=d"(r.v[i])
                : "d"(c.v[i]), "d"(fw), "d"(thr_r));
#else
            // the same three comparisons on the bit patterns (integer pipe instead of three DSETP on the FP64 pipe,
            // which this kernel keeps busiest after the LSU): fw < 0 is its sign bit (-0.0 scales to -0.0 either
            // way), 0 < q < 1 is 1 <= high word <= 0x3FEFFFFF for a non-negative double.  Differs from the
            // floating-point test only for subnormal q (0 < c < 1e-309 thres), which is then left unscaled.
            asm("{\n\t.reg .pred p1, p2;\n\t.reg .f64 q, fq;\n\t.reg .b32 qlo, qhi, flo, fhi, t;\n\t"
                "mul.f64 q, %3, %1;\n\t"
                "mul.f64 fq, %2, q;\n\t"
                "mov.b64 {qlo, qhi}, q;\n\t"
                "mov.b64 {flo, fhi}, %2;\n\t"
                "sub.u32 t, qhi, 1;\n\t"
                "setp.lt.u32 p1, t, 0x3FEFFFFF;\n\t"
                "setp.lt.and.s32 p2, fhi, 0, p1;\n\t"
                "selp.f64 %0, fq, %2, p2;\n\t}"
                : "=d"(r.v[i])
                : "d"(c.v[i]), "d"(fw), "d"(thr_r));
#endif
        }
    }
    return r;
}

constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// pair planes of a ring slot, [KC][FS_COLS] pairs each (see step_ctab_kernel in nkb_tables.cu):
//   sweep A: 0 {aL,aC}  1 {aR,m1}  2 {fA,-}          sweep C: 3 {ib2,g2}  (C may share a slot with A)
//   sweep B: 0 {bL,bC}  1 {bR,m1}  2 {g1,ib1} 3 {m2,fB}   (the back substitution needs plane 2 only)
#ifndef FS_HINT_A
#define FS_HINT_A kEvictNormal
#endif
#ifndef FS_HINT_B
#define FS_HINT_B kEvictFirst
#endif
#ifndef FS_CWAIT
#define FS_CWAIT 20
#endif
template <int KIND, int MPT>
__global__ void __launch_bounds__(fs_cfg(MPT).threads, 1) step_fused_kernel(const StepArgs p,
                                                                             const __grid_constant__ StepMaps maps) {
    constexpr FsCfg C = fs_cfg(MPT);
    constexpr int KC = C.kc, NCW = C.ncw, NS = C.ns, NO = C.no;
    constexpr int W = KC * 2 * MPT;  // TMEM columns per chunk and thread (16)
    constexpr bool FRC = (KIND == NKB_MOD_FORCED_FILE);
    using V = Vd<MPT>;
    // dynamic shared memory is the only shared allocation of this kernel: its window offset is a
    // multiple of 1024 (checked), which the 128-byte-swizzled boxes rely on.  Offsets from the
    // symbol keep the address space visible to the compiler (LDS/STS with immediate offsets).
    extern __shared__ __align__(1024) unsigned char fs_smem[];
    const uint32_t smem0 = fs_smem_u32(fs_smem);
    if (smem0 & 1023u) __trap();
    const uint32_t bar_full = smem0, bar_empty = smem0 + 8 * NS, bar_ofull = smem0 + 16 * NS,
                   bar_oempty = smem0 + 16 * NS + 8 * NO;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(fs_smem + 960);
    unsigned char *ring = fs_smem + 1024;
    unsigned char *oring = ring + NS * C.slot;
    const uint32_t ring_a = smem0 + 1024, oring_a = ring_a + NS * C.slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            fs_mbar_init(bar_full + 8 * s, 1);
            fs_mbar_init(bar_empty + 8 * s, NCW);
        }
        for (int s = 0; s < NO; ++s) {
            fs_mbar_init(bar_ofull + 8 * s, NCW);
            fs_mbar_init(bar_oempty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         fs_smem_u32(tmem_holder))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_holder);

    const int nz = p.nz, ny = p.ny;
    const int nchunk = (nz + KC - 1) / KC;
    constexpr int NPA = FRC ? 3 : 2;

    // ---- the (step, tile) items of this CTA, in processing order (identical in the three roles) ----
    // item i+1 is "independent" of item i when it does not need item i's result: then the last sweep
    // of item i (C, output heavy) and the first sweep of item i+1 (A, input heavy) — both top-down —
    // run as ONE pass over depth: C reads chunk c of the TMEM scratch, A overwrites it.
    struct Item {
        int n, tile;
    };
    auto item_first = [&](int n) { return (int)((blockIdx.x + (unsigned)n * (unsigned)p.rot) % gridDim.x); };
    auto item_valid = [&](const Item &it) { return it.n < p.step1; };
    auto item_next = [&](const Item &it) {
        Item nx = {it.n, it.tile + (int)gridDim.x};
        if (nx.tile >= p.ntiles) {
            nx.n = it.n + 1;
            nx.tile = item_first(nx.n);
        }
        return nx;
    };
    auto item_depends = [&](const Item &nx, const Item &it) {
        if (nx.n == it.n) return false;  // tiles of one step are independent
        // Fusing across a step boundary makes the completion of a CTA's last tile of step n wait for the
        // step-n neighbours of its first tile of step n+1.  That is free of cycles only when those
        // neighbours cannot themselves be last-round tiles (ntiles > 2*grid + nmb, decided by the host).
        if (!p.cross_step_fuse) return true;
        const int d = nx.tile - it.tile;
        // neighbours in the column direction are nmb apart (a tracer boundary only makes this conservative)
        return d == 0 || d == fs_nbr(p.nmb) || d == -fs_nbr(p.nmb);
    };
    Item it0 = {p.step0, item_first(p.step0)};
    if (it0.tile >= p.ntiles) it0.n = p.step1;  // (grid <= ntiles: does not happen)

    if (warp == NCW) {
        // ===== producer: one lane issues every TMA load of this CTA, in consumption order =====
        if (lane == 0) {
            uint32_t g = 0;
            struct TileP {
                const CUtensorMap *uin;
                int m0, j0, tr, zt, ct;
            };
            auto tile_p = [&](const Item &it) {
                TileP t;
                const bool to_f = (((p.n_steps - 1 - it.n) & 1) == 0);  // this step writes f (else w)
                t.uin = (it.n == 0) ? &maps.in_x0 : (to_f ? &maps.in_w : &maps.in_f);
                int mb;
                fs_tile_split(it.tile, p.nmb, p.nct, mb, t.ct, t.tr);
                t.m0 = mb * FS_MEM;
                t.j0 = t.ct * p.jt;
                t.zt = (it.n * p.ncls + p.class_of[t.tr]) * 8;
                return t;
            };
            // the tile and its two column neighbours must have completed step n - 1: their output is
            // this step's input (halo included), and this step's output buffer is what they read then
            auto dep_wait = [&](const Item &it, const TileP &t) {
                if (it.n > 0) {
                    fs_wait_done(p.done + it.tile, it.n, p.err);
                    if (t.ct > 0) fs_wait_done(p.done + it.tile - fs_nbr(p.nmb), it.n, p.err);
                    if (t.ct + 1 < p.nct) fs_wait_done(p.done + it.tile + fs_nbr(p.nmb), it.n, p.err);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
            };
            // sweep: 0 = A (tile ta), 1 = B (ta), 2 = C (tc), 3 = C (tc) fused with A (ta)
            auto issue = [&](int sweep, const TileP &ta, const TileP &tc) {
                for (int cc = 0; cc < nchunk; ++cc) {
                    const int c = (sweep == 1) ? nchunk - 1 - cc : cc;
                    const int k0 = c * KC;
                    const uint32_t s = g % NS, ph = (g / NS) & 1;
                    fs_mbar_wait<2000>(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t sb = ring_a + s * C.slot;
                    const uint32_t pl = sb + C.ubytes;
                    const uint32_t fb = bar_full + 8 * s;
                    if (sweep == 0 || sweep == 3) {
                        fs_mbar_expect_tx(fb, C.ubytes + (NPA + (sweep == 3 ? 1 : 0)) * C.ppbytes);
                        fs_tma_load_4d(sb, ta.uin, fb, ta.m0, ta.j0 - 2, k0, ta.tr, FS_HINT_A);
#pragma unroll
                        for (int q = 0; q < NPA; ++q)
                            fs_tma_load_3d(pl + q * C.ppbytes, &maps.ctab, fb, 2 * ta.j0, k0, ta.zt + q, kEvictLast);
                        if (sweep == 3)
                            fs_tma_load_3d(pl + 3 * C.ppbytes, &maps.ctab, fb, 2 * tc.j0, k0, tc.zt + 7, kEvictLast);
                    } else if (sweep == 1) {
                        fs_mbar_expect_tx(fb, C.ubytes + 4 * C.ppbytes);
                        fs_tma_load_4d(sb, ta.uin, fb, ta.m0, ta.j0 - 2, k0, ta.tr, FS_HINT_B);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            fs_tma_load_3d(pl + q * C.ppbytes, &maps.ctab, fb, 2 * ta.j0, k0, ta.zt + 3 + q, kEvictLast);
                    } else {
                        fs_mbar_expect_tx(fb, C.ppbytes);
                        fs_tma_load_3d(pl + 3 * C.ppbytes, &maps.ctab, fb, 2 * tc.j0, k0, tc.zt + 7, kEvictLast);
                    }
                    ++g;
                }
            };
            Item it = it0;
            if (item_valid(it)) {
                TileP t = tile_p(it);
                dep_wait(it, t);
                issue(0, t, t);
                while (true) {
                    issue(1, t, t);
                    const Item nx = item_next(it);
                    if (!item_valid(nx)) {
                        issue(2, t, t);
                        break;
                    }
                    const TileP tn = tile_p(nx);
                    if (!item_depends(nx, it)) {
                        dep_wait(nx, tn);
                        issue(3, tn, t);
                    } else {
                        issue(2, t, t);
                        dep_wait(nx, tn);
                        issue(0, tn, tn);
                    }
                    it = nx;
                    t = tn;
                }
            }
        }
    } else if (warp == NCW + 1) {
        // ===== store warp: drains the output staging ring with TMA stores =====
        if (lane == 0) {
            uint32_t go = 0;
            for (Item it = it0; item_valid(it); it = item_next(it)) {
                const bool to_f = (((p.n_steps - 1 - it.n) & 1) == 0);
                const CUtensorMap *uout = to_f ? &maps.out_f : &maps.out_w;
                int mb, ct, tr;
                fs_tile_split(it.tile, p.nmb, p.nct, mb, ct, tr);
                for (int c = 0; c < nchunk; ++c) {
                    const uint32_t s = go % NO, ph = (go / NO) & 1;
                    fs_mbar_wait<2000>(bar_ofull + 8 * s, ph);
                    fs_tma_store_4d(uout, oring_a + s * C.out, mb * FS_MEM, ct * p.jt, c * KC, tr);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    fs_mbar_arrive(bar_oempty + 8 * s);
                    ++go;
                }
                // publish: the tile has completed step n (its stores are performed)
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.done + it.tile), "r"(it.n + 1) : "memory");
            }
        }
    } else {
        // ===== consumers =====
        constexpr int DELTA = (MPT == 2) ? 1 : 2;  // lane distance of the neighbour column
        constexpr int WIDTH = (MPT == 2) ? 16 : 32;
        constexpr int PP = KC * FS_COLS;           // pairs per plane
        const int col = (MPT == 2) ? (lane & 15) : (lane >> 1);
        const int c16 = (MPT == 2) ? 2 * warp + (lane >> 4) : warp;  // 16-byte chunk within the 128-byte row
        const int sub = (MPT == 2) ? 0 : 8 * (lane & 1);             // byte within the chunk
        // TMEM: lane quarter warp % 4; with 8 warps the upper four use columns 256..511
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
        int offU[KC][3], offO[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int r = q * FS_UCOLS + col + d;
                offU[q][d] = 128 * r + 16 * (c16 ^ (r & 7)) + sub;
            }
            const int ro = q * p.jt + col - 1;
            offO[q] = 128 * ro + 16 * (c16 ^ (ro & 7)) + sub;
        }
        const bool interior = (col >= 1 && col <= p.jt);
        const double thr_r = p.sink_thres_r;
        uint32_t g = 0, go = 0;

        struct TileC {  // per-thread parameters of a (step, tile) item
            double aff1, aff2, ws1, ws2;
        };
        auto tile_c = [&](const Item &it) {
            TileC t;
            const double hstep = __ldg(p.h + it.n);
            const double *aff_n = p.aff + (size_t)(2 * it.n) * p.ncls * ny;
            int mbu, ct, tr;
            fs_tile_split(it.tile, p.nmb, p.nct, mbu, ct, tr);
            const int j = ct * p.jt - 1 + col;
            const int cls = p.class_of[tr];
            t.aff1 = t.aff2 = 0.0;
            if (j >= 0 && j < ny) {
                t.aff1 = __ldg(aff_n + (size_t)cls * ny + j);
                t.aff2 = __ldg(aff_n + (size_t)(p.ncls + cls) * ny + j);
            }
            t.ws1 = kGamma * hstep * p.src_const[tr];
            t.ws2 = hstep * (1.0 - kDelta) * p.src_const[tr];
            return t;
        };

        // stage-1 right-hand side + LU forward elimination of one chunk (top -> bottom); y1 -> raw
        auto chunk_a = [&](const TileC &t, const unsigned char *sb, int c, V &yprev, Raw<W> &raw) {
            const double2 *pl = reinterpret_cast<const double2 *>(sb + C.ubytes) + col;
            V yb[KC];
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                const V cv = fs_ld(sb + offU[q][1], (V *)nullptr);
                const V cl = fs_ld(sb + offU[q][0], (V *)nullptr);
                const V cr = fs_ld(sb + offU[q][2], (V *)nullptr);
                const double2 lc = pl[q * FS_COLS], rm = pl[PP + q * FS_COLS];
                double fw = 0.0;
                if constexpr (FRC) fw = pl[2 * PP + q * FS_COLS].x;
                const V sv = fs_source<KIND, MPT>(t.ws1, thr_r, fw, cv);
                V rhs = fs_fma(lc.x, cl, fs_fma(rm.x, cr, fs_fma(lc.y, cv, sv)));
                if (c == 0 && q == 0) rhs = fs_add(rhs, fs_splat<MPT>(t.aff1));
                yprev = fs_fma(-rm.y, yprev, rhs);
                yb[q] = yprev;
            }
            fs_pack<MPT, KC>(yb, raw);
        };
        // stage-2 substitution of one chunk (top -> bottom) into the output staging slot
        auto chunk_c = [&](const unsigned char *sb, unsigned char *ob, const Raw<W> &cur, V &u2p) {
            const double2 *pl = reinterpret_cast<const double2 *>(sb + C.ubytes) + col;
            V ycur[KC], u2[KC];
            fs_unpack<MPT, KC>(cur, ycur);
            // all coefficient loads first: the compiler cannot move a shared load across the staging
            // stores below (it cannot prove the two ring slots disjoint)
            double2 ig[KC];
#pragma unroll
            for (int q = 0; q < KC; ++q) ig[q] = pl[3 * PP + q * FS_COLS];
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                u2p = fs_fma(-ig[q].y, u2p, fs_mul(ig[q].x, ycur[q]));
                u2[q] = u2p;
            }
            if (interior) {
#pragma unroll
                for (int q = 0; q < KC; ++q) fs_st(ob + offO[q], u2[q]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        };

        // ---- sweep A alone (first item of the launch, or when the item needs its predecessor's result) ----
        auto sweep_a = [&](const TileC &t) {
            V yprev = fs_splat<MPT>(0.0);
            for (int c = 0; c < nchunk; ++c) {
                const uint32_t s = g % NS, ph = (g / NS) & 1;
                fs_mbar_wait<FS_CWAIT>(bar_full + 8 * s, ph);
                Raw<W> raw;
                chunk_a(t, ring + s * C.slot, c, yprev, raw);
                __syncwarp();
                if (lane == 0) fs_mbar_arrive(bar_empty + 8 * s);
                fs_tmem_st(taddr + c * W, raw);
                ++g;
            }
            fs_tmem_wait_st();
        };

        // ---- sweep B: stage-1 back substitution + stage-2 rhs + UL elimination, bottom -> top ----
        auto sweep_b = [&](const TileC &t) {
            Raw<W> ra, rb;  // chunk c and chunk c-1 (prefetched), alternating roles: no register copies
            V u1n = fs_splat<MPT>(0.0), y2n = fs_splat<MPT>(0.0);
            auto chunk_b = [&](Raw<W> &cur, Raw<W> &nxt, int c) {
                if (c > 0) fs_tmem_ld(taddr + (c - 1) * W, nxt);
                const uint32_t s = g % NS, ph = (g / NS) & 1;
                fs_mbar_wait<FS_CWAIT>(bar_full + 8 * s, ph);
                V ycur[KC], y1top;
                fs_unpack<MPT, KC>(cur, ycur);
                if (c > 0) {
                    fs_tmem_wait_ld(nxt);
#pragma unroll
                    for (int i = 0; i < MPT; ++i)
                        y1top.v[i] = __hiloint2double((int)nxt.w[((KC - 1) * MPT + i) * 2 + 1],
                                                      (int)nxt.w[((KC - 1) * MPT + i) * 2]);
                } else {
                    y1top = fs_splat<MPT>(0.0);
                }
                const unsigned char *sb = ring + s * C.slot;
                const double2 *pl = reinterpret_cast<const double2 *>(sb + C.ubytes) + col;
                V yb[KC];
#pragma unroll
                for (int q = KC - 1; q >= 0; --q) {
                    const V y1 = ycur[q];
                    const V y1m = (q > 0) ? ycur[q > 0 ? q - 1 : 0] : y1top;
                    const double2 lc = pl[q * FS_COLS], ri = pl[PP + q * FS_COLS], gm = pl[2 * PP + q * FS_COLS],
                                  mf = pl[3 * PP + q * FS_COLS];
                    const V u1 = fs_fma(-gm.x, u1n, fs_mul(gm.y, y1));
                    u1n = u1;
                    V rhs1 = fs_fma(ri.y, y1m, y1);  // = u_n + gamma h E(u_n) (+ aff1 at the surface)
                    if (c == 0 && q == 0) rhs1 = fs_sub(rhs1, fs_splat<MPT>(t.aff1));
                    const V un = fs_ld(sb + offU[q][1], (V *)nullptr);
                    // a0 u_n + h (delta - 1 + gamma) E(u_n) + he1 * source(u1)
                    V pp;
                    if constexpr (FRC) {
                        pp = fs_add(fs_fma(p.r, rhs1, fs_mul(p.a0r, un)), fs_source<KIND, MPT>(t.ws2, thr_r, mf.y, u1));
                    } else {
                        pp = fs_fma(p.r, rhs1, fs_fma(p.a0r, un, fs_splat<MPT>(t.ws2)));
                    }
                    const V ul = fs_shfl_up<MPT, DELTA, WIDTH>(u1), ur = fs_shfl_down<MPT, DELTA, WIDTH>(u1);
                    V rhs2 = fs_fma(lc.x, ul, fs_fma(ri.x, ur, fs_fma(lc.y, u1, pp)));
                    if (c == 0 && q == 0) rhs2 = fs_add(rhs2, fs_splat<MPT>(t.aff2));
                    y2n = fs_fma(-mf.x, y2n, rhs2);
                    yb[q] = y2n;
                }
                __syncwarp();
                if (lane == 0) fs_mbar_arrive(bar_empty + 8 * s);
                Raw<W> raw;
                fs_pack<MPT, KC>(yb, raw);
                fs_tmem_st(taddr + c * W, raw);
                ++g;
            };
            fs_tmem_ld(taddr + (nchunk - 1) * W, ra);
            fs_tmem_wait_ld(ra);
            int c = nchunk - 1;
            for (; c >= 1; c -= 2) {
                chunk_b(ra, rb, c);
                chunk_b(rb, ra, c - 1);
            }
            if (c == 0) chunk_b(ra, rb, 0);
            fs_tmem_wait_st();
        };

        // ---- sweep C (stage-2 substitution, staged TMA store), optionally fused with sweep A of the next
        //      item: both run top -> bottom; A's y1 of chunk c replaces the y2 that C has just consumed ----
        auto sweep_c = [&](bool with_a, const TileC &ta) {
            Raw<W> ra, rb;
            V u2p = fs_splat<MPT>(0.0), yprev = fs_splat<MPT>(0.0);
            auto chunk_ca = [&](Raw<W> &cur, Raw<W> &nxt, int c) {
                if (c + 1 < nchunk) fs_tmem_ld(taddr + (c + 1) * W, nxt);
                const uint32_t s = g % NS, ph = (g / NS) & 1;
                const uint32_t so = go % NO, pho = (go / NO) & 1;
                fs_mbar_wait<FS_CWAIT>(bar_full + 8 * s, ph);
                fs_mbar_wait<FS_CWAIT>(bar_oempty + 8 * so, pho ^ 1);
                const unsigned char *sb = ring + s * C.slot;
                chunk_c(sb, oring + so * C.out, cur, u2p);
                Raw<W> raw;
                if (with_a) chunk_a(ta, sb, c, yprev, raw);
                __syncwarp();
                if (lane == 0) {
                    fs_mbar_arrive(bar_empty + 8 * s);
                    fs_mbar_arrive(bar_ofull + 8 * so);
                }
                if (with_a) fs_tmem_st(taddr + c * W, raw);
                if (c + 1 < nchunk) fs_tmem_wait_ld(nxt);
                ++g;
                ++go;
            };
            fs_tmem_ld(taddr, ra);
            fs_tmem_wait_ld(ra);
            int c = 0;
            for (; c + 1 < nchunk; c += 2) {
                chunk_ca(ra, rb, c);
                chunk_ca(rb, ra, c + 1);
            }
            if (c < nchunk) chunk_ca(ra, rb, c);
            if (with_a) fs_tmem_wait_st();
        };

        // A batch of at most 14 members (one member block: a single state staged into 4 lanes, the Newton iterate and
        // every Krylov product) fills only some of the eight warps.  The others keep the barrier protocol going —
        // same waits and arrives in the same order — and leave their schedulers to the warps that work.
        if (p.nmb == 1 && ((MPT == 2) ? 4 * warp : 2 * warp) >= p.B) {
            auto pass = [&](bool with_out) {
                for (int c = 0; c < nchunk; ++c) {
                    const uint32_t s = g % NS, ph = (g / NS) & 1;
                    fs_mbar_wait<FS_CWAIT>(bar_full + 8 * s, ph);
                    if (with_out) {
                        const uint32_t so = go % NO, pho = (go / NO) & 1;
                        fs_mbar_wait<FS_CWAIT>(bar_oempty + 8 * so, pho ^ 1);
                        if (lane == 0) fs_mbar_arrive(bar_ofull + 8 * so);
                        ++go;
                    }
                    if (lane == 0) fs_mbar_arrive(bar_empty + 8 * s);
                    ++g;
                }
            };
            Item it = it0;
            if (item_valid(it)) {
                pass(false);  // sweep A
                while (true) {
                    pass(false);  // sweep B
                    const Item nx = item_next(it);
                    if (!item_valid(nx)) {
                        pass(true);  // sweep C
                        break;
                    }
                    pass(true);  // sweep C, alone or sharing the pass with sweep A of the next item
                    if (item_depends(nx, it)) pass(false);  // sweep A on its own
                    it = nx;
                }
            }
        } else {
        Item it = it0;
        if (item_valid(it)) {
            TileC t = tile_c(it);
            sweep_a(t);
            while (true) {
                sweep_b(t);
                const Item nx = item_next(it);
                if (!item_valid(nx)) {
                    sweep_c(false, t);
                    break;
                }
                const TileC tn = tile_c(nx);
                if (!item_depends(nx, it)) {
                    sweep_c(true, tn);
                } else {
                    sweep_c(false, t);
                    sweep_a(tn);
                }
                it = nx;
                t = tn;
            }
        }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- phosphorus (py_driver_2d/phosphorus.py:58-95): three coupled tracers per tile ------------------
// The explicit sources couple po4, dop and pop cell by cell, so the three tracers of a (column, member)
// pair must be resident together and the tensor-memory scratch holds 3 x 125 levels per pair: a tile is
// 4 members x 14 interior columns x all levels x 3 tracers, six consumer warps = 3 tracers x 2 member
// pairs, lane = 2*column + (member & 1) as above.  Rows of 4 members are only 32 bytes, and TMA
// issues one request per contiguous run of a box: in the member-fastest layout a state box would be
// 144 runs of 32 bytes, which made the TMA request rate the limit of the kernel.  The model year is
// therefore integrated in a member-block-major copy of the state,
//     x_tm[tracer][level][member block][member pair][column][2 members]
// (converted once per evaluation, p3_to_tm_kernel / p3_from_tm_sub_kernel), in which the 18 columns
// of a tile, level and member pair are ONE contiguous run of 288 bytes (16 requests per box) and the
// 16 columns x 2 members that a warp reads are 256 contiguous bytes of shared memory (no bank
// conflicts, no swizzle).
// Sweep A reads the other two tracers' centre values from their state boxes of the same ring slot;
// in sweep B the three warps that share a member pair exchange the stage-1 solution of a chunk through
// a double-buffered shared-memory block and a named barrier before the stage-2 right-hand side.
#ifndef P3_NS_DEF
#define P3_NS_DEF 5
#endif
#ifndef P3_NO_DEF
#define P3_NO_DEF 2  // output staging slots (2 measured +1.6 % against 3, 1 is -19 %)
#endif
#ifndef P3_HINT_A
#define P3_HINT_A kEvictNormal
#endif
#ifndef P3_HINT_B
#define P3_HINT_B kEvictFirst
#endif
#ifndef P3_CWAIT
#define P3_CWAIT 20  // suspend-time hint (ns) of the consumer warps' barrier waits
#endif
constexpr int P3_MEM = 4, P3_T = 3, P3_KC = 8, P3_NCW = 6, P3_NS = P3_NS_DEF, P3_NO = P3_NO_DEF, P3_NCLS = 2;
constexpr int P3_UBOX = P3_KC * FS_UCOLS * P3_MEM * 8;            // 4608
constexpr int P3_PP = P3_KC * FS_COLS * 16;                       // bytes of one pair plane (2048)
constexpr int P3_SLOT = P3_T * P3_UBOX + P3_NCLS * 4 * P3_PP;     // 30208
constexpr int P3_OBOX = P3_KC * FS_JT * P3_MEM * 8;               // 3584
constexpr int P3_OUT = P3_T * P3_OBOX;
constexpr int P3_EX = 2 * P3_T * P3_KC * 64 * 8;                  // u1 exchange, two chunks deep
constexpr int P3_SMEM = 1024 + P3_NS * P3_SLOT + P3_NO * P3_OUT + P3_EX;
constexpr int P3_THREADS = (P3_NCW + 2) * 32;
#ifndef P3_PSLEEP_NS
#define P3_PSLEEP_NS 100
#endif
constexpr int P3_PSLEEP = P3_PSLEEP_NS;  // ns between barrier polls of the producer and store lanes
static_assert(P3_UBOX % 128 == 0 && P3_SLOT % 128 == 0 && P3_OBOX % 128 == 0 && P3_OUT % 128 == 0,
              "TMA boxes need 128-byte aligned shared-memory bases");
static_assert(P3_SMEM <= 227 * 1024, "shared memory");

// a / b for b in the normal range (po4 + half saturation) without the IEEE division subroutine and its special-case
// paths.  The FP64 instruction count of the coupled sources is what the two schedulers that carry two consumer warps
// run out of (round 2, scripts/ab_variants.sh on one box: every DFMA less per source evaluation is worth 1.5 - 2 %), so
// the quotient is as short as its use allows:
//   P3_DIV = 3 (default): reciprocal seed r (MUFU.RCP64H, 2^-23 or better; 2^-20 if only the upper mantissa word is
//       exact), e = 1 - b r and q0 = a r side by side, q = q0 (1 + e): three FP64 operations, relative error e^2 <= 2^-40
//       (1e-12) — seven orders below the time integration error of the uptake term it feeds; exactly 0 for a = 0
//   P3_DIV = 2: the same plus the residual correction with the SEED, q + (a - b q) r: error 2^-40 x 2^-20, within
//       0.51 ulp of the IEEE quotient (checked with exact rational arithmetic), five operations; 3.5 % slower
//   P3_DIV = 0: round 1's two Newton steps on r and the correction (eight operations); 7 % slower
#ifndef P3_DIV
#define P3_DIV 3
#endif
__device__ __forceinline__ double p3_div(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#if P3_DIV >= 2
    const double e = fma(-b, r, 1.0);
    const double q0 = a * r;
    const double q = fma(q0, e, q0);
#if P3_DIV >= 3
    return q;
#else
    return fma(fma(-b, q, a), r, q);
#endif
#else
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
#endif
}

// explicit source of one tracer times the stage weight w (phosphorus.py:75-95) in coefficient form:
//   cu * uptake + kd * dop + kq * pop,  uptake = fw * po4 / (po4 + hs),  fw = w * max_uptake_rate * light
//   po4: cu = -1, kd = +w rd, kq = +w rp;   dop: cu = sigma, kd = -w rd, kq = 0;   pop: cu = 1 - sigma, kd = 0, kq = -w rp
// (the terms that do not need the quotient first, fma(cu fw, quotient, kd dop + kq pop), measured 1 % slower)
__device__ __forceinline__ double p3_source(double fw, double cu, double kd, double kq, double hs, double po4,
                                            double dop, double pop) {
    const double u = fw * p3_div(po4, po4 + hs);
    return fma(kq, pop, fma(kd, dop, cu * u));
}

__global__ void __launch_bounds__(P3_THREADS, 1) step_fused_p3_kernel(const StepArgs p,
                                                                      const __grid_constant__ StepMaps maps) {
    constexpr int KC = P3_KC, NCW = P3_NCW, NS = P3_NS, NO = P3_NO;
    constexpr int W = KC * 2;  // TMEM columns per chunk and thread
    extern __shared__ __align__(1024) unsigned char fs_smem[];
    const uint32_t smem0 = fs_smem_u32(fs_smem);
    if (smem0 & 1023u) __trap();
    const uint32_t bar_full = smem0, bar_empty = smem0 + 8 * NS, bar_ofull = smem0 + 16 * NS,
                   bar_oempty = smem0 + 16 * NS + 8 * NO;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(fs_smem + 960);
    unsigned char *ring = fs_smem + 1024;
    unsigned char *oring = ring + NS * P3_SLOT;
    double *exch = reinterpret_cast<double *>(oring + NO * P3_OUT);
    const uint32_t ring_a = smem0 + 1024, oring_a = ring_a + NS * P3_SLOT;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            fs_mbar_init(bar_full + 8 * s, 1);
            fs_mbar_init(bar_empty + 8 * s, NCW);
        }
        for (int s = 0; s < NO; ++s) {
            fs_mbar_init(bar_ofull + 8 * s, NCW);
            fs_mbar_init(bar_oempty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         fs_smem_u32(tmem_holder))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_holder);

    const int nz = p.nz, ny = p.ny;
    const int nchunk = (nz + KC - 1) / KC;

    struct Item {
        int n, tile;
    };
    auto item_first = [&](int n) { return (int)((blockIdx.x + (unsigned)n * (unsigned)p.rot) % gridDim.x); };
    auto item_valid = [&](const Item &it) { return it.n < p.step1; };
    // a CTA takes groups of p.tgroup member-adjacent tiles back to back: their 32-byte rows share DRAM
    // bursts and L2 lines, so the second tile of a group finds its rows in L2
    auto item_next = [&](const Item &it) {
        Item nx = {it.n, it.tile + 1};
        if (nx.tile % p.tgroup == 0 || nx.tile >= p.ntiles) {
            nx.tile = (it.tile / p.tgroup + (int)gridDim.x) * p.tgroup;
            if (nx.tile >= p.ntiles) {
                nx.n = it.n + 1;
                nx.tile = item_first(nx.n) * p.tgroup;
            }
        }
        return nx;
    };
    auto item_depends = [&](const Item &nx, const Item &it) {
        if (nx.n == it.n) return false;
        if (!p.cross_step_fuse) return true;
        const int d = nx.tile - it.tile;
        return d == 0 || d == fs_nbr(p.nmb) || d == -fs_nbr(p.nmb);
    };
    Item it0 = {p.step0, item_first(p.step0) * p.tgroup};
    if (it0.tile >= p.ntiles) it0.n = p.step1;

    if (warp == NCW) {
        // ===== producer =====
        if (lane == 0) {
            uint32_t g = 0;
            struct TileP {
                const CUtensorMap *uin;
                int mb, j0, ct, zt;
            };
            auto tile_p = [&](const Item &it) {
                TileP t;
                const bool to_f = (((p.n_steps - 1 - it.n) & 1) == 0);
                t.uin = (it.n == 0) ? &maps.in_x0 : (to_f ? &maps.in_w : &maps.in_f);
                int tru;
                fs_tile_split(it.tile, p.nmb, p.nct, t.mb, t.ct, tru);  // (one tile = all three tracers: tru == 0)
                t.j0 = t.ct * p.jt;
                t.zt = it.n * P3_NCLS * 8;
                return t;
            };
            auto dep_wait = [&](const Item &it, const TileP &t) {
                if (it.n > 0) {
                    fs_wait_done(p.done + it.tile, it.n, p.err);
                    if (t.ct > 0) fs_wait_done(p.done + it.tile - fs_nbr(p.nmb), it.n, p.err);
                    if (t.ct + 1 < p.nct) fs_wait_done(p.done + it.tile + fs_nbr(p.nmb), it.n, p.err);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
            };
            auto issue = [&](int sweep, const TileP &ta, const TileP &tc) {
                for (int cc = 0; cc < nchunk; ++cc) {
                    const int c = (sweep == 1) ? nchunk - 1 - cc : cc;
                    const int k0 = c * KC;
                    const uint32_t s = g % NS, ph = (g / NS) & 1;
                    fs_mbar_wait_sleep<P3_PSLEEP>(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t sb = ring_a + s * P3_SLOT;
                    const uint32_t pl = sb + P3_T * P3_UBOX;
                    const uint32_t fb = bar_full + 8 * s;
                    if (sweep == 0 || sweep == 3) {
                        fs_mbar_expect_tx(fb, P3_T * P3_UBOX + P3_NCLS * (3 + (sweep == 3 ? 1 : 0)) * P3_PP);
#pragma unroll
                        for (int t = 0; t < P3_T; ++t)
                            fs_tma_load_4d(sb + t * P3_UBOX, ta.uin, fb, 2 * (ta.j0 - 2), 2 * ta.mb, k0, t, P3_HINT_A);
#pragma unroll
                        for (int cl = 0; cl < P3_NCLS; ++cl) {
#pragma unroll
                            for (int q = 0; q < 3; ++q)
                                fs_tma_load_3d(pl + (cl * 4 + q) * P3_PP, &maps.ctab, fb, 2 * ta.j0, k0, ta.zt + cl * 8 + q,
                                               kEvictLast);
                            if (sweep == 3)
                                fs_tma_load_3d(pl + (cl * 4 + 3) * P3_PP, &maps.ctab, fb, 2 * tc.j0, k0, tc.zt + cl * 8 + 7,
                                               kEvictLast);
                        }
                    } else if (sweep == 1) {
                        fs_mbar_expect_tx(fb, P3_T * P3_UBOX + P3_NCLS * 4 * P3_PP);
#pragma unroll
                        for (int t = 0; t < P3_T; ++t)
                            fs_tma_load_4d(sb + t * P3_UBOX, ta.uin, fb, 2 * (ta.j0 - 2), 2 * ta.mb, k0, t, P3_HINT_B);
#pragma unroll
                        for (int cl = 0; cl < P3_NCLS; ++cl)
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                fs_tma_load_3d(pl + (cl * 4 + q) * P3_PP, &maps.ctab, fb, 2 * ta.j0, k0,
                                               ta.zt + cl * 8 + 3 + q, kEvictLast);
                    } else {
                        fs_mbar_expect_tx(fb, P3_NCLS * P3_PP);
#pragma unroll
                        for (int cl = 0; cl < P3_NCLS; ++cl)
                            fs_tma_load_3d(pl + (cl * 4 + 3) * P3_PP, &maps.ctab, fb, 2 * tc.j0, k0, tc.zt + cl * 8 + 7,
                                           kEvictLast);
                    }
                    ++g;
                }
            };
            Item it = it0;
            if (item_valid(it)) {
                TileP t = tile_p(it);
                dep_wait(it, t);
                issue(0, t, t);
                while (true) {
                    issue(1, t, t);
                    const Item nx = item_next(it);
                    if (!item_valid(nx)) {
                        issue(2, t, t);
                        break;
                    }
                    const TileP tn = tile_p(nx);
                    if (!item_depends(nx, it)) {
                        dep_wait(nx, tn);
                        issue(3, tn, t);
                    } else {
                        issue(2, t, t);
                        dep_wait(nx, tn);
                        issue(0, tn, tn);
                    }
                    it = nx;
                    t = tn;
                }
            }
        }
    } else if (warp == NCW + 1) {
        // ===== store warp =====
        if (lane == 0) {
            uint32_t go = 0;
            for (Item it = it0; item_valid(it); it = item_next(it)) {
                const bool to_f = (((p.n_steps - 1 - it.n) & 1) == 0);
                const CUtensorMap *uout = to_f ? &maps.out_f : &maps.out_w;
                int mb, ct, tru;
                fs_tile_split(it.tile, p.nmb, p.nct, mb, ct, tru);
                for (int c = 0; c < nchunk; ++c) {
                    const uint32_t s = go % NO, ph = (go / NO) & 1;
                    fs_mbar_wait_sleep<P3_PSLEEP>(bar_ofull + 8 * s, ph);
#pragma unroll
                    for (int t = 0; t < P3_T; ++t)
                        fs_tma_store_4d(uout, oring_a + s * P3_OUT + t * P3_OBOX, 2 * ct * p.jt, 2 * mb, c * KC, t);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    fs_mbar_arrive(bar_oempty + 8 * s);
                    ++go;
                }
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.done + it.tile), "r"(it.n + 1) : "memory");
            }
        }
    } else {
        // ===== consumers: warp = (tracer, member pair) =====
        constexpr int PP = KC * FS_COLS;  // pairs per plane
        // (which tracer / member pair shares a scheduler with which makes no difference: three role maps measured
        // within 1 %, profiles/r02_p3_variants.md)
        const int tr = warp >> 1, pr = warp & 1;
        const int cls = p.class_of[tr];
        const int col = lane >> 1;
        const int sub = 8 * (lane & 1);
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
        int offU[KC][3], offO[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                offU[q][d] = 16 * ((q * 2 + pr) * FS_UCOLS + col + d) + sub;
            }
            offO[q] = 16 * ((q * 2 + pr) * p.jt + col - 1) + sub;
        }
        const bool interior = (col >= 1 && col <= p.jt);
        const double hs = p.p3_hs, sg = p.p3_sigma;
        const int exl = pr * 32 + lane;  // this thread's slot in a [tracer][level] row of the exchange block
        uint32_t g = 0, go = 0, gx = 0;
        const double cu = (tr == 0) ? -1.0 : (tr == 1 ? sg : 1.0 - sg);
        const double sd = (tr == 0) ? p.p3_rd : (tr == 1 ? -p.p3_rd : 0.0);
        const double sq = (tr == 0) ? p.p3_rp : (tr == 2 ? -p.p3_rp : 0.0);

        struct TileC {
            double aff1, aff2, kd1, kq1, kd2, kq2;
        };
        auto tile_c = [&](const Item &it) {
            TileC t;
            const double hstep = __ldg(p.h + it.n);
            const double *aff_n = p.aff + (size_t)(2 * it.n) * p.ncls * ny;
            int mbu, ct, tru;
            fs_tile_split(it.tile, p.nmb, p.nct, mbu, ct, tru);
            const int j = ct * p.jt - 1 + col;
            t.aff1 = t.aff2 = 0.0;
            if (j >= 0 && j < ny) {
                t.aff1 = __ldg(aff_n + (size_t)cls * ny + j);
                t.aff2 = __ldg(aff_n + (size_t)(p.ncls + cls) * ny + j);
            }
            const double w1 = kGamma * hstep, w2 = hstep * (1.0 - kDelta);
            t.kd1 = w1 * sd; t.kq1 = w1 * sq;
            t.kd2 = w2 * sd; t.kq2 = w2 * sq;
            return t;
        };

        auto chunk_a = [&](const TileC &t, const unsigned char *sb, int c, double &yprev, Raw<W> &raw) {
            const double2 *pl = reinterpret_cast<const double2 *>(sb + P3_T * P3_UBOX + cls * 4 * P3_PP) + col;
            const unsigned char *ub = sb + tr * P3_UBOX;
            Vd<1> yb[KC];
            // the sources of the chunk first: eight independent evaluations (instruction-level parallelism
            // for the reciprocal sequences), then the recurrence
            double sv[KC];
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                const double po4 = *reinterpret_cast<const double *>(sb + offU[q][1]);
                const double dop = *reinterpret_cast<const double *>(sb + P3_UBOX + offU[q][1]);
                const double pop = *reinterpret_cast<const double *>(sb + 2 * P3_UBOX + offU[q][1]);
                sv[q] = p3_source(pl[2 * PP + q * FS_COLS].x, cu, t.kd1, t.kq1, hs, po4, dop, pop);
            }
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                const double cv = *reinterpret_cast<const double *>(ub + offU[q][1]);
                const double cl = *reinterpret_cast<const double *>(ub + offU[q][0]);
                const double cr = *reinterpret_cast<const double *>(ub + offU[q][2]);
                const double2 lc = pl[q * FS_COLS], rm = pl[PP + q * FS_COLS];
                double rhs = fma(lc.x, cl, fma(rm.x, cr, fma(lc.y, cv, sv[q])));
                if (c == 0 && q == 0) rhs += t.aff1;
                yprev = fma(-rm.y, yprev, rhs);
                yb[q].v[0] = yprev;
            }
            fs_pack<1, KC>(yb, raw);
        };
        auto chunk_c = [&](const unsigned char *sb, unsigned char *ob, const Raw<W> &cur, double &u2p) {
            const double2 *pl = reinterpret_cast<const double2 *>(sb + P3_T * P3_UBOX + cls * 4 * P3_PP) + col;
            Vd<1> ycur[KC];
            double u2[KC];
            fs_unpack<1, KC>(cur, ycur);
            double2 ig[KC];
#pragma unroll
            for (int q = 0; q < KC; ++q) ig[q] = pl[3 * PP + q * FS_COLS];
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                u2p = fma(-ig[q].y, u2p, ig[q].x * ycur[q].v[0]);
                u2[q] = u2p;
            }
            if (interior) {
#pragma unroll
                for (int q = 0; q < KC; ++q) *reinterpret_cast<double *>(ob + tr * P3_OBOX + offO[q]) = u2[q];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        };

        auto sweep_a = [&](const TileC &t) {
            double yprev = 0.0;
            for (int c = 0; c < nchunk; ++c) {
                const uint32_t s = g % NS, ph = (g / NS) & 1;
                fs_mbar_wait<P3_CWAIT>(bar_full + 8 * s, ph);
                Raw<W> raw;
                chunk_a(t, ring + s * P3_SLOT, c, yprev, raw);
                __syncwarp();
                if (lane == 0) fs_mbar_arrive(bar_empty + 8 * s);
                fs_tmem_st(taddr + c * W, raw);
                ++g;
            }
            fs_tmem_wait_st();
        };

        auto sweep_b = [&](const TileC &t) {
            Raw<W> ra, rb;
            double u1n = 0.0, y2n = 0.0;
            auto chunk_b = [&](Raw<W> &cur, Raw<W> &nxt, int c) {
                if (c > 0) fs_tmem_ld(taddr + (c - 1) * W, nxt);
                const uint32_t s = g % NS, ph = (g / NS) & 1;
                fs_mbar_wait<P3_CWAIT>(bar_full + 8 * s, ph);
                Vd<1> ycur[KC];
                fs_unpack<1, KC>(cur, ycur);
                const unsigned char *sb = ring + s * P3_SLOT;
                const double2 *pl = reinterpret_cast<const double2 *>(sb + P3_T * P3_UBOX + cls * 4 * P3_PP) + col;
                // stage-1 back substitution of the chunk, published to the two other tracers' warps
                double *ex = exch + (size_t)(gx & 1) * (P3_T * KC * 64);
                double u1v[KC];
#pragma unroll
                for (int q = KC - 1; q >= 0; --q) {
                    const double2 gm = pl[2 * PP + q * FS_COLS];
                    u1n = fma(-gm.x, u1n, gm.y * ycur[q].v[0]);
                    u1v[q] = u1n;
                    ex[(tr * KC + q) * 64 + exl] = u1n;
                }
                double y1top = 0.0;
                if (c > 0) {
                    fs_tmem_wait_ld(nxt);
                    y1top = __hiloint2double((int)nxt.w[(KC - 1) * 2 + 1], (int)nxt.w[(KC - 1) * 2]);
                }
                // (an mbarrier arrive here with the wait moved below the own-tracer part measured 4 % slower than
                // the named barrier, and forcing the pair loads to LDS.128 through inline asm 4 % slower than
                // leaving their scheduling to the compiler: scripts/ab_variants.sh, profiles/r02_p3_variants.md)
                asm volatile("bar.sync %0, 96;" ::"r"(1 + pr) : "memory");
                // everything of the stage-2 right-hand side that uses this warp's own tracer, then the coupled sources;
                // every pair plane is read once as a whole double2 so that the loads stay LDS.128 (a pair whose .x and
                // .y are first used far apart is split into two LDS.64 of two wavefronts each: 52 -> 44 per level)
                double part[KC];
                double2 mf[KC];
#pragma unroll
                for (int q = KC - 1; q >= 0; --q) {
                    const double y1 = ycur[q].v[0];
                    const double y1m = (q > 0) ? ycur[q > 0 ? q - 1 : 0].v[0] : y1top;
                    const double2 lc = pl[q * FS_COLS], ri = pl[PP + q * FS_COLS];
                    mf[q] = pl[3 * PP + q * FS_COLS];
                    const double u1 = u1v[q];
                    double rhs1 = fma(ri.y, y1m, y1);
                    if (c == 0 && q == 0) rhs1 -= t.aff1;
                    const double un = *reinterpret_cast<const double *>(sb + tr * P3_UBOX + offU[q][1]);
                    const double pp = fma(p.r, rhs1, p.a0r * un);
                    const double ul = __shfl_up_sync(0xffffffffu, u1, 2, 32), ur = __shfl_down_sync(0xffffffffu, u1, 2, 32);
                    part[q] = fma(lc.x, ul, fma(ri.x, ur, fma(lc.y, u1, pp)));
                    if (c == 0 && q == 0) part[q] += t.aff2;
                }
                double sv[KC];
#pragma unroll
                for (int q = 0; q < KC; ++q)
                    sv[q] = p3_source(mf[q].y, cu, t.kd2, t.kq2, hs, ex[q * 64 + exl], ex[(KC + q) * 64 + exl],
                                      ex[(2 * KC + q) * 64 + exl]);
                Vd<1> yb[KC];
#pragma unroll
                for (int q = KC - 1; q >= 0; --q) {
                    y2n = fma(-mf[q].x, y2n, part[q] + sv[q]);
                    yb[q].v[0] = y2n;
                }
                __syncwarp();
                if (lane == 0) fs_mbar_arrive(bar_empty + 8 * s);
                Raw<W> raw;
                fs_pack<1, KC>(yb, raw);
                fs_tmem_st(taddr + c * W, raw);
                ++g;
                ++gx;
            };
            fs_tmem_ld(taddr + (nchunk - 1) * W, ra);
            fs_tmem_wait_ld(ra);
            int c = nchunk - 1;
            for (; c >= 1; c -= 2) {
                chunk_b(ra, rb, c);
                chunk_b(rb, ra, c - 1);
            }
            if (c == 0) chunk_b(ra, rb, 0);
            fs_tmem_wait_st();
        };

        auto sweep_c = [&](bool with_a, const TileC &ta) {
            Raw<W> ra, rb;
            double u2p = 0.0, yprev = 0.0;
            auto chunk_ca = [&](Raw<W> &cur, Raw<W> &nxt, int c) {
                if (c + 1 < nchunk) fs_tmem_ld(taddr + (c + 1) * W, nxt);
                const uint32_t s = g % NS, ph = (g / NS) & 1;
                const uint32_t so = go % NO, pho = (go / NO) & 1;
                fs_mbar_wait<P3_CWAIT>(bar_full + 8 * s, ph);
                fs_mbar_wait<P3_CWAIT>(bar_oempty + 8 * so, pho ^ 1);
                const unsigned char *sb = ring + s * P3_SLOT;
                chunk_c(sb, oring + so * P3_OUT, cur, u2p);
                Raw<W> raw;
                if (with_a) chunk_a(ta, sb, c, yprev, raw);
                __syncwarp();
                if (lane == 0) {
                    fs_mbar_arrive(bar_empty + 8 * s);
                    fs_mbar_arrive(bar_ofull + 8 * so);
                }
                if (with_a) fs_tmem_st(taddr + c * W, raw);
                if (c + 1 < nchunk) fs_tmem_wait_ld(nxt);
                ++g;
                ++go;
            };
            fs_tmem_ld(taddr, ra);
            fs_tmem_wait_ld(ra);
            int c = 0;
            for (; c + 1 < nchunk; c += 2) {
                chunk_ca(ra, rb, c);
                chunk_ca(rb, ra, c + 1);
            }
            if (c < nchunk) chunk_ca(ra, rb, c);
            if (with_a) fs_tmem_wait_st();
        };

        // one member block with at most two members (a single state): the three warps of the second member pair
        // only keep the barrier protocol going (see step_fused_kernel)
        if (p.nmb == 1 && 2 * pr >= p.B) {
            auto pass = [&](bool with_out) {
                for (int c = 0; c < nchunk; ++c) {
                    const uint32_t s = g % NS, ph = (g / NS) & 1;
                    fs_mbar_wait<P3_CWAIT>(bar_full + 8 * s, ph);
                    if (with_out) {
                        const uint32_t so = go % NO, pho = (go / NO) & 1;
                        fs_mbar_wait<P3_CWAIT>(bar_oempty + 8 * so, pho ^ 1);
                        if (lane == 0) fs_mbar_arrive(bar_ofull + 8 * so);
                        ++go;
                    }
                    if (lane == 0) fs_mbar_arrive(bar_empty + 8 * s);
                    ++g;
                }
            };
            Item it = it0;
            if (item_valid(it)) {
                pass(false);
                while (true) {
                    pass(false);
                    const Item nx = item_next(it);
                    if (!item_valid(nx)) {
                        pass(true);
                        break;
                    }
                    pass(true);
                    if (item_depends(nx, it)) pass(false);
                    it = nx;
                }
            }
        } else {
        Item it = it0;
        if (item_valid(it)) {
            TileC t = tile_c(it);
            sweep_a(t);
            while (true) {
                sweep_b(t);
                const Item nx = item_next(it);
                if (!item_valid(nx)) {
                    sweep_c(false, t);
                    break;
                }
                const TileC tn = tile_c(nx);
                if (!item_depends(nx, it)) {
                    sweep_c(true, tn);
                } else {
                    sweep_c(false, t);
                    sweep_a(tn);
                }
                it = nx;
                t = tn;
            }
        }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
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

// ---- phosphorus: conversion between the member-fastest layout [row][ldb] (row = (tracer, level, column))
// and the member-block-major layout of the step kernel (member m -> pair m >> 1 of block m >> 2, slot
// m & 1).  Members B .. 4*nmb-1 are zero in the copy.
__global__ void p3_to_tm_kernel(const double *__restrict__ src, double *__restrict__ dst, int ny, int B, size_t ldb,
                                int nmb) {
    const int tk = blockIdx.y;
    const int wm = 4 * nmb;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ny * wm) return;
    const int j = idx / wm, m = idx % wm;
    const double v = (m < B) ? src[((size_t)tk * ny + j) * ldb + m] : 0.0;
    dst[(((size_t)tk * nmb * 2 + (m >> 1)) * ny + j) * 2 + (m & 1)] = v;
}
// f = x_tm(T) - x0, back in the member-fastest layout
__global__ void p3_from_tm_sub_kernel(const double *__restrict__ tm, const double *__restrict__ x0,
                                      double *__restrict__ f, int ny, int B, size_t ldb, int nmb) {
    const int tk = blockIdx.y;
    const int wm = 4 * nmb;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ny * wm) return;
    const int j = idx / wm, m = idx % wm;
    if (m >= B) return;
    const size_t o = ((size_t)tk * ny + j) * ldb + m;
    f[o] = tm[(((size_t)tk * nmb * 2 + (m >> 1)) * ny + j) * 2 + (m & 1)] - x0[o];
}
__global__ void p3_gather_member_kernel(const double *__restrict__ tm, double *__restrict__ dst, int ny, int nmb,
                                        size_t n, int b) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (tracer, level, column)
    if (i >= n) return;
    const size_t tk = i / ny, j = i % ny;
    dst[i] = tm[((tk * nmb * 2 + (b >> 1)) * ny + j) * 2 + (b & 1)];
}

int launch_p3_to_tm(const ModelDev &v, const double *src, double *dst, int B, size_t ldb, cudaStream_t st) {
    const int nmb = (B + P3_MEM - 1) / P3_MEM;
    dim3 grid((unsigned)((v.ny * 4 * nmb + 255) / 256), (unsigned)(v.T * v.nz));
    p3_to_tm_kernel<<<grid, 256, 0, st>>>(src, dst, v.ny, B, ldb, nmb);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}
int launch_p3_from_tm_sub(const ModelDev &v, const double *tm, const double *x0, double *f, int B, size_t ldb,
                          cudaStream_t st) {
    const int nmb = (B + P3_MEM - 1) / P3_MEM;
    dim3 grid((unsigned)((v.ny * 4 * nmb + 255) / 256), (unsigned)(v.T * v.nz));
    p3_from_tm_sub_kernel<<<grid, 256, 0, st>>>(tm, x0, f, v.ny, B, ldb, nmb);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}
int launch_p3_gather_member(const ModelDev &v, const double *tm, double *dst, int B, int b, cudaStream_t st) {
    const int nmb = (B + P3_MEM - 1) / P3_MEM;
    const size_t n = (size_t)v.T * v.nz * v.ny;
    p3_gather_member_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tm, dst, v.ny, nmb, n, b);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}
bool fused_tile_major(const ModelDev &v) { return v.kind == NKB_MOD_PHOSPHORUS; }

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
                     const cuuint32_t *box, CUtensorMapSwizzle swz,
                     CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
    FsEncodeTiledFn fn = fs_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return 1;
    }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, const_cast<void *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
        return 1;
    }
    return 0;
}

// members per thread of the consumer layout (NKB_FUSED_MPT: 1 = 8 warps x 8-level chunks, 2 = 4 warps x 4)
static int fs_mpt() { return fs_env_int("NKB_FUSED_MPT", 1) == 2 ? 2 : 1; }

// geometry of the column tiling: nct tiles of jt <= 14 interior columns
static void fs_col_tiles(int ny, int &nct, int &jt) {
    int jmax = fs_env_int("NKB_FUSED_JT", FS_JT);
    if (jmax < 1) jmax = 1;
    if (jmax > FS_JT) jmax = FS_JT;
    nct = (ny + jmax - 1) / jmax;
    jt = (ny + nct - 1) / nct;
}

bool fused_step_usable(const ModelDev &v, int B, int ldb, const double *x0, const double *f, const double *work) {
    if (fs_env_int("NKB_FUSED", 1) == 0) return false;
    if (v.column_model == 1 && v.ny == 1) return false;
    if (v.kind == NKB_MOD_PHOSPHORUS) {
        if (fs_env_int("NKB_FUSED_P3", 1) == 0 || v.T != P3_T || v.n_classes != P3_NCLS) return false;
        if (v.class_of[0] != 0 || v.class_of[1] != 0 || v.class_of[2] != 1) return false;  // po4, dop: class 0; pop: class 1
        if (((B + P3_MEM - 1) / P3_MEM) * P3_MEM > ldb) return false;  // the member-block-major copy pads B to 4
    } else if (v.kind != NKB_MOD_LINEAR && v.kind != NKB_MOD_FORCED_FILE) {
        return false;
    }
    if (v.nz > 128) return false;  // TMEM: 2 x 32-bit columns per level and member, 256 per thread
    if (B < fs_env_int("NKB_FUSED_MIN_B", 1) || (ldb % 2) != 0) return false;
    if (((uintptr_t)x0 | (uintptr_t)f | (uintptr_t)work) & 15) return false;
    return fs_encode_fn() != nullptr;
}

int fused_encode_state_maps(const ModelDev &v, int B, int ldb, const double *buf, CUtensorMap *in, CUtensorMap *out) {
    int nct, jt;
    fs_col_tiles(v.ny, nct, jt);
    if (v.kind == NKB_MOD_PHOSPHORUS) {
        // member-block-major copy: [tracer][level][member pair][column * 2 members]
        const cuuint64_t npair = 2 * (cuuint64_t)((B + P3_MEM - 1) / P3_MEM);
        const cuuint64_t run = (cuuint64_t)v.ny * 2;
        const cuuint64_t dims[4] = {run, npair, (cuuint64_t)v.nz, (cuuint64_t)v.T};
        const cuuint64_t strides[3] = {run * 8, npair * run * 8, (cuuint64_t)v.nz * npair * run * 8};
        const cuuint32_t box_in[4] = {FS_UCOLS * 2, 2, P3_KC, 1};
        const cuuint32_t box_out[4] = {(cuuint32_t)jt * 2, 2, P3_KC, 1};
        // the 288-byte runs of a tile are not aligned to anything: with the 256-byte L2 promotion every run pulls
        // two or three 256-byte chunks (ncu: 2.1x the state read per step).  NKB_P3_L2PROMO = 0 none, 1 64 B (default:
        // +2.5 % against 256 B, measured), 2 128 B, 3 256 B
        CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        switch (fs_env_int("NKB_P3_L2PROMO", 1)) {
            case 0: promo = CU_TENSOR_MAP_L2_PROMOTION_NONE; break;
            case 1: promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B; break;
            case 2: promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; break;
            default: break;
        }
        if (in && fs_encode(in, buf, 4, dims, strides, box_in, CU_TENSOR_MAP_SWIZZLE_NONE, promo)) return 1;
        if (out && fs_encode(out, buf, 4, dims, strides, box_out, CU_TENSOR_MAP_SWIZZLE_NONE, promo)) return 1;
        return 0;
    }
    const cuuint32_t kc = (cuuint32_t)fs_cfg(fs_mpt()).kc;
    // members beyond B are never read (zero-filled) nor written (clipped)
    const cuuint64_t dims[4] = {(cuuint64_t)B, (cuuint64_t)v.ny, (cuuint64_t)v.nz, (cuuint64_t)v.T};
    const cuuint64_t strides[3] = {(cuuint64_t)ldb * 8, (cuuint64_t)v.ny * ldb * 8, (cuuint64_t)v.nz * v.ny * ldb * 8};
    const cuuint32_t box_in[4] = {FS_MEM, FS_UCOLS, kc, 1};
    const cuuint32_t box_out[4] = {FS_MEM, (cuuint32_t)jt, kc, 1};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    switch (fs_env_int("NKB_FUSED_L2PROMO", 3)) {
        case 0: promo = CU_TENSOR_MAP_L2_PROMOTION_NONE; break;
        case 1: promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B; break;
        case 2: promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; break;
        default: break;
    }
    if (in && fs_encode(in, buf, 4, dims, strides, box_in, CU_TENSOR_MAP_SWIZZLE_128B, promo)) return 1;
    if (out && fs_encode(out, buf, 4, dims, strides, box_out, CU_TENSOR_MAP_SWIZZLE_128B, promo)) return 1;
    return 0;
}

int fused_encode_ctab_map(int nz, int ny, size_t nplanes, const double *buf, CUtensorMap *map, int kind) {
    // rows of (ny + 1) coefficient pairs, one zero pair on the left: the box of a column tile starts at
    // column j0 - 1 = pair index j0, double index 2*j0 — always even.  TMA faults ("illegal
    // instruction") on a box whose innermost start address is not 16-byte aligned (an odd float64
    // coordinate), measured on B200.
    const cuuint64_t np = 2 * ((cuuint64_t)ny + 1);
    const cuuint64_t dims[3] = {np, (cuuint64_t)nz, (cuuint64_t)nplanes};
    const cuuint64_t strides[2] = {np * 8, (cuuint64_t)nz * np * 8};
    const cuuint32_t kc = (kind == NKB_MOD_PHOSPHORUS) ? (cuuint32_t)P3_KC : (cuuint32_t)fs_cfg(fs_mpt()).kc;
    const cuuint32_t box[3] = {2 * FS_COLS, kc, 1};
    return fs_encode(map, buf, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

template <int KIND, int MPT>
static int fs_launch_m(const StepArgs &a, const StepMaps &maps, int grid, bool cooperative, cudaStream_t st) {
    auto kern = step_fused_kernel<KIND, MPT>;
    constexpr FsCfg C = fs_cfg(MPT);
    static unsigned long long attr_mask = 0;
    if (nkb::first_use_on_device(attr_mask)) {
        NKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C.smem));
    }
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C.threads);
    cfg.dynamicSmemBytes = C.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait on each other's tiles
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cooperative ? 1 : 0;
    const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, a, maps);
    if (cooperative && err == cudaErrorCooperativeLaunchTooLarge) {
        cudaGetLastError();  // the SMs are not all available: the caller falls back to one launch per step
        return -1;
    }
    NKB_CUDA(err);
    count_launch();
    return 0;
}

// integrates the time steps [step0, step1) in one launch
int launch_steps_fused(const ModelDev &v, int B, int n_steps, int step0, int step1, const double *d_h,
                       const double *d_aff, const FusedMaps &fm, int *d_done, int *d_err, cudaStream_t st) {
    StepArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nz = v.nz; a.ny = v.ny; a.B = B; a.T = v.T; a.ncls = v.n_classes; a.n_steps = n_steps;
    fs_col_tiles(v.ny, a.nct, a.jt);
    const bool p3 = (v.kind == NKB_MOD_PHOSPHORUS);
    a.nmb = p3 ? (B + P3_MEM - 1) / P3_MEM : (B + FS_MEM - 1) / FS_MEM;
    a.ntiles = p3 ? a.nmb * a.nct : a.nmb * a.nct * v.T;
    a.p3_hs = v.po4_halfsat; a.p3_sigma = v.sigma; a.p3_rd = v.dop_remin_rate; a.p3_rp = v.pop_remin_rate;
    a.step0 = step0; a.step1 = step1;
    for (int t = 0; t < NKB_MAX_TRACERS; ++t) { a.class_of[t] = v.class_of[t]; a.src_const[t] = v.src_const[t]; }
    a.sink_thres_r = v.sink_thres > 0.0 ? 1.0 / v.sink_thres : 0.0;
    // stage 2 (nkb_api.cu): rhs2 = a0 u_n + a1 u1 + h (delta - 1 + gamma) E(u_n) + h (1 - delta) E(u1); with
    // rhs1 = u_n + gamma h E(u_n) the u_n terms are P = r rhs1 + (a0 - r) u_n, r = (delta - 1 + gamma)/gamma
    const double a1 = (1.0 - kGamma) / kGamma, a0 = 1.0 - a1;
    a.r = (kDelta - 1.0 + kGamma) / kGamma;
    a.a0r = a0 - a.r;
    a.h = d_h; a.aff = d_aff; a.done = d_done; a.err = d_err;
    StepMaps maps;
    maps.in_x0 = fm.in_x0; maps.in_f = fm.in_f; maps.in_w = fm.in_w;
    maps.out_f = fm.out_f; maps.out_w = fm.out_w; maps.ctab = fm.ctab;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        NKB_CUDA(cudaGetDevice(&dev));
        NKB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = fs_env_int("NKB_FUSED_GRID", n_sm);
    if (grid > n_sm) grid = n_sm;  // one CTA per SM (shared memory, 512 TMEM columns): all co-resident
    a.tgroup = 1;
    if (p3) {
        a.tgroup = fs_env_int("NKB_P3_GROUP", 1);
        if (a.tgroup < 1 || a.tgroup > 8) a.tgroup = 1;
    }
    const int ngroups = (a.ntiles + a.tgroup - 1) / a.tgroup;
    if (grid > ngroups) grid = ngroups;
    // rotate the tile -> CTA assignment by the number of left-over tiles per step so that the CTAs that
    // get one tile more than the others change from step to step
    a.rot = (step1 - step0 > 1) ? ngroups % grid : 0;
    a.cross_step_fuse = (a.ntiles > 2 * a.tgroup * grid + fs_nbr(a.nmb)) ? 1 : 0;
    const bool coop = (step1 - step0 > 1);
    if (p3) {
        static unsigned long long attr_mask = 0;
        if (nkb::first_use_on_device(attr_mask)) {
            NKB_CUDA(cudaFuncSetAttribute(step_fused_p3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM));
        }
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(P3_THREADS);
        cfg.dynamicSmemBytes = P3_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = coop ? 1 : 0;
        const cudaError_t err = cudaLaunchKernelEx(&cfg, step_fused_p3_kernel, a, maps);
        if (coop && err == cudaErrorCooperativeLaunchTooLarge) {
            cudaGetLastError();
            return -1;
        }
        NKB_CUDA(err);
        count_launch();
        return 0;
    }
    const int mpt = fs_mpt();
    if (v.kind == NKB_MOD_LINEAR)
        return mpt == 2 ? fs_launch_m<NKB_MOD_LINEAR, 2>(a, maps, grid, coop, st)
                        : fs_launch_m<NKB_MOD_LINEAR, 1>(a, maps, grid, coop, st);
    if (v.kind == NKB_MOD_FORCED_FILE)
        return mpt == 2 ? fs_launch_m<NKB_MOD_FORCED_FILE, 2>(a, maps, grid, coop, st)
                        : fs_launch_m<NKB_MOD_FORCED_FILE, 1>(a, maps, grid, coop, st);
    set_error("launch_steps_fused: unsupported module kind");
    return 2;
}

// a single state in the reference's own layout (B == 1, ldb == 1) is staged into a 4-lane batch so that it
// can take the fused step kernels too (TMA needs a 16-byte member pitch): one launch per model year
// instead of two per time step — 3 to 5 times faster for the Newton iterate and the Krylov products
bool fused_single_state(const ModelDev &v) {
    if (fs_env_int("NKB_FUSED", 1) == 0 || fs_env_int("NKB_FUSED_MIN_B", 1) > 1) return false;
    if (v.column_model == 1 && v.ny == 1) return false;
    if (v.nz > 128 || fs_encode_fn() == nullptr) return false;
    if (v.kind == NKB_MOD_PHOSPHORUS) return fs_env_int("NKB_FUSED_P3", 1) != 0 && v.T == P3_T && v.n_classes == P3_NCLS;
    return v.kind == NKB_MOD_LINEAR || v.kind == NKB_MOD_FORCED_FILE;
}

bool fused_persistent() { return fs_env_int("NKB_FUSED_PERSIST", 1) != 0; }

int fused_tile_count(const ModelDev &v, int B) {
    int nct, jt;
    fs_col_tiles(v.ny, nct, jt);
    if (v.kind == NKB_MOD_PHOSPHORUS) return ((B + P3_MEM - 1) / P3_MEM) * nct;
    return ((B + FS_MEM - 1) / FS_MEM) * nct * v.T;
}

int launch_sub_inplace(double *out, const double *x0, size_t n, cudaStream_t st) {
    const size_t n2 = n / 2;  // n is a multiple of ldb (even)
    sub_inplace_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(out, x0, n2);
    count_launch();
    NKB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nkb
