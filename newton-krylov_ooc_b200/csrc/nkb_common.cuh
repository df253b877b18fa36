// Shared declarations of the nkb200 CUDA library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "nkb200.h"

namespace nkb {

void set_error(const std::string &msg);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define NKB_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            ::nkb::set_error(std::string(#call) + ": " + cudaGetErrorString(err__) + " at " + \
                             __FILE__ + ":" + std::to_string(__LINE__));                      \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

#define NKB_REQUIRE(cond, msg)                   \
    do {                                         \
        if (!(cond)) {                           \
            ::nkb::set_error(std::string(msg));  \
            return 2;                            \
        }                                        \
    } while (0)

// ARS(2,2,2) IMEX Runge-Kutta constants (Ascher, Ruuth, Spiteri 1997)
constexpr double kGamma = 0.29289321881345247559915563789515;   // 1 - 1/sqrt(2)
constexpr double kDelta = -0.70710678118654752440084436210485;  // 1 - 1/(2 gamma)

// device-side description of one tracer module (time-invariant tables live in HBM)
struct ModelDev {
    int nz, ny, T, kind, n_classes, column_model;
    int class_of[NKB_MAX_TRACERS];
    double t0, t1;
    const double *depth_edges;  // [nz+1]
    const double *depth_mid;    // [nz]
    const double *dz_r;         // [nz]
    const double *dz_mid;       // [nz-1]
    const double *dz_mid_r;     // [nz-1]
    const double *wvel;         // [nz+1][ny] (boundary rows zeroed)
    const double *estencil;     // [nz][ny][4] {eL, eC, eR, 0} or nullptr
    const double *bld_max;      // [ny]
    double surf_diag[NKB_MAX_CLASSES], surf_aff[NKB_MAX_CLASSES], decay[NKB_MAX_CLASSES],
        sink_vel[NKB_MAX_CLASSES];
    int n_flux_pts;
    double flux_t[8], flux_v[8];
    double src_const[NKB_MAX_TRACERS];
    double sink_thres;
    int n_frc;
    const double *frc_time;  // [n_frc]
    const double *frc_data;  // [n_frc][nz][ny]
    const double *light;     // [nz][ny]
    double po4_halfsat, max_uptake_rate, sigma, dop_remin_rate, pop_remin_rate;
    int po4_s_restoring_opt;
    int n_srf;
    const double *srf_time;  // [n_srf]
    const double *srf_data;  // [n_srf][ny]
    double srf_rate[NKB_MAX_CLASSES];
};

// arguments of one fused stage launch (K1+K2)
struct StageArgs {
    const double *u[2];   // stage inputs [T][nz][ny][ldb]
    double *out;          // [T][nz][ny][ldb]
    const double *sub;    // optional: out = x - sub (final F = x(T) - x(0)), else nullptr
    double a[2];          // rhs = sum_i a[i]*u[i] + he[i]*E(u[i])
    double he[2];
    const double *est;    // [nz][ny][4] {eL, eC, eR, 0} or nullptr
    const double *tri;    // [ncls][nz][ny][4] {ib, g, m, 0} of this stage ({sub, diag, sup, 0} raw, for tend)
    const double *aff;    // [ncls][ny]  hg*affine surface source
    const double *src2;   // [nz][ny][2] forcing at the explicit times of inputs 0 and 1 (FORCED_FILE)
    const double *light;  // [nz][ny]
    int nz, ny, B, ldb, T;
    int hints;            // bit0: evict_first on streamed data, bit1: evict_last on intermediates
    int ksm;              // levels [0, ksm) keep the forward-sweep intermediates in shared memory
    int class_of[NKB_MAX_TRACERS];
    double src_const[NKB_MAX_TRACERS];
    double sink_thres_r;  // 1/sink_thres or 0
    double halfsat, umax, sigma, rdop, rpop;
    int restoring_opt;    // test_problem phosphorus: po4_s restoring option (tendency kernel only)
};

// arguments of the persistent column-model year kernel (nkb_column.cu)
struct ColumnArgs {
    const double *x0;      // [T][nz][ldb]
    double *out;           // [T][nz][ldb]  F = x(T) - x(0)
    const double *tri;     // [n_stages][ncls][nz][4] {ib, g, m, 0}
    const double *aff;     // [n_stages][ncls]
    const double *h;       // [n_steps]
    const double *light;   // [nz] (phosphorus)
    const int *hist_slot;  // [n_steps+1] slot index or -1 (nullptr: no hist)
    double *hist;          // [n_hist][T][nz]
    int nz, B, ldb, T, n_steps, ncls;
    int class_of[NKB_MAX_TRACERS];
    double src_const[NKB_MAX_TRACERS];
    int restoring_opt;
};

int launch_column_year(int kind, const ColumnArgs &a, cudaStream_t st);
int launch_stage_tables(const ModelDev &m, int n_stages, const double *d_t, const double *d_hg, int mode,
                        double *tri, double *aff, cudaStream_t st);
int launch_mixing_coeff(const ModelDev &m, double time, double *out, cudaStream_t st);
int launch_forcing_tables(const ModelDev &m, int n_times, const double *d_t, double *src, cudaStream_t st);
int launch_gather_member(const double *src, double *dst, size_t n, size_t ldb, int b, cudaStream_t st);
int launch_scatter_member(const double *src, double *dst, size_t n, size_t ldb, cudaStream_t st);
bool fused_single_state(const ModelDev &v);
int launch_pack(const double *src, double *dst, int n, int B, int ldb, cudaStream_t st);
int launch_unpack(const double *src, double *dst, int n, int B, int ldb, cudaStream_t st);
int launch_stage(int kind, int nin, const StageArgs &a, cudaStream_t st);
int launch_tend(int kind, const StageArgs &a, cudaStream_t st);
int launch_step_ctab(const ModelDev &m, int n_steps, const double *d_t, const double *d_hg, const double *d_texp,
                     double *ctab, cudaStream_t st);
// fused step kernel (nkb_step_fused.cu)
bool fused_step_usable(const ModelDev &v, int B, int ldb, const double *x0, const double *f, const double *work);
int fused_encode_state_maps(const ModelDev &v, int B, int ldb, const double *buf, CUtensorMap *in, CUtensorMap *out);
int fused_encode_ctab_map(int nz, int ny, size_t nplanes, const double *buf, CUtensorMap *map, int kind);
struct FusedMaps {
    CUtensorMap in_x0, in_f, in_w, out_f, out_w, ctab;
};
int launch_steps_fused(const ModelDev &v, int B, int n_steps, int step0, int step1, const double *d_h,
                       const double *d_aff, const FusedMaps &fm, int *d_done, int *d_err, cudaStream_t st);
bool fused_persistent();
int fused_tile_count(const ModelDev &v, int B);
int launch_sub_inplace(double *out, const double *x0, size_t n, cudaStream_t st);
// phosphorus step kernel: the year is integrated in a member-block-major copy of the state
bool fused_tile_major(const ModelDev &v);
int launch_p3_to_tm(const ModelDev &v, const double *src, double *dst, int B, size_t ldb, cudaStream_t st);
int launch_p3_from_tm_sub(const ModelDev &v, const double *tm, const double *x0, double *f, int B, size_t ldb,
                          cudaStream_t st);
int launch_p3_gather_member(const ModelDev &v, const double *tm, double *dst, int B, int b, cudaStream_t st);
bool tma_path_usable(const StageArgs &a);
int launch_stage_tma(int kind, int nin, const StageArgs &a, cudaStream_t st);


// cudaFuncSetAttribute is per device (and context): a process that switches devices must repeat it there.
// `mask` is a function-local static, one bit per device ordinal.
inline bool first_use_on_device(unsigned long long &mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

// device temporary released on every exit path (the NKB_CUDA / NKB_REQUIRE macros return early)
template <class T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, count * sizeof(T)); }
};
}  // namespace nkb

// the opaque handle of the C ABI
struct nkb_model {
    nkb::ModelDev dev;
    void *arena = nullptr;      // one allocation holding all time-invariant tables
    // schedule
    int n_steps = 0;
    double *h_t_start = nullptr, *h_h = nullptr;  // host copies
    double *d_h = nullptr;                         // device copy of the step sizes
    // per-stage tables, stage s = 2*step + {0,1}
    double *tri = nullptr;      // [n_stages][n_classes][nz][ny][4]  {ib, g, m, 0}
    double *aff = nullptr;      // [n_stages][n_classes][ny]  h*gamma*(affine surface source) per column
    double *src = nullptr;      // [n_steps][nz][ny][2] forcing at the two explicit stage times (or nullptr)
    // fused-step coefficient table (built on first use, nkb_tables.cu:step_ctab_kernel)
    double *ctab = nullptr;     // [n_steps][n_classes][8][nz][2*(ny+1)]
    CUtensorMap map_ctab;
    int *d_done = nullptr;      // per-tile step counters of the persistent step kernel
    size_t done_cap = 0;
    int *h_err = nullptr, *d_err = nullptr;  // host-mapped error flag (dependency-wait timeout)
    int *d_hist_slot = nullptr;              // column model: hist slot of every step (or -1)
    size_t hist_slot_cap = 0;
    // scratch for tend()/mixing_coeff()
    double *tri_raw = nullptr;  // [n_classes][nz][ny][4]
    double *aff_raw = nullptr;  // [n_classes][ny]
    double *src_raw = nullptr;  // [nz][ny][2]
    // host-buffer path
    double *d_stage_major = nullptr, *d_stage_x = nullptr, *d_stage_f = nullptr, *d_stage_work = nullptr;
    size_t stage_cap = 0;
    cudaGraphExec_t graph = nullptr;
    const double *graph_x0 = nullptr;
    double *graph_f = nullptr, *graph_work = nullptr;
    int graph_B = 0, graph_ldb = 0;
    cudaStream_t own_stream = nullptr;
};
