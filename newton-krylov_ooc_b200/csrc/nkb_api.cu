// C ABI of the nkb200 library: model handles, schedule tables, the model-year evaluation loop.
#include <cstring>
#include <vector>

#include "nkb_common.cuh"

namespace nkb {

static thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};
void set_error(const std::string &msg) { g_error = msg; }

}  // namespace nkb

using nkb::ModelDev;
using nkb::StageArgs;

extern "C" {

const char *nkb_last_error(void) { return nkb::g_error.c_str(); }
int nkb_version(void) { return 100; }
uint64_t nkb_launch_count(void) { return nkb::g_launches.load(); }

int nkb_model_create(nkb_model **out, const nkb_model_desc *d) {
    NKB_REQUIRE(out && d, "nkb_model_create: null argument");
    NKB_REQUIRE(d->nz >= 2 && d->ny >= 1, "nkb_model_create: need nz >= 2, ny >= 1");
    NKB_REQUIRE(d->n_tracers >= 1 && d->n_tracers <= NKB_MAX_TRACERS, "nkb_model_create: bad n_tracers");
    NKB_REQUIRE(d->n_classes >= 1 && d->n_classes <= NKB_MAX_CLASSES, "nkb_model_create: bad n_classes");
    NKB_REQUIRE(d->h_depth_edges, "nkb_model_create: depth edges missing");
    NKB_REQUIRE(d->column_model == 1 || (d->h_bld_max && d->h_ypos_mid), "nkb_model_create: bld_max missing");
    NKB_REQUIRE(d->kind != NKB_MOD_PHOSPHORUS || (d->n_tracers == 3 && d->h_light),
                "nkb_model_create: phosphorus needs 3 tracers and a light table");
    NKB_REQUIRE(d->kind != NKB_MOD_PHOSPHORUS_1D || (d->n_tracers == 6 && d->h_light && d->ny == 1),
                "nkb_model_create: test_problem phosphorus needs 6 tracers, ny == 1 and a light table");
    NKB_REQUIRE(d->kind != NKB_MOD_FORCED_FILE || (d->n_frc >= 2 && d->h_frc_time && d->h_frc_data),
                "nkb_model_create: forced file module needs >= 2 forcing records");
    NKB_REQUIRE(d->n_srf == 0 || (d->n_srf >= 2 && d->h_srf_time && d->h_srf_data),
                "nkb_model_create: a surface restoring record needs >= 2 times");
    for (int t = 0; t < d->n_tracers; ++t)
        NKB_REQUIRE(d->class_of[t] >= 0 && d->class_of[t] < d->n_classes, "nkb_model_create: bad class_of");

    const int nz = d->nz, ny = d->ny;
    const size_t plane = (size_t)nz * ny;
    // host staging of every time-invariant table, then one upload
    std::vector<double> host;
    auto push = [&](const double *src, size_t n) {
        size_t off = host.size();
        host.insert(host.end(), src, src + n);
        while (host.size() % 2) host.push_back(0.0);
        return off;
    };
    std::vector<double> mid(nz), dzr(nz), dzm(nz - 1), dzmr(nz - 1);
    for (int k = 0; k < nz; ++k) {
        mid[k] = 0.5 * (d->h_depth_edges[k] + d->h_depth_edges[k + 1]);  // spatial_axis.py:35
        dzr[k] = 1.0 / (d->h_depth_edges[k + 1] - d->h_depth_edges[k]);  // :36-37
    }
    for (int k = 0; k < nz - 1; ++k) {
        dzm[k] = mid[k + 1] - mid[k];  // :38
        dzmr[k] = 1.0 / dzm[k];        // :39
    }
    const size_t o_edges = push(d->h_depth_edges, nz + 1);
    const size_t o_mid = push(mid.data(), nz);
    const size_t o_dzr = push(dzr.data(), nz);
    const size_t o_dzm = push(dzm.data(), nz - 1);
    const size_t o_dzmr = push(dzmr.data(), nz - 1);
    std::vector<double> w((size_t)(nz + 1) * ny, 0.0);
    if (d->h_wvel) {
        std::memcpy(w.data(), d->h_wvel, w.size() * sizeof(double));
        for (int j = 0; j < ny; ++j) {  // boundary faces carry no flux (advection.py:67-70)
            w[j] = 0.0;
            w[(size_t)nz * ny + j] = 0.0;
        }
    }
    const size_t o_w = push(w.data(), w.size());
    size_t o_est = 0, o_bld = 0, o_ft = 0, o_fd = 0, o_light = 0, o_st = 0, o_sd = 0;
    if (d->h_estencil) {  // [3][nz][ny] -> packed [nz][ny][4] {eL, eC, eR, 0}
        std::vector<double> e4(4 * plane, 0.0);
        for (size_t c = 0; c < plane; ++c)
            for (int q = 0; q < 3; ++q) e4[4 * c + q] = d->h_estencil[(size_t)q * plane + c];
        o_est = push(e4.data(), e4.size());
    }
    if (d->h_bld_max) o_bld = push(d->h_bld_max, ny);
    if (d->n_frc > 0) {
        o_ft = push(d->h_frc_time, d->n_frc);
        o_fd = push(d->h_frc_data, (size_t)d->n_frc * plane);
    }
    if (d->h_light) o_light = push(d->h_light, plane);
    if (d->n_srf > 0) {
        o_st = push(d->h_srf_time, d->n_srf);
        o_sd = push(d->h_srf_data, (size_t)d->n_srf * ny);
    }

    nkb_model *m = new nkb_model();
    if (cudaMalloc(&m->arena, host.size() * sizeof(double)) != cudaSuccess) {
        delete m;
        nkb::set_error("nkb_model_create: cudaMalloc failed (is a CUDA device present?)");
        return 1;
    }
    NKB_CUDA(cudaMemcpy(m->arena, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
    const double *base = static_cast<const double *>(m->arena);
    ModelDev &v = m->dev;
    std::memset(&v, 0, sizeof(v));
    v.nz = nz; v.ny = ny; v.T = d->n_tracers; v.kind = d->kind; v.n_classes = d->n_classes;
    v.column_model = d->column_model;
    for (int t = 0; t < NKB_MAX_TRACERS; ++t) { v.class_of[t] = d->class_of[t]; v.src_const[t] = d->src_const[t]; }
    v.t0 = d->t0; v.t1 = d->t1;
    v.depth_edges = base + o_edges; v.depth_mid = base + o_mid; v.dz_r = base + o_dzr;
    v.dz_mid = base + o_dzm; v.dz_mid_r = base + o_dzmr; v.wvel = base + o_w;
    v.estencil = d->h_estencil ? base + o_est : nullptr;
    v.bld_max = d->h_bld_max ? base + o_bld : nullptr;
    for (int c = 0; c < NKB_MAX_CLASSES; ++c) {
        v.surf_diag[c] = d->surf_diag[c]; v.surf_aff[c] = d->surf_aff[c];
        v.decay[c] = d->decay[c]; v.sink_vel[c] = d->sink_vel[c];
    }
    v.n_flux_pts = d->n_flux_pts;
    NKB_REQUIRE(v.n_flux_pts >= 0 && v.n_flux_pts <= 8, "nkb_model_create: n_flux_pts > 8");
    for (int i = 0; i < 8; ++i) { v.flux_t[i] = d->flux_t[i]; v.flux_v[i] = d->flux_v[i]; }
    v.sink_thres = d->sink_thres;
    v.n_frc = d->n_frc;
    v.frc_time = d->n_frc > 0 ? base + o_ft : nullptr;
    v.frc_data = d->n_frc > 0 ? base + o_fd : nullptr;
    v.light = d->h_light ? base + o_light : nullptr;
    v.po4_halfsat = d->po4_halfsat; v.max_uptake_rate = d->max_uptake_rate; v.sigma = d->sigma;
    v.dop_remin_rate = d->dop_remin_rate; v.pop_remin_rate = d->pop_remin_rate;
    v.po4_s_restoring_opt = d->po4_s_restoring_opt;
    v.n_srf = d->n_srf;
    v.srf_time = d->n_srf > 0 ? base + o_st : nullptr;
    v.srf_data = d->n_srf > 0 ? base + o_sd : nullptr;
    for (int c = 0; c < NKB_MAX_CLASSES; ++c) v.srf_rate[c] = d->srf_rate[c];

    NKB_CUDA(cudaMalloc(&m->tri_raw, (size_t)v.n_classes * 4 * plane * sizeof(double)));
    NKB_CUDA(cudaMalloc(&m->aff_raw, (size_t)v.n_classes * ny * sizeof(double)));
    NKB_CUDA(cudaMalloc(&m->src_raw, 2 * plane * sizeof(double)));
    NKB_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    *out = m;
    return 0;
}

void nkb_model_destroy(nkb_model *m) {
    if (!m) return;
    cudaFree(m->arena); cudaFree(m->tri); cudaFree(m->aff); cudaFree(m->src); cudaFree(m->d_h);
    cudaFree(m->tri_raw); cudaFree(m->aff_raw); cudaFree(m->src_raw);
    cudaFree(m->ctab); cudaFree(m->d_done); cudaFree(m->d_hist_slot);
    if (m->h_err) cudaFreeHost(m->h_err);
    cudaFree(m->d_stage_major); cudaFree(m->d_stage_x); cudaFree(m->d_stage_f); cudaFree(m->d_stage_work);
    if (m->graph) cudaGraphExecDestroy(m->graph);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete[] m->h_t_start; delete[] m->h_h;
    delete m;
}

int nkb_model_set_schedule(nkb_model *m, int n_steps, const double *h_t_start, const double *h_h) {
    NKB_REQUIRE(m && n_steps >= 1 && h_t_start && h_h, "nkb_model_set_schedule: bad argument");
    const ModelDev &v = m->dev;
    const size_t plane = (size_t)v.nz * v.ny;
    const int n_stages = 2 * n_steps;
    cudaFree(m->tri); cudaFree(m->aff); cudaFree(m->src);
    m->tri = m->aff = m->src = nullptr;
    cudaFree(m->ctab);
    m->ctab = nullptr;
    if (m->graph) { cudaGraphExecDestroy(m->graph); m->graph = nullptr; }
    delete[] m->h_t_start; delete[] m->h_h;
    m->h_t_start = new double[n_steps]; m->h_h = new double[n_steps];
    std::memcpy(m->h_t_start, h_t_start, n_steps * sizeof(double));
    std::memcpy(m->h_h, h_h, n_steps * sizeof(double));
    m->n_steps = n_steps;

    std::vector<double> t_imp(n_stages), hg(n_stages), t_exp(n_stages);
    for (int n = 0; n < n_steps; ++n) {
        const double t = h_t_start[n], h = h_h[n];
        // implicit stage times c = (gamma, 1); the end of the last step is t1 exactly
        t_imp[2 * n] = t + nkb::kGamma * h;
        t_imp[2 * n + 1] = (n + 1 < n_steps) ? h_t_start[n + 1] : v.t1;
        hg[2 * n] = hg[2 * n + 1] = nkb::kGamma * h;
        // explicit stage times c = (0, gamma)
        t_exp[2 * n] = t;
        t_exp[2 * n + 1] = t + nkb::kGamma * h;
    }
    nkb::DevBuf<double> t_buf, hg_buf;  // released on every exit path
    NKB_CUDA(t_buf.alloc(n_stages));
    NKB_CUDA(hg_buf.alloc(n_stages));
    double *d_t = t_buf.p, *d_hg = hg_buf.p;
    NKB_CUDA(cudaMemcpy(d_t, t_imp.data(), n_stages * sizeof(double), cudaMemcpyHostToDevice));
    NKB_CUDA(cudaMemcpy(d_hg, hg.data(), n_stages * sizeof(double), cudaMemcpyHostToDevice));
    NKB_CUDA(cudaMalloc(&m->tri, (size_t)n_stages * v.n_classes * 4 * plane * sizeof(double)));
    NKB_CUDA(cudaMalloc(&m->aff, (size_t)n_stages * v.n_classes * v.ny * sizeof(double)));
    const int chunk = 32768;
    for (int s0 = 0; s0 < n_stages; s0 += chunk) {
        const int ns = (n_stages - s0 < chunk) ? n_stages - s0 : chunk;
        if (nkb::launch_stage_tables(v, ns, d_t + s0, d_hg + s0, 1,
                                     m->tri + (size_t)s0 * v.n_classes * 4 * plane,
                                     m->aff + (size_t)s0 * v.n_classes * v.ny, 0))
            return 1;
    }
    if (v.kind == NKB_MOD_FORCED_FILE) {
        NKB_CUDA(cudaMemcpy(d_t, t_exp.data(), n_stages * sizeof(double), cudaMemcpyHostToDevice));
        NKB_CUDA(cudaMalloc(&m->src, (size_t)n_stages * plane * sizeof(double)));
        for (int s0 = 0; s0 < n_stages; s0 += chunk) {
            const int ns = (n_stages - s0 < chunk) ? n_stages - s0 : chunk;
            if (nkb::launch_forcing_tables(v, ns, d_t + s0, m->src + (size_t)s0 * plane, 0)) return 1;  // [step][cell][2]
        }
    }
    cudaFree(m->d_h);
    m->d_h = nullptr;
    NKB_CUDA(cudaMalloc(&m->d_h, n_steps * sizeof(double)));
    NKB_CUDA(cudaMemcpy(m->d_h, h_h, n_steps * sizeof(double), cudaMemcpyHostToDevice));
    NKB_CUDA(cudaDeviceSynchronize());
    return 0;
}

// Coefficient table of the fused step kernel, built on the first evaluation that uses it.
static int ensure_fused_tables(nkb_model *m) {
    if (m->ctab) return 0;
    const ModelDev &v = m->dev;
    const int n_steps = m->n_steps, n_stages = 2 * n_steps;
    std::vector<double> t_imp(n_stages), hg(n_stages), t_exp(n_stages);
    for (int n = 0; n < n_steps; ++n) {  // same stage times as nkb_model_set_schedule
        const double t = m->h_t_start[n], h = m->h_h[n];
        t_imp[2 * n] = t + nkb::kGamma * h;
        t_imp[2 * n + 1] = (n + 1 < n_steps) ? m->h_t_start[n + 1] : v.t1;
        hg[2 * n] = hg[2 * n + 1] = nkb::kGamma * h;
        t_exp[2 * n] = t;
        t_exp[2 * n + 1] = t + nkb::kGamma * h;
    }
    double *d_t = nullptr, *d_hg = nullptr, *d_te = nullptr;
    NKB_CUDA(cudaMalloc(&d_t, n_stages * sizeof(double)));
    NKB_CUDA(cudaMalloc(&d_hg, n_stages * sizeof(double)));
    NKB_CUDA(cudaMalloc(&d_te, n_stages * sizeof(double)));
    NKB_CUDA(cudaMemcpy(d_t, t_imp.data(), n_stages * sizeof(double), cudaMemcpyHostToDevice));
    NKB_CUDA(cudaMemcpy(d_hg, hg.data(), n_stages * sizeof(double), cudaMemcpyHostToDevice));
    NKB_CUDA(cudaMemcpy(d_te, t_exp.data(), n_stages * sizeof(double), cudaMemcpyHostToDevice));
    const size_t per_step = (size_t)v.n_classes * 8 * v.nz * 2 * (v.ny + 1);
    NKB_CUDA(cudaMalloc(&m->ctab, (size_t)n_steps * per_step * sizeof(double)));
    NKB_CUDA(cudaMemset(m->ctab, 0, (size_t)n_steps * per_step * sizeof(double)));
    const int chunk = 32768;
    for (int s0 = 0; s0 < n_steps; s0 += chunk) {
        const int ns = (n_steps - s0 < chunk) ? n_steps - s0 : chunk;
        if (nkb::launch_step_ctab(v, ns, d_t + 2 * s0, d_hg + 2 * s0, d_te + 2 * s0, m->ctab + (size_t)s0 * per_step, 0))
            return 1;
    }
    if (nkb::fused_encode_ctab_map(v.nz, v.ny, (size_t)n_steps * v.n_classes * 8, m->ctab, &m->map_ctab, v.kind)) return 1;
    NKB_CUDA(cudaDeviceSynchronize());
    cudaFree(d_t); cudaFree(d_hg); cudaFree(d_te);
    return 0;
}

int nkb_model_mixing_coeff(nkb_model *m, double time, double *d_out, void *stream) {
    NKB_REQUIRE(m && d_out, "nkb_model_mixing_coeff: null argument");
    return nkb::launch_mixing_coeff(m->dev, time, d_out, (cudaStream_t)stream);
}

static void fill_args(const nkb_model *m, StageArgs &a, int B, int ldb) {
    const ModelDev &v = m->dev;
    std::memset(&a, 0, sizeof(a));
    a.nz = v.nz; a.ny = v.ny; a.B = B; a.ldb = ldb; a.T = v.T;
    a.est = v.estencil; a.light = v.light;
    for (int t = 0; t < NKB_MAX_TRACERS; ++t) { a.class_of[t] = v.class_of[t]; a.src_const[t] = v.src_const[t]; }
    a.sink_thres_r = v.sink_thres > 0.0 ? 1.0 / v.sink_thres : 0.0;
    a.halfsat = v.po4_halfsat; a.umax = v.max_uptake_rate; a.sigma = v.sigma;
    a.rdop = v.dop_remin_rate; a.rpop = v.pop_remin_rate;
    a.restoring_opt = v.po4_s_restoring_opt;
}

int nkb_model_tend(nkb_model *m, double time, const double *d_x, double *d_tend, int B, int ldb, void *stream) {
    NKB_REQUIRE(m && d_x && d_tend && B >= 1 && ldb >= B, "nkb_model_tend: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const ModelDev &v = m->dev;
    nkb::DevBuf<double> t_buf;  // released on every exit path
    NKB_CUDA(t_buf.alloc(2));
    double *d_t = t_buf.p;
    const double th[2] = {time, time};
    NKB_CUDA(cudaMemcpyAsync(d_t, th, 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nkb::launch_stage_tables(v, 1, d_t, d_t + 1, 0, m->tri_raw, m->aff_raw, st)) return 1;
    if (v.kind == NKB_MOD_FORCED_FILE && nkb::launch_forcing_tables(v, 2, d_t, m->src_raw, st)) return 1;
    StageArgs a;
    fill_args(m, a, B, ldb);
    a.u[0] = d_x; a.out = d_tend; a.tri = m->tri_raw; a.aff = m->aff_raw; a.src2 = m->src_raw;
    const int rc = nkb::launch_tend(v.kind, a, st);
    NKB_CUDA(cudaStreamSynchronize(st));
    return rc;
}

size_t nkb_model_work_doubles(const nkb_model *m, int B, int ldb) {
    const size_t n1 = (size_t)m->dev.T * m->dev.nz * m->dev.ny;
    // a single state is staged into a 4-lane batch for the fused step kernels: x, f and two work copies
    if (B == 1 && ldb == 1) return 16 * n1;
    return 2 * n1 * (size_t)ldb;
}

int nkb_model_eval(nkb_model *m, const double *d_x0, double *d_f, double *d_work, int B, int ldb, int n_hist,
                   const int *h_hist_steps, double *d_hist, void *stream) {
    NKB_REQUIRE(m && d_x0 && d_f && d_work, "nkb_model_eval: null argument");
    NKB_REQUIRE(m->n_steps > 0 && m->tri, "nkb_model_eval: call nkb_model_set_schedule first");
    NKB_REQUIRE(B >= 1 && ldb >= B, "nkb_model_eval: need 1 <= B <= ldb");
    NKB_REQUIRE(B == 1 || ldb % 2 == 0, "nkb_model_eval: ldb must be even");
    cudaStream_t st = (cudaStream_t)stream;
    const ModelDev &v = m->dev;
    const size_t plane = (size_t)v.nz * v.ny;
    if (B == 1 && ldb == 1 && nkb::fused_single_state(v)) {
        // the reference's own layout [tracer, depth, ypos]: stage it as member 0 of a 4-lane batch
        const size_t n1 = (size_t)v.T * plane, ls = 4;
        double *xs = d_work, *fs = d_work + n1 * ls, *wk = d_work + 2 * n1 * ls;
        if (nkb::launch_scatter_member(d_x0, xs, n1, ls, st)) return 1;
        if (nkb_model_eval(m, xs, fs, wk, 1, (int)ls, n_hist, h_hist_steps, d_hist, stream)) return 1;
        return nkb::launch_gather_member(fs, d_f, n1, ls, 0, st);
    }
    const size_t nstate = (size_t)v.T * plane * ldb;
    double *w_u1 = d_work, *w_alt = d_work + nstate;
    const int S = m->n_steps;
    const size_t tri_stride = (size_t)v.n_classes * 4 * plane, aff_stride = (size_t)v.n_classes * v.ny;
    const double a1 = (1.0 - nkb::kGamma) / nkb::kGamma, a0 = 1.0 - a1;

    if (v.column_model == 1 && v.ny == 1) {
        // test_problem: the whole year in one persistent kernel, state resident in shared memory
        nkb::ColumnArgs c;
        std::memset(&c, 0, sizeof(c));
        c.x0 = d_x0; c.out = d_f; c.tri = m->tri; c.aff = m->aff; c.h = m->d_h; c.light = v.light;
        c.nz = v.nz; c.B = B; c.ldb = ldb; c.T = v.T; c.n_steps = S; c.ncls = v.n_classes;
        for (int t = 0; t < NKB_MAX_TRACERS; ++t) { c.class_of[t] = v.class_of[t]; c.src_const[t] = v.src_const[t]; }
        c.restoring_opt = v.po4_s_restoring_opt;
        if (n_hist > 0) {
            std::vector<int> slot(S + 1, -1);
            for (int i = 0; i < n_hist; ++i) {
                NKB_REQUIRE(h_hist_steps[i] >= 0 && h_hist_steps[i] <= S, "nkb_model_eval: hist step out of range");
                slot[h_hist_steps[i]] = i;
            }
            // the slot table lives with the model (no allocation, no host synchronisation per evaluation: the
            // two legs of a Richardson pair run concurrently on two streams)
            if ((size_t)(S + 1) > m->hist_slot_cap) {
                cudaFree(m->d_hist_slot);
                m->d_hist_slot = nullptr;
                NKB_CUDA(cudaMalloc(&m->d_hist_slot, (S + 1) * sizeof(int)));
                m->hist_slot_cap = (size_t)(S + 1);
            }
            NKB_CUDA(cudaMemcpyAsync(m->d_hist_slot, slot.data(), (S + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
            c.hist_slot = m->d_hist_slot;
            c.hist = d_hist;
        }
        return nkb::launch_column_year(v.kind, c, st);
    }

    int hist_i = 0;
    auto emit_hist = [&](int step, const double *state) -> int {
        while (hist_i < n_hist && h_hist_steps[hist_i] == step) {
            if (nkb::launch_gather_member(state, d_hist + (size_t)hist_i * v.T * plane, (size_t)v.T * plane, ldb, 0, st))
                return 1;
            ++hist_i;
        }
        return 0;
    };
    if (emit_hist(0, d_x0)) return 1;

    if (nkb::fused_step_usable(v, B, ldb, d_x0, d_f, d_work)) {
        // fused step kernel (nkb_step_fused.cu): both implicit stages of a time step in one pass; by
        // default ONE persistent launch integrates all steps between two hist snapshots (CTAs wait
        // on per-tile step counters instead of kernel boundaries), NKB_FUSED_PERSIST=0: one launch
        // per step
        if (ensure_fused_tables(m)) return 1;
        if (m->h_err && *m->h_err) {
            *m->h_err = 0;
            nkb::set_error("nkb_model_eval: a previous persistent step launch timed out waiting for a tile");
            return 1;
        }
        if (!m->h_err) {
            NKB_CUDA(cudaHostAlloc(&m->h_err, sizeof(int), cudaHostAllocMapped));
            *m->h_err = 0;
            NKB_CUDA(cudaHostGetDevicePointer(&m->d_err, m->h_err, 0));
        }
        const size_t ntiles = (size_t)nkb::fused_tile_count(v, B);
        if (ntiles > m->done_cap) {
            cudaFree(m->d_done);
            m->d_done = nullptr;
            NKB_CUDA(cudaMalloc(&m->d_done, ntiles * sizeof(int)));
            m->done_cap = ntiles;
        }
        NKB_CUDA(cudaMemsetAsync(m->d_done, 0, ntiles * sizeof(int), st));
        nkb::FusedMaps fm;
        // buffers the steps alternate between: the last step writes buf_f.  Phosphorus integrates in a
        // member-block-major copy (nkb_step_fused.cu): x0 is converted into the work buffer that step 0
        // does not write, the result is converted back (minus x0) at the end.
        const bool tm = nkb::fused_tile_major(v);
        double *buf_f = tm ? w_alt : d_f, *buf_w = tm ? w_u1 : w_alt;
        const double *buf_x0 = d_x0;
        if (tm) {
            double *x0_tm = (((S - 1) & 1) == 0) ? buf_w : buf_f;
            if (nkb::launch_p3_to_tm(v, d_x0, x0_tm, B, ldb, st)) return 1;
            buf_x0 = x0_tm;
        }
        if (nkb::fused_encode_state_maps(v, B, ldb, buf_x0, &fm.in_x0, nullptr)) return 1;
        if (nkb::fused_encode_state_maps(v, B, ldb, buf_f, &fm.in_f, &fm.out_f)) return 1;
        if (nkb::fused_encode_state_maps(v, B, ldb, buf_w, &fm.in_w, &fm.out_w)) return 1;
        // encoded per evaluation: the box shape follows the thread layout in use (NKB_FUSED_MPT)
        if (nkb::fused_encode_ctab_map(v.nz, v.ny, (size_t)S * v.n_classes * 8, m->ctab, &fm.ctab, v.kind)) return 1;
        bool persist = nkb::fused_persistent();
        int n = 0;
        while (n < S) {
            // the segment ends after the next step whose result is a hist snapshot (or at S)
            int end = S;
            if (!persist) end = n + 1;
            else if (hist_i < n_hist && h_hist_steps[hist_i] < S) end = h_hist_steps[hist_i] > n ? h_hist_steps[hist_i] : n + 1;
            const int rc = nkb::launch_steps_fused(v, B, S, n, end, m->d_h, m->aff, fm, m->d_done, m->d_err, st);
            if (rc == -1) {  // cooperative launch refused (SMs in use by someone else): one launch per step
                persist = false;
                continue;
            }
            if (rc) return 1;
            n = end;
            if (n_hist > 0 && n < S) {
                double *dest = (((S - n) & 1) == 0) ? buf_f : buf_w;  // buffer written by step n - 1
                if (tm) {
                    while (hist_i < n_hist && h_hist_steps[hist_i] == n) {
                        if (nkb::launch_p3_gather_member(v, dest, d_hist + (size_t)hist_i * v.T * plane, B, 0, st)) return 1;
                        ++hist_i;
                    }
                } else if (emit_hist(n, dest)) {
                    return 1;
                }
            }
        }
        if (tm) {
            if (nkb::launch_p3_from_tm_sub(v, buf_f, d_x0, d_f, B, ldb, st)) return 1;
        } else if (nkb::launch_sub_inplace(d_f, d_x0, nstate, st)) {
            return 1;
        }
        NKB_CUDA(cudaGetLastError());
        return 0;
    }

    StageArgs a;
    fill_args(m, a, B, ldb);
    const double *un = d_x0;
    for (int n = 0; n < S; ++n) {
        const double h = m->h_h[n];
        const bool last = (n == S - 1);
        // destination of this step's result alternates so that the last step lands in d_f
        double *dest = (((S - 1 - n) & 1) == 0) ? d_f : w_alt;
        // stage 1: (I - h*gamma*L(t_n + gamma h)) u1 = u_n + h*gamma*E(t_n, u_n)
        a.u[0] = un; a.u[1] = nullptr; a.out = w_u1; a.sub = nullptr;
        a.a[0] = 1.0; a.he[0] = nkb::kGamma * h; a.a[1] = 0.0; a.he[1] = 0.0;
        a.tri = m->tri + (size_t)(2 * n) * tri_stride;
        a.aff = m->aff + (size_t)(2 * n) * aff_stride;
        a.src2 = m->src ? m->src + (size_t)(2 * n) * plane : nullptr;
        if (nkb::launch_stage(v.kind, 1, a, st)) return 1;
        // stage 2: (I - h*gamma*L(t_n + h)) u2 = a0 u_n + a1 u1 + h(delta-1+gamma) E_n + h(1-delta) E(u1)
        a.u[0] = un; a.u[1] = w_u1; a.out = dest; a.sub = last ? d_x0 : nullptr;
        a.a[0] = a0; a.a[1] = a1;
        a.he[0] = h * (nkb::kDelta - 1.0 + nkb::kGamma); a.he[1] = h * (1.0 - nkb::kDelta);
        a.tri = m->tri + (size_t)(2 * n + 1) * tri_stride;
        a.aff = m->aff + (size_t)(2 * n + 1) * aff_stride;
        if (nkb::launch_stage(v.kind, 2, a, st)) return 1;
        un = dest;
        if (n_hist > 0 && !last && emit_hist(n + 1, dest)) return 1;
    }
    NKB_CUDA(cudaGetLastError());
    // the final state (not the difference) for hist: x(T) = f + x0 is assembled by the host
    return 0;
}

int nkb_model_poll_error(nkb_model *m) {
    if (!m || !m->h_err) return 0;
    const int e = *reinterpret_cast<volatile int *>(m->h_err);
    if (e) *m->h_err = 0;
    return e ? 1 : 0;
}

int nkb_model_eval_host(nkb_model *m, const double *h_x0, double *h_f, int B) {
    NKB_REQUIRE(m && h_x0 && h_f && B >= 1, "nkb_model_eval_host: bad argument");
    const ModelDev &v = m->dev;
    const size_t n = (size_t)v.T * v.nz * v.ny;
    const int ldb = (B == 1) ? 1 : ((B + 31) / 32) * 32;
    const size_t need = n * (size_t)ldb;
    if (need > m->stage_cap) {
        cudaFree(m->d_stage_major); cudaFree(m->d_stage_x); cudaFree(m->d_stage_f); cudaFree(m->d_stage_work);
        m->d_stage_major = m->d_stage_x = m->d_stage_f = m->d_stage_work = nullptr;
        m->stage_cap = 0;
        // every buffer is sized for the PADDED member count: a later call with another B that rounds to the
        // same ldb reuses them (the capacity test above is on n * ldb)
        NKB_CUDA(cudaMalloc(&m->d_stage_major, need * sizeof(double)));
        NKB_CUDA(cudaMalloc(&m->d_stage_x, need * sizeof(double)));
        NKB_CUDA(cudaMalloc(&m->d_stage_f, need * sizeof(double)));
        NKB_CUDA(cudaMalloc(&m->d_stage_work, nkb_model_work_doubles(m, ldb, ldb) * sizeof(double)));
        // ordered before the pack kernel: own_stream is non-blocking, so a legacy-stream memset would not be
        NKB_CUDA(cudaMemsetAsync(m->d_stage_x, 0, need * sizeof(double), m->own_stream));
        m->stage_cap = need;
    }
    cudaStream_t st = m->own_stream;
    NKB_CUDA(cudaMemcpyAsync(m->d_stage_major, h_x0, n * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nkb::launch_pack(m->d_stage_major, m->d_stage_x, (int)n, B, ldb, st)) return 1;
    if (nkb_model_eval(m, m->d_stage_x, m->d_stage_f, m->d_stage_work, B, ldb, 0, nullptr, nullptr, st)) return 1;
    if (nkb::launch_unpack(m->d_stage_f, m->d_stage_major, (int)n, B, ldb, st)) return 1;
    NKB_CUDA(cudaMemcpyAsync(h_f, m->d_stage_major, n * (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
    NKB_CUDA(cudaStreamSynchronize(st));
    if (nkb_model_poll_error(m)) {
        nkb::set_error("nkb_model_eval_host: the persistent step kernel timed out waiting for a tile");
        return 1;
    }
    return 0;
}

}  // extern "C"
