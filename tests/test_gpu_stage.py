"""GPU parity tests of the fused stage kernel / coefficient tables against the oracle.

All calls go through the C ABI (ctypes).  Tolerances: the CUDA path and the numpy
restatement of the SAME scheme (oracle/imex_oracle.py) agree to rounding error
(rtol 1e-11); the tendency agrees with the reference's comp_tend restatement
(oracle/nk_oracle.py) to rtol 1e-12 of the field maximum.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _grid(nz, ny, ratio=19.0, vvel=0.1, kh=1000.0):
    from oracle import nk_oracle as o
    from nk_ooc_b200.py_driver_2d.modules import Transport2D
    from nk_ooc_b200.spatial_axis import SpatialAxis

    ze = o.stretched_edges(nz, 0.0, 4000.0, ratio)
    ye = o.stretched_edges(ny, 0.0, 50.0e5, 1.0)
    g = o.Grid2D(ze, ye, vvel, kh)
    tr = Transport2D(SpatialAxis("depth", ze), SpatialAxis("ypos", ye), vvel, kh)
    return g, tr


def _to_dev(x):  # [T, nz, ny, B] host -> padded member-fastest device tensor
    from nk_ooc_b200.engine import padded_members

    B = x.shape[-1]
    ldb = padded_members(B)
    out = torch.zeros(x.shape[:-1] + (ldb,), dtype=torch.float64, device="cuda")
    out[..., :B] = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return out


def _forcing(g, rng):
    nt = 13
    times = np.linspace(0.0, 365.0 * 86400.0, nt)
    data = -1.0e-8 * np.abs(rng.normal(size=(nt, g.nz, g.ny)))
    return times, data


@pytest.mark.parametrize("nz,ny", [(12, 9), (30, 30)])
def test_mixing_coeff_matches_oracle(nz, ny):
    from nk_ooc_b200.py_driver_2d.modules import iage_model

    g, tr = _grid(nz, ny)
    m = iage_model(tr)
    for frac in [0.0, 0.26, 0.3, 0.349, 0.5, 0.7, 0.99]:
        t = frac * 365.0 * 86400.0
        got = m.mixing_coeff(t).cpu().numpy()
        want = g.vert_mixing_coeff(t)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("kind", ["iage", "phosphorus", "forced"])
@pytest.mark.parametrize("B", [1, 3, 40])
def test_tend_matches_reference_comp_tend(kind, B):
    from oracle import nk_oracle as o
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(5)
    g, tr = _grid(14, 11)
    if kind == "iage":
        om, m = o.Iage2D(g), modules.iage_model(tr)
    elif kind == "phosphorus":
        om, m = o.Phosphorus2D(g), modules.phosphorus_model(tr)
    else:
        times, data = _forcing(g, rng)
        om = o.Forced2D(g, restore_rate_10m=1.0 / 3600.0, restore_const=1.0, sms_opt="file", sms_times=times,
                        sms_data=data, sink_thres=0.05)
        m = modules.forced_model(tr, "const", 1.0, 1.0 / 3600.0, "file", sms_times=times, sms_data=data,
                                 sink_thres=0.05)
    x = np.abs(rng.normal(size=(om.tracer_cnt, g.nz, g.ny, B))) * 0.08
    for t in [0.0, 0.31 * 365 * 86400.0, 0.8 * 365 * 86400.0]:
        got = m.tend(t, _to_dev(x), B).cpu().numpy()[..., :B]
        want = np.stack([om.comp_tend(t, x[..., b].reshape(-1)).reshape(x.shape[:-1]) for b in range(B)], axis=-1)
        scale = np.abs(want).max()
        np.testing.assert_allclose(got, want, rtol=0.0, atol=1e-12 * scale)


@pytest.mark.parametrize("kind", ["iage", "phosphorus", "forced"])
@pytest.mark.parametrize("B", [1, 5, 70])
def test_model_year_matches_scheme_oracle(kind, B):
    """same ARS(2,2,2) schedule in numpy and in CUDA: agreement to rounding"""
    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(11)
    g, tr = _grid(10, 7)
    # phosphorus: the explicit uptake term (1/(3 days)) needs h well below 3 days, otherwise the
    # scheme is unstable and rounding differences are amplified
    nsteps = 240 if kind == "phosphorus" else 24
    if kind == "iage":
        mod, m = im.Module2D("iage", g), modules.iage_model(tr)
    elif kind == "phosphorus":
        mod, m = im.Module2D("phosphorus", g, phos=o.Phosphorus2D(g)), modules.phosphorus_model(tr)
    else:
        times, data = _forcing(g, rng)
        f = o.Forced2D(g, restore_rate_10m=1.0 / 3600.0, restore_const=1.0, sms_opt="file", sms_times=times,
                       sms_data=data, sink_thres=0.05)
        mod = im.Module2D("forced", g, forced=f)
        m = modules.forced_model(tr, "const", 1.0, 1.0 / 3600.0, "file", sms_times=times, sms_data=data,
                                 sink_thres=0.05)
    x = np.abs(rng.normal(size=(mod.T, g.nz, g.ny, B))) * 0.5
    m.set_uniform_schedule(nsteps)
    got = m.eval(_to_dev(x), B).cpu().numpy()[..., :B]
    want = im.model_year_2d(mod, x, nsteps)
    np.testing.assert_allclose(got, want, rtol=0.0, atol=1e-9 * np.abs(want).max())


def test_hist_snapshots_and_host_path():
    from oracle import imex_oracle as im
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(3)
    g, tr = _grid(10, 7)
    m = modules.iage_model(tr)
    m.set_uniform_schedule(12)
    B = 6
    x = rng.normal(size=(2, g.nz, g.ny, B))
    snaps = []
    want = im.model_year_2d(im.Module2D("iage", g), x, 12, snapshots=snaps)
    f, hist = m.eval(_to_dev(x), B, hist_steps=[0, 4, 8, 12])
    hist = hist.cpu().numpy()
    np.testing.assert_allclose(hist[0], x[..., 0], rtol=0, atol=0)
    np.testing.assert_allclose(hist[1], snaps[3][1][..., 0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(hist[2], snaps[7][1][..., 0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(hist[3], x[..., 0] + want[..., 0], rtol=0, atol=1e-12)
    # host-buffer entry point (member-major layout)
    xm = np.ascontiguousarray(np.moveaxis(x, -1, 0))
    fh = m.eval_host(xm).numpy()
    np.testing.assert_allclose(np.moveaxis(fh, 0, -1), want, rtol=0, atol=1e-11 * np.abs(want).max())


@pytest.mark.parametrize("kind", ["iage", "forced", "iage_columns"])
@pytest.mark.parametrize("nz,ny,B", [(21, 33, 20), (37, 15, 50), (128, 16, 9), (10, 7, 70)])
def test_fused_step_kernel_matches_scheme_oracle(kind, nz, ny, B, monkeypatch):
    """the fused step kernel (one launch per time step: TMEM-resident intermediates, TMA-fed
    sweeps; nkb_step_fused.cu) against the numpy statement of the scheme, and against the
    stage-per-launch path.  Grids cover several column tiles (ny > 14), a partial last tile,
    a level count that is not a multiple of the chunk (4) and the TMEM capacity limit (128)."""
    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200 import _lib
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(21)
    if kind == "iage_columns":
        g, tr = _grid(nz, ny, vvel=0.0, kh=0.0)
    else:
        g, tr = _grid(nz, ny)
    nsteps = 6
    if kind in ("iage", "iage_columns"):
        mod, m = im.Module2D("iage", g), modules.iage_model(tr)
    else:
        times, data = _forcing(g, rng)
        f = o.Forced2D(g, restore_rate_10m=1.0 / 3600.0, restore_const=1.0, sms_opt="file", sms_times=times,
                       sms_data=data, sink_thres=0.05)
        mod = im.Module2D("forced", g, forced=f)
        m = modules.forced_model(tr, "const", 1.0, 1.0 / 3600.0, "file", sms_times=times, sms_data=data,
                                 sink_thres=0.05)
    x = np.abs(rng.normal(size=(mod.T, g.nz, g.ny, B))) * 0.5
    m.set_uniform_schedule(nsteps)
    lib = _lib.load()
    xd = _to_dev(x)
    monkeypatch.setenv("NKB_FUSED", "1")
    m.eval(xd, B)  # the first fused evaluation also builds the step tables
    n0 = lib.nkb_launch_count()
    got = m.eval(xd, B).cpu().numpy()[..., :B]
    assert lib.nkb_launch_count() - n0 == 2, "the persistent fused step kernel did not run"
    # one launch per step instead of one persistent launch with per-tile dependencies: same bits
    monkeypatch.setenv("NKB_FUSED_PERSIST", "0")
    n0 = lib.nkb_launch_count()
    per_step = m.eval(xd, B).cpu().numpy()[..., :B]
    assert lib.nkb_launch_count() - n0 == nsteps + 1
    np.testing.assert_array_equal(got, per_step)
    monkeypatch.delenv("NKB_FUSED_PERSIST")
    monkeypatch.setenv("NKB_FUSED", "0")
    n0 = lib.nkb_launch_count()
    unfused = m.eval(xd, B).cpu().numpy()[..., :B]
    assert lib.nkb_launch_count() - n0 == 2 * nsteps
    want = im.model_year_2d(mod, x, nsteps)
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=0.0, atol=1e-10 * scale)
    np.testing.assert_allclose(got, unfused, rtol=0.0, atol=1e-10 * scale)


@pytest.mark.parametrize("nz,ny,B", [(21, 33, 20), (37, 15, 50), (125, 16, 9), (10, 7, 70), (40, 50, 13)])
def test_fused_phosphorus_step_kernel_matches_scheme_oracle(nz, ny, B, monkeypatch):
    """three coupled tracers per tile (step_fused_p3_kernel: 4-member rows, cross-tracer exchange of
    the stage-1 solution between warps) against the numpy statement of the scheme and against the
    stage-per-launch kernels; persistent launch and one launch per step give the same bits.  Member
    counts that are not multiples of 4, several column tiles, level counts that are not multiples
    of the chunk, the refined grid's 125 levels."""
    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200 import _lib
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(31)
    g, tr = _grid(nz, ny)
    nsteps = 240  # the explicit uptake term (1/(3 days)) needs h well below 3 days
    mod, m = im.Module2D("phosphorus", g, phos=o.Phosphorus2D(g)), modules.phosphorus_model(tr)
    x = np.abs(rng.normal(size=(3, g.nz, g.ny, B))) * 0.5
    m.set_uniform_schedule(nsteps)
    lib = _lib.load()
    xd = _to_dev(x)
    monkeypatch.setenv("NKB_FUSED_P3", "1")
    m.eval(xd, B)
    n0 = lib.nkb_launch_count()
    got = m.eval(xd, B).cpu().numpy()[..., :B]
    # layout conversion in, ONE persistent step launch, conversion out (minus x0)
    assert lib.nkb_launch_count() - n0 == 3, "the persistent fused phosphorus kernel did not run"
    m.check_health()
    monkeypatch.setenv("NKB_FUSED_PERSIST", "0")
    per_step = m.eval(xd, B).cpu().numpy()[..., :B]
    np.testing.assert_array_equal(got, per_step)
    monkeypatch.delenv("NKB_FUSED_PERSIST")
    monkeypatch.setenv("NKB_FUSED_P3", "0")
    n0 = lib.nkb_launch_count()
    unfused = m.eval(xd, B).cpu().numpy()[..., :B]
    assert lib.nkb_launch_count() - n0 == 2 * nsteps
    want = im.model_year_2d(mod, x, nsteps)
    scale = np.abs(want).max()
    np.testing.assert_allclose(unfused, want, rtol=0.0, atol=1e-10 * scale)
    np.testing.assert_allclose(got, want, rtol=0.0, atol=1e-10 * scale)


def test_fused_kernel_hist_snapshots_and_two_members_per_thread(monkeypatch):
    """B >= 8: hist snapshots cut the persistent launch into segments (one cooperative launch per
    interval between snapshots); the alternative thread layout (NKB_FUSED_MPT=2: 4 consumer warps x 2
    members, 4-level chunks) gives the same result to rounding"""
    from oracle import imex_oracle as im
    from nk_ooc_b200 import _lib
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(31)
    g, tr = _grid(13, 19)
    m = modules.iage_model(tr)
    m.set_uniform_schedule(12)
    B = 21
    x = rng.normal(size=(2, g.nz, g.ny, B))
    snaps = []
    want = im.model_year_2d(im.Module2D("iage", g), x, 12, snapshots=snaps)
    lib = _lib.load()
    xd = _to_dev(x)
    m.eval(xd, B)
    n0 = lib.nkb_launch_count()
    f, hist = m.eval(xd, B, hist_steps=[0, 4, 8, 12])
    # gathers of member 0 (4) + persistent segments [0,4) [4,8) [8,12) + final difference
    assert lib.nkb_launch_count() - n0 <= 4 + 3 + 1
    hist = hist.cpu().numpy()
    scale = np.abs(want).max()
    np.testing.assert_allclose(hist[0], x[..., 0], rtol=0, atol=0)
    np.testing.assert_allclose(hist[1], snaps[3][1][..., 0], rtol=0, atol=1e-11 * scale)
    np.testing.assert_allclose(hist[2], snaps[7][1][..., 0], rtol=0, atol=1e-11 * scale)
    np.testing.assert_allclose(f.cpu().numpy()[..., :B], want, rtol=0, atol=1e-10 * scale)
    monkeypatch.setenv("NKB_FUSED_MPT", "2")
    m2 = modules.iage_model(tr)  # box shapes depend on the layout: fresh tables
    m2.set_uniform_schedule(12)
    got2 = m2.eval(xd, B).cpu().numpy()[..., :B]
    np.testing.assert_allclose(got2, want, rtol=0, atol=1e-10 * scale)



def test_fused_phosphorus_hist_snapshots_odd_and_even_step_counts():
    """hist snapshots of the phosphorus step kernel are gathered from the member-block-major work
    buffers (segmented persistent launches); an odd and an even number of steps put x(0)'s copy and
    the final state in different work buffers; a member count that is not a multiple of 4"""
    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(41)
    g, tr = _grid(13, 19)
    mod = im.Module2D("phosphorus", g, phos=o.Phosphorus2D(g))
    B = 10
    x = np.abs(rng.normal(size=(3, g.nz, g.ny, B))) * 0.5
    for nsteps, marks in ((240, [0, 80, 81, 240]), (241, [0, 7, 240, 241])):
        m = modules.phosphorus_model(tr)
        m.set_uniform_schedule(nsteps)
        snaps = []
        want = im.model_year_2d(mod, x, nsteps, snapshots=snaps)
        f, hist = m.eval(_to_dev(x), B, hist_steps=marks)
        m.check_health()
        hist = hist.cpu().numpy()
        scale = np.abs(want).max()
        np.testing.assert_allclose(f.cpu().numpy()[..., :B], want, rtol=0, atol=1e-10 * scale)
        np.testing.assert_array_equal(hist[0], x[..., 0])
        for i, step in enumerate(marks[1:-1], start=1):
            np.testing.assert_allclose(hist[i], snaps[step - 1][1][..., 0], rtol=0, atol=1e-10 * scale)
        np.testing.assert_allclose(hist[-1], x[..., 0] + want[..., 0], rtol=0, atol=1e-10 * scale)


def test_full_size_properties_refined_grid():
    """BASELINE.json's headline size (refined 125 x 150 grid, 4096 members; 8 steps instead of a year)
    through size-independent properties, since the numpy oracle takes minutes there:
    (i) members are independent — a permutation of the members permutes the results, bit for bit;
    (ii) the linear module is affine — F(a x + b y) - F(0) = a (F(x) - F(0)) + b (F(y) - F(0));
    (iii) the persistent fused kernel, one launch per step and the stage-per-launch kernels agree;
    (iv) one member against the numpy statement of the scheme."""
    import os

    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200.py_driver_2d import modules
    from nk_ooc_b200.spatial_axis import SpatialAxis, edges_from_defn

    nz, ny, B, nsteps = 125, 150, 4096, 8
    ze = edges_from_defn(nz, 0.0, 4000.0, 11.8)
    ye = edges_from_defn(ny, 0.0, 50.0e5, 1.0)
    tr = modules.Transport2D(SpatialAxis("depth", ze), SpatialAxis("ypos", ye), 0.1, 1000.0)
    m = modules.iage_model(tr)
    # the first 8 steps of the production schedule (h = 1/1200 yr): stable for non-smooth random states
    from nk_ooc_b200.engine import graded_schedule

    t_all, h_all = graded_schedule(0.0, 365.0 * 86400.0)
    sched = (t_all[:nsteps].copy(), h_all[:nsteps].copy())
    m.set_schedule(*sched)
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((2, nz, ny, B), dtype=torch.float64, device="cuda", generator=gen)
    f = m.eval(x, B).clone()
    # (i)
    perm = torch.randperm(B, device="cuda", generator=gen)
    fp = m.eval(x[..., perm].contiguous(), B)
    assert torch.equal(fp, f[..., perm])
    # (ii) members 0..B/2-1 = x, B/2.. = y; third batch = a x + b y; F(0) from a zero batch
    h = B // 2
    a, b = 0.75, -1.5
    z = torch.zeros_like(x)
    z[..., :h] = a * x[..., :h] + b * x[..., h:]
    fz = m.eval(z, B)
    f0 = fz[..., h:h + 1]  # members h.. of z are zero states
    lhs = fz[..., :h] - f0
    rhs = a * (f[..., :h] - f0) + b * (f[..., h:] - f0)
    scale = float(f.abs().max())
    assert float((lhs - rhs).abs().max()) <= 1e-11 * scale
    # (iii)
    os.environ["NKB_FUSED_PERSIST"] = "0"
    try:
        assert torch.equal(m.eval(x, B), f)
    finally:
        del os.environ["NKB_FUSED_PERSIST"]
    os.environ["NKB_FUSED"] = "0"
    try:
        fu = m.eval(x, B)
    finally:
        del os.environ["NKB_FUSED"]
    assert float((fu - f).abs().max()) <= 1e-11 * scale
    # (iv)
    g = o.Grid2D(ze, ye, 0.1, 1000.0)
    want = im.model_year_2d(im.Module2D("iage", g), x[..., 5:6].cpu().numpy(), schedule=sched)
    np.testing.assert_allclose(f[..., 5:6].cpu().numpy(), want, rtol=0, atol=1e-11 * np.abs(want).max())


def test_full_size_properties_refined_grid_phosphorus():
    """the three-tracer module at BASELINE.json's headline size (refined 125 x 150, 4096 members; the
    first 8 steps of the production schedule): members are independent bit for bit (a permutation of
    the members permutes the results — members share 4-member tiles with different neighbours), the
    persistent launch equals one launch per step, the stage-per-launch kernels agree to rounding, one
    member against the numpy statement of the scheme, and total phosphorus (po4 + dop + pop, volume
    weighted) is conserved by the step kernels as it is by the reference's tendencies"""
    import os

    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200.engine import graded_schedule
    from nk_ooc_b200.py_driver_2d import modules
    from nk_ooc_b200.spatial_axis import SpatialAxis, edges_from_defn

    nz, ny, B, nsteps = 125, 150, 4096, 8
    ze = edges_from_defn(nz, 0.0, 4000.0, 11.8)
    ye = edges_from_defn(ny, 0.0, 50.0e5, 1.0)
    depth, ypos = SpatialAxis("depth", ze), SpatialAxis("ypos", ye)
    tr = modules.Transport2D(depth, ypos, 0.1, 1000.0)
    m = modules.phosphorus_model(tr)
    t_all, h_all = graded_schedule(0.0, 365.0 * 86400.0)
    sched = (t_all[:nsteps].copy(), h_all[:nsteps].copy())
    m.set_schedule(*sched)
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand((3, nz, ny, B), dtype=torch.float64, device="cuda", generator=gen) + 0.1
    f = m.eval(x, B).clone()
    m.check_health()
    perm = torch.randperm(B, device="cuda", generator=gen)
    fp = m.eval(x[..., perm].contiguous(), B)
    assert torch.equal(fp, f[..., perm])
    scale = float(f.abs().max())
    os.environ["NKB_FUSED_PERSIST"] = "0"
    try:
        assert torch.equal(m.eval(x, B), f)
    finally:
        del os.environ["NKB_FUSED_PERSIST"]
    os.environ["NKB_FUSED_P3"] = "0"
    try:
        fu = m.eval(x, B)
    finally:
        del os.environ["NKB_FUSED_P3"]
    assert float((fu - f).abs().max()) <= 1e-11 * scale
    g = o.Grid2D(ze, ye, 0.1, 1000.0)
    mod = im.Module2D("phosphorus", g, phos=o.Phosphorus2D(g))
    want = im.model_year_2d(mod, x[..., 7:8].cpu().numpy(), schedule=sched)
    np.testing.assert_allclose(f[..., 7:8].cpu().numpy(), want, rtol=0, atol=1e-11 * np.abs(want).max())
    # conservation: sources sum to zero over the tracers, transport and sinking are flux divergences
    # with closed boundaries -> the volume integral of F summed over the tracers vanishes
    vol = torch.from_numpy(np.outer(depth.delta, ypos.delta)).cuda()
    total = (f[..., :64].sum(dim=0) * vol[..., None]).sum(dim=(0, 1))
    content = (x[..., :64].sum(dim=0) * vol[..., None]).sum(dim=(0, 1))
    assert float((total / content).abs().max()) <= 1e-12


@pytest.mark.parametrize("B", [1, 24])
def test_forced_surface_restoring_to_a_record(B, golden_dir, tmp_path):
    """forced_surf_restore_opt = file (py_driver_2d/forced.py:46-51,124-130): the restoring value is a
    record in time and ypos (K3 interpolates it per implicit stage into the affine surface source).
    (i) the device tendency against the REFERENCE's own comp_tend (golden vectors, record on a coarser
    ypos axis read through the host mirror's forcing reader), (ii) the model year — fused step kernel
    for B >= 8, stage kernels for B = 1 — against the numpy statement of the scheme"""
    import os

    from scipy.io import netcdf_file

    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200.py_driver_2d import modules
    from nk_ooc_b200.py_driver_2d.model_state import read_forcing
    from nk_ooc_b200.spatial_axis import SpatialAxis

    gv = np.load(os.path.join(golden_dir, "forced_restore_file.npz"))
    nz, ny, ratio, vvel, kh = gv["params"]
    depth, ypos = SpatialAxis("depth", gv["depth_edges"]), SpatialAxis("ypos", gv["ypos_edges"])
    fname = str(tmp_path / "surf_restore.nc")
    with netcdf_file(fname, "w", version=2) as f:
        f.createDimension("time", len(gv["rtimes"]))
        f.createDimension("ypos", len(gv["ypos_file"]))
        for name, vals in (("time", gv["rtimes"]), ("ypos", gv["ypos_file"])):
            v = f.createVariable(name, "f8", (name,))
            v[:] = vals
        v = f.createVariable("surf_vals", "f8", ("time", "ypos"))
        v[:] = gv["rdata"]
    rtimes, rdata = read_forcing(fname, "surf_vals", [ypos.mid])
    tr = modules.Transport2D(depth, ypos, float(vvel), float(kh))
    m = modules.forced_model(tr, "file", surf_restore_rate_10m=1.0 / 7200.0, sms_opt="const", sms_const=-1.0e-9,
                             surf_restore_times=rtimes, surf_restore_data=rdata)
    x1 = gv["x"]
    for i, t in enumerate(gv["times"]):
        got = m.tend(float(t), _to_dev(x1[..., None]), 1).cpu().numpy()[..., 0]
        np.testing.assert_allclose(got, gv["tend"][i], rtol=0, atol=1e-12 * np.abs(gv["tend"][i]).max())
    g = o.Grid2D(gv["depth_edges"], gv["ypos_edges"], float(vvel), float(kh))
    f2 = o.Forced2D(g, restore_rate_10m=1.0 / 7200.0, restore_const=None, sms_opt="const", sms_const=-1.0e-9,
                    restore_times=rtimes, restore_data=rdata)
    rng = np.random.default_rng(B)
    x = np.abs(rng.normal(size=(1, g.nz, g.ny, B)))
    nsteps = 24
    m.set_uniform_schedule(nsteps)
    got = m.eval(_to_dev(x), B).cpu().numpy()[..., :B]
    want = im.model_year_2d(im.Module2D("forced", g, forced=f2), x, nsteps)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * np.abs(want).max())
    with pytest.raises(ValueError):
        modules.forced_model(tr, "file", surf_restore_times=rtimes[:1], surf_restore_data=rdata[:1])


@pytest.mark.parametrize("kind", ["iage", "phosphorus"])
def test_single_state_takes_the_fused_step_kernel(kind, monkeypatch):
    """a single state in the reference's own layout (B = 1, member pitch 1: the Newton iterate, every
    Krylov product) is staged into a 4-lane batch and integrated by ONE persistent launch instead of
    two launches per time step; same result as the stage-per-launch kernels to rounding, hist
    snapshots included"""
    from oracle import imex_oracle as im
    from oracle import nk_oracle as o
    from nk_ooc_b200 import _lib
    from nk_ooc_b200.py_driver_2d import modules

    rng = np.random.default_rng(51)
    g, tr = _grid(21, 33)
    nsteps = 240
    if kind == "iage":
        mod, m = im.Module2D("iage", g), modules.iage_model(tr)
    else:
        mod, m = im.Module2D("phosphorus", g, phos=o.Phosphorus2D(g)), modules.phosphorus_model(tr)
    m.set_uniform_schedule(nsteps)
    x = np.abs(rng.normal(size=(mod.T, g.nz, g.ny, 1))) * 0.5
    xd = torch.from_numpy(x).cuda()
    assert xd.shape[-1] == 1
    lib = _lib.load()
    m.eval(xd, 1)
    n0 = lib.nkb_launch_count()
    got = m.eval(xd, 1).cpu().numpy()
    m.check_health()
    # scatter, [layout conversion,] persistent step launch, final difference, gather
    assert lib.nkb_launch_count() - n0 == (4 if kind == "iage" else 5)
    snaps = []
    want = im.model_year_2d(mod, x, nsteps, snapshots=snaps)
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * scale)
    f, hist = m.eval(xd, 1, hist_steps=[0, 100, 240])
    np.testing.assert_allclose(hist[1].cpu().numpy(), snaps[99][1][..., 0], rtol=0, atol=1e-10 * scale)
    np.testing.assert_allclose(hist[2].cpu().numpy(), x[..., 0] + want[..., 0], rtol=0, atol=1e-10 * scale)
    monkeypatch.setenv("NKB_FUSED_MIN_B", "8")
    n0 = lib.nkb_launch_count()
    unfused = m.eval(xd, 1).cpu().numpy()
    assert lib.nkb_launch_count() - n0 == 2 * nsteps
    np.testing.assert_allclose(got, unfused, rtol=0, atol=1e-10 * scale)
