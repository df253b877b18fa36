"""GPU parity tests of the Krylov vector kernels and the banded solver against the oracle"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _dev(x):
    from nk_ooc_b200.engine import padded_members

    B = x.shape[-1]
    out = torch.zeros(x.shape[:-1] + (padded_members(B),), dtype=torch.float64, device="cuda")
    out[..., :B] = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return out


@pytest.mark.parametrize("B", [1, 7, 33, 100])
def test_pack_unpack_roundtrip(B):
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(B)
    x = rng.normal(size=(B, 2, 5, 9))
    xd = torch.from_numpy(x).cuda()
    fast = engine.pack(xd)
    assert fast.shape == (2, 5, 9, engine.padded_members(B))
    np.testing.assert_array_equal(fast.cpu().numpy()[..., :B], np.moveaxis(x, 0, -1))
    back = engine.unpack(fast, B)
    np.testing.assert_array_equal(back.cpu().numpy(), x)


@pytest.mark.parametrize("B", [1, 4, 50])
@pytest.mark.parametrize("regions", ["one", "columns", "masked"])
def test_wdot_and_mean(B, regions):
    from oracle import nk_oracle as o
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(7)
    nz, ny, T = 20, 13, 2
    wgt = np.outer(rng.uniform(1, 5, nz), rng.uniform(1, 2, ny))
    if regions == "one":
        mask = np.ones((nz, ny), dtype=np.int32)
    elif regions == "columns":
        mask = o.column_region_mask(nz, ny, 0.0, 0.0)
    else:
        mask = rng.integers(0, 4, size=(nz, ny)).astype(np.int32)
    w = o.region_weights(mask, wgt)
    rw = engine.RegionWeights(mask, wgt)
    a = rng.normal(size=(T, nz, ny, B))
    b = rng.normal(size=(T, nz, ny, B))
    got = rw.dot(_dev(a), _dev(b), B).cpu().numpy()
    want = np.stack([o.dot_prod(w, a[..., i], b[..., i]) for i in range(B)], axis=-1)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)
    got = rw.dot(_dev(a), None, B).cpu().numpy()
    want = np.stack([o.mean(w, a[..., i]) for i in range(B)], axis=-1)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)


def test_axpby_region_scalars():
    from oracle import nk_oracle as o
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(9)
    nz, ny, T, B = 8, 6, 2, 5
    mask = rng.integers(0, 3, size=(nz, ny)).astype(np.int32)
    wgt = np.ones((nz, ny))
    rw = engine.RegionWeights(mask, wgt)
    R = rw.region_cnt
    x = rng.normal(size=(T, nz, ny, B))
    y = rng.normal(size=(T, nz, ny, B))
    alpha = rng.normal(size=(R, B))
    beta = rng.normal(size=(R, B))
    yd = _dev(y)
    rw.axpby(torch.from_numpy(alpha).cuda(), _dev(x), torch.from_numpy(beta).cuda(), yd, B)
    want = np.empty_like(y)
    for i in range(B):
        al = o.broadcast_region_vals(mask, alpha[:, i])
        be = o.broadcast_region_vals(mask, beta[:, i])
        want[..., i] = al * x[..., i] + be * y[..., i]
    np.testing.assert_allclose(yd.cpu().numpy()[..., :B], want, rtol=1e-14, atol=0)


@pytest.mark.parametrize("n,kl,ku", [(20, 1, 1), (60, 3, 3), (120, 9, 9), (50, 4, 2)])
def test_banded_solve_matches_scipy(n, kl, ku):
    from scipy import linalg
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(n)
    ab = rng.normal(size=(kl + ku + 1, n))
    ab[ku] += 0.5  # not diagonally dominant: exercises the pivoting
    B = 9
    y = rng.normal(size=(n, B))
    want = linalg.solve_banded((kl, ku), ab, y)
    f = engine.BandedFactor(ab, kl, ku)
    got = f.solve(_dev(y), B).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9 * np.abs(want).max())
    got = f.solve(_dev(y), B, scale=0.25, subtract_rhs=True).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, 0.25 * want - y, rtol=0, atol=1e-9 * np.abs(want).max())


@pytest.mark.parametrize("B", [1, 7])
def test_limiter_scalef_matches_oracle(B):
    """nkb_limiter_scalef against the restated comp_scalef_lob/upb (utils.py:561-600), incl. the
    reference's own known-answer cases (tests/test_utils.py:144-224) in member 0"""
    from oracle import nk_oracle as o
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(13)
    region_cnt = 7
    shape = (3, region_cnt)
    mask = np.zeros(shape, dtype=np.int32)
    for r in range(region_cnt):
        mask[:, r] = r + 1
    base = np.ones((1,) + shape + (B,))
    inc = np.ones((1,) + shape + (B,))
    inc[0, 0, 1, 0] = -0.5
    inc[0, 0, 2, 0], inc[0, 1, 2, 0] = -0.5, -1.0
    inc[0, 0, 3, 0], inc[0, 1, 3, 0], inc[0, 2, 3, 0] = -0.5, -1.0, -2.0
    base[0, :, 4:, 0] = 0.0
    inc[0, 0, 5, 0] = 0.0
    inc[0, 0, 6, 0], inc[0, 1, 6, 0] = 0.0, -1.0
    for b in range(1, B):
        base[..., b] = rng.random(size=(1,) + shape) + 0.1
        inc[..., b] = rng.normal(size=(1,) + shape)
    rw = engine.RegionWeights(mask, np.ones(shape))
    got = rw.limiter_scalef(_dev(base), _dev(inc), 0.0, None, B).cpu().numpy()
    assert (got[:, 0] == np.array([1.0, 1.0, 1.0, 0.5, 1.0, 1.0, 0.0])).all()
    for b in range(B):
        want = o.comp_scalef_lob(region_cnt, mask, base[0, ..., b], inc[0, ..., b], 0.0)
        np.testing.assert_array_equal(np.minimum(got[:, b], 1.0), want)
        want = o.comp_scalef_upb(region_cnt, mask, -base[0, ..., b], -inc[0, ..., b], 0.0)
        gotu = rw.limiter_scalef(_dev(-base), _dev(-inc), None, 0.0, B).cpu().numpy()
        np.testing.assert_array_equal(np.minimum(gotu[:, b], 1.0), want)
    bad = base.copy()
    bad[0, 1, 1, 0] = -1.0
    with pytest.raises(ValueError, match="base < lob"):
        rw.limiter_scalef(_dev(bad), _dev(inc), 0.0, None, B)
    with pytest.raises(ValueError, match="base > upb"):
        rw.limiter_scalef(_dev(-bad), _dev(-inc), None, 0.0, B)
    # base a hair below the bound with an increment that never needs limiting: the reference returns 1
    # everywhere and does NOT raise (utils.py:571-573: the all-clear test comes first)
    hair = np.ones((1,) + shape + (B,))
    hair[0, 1, 1, 0] = -1.0e-12
    up = np.full((1,) + shape + (B,), 0.25)
    for b in range(B):
        want = o.comp_scalef_lob(region_cnt, mask, hair[0, ..., b], up[0, ..., b], 0.0)
        assert (want == 1.0).all()
    got = rw.limiter_scalef(_dev(hair), _dev(up), 0.0, None, B).cpu().numpy()
    assert (np.minimum(got, 1.0) == 1.0).all()


# ---- wide batches (two-members-per-lane kernels) and the windowed banded solver -------------------
@pytest.mark.parametrize("B", [64, 101, 256])
@pytest.mark.parametrize("regions", ["one", "columns", "masked"])
def test_wide_batch_wdot_axpby_limiter_match_narrow_path(B, regions):
    """B >= 64 takes the vectorised kernels: compared with the oracle (dot, mean, axpby, limiter) on
    ragged member counts (odd B: the second member of the last lane pair does not exist)"""
    from oracle import nk_oracle as o
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(B)
    nz, ny, T = 21, 13, 2  # 273 cells: not a multiple of the cells per warp
    wgt = np.outer(rng.uniform(1, 5, nz), rng.uniform(1, 2, ny))
    if regions == "one":
        mask = np.ones((nz, ny), dtype=np.int32)
    elif regions == "columns":
        mask = o.column_region_mask(nz, ny, 0.0, 0.0)
    else:
        mask = rng.integers(0, 4, size=(nz, ny)).astype(np.int32)
    w = o.region_weights(mask, wgt)
    rw = engine.RegionWeights(mask, wgt)
    R = rw.region_cnt
    a = rng.normal(size=(T, nz, ny, B))
    b = rng.normal(size=(T, nz, ny, B))
    got = rw.dot(_dev(a), _dev(b), B).cpu().numpy()
    want = np.stack([o.dot_prod(w, a[..., i], b[..., i]) for i in range(B)], axis=-1)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)
    got = rw.dot(_dev(a), None, B).cpu().numpy()
    want = np.stack([o.mean(w, a[..., i]) for i in range(B)], axis=-1)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)
    # axpby with per-(region, member) scalars, with x = None and with beta = 0 over NaN
    alpha, beta = rng.normal(size=(R, B)), rng.normal(size=(R, B))
    yd = _dev(b)
    rw.axpby(torch.from_numpy(alpha).cuda(), _dev(a), torch.from_numpy(beta).cuda(), yd, B)
    want = np.empty_like(b)
    for i in range(B):
        al, be = o.broadcast_region_vals(mask, alpha[:, i]), o.broadcast_region_vals(mask, beta[:, i])
        want[..., i] = al * a[..., i] + be * b[..., i]
    # (the kernel fuses be*y + (al*x) into one fma: differences of one rounding, amplified where the two
    # terms cancel)
    np.testing.assert_allclose(yd.cpu().numpy()[..., :B], want, rtol=1e-12, atol=1e-15)
    yd = _dev(b)
    rw.axpby(None, None, torch.from_numpy(beta).cuda(), yd, B)
    for i in range(B):
        want[..., i] = o.broadcast_region_vals(mask, beta[:, i]) * b[..., i]
    np.testing.assert_allclose(yd.cpu().numpy()[..., :B], want, rtol=1e-14, atol=0)
    yd = torch.full_like(_dev(b), float("nan"))
    rw.axpby(2.0, _dev(a), 0.0, yd, B)
    np.testing.assert_array_equal(yd.cpu().numpy()[..., :B], 2.0 * a)
    # limiter
    base = rng.random(size=(T, nz, ny, B)) + 0.05
    inc = rng.normal(size=(T, nz, ny, B))
    got = rw.limiter_scalef(_dev(base), _dev(inc), 0.0, None, B).cpu().numpy()
    for i in range(0, B, 7):
        want = np.minimum.reduce([o.comp_scalef_lob(R, mask, base[t, ..., i], inc[t, ..., i], 0.0) for t in range(T)])
        np.testing.assert_array_equal(np.minimum(got[:, i], 1.0), want)
    bad = base.copy()
    bad[1, 3, 3, B - 1] = -1.0
    if mask[3, 3] > 0:
        with pytest.raises(ValueError):
            rw.limiter_scalef(_dev(bad), _dev(inc), 0.0, None, B)


@pytest.mark.parametrize("n,kl,ku,B", [(900, 90, 90, 1), (900, 90, 90, 5), (400, 61, 33, 40), (2000, 1, 1, 130),
                                       (300, 2, 5, 33), (64, 63, 63, 3)])
def test_banded_window_solver(n, kl, ku, B):
    """the shared-memory window kernel: wide bands with one right-hand side (band-parallel), narrow
    bands with many (member-parallel), pivoting, scale/subtract epilogue, in place"""
    from scipy import linalg
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(n + kl)
    ab = rng.normal(size=(kl + ku + 1, n))
    ab[ku] += 0.5 * np.sqrt(kl + ku)
    y = rng.normal(size=(n, B))
    want = linalg.solve_banded((kl, ku), ab, y)
    f = engine.BandedFactor(ab, kl, ku)
    assert f.n_blocks == 1
    tol = 1e-9 * np.abs(want).max()
    got = f.solve(_dev(y), B).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, want, rtol=0, atol=tol)
    got = f.solve(_dev(y), B, scale=0.25, subtract_rhs=True).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, 0.25 * want - y, rtol=0, atol=tol)
    yd = _dev(y)
    ldb = yd.shape[-1]
    engine.check(f.lib.nkb_banded_solve(f.handle, yd.data_ptr(), yd.data_ptr(), B, ldb, 1.0, 0, None), "in place")
    np.testing.assert_allclose(yd.cpu().numpy()[:, :B], want, rtol=0, atol=tol)


@pytest.mark.parametrize("n,kl,ku,B", [(2000, 40, 40, 1), (1999, 17, 33, 5), (1003, 90, 21, 32), (48, 16, 16, 3),
                                       (4500, 450, 450, 1), (700, 130, 130, 70)])
def test_banded_panel_solver(n, kl, ku, B, monkeypatch):
    """the panel (blocked) substitution kernel: one wide block factored without row interchanges (diagonally
    dominant, like the 2-D preconditioners I - dt J): 16 rows per barrier pair with inverted diagonal blocks.
    Against scipy and bit-for-bit-irrelevant but rounding-level against the row-by-row window kernel; sizes that
    are not multiples of the panel, kl != ku, several member groups, scale/subtract epilogue, in place"""
    from scipy import linalg
    from nk_ooc_b200 import _lib, engine

    rng = np.random.default_rng(n + kl + B)
    ab = rng.normal(size=(kl + ku + 1, n))
    ab[ku] = 1.5 * np.abs(ab).sum(axis=0) + 1.0  # strictly diagonally dominant by columns: no interchanges
    y = rng.normal(size=(n, B))
    want = linalg.solve_banded((kl, ku), ab, y)
    lib = _lib.load()
    f = engine.BandedFactor(ab, kl, ku)
    assert f.n_blocks == 1 and f.path == "panel"
    tol = 1e-12 * np.abs(want).max()
    n0 = lib.nkb_launch_count()
    got = f.solve(_dev(y), B).cpu().numpy()[:, :B]
    assert lib.nkb_launch_count() - n0 == 1
    np.testing.assert_allclose(got, want, rtol=0, atol=tol)
    got = f.solve(_dev(y), B, scale=0.25, subtract_rhs=True).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, 0.25 * want - y, rtol=0, atol=tol)
    yd = _dev(y)
    ldb = yd.shape[-1]
    engine.check(f.lib.nkb_banded_solve(f.handle, yd.data_ptr(), yd.data_ptr(), B, ldb, 1.0, 0, None), "in place")
    np.testing.assert_allclose(yd.cpu().numpy()[:, :B], want, rtol=0, atol=tol)
    # the same factor through the window kernel (panel data not built)
    monkeypatch.setenv("NKB_BANDED_PANEL", "0")
    f2 = engine.BandedFactor(ab, kl, ku)
    assert f2.path == "window"
    got2 = f2.solve(_dev(y), B).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, 0.25 * got2 - y, rtol=0, atol=tol)
    # a matrix that needs interchanges keeps the pivoting window kernel
    monkeypatch.delenv("NKB_BANDED_PANEL")
    ab3 = rng.normal(size=(kl + ku + 1, n))
    ab3[ku] += 0.5 * np.sqrt(kl + ku)
    f3 = engine.BandedFactor(ab3, kl, ku)
    assert f3.path == "window"
    got3 = f3.solve(_dev(y), B).cpu().numpy()[:, :B]
    want3 = linalg.solve_banded((kl, ku), ab3, y)
    np.testing.assert_allclose(got3, want3, rtol=0, atol=1e-9 * np.abs(want3).max())


@pytest.mark.parametrize("kl,ku", [(1, 1), (2, 1), (1, 3), (4, 4)])
@pytest.mark.parametrize("n", [7, 125, 700])
def test_banded_thomas_kernel_narrow_bands(n, kl, ku):
    """diagonally dominant narrow bands are factored without row interchanges and take the batched
    Thomas kernel (B >= 16; z in shared memory up to 256 rows, through the output buffer beyond);
    compared with scipy and with the window kernel on the same factor"""
    from scipy import linalg
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(n * 10 + kl)
    ab = rng.normal(size=(kl + ku + 1, n))
    ab[ku] = 6.0 + rng.random(n)
    B = 45
    y = rng.normal(size=(n, B))
    want = linalg.solve_banded((kl, ku), ab, y)
    f = engine.BandedFactor(ab, kl, ku)
    tol = 1e-12 * np.abs(want).max()
    got = f.solve(_dev(y), B).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, want, rtol=0, atol=tol)
    got = f.solve(_dev(y), B, scale=3.0, subtract_rhs=True).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, 3.0 * want - y, rtol=0, atol=10 * tol)
    yd = _dev(y)
    engine.check(f.lib.nkb_banded_solve(f.handle, yd.data_ptr(), yd.data_ptr(), B, yd.shape[-1], 1.0, 0, None), "in place")
    np.testing.assert_allclose(yd.cpu().numpy()[:, :B], want, rtol=0, atol=tol)
    got1 = f.solve(_dev(y[:, :3]), 3, scale=3.0, subtract_rhs=True).cpu().numpy()[:, :3]  # B < 16: window kernel
    np.testing.assert_allclose(got1, 3.0 * want[:, :3] - y[:, :3], rtol=0, atol=10 * tol)


@pytest.mark.parametrize("B", [1, 70])
def test_banded_block_diagonal_systems_are_found_and_solved(B):
    """per-column tridiagonal systems stored as ONE band (grid without lateral processes): the blocks
    are detected, factored and solved in parallel; ragged block sizes; a wider block-diagonal band"""
    from scipy import linalg
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(3)
    sizes = [20, 1, 33, 20, 7]
    n = sum(sizes)
    ab = rng.normal(size=(3, n))
    ab[1] += 3.0
    edge = np.cumsum(sizes)[:-1]
    ab[0, edge] = 0.0      # A(edge-1, edge)
    ab[2, edge - 1] = 0.0  # A(edge, edge-1)
    f = engine.BandedFactor(ab, 1, 1)
    assert f.n_blocks == len(sizes)
    y = rng.normal(size=(n, B))
    want = linalg.solve_banded((1, 1), ab, y)
    got = f.solve(_dev(y), B, scale=2.0, subtract_rhs=True).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, 2.0 * want - y, rtol=0, atol=1e-11 * np.abs(want).max())
    # dense diagonal blocks of size 12 inside a band of half-width 11
    nb, m = 9, 12
    dense = np.zeros((nb * m, nb * m))
    for i in range(nb):
        dense[i * m:(i + 1) * m, i * m:(i + 1) * m] = rng.normal(size=(m, m)) + 4.0 * np.eye(m)
    kl = ku = m - 1
    ab = np.zeros((kl + ku + 1, nb * m))
    r, c = np.nonzero(dense)
    ab[ku + r - c, c] = dense[r, c]
    f = engine.BandedFactor(ab, kl, ku)
    assert f.n_blocks == nb
    y = rng.normal(size=(nb * m, B))
    want = np.linalg.solve(dense, y)
    got = f.solve(_dev(y), B).cpu().numpy()[:, :B]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * np.abs(want).max())


def test_production_size_vector_kernels_and_column_solves():
    """BASELINE.json's headline size (125 x 150 cells, 4096 members) for the Krylov vector kernels and the
    per-column tridiagonal preconditioner solves, checked on sampled members against numpy / scipy and
    through size-independent properties (linearity of the solve, dot(a, a) = norm^2, axpby inverse)"""
    from scipy import linalg
    from nk_ooc_b200 import engine

    nz, ny, B = 125, 150, 4096
    n = nz * ny
    gen = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((1, n, B), dtype=torch.float64, device="cuda", generator=gen)
    b = torch.randn((1, n, B), dtype=torch.float64, device="cuda", generator=gen)
    rng = np.random.default_rng(0)
    wgt = np.outer(rng.uniform(1, 5, nz), rng.uniform(1, 2, ny))
    mask = np.ones((nz, ny), dtype=np.int32)
    rw = engine.RegionWeights(mask, wgt)
    w = (wgt / wgt.sum()).reshape(-1)
    sample = [0, 1, 777, 4095]
    dots = rw.dot(a, b, B).cpu().numpy()
    means = rw.dot(a, None, B).cpu().numpy()
    for m in sample:
        am, bm = a[0, :, m].cpu().numpy(), b[0, :, m].cpu().numpy()
        np.testing.assert_allclose(dots[0, m], np.sum(w * am * bm), rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(means[0, m], np.sum(w * am), rtol=0, atol=1e-13)
    # y <- alpha x + beta y, then its inverse: back to y to rounding
    alpha = torch.rand((1, B), dtype=torch.float64, device="cuda", generator=gen) + 0.5
    beta = torch.rand((1, B), dtype=torch.float64, device="cuda", generator=gen) + 0.5
    y = b.clone()
    rw.axpby(alpha, a, beta, y, B)
    np.testing.assert_allclose(y[0, :, 777].cpu().numpy(),
                               float(alpha[0, 777]) * a[0, :, 777].cpu().numpy() + float(beta[0, 777]) * b[0, :, 777].cpu().numpy(),
                               rtol=1e-13, atol=1e-15)
    rw.axpby(-alpha / beta, a, 1.0 / beta, y, B)
    assert float((y - b).abs().max()) <= 1e-12
    # limiter: base + scalef * inc stays >= 0 for every member, and is tight for the binding cell
    base = torch.rand((1, n, B), dtype=torch.float64, device="cuda", generator=gen) + 0.01
    sc = rw.limiter_scalef(base, a, 0.0, None, B)
    low = (base + torch.minimum(sc, torch.ones_like(sc)).reshape(1, 1, B) * a).amin(dim=(0, 1))
    assert float(low.min()) >= -1e-12 and float(low.abs().max()) <= 1.0
    np.testing.assert_allclose(low[sample].cpu().numpy(), 0.0, atol=1e-12)
    # 150 per-column tridiagonal systems as one band: blocks found, sampled members against scipy, linearity
    ab = np.zeros((3, n))
    ab[1] = 2.0 + rng.random(n)
    ab[0, 1:] = -rng.random(n - 1)
    ab[2, :-1] = -rng.random(n - 1)
    edge = np.arange(nz, n, nz)
    ab[0, edge] = 0.0
    ab[2, edge - 1] = 0.0
    fac = engine.BandedFactor(ab, 1, 1)
    assert fac.n_blocks == ny
    ya, yb = a.reshape(n, B), b.reshape(n, B)
    xa = fac.solve(ya, B)
    for m in sample:
        want = linalg.solve_banded((1, 1), ab, ya[:, m].cpu().numpy())
        np.testing.assert_allclose(xa[:, m].cpu().numpy(), want, rtol=0, atol=1e-12 * np.abs(want).max())
    xsum = fac.solve((2.0 * ya - 3.0 * yb).contiguous(), B)
    lin = 2.0 * xa - 3.0 * fac.solve(yb, B)
    assert float((xsum - lin).abs().max()) <= 1e-11 * float(lin.abs().max())
    res = fac.solve(ya, B, scale=0.5, subtract_rhs=True)
    assert float((res - (0.5 * xa - ya)).abs().max()) <= 1e-12 * float(xa.abs().max())


@pytest.mark.parametrize("n,kl,ku", [(900, 90, 90), (300, 40, 17), (257, 3, 200), (2000, 151, 151)])
def test_banded_cooperative_factorisation(n, kl, ku, monkeypatch):
    """one wide block factored by all SMs (cooperative launch, two grid barriers per column) gives the
    same pivots and — up to the rounding of the fused update — the same factor as the one-CTA kernel:
    solutions against scipy with pivoting exercised (not diagonally dominant), 1 and 9 right-hand sides"""
    from scipy import linalg
    from nk_ooc_b200 import engine

    rng = np.random.default_rng(n)
    ab = rng.normal(size=(kl + ku + 1, n))
    ab[ku] += 0.3
    y = rng.normal(size=(n, 9))
    want = linalg.solve_banded((kl, ku), ab, y)
    tol = 1e-8 * np.abs(want).max()
    got = {}
    for coop in ("1", "0"):
        monkeypatch.setenv("NKB_BANDED_COOP", coop)
        f = engine.BandedFactor(ab, kl, ku)
        got[coop] = f.solve(_dev(y), 9).cpu().numpy()[:, :9]
        np.testing.assert_allclose(got[coop], want, rtol=0, atol=tol)
        np.testing.assert_allclose(f.solve(_dev(y[:, :1]), 1).cpu().numpy()[:, 0], want[:, 0], rtol=0, atol=tol)
    np.testing.assert_allclose(got["1"], got["0"], rtol=0, atol=1e-3 * tol)


@pytest.mark.parametrize("resident", [True, False])
@pytest.mark.parametrize("B", [1, 5, 40])
@pytest.mark.parametrize("regions", ["one", "columns", "masked"])
def test_fused_mgs_and_lin_comb_match_the_vector_loop(B, regions, resident, monkeypatch):
    """nkb_mgs (one cooperative launch, w resident on the chip; or its general path) and nkb_lin_comb against
    the loop of the reference (model_state_base.py:365-377: h_i = dot(w, v_i); w -= h_i v_i; :619-624) done with
    numpy in float64 on the host, and against the k x (nkb_wdot, nkb_axpby) loop they replace"""
    from oracle import nk_oracle as o
    from nk_ooc_b200 import engine

    if not resident:
        monkeypatch.setenv("NKB_MGS_RESIDENT", "0")
    rng = np.random.default_rng(11)
    nz, ny, T, k = 20, 13, 2, 6
    wgt = np.outer(rng.uniform(1, 5, nz), rng.uniform(1, 2, ny))
    if regions == "one":
        mask = np.ones((nz, ny), dtype=np.int32)
    elif regions == "columns":
        mask = o.column_region_mask(nz, ny, 0.0, 0.0)
    else:
        mask = rng.integers(0, 4, size=(nz, ny)).astype(np.int32)
    w = o.region_weights(mask, wgt)
    rw = engine.RegionWeights(mask, wgt)
    R = rw.region_cnt
    basis = [rng.normal(size=(T, nz, ny, B)) for _ in range(k)]
    x = rng.normal(size=(T, nz, ny, B))
    # host loop (the reference's algorithm)
    want_w = x.copy()
    want_h = np.zeros((k, R, B))
    reg = np.where(mask > 0, mask - 1, 0)
    inside = (mask > 0)[None, :, :, None]
    for i in range(k):
        for b in range(B):
            want_h[i, :, b] = o.dot_prod(w, want_w[..., b], basis[i][..., b])
        want_w = want_w - np.where(inside, want_h[i][reg][None] * basis[i], 0.0)
    flat = lambda t: t.reshape(T, nz * ny, t.shape[-1])
    wd = _dev(x)
    bd = [_dev(v) for v in basis]
    n0 = engine._lib.load().nkb_launch_count()
    h = rw.mgs(flat(wd), [flat(v) for v in bd], B)
    launches = engine._lib.load().nkb_launch_count() - n0
    assert launches == (1 if resident else 3 * k)  # (the dot is a two-pass reduction)
    np.testing.assert_allclose(h.cpu().numpy(), want_h, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(wd.cpu().numpy()[..., :B], want_w, rtol=1e-12, atol=1e-13)
    # the loop of library calls it replaces
    w2 = _dev(x)
    for i in range(k):
        hi = rw.dot(flat(w2), flat(bd[i]), B)
        rw.axpby(-hi, flat(bd[i]), 1.0, flat(w2), B, fill_alpha=0.0)
        np.testing.assert_allclose(hi.cpu().numpy(), h[i].cpu().numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(w2.cpu().numpy(), wd.cpu().numpy(), rtol=1e-12, atol=1e-13)
    # run-to-run determinism of the fused kernel (fixed summation order)
    w3 = _dev(x)
    h3 = rw.mgs(flat(w3), [flat(v) for v in bd], B)
    assert torch.equal(h3, h) and torch.equal(w3, wd)
    # lin_comb
    coeff = rng.normal(size=(k, R, B))
    got = rw.lin_comb(torch.from_numpy(coeff).cuda(), [flat(v) for v in bd], B).cpu().numpy().reshape(T, nz, ny, -1)
    want = np.zeros((T, nz, ny, B))
    for i in range(k):
        want += np.where(inside, coeff[i][reg][None], 1.0) * basis[i]
    np.testing.assert_allclose(got[..., :B], want, rtol=1e-12, atol=1e-13)


def test_interleave_blocks():
    """[G][n][W] all-gather output -> member-fastest [n][ldo] with the first B of G*W members"""
    from nk_ooc_b200 import engine

    lib = engine._lib.load()
    rng = np.random.default_rng(3)
    G, n, W, B = 3, 50, 32, 70
    src = rng.normal(size=(G, n, W))
    ldo = engine.padded_members(B)
    out = torch.full((n, ldo), -7.0, dtype=torch.float64, device="cuda")
    sd = torch.from_numpy(src).cuda()
    engine.check(lib.nkb_interleave_blocks(sd.data_ptr(), out.data_ptr(), n, G, W, ldo, B, engine._stream_ptr()),
                 "nkb_interleave_blocks")
    want = np.moveaxis(src, 0, 1).reshape(n, G * W)[:, :B]
    got = out.cpu().numpy()
    np.testing.assert_array_equal(got[:, :B], want)
    assert (got[:, B:] == -7.0).all()
