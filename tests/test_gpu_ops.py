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
    with pytest.raises(ValueError):
        rw.limiter_scalef(_dev(bad), _dev(inc), 0.0, None, B)
