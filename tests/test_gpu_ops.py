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
