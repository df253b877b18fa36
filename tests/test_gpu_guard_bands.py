"""Out-of-bounds writes (compute-sanitizer is closed on the GPU pool: the checks are our own).  Every buffer a C-ABI
call gets — state, result, work space, right-hand sides — is carved out of ONE arena filled with a canary bit
pattern, with guard bands on both sides; after the call the guard bands (and the inputs) must be bit-identical, and
padding members beyond B must not leak into the members' results.  Sizes are ragged on purpose (columns not a multiple
of the tile width, members not a multiple of the member block)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CANARY = -7.123456789e301
GUARD = 4096  # doubles on each side of a buffer


class Arena:
    """consecutive buffers in one canary-filled allocation: guard | buf0 | guard | buf1 | guard ..."""

    def __init__(self, sizes):
        self.sizes = [int(s) for s in sizes]
        pad = lambda n: (n + 63) // 64 * 64  # noqa: E731  (512-byte aligned starts)
        self.offsets, off = [], GUARD
        for n in self.sizes:
            self.offsets.append(off)
            off += pad(n) + GUARD
        self.mem = torch.full((off,), CANARY, dtype=torch.float64, device="cuda")

    def buf(self, i):
        return self.mem[self.offsets[i]: self.offsets[i] + self.sizes[i]]

    def assert_guards_intact(self, what):
        torch.cuda.synchronize()
        keep = torch.ones(self.mem.numel(), dtype=torch.bool, device="cuda")
        for off, n in zip(self.offsets, self.sizes):
            keep[off: off + n] = False
        guards = self.mem[keep]
        bad = int((guards.view(torch.int64) != torch.tensor(CANARY, dtype=torch.float64).view(torch.int64).item()).sum())
        assert bad == 0, f"{what}: {bad} doubles outside the declared buffers were overwritten"


def _model(kind, nz, ny):
    from test_gpu_stage import _forcing, _grid
    from nk_ooc_b200.py_driver_2d import modules

    g, tr = _grid(nz, ny)
    if kind == "iage":
        return modules.iage_model(tr)
    if kind == "phosphorus":
        return modules.phosphorus_model(tr)
    times, data = _forcing(g, np.random.default_rng(5))
    return modules.forced_model(tr, "const", 1.0, 1.0 / 3600.0, "file", sms_times=times, sms_data=data, sink_thres=0.05)


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("B", [1, 5, 37])
@pytest.mark.parametrize("kind", ["iage", "forced", "phosphorus"])
def test_model_year_writes_only_its_buffers(kind, B, fused, monkeypatch):
    from nk_ooc_b200.engine import padded_members

    monkeypatch.setenv("NKB_FUSED", fused)
    nz, ny, nsteps = 13, 19, 12  # 19 columns: one full tile of 14 and a ragged one
    model = _model(kind, nz, ny)
    model.set_uniform_schedule(nsteps)
    ldb = padded_members(B)
    n = model.T * nz * ny * ldb
    need = model.lib.nkb_model_work_doubles(model.handle, B, ldb)
    n_hist = 2
    arena = Arena([n, n, need, n_hist * model.T * nz * ny])
    x, out, work, hist = (arena.buf(i) for i in range(4))
    rng = np.random.default_rng(11)
    x.copy_(torch.from_numpy(np.abs(rng.normal(size=n)) * 0.5 + 0.05).cuda())
    xv = x.view(model.T, nz, ny, ldb)
    x_before = x.clone()
    steps = (ctypes.c_int * n_hist)(3, 7)
    from nk_ooc_b200.engine import _stream_ptr, check

    for _ in range(2):  # (the first call also builds the coefficient tables)
        check(model.lib.nkb_model_eval(model.handle, x.data_ptr(), out.data_ptr(), work.data_ptr(), B, ldb, n_hist,
                                       steps, hist.data_ptr(), _stream_ptr()), "nkb_model_eval")
    arena.assert_guards_intact(f"nkb_model_eval {kind} B={B} fused={fused}")
    model.check_health()
    assert torch.equal(x.view(torch.int64), x_before.view(torch.int64)), "the input state was modified"
    got = out.view(model.T, nz, ny, ldb)[..., :B].clone()
    assert torch.isfinite(got).all() and torch.isfinite(hist).all()
    # the padding members must not influence the members: the same call with other padding values
    if ldb > B:
        xv[..., B:] = 3.0e7
        check(model.lib.nkb_model_eval(model.handle, x.data_ptr(), out.data_ptr(), work.data_ptr(), B, ldb, 0, None,
                                       None, _stream_ptr()), "nkb_model_eval")
        torch.cuda.synchronize()
        assert torch.equal(out.view(model.T, nz, ny, ldb)[..., :B], got)
        arena.assert_guards_intact("second call")


@pytest.mark.parametrize("n,kl,ku,B", [(133, 1, 1, 5), (133, 3, 2, 37), (997, 61, 33, 5), (1003, 48, 48, 37),
                                      (2000, 40, 40, 1)])
def test_banded_solve_writes_only_its_buffers(n, kl, ku, B):
    from nk_ooc_b200.engine import BandedFactor, _stream_ptr, check, padded_members

    rng = np.random.default_rng(3)
    ab = rng.normal(size=(kl + ku + 1, n))
    ab[ku] += 4.0 * (kl + ku + 1)  # diagonally dominant: no interchanges, every solver path is eligible
    fac = BandedFactor(ab, kl, ku)
    ldb = padded_members(B)
    arena = Arena([n * ldb, n * ldb])
    y, x = arena.buf(0), arena.buf(1)
    y.copy_(torch.from_numpy(rng.normal(size=n * ldb)).cuda())
    y_before = y.clone()
    check(fac.lib.nkb_banded_solve(fac.handle, y.data_ptr(), x.data_ptr(), B, ldb, 1.0, 0, _stream_ptr()),
          "nkb_banded_solve")
    arena.assert_guards_intact(f"nkb_banded_solve n={n} kl={kl} ku={ku} B={B} path={fac.path}")
    assert torch.equal(y.view(torch.int64), y_before.view(torch.int64))
    from scipy.linalg import solve_banded

    want = solve_banded((kl, ku), ab, y.view(n, ldb)[:, :B].cpu().numpy())
    np.testing.assert_allclose(x.view(n, ldb)[:, :B].cpu().numpy(), want, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("B", [1, 5, 37])
def test_krylov_vector_kernels_write_only_their_buffers(B):
    """nkb_mgs (w in place + scratch), nkb_lin_comb (out), nkb_axpby (y in place) with regions and excluded cells"""
    from nk_ooc_b200.engine import RegionWeights, padded_members

    T, nz, ny, k = 2, 13, 19, 3
    rng = np.random.default_rng(17)
    mask = np.tile(np.arange(ny) % 3 + 1, (nz, 1)).astype(np.int32)
    mask[rng.random(size=mask.shape) < 0.1] = 0
    rw = RegionWeights(mask, np.abs(rng.normal(size=(nz, ny))) + 0.1)
    ldb = padded_members(B)
    n = T * nz * ny * ldb
    from nk_ooc_b200 import _lib

    need = _lib.load().nkb_mgs_scratch_doubles(rw.region_cnt, B, rw.max_row)
    arena = Arena([n] * (k + 3) + [need])
    bufs = [arena.buf(i) for i in range(k + 3)]
    for b in bufs:
        b.copy_(torch.from_numpy(rng.normal(size=n)).cuda())
    shape = (T, nz, ny, ldb)
    basis = [b.view(shape) for b in bufs[:k]]
    w, out, y = (b.view(shape) for b in bufs[k:])
    rw._mgs_scratch = arena.buf(k + 3)  # pylint: disable=protected-access
    before = [b.clone() for b in bufs[:k]]
    h = rw.mgs(w, basis, B)
    arena.assert_guards_intact("nkb_mgs")
    rw.lin_comb(h, basis, B, add=w, out=out)
    arena.assert_guards_intact("nkb_lin_comb")
    alpha = torch.from_numpy(rng.normal(size=(rw.region_cnt, B))).cuda()
    rw.axpby(alpha, w, 0.5, y, B)
    arena.assert_guards_intact("nkb_axpby")
    for b, b0 in zip(bufs[:k], before):
        assert torch.equal(b.view(torch.int64), b0.view(torch.int64)), "a basis vector was modified"
    assert torch.isfinite(h).all() and torch.isfinite(out[..., :B]).all() and torch.isfinite(y[..., :B]).all()
