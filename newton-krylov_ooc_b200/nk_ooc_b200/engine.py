"""Thin host wrappers over the C ABI: device memory and streams come from torch, every
computation is a call into libnkb200.so (no torch math on the hot path, no CPU fallback)."""

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ModelDesc, check, dptr

GAMMA = 1.0 - 1.0 / np.sqrt(2.0)


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.NkbError("no CUDA device: the nk_ooc_b200 hot path has no CPU fallback")


def _stream_ptr():
    """cudaStream_t of torch's current stream (the raw handle: constructing a torch.cuda.Stream wrapper per library
    call cost 13 % of a Newton step on the CI grid)"""
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # pylint: disable=protected-access
    if raw is not None:
        return ctypes.c_void_p(raw(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def padded_members(B):
    """leading dimension (member stride) used for a batch of B members"""
    return 1 if B == 1 else ((B + 31) // 32) * 32


def uniform_schedule(n_steps, t0, t1):
    """step start times t0 + k*h and sizes h (times are not accumulated so that the hist
    times k/60 yr (py_driver_2d/model_state.py:81) and the forcing kinks are hit exactly)"""
    h = (t1 - t0) / n_steps
    return t0 + h * np.arange(n_steps), np.full(n_steps, h)


def piecewise_schedule(breaks, steps_per_piece, t0, t1):
    """uniform steps inside each [breaks[i], breaks[i+1]] (fractions of the time range)"""
    ts, hs = [], []
    for (a, b), n in zip(zip(breaks[:-1], breaks[1:]), steps_per_piece):
        h = (b - a) * (t1 - t0) / n
        ts.append(t0 + a * (t1 - t0) + h * np.arange(n))
        hs.append(np.full(n, h))
    return np.concatenate(ts), np.concatenate(hs)


def graded_schedule(t0, t1, flat=20, ramp=120, ramp_first=240):
    """default py_driver_2d schedule: the year is cut into the 60 hist intervals
    (py_driver_2d/model_state.py:81); `flat` steps per interval where the mixed-layer depth is
    constant, `ramp` steps per interval while it moves (0.25-0.35 yr and 0.65-0.75 yr,
    py_driver_2d/vert_mix.py:98-99) and `ramp_first` in the first interval of each ramp.
    2640 steps/yr by default: meets the reference's CI tolerance (rtol 1e-3, atol 1e-6) on the
    final state AND on all 61 hist snapshots of baselines/ci_py_driver_2d_iage."""
    counts = [flat] * 60
    for k in list(range(15, 21)) + list(range(39, 45)):
        counts[k] = ramp
    counts[15] = counts[39] = ramp_first
    return piecewise_schedule([k / 60.0 for k in range(61)], counts, t0, t1)


class Model:
    """one tracer module on one grid: owns the device tables (nkb_model handle)"""

    def __init__(self, desc, keepalive):
        require_cuda()
        self.lib = _lib.load()
        self.desc = desc
        self._keepalive = keepalive
        self.nz, self.ny, self.T = desc.nz, desc.ny, desc.n_tracers
        self.n = self.T * self.nz * self.ny
        self.t0, self.t1 = desc.t0, desc.t1
        self.handle = ctypes.c_void_p()
        check(self.lib.nkb_model_create(ctypes.byref(self.handle), ctypes.byref(desc)), "nkb_model_create")
        self.n_steps = 0
        self.t_start = None
        self.h = None
        self._work = None

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.nkb_model_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:  # pylint: disable=broad-except
            pass

    # ---- schedule -------------------------------------------------------------------
    def set_schedule(self, t_start, h):
        t_start = np.ascontiguousarray(t_start, dtype=np.float64)
        h = np.ascontiguousarray(h, dtype=np.float64)
        check(
            self.lib.nkb_model_set_schedule(self.handle, len(h), dptr(t_start), dptr(h)),
            "nkb_model_set_schedule",
        )
        self.n_steps, self.t_start, self.h = len(h), t_start, h

    def set_uniform_schedule(self, n_steps):
        self.set_schedule(*uniform_schedule(n_steps, self.t0, self.t1))

    def set_graded_schedule(self, **kw):
        self.set_schedule(*graded_schedule(self.t0, self.t1, **kw))

    def step_index_of_times(self, times):
        """step indices (0..n_steps) whose END time equals each requested time"""
        ends = np.concatenate(([self.t0], self.t_start[1:], [self.t1]))
        idx = []
        for t in times:
            i = int(np.argmin(np.abs(ends - t)))
            if abs(ends[i] - t) > 1.0e-9 * max(1.0, abs(self.t1 - self.t0)):
                raise ValueError(f"time {t} is not a step boundary of the schedule")
            idx.append(i)
        return idx

    # ---- device-side operations -------------------------------------------------------
    def state_shape(self, B):
        return (self.T, self.nz, self.ny, padded_members(B))

    def mixing_coeff(self, time):
        out = torch.empty((self.nz - 1, self.ny), dtype=torch.float64, device="cuda")
        check(self.lib.nkb_model_mixing_coeff(self.handle, float(time), out.data_ptr(), _stream_ptr()),
              "nkb_model_mixing_coeff")
        return out

    def tend(self, time, x, B):
        """dc/dt(time, x) for a member-fastest batch x [T, nz, ny, ldb]"""
        ldb = x.shape[-1]
        out = torch.zeros_like(x)
        check(self.lib.nkb_model_tend(self.handle, float(time), x.data_ptr(), out.data_ptr(), B, ldb, _stream_ptr()),
              "nkb_model_tend")
        return out

    def eval(self, x, B, out=None, hist_steps=None):
        """F(x) = x(T) - x(0) for a member-fastest batch x [T, nz, ny, ldb] (device tensor).
        Returns F (and the [n_hist, T, nz, ny] history of member 0 when hist_steps is given)."""
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
        ldb = x.shape[-1]
        if out is None:
            out = torch.empty_like(x)
        need = self.lib.nkb_model_work_doubles(self.handle, B, ldb)
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.float64, device="cuda")
        n_hist, steps_arr, hist = 0, None, None
        hist_ptr = None
        if hist_steps is not None:
            # the library fills the slots in the order it meets the steps: the list must be strictly
            # increasing and within [0, n_steps] or snapshots would land in the wrong slots
            hs = [int(s) for s in hist_steps]
            if any(b <= a for a, b in zip(hs[:-1], hs[1:])) or (hs and (hs[0] < 0 or hs[-1] > self.n_steps)):
                raise ValueError("hist_steps must be strictly increasing and within [0, n_steps]")
            # the final state is x0 + F; the library records steps < n_steps
            inner = [s for s in hs if s < self.n_steps]
            n_hist = len(inner)
            steps_arr = (ctypes.c_int * max(n_hist, 1))(*inner)
            hist = torch.full((len(hist_steps), self.T, self.nz, self.ny), float("nan"), dtype=torch.float64, device="cuda")
            hist_ptr = hist.data_ptr()
        check(
            self.lib.nkb_model_eval(self.handle, x.data_ptr(), out.data_ptr(), self._work.data_ptr(), B, ldb,
                                    n_hist, steps_arr, hist_ptr, _stream_ptr()),
            "nkb_model_eval",
        )
        if hist_steps is not None:
            for i, s in enumerate(hist_steps):
                if s == self.n_steps:
                    hist[i] = x[..., 0] + out[..., 0]
            return out, hist
        return out

    def check_health(self):
        """after a synchronisation: raise if the persistent step kernel reported a dependency time-out"""
        if self.lib.nkb_model_poll_error(self.handle):
            raise _lib.NkbError("the persistent step kernel timed out waiting for a tile: results are invalid")

    def eval_host(self, x_host, out_host=None):
        """F for member-major host arrays [B, T, nz, ny] (pinned torch tensors or numpy)"""
        xt = x_host if isinstance(x_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_host))
        B = xt.shape[0]
        if out_host is None:
            out_host = torch.empty_like(xt)
        check(self.lib.nkb_model_eval_host(self.handle, xt.data_ptr(), out_host.data_ptr(), B), "nkb_model_eval_host")
        return out_host


def pack(x_major):
    """member-major device tensor [B, ...] -> member-fastest [..., ldb]"""
    lib = _lib.load()
    B = x_major.shape[0]
    n = x_major[0].numel()
    ldb = padded_members(B)
    out = torch.zeros(tuple(x_major.shape[1:]) + (ldb,), dtype=torch.float64, device="cuda")
    check(lib.nkb_pack_members(x_major.contiguous().data_ptr(), out.data_ptr(), n, B, ldb, _stream_ptr()),
          "nkb_pack_members")
    return out


def unpack(x_fast, B):
    """member-fastest device tensor [..., ldb] -> member-major [B, ...]"""
    lib = _lib.load()
    ldb = x_fast.shape[-1]
    n = x_fast.numel() // ldb
    out = torch.empty((B,) + tuple(x_fast.shape[:-1]), dtype=torch.float64, device="cuda")
    check(lib.nkb_unpack_members(x_fast.data_ptr(), out.data_ptr(), n, B, ldb, _stream_ptr()), "nkb_unpack_members")
    return out


class RegionWeights:
    """CSR region-mean matrix on the device (model_config.py:292-315) + region ids per cell"""

    def __init__(self, region_mask, grid_weight):
        require_cuda()
        mask = np.where(grid_weight == 0.0, 0, region_mask).astype(np.int32)
        wgt = np.where(mask == 0, 0.0, grid_weight).astype(np.float64)
        self.region_cnt = int(mask.max())
        flat_m, flat_w = mask.reshape(-1), wgt.reshape(-1)
        indices, indptr, data = [], [0], []
        for r in range(self.region_cnt):
            idx = np.nonzero(flat_m == r + 1)[0]
            indices.extend(idx.tolist())
            indptr.append(len(indices))
            row = flat_w[idx]
            inv = 1.0 / sum(row)
            data.extend((inv * row).tolist())
        self.ncell = flat_m.size
        self.mask_host = mask
        self.indptr = torch.tensor(indptr, dtype=torch.int32, device="cuda")
        self.indices = torch.tensor(indices, dtype=torch.int32, device="cuda")
        self.data = torch.tensor(data, dtype=torch.float64, device="cuda")
        self.region = torch.tensor(flat_m, dtype=torch.int32, device="cuda")
        self.max_row = int(np.diff(indptr).max())
        # dense form of the same weights (the fused Gram-Schmidt kernel walks the cells in storage order)
        cellw = np.zeros(self.ncell)
        cellw[np.asarray(indices, dtype=np.int64)] = data
        self.cellw = torch.tensor(cellw, dtype=torch.float64, device="cuda")
        self._mgs_scratch = None

    def dot(self, a, b, B):
        """[region_cnt, B] region-weighted dot products of member-fastest a, b [T, cells..., ldb];
        b=None gives the region means of a"""
        lib = _lib.load()
        ldb = a.shape[-1]
        T = a.shape[0]
        nch = lib.nkb_wdot_chunks(self.max_row)
        partial = torch.empty(nch * self.region_cnt * B, dtype=torch.float64, device="cuda")
        out = torch.empty((self.region_cnt, B), dtype=torch.float64, device="cuda")
        check(
            lib.nkb_wdot(self.indptr.data_ptr(), self.indices.data_ptr(), self.data.data_ptr(), self.region_cnt, T,
                         self.ncell, a.data_ptr(), None if b is None else b.data_ptr(), B, ldb, partial.data_ptr(),
                         nch, out.data_ptr(), _stream_ptr()),
            "nkb_wdot",
        )
        return out

    def axpby(self, alpha, x, beta, y, B, fill_alpha=1.0, fill_beta=1.0):
        """y <- alpha[r, b]*x + beta[r, b]*y in place (alpha/beta: None -> fill value everywhere,
        float -> that value, or device tensors [region_cnt, B])"""
        lib = _lib.load()
        ldb = y.shape[-1]
        T = y.shape[0]

        def prep(v, fill):
            if v is None:
                return None, fill
            if isinstance(v, (int, float)):
                return None, float(v)
            return v.contiguous(), fill

        al, fa = prep(alpha, fill_alpha)
        be, fb = prep(beta, fill_beta)
        # scalars given as plain floats apply everywhere (also outside regions)
        region_ptr = self.region.data_ptr()
        check(
            lib.nkb_axpby(region_ptr, self.region_cnt, T, self.ncell, None if al is None else al.data_ptr(),
                          None if x is None else x.data_ptr(), None if be is None else be.data_ptr(), y.data_ptr(),
                          fa, fb, B, ldb, _stream_ptr()),
            "nkb_axpby",
        )
        return y


    def mgs(self, w, basis, B):
        """modified Gram-Schmidt of w (in place) against the list `basis` of member-fastest tensors
        (model_state_base.py:365-377); returns the device tensor h [k, region_cnt, B].  One launch when w fits on
        the chip; no host synchronisation either way."""
        lib = _lib.load()
        k = len(basis)
        T, ldb = w.shape[0], w.shape[-1]
        h = torch.empty((k, self.region_cnt, B), dtype=torch.float64, device="cuda")
        if k == 0:
            return h
        need = lib.nkb_mgs_scratch_doubles(self.region_cnt, B, self.max_row)
        if self._mgs_scratch is None or self._mgs_scratch.numel() < need:
            self._mgs_scratch = torch.empty(need, dtype=torch.float64, device="cuda")
        ptrs = (ctypes.c_void_p * k)(*[v.data_ptr() for v in basis])
        check(
            lib.nkb_mgs(self.indptr.data_ptr(), self.indices.data_ptr(), self.data.data_ptr(), self.region.data_ptr(),
                        self.cellw.data_ptr(), self.region_cnt, T, self.ncell, self.max_row, w.data_ptr(), ptrs, k, B,
                        ldb, self._mgs_scratch.data_ptr(), self._mgs_scratch.numel(), h.data_ptr(), _stream_ptr()),
            "nkb_mgs",
        )
        return h

    def lin_comb(self, coeff, basis, B, add=None, out=None, fill=1.0):
        """out = sum_i coeff[i][r, b] * basis[i] (+ add) in one pass (model_state_base.py:619-624);
        coeff: device tensor [k, region_cnt, B]"""
        lib = _lib.load()
        k = len(basis)
        T, ldb = basis[0].shape[0], basis[0].shape[-1]
        if out is None:
            out = torch.empty_like(basis[0])
        ptrs = (ctypes.c_void_p * k)(*[v.data_ptr() for v in basis])
        check(
            lib.nkb_lin_comb(self.region.data_ptr(), self.region_cnt, T, self.ncell, coeff.contiguous().data_ptr(), ptrs,
                             k, None if add is None else add.data_ptr(), out.data_ptr(), float(fill), B, ldb,
                             _stream_ptr()),
            "nkb_lin_comb",
        )
        return out

    def limiter_scalef(self, base, inc, lob, upb, B):
        """[region_cnt, B] largest scale factors in [0, 1] keeping base + scalef*inc in [lob, upb]
        (utils.py:561-600); raises ValueError when base itself is out of bounds"""
        lib = _lib.load()
        ldb = base.shape[-1]
        T = base.shape[0]
        out = torch.full((self.region_cnt, B), float("inf"), dtype=torch.float64, device="cuda")
        flag = torch.zeros(padded_members(B) + 1, dtype=torch.int32, device="cuda")
        check(
            lib.nkb_limiter_scalef(self.region.data_ptr(), self.region_cnt, T, self.ncell, base.data_ptr(),
                                   inc.data_ptr(), 0.0 if lob is None else float(lob), 0 if lob is None else 1,
                                   0.0 if upb is None else float(upb), 0 if upb is None else 1, B, ldb,
                                   out.data_ptr(), flag.data_ptr(), _stream_ptr()),
            "nkb_limiter_scalef",
        )
        bits = flag[:B].cpu().numpy()
        # comp_scalef_lob / comp_scalef_upb (utils.py:561-600): an error only when a bound has to be enforced
        # (some base + increment violates it) and base itself violates it already
        if (((bits & 1) != 0) & ((bits & 2) != 0)).any():
            raise ValueError("base < lob")
        if (((bits & 4) != 0) & ((bits & 8) != 0)).any():
            raise ValueError("base > upb")
        return out


class BandedFactor:
    """member-shared banded LU (nkb_banded)"""

    def __init__(self, ab, kl, ku):
        require_cuda()
        self.lib = _lib.load()
        ab = np.ascontiguousarray(ab, dtype=np.float64)
        assert ab.shape[0] == kl + ku + 1
        self.n = ab.shape[1]
        self.handle = ctypes.c_void_p()
        check(self.lib.nkb_banded_create(ctypes.byref(self.handle), self.n, kl, ku, dptr(ab)), "nkb_banded_create")
        self.n_blocks = self.lib.nkb_banded_blocks(self.handle)  # independent diagonal blocks (solved in parallel)
        self.path = {3: "panel", 2: "thomas", 1: "window"}.get(self.lib.nkb_banded_path(self.handle), "?")

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.nkb_banded_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:  # pylint: disable=broad-except
            pass

    def solve(self, y, B, scale=1.0, subtract_rhs=False):
        """x = A^-1 (scale*y) [- y]  for member-fastest y [n, ldb]"""
        ldb = y.shape[-1]
        x = torch.empty_like(y)
        check(self.lib.nkb_banded_solve(self.handle, y.data_ptr(), x.data_ptr(), B, ldb, float(scale),
                                        1 if subtract_rhs else 0, _stream_ptr()), "nkb_banded_solve")
        return x
