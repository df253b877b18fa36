"""CPU tests: the C-ABI library loads here (no GPU) and exports every symbol declared in
include/nkb200.h; ctypes prototypes and the header agree; calls fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "nkb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nkb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nk_ooc_b200 import _lib

    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nkb200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype in _lib.SYMBOLS"
    for name in _lib.SYMBOLS:
        assert name in declared, f"{name} bound in _lib.py but not declared in include/nkb200.h"
    assert lib.nkb_version() >= 100


def test_model_desc_layout_matches_header():
    """sizeof(nkb_model_desc) as compiled by gcc from the header == ctypes.sizeof(ModelDesc)"""
    import subprocess
    import tempfile

    from nk_ooc_b200 import _lib

    src = '#include <stdio.h>\n#include "nkb200.h"\nint main(void){printf("%zu\\n", sizeof(nkb_model_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as tmp:
        cfile = os.path.join(tmp, "s.c")
        open(cfile, "w").write(src)
        exe = os.path.join(tmp, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe], check=True)
        size = int(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    assert size == ctypes.sizeof(_lib.ModelDesc)


def test_calls_fail_loudly_without_cuda():
    """no CPU fallback: on a box without a GPU the product path raises"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from nk_ooc_b200 import _lib
    from nk_ooc_b200.py_driver_2d import modules
    from nk_ooc_b200.spatial_axis import spatial_axis_from_defn

    depth = spatial_axis_from_defn("depth", nlevs=6, edge_end=4000.0, delta_ratio_max=19.0)
    ypos = spatial_axis_from_defn("ypos", nlevs=5, edge_end=50.0e5, delta_ratio_max=1.0)
    tr = modules.Transport2D(depth, ypos)
    with pytest.raises(_lib.NkbError):
        modules.iage_model(tr)


def test_bad_arguments_are_rejected_with_a_message():
    from nk_ooc_b200 import _lib

    lib = _lib.load()
    assert lib.nkb_model_create(None, None) != 0
    assert b"null" in lib.nkb_last_error()
    assert lib.nkb_pack_members(None, None, 0, 0, 0, None) != 0
    assert lib.nkb_wdot(None, None, None, 1, 1, 1, None, None, 1, 1, None, 1, None, None) != 0


def test_host_set_up_matches_oracle():
    """host-side time-invariant fields (velocity, Peclet-limited mixing, explicit stencil)"""
    from nk_ooc_b200.py_driver_2d.modules import Transport2D
    from nk_ooc_b200.spatial_axis import SpatialAxis, edges_from_defn
    from oracle import imex_oracle as im
    from oracle import nk_oracle as o

    ze = edges_from_defn(30, 0.0, 4000.0, 19.0)
    ye = edges_from_defn(30, 0.0, 50.0e5, 1.0)
    np.testing.assert_array_equal(ze, o.stretched_edges(30, 0.0, 4000.0, 19.0))
    tr = Transport2D(SpatialAxis("depth", ze), SpatialAxis("ypos", ye))
    g = o.Grid2D(ze, ye)
    np.testing.assert_array_equal(tr.advection.vvel, g.vvel)
    np.testing.assert_array_equal(tr.advection.wvel, g.wvel)
    np.testing.assert_array_equal(tr.horiz_mix.mixing_coeff, g.hmix)
    e_l, e_c, e_r = im.explicit_stencil_2d(g)
    np.testing.assert_array_equal(tr.estencil, np.stack([e_l, e_c, e_r]))
    # the stencil form equals the reference's flux form (advection + horizontal mixing)
    rng = np.random.default_rng(0)
    c = rng.normal(size=(1, 30, 30))
    full = g.transport_tend(0.0, c)
    flux_form = full - _vertical_part(g, c)  # cancellation: compare at the scale of the full tendency
    sten = e_c * c[0]
    sten[:, 1:] += e_l[:, 1:] * c[0, :, :-1]
    sten[:, :-1] += e_r[:, :-1] * c[0, :, 1:]
    np.testing.assert_allclose(sten, flux_form[0], rtol=0, atol=1e-13 * np.abs(full).max())


def _vertical_part(g, c):
    from oracle import imex_oracle as im

    sub, diag, sup = im.implicit_tridiag_2d(g, 0.0)
    out = diag[None] * c
    out[:, 1:] += sub[None, 1:] * c[:, :-1]
    out[:, :-1] += sup[None, :-1] * c[:, 1:]
    return out


def test_spatial_axis_roundtrip(tmp_path):
    from nk_ooc_b200.spatial_axis import spatial_axis_from_defn, spatial_axis_from_file

    ax = spatial_axis_from_defn("depth", nlevs=20)
    fname = str(tmp_path / "depth_axis.nc")
    ax.dump(fname, "test")
    back = spatial_axis_from_file(fname, "depth")
    np.testing.assert_array_equal(back.edges, ax.edges)
    assert back.units == "m" and len(back) == 20
    vals = np.ones((3, 20))
    np.testing.assert_allclose(ax.int_vals_mid(vals, -1), 900.0)
    with pytest.raises(ValueError):
        ax.int_vals_mid(np.ones(19), 0)
