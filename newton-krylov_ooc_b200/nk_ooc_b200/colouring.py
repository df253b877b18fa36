"""Column regions, index maps and distance-2 colouring of the perturbation probes.

Product-side implementation (vectorised: sparse boolean matrix products instead of the
notebook's per-cell Python sets) of notebooks/IRF_coloring_dev.ipynb of the reference:
connectivity `conn_nd` (cell 4-5: MOM6 9-point stencil, periodic in i, tripole seam at the last
j row) or any list of index offsets, distance-2 connectivity `conn2_nd` (cell 6-7), greedy
first-fit colouring in C order / reverse / largest-degree-first (cells 9-13), int32 index maps
`nd_to_flat` / `flat_to_nd` (cell 17), DIMACS export (cell 19) and the reader + properness check
of an external (gCol HybridEA) solution (cell 23).  All results are integers and must be
bit-identical to the notebook's; tests/test_colouring.py checks them against the loop-based
restatement in oracle/nk_oracle.py and the notebook's printed answers.

Two cells of the same colour are more than two stencil steps apart, so unit perturbations placed
on all cells (here: all columns) of one colour can be evaluated in ONE member of a batched
function evaluation and their responses separated afterwards (`probe_batch`, `decode_probes`).
"""

import numpy as np
from scipy import sparse


def index_maps(mask):
    """(nd_to_flat int32 [-1 where masked], flat_to_nd int32 [n_cells, ndim]), C order over
    mask != 0 (IRF_coloring_dev.ipynb cell 17)"""
    mask = np.asarray(mask)
    flat_to_nd = np.argwhere(mask != 0).astype(np.int32)
    nd_to_flat = np.full(mask.shape, -1, dtype=np.int32)
    nd_to_flat[mask != 0] = np.arange(len(flat_to_nd), dtype=np.int32)
    return nd_to_flat, flat_to_nd


def _wrap_mom6(shape, idx):
    """ind_wrap of cell 4 on index arrays [n, ndim]: periodic in the last dimension, tripole fold
    across the end of the second-to-last one; returns (indices, in_bounds)"""
    idx = idx.copy()
    ni, nj = shape[-1], shape[-2]
    i, j = idx[:, -1], idx[:, -2]
    i = np.where(i < 0, i + ni, i)
    i = np.where(i >= ni, i - ni, i)
    fold = j >= nj
    i = np.where(fold, ni - 1 - i, i)
    j = np.where(fold, 2 * nj - 1 - j, j)
    idx[:, -1], idx[:, -2] = i, j
    ok = np.ones(len(idx), dtype=bool)
    for d, n in enumerate(shape):
        ok &= (idx[:, d] >= 0) & (idx[:, d] < n)
    return idx, ok


def _shift_matrix(mask, nd_to_flat, flat_to_nd, offset, wrap):
    """boolean CSR S with S[a, b] = 1 iff cell b = cell a + offset (after wrapping) and both wet"""
    n = len(flat_to_nd)
    tgt = flat_to_nd.astype(np.int64) + np.asarray(offset, dtype=np.int64)[None, -mask.ndim:]
    if wrap == "mom6":
        tgt, ok = _wrap_mom6(mask.shape, tgt)
    else:
        ok = np.ones(n, dtype=bool)
        for d, size in enumerate(mask.shape):
            ok &= (tgt[:, d] >= 0) & (tgt[:, d] < size)
    rows = np.nonzero(ok)[0]
    cols = nd_to_flat[tuple(tgt[rows].T)]
    keep = cols >= 0
    return sparse.csr_matrix((np.ones(keep.sum(), dtype=np.int8), (rows[keep], cols[keep])), shape=(n, n))


def connectivity_offsets(mask, offsets, wrap=None):
    """conn: cells reachable by one of `offsets` (the cell itself is NOT included unless (0,..) is)"""
    mask = np.asarray(mask)
    nd_to_flat, flat_to_nd = index_maps(mask)
    n = len(flat_to_nd)
    conn = sparse.csr_matrix((n, n), dtype=np.int8)
    for off in offsets:
        conn = conn + _shift_matrix(mask, nd_to_flat, flat_to_nd, off, wrap)
    return _boolean(conn)


def connectivity_mom6(mask):
    """gen_conn_nd of cell 4: 3x3 in the horizontal built as (x then y) | (y then x) through wet
    cells only, plus the 5-point stencil of the layers above and below for 3-D masks"""
    mask = np.asarray(mask)
    nd_to_flat, flat_to_nd = index_maps(mask)
    nd = mask.ndim

    def s(off):
        return _shift_matrix(mask, nd_to_flat, flat_to_nd, (0,) * (nd - len(off)) + tuple(off), "mom6")

    sx = _boolean(s((0, -1)) + s((0, 0)) + s((0, 1)))
    sy = _boolean(s((-1, 0)) + s((0, 0)) + s((1, 0)))
    conn = _boolean(sx @ sy + sy @ sx)
    if nd == 3:
        for dk in (-1, 1):
            for off in ((dk, -1, 0), (dk, 0, -1), (dk, 0, 0), (dk, 0, 1), (dk, 1, 0)):
                # the notebook does not wrap the layer index: out-of-range layers are dropped by the
                # bounds check of apply_ind_offsets
                conn = conn + _shift_matrix(mask, nd_to_flat, flat_to_nd, off, "mom6")
        conn = _boolean(conn)
    return conn


def _boolean(mat):
    mat = sparse.csr_matrix(mat)
    mat.data[:] = 1
    mat.sum_duplicates()
    mat.sort_indices()
    return mat.astype(np.int8)


def distance2(conn):
    """conn2[a] = union of conn[b] over b in conn[a] (cell 6); contains a itself when a is in conn[a]"""
    return _boolean(conn @ conn)


def greedy_colouring(conn2, order=None):
    """first-fit colouring (cells 9-13): vertices visited in `order` (default C order of the flat
    index) get the smallest colour (0-based) absent from their conn2 neighbourhood.
    Returns int32 [n_cells]."""
    n = conn2.shape[0]
    indptr, indices = conn2.indptr, conn2.indices
    colour = np.full(n, -1, dtype=np.int32)
    order = range(n) if order is None else order
    for v in order:
        used = colour[indices[indptr[v]:indptr[v + 1]]]
        used = used[used >= 0]
        if used.size == 0:
            colour[v] = 0
            continue
        present = np.zeros(used.max() + 2, dtype=bool)
        present[used] = True
        colour[v] = int(np.argmin(present))
    return colour


def degree_order(conn2):
    """largest conn2 neighbourhood first, ties in C order (cell 13: list.sort is stable)"""
    deg = np.diff(conn2.indptr)
    return np.argsort(-deg, kind="stable")


def to_nd(mask, flat_vals, fill=-1):
    out = np.full(np.asarray(mask).shape, fill, dtype=np.int32)
    out[np.asarray(mask) != 0] = flat_vals
    return out


def dimacs_lines(conn2, comment="adjacent graph for IRF tracers"):
    """DIMACS 'p edge n m' + 'e i j' (1-based, i < j) in the notebook's order: by vertex, then by
    neighbour index ascending (cell 19 iterates Python sets, whose order is not defined; the edge
    SET is what gCol reads)"""
    coo = sparse.triu(conn2, k=1).tocsr()
    coo.sort_indices()
    lines = [f"c {comment}", f"p edge {conn2.shape[0]} {coo.nnz}"]
    for i in range(coo.shape[0]):
        for j in coo.indices[coo.indptr[i]:coo.indptr[i + 1]]:
            lines.append(f"e {i + 1} {j + 1}")
    return lines


def read_solution(lines, conn2):
    """external colouring (gCol solution.txt: a header line, then one colour per vertex in flat
    order; cell 23).  Raises ValueError on an improper colouring."""
    vals = np.array([int(v) for v in lines[1:1 + conn2.shape[0]]], dtype=np.int32)
    if len(vals) != conn2.shape[0]:
        raise ValueError("solution has too few vertices")
    check_proper(vals, conn2)
    return vals


def check_proper(colour, conn2):
    coo = sparse.triu(conn2, k=1).tocoo()
    bad = np.nonzero(colour[coo.row] == colour[coo.col])[0]
    if bad.size:
        raise ValueError(f"improper coloring at flat index {int(coo.row[bad[0]])}")


# ---- probes of a batched function evaluation ------------------------------------------------------
def column_colouring(ny, reach=1):
    """colours of the ypos columns of py_driver_2d for perturbation probes: columns interact over
    `reach` columns per stencil application (3-point y stencil: reach 1), two probes may share a
    member when they are more than 2*reach columns apart.  Greedy first-fit on the 1-D distance-2
    graph = j mod (2*reach + 1)."""
    mask = np.ones((1, ny), dtype=np.int32)
    offs = [(0, d) for d in range(-reach, reach + 1)]
    conn2 = distance2(connectivity_offsets(mask, offs))
    return greedy_colouring(conn2)


def probe_batch(x0, colour_of_column, eps):
    """members of a batched evaluation that probe every (tracer, level) of every column:
    member m = (c*T + t)*nz + k carries x0 + eps at tracer t, level k of ALL columns of colour c.
    x0 [T, nz, ny] -> [n_colours*T*nz, T, nz, ny] (member-major host layout)"""
    T, nz, ny = x0.shape
    ncol = int(colour_of_column.max()) + 1
    out = np.broadcast_to(x0, (ncol * T * nz,) + x0.shape).copy()
    for c in range(ncol):
        cols = np.nonzero(colour_of_column == c)[0]
        for t in range(T):
            for k in range(nz):
                out[(c * T + t) * nz + k, t, k, cols] += eps
    return out


def decode_probes(f0, fprobe, colour_of_column, eps, reach=1):
    """column-block Jacobian from the probe responses: jac[j][(t', k'), (t, k)] = d F[t', k', j] /
    d x[t, k, j] and the coupling blocks to the columns j-reach..j+reach:
    returns [ny, 2*reach+1, T*nz, T*nz] (neighbour index reach = the column itself)"""
    T, nz, ny = f0.shape
    n = T * nz
    colour_of_column = np.asarray(colour_of_column)
    ncol = int(colour_of_column.max()) + 1
    # responses of all probes at once: [colour, probe (t, k), response (t', k'), column]
    resp = ((np.asarray(fprobe) - f0[np.newaxis]) / eps).reshape(ncol, n, n, ny)
    jac = np.zeros((ny, 2 * reach + 1, n, n))
    cols = np.arange(ny)
    for d in range(-reach, reach + 1):
        jv = cols[(cols + d >= 0) & (cols + d < ny)]
        # jac[j, d + reach, r, p] = resp[colour[j], p, r, j + d]
        jac[jv, d + reach] = np.transpose(resp[colour_of_column[jv], :, :, jv + d], (0, 2, 1))
    return jac
