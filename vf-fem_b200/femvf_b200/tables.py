"""
Host-side (setup-time) index tables for the device assembly.

DOLFIN builds the equivalent structures (dofmap, sparsity pattern, facet-cell
connectivity, Dirichlet dof lists) when a form is first assembled
(``/root/reference/src/femvf/models/assemblyutils.py:49-50``,
``residuals/base.py:54-65``).  They depend on the mesh only, so they are computed
once with numpy and uploaded; nothing here runs inside a time step.

The canonical CSR pattern: rows are node-major interleaved vector
DOFs, columns ascending, one full d x d block per vertex pair sharing a cell,
explicit zeros kept.  It is represented by the node graph ``(brptr, bcol)``; the
scalar ``(rowptr, colidx)`` is derived from it.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def node_graph(nn: int, cells: np.ndarray):
    nen = cells.shape[1]
    ii = np.repeat(cells, nen, axis=1).ravel()
    jj = np.tile(cells, (1, nen)).ravel()
    g = sp.coo_matrix((np.ones(len(ii), dtype=np.int8), (ii, jj)), shape=(nn, nn)).tocsr()
    g.sort_indices()
    return g.indptr.astype(np.int32), g.indices.astype(np.int32)


def scalar_csr_from_graph(brptr: np.ndarray, bcol: np.ndarray, d: int):
    """Scalar CSR ``(rowptr, colidx)`` implied by the node graph and block size ``d``."""
    nn = len(brptr) - 1
    deg = np.diff(brptr).astype(np.int64)
    rowptr = np.zeros(d * nn + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.repeat(deg * d, d))
    node_of_blk = np.repeat(np.arange(nn), deg)
    k_in_row = np.arange(len(bcol)) - brptr[node_of_blk].astype(np.int64)
    colidx = np.empty(rowptr[-1], dtype=np.int32)
    for a in range(d):
        start = rowptr[d * node_of_blk + a] + k_in_row * d
        for b in range(d):
            colidx[start + b] = d * bcol + b
    return rowptr.astype(np.int32), colidx


def _csr_from_pairs(keys: np.ndarray, vals: np.ndarray, n: int):
    order = np.lexsort((vals, keys))
    keys = keys[order]
    vals = vals[order]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, keys + 1, 1)
    return np.cumsum(ptr).astype(np.int32), vals.astype(np.int32)


def build_tables(coords, cells, pf_cell, pf_opp, fixed_dofs):
    """
    Parameters
    ----------
    coords : (nn, d) float64;  cells : (ne, d+1) int
    pf_cell, pf_opp : parent cell and opposite local vertex of every exterior facet in
        the 'pressure' subdomain (``residuals/solid.py:179-180``)
    fixed_dofs : vector DOFs in the closure of the 'fixed' facets (``residuals/base.py:47-65``)
    """
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.int64)
    nn, d = coords.shape
    ne, nen = cells.shape
    if ne >= 2**29:
        raise ValueError("too many cells for the packed adjacency tables")
    brptr, bcol = node_graph(nn, cells)
    rowptr, colidx = scalar_csr_from_graph(brptr, bcol, d)

    e_idx = np.repeat(np.arange(ne, dtype=np.int64), nen)
    a_idx = np.tile(np.arange(nen, dtype=np.int64), ne)
    n2e_ptr, n2e = _csr_from_pairs(cells.ravel(), e_idx * 4 + a_idx, nn)

    pf_cell = np.asarray(pf_cell, dtype=np.int64).reshape(-1)
    pf_opp = np.asarray(pf_opp, dtype=np.int64).reshape(-1)
    nfp = len(pf_cell)
    if nfp:
        f_idx = np.repeat(np.arange(nfp, dtype=np.int64), nen)
        fa_idx = np.tile(np.arange(nen, dtype=np.int64), nfp)
        n2f_ptr, n2f = _csr_from_pairs(cells[pf_cell].ravel(), f_idx * 4 + fa_idx, nn)
    else:
        n2f_ptr, n2f = np.zeros(nn + 1, dtype=np.int32), np.zeros(0, dtype=np.int32)

    bc = np.zeros(d * nn, dtype=np.uint8)
    bc[np.asarray(fixed_dofs, dtype=np.int64)] = 1

    return {
        'dim': d, 'nn': nn, 'ne': ne, 'nfp': nfp,
        'xyz': np.ascontiguousarray(coords.T),  # SoA (d, nn)
        'cells': np.ascontiguousarray(cells.T.astype(np.int32)),  # SoA (nen, ne)
        'brptr': brptr, 'bcol': bcol, 'rowptr': rowptr, 'colidx': colidx,
        'n2e_ptr': n2e_ptr, 'n2e': n2e, 'n2f_ptr': n2f_ptr, 'n2f': n2f,
        'pf_cell': pf_cell.astype(np.int32), 'pf_opp': pf_opp.astype(np.int32),
        'bc': bc,
    }


def tile_partition(brptr: np.ndarray, d: int, nodes_per_tile: int, max_tile_values: int):
    """
    Split the node range into contiguous tiles of at most ``nodes_per_tile`` nodes whose
    block rows hold at most ``max_tile_values`` doubles (the shared-memory CSR slice of
    one CTA).  Returns the tile start nodes (ntiles + 1).
    """
    nn = len(brptr) - 1
    vals = d * d * brptr.astype(np.int64)
    starts = [0]
    i = 0
    while i < nn:
        hi = min(i + nodes_per_tile, nn)
        # largest j <= hi with vals[j] - vals[i] <= max_tile_values
        j = int(np.searchsorted(vals, vals[i] + max_tile_values, side='right')) - 1
        j = min(j, hi)
        if j <= i:
            raise ValueError("a single block row exceeds the shared-memory tile")
        starts.append(j)
        i = j
    return np.asarray(starts, dtype=np.int32)
