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


def order_fans_2d(cells: np.ndarray, n2e_ptr: np.ndarray, n2e: np.ndarray, nn: int):
    """
    Reorder every vertex's adjacent-cell list counter-clockwise around the vertex (cells are
    positively oriented), so that consecutive cells share the edge (vertex, prev(cell)) ==
    (vertex, next(next cell)).  Open fans (boundary vertices) start at the cell whose ``next``
    edge is a boundary edge.  Returns (n2e_reordered, ok); ``ok`` is False when some vertex
    star is not a single fan (non-manifold vertex), in which case the input order is kept.
    """
    n2e_ptr = n2e_ptr.astype(np.int64)
    npair = len(n2e)
    e, a = n2e.astype(np.int64) >> 2, n2e.astype(np.int64) & 3
    node = np.repeat(np.arange(nn), np.diff(n2e_ptr))
    nxt = cells[e, (a + 1) % 3]
    prv = cells[e, (a + 2) % 3]
    # successor of pair q: the pair of the same vertex whose next vertex is prev(q)
    key_next = node * nn + nxt
    order = np.argsort(key_next, kind='stable')
    ks = key_next[order]
    if np.any(ks[1:] == ks[:-1]):
        return n2e, False                      # an edge with two cells on the same side
    want = node * nn + prv
    pos = np.searchsorted(ks, want)
    pos_c = np.minimum(pos, npair - 1)
    has_succ = ks[pos_c] == want
    succ = np.where(has_succ, order[pos_c], -1)
    has_pred = np.zeros(npair, dtype=bool)
    has_pred[succ[has_succ]] = True
    # start of each fan: the pair without predecessor (open fan) or the first listed (closed)
    start = n2e_ptr[:-1].copy()
    nopred = np.nonzero(~has_pred)[0]
    cnt = np.zeros(nn, dtype=np.int64)
    np.add.at(cnt, node[nopred], 1)
    if np.any(cnt > 1):
        return n2e, False                      # more than one open fan at a vertex
    start[node[nopred]] = nopred
    deg = np.diff(n2e_ptr)
    rank = -np.ones(npair, dtype=np.int64)
    cur = start.copy()
    alive = deg > 0
    for r in range(int(deg.max()) if npair else 0):
        idx = np.nonzero(alive & (r < deg))[0]
        c = cur[idx]
        good = c >= 0
        if not np.all(good):
            return n2e, False
        rank[c] = r
        cur[idx] = succ[c]
    if np.any(rank < 0):
        return n2e, False
    new_pos = n2e_ptr[node] + rank
    if len(np.unique(new_pos)) != npair:
        return n2e, False
    out = np.empty_like(n2e)
    out[new_pos] = n2e
    return out, True


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
    fan_ok = False
    if d == 2:
        n2e, fan_ok = order_fans_2d(cells, n2e_ptr, n2e, nn)

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
        'bc': bc, 'fan_ok': bool(fan_ok),
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


def build_tile_elem_tables(T: dict, tile_start: np.ndarray):
    """
    Tables of the two-phase tile assembly kernel (``asm_tile2_kernel``, triangles only).

    te_ptr/te_elem : cells touching each node tile (owned + halo), ascending
    te_quad        : per (tile, cell): the LOCAL slots of the cell's 3 vertices in the tile's
                     staged vertex list (own vertices first, then ``tile_halo``) + the cell id
    tile_halo      : halo vertices (global ids) of each tile, ascending per tile
    tile_desc      : 8 int32 per tile (see below)
    pair_info      : one uint32 per (node, adjacent cell) pair in ``n2e`` order:
                     bits [0,12) index of the cell in its node's tile list, [12,14) local
                     index a of the node in the cell, [14,20) [20,26) [26,32) CSR slots (within
                     the node's block row) of the cell's vertices a, (a+1)%3, (a+2)%3
    Returns None when the packing limits (4096 cells per tile, 64 blocks per row) do not hold.
    """
    d, nn, ne = T['dim'], T['nn'], T['ne']
    if d != 2:
        return None
    cells = np.ascontiguousarray(T['cells'].T.astype(np.int64))  # (ne, 3)
    brptr = T['brptr'].astype(np.int64)
    bcol = T['bcol'].astype(np.int64)
    if np.max(np.diff(brptr)) >= 64:
        return None
    ntiles = len(tile_start) - 1
    tile_of_node = np.searchsorted(tile_start.astype(np.int64), np.arange(nn), side='right') - 1
    # (tile, cell) incidences, unique and sorted by tile then cell
    tc = tile_of_node[cells]  # (ne, 3)
    key = (tc * ne + np.arange(ne)[:, None]).ravel()
    key = np.unique(key)
    te_tile = key // ne
    te_elem = (key % ne).astype(np.int32)
    te_ptr = np.zeros(ntiles + 1, dtype=np.int64)
    np.add.at(te_ptr, te_tile + 1, 1)
    te_ptr = np.cumsum(te_ptr)
    max_tile_elems = int(np.max(np.diff(te_ptr)))
    if max_tile_elems >= 4096:
        return None

    # per (node, cell) pair
    n2e_ptr = T['n2e_ptr'].astype(np.int64)
    n2e = T['n2e'].astype(np.int64)
    pair_node = np.repeat(np.arange(nn), np.diff(n2e_ptr))
    pe, pa = n2e >> 2, n2e & 3
    ptile = tile_of_node[pair_node]
    # local index of the cell in the tile list: position of (tile, cell) in the sorted keys
    local = np.searchsorted(key, ptile * ne + pe) - te_ptr[ptile]
    # CSR slot of each vertex of the cell in the node's block row
    gkey = np.repeat(np.arange(nn), np.diff(brptr)) * nn + bcol  # sorted globally
    slots = []
    for shift in range(3):   # the pair's own vertex, then next and prev in the cell's CCW order
        col = cells[pe, (pa + shift) % 3]
        pos = np.searchsorted(gkey, pair_node * nn + col)
        slots.append(pos - brptr[pair_node])
    info = (local.astype(np.uint64) | (pa.astype(np.uint64) << np.uint64(12))
            | (slots[0].astype(np.uint64) << np.uint64(14))
            | (slots[1].astype(np.uint64) << np.uint64(20))
            | (slots[2].astype(np.uint64) << np.uint64(26)))
    # halo vertices of each tile: vertices of the tile's cells outside its own node range,
    # ascending per tile.  The kernel stages own + halo vertices in shared memory, so a cell
    # addresses its vertices by LOCAL slot: own vertex v -> v - i0, halo -> nT + rank in the list
    ts = tile_start.astype(np.int64)
    tcells = cells[te_elem]                                   # (n_te, 3) global vertex ids
    vkey = np.unique((te_tile[:, None] * nn + tcells).ravel())
    vtile, vvert = vkey // nn, vkey % nn
    is_halo = tile_of_node[vvert] != vtile
    hkey = vkey[is_halo]                                      # sorted by tile, then vertex
    th_ptr = np.zeros(ntiles + 1, dtype=np.int64)
    np.add.at(th_ptr, vtile[is_halo] + 1, 1)
    th_ptr = np.cumsum(th_ptr)
    tile_halo = vvert[is_halo].astype(np.int32)
    nown = np.diff(ts)
    own = tile_of_node[tcells] == te_tile[:, None]
    slot_own = tcells - ts[te_tile][:, None]
    slot_halo = (np.searchsorted(hkey, te_tile[:, None] * nn + tcells)
                 - th_ptr[te_tile][:, None] + nown[te_tile][:, None])
    te_quad = np.empty((len(te_elem), 4), dtype=np.int32)
    te_quad[:, :3] = np.where(own, slot_own, slot_halo)
    te_quad[:, 3] = te_elem
    # flat per-tile descriptor: everything a CTA needs is addressable after ONE dependent load,
    # instead of a chain tile_start -> te_ptr -> te_elem -> cells.  Twelve int32 (8-11: prefetch hint):
    #   i0, te0, pair0, blk0, halo0, nT | nH << 16, n_cells | n_pairs << 16, n_blocks
    npairs = n2e_ptr[ts[1:]] - n2e_ptr[ts[:-1]]
    ncell = np.diff(te_ptr)
    nhalo = np.diff(th_ptr)
    if np.max(npairs) >= 65536 or np.max(nown + nhalo) >= 32768:
        return None
    # window between the 10% and 90% quantile of the tile's (ascending) cell ids: where the
    # material data of the bulk of its cells lives.  Only an L2 prefetch hint (descriptor
    # words 8 and 9)
    es = te_elem.astype(np.int64)
    n_b = np.diff(te_ptr)
    lo_i = te_ptr[:-1] + n_b // 10
    hi_i = te_ptr[:-1] + np.maximum(n_b - n_b // 10 - 1, 0)
    e_lo = es[lo_i]
    e_cnt = es[hi_i] - es[lo_i] + 1
    e_cnt = np.minimum(e_cnt, 4 * np.diff(te_ptr))   # a sparse window is not worth fetching
    desc = np.zeros((ntiles, 12), dtype=np.int64)
    desc[:, 8] = e_lo
    desc[:, 9] = e_cnt
    desc[:, 0] = ts[:-1]
    desc[:, 1] = te_ptr[:-1]
    desc[:, 2] = n2e_ptr[ts[:-1]]
    desc[:, 3] = brptr[ts[:-1]]
    desc[:, 4] = th_ptr[:-1]
    desc[:, 5] = nown | (nhalo << 16)
    desc[:, 6] = ncell | (npairs << 16)
    desc[:, 7] = brptr[ts[1:]] - brptr[ts[:-1]]
    desc = desc.astype(np.uint32).view(np.int32)
    max_tile_pairs = int(np.max(npairs))
    max_tile_verts = int(np.max(nown + nhalo))
    return {
        'te_ptr': te_ptr.astype(np.int32), 'te_elem': te_elem,
        'pair_info': info.astype(np.uint32), 'max_tile_elems': max_tile_elems,
        'tile_desc': np.ascontiguousarray(desc), 'te_quad': np.ascontiguousarray(te_quad),
        'max_tile_pairs': max_tile_pairs,
        'tile_halo': tile_halo if len(tile_halo) else np.zeros(1, np.int32),
        'n_tile_halo': int(len(tile_halo)), 'max_tile_verts': max_tile_verts,
    }


def build_fan_tables(T: dict, tile_nodes: int):
    """
    Tables of the node-centric fan assembly kernel (``asm_fan_kernel``, triangles with
    counter-clockwise vertex fans, ``T['fan_ok']``).

    Every vertex n owns its block row of J and its residual entries.  Its adjacent cells are
    walked counter-clockwise: cell j is (n, p_j, p_{j+1}), so walking the fan loads ONE new ring
    vertex per cell and every off-diagonal block (n, p_j) is the sum of two consecutive cells.
    Nodes are grouped in contiguous tiles of ``tile_nodes`` (the CTA size); a tile stages in
    shared memory its own vertices, the ring vertices outside its range (``halo``), the DG0
    properties of the cells it touches (``tcell``) and its slice of the CSR value array.

    ring : uint32, per tile a block of ``rows`` x ``tile_nodes`` words, row-major (word (r, t)
        belongs to the tile's t-th node: consecutive threads read consecutive words).  Row 0 is
        the node header
            (brptr[n] - brptr[i0]) | deg << 12 | self_slot << 17 | ncell << 22 | closed << 27
        rows 1 + j, j = 0..ncell are the ring vertices p_j
            staged slot of p_j | CSR slot of p_j in n's block row << 10 | local cell << 15
        where the local cell indexes the tile's ``tcell`` list and is that of cell
        j = (n, p_j, p_{j+1}) (0xfff for the last entry; a closed fan repeats p_0 there).
    tcell : int32, cells touched by each tile (ascending), every tile's list padded to an even
        count with repeats of its last cell (16-byte granularity of the bulk copies).
    halo : int32, ring vertices outside each tile's own range, ascending per tile.
    desc : int32 (ntiles, 12): i0, nT | nH << 16, halo0, ring0 (words), rows, tcell0,
        padded cell count, brptr[i0], number of blocks of the tile, 0, 0, 0.
    Returns None when the packing limits do not hold.
    """
    if T['dim'] != 2 or not T.get('fan_ok', False):
        return None
    nn, ne = T['nn'], T['ne']
    TN = int(tile_nodes)
    cells = np.ascontiguousarray(T['cells'].T.astype(np.int64))
    brptr = T['brptr'].astype(np.int64)
    bcol = T['bcol'].astype(np.int64)
    n2e_ptr = T['n2e_ptr'].astype(np.int64)
    n2e = T['n2e'].astype(np.int64)
    deg = np.diff(brptr)
    ncell = np.diff(n2e_ptr)
    if deg.max() >= 32 or ncell.max() >= 31 or np.any(ncell == 0):
        return None
    ntiles = -(-nn // TN)
    node = np.repeat(np.arange(nn), ncell)
    pe, pa = n2e >> 2, n2e & 3
    vp = cells[pe, (pa + 1) % 3]            # p_j of every (node, cell) pair
    vq = cells[pe, (pa + 2) % 3]            # p_{j+1}
    first = n2e_ptr[:-1]
    last = n2e_ptr[1:] - 1
    closed = vq[last] == vp[first]
    tile_of = np.arange(nn) // TN
    i0 = np.arange(ntiles, dtype=np.int64) * TN
    nT = np.minimum(i0 + TN, nn) - i0
    # cells touched by each tile, ascending; local index of every (node, cell) pair
    pt = tile_of[node]
    ckey = np.unique(pt * ne + pe)
    ctile = ckey // ne
    tc_cnt = np.zeros(ntiles, dtype=np.int64)
    np.add.at(tc_cnt, ctile, 1)
    tc_raw = np.zeros(ntiles + 1, dtype=np.int64)
    tc_raw[1:] = np.cumsum(tc_cnt)
    lcell_pair = np.searchsorted(ckey, pt * ne + pe) - tc_raw[pt]
    ncp = tc_cnt + (tc_cnt & 1)             # padded to an even count
    if ncp.max() >= 0xfff:
        return None
    tc_ptr = np.zeros(ntiles + 1, dtype=np.int64)
    tc_ptr[1:] = np.cumsum(ncp)
    tcell = np.empty(int(tc_ptr[-1]), dtype=np.int64)
    tcell[tc_ptr[ctile] + (np.arange(len(ckey)) - tc_raw[ctile])] = ckey % ne
    odd = np.nonzero(tc_cnt & 1)[0]
    tcell[tc_ptr[odd] + tc_cnt[odd]] = tcell[tc_ptr[odd] + tc_cnt[odd] - 1]
    # ring entries: ncell + 1 per node
    rptr = np.zeros(nn + 1, dtype=np.int64)
    rptr[1:] = np.cumsum(ncell + 1)
    nent = int(rptr[-1])
    ent_node = np.repeat(np.arange(nn), ncell + 1)
    ent_j = np.arange(nent) - rptr[ent_node]
    is_last = ent_j == ncell[ent_node]
    pair_of = n2e_ptr[ent_node] + np.minimum(ent_j, ncell[ent_node] - 1)
    ent_v = np.where(is_last, vq[pair_of], vp[pair_of])
    ent_cell = np.where(is_last, 0xfff, lcell_pair[pair_of])
    # CSR slot of the ring vertex in the node's block row
    gkey = np.repeat(np.arange(nn), deg) * nn + bcol
    cslot = np.searchsorted(gkey, ent_node * nn + ent_v) - brptr[ent_node]
    self_slot = np.searchsorted(gkey, np.arange(nn) * (nn + 1)) - brptr[:-1]
    # halo vertices per tile and staged slots
    ent_tile = tile_of[ent_node]
    halo_mask = tile_of[ent_v] != ent_tile
    hkey = np.unique(ent_tile[halo_mask] * nn + ent_v[halo_mask])
    htile, hvert = hkey // nn, hkey % nn
    th_ptr = np.zeros(ntiles + 1, dtype=np.int64)
    np.add.at(th_ptr, htile + 1, 1)
    th_ptr = np.cumsum(th_ptr)
    nH = np.diff(th_ptr)
    if (nT + nH).max() >= 1024:
        return None
    vslot = np.where(halo_mask,
                     np.searchsorted(hkey, ent_tile * nn + ent_v) - th_ptr[ent_tile] + nT[ent_tile],
                     ent_v - i0[ent_tile])
    bbase = brptr[i0]
    nblk = brptr[np.minimum(i0 + TN, nn)] - bbase
    rel_b0 = brptr[:-1] - bbase[tile_of]
    if rel_b0.max() >= 4096:
        return None
    # per-tile ring blocks
    rows_node = ncell + 2                                  # header + ncell + 1 entries
    rows = np.zeros(ntiles, dtype=np.int64)
    np.maximum.at(rows, tile_of, rows_node)
    ring0 = np.zeros(ntiles + 1, dtype=np.int64)
    ring0[1:] = np.cumsum(rows * TN)
    if ring0[-1] >= 2**31:
        return None
    ring = np.full(int(ring0[-1]), 0xfff << 15, dtype=np.uint32)
    t_in = np.arange(nn) - i0[tile_of]
    hdr = ring0[tile_of] + t_in
    ring[hdr] = (rel_b0 | (deg << 12) | (self_slot << 17) | (ncell << 22)
                 | (closed.astype(np.int64) << 27))
    pos = ring0[ent_tile] + (1 + ent_j) * TN + t_in[ent_node]
    ring[pos] = vslot | (cslot << 10) | (ent_cell << 15)
    desc = np.zeros((ntiles, 12), dtype=np.int64)
    desc[:, 0] = i0
    desc[:, 1] = nT | (nH << 16)
    desc[:, 2] = th_ptr[:-1]
    desc[:, 3] = ring0[:-1]
    desc[:, 4] = rows
    desc[:, 5] = tc_ptr[:-1]
    desc[:, 6] = ncp
    desc[:, 7] = bbase
    desc[:, 8] = nblk
    return {
        'tile_nodes': TN, 'ntiles': int(ntiles),
        'desc': np.ascontiguousarray(desc.astype(np.uint32).view(np.int32)),
        'ring': np.ascontiguousarray(ring),
        'tcell': np.ascontiguousarray(tcell.astype(np.int32)),
        'halo': hvert.astype(np.int32) if len(hvert) else np.zeros(1, np.int32),
        'n_halo': int(len(hvert)), 'max_verts': int((nT + nH).max()),
        'max_rows': int(rows.max()), 'max_cells': int(ncp.max()), 'max_blocks': int(nblk.max()),
    }


def color_node_graph(brptr: np.ndarray, bcol: np.ndarray, node0: int, node1: int, seed: int = 0):
    """
    Distance-1 colouring of the node graph restricted to nodes [node0, node1) for the multicolour
    block ILU(0) (``csrc/ilu.cu``): no two nodes of a colour are adjacent (share a cell).

    Jones-Plassmann rounds, vectorised: an uncoloured node whose random priority beats every
    uncoloured neighbour takes the smallest colour not used by its coloured neighbours.
    Returns (color (nn,) int32 with -1 outside the range, rows: the nodes of the range grouped by
    colour, color_ptr (ncolors + 1,) int32).
    """
    nn = len(brptr) - 1
    brptr = np.asarray(brptr, dtype=np.int64)
    bcol = np.asarray(bcol, dtype=np.int64)
    row = np.repeat(np.arange(nn, dtype=np.int64), np.diff(brptr))
    keep = (row >= node0) & (row < node1) & (bcol >= node0) & (bcol < node1) & (bcol != row)
    r, c = row[keep] - node0, bcol[keep] - node0
    n = node1 - node0
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, r + 1, 1)
    ptr = np.cumsum(ptr)
    has = ptr[1:] > ptr[:-1]
    start = ptr[:-1][has]
    rng = np.random.default_rng(seed)
    prio = rng.permutation(n).astype(np.int64) + 1
    color = np.full(n, -1, dtype=np.int64)
    while True:
        unc = color < 0
        if not unc.any():
            break
        nb_prio = np.where(color[c] < 0, prio[c], 0)
        maxn = np.zeros(n, dtype=np.int64)
        if len(start):
            maxn[has] = np.maximum.reduceat(nb_prio, start)
        cand = unc & (prio > maxn)
        bits = np.where(color[c] >= 0, np.left_shift(np.int64(1), np.maximum(color[c], 0)), 0)
        forb = np.zeros(n, dtype=np.int64)
        if len(start):
            forb[has] = np.bitwise_or.reduceat(bits, start)
        f = forb[cand]
        lowest_free = (~f) & (f + 1)            # isolates the lowest zero bit of f
        color[cand] = np.round(np.log2(lowest_free.astype(np.float64))).astype(np.int64)
    if color.max() >= 62:
        raise ValueError("node graph needs more than 62 colours")
    order = np.argsort(color, kind='stable')
    ncol = int(color.max()) + 1
    cptr = np.zeros(ncol + 1, dtype=np.int32)
    cptr[1:] = np.cumsum(np.bincount(color, minlength=ncol))
    full = np.full(nn, -1, dtype=np.int32)
    full[node0:node1] = color
    return full, np.ascontiguousarray((order + node0).astype(np.int32)), cptr
