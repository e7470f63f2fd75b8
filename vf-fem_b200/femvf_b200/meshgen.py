"""
Synthetic mesh generation for the benchmark configurations.

The reference meshes its M5_CB vocal-fold geometry with gmsh from STEP files
(``/root/reference/meshes/genmesh_M5_CB.py:10-33``); neither gmsh nor any ``.msh``
file is available, so the outline of ``meshes/stp/M5_CB_GA0.STEP`` (SURVEY.md
App. B, units cm) is re-meshed here with a Delaunay triangulation of a point
cloud.  The same physical groups are produced: cells 'body'/'cover', facets
'pressure'/'fixed', vertices 'separation-inf'/'separation-sup'.

Host-side setup code only (numpy/scipy); nothing here runs in a time step.
"""

from __future__ import annotations

import numpy as np

from .mesh import Mesh, MeshFunction

# --- M5_CB outline -----------------------------------------------------------

_BODY_TAG, _COVER_TAG = 2, 1
_PRESSURE_TAG, _FIXED_TAG = 3, 4
_SEP_INF_TAG, _SEP_SUP_TAG = 5, 6


def _arc(center, radius, p_start, p_end, h):
    """Points on the shorter arc from ``p_start`` to ``p_end`` (excluding the end point)."""
    c = np.asarray(center)
    a0 = np.arctan2(p_start[1] - c[1], p_start[0] - c[0])
    a1 = np.arctan2(p_end[1] - c[1], p_end[0] - c[0])
    da = (a1 - a0 + np.pi) % (2 * np.pi) - np.pi
    n = max(int(np.ceil(abs(da) * radius / h)), 1)
    t = a0 + da * np.arange(n) / n
    return np.stack([c[0] + radius * np.cos(t), c[1] + radius * np.sin(t)], axis=1)


def _line(p0, p1, h):
    p0 = np.asarray(p0, dtype=float)
    p1 = np.asarray(p1, dtype=float)
    n = max(int(np.ceil(np.linalg.norm(p1 - p0) / h)), 1)
    t = np.arange(n)[:, None] / n
    return p0 + t * (p1 - p0)


def m5_outline(h: float):
    """Counter-clockwise boundary points of the M5_CB_GA0 outline and per-segment tags."""
    P0 = (0.0, 0.0)
    P1 = (0.7895, 0.0)
    P2 = (0.7895, 0.4013)
    P3 = (0.6908, 0.5)
    P4 = (0.4895, 0.5)
    P5 = (0.37459, 0.44642)
    segs = [
        (_line(P0, P1, h), _FIXED_TAG),
        (_line(P1, P2, h), _PRESSURE_TAG),
        (_arc((0.6908, 0.4013), 0.0987, P2, P3, h), _PRESSURE_TAG),
        (_line(P3, P4, h), _PRESSURE_TAG),
        (_arc((0.4895, 0.35), 0.15, P4, P5, h), _PRESSURE_TAG),
        (_line(P5, P0, h), _PRESSURE_TAG),
    ]
    pts = np.concatenate([s for s, _ in segs], axis=0)
    tags = np.concatenate([np.full(len(s), t) for s, t in segs])
    return pts, tags


def _cover_interface(h: float):
    """Polyline of the body/cover interface."""
    Q0 = (0.06527, 0.0)
    Q1 = (0.41289, 0.41428)
    Q2 = (0.4895, 0.45)
    Q3 = (0.6908, 0.45)
    Q4 = (0.7395, 0.4013)
    Q5 = (0.7395, 0.0)
    parts = [
        _line(Q0, Q1, h),
        _arc((0.4895, 0.35), 0.10, Q1, Q2, h),
        _line(Q2, Q3, h),
        _arc((0.6908, 0.4013), 0.0487, Q3, Q4, h),
        _line(Q4, Q5, h),
        np.array([Q5]),
    ]
    return np.concatenate(parts, axis=0)


def _points_in_polygon(pts, poly):
    """Even-odd rule point-in-polygon test (vectorised)."""
    x, y = pts[:, 0], pts[:, 1]
    inside = np.zeros(len(pts), dtype=bool)
    n = len(poly)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        cond = (y0 > y) != (y1 > y)
        with np.errstate(divide='ignore', invalid='ignore'):
            xint = (x1 - x0) * (y - y0) / (y1 - y0) + x0
        inside ^= cond & (x < xint)
    return inside


def _dist_to_polyline(pts, poly, closed=True):
    n = len(poly)
    dmin = np.full(len(pts), np.inf)
    last = n if closed else n - 1
    for i in range(last):
        a = poly[i]
        b = poly[(i + 1) % n]
        ab = b - a
        t = np.clip(((pts - a) @ ab) / (ab @ ab), 0.0, 1.0)
        proj = a + t[:, None] * ab
        dmin = np.minimum(dmin, np.linalg.norm(pts - proj, axis=1))
    return dmin


def m5_cb_mesh(h: float = 0.05):
    """
    Triangulate the M5_CB outline with target edge length ``h`` (cm).

    Returns the ``(mesh, mesh_functions, mesh_subdomains)`` tuple consumed by
    ``load.load_fsi_model`` (``/root/reference/src/femvf/load.py:45-54``).
    """
    from scipy.spatial import Delaunay

    bpts, btags = m5_outline(h)
    # interior: hexagonal lattice clipped to the outline, kept away from the boundary
    dy = h * np.sqrt(3.0) / 2.0
    ys = np.arange(dy * 0.5, 0.5, dy)
    rows = []
    for k, yv in enumerate(ys):
        xs = np.arange(((k % 2) * 0.5 + 0.25) * h, 0.7895, h)
        rows.append(np.stack([xs, np.full_like(xs, yv)], axis=1))
    ipts = np.concatenate(rows, axis=0)
    keep = _points_in_polygon(ipts, bpts) & (_dist_to_polyline(ipts, bpts) > 0.55 * h)
    ipts = ipts[keep]

    pts = np.concatenate([bpts, ipts], axis=0)
    tri = Delaunay(pts)
    cells = tri.simplices.astype(np.int64)
    x = pts[cells]
    e1, e2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
    area = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
    cen = x.mean(axis=1)
    good = (area > 1e-6 * h * h) & _points_in_polygon(cen, bpts)
    cells = cells[good]

    # drop unreferenced points (none expected) and renumber
    used = np.unique(cells)
    remap = -np.ones(len(pts), dtype=np.int64)
    remap[used] = np.arange(len(used))
    pts = pts[used]
    cells = remap[cells]
    nb = len(bpts)
    assert np.all(used[:nb] == np.arange(nb)), "a boundary point was dropped"

    mesh = Mesh(pts, cells)
    return _tag_m5(mesh, nb, btags, h)


def _tag_m5(mesh: Mesh, nb: int, btags: np.ndarray, h: float):
    d = 2
    facets = mesh.facets
    facet_mf = MeshFunction(1, np.zeros(len(facets), dtype=np.int64))
    ext = mesh.exterior_facets
    # exterior facet (i, i+1 mod nb) carries the tag of boundary point i
    f = facets[ext]
    lo, hi = f[:, 0], f[:, 1]
    wrap = (lo == 0) & (hi == nb - 1)
    start = np.where(wrap, hi, lo)
    ok = wrap | (hi == lo + 1)
    if not np.all(ok) or np.any(hi >= nb):
        raise RuntimeError("exterior facets do not follow the outline")
    facet_mf.array()[ext] = btags[start]

    interface = _cover_interface(h * 0.25)
    body_poly = np.concatenate([interface, np.array([[0.06527, 0.0]])], axis=0)
    cen = mesh.coordinates()[mesh.cells()].mean(axis=1)
    in_body = _points_in_polygon(cen, body_poly)
    cell_mf = MeshFunction(2, np.where(in_body, _BODY_TAG, _COVER_TAG))

    vert_mf = MeshFunction(0, np.zeros(mesh.num_vertices(), dtype=np.int64))
    x = mesh.coordinates()
    i_inf = int(np.argmin(np.linalg.norm(x[:nb] - np.array([0.4895, 0.5]), axis=1)))
    i_sup = int(np.argmin(np.linalg.norm(x[:nb] - np.array([0.6908, 0.5]), axis=1)))
    vert_mf[i_inf] = _SEP_INF_TAG
    vert_mf[i_sup] = _SEP_SUP_TAG

    mesh_functions = (vert_mf, facet_mf, cell_mf)
    mesh_subdomains = (
        {'separation-inf': _SEP_INF_TAG, 'separation-sup': _SEP_SUP_TAG},
        {'pressure': _PRESSURE_TAG, 'fixed': _FIXED_TAG},
        {'body': _BODY_TAG, 'cover': _COVER_TAG},
    )
    return mesh, mesh_functions, mesh_subdomains


# --- refinement / renumbering ----------------------------------------------------


def refine_red(mesh_tuple):
    """Uniform red refinement of a tagged triangle mesh (each triangle -> 4)."""
    mesh, mfs, subdomains = mesh_tuple
    if mesh.topology().dim() != 2:
        raise ValueError("red refinement implemented for triangles")
    x = mesh.coordinates()
    c = mesh.cells()
    facets = mesh.facets
    nn = len(x)
    cf = mesh.cell_facets  # facet opposite local vertex k
    mid = nn + cf  # midpoint node ids per (cell, local facet)
    xm = x[facets].mean(axis=1)
    newx = np.concatenate([x, xm], axis=0)
    v0, v1, v2 = c[:, 0], c[:, 1], c[:, 2]
    m0, m1, m2 = mid[:, 0], mid[:, 1], mid[:, 2]  # m0 on edge (v1,v2), m1 on (v0,v2), m2 on (v0,v1)
    newc = np.concatenate([
        np.stack([v0, m2, m1], axis=1),
        np.stack([m2, v1, m0], axis=1),
        np.stack([m1, m0, v2], axis=1),
        np.stack([m0, m1, m2], axis=1),
    ], axis=0)
    new_mesh = Mesh(newx, newc)

    vert_mf, facet_mf, cell_mf = mfs
    new_cell_mf = MeshFunction(2, np.tile(cell_mf.array(), 4))
    new_vert = np.zeros(len(newx), dtype=np.int64)
    new_vert[:nn] = vert_mf.array()
    new_vert_mf = MeshFunction(0, new_vert)

    # child facets of a tagged parent facet inherit its tag: (lo, mid), (mid, hi)
    nf_new = new_mesh.facets
    new_facet = np.zeros(len(nf_new), dtype=np.int64)
    tagged = np.nonzero(facet_mf.array() != 0)[0]
    if len(tagged):
        key = {}
        for k, fidx in enumerate(tagged):
            a, b = facets[fidx]
            m = nn + fidx
            t = facet_mf.array()[fidx]
            key[(min(a, m), max(a, m))] = t
            key[(min(b, m), max(b, m))] = t
        # vectorised lookup through a sorted structured view
        keys = np.array(list(key.keys()), dtype=np.int64)
        vals = np.array(list(key.values()), dtype=np.int64)
        big = nf_new[:, 0] * (len(newx) + 1) + nf_new[:, 1]
        kbig = keys[:, 0] * (len(newx) + 1) + keys[:, 1]
        order = np.argsort(kbig)
        pos = np.searchsorted(kbig[order], big)
        pos = np.clip(pos, 0, len(kbig) - 1)
        hit = kbig[order][pos] == big
        new_facet[hit] = vals[order][pos[hit]]
    new_facet_mf = MeshFunction(1, new_facet)
    return new_mesh, (new_vert_mf, new_facet_mf, new_cell_mf), subdomains


def morton_order(x: np.ndarray, bits: int = 16) -> np.ndarray:
    """Permutation sorting points along a Z-order (Morton) space-filling curve."""
    d = x.shape[1]
    lo, hi = x.min(axis=0), x.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((x - lo) / span * (2**bits - 1)).astype(np.uint64), 2**bits - 1)
    code = np.zeros(len(x), dtype=np.uint64)
    for b in range(bits):
        for k in range(d):
            code |= ((q[:, k] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * d + k)
    return np.argsort(code, kind='stable')


def renumber_for_locality(mesh_tuple, method: str = 'morton'):
    """
    Renumber vertices along a space-filling curve (compact node tiles: a contiguous index
    range is a blob, not a strip, so the cells touching it are mostly interior) or by
    reverse Cuthill-McKee, and sort cells by their lowest vertex, so that gathers of nodal
    data and the CSR rows written by one CTA are close in memory.  Tags follow their
    entities.
    """
    mesh, mfs, subdomains = mesh_tuple
    d = mesh.topology().dim()
    x = mesh.coordinates()
    c = mesh.cells()
    nn = len(x)
    if method == 'morton':
        perm = morton_order(x, 16 if d == 2 else 10).astype(np.int64)
    else:
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        e = mesh.edges
        g = coo_matrix((np.ones(2 * len(e)), (np.r_[e[:, 0], e[:, 1]], np.r_[e[:, 1], e[:, 0]])),
                       shape=(nn, nn)).tocsr()
        perm = np.asarray(reverse_cuthill_mckee(g, symmetric_mode=True), dtype=np.int64)
    inv = np.empty(nn, dtype=np.int64)
    inv[perm] = np.arange(nn)
    newx = x[perm]
    newc = inv[c]
    corder = np.argsort(newc.min(axis=1), kind='stable')
    newc = newc[corder]

    old_facets = mesh.facets
    old_facet_tags = mfs[d - 1].array()
    new_mesh = Mesh(newx, newc)
    new_cell_mf = MeshFunction(d, mfs[d].array()[corder])
    new_vert_mf = None
    if mfs[0] is not None and mfs[0].dim() == 0:
        v = np.zeros(nn, dtype=np.int64)
        v[inv] = mfs[0].array()
        new_vert_mf = MeshFunction(0, v)

    nf_new = new_mesh.facets
    new_facet = np.zeros(len(nf_new), dtype=np.int64)
    tagged = np.nonzero(old_facet_tags != 0)[0]
    if len(tagged):
        tf = np.sort(inv[old_facets[tagged]], axis=1)
        base = nn + 1
        if d == 2:
            kbig = tf[:, 0] * base + tf[:, 1]
            big = nf_new[:, 0] * base + nf_new[:, 1]
        else:
            kbig = (tf[:, 0] * base + tf[:, 1]) * base + tf[:, 2]
            big = (nf_new[:, 0] * base + nf_new[:, 1]) * base + nf_new[:, 2]
        order = np.argsort(kbig)
        pos = np.clip(np.searchsorted(kbig[order], big), 0, len(kbig) - 1)
        hit = kbig[order][pos] == big
        new_facet[hit] = old_facet_tags[tagged][order][pos[hit]]
    new_facet_mf = MeshFunction(d - 1, new_facet)

    out_mfs = list(mfs)
    out_mfs[d] = new_cell_mf
    out_mfs[d - 1] = new_facet_mf
    if new_vert_mf is not None:
        out_mfs[0] = new_vert_mf
    if d == 3 and len(out_mfs) > 1 and out_mfs[1] is not None and out_mfs[1].dim() == 1:
        out_mfs[1] = MeshFunction(1, np.zeros(new_mesh.num_entities(1), dtype=np.int64))
    return new_mesh, tuple(out_mfs), subdomains


def m5_cb_refined(h: float, levels: int, renumber: bool = True):
    """M5_CB mesh at size ``h`` red-refined ``levels`` times (config 3 of BASELINE.json)."""
    mt = m5_cb_mesh(h)
    for _ in range(levels):
        mt = refine_red(mt)
    if renumber:
        mt = renumber_for_locality(mt)
    return mt


# --- extrusion to tetrahedra ------------------------------------------------------


def extrude_to_tets(mesh_tuple, length: float, nz: int):
    """
    Extrude a tagged triangle mesh along z into prisms split into 3 tets each
    (config 5 of BASELINE.json; fixed z-faces as in ``tests/fixture_mesh.py:74-83``).

    The prism split uses the global vertex order so that neighbouring prisms
    agree on the diagonals of their shared quadrilateral faces.
    """
    mesh, mfs, subdomains = mesh_tuple
    x2 = mesh.coordinates()
    c2 = mesh.cells()
    nn2 = len(x2)
    zs = np.linspace(0.0, length, nz + 1)
    coords = np.concatenate(
        [np.concatenate([x2, np.full((nn2, 1), z)], axis=1) for z in zs], axis=0
    )
    cs = np.sort(c2, axis=1)  # a<b<c global order -> consistent diagonals
    tets = []
    for k in range(nz):
        a0, b0, c0 = (cs[:, j] + k * nn2 for j in range(3))
        a1, b1, c1 = (cs[:, j] + (k + 1) * nn2 for j in range(3))
        tets.append(np.stack([a0, b0, c0, c1], axis=1))
        tets.append(np.stack([a0, b0, b1, c1], axis=1))
        tets.append(np.stack([a0, a1, b1, c1], axis=1))
    tets = np.concatenate(tets, axis=0)
    mesh3 = Mesh(coords, tets)

    cell_mf = MeshFunction(3, np.tile(np.repeat(mfs[2].array()[None, :], 3, axis=0).reshape(-1), nz))
    # facet tags: z end caps -> fixed; lateral faces inherit the tag of the 2D edge
    f3 = mesh3.facets
    facet_tag = np.zeros(len(f3), dtype=np.int64)
    fixed = subdomains[1]['fixed']
    ext = mesh3.exterior_facets
    xf = coords[f3[ext]]
    zmin = xf[:, :, 2].min(axis=1)
    zmax = xf[:, :, 2].max(axis=1)
    cap = (zmax < 1e-12) | (zmin > length - 1e-12)
    e2 = mesh.facets
    etag = mfs[1].array()
    base = nn2 + 1
    ekey = e2[:, 0] * base + e2[:, 1]
    order = np.argsort(ekey)
    v2 = np.sort(f3[ext] % nn2, axis=1)  # project vertices to the 2D mesh
    # a lateral triangle projects to two distinct 2D vertices
    lo = v2[:, 0]
    hi = v2[:, 2]
    key = lo * base + hi
    pos = np.clip(np.searchsorted(ekey[order], key), 0, len(ekey) - 1)
    hit = (ekey[order][pos] == key) & ~cap
    tags_ext = np.zeros(len(ext), dtype=np.int64)
    tags_ext[hit] = etag[order][pos[hit]]
    tags_ext[cap] = fixed
    facet_tag[ext] = tags_ext
    facet_mf = MeshFunction(2, facet_tag)
    edge_mf = MeshFunction(1, np.zeros(mesh3.num_entities(1), dtype=np.int64))
    mesh_functions = (None, edge_mf, facet_mf, cell_mf)
    mesh_subdomains = ({}, {}, dict(subdomains[1]), dict(subdomains[2]))
    return mesh3, mesh_functions, mesh_subdomains
