"""
P2 (6-node) triangle assembly on the device: host side of ``vf_p2_*`` (``csrc/p2.cu``).

The reference is P1 only (``/root/reference/src/femvf/equations/form.py:521-524``); this is the
P2 extension BASELINE.json names (north_star subsystem 1, configs[2]: "~1M P2 triangles").  It
stands in for the same ``dfn.assemble`` calls (``models/assemblyutils.py:49-50``,
``models/transient.py:363-406``) on a P2 space: Newmark inertia, linear elastic and Kelvin-Voigt
viscous terms, follower pressure on the 'pressure' edges, Dirichlet rows on the 'fixed' edges.

``P2Assembler`` takes the same mesh tuple as the P1 residuals, builds the mid-edge nodes and the
tables of the kernel, and assembles into caller-visible device tensors.  No CPU fallback.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import check
from . import tables as _tables

# local edge (pair of local vertices) of the mid-edge nodes 3, 4, 5 (DOLFIN / UFC order)
_EDGE_LOCAL = np.array([[1, 2], [0, 2], [0, 1]])
# mid-edge local node of the edge between two local vertices
_MID_OF = np.array([[-1, 5, 4], [5, -1, 3], [4, 3, -1]])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def p2_nodes(coords: np.ndarray, cells: np.ndarray, interleave: bool = True):
    """Mid-edge nodes for a P1 triangle mesh.  Returns (coords6, cells6, vertex_ids, edges,
    edge_node): with ``interleave`` the nodes are renumbered so that a mid-edge node sits next to
    its lower vertex (locality of the block rows), else they are appended after the vertices."""
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    nv = coords.shape[0]
    ev = np.sort(cells[:, _EDGE_LOCAL], axis=2)
    key = ev[..., 0] * nv + ev[..., 1]
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    edges = np.stack([uniq // nv, uniq % nv], axis=1)
    ne_ = len(edges)
    if interleave:
        # order: vertex v, then the edges whose lower vertex is v
        sort_key = np.concatenate([2 * np.arange(nv), 2 * edges[:, 0] + 1])
        order = np.argsort(sort_key, kind='stable')
        new_id = np.empty(nv + ne_, dtype=np.int64)
        new_id[order] = np.arange(nv + ne_)
    else:
        new_id = np.arange(nv + ne_)
    vertex_ids = new_id[:nv]
    edge_node = new_id[nv:]
    coords6 = np.empty((nv + ne_, 2))
    coords6[vertex_ids] = coords
    coords6[edge_node] = 0.5 * (coords[edges[:, 0]] + coords[edges[:, 1]])
    cells6 = np.concatenate([vertex_ids[cells], edge_node[inv.reshape(-1, 3)]], axis=1)
    return coords6, cells6, vertex_ids, edges, edge_node


def build_p2_tables(coords, cells, pfacets=None, pfacet_cells=None, fixed_edges=None,
                    interleave: bool = True) -> dict:
    """Host tables of the P2 kernels (arguments as for ``P2Assembler``): mid-edge nodes, node
    graph, (node, cell) pairs with the CSR slots of the cell's six nodes, pressure-edge records,
    Dirichlet flags and the class-sorted thread -> node map.  ``keep`` is the argument list of
    ``vf_p2_create`` in order."""
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    nv = coords.shape[0]
    c6, cells6, vid, edges, enode = p2_nodes(coords, cells, interleave)
    nn, ne = c6.shape[0], cells6.shape[0]
    brptr, bcol = _tables.node_graph(nn, cells6)
    brptr = brptr.astype(np.int64); bcol = bcol.astype(np.int64)
    deg = np.diff(brptr)
    if deg.max() > 31:
        raise ValueError("a P2 node couples to more than 31 nodes")
    # (node, cell) pairs grouped by node
    pair_node = cells6.ravel()
    pair_ref = (np.repeat(np.arange(ne), 6) * 8 + np.tile(np.arange(6), ne))
    order = np.argsort(pair_node, kind='stable')
    n2e = pair_ref[order]
    n2e_ptr = np.zeros(nn + 1, dtype=np.int64)
    np.add.at(n2e_ptr, pair_node + 1, 1)
    n2e_ptr = np.cumsum(n2e_ptr)
    # CSR slots of the six nodes of every pair's cell in the pair's node row
    pn = pair_node[order]
    gkey = np.repeat(np.arange(nn), deg) * nn + bcol
    slots = np.zeros(len(n2e), dtype=np.uint64)
    for b in range(6):
        nb = cells6[n2e >> 3, b]
        s = np.searchsorted(gkey, pn * nn + nb) - brptr[pn]
        slots |= s.astype(np.uint64) << np.uint64(5 * b)
    # pressure edges
    nfp = 0 if pfacets is None else len(pfacets)
    if nfp:
        pf = np.asarray(pfacets, dtype=np.int64).reshape(-1, 2)
        pc = np.asarray(pfacet_cells, dtype=np.int64)
        tri = cells[pc]
        loc = np.argmax(tri[:, None, :] == pf[:, :, None], axis=2)          # (nfp, 2)
        mid = _MID_OF[loc[:, 0], loc[:, 1]]
        pf_loc = np.concatenate([loc, mid[:, None]], axis=1)
        opp = tri[np.arange(nfp), 3 - loc[:, 0] - loc[:, 1]]
        t = coords[pf[:, 1]] - coords[pf[:, 0]]
        length = np.linalg.norm(t, axis=1)
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / length[:, None]
        sgn = np.sign(((coords[pf[:, 0]] - coords[opp]) * nrm).sum(axis=1))
        pf_geo = np.concatenate([nrm * sgn[:, None], length[:, None]], axis=1)
        fnodes = cells6[pc[:, None], pf_loc]                                 # (nfp, 3) P2 ids
        f_node = fnodes.ravel()
        f_ref = np.repeat(np.arange(nfp), 3) * 4 + np.tile(np.arange(3), nfp)
        f_cell = np.repeat(pc, 3)
        fo = np.argsort(f_node, kind='stable')
        n2f = f_ref[fo]
        n2f_ptr = np.zeros(nn + 1, dtype=np.int64)
        np.add.at(n2f_ptr, f_node + 1, 1)
        n2f_ptr = np.cumsum(n2f_ptr)
        # index of the (node, parent cell) pair in n2e
        pair_key = pn * ne + (n2e >> 3)                                      # sorted by node
        n2f_pair = np.searchsorted(pair_key, f_node[fo] * ne + f_cell[fo])
    else:
        pc = np.zeros(0, np.int64); pf_loc = np.zeros((0, 3), np.int64)
        pf_geo = np.zeros((0, 3)); n2f = np.zeros(0, np.int64)
        n2f_ptr = np.zeros(nn + 1, np.int64); n2f_pair = np.zeros(0, np.int64)
    # Dirichlet nodes: closure of the fixed edges
    fixed = np.zeros(nn, dtype=np.uint8)
    if fixed_edges is not None and len(fixed_edges):
        fe = np.sort(np.asarray(fixed_edges, dtype=np.int64).reshape(-1, 2), axis=1)
        ekey = edges[:, 0] * nv + edges[:, 1]
        idx = np.searchsorted(ekey, fe[:, 0] * nv + fe[:, 1])
        fixed[vid[fe.ravel()]] = 1
        fixed[enode[idx]] = 1
    # thread -> node map: vertex nodes, then mid-edge nodes (each class in node order)
    node_order = np.concatenate([np.sort(vid), np.sort(enode)])
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    keep = [np.ascontiguousarray(c6), i32(cells6), i32(brptr), i32(bcol), i32(n2e_ptr),
            i32(n2e), np.ascontiguousarray(slots.astype(np.uint32)), i32(n2f_ptr), i32(n2f),
            i32(n2f_pair), i32(pc), i32(pf_loc), np.ascontiguousarray(pf_geo),
            np.ascontiguousarray(fixed), i32(node_order)]
    return dict(coords6=c6, cells6=cells6, vertex_ids=vid, edges=edges, edge_node=enode, nn=nn,
                ne=ne, nv=nv, nfp=nfp, brptr=brptr, bcol=bcol,
                fixed_nodes=np.nonzero(fixed)[0], keep=keep)


class P2Assembler:
    def __init__(self, coords, cells, pfacets=None, pfacet_cells=None, fixed_edges=None,
                 device=None, interleave: bool = True):
        """
        coords (nv, 2), cells (ne, 3): positively oriented P1 triangles
        pfacets (nfp, 2): vertex ids of the 'pressure' edges, pfacet_cells (nfp,) parent cells
        fixed_edges (nfix, 2): vertex ids of the Dirichlet ('fixed') edges
        """
        self._lib = _cabi.load_library()
        if not torch.cuda.is_available() or self._lib.vf_device_count() <= 0:
            raise _cabi.VFError("femvf_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        coords = np.asarray(coords, dtype=np.float64)
        cells = np.asarray(cells, dtype=np.int64)
        self.nv = coords.shape[0]
        T = build_p2_tables(coords, cells, pfacets, pfacet_cells, fixed_edges, interleave)
        self.coords6, self.cells6 = T['coords6'], T['cells6']
        self.vertex_ids, self.edges, self.edge_node = T['vertex_ids'], T['edges'], T['edge_node']
        nn, ne, nfp = T['nn'], T['ne'], T['nfp']
        self.nn, self.ne, self.N = nn, ne, 2 * nn
        self.fixed_nodes = T['fixed_nodes']
        self.brptr, self.bcol = T['brptr'], T['bcol']
        keep = T['keep']
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            check(self._lib.vf_p2_create(
                nn, ne, _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]), _ptr(keep[4]),
                _ptr(keep[5]), _ptr(keep[6]), _ptr(keep[7]), _ptr(keep[8]), _ptr(keep[9]), nfp,
                _ptr(keep[10]), _ptr(keep[11]), _ptr(keep[12]), _ptr(keep[13]), _ptr(keep[14]),
                self.nv, stream, C.byref(handle)))
        self._h = handle
        self.nnz = int(self._lib.vf_p2_nnz(self._h))
        f64 = dict(dtype=torch.float64, device=self.device)
        self.F = torch.zeros(self.N, **f64)
        self.J = torch.zeros(self.nnz, **f64)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            self._lib.vf_p2_destroy(h)
            self._h = None

    def csr_pattern(self):
        """Scalar CSR (indptr, indices) of the node-major interleaved DOFs."""
        return _tables.scalar_csr_from_graph(self.brptr.astype(np.int32),
                                             self.bcol.astype(np.int32), 2)

    def assemble(self, u1, u0, v0, a0, p1, emod, eta, rho, nu: float, dt: float,
                 res: bool = True, jac: bool = True):
        """All vectors are fp64 device tensors: nodal (2 nn) / (nn) / per cell (ne).  Results in
        ``self.F`` and ``self.J``."""
        flags = (1 if res else 0) | (2 if jac else 0)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(self._lib.vf_p2_assemble(
            self._h, flags, float(dt), float(nu), emod.data_ptr(), eta.data_ptr(), rho.data_ptr(),
            u1.data_ptr(), u0.data_ptr(), v0.data_ptr(), a0.data_ptr(), p1.data_ptr(),
            self.F.data_ptr(), self.J.data_ptr(), stream))
        return self.F, self.J
