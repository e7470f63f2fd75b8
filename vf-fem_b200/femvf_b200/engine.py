"""
Python owner of one device engine (``vf_engine`` of ``include/vffem_b200.h``).

PyTorch supplies the device memory (one ``uint8`` arena tensor), the stream and, for
ensembles sharded over several GPUs, ``torch.distributed``; every computation is a call
through the C ABI.  There is no CPU fallback: constructing an ``Engine`` without a CUDA
device raises.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import ARRAY_IDS, ProblemDesc, SolverOpts, check
from . import tables as _tables
from .solverconst import DEFAULT_NEWTON_SOLVER_PRM, DEFAULT_LINEAR_SOLVER_PRM

SCAL = {'nu': 0, 'ycontact': 1, 'kcontact': 2, 'ncontact': 3, 'ymid': 6, 'rayleigh_m': 7,
        'rayleigh_k': 8}
SCAL_COUNT = 10
DAMPING = {'kelvin_voigt': 0, 'rayleigh': 1}
FPROP = {'rho_air': 0, 'r_sep': 1, 'area_lb': 2, 'zeta_min': 3, 'zeta_sep': 4}
FPROP_COUNT = 5


def make_solver_opts(options: Optional[dict] = None, is_static: bool = False) -> SolverOpts:
    """Newton options with the reference's keys (``solverconst.py:1-6``); ``linear_solver``
    is accepted and ignored."""
    prm = dict(DEFAULT_NEWTON_SOLVER_PRM)
    lin = dict(DEFAULT_LINEAR_SOLVER_PRM)
    if options:
        for key, value in options.items():
            if key in lin:
                lin[key] = value
            else:
                prm[key] = value
    return SolverOpts(
        float(prm['absolute_tolerance']), float(prm['relative_tolerance']),
        int(prm['maximum_iterations']),
        float(lin['gmres_relative_tolerance']), float(lin['gmres_absolute_tolerance']),
        int(lin['gmres_maximum_iterations']), int(bool(is_static)),
        int(lin['polynomial_degree']), 0,
    )


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(
        self,
        tables: dict,
        s: Optional[np.ndarray] = None,
        fsi_solid: Optional[np.ndarray] = None,
        fsi_fluid: Optional[np.ndarray] = None,
        fluid_kind: int = 0,
        idx_sep: int = 0,
        contact: bool = False,
        membrane: bool = False,
        damping: str = 'kelvin_voigt',
        n_members: int = 1,
        gmres_restart: int = 40,
        device: Optional[torch.device] = None,
    ):
        self._lib = _cabi.load_library()
        if not torch.cuda.is_available() or self._lib.vf_device_count() <= 0:
            raise _cabi.VFError(
                "femvf_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.tables = tables
        d = tables['dim']
        self.dim, self.nn, self.ne = d, tables['nn'], tables['ne']
        self.N = d * self.nn
        self.nfp = int(tables['nfp'])
        self.n_members = int(n_members)

        import os
        if d == 2:
            nodes_per_tile = int(os.environ.get('VF_TILE_NODES', '80'))
            max_vals, tile_threads = 64 * nodes_per_tile, max(nodes_per_tile, 32)
        else:
            nodes_per_tile, max_vals, tile_threads = 64, 12288, 64
        tile_start = _tables.tile_partition(tables['brptr'], d, nodes_per_tile, max_vals)
        vals = d * d * tables['brptr'].astype(np.int64)
        tile_max = int(np.max(vals[tile_start[1:]] - vals[tile_start[:-1]]))
        # two-phase element-centric tile kernel (triangles)
        tile2 = _tables.build_tile_elem_tables(tables, tile_start) \
            if os.environ.get('VF_TILE2', '1') == '1' else None
        tile2_threads = 0
        if tile2 is not None:
            want = max(tile2['max_tile_elems'], d * nodes_per_tile)
            tile2_threads = int(os.environ.get('VF_TILE2_THREADS', str(-(-want // 32) * 32)))
            smem = 8 * (18 * tile2['max_tile_elems'] + max(tile_max, 8 * tile2['max_tile_verts'])
                        + d * nodes_per_tile) \
                + 4 * (tile2['max_tile_pairs'] + 2 * nodes_per_tile + 8)
            if smem > 200 * 1024 or tile2_threads > 320:
                tile2, tile2_threads = None, 0
        self.tile_info = {'nodes_per_tile': nodes_per_tile, 'ntiles': len(tile_start) - 1,
                          'tile_max_values': tile_max, 'two_phase': tile2 is not None,
                          'max_tile_elems': tile2['max_tile_elems'] if tile2 else 0,
                          'max_tile_verts': tile2['max_tile_verts'] if tile2 else 0,
                          'tile2_threads': tile2_threads}

        if s is None:
            s = np.zeros((0, 0))
        s = np.ascontiguousarray(np.atleast_2d(np.asarray(s, dtype=np.float64)))
        self.n_fluid, self.ns = (s.shape[0], s.shape[1]) if s.size else (0, 0)
        fsi_solid = np.zeros(0, np.int32) if fsi_solid is None else \
            np.ascontiguousarray(fsi_solid, dtype=np.int32)
        fsi_fluid = np.zeros(0, np.int32) if fsi_fluid is None else \
            np.ascontiguousarray(fsi_fluid, dtype=np.int32)
        if len(fsi_solid) != len(fsi_fluid):
            raise ValueError("solid/fluid FSI dof arrays must have the same length")

        def keep_last(key):
            # numpy fancy assignment x[key] = v keeps the LAST value written to a repeated
            # index; a parallel scatter needs unique targets to reproduce that
            _, first_rev = np.unique(key[::-1], return_index=True)
            return np.sort(len(key) - 1 - first_rev)
        ka = keep_last(fsi_fluid) if len(fsi_fluid) else np.zeros(0, np.int64)
        kp = keep_last(fsi_solid) if len(fsi_solid) else np.zeros(0, np.int64)
        fsia_solid, fsia_fluid = np.ascontiguousarray(fsi_solid[ka]), np.ascontiguousarray(fsi_fluid[ka])
        fsip_solid, fsip_fluid = np.ascontiguousarray(fsi_solid[kp]), np.ascontiguousarray(fsi_fluid[kp])
        self.state_size = 3 * self.N + self.n_fluid + self.n_fluid * self.ns

        # keep every host array referenced by the descriptor alive until vf_create returns
        keep = dict(tables)
        keep.update(tile_start=tile_start, s=s, fsi_solid=fsia_solid, fsi_fluid=fsia_fluid,
                    fsip_solid=fsip_solid, fsip_fluid=fsip_fluid)
        # pair info of one tile spanning the whole mesh: record-based assembly inside the
        # per-member time-loop kernel (small meshes only: the cell index is packed in 12 bits)
        if d == 2 and self.ne < 4096 and tables.get('fan_ok', False):
            whole = _tables.build_tile_elem_tables(tables, np.array([0, self.nn], dtype=np.int32))
            if whole is not None:
                keep['gpair'] = np.ascontiguousarray(whole['pair_info'])
        if tile2 is not None:
            keep.update(te_ptr=tile2['te_ptr'], te_elem=tile2['te_elem'],
                        pair_info=tile2['pair_info'], tile_desc=tile2['tile_desc'],
                        te_quad=tile2['te_quad'], tile_halo=tile2['tile_halo'])
        desc = ProblemDesc(
            d, self.nn, self.ne, tables['nfp'],
            _ptr(keep['xyz']), _ptr(keep['cells']), _ptr(keep['brptr']), _ptr(keep['bcol']),
            _ptr(keep['n2e_ptr']), _ptr(keep['n2e']), _ptr(keep['n2f_ptr']), _ptr(keep['n2f']),
            _ptr(keep['pf_cell']), _ptr(keep['pf_opp']), _ptr(keep['bc']),
            _ptr(tile_start), len(tile_start) - 1, tile_max, tile_threads,
            _ptr(keep.get('te_ptr')), _ptr(keep.get('te_elem')), _ptr(keep.get('pair_info')),
            _ptr(keep.get('tile_desc')), _ptr(keep.get('te_quad')), _ptr(keep.get('tile_halo')),
            tile2['max_tile_elems'] if tile2 else 0, tile2['max_tile_pairs'] if tile2 else 0,
            tile2['n_tile_halo'] if tile2 else 0, tile2['max_tile_verts'] if tile2 else 0,
            tile2_threads, int(bool(tables.get('fan_ok', False))),
            _ptr(keep.get('gpair')),
            self.n_fluid, self.ns, len(fsia_solid), _ptr(s), _ptr(fsia_solid), _ptr(fsia_fluid),
            len(fsip_solid), _ptr(fsip_solid), _ptr(fsip_fluid),
            int(fluid_kind), int(idx_sep), int(bool(contact)), int(bool(membrane)),
            DAMPING[damping],
            self.n_members, int(gmres_restart),
        )
        nbytes = self._lib.vf_arena_bytes(C.byref(desc))
        if nbytes == 0:
            raise _cabi.VFError(self._lib.vf_last_error().decode())
        with torch.cuda.device(self.device):
            self.arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            handle = C.c_void_p()
            check(self._lib.vf_create(C.byref(desc), self.arena.data_ptr(), nbytes,
                                      self._stream(), C.byref(handle)))
        self._h = handle
        self.nnz = int(self._lib.vf_nnz(self._h))
        # node-centric fan kernel (triangles with ordered vertex fans): the default 2D assembly
        self.fan_info = None
        if d == 2 and tables.get('fan_ok', False) and os.environ.get('VF_FAN', '1') != '0':
            fan = _tables.build_fan_tables(tables, int(os.environ.get('VF_FAN_NODES', '128')))
            if fan is not None:
                with torch.cuda.device(self.device):
                    check(self._lib.vf_set_fan_tables(
                        self._h, fan['tile_nodes'], fan['ntiles'], _ptr(fan['desc']),
                        _ptr(fan['ring']), fan['ring'].size, _ptr(fan['halo']), fan['n_halo'],
                        _ptr(fan['tcell']), fan['tcell'].size, fan['max_verts'],
                        fan['max_rows'], fan['max_cells'], fan['max_blocks'], self._stream()))
                self.fan_info = {k: fan[k] for k in ('tile_nodes', 'ntiles', 'max_verts',
                                                     'max_rows', 'max_cells', 'max_blocks',
                                                     'n_halo')}
        self._views = {}
        self._pinned = {}
        self._pinned_up = {}
        self._registered = {}

    # --- plumbing -------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def __del__(self):
        for key in list(getattr(self, '_registered', {})):
            try:
                torch.cuda.cudart().cudaHostUnregister(key)
            except Exception:
                pass
            self._registered.pop(key, None)
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            self._lib.vf_destroy(h)
            self._h = None

    def view(self, name: str, member: int = 0) -> torch.Tensor:
        """fp64 view into the arena of a named per-member array."""
        if name in ('rho', 'eta', 'emod'):
            self.props_changed(member)   # the caller may write through the view
        key = (name, member)
        if key not in self._views:
            off, cnt = C.c_size_t(), C.c_size_t()
            check(self._lib.vf_array_info(self._h, ARRAY_IDS[name], member, C.byref(off),
                                          C.byref(cnt)))
            self._views[key] = self.arena[off.value:off.value + 8 * cnt.value].view(torch.float64)
        return self._views[key]

    def props_changed(self, member: int = -1):
        """DG0 properties (rho, eta, emod) were written in the arena directly: the tile-ordered
        copy used by the fan assembly kernel is refreshed at the next assembly."""
        h = getattr(self, '_h', None)
        if h is not None:
            check(self._lib.vf_props_changed(h, int(member)))

    def member_view(self, name: str) -> torch.Tensor:
        """(n_members, count) strided fp64 view of a named array across all members."""
        if name in ('rho', 'eta', 'emod'):
            self.props_changed(-1)
        v0 = self.view(name, 0)
        if self.n_members == 1:
            return v0.unsqueeze(0)
        off0, off1, cnt = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(self._lib.vf_array_info(self._h, ARRAY_IDS[name], 0, C.byref(off0), C.byref(cnt)))
        check(self._lib.vf_array_info(self._h, ARRAY_IDS[name], 1, C.byref(off1), C.byref(cnt)))
        stride = (off1.value - off0.value) // 8
        flat = self.arena.view(torch.float64)
        return torch.as_strided(flat, (self.n_members, cnt.value), (stride, 1), off0.value // 8)

    def _page_lock(self, arr: np.ndarray) -> bool:
        """cudaHostRegister a caller-owned array once (the engine keeps it alive)."""
        key = arr.ctypes.data
        hit = self._registered.get(key)
        if hit is not None:
            return hit[1] >= arr.nbytes
        rc = torch.cuda.cudart().cudaHostRegister(key, arr.nbytes, 0)
        ok = int(rc) == 0
        if ok:
            self._registered[key] = (arr, arr.nbytes)
        return ok

    def upload(self, name: str, value, member: int = 0, persistent: bool = False):
        """Host -> device copy of a named array (scalars are broadcast).  Large arrays move at
        PCIe DMA rate: ``persistent=True`` page-locks the caller's own (long-lived, contiguous
        float64) array in place; otherwise they are staged through a cached pinned buffer."""
        n = self.view(name, member).numel()
        a = None
        if n * 8 >= (1 << 20):
            if persistent and isinstance(value, np.ndarray) and value.dtype == np.float64 \
                    and value.flags.c_contiguous and value.size == n and self._page_lock(value):
                a = value
            else:
                if name not in self._pinned_up:
                    self._pinned_up[name] = torch.empty(n, dtype=torch.float64, pin_memory=True)
                a = self._pinned_up[name].numpy()
                a[:] = np.ravel(value) if np.ndim(value) else value
        else:
            a = np.empty(n, dtype=np.float64)
            a[:] = np.ravel(value) if np.ndim(value) else value
        check(self._lib.vf_upload(self._h, ARRAY_IDS[name], member, _ptr(a), a.size,
                                  self._stream()))

    def download(self, name: str, member: int = 0, pinned: bool = False) -> np.ndarray:
        """Device -> host copy of a named array.  ``pinned=True`` returns a view of a cached
        page-locked staging buffer (valid until the next pinned download of that array): large
        results such as the CSR values then move at full PCIe rate."""
        v = self.view(name, member)
        if pinned:
            key = (name, member)
            if key not in self._pinned:
                self._pinned[key] = torch.empty(v.numel(), dtype=torch.float64, pin_memory=True)
            out = self._pinned[key].numpy()
        else:
            out = np.empty(v.numel(), dtype=np.float64)
        check(self._lib.vf_download(self._h, ARRAY_IDS[name], member, _ptr(out), out.size,
                                    self._stream()))
        return out

    def csr_pattern(self):
        rowptr = np.empty(self.N + 1, dtype=np.int32)
        colidx = np.empty(self.nnz, dtype=np.int32)
        check(self._lib.vf_csr_pattern(self._h, _ptr(rowptr), _ptr(colidx)))
        return rowptr, colidx

    @property
    def launch_count(self) -> int:
        return int(self._lib.vf_launch_count(self._h))

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()

    # --- the hot path -----------------------------------------------------------------
    def assemble(self, member: int = 0, res: bool = True, jac: bool = True, dt: float = 1.0,
                 is_static: bool = False):
        flags = (1 if res else 0) | (2 if jac else 0)
        check(self._lib.vf_assemble(self._h, member, flags, float(dt), int(is_static),
                                    self._stream()))

    def assemble_mix(self, coef4, dt: float, member: int = 0, apply_bc: bool = False):
        """J := coef4 . (K, C, M, K_p) on the Jacobian's pattern (``vf_assemble_mix``)."""
        c = np.ascontiguousarray(coef4, dtype=np.float64)
        assert c.size == 4
        check(self._lib.vf_assemble_mix(self._h, member, float(dt), _ptr(c), int(apply_bc),
                                        self._stream()))

    def pressure_control_blocks(self, member: int = 0):
        """(rows, cols, blocks): vertex pairs (a, b) of the pressure facets and the (dim,)
        blocks d F_u[a] / d p1[b] (``vf_pressure_control_blocks``)."""
        d = self.dim
        n = self.nfp * d * d
        rows = np.zeros(n, dtype=np.int32)
        cols = np.zeros(n, dtype=np.int32)
        out = torch.zeros(max(n * d, 1), dtype=torch.float64, device=self.device)
        check(self._lib.vf_pressure_control_blocks(self._h, member, out.data_ptr(), _ptr(rows),
                                                   _ptr(cols), self._stream()))
        return rows, cols, out[:n * d].cpu().numpy().reshape(n, d)

    def spmv(self, x: torch.Tensor, y: torch.Tensor, member: int = 0):
        assert x.dtype == torch.float64 and y.dtype == torch.float64
        assert x.numel() == self.N and y.numel() == self.N and x.is_cuda and y.is_cuda
        check(self._lib.vf_spmv(self._h, member, x.data_ptr(), y.data_ptr(), self._stream()))

    # --- grid-wide Krylov building blocks (large meshes / mesh partitions) --------------
    def spmv_rows(self, x: torch.Tensor, y: torch.Tensor, node0: int, node1: int, member: int = 0):
        check(self._lib.vf_spmv_rows(self._h, member, x.data_ptr(), y.data_ptr(), int(node0),
                                     int(node1), self._stream()))

    def block_jacobi_setup(self, node0: int, node1: int, member: int = 0):
        check(self._lib.vf_block_jacobi_setup(self._h, member, int(node0), int(node1),
                                              self._stream()))

    def block_jacobi_apply(self, r: torch.Tensor, z: torch.Tensor, node0: int, node1: int,
                           member: int = 0):
        check(self._lib.vf_block_jacobi_apply(self._h, member, r.data_ptr(), z.data_ptr(),
                                              int(node0), int(node1), self._stream()))

    def band_setup(self):
        """Reverse Cuthill-McKee ordering of the node graph and the band storage of the banded LU
        (``csrc/band.cu``); returns the half bandwidth in scalar DOFs."""
        import scipy.sparse as sp
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        brptr, bcol = self.tables['brptr'], self.tables['bcol']
        nn, d = self.nn, self.dim
        g = sp.csr_matrix((np.ones(len(bcol), dtype=np.int8), bcol, brptr), shape=(nn, nn))
        order = reverse_cuthill_mckee(g, symmetric_mode=True)        # new position -> old node
        pos = np.empty(nn, dtype=np.int64)
        pos[order] = np.arange(nn)
        row = np.repeat(np.arange(nn), np.diff(brptr))
        hb_nodes = int(np.max(np.abs(pos[row] - pos[bcol])))
        hb = d * hb_nodes + d - 1
        perm = (d * pos[:, None] + np.arange(d)[None, :]).ravel().astype(np.int32)
        check(self._lib.vf_band_setup(self._h, _ptr(perm), hb, self._stream()))
        self.band_half_bandwidth = hb
        return hb

    def band_factor(self, member: int = 0):
        check(self._lib.vf_band_factor(self._h, member, self._stream()))

    def band_solve(self, b: torch.Tensor, x: torch.Tensor):
        check(self._lib.vf_band_solve(self._h, b.data_ptr(), x.data_ptr(), self._stream()))

    def ilu_setup(self, node0: int = 0, node1: Optional[int] = None) -> int:
        """Colour the node graph of rows [node0, node1) and allocate the block ILU(0) storage;
        returns the number of colours."""
        node1 = self.nn if node1 is None else node1
        color, rows, cptr = _tables.color_node_graph(self.tables['brptr'], self.tables['bcol'],
                                                     node0, node1)
        check(self._lib.vf_ilu_setup(self._h, node0, node1, len(cptr) - 1, _ptr(rows), _ptr(cptr),
                                     _ptr(color), self._stream()))
        self.ilu_colors = len(cptr) - 1
        return self.ilu_colors

    def ilu_factor(self, member: int = 0):
        check(self._lib.vf_ilu_factor(self._h, member, self._stream()))

    def ilu_apply(self, r: torch.Tensor, z: torch.Tensor):
        """z = U^-1 L^-1 r on the DOFs of the factorised node range (full local vectors)."""
        check(self._lib.vf_ilu_apply(self._h, r.data_ptr(), z.data_ptr(), self._stream()))

    def multidot(self, V: torch.Tensor, nvec: int, w: torch.Tensor, n: int, out: torch.Tensor,
                 scratch: torch.Tensor):
        check(self._lib.vf_multidot(self._h, V.data_ptr(), V.stride(0), int(nvec), w.data_ptr(),
                                    int(n), out.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                    self._stream()))

    def multi_axpy(self, V: torch.Tensor, nvec: int, h: torch.Tensor, w: torch.Tensor, n: int):
        check(self._lib.vf_multi_axpy(self._h, V.data_ptr(), V.stride(0), int(nvec), h.data_ptr(),
                                      w.data_ptr(), int(n), self._stream()))

    def axpby(self, alpha: float, x: torch.Tensor, beta: float, y: torch.Tensor, n: int):
        check(self._lib.vf_axpby(self._h, float(alpha), x.data_ptr(), float(beta), y.data_ptr(),
                                 int(n), self._stream()))

    def glottal_width_series(self, u_hist: np.ndarray, member: int = 0) -> np.ndarray:
        """min fluid area of every stored displacement state (rows of ``u_hist``, host)."""
        u_hist = np.ascontiguousarray(u_hist, dtype=np.float64).reshape(-1, self.N)
        ud = torch.as_tensor(u_hist).to(self.device)
        out = torch.empty(u_hist.shape[0], dtype=torch.float64, device=self.device)
        check(self._lib.vf_glottal_width_series(self._h, member, u_hist.shape[0], ud.data_ptr(),
                                                self.N, out.data_ptr(), self._stream()))
        return out.cpu().numpy()

    def scale_rsqrt(self, x: torch.Tensor, s2: torch.Tensor, y: torch.Tensor, n: int,
                    sub: Optional[torch.Tensor] = None, s_out: Optional[torch.Tensor] = None):
        """y = x / sqrt(s2[0] - sum(sub**2)), scalars read on the device; the radicand is
        stored to ``s_out[0]`` when given."""
        check(self._lib.vf_scale_rsqrt(
            self._h, x.data_ptr(), s2.data_ptr(), sub.data_ptr() if sub is not None else None,
            int(sub.numel()) if sub is not None else 0,
            s_out.data_ptr() if s_out is not None else None, y.data_ptr(), int(n),
            self._stream()))

    def newmark_residual(self, dt: float, member: int = 0, pinned: bool = False):
        """Host copies of F_v, F_a (``transient.py:374-377``) computed on the device from the
        member's resident state.  ``pinned=True``: views of cached page-locked buffers."""
        if not hasattr(self, '_nmk_dev'):
            self._nmk_dev = torch.empty((2, self.N), dtype=torch.float64, device=self.device)
        d = self._nmk_dev
        check(self._lib.vf_newmark_residual(self._h, member, float(dt), d[0].data_ptr(),
                                            d[1].data_ptr(), self._stream()))
        if pinned:
            if not hasattr(self, '_nmk_host'):
                self._nmk_host = torch.empty((2, self.N), dtype=torch.float64, pin_memory=True)
            h = self._nmk_host
        else:
            h = torch.empty((2, self.N), dtype=torch.float64)
        h.copy_(d)
        return h[0].numpy(), h[1].numpy()

    def linear_solve(self, b: torch.Tensor, x: torch.Tensor, member: int = 0, options=None):
        info = np.zeros(3)
        opts = make_solver_opts(options)
        check(self._lib.vf_linear_solve(self._h, member, b.data_ptr(), x.data_ptr(),
                                        C.byref(opts), _ptr(info), self._stream()))
        return {'iterations': int(info[0]), 'residual': float(info[1]), 'bnorm': float(info[2])}

    def solve_state1(self, dt: float, member0: int = 0, count: int = 1, options=None,
                     is_static: bool = False):
        opts = make_solver_opts(options, is_static)
        check(self._lib.vf_solve_state1(self._h, member0, count, float(dt), C.byref(opts),
                                        self._stream()))

    def fluid_solve(self, member0: int = 0, count: int = 1):
        check(self._lib.vf_fluid_solve(self._h, member0, count, self._stream()))

    def integrate(self, dts, controls, options=None, store_states: bool = False,
                  store_info: bool = True):
        """
        ``nsteps`` coupled steps for all members from ``state0``.

        controls : (ncontrols, 2, n_fluid) array of (psub, psup)
        Returns (hist_state or None, hist_info or None) as device tensors shaped
        (n_members, nsteps+1, state_size) and (n_members, nsteps+1, 4).
        """
        dts = np.ascontiguousarray(dts, dtype=np.float64).reshape(-1)
        controls = np.ascontiguousarray(controls, dtype=np.float64).reshape(-1, 2, self.n_fluid)
        nsteps = len(dts)
        hist_state = hist_info = None
        if store_states:
            hist_state = torch.empty((self.n_members, nsteps + 1, self.state_size),
                                     dtype=torch.float64, device=self.device)
        if store_info:
            hist_info = torch.empty((self.n_members, nsteps + 1, 4), dtype=torch.float64,
                                    device=self.device)
        opts = make_solver_opts(options)
        check(self._lib.vf_integrate(
            self._h, nsteps, _ptr(dts), controls.shape[0], _ptr(controls), C.byref(opts),
            None if hist_state is None else hist_state.data_ptr(),
            None if hist_info is None else hist_info.data_ptr(), self._stream()))
        return hist_state, hist_info

    def integrate_host(self, dts, controls, ini_state: np.ndarray, emod=None, eta=None,
                       options=None, fin_state=None, info_series=None):
        """End-to-end variant with host buffers (host<->device copies inside the call)."""
        dts = np.ascontiguousarray(dts, dtype=np.float64).reshape(-1)
        controls = np.ascontiguousarray(controls, dtype=np.float64).reshape(-1, 2, self.n_fluid)
        nsteps = len(dts)
        B = self.n_members
        assert ini_state.shape == (B, self.state_size) and ini_state.flags.c_contiguous
        if fin_state is None:
            fin_state = np.empty((B, self.state_size))
        if info_series is None:
            info_series = np.empty((B, nsteps + 1, 4))
        for a in (emod, eta):
            assert a is None or (a.shape == (B, self.ne) and a.flags.c_contiguous)
        opts = make_solver_opts(options)
        check(self._lib.vf_integrate_host(
            self._h, nsteps, _ptr(dts), controls.shape[0], _ptr(controls), C.byref(opts),
            _ptr(ini_state), _ptr(emod), _ptr(eta), _ptr(fin_state), _ptr(info_series),
            self._stream()))
        return fin_state, info_series
