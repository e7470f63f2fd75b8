"""
Fluid <-> solid DOF correspondence: mirror of ``/root/reference/src/femvf/models/fsi.py``.

``FSIMap`` keeps the reference's gather/scatter semantics (``fsi.py:66-70``) on host arrays;
inside the device time loop the same two index arrays drive the gather/scatter kernels
(``csrc/member_solver.cuh``, ``blk_fluid`` and the ``p1`` scatter of ``member_kernel``).
The constant coupling Jacobians (``fsi.py:72-88, 250-263``) are scipy CSR matrices instead of
PETSc ``Mat`` objects.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class FSIMap:
    """1-to-1 correspondence between DOFs of vectors on the fluid and solid domains."""

    def __init__(self, ndof_fluid: int, ndof_solid: int, fluid_dofs, solid_dofs, comm=None):
        self.N_FLUID = ndof_fluid
        self.N_SOLID = ndof_solid
        self.dofs_fluid = np.asarray(fluid_dofs)
        self.dofs_solid = np.asarray(solid_dofs)
        self.fluid_to_solid_idx = {f: s for f, s in zip(self.dofs_fluid, self.dofs_solid)}
        self.solid_to_fluid_idx = {s: f for f, s in zip(self.dofs_fluid, self.dofs_solid)}
        self.dsolid_dfluid = self.assem_dsolid_dfluid(comm)
        self.dfluid_dsolid = self.assem_dfluid_dsolid(comm)

    def map_fluid_to_solid(self, fluid_vec, solid_vec):
        solid_vec[self.dofs_solid] = fluid_vec[self.dofs_fluid]

    def map_solid_to_fluid(self, solid_vec, fluid_vec):
        fluid_vec[self.dofs_fluid] = solid_vec[self.dofs_solid]

    def assem_dsolid_dfluid(self, comm=None):
        jj = np.fromiter(self.fluid_to_solid_idx.keys(), dtype=np.int64)
        ii = np.fromiter(self.fluid_to_solid_idx.values(), dtype=np.int64)
        return sp.csr_matrix((np.ones(len(ii)), (ii, jj)), shape=(self.N_SOLID, self.N_FLUID))

    def assem_dfluid_dsolid(self, comm=None):
        jj = np.fromiter(self.solid_to_fluid_idx.keys(), dtype=np.int64)
        ii = np.fromiter(self.solid_to_fluid_idx.values(), dtype=np.int64)
        return sp.csr_matrix((np.ones(len(ii)), (ii, jj)), shape=(self.N_FLUID, self.N_SOLID))


def make_dslarea_dslu(n_area: int, n_dis: int, ndim: int = 2):
    """Sensitivity of the channel area to the displacement vector: -2 on the y component
    (``fsi.py:250-263``)."""
    ii = np.arange(n_area)
    return sp.csr_matrix((np.full(n_area, -2.0), (ii, ndim * ii + 1)), shape=(n_area, n_dis))


def make_fsimap(solid, fluid, solid_fsi_dofs, fluid_fsi_dofs) -> FSIMap:
    """``fsi.py:165-184``."""
    n_solid = solid.residual.form['control/p1'].function_space().dim()
    return FSIMap(fluid.state0['p'].size, n_solid, fluid_fsi_dofs, solid_fsi_dofs)


def make_coupling_stuff(solid, fluid, solid_fsi_dofs, fluid_fsi_dofs):
    """``fsi.py:106-162``: FSI map, solid area work vector and the constant coupling
    Jacobians d(fluid area)/d(solid u) and d(solid p)/d(fluid p)."""
    fsimap = make_fsimap(solid, fluid, solid_fsi_dofs, fluid_fsi_dofs)
    solid_area = np.zeros(fsimap.N_SOLID)
    ndim = solid.residual.mesh().topology().dim()
    n_u = solid.state0['u'].size
    dslarea_dslu = make_dslarea_dslu(n_u // ndim, n_u, ndim)
    dflarea_dslu = fsimap.dfluid_dsolid @ dslarea_dslu
    dslp_dflp = fsimap.dsolid_dfluid
    return fsimap, solid_area, dflarea_dslu, dslp_dflp, None
