"""GPU parity of the kernel variants selected by environment switches, both settings of each:

* ``VF_NODE_WARP``: thread-per-node assembly of tetrahedra with the block rows in global memory
  (``asm_node_global_kernel``) or in a warp's packed shared-memory slice written by one coalesced
  copy (``asm_node_warp_kernel``), ``csrc/assembly.cu``;
* ``VF_TET_TABLES``: generic structure-of-arrays gathers + column-list scans, or the per-cell /
  per-node records and precomputed CSR slots of ``csrc/tet_tables.h``;
* ``VF_P2_WARP``, ``VF_P2_PACK``: first / second version of the P2 triangle kernel
  (``csrc/p2.cu``), the second with and without the packed nodal-state pre-pass.

Each setting is compared with the oracle (<= 1e-12 relative, CSR pattern bit-exact) and the two
settings with each other (same summation order per entry: equal to rounding of the last bits).
"""

import numpy as np
import pytest

from helpers import mesh_tuples
from test_gpu_assembly import _assemble_and_compare
from test_gpu_p2 import test_p2_assembly_parity as _p2_parity, _fixed_edges

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-14):
    return np.max(np.abs(a - b)) <= rtol * np.max(np.abs(a))


def _tet_mesh(name):
    from femvf_b200 import meshgen
    if name == 'cube332':
        return mesh_tuples()['cube332']()
    # 1036 nodes = 32 full warps + 12 nodes, 4410 tetrahedra, pressure / fixed facets tagged
    return meshgen.renumber_for_locality(
        meshgen.extrude_to_tets(meshgen.m5_cb_mesh(0.05), 1.5, 6))


@pytest.mark.parametrize('mesh_name', ['cube332', 'm5_extruded'])
@pytest.mark.parametrize('variant', ['kv', 'epithelium_contact'])
def test_tet_assembly_both_node_kernels(monkeypatch, mesh_name, variant):
    import torch
    assert torch.cuda.is_available()
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    membrane = contact = variant == 'epithelium_contact'
    Residual = slr.KelvinVoigtWEpithelium if membrane else slr.KelvinVoigt
    Model = transient.NodalContactModel if contact else transient.FenicsModel
    out = {}
    for setting in ('00', '10', '01', '11'):
        monkeypatch.setenv('VF_NODE_WARP', setting[0])
        monkeypatch.setenv('VF_TET_TABLES', setting[1])
        model = Model(Residual(*_tet_mesh(mesh_name)))
        assert model.residual.mesh().topology().dim() == 3
        _assemble_and_compare(model, np.random.default_rng(5), contact=contact, membrane=membrane)
        J = model.assem_dres_dstate1().sub['u', 'state/u1'].data.copy()
        # residual and Jacobian in ONE launch (the <JAC, RES> instantiation), run twice
        e = model.engine
        e.assemble(0, res=True, jac=True, dt=model.dt)
        F_both, J_both = e.download('F').copy(), e.download('J').copy()
        e.assemble(0, res=True, jac=True, dt=model.dt)
        assert np.array_equal(e.download('J'), J_both) and np.array_equal(e.download('F'), F_both)
        # other template instantiations of the same element code: equal up to fma contraction
        assert _close(J_both, J)
        assert _close(F_both, np.asarray(model.assem_res()['u']))
        out[setting] = (F_both, J_both)
    for setting in ('10', '01', '11'):
        assert _close(out['00'][0], out[setting][0])
        assert _close(out['00'][1], out[setting][1])


@pytest.mark.parametrize('setting', ['0', '1', '1-nopack'])
@pytest.mark.parametrize('mesh_name,levels', [('square5', 0), ('m5', 0), ('m5r', 2)])
def test_p2_both_kernels(monkeypatch, setting, mesh_name, levels):
    monkeypatch.setenv('VF_P2_WARP', setting[0])
    monkeypatch.setenv('VF_P2_PACK', '0' if setting.endswith('nopack') else '1')
    _p2_parity(mesh_name, levels, True)


def test_p2_kernels_agree(monkeypatch):
    """Version 1 and version 2 on the same inputs: same terms in the same order."""
    import torch
    from femvf_b200 import meshgen
    from femvf_b200.p2 import P2Assembler
    from femvf_b200.residuals import solid as slr
    from helpers import oracle_problem
    res = slr.KelvinVoigt(*meshgen.m5_cb_refined(0.05, 2))
    p1prob = oracle_problem(res)
    mesh = res.mesh()
    fe = _fixed_edges(p1prob, mesh.coordinates(), mesh.cells())
    asm = P2Assembler(mesh.coordinates(), mesh.cells(), p1prob.pfacets, p1prob.pfacet_cells, fe,
                      interleave=True)
    indptr, _ = asm.csr_pattern()
    N = len(indptr) - 1
    nn, ne = N // 2, len(mesh.cells())
    rng = np.random.default_rng(11)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    args = [t(rng.uniform(-1e-2, 1e-2, N)), t(rng.uniform(-1e-2, 1e-2, N)),
            t(rng.uniform(-1, 1, N)), t(rng.uniform(-1e3, 1e3, N)), t(rng.uniform(0, 8e3, nn)),
            t(rng.uniform(2.5e4, 1e5, ne)), t(rng.uniform(1, 5, ne)), t(rng.uniform(0.9, 1.1, ne)),
            0.45, 1e-4]
    out = {}
    for setting in ('00', '10', '11'):
        monkeypatch.setenv('VF_P2_WARP', setting[0])
        monkeypatch.setenv('VF_P2_PACK', setting[1])
        F, J = asm.assemble(*args)
        out[setting] = (F.cpu().numpy().copy(), J.cpu().numpy().copy())
        # residual-only and Jacobian-only launches of the same kernel
        F1 = asm.assemble(*args, res=True, jac=False)[0].cpu().numpy().copy()
        J1 = asm.assemble(*args, res=False, jac=True)[1].cpu().numpy().copy()
        assert _close(out[setting][0], F1) and _close(out[setting][1], J1)
    for setting in ('10', '11'):
        assert _close(out['00'][0], out[setting][0]) and _close(out['00'][1], out[setting][1])
