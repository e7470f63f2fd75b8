"""CPU tests of the host-side mirror of the reference interface (no device calls)."""

import numpy as np
import pytest

from helpers import mesh_tuples
from femvf_b200 import blockvec as bv, forward, load, mesh as M, meshgen, statefile as sf, tables
from femvf_b200.models import transient
from femvf_b200.residuals import solid as slr, fluid as flr
from oracle import model as om


def test_blockvector_contract():
    a = bv.BlockVector([np.arange(3.0), np.zeros(2), np.ones(1)], labels=[('u', 'v', 'q')])
    assert a.size == 3 and a.bshape == ((3, 2, 1),) and a.keys() == ['u', 'v', 'q']
    a['v'][:] = 5
    assert np.all(a.sub['v'] == 5) and np.all(a[1] == 5)
    b = a[:2]
    b['u'][0] = -1.0
    assert a['u'][0] == -1.0  # slices are views
    c = a.copy(); c[:] = 0
    assert a['v'][0] == 5 and c.norm() == 0
    c[:] = a
    assert np.array_equal(c.to_mono_ndarray(), a.to_mono_ndarray())
    c['q'] = 7
    assert c['q'][0] == 7
    d = bv.concatenate([a[:1], a[1:]])
    assert d.labels == a.labels
    sl, fl = bv.chunk(a, (2, 1))
    assert sl.keys() == ['u', 'v'] and fl.keys() == ['q']
    assert np.allclose((a - a * 2 + a).to_mono_ndarray(), 0)
    assert a[['q', 'u']].keys() == ['q', 'u']
    with pytest.raises(KeyError):
        a['nope']


def test_property_label_order_matches_reference():
    mt = mesh_tuples()['square5']()
    kv = transient.FenicsModel(slr.KelvinVoigt(*mt))
    assert kv.prop.keys() == ['rho', 'eta', 'emod', 'nu', 'ycontact', 'ncontact', 'kcontact']
    epi = transient.FenicsModel(slr.KelvinVoigtWEpithelium(*mesh_tuples()['square5']()))
    assert epi.prop.keys() == ['rho', 'emod_membrane', 'nu_membrane', 'th_membrane', 'emod', 'nu',
                               'eta', 'ycontact', 'ncontact', 'kcontact']
    assert kv.prop['nu'][0] == 0.45 and np.isinf(kv.prop['ycontact'][0])
    assert np.array_equal(kv.prop['ncontact'], [0.0, 1.0]) and kv.prop['kcontact'][0] == 1.0
    assert kv.state0.keys() == ['u', 'v', 'a'] and kv.control.keys() == ['p']


def build_fsi(name='square5', Fluid=flr.BernoulliAreaRatioSep, zs=None):
    mt = mesh_tuples()[name]()
    d = mt[0].topology().dim()
    return load.load_fsi_model(
        mt, slr.KelvinVoigt, Fluid,
        {'dirichlet_bcs': {'state/u1': [(np.zeros(d), 'facet', 'fixed')]}}, {}, zs=zs)


def test_load_fsi_model_layout():
    model = build_fsi('m5')
    assert model.state0.keys() == ['u', 'v', 'a', 'q', 'p']
    assert model.control.keys() == ['psub', 'psup']
    assert model.prop.keys()[-4:] == ['rho_air', 'r_sep', 'area_lb', 'ymid']
    ns = model.fluid.state0['p'].size
    s = model.fluid.residual.mesh()
    assert s.shape == (ns,) and s[0] == 0 and np.all(np.diff(s) > 0)
    # the interface follows the loaded surface from the origin towards +x
    x = model.solid.residual.mesh().coordinates()[model.fsimap.dofs_solid]
    assert np.allclose(x[0], [0, 0]) and np.isclose(x[-1, 0], 0.7895)
    assert np.array_equal(model.fsimap.dofs_fluid, np.arange(ns))
    # nearest-neighbour ordering agrees with the oracle's independent implementation
    res = model.solid.residual
    fids, _, _ = res.pressure_facets()
    s_o, verts_o = om.interface_from_edges(res.mesh().coordinates(), res.mesh().facets[fids])
    assert np.array_equal(verts_o, model.fsimap.dofs_solid) and np.allclose(s_o, s)


def test_explicit_coupling_host_mirrors():
    """_set_ini_fluid_state / _set_fin_solid_state semantics (transient.py:833-858)."""
    model = build_fsi('m5')
    prop = model.prop.copy(); prop['ymid'][:] = 0.8
    model.set_prop(prop)
    st = model.state0.copy(); st[:] = 0
    st['p'][:] = np.arange(st['p'].size) + 1.0
    model.set_ini_state(st)
    p_solid = model.solid.control['p']
    assert np.array_equal(p_solid[model.fsimap.dofs_solid], st['p'])
    mask = np.ones(p_solid.size, bool); mask[model.fsimap.dofs_solid] = False
    assert not p_solid[mask].any()
    st['u'][1::2] = 0.01
    model.set_fin_state(st)
    y = model.solid.residual.mesh().coordinates()[model.fsimap.dofs_solid, 1]
    assert np.allclose(model.fluid.control['area'], 2 * (0.8 - (y + 0.01)))


def test_fixture_interface_includes_interior_vertices_quirk_q8():
    model = build_fsi('square5')
    # 'pressure' = facet value 0 = every unmarked facet, interior ones included (App. C, Q8)
    assert model.fluid.state0['p'].size == 36


def test_3d_interface_planes():
    model = build_fsi('cube332', zs=np.linspace(0, 1, 3))
    s = model.fluid.residual.mesh()
    assert s.ndim == 2 and s.shape[0] == 3
    assert model.fluid.state0['q'].size == 3
    assert model.control['psub'].size == 3


def test_3d_fsi_map_gives_every_plane_its_own_channel():
    """reference load.py:204-214: solid_dofs is interface_vertices.flat, fluid_dofs = arange of
    the same length, so z-plane k maps to fluid DOFs k*ns .. (k+1)*ns - 1."""
    zs = np.linspace(0, 1, 3)
    model = build_fsi('cube332', zs=zs)
    s = model.fluid.residual.mesh()
    nz, ns = s.shape
    fd, sd = model.fsimap.dofs_fluid, model.fsimap.dofs_solid
    assert np.array_equal(np.sort(fd), np.arange(nz * ns))
    coords = model.solid.residual.mesh().coordinates()
    # every plane's solid vertices lie on that plane
    for k in range(nz):
        assert np.allclose(coords[sd[fd // ns == k], 2], zs[k])
    # a displacement that differs per plane reaches each channel's area separately
    area_solid = 2.0 * (1.05 - (coords[:, 1] + 0.01 * coords[:, 2]))
    area = np.ones(nz * ns)
    model.fsimap.map_solid_to_fluid(area_solid, area)
    area = area.reshape(nz, ns)
    for k in range(nz):
        assert np.allclose(area[k], 2.0 * (1.05 - (coords[sd[k * ns:(k + 1) * ns], 1]
                                                   + 0.01 * zs[k])))
    assert not np.allclose(area[0], area[1])


def test_integrate_argument_validation():
    model = build_fsi('square5')
    st = model.state0.copy(); ctl = model.control.copy(); prop = model.prop.copy()
    with pytest.raises(ValueError):
        forward.integrate(model, None, st, [ctl], prop, [], write=False)
    with pytest.raises(ValueError):
        forward.integrate(model, None, st, [ctl], prop, [0.0, 0.0], write=False)
    with pytest.raises(ValueError):
        load.load_fsi_model(mesh_tuples()['square5'](), slr.KelvinVoigt,
                            flr.BernoulliAreaRatioSep, {}, {}, coupling='monolithic')
    with pytest.raises(ValueError):
        load.load_fenics_model('mesh.xml', slr.KelvinVoigt)


def test_statefile_layout_and_roundtrip(tmp_path):
    model = build_fsi('square5')
    path = str(tmp_path / 'state.h5')
    st = model.state0.copy(); st[:] = 0
    ctl = model.control.copy(); ctl['psub'][:] = 8e3; ctl['psup'][:] = 0
    with sf.StateFile(model, path, mode='w') as f:
        f.init_layout()
        for n in range(3):
            st['u'][:] = n
            forward.append_step_result(f, st, ctl, 0.1 * n, {'num_iter': n, 'abs_err': 0.0})
        f.append_prop(model.prop)
        assert f.size == 3
        for key in ('time', 'meas_indices', 'mesh/solid/coordinates', 'mesh/solid/connectivity',
                    'mesh/solid/dim', 'dofmap/CG1', 'state/u', 'state/q', 'control/psub',
                    'properties/emod', 'solver_info/num_iter', 'solver_info/rel_err'):
            assert key in f.file, key
        assert f.file['state/u'].shape == (3, st['u'].size)
        assert np.isnan(f.get_solver_info(2)['rel_err'])  # missing keys are written as NaN
        assert f.file['dofmap/CG1'].shape == (model.solid.residual.mesh().num_cells(), 6)
    with sf.StateFile(model, path, mode='r') as f:
        assert f.size == 3
        assert np.all(f.get_state(2)['u'] == 2) and f.get_time(1) == 0.1
        assert f.get_control(10)['psub'][0] == 8e3
        assert np.array_equal(f.get_prop()['emod'], model.prop['emod'])


def test_tables_and_tiles():
    for name in ('square5', 'cube332', 'm5'):
        mt = mesh_tuples()[name]()
        res = slr.KelvinVoigt(*mt)
        mesh = res.mesh(); d = mesh.topology().dim()
        fids, pfc, pfo = res.pressure_facets()
        T = tables.build_tables(mesh.coordinates(), mesh.cells(), pfc, pfo, res.fixed_dofs())
        # every cell appears once per vertex in the node->cell table, with the right local index
        e, a = T['n2e'] >> 2, T['n2e'] & 3
        node = np.repeat(np.arange(T['nn']), np.diff(T['n2e_ptr']))
        assert np.array_equal(mesh.cells()[e, a], node) and len(e) == (d + 1) * T['ne']
        # facet table: opposite vertex is not on the facet; outward normal convention
        f, fa = T['n2f'] >> 2, T['n2f'] & 3
        nodef = np.repeat(np.arange(T['nn']), np.diff(T['n2f_ptr']))
        assert np.array_equal(mesh.cells()[T['pf_cell'][f], fa], nodef)
        for k, fid in enumerate(fids):
            opp = mesh.cells()[pfc[k], pfo[k]]
            assert opp not in mesh.facets[fid]
        ts = tables.tile_partition(T['brptr'], d, 16, 4096)
        assert ts[0] == 0 and ts[-1] == T['nn'] and np.all(np.diff(ts) > 0)
        assert np.all(np.diff(ts) <= 16)
        vals = d * d * T['brptr'].astype(np.int64)
        assert np.all(vals[ts[1:]] - vals[ts[:-1]] <= 4096)


def test_mesh_generators():
    mesh, mfs, sd = meshgen.m5_cb_mesh(0.05)
    assert abs(mesh.signed_volumes().sum() - 0.2868) < 2e-3
    assert np.all(mesh.signed_volumes() > 0)
    tags = mfs[1].array(); ext = mesh.exterior_facets
    assert set(np.unique(tags[ext])) == {sd[1]['pressure'], sd[1]['fixed']}
    mt2 = meshgen.refine_red((mesh, mfs, sd))
    assert mt2[0].num_cells() == 4 * mesh.num_cells()
    assert (mt2[1][1].array() == sd[1]['pressure']).sum() == 2 * (tags == sd[1]['pressure']).sum()
    assert abs(mt2[0].signed_volumes().sum() - mesh.signed_volumes().sum()) < 1e-12
    mt3 = meshgen.renumber_for_locality(mt2)
    assert (mt3[1][1].array() == sd[1]['fixed']).sum() == (mt2[1][1].array() == sd[1]['fixed']).sum()
    m3 = meshgen.extrude_to_tets((mesh, mfs, sd), 1.5, 3)
    assert np.isclose(m3[0].signed_volumes().sum(), 1.5 * mesh.signed_volumes().sum())
    assert len(m3[0].exterior_facets) == 2 * mesh.num_cells() + 2 * 3 * len(ext)


def test_grid_gmres_host_logic_with_stub_engine():
    """GridGMRES control flow (blocked host read-back, Givens, restarts) against a dense solve.
    The engine is replaced by a torch-CPU stub of the C-ABI vector kernels, so only the host
    logic of femvf_b200/distributed.py is exercised here; the kernels are covered by -m gpu."""
    import scipy.sparse as sp
    import torch
    from femvf_b200.distributed import GridGMRES

    class StubEngine:
        def __init__(self, A):
            self.A = torch.as_tensor(A.toarray())
            self.N, self.dim, self.device = A.shape[0], 2, 'cpu'
            self.dinv = torch.as_tensor(1.0 / A.diagonal())

        def block_jacobi_setup(self, a, b): pass
        def block_jacobi_apply(self, x, y, a, b): y.copy_(self.dinv * x)
        def spmv_rows(self, x, y, a, b): y.copy_(self.A @ x)
        def multidot(self, V, nvec, w, n, out, scratch): out.copy_(V[:nvec] @ w)
        def multi_axpy(self, V, nvec, h, w, n): w.sub_(V[:nvec].T @ h)
        def axpby(self, alpha, x, beta, y, n): y.copy_(alpha * x + beta * y)
        def scale_rsqrt(self, x, s2, y, n, sub=None, s_out=None):
            s = s2[0] - (torch.dot(sub, sub) if sub is not None else 0.0)
            if s_out is not None:
                s_out[0] = s
            y.copy_(x / torch.sqrt(s))

    rng = np.random.default_rng(0)
    n = 60
    A = sp.random(n, n, 0.2, random_state=1) + sp.identity(n) * 5
    A = (A + A.T).tocsr()
    b = rng.standard_normal(n)
    for restart, every in ((10, 8), (10, 3), (7, 1), (64, 8)):
        g = GridGMRES(StubEngine(A), n // 2, None, restart=restart)
        x = torch.zeros(n, dtype=torch.float64)
        info = g.solve(torch.as_tensor(b), x, rtol=1e-12, check_every=every)
        assert np.linalg.norm(A @ x.numpy() - b) <= 1e-10 * np.linalg.norm(b), info
        assert info['iterations'] < 200
    # iteration cap honoured exactly (the bench runs a fixed count)
    g = GridGMRES(StubEngine(A), n // 2, None, restart=10)
    info = g.solve(torch.as_tensor(b), torch.zeros(n, dtype=torch.float64), rtol=0.0, maxiter=25)
    assert info['iterations'] == 25 and g.spmv_count == 27


def test_node_colouring_and_band_ordering_tables():
    """Host tables of the whole-GPU solvers: the multicolour ordering of csrc/ilu.cu (no two
    adjacent nodes share a colour, classes cover the range) and the reverse Cuthill-McKee band
    ordering handed to csrc/band.cu (a permutation whose bandwidth bounds every non-zero)."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    from femvf_b200 import meshgen, tables
    mesh = meshgen.m5_cb_refined(0.05, 2)[0]
    T = tables.build_tables(mesh.coordinates(), mesh.cells(), [], [], [])
    nn, brptr, bcol = T['nn'], T['brptr'].astype(np.int64), T['bcol']
    row = np.repeat(np.arange(nn), np.diff(brptr))
    for node0, node1 in ((0, nn), (100, nn - 57)):
        color, rows, cptr = tables.color_node_graph(T['brptr'], T['bcol'], node0, node1)
        inside = (row >= node0) & (row < node1) & (bcol >= node0) & (bcol < node1) & (row != bcol)
        assert not np.any(color[row[inside]] == color[bcol[inside]])
        assert np.all(color[:node0] == -1) and np.all(color[node1:] == -1)
        assert np.array_equal(np.sort(rows), np.arange(node0, node1))
        assert cptr[0] == 0 and cptr[-1] == node1 - node0 and len(cptr) - 1 <= 12
        for c in range(len(cptr) - 1):
            assert np.all(color[rows[cptr[c]:cptr[c + 1]]] == c)
    # band ordering as built by Engine.band_setup
    g = sp.csr_matrix((np.ones(len(bcol), dtype=np.int8), bcol, T['brptr']), shape=(nn, nn))
    order = reverse_cuthill_mckee(g, symmetric_mode=True)
    pos = np.empty(nn, dtype=np.int64)
    pos[order] = np.arange(nn)
    hb = int(np.max(np.abs(pos[row] - pos[bcol])))
    assert np.array_equal(np.sort(pos), np.arange(nn))
    assert hb < nn // 10          # RCM finds the thin direction of the M5_CB outline


def test_kelvin_voigt_w_shape_moves_the_mesh():
    """``KelvinVoigtWShape`` (reference ``residuals/solid.py:192-215``): 'umesh' is the last
    property and ``set_prop`` displaces the mesh coordinates by it
    (``models/transient.py:347-360``); the device tables follow and the engine is rebuilt."""
    mt = mesh_tuples()['m5']()
    model = transient.FenicsModel(slr.KelvinVoigtWShape(*mt))
    assert model.prop.keys() == ['rho', 'emod', 'nu', 'eta', 'ycontact', 'ncontact', 'kcontact',
                                 'umesh']
    mesh = model.residual.mesh()
    ref = mesh.coordinates().copy()
    nn = ref.shape[0]
    assert model.prop['umesh'].size == 2 * nn and not np.any(model.prop['umesh'])
    tables0 = model.assembly_tables
    model._engine = sentinel = object()       # stands for a live engine (no device here)
    prop = model.prop.copy()
    prop['emod'][:] = 5e4
    model.set_prop(prop)                      # umesh still zero: geometry and engine are kept
    assert model._engine is sentinel and model.assembly_tables is tables0
    rng = np.random.default_rng(0)
    du = 1e-3 * rng.standard_normal((nn, 2))
    prop['umesh'][:] = du.ravel()
    model.set_prop(prop)
    assert np.array_equal(mesh.coordinates(), ref + du)
    assert np.array_equal(model.XREF, (ref + du).ravel())
    assert model._engine is None              # next device call builds an engine on the new mesh
    assert np.array_equal(model.assembly_tables['xyz'], (ref + du).T)
    assert all(model._dirty.values())
    # the pattern and the pressure / Dirichlet tables are topological: unchanged
    for key in ('brptr', 'bcol', 'bc', 'pf_cell'):
        assert np.array_equal(model.assembly_tables[key], tables0[key])
    prop['umesh'][:] = 0.0
    model.set_prop(prop)
    assert np.array_equal(mesh.coordinates(), ref)


def test_shape_change_drops_the_shared_fsi_engine():
    mt = mesh_tuples()['m5']()
    model = load.load_fsi_model(
        mt, slr.KelvinVoigtWShape, flr.BernoulliAreaRatioSep,
        {'dirichlet_bcs': {'state/u1': [(np.zeros(2), 'facet', 'fixed')]}}, {})
    assert model.prop.keys()[:8] == ['rho', 'emod', 'nu', 'eta', 'ycontact', 'ncontact',
                                     'kcontact', 'umesh']
    model._engine = model.solid._engine = model.fluid._engine = object()
    prop = model.prop.copy()
    prop['umesh'][:] = 1e-3
    model.set_prop(prop)
    assert model._engine is None and model.solid._engine is None and model.fluid._engine is None
    # an engine shared by an ensemble cannot follow a shape change
    model.solid._engine_attached = True
    prop['umesh'][:] = 2e-3
    with pytest.raises(NotImplementedError):
        model.set_prop(prop)


def test_large_blockvector_assignment_is_chunked_and_exact():
    """Assignments of >= 2 M elements go through the threaded, chunked copy (blockvec._parallel_copy):
    same values, for one or several blocks, odd sizes and a 2-D block."""
    rng = np.random.default_rng(3)
    a = bv.BlockVector([rng.random(1500001), rng.random(7), rng.random((300000, 2))],
                       labels=[('x', 'y', 'z')])
    b = a.copy()
    for v in b.vecs:
        v[...] = -1.0
    b[:] = a
    assert all(np.array_equal(x, y) for x, y in zip(a.vecs, b.vecs))
    c = bv.BlockVector([rng.random(2200003)], labels=[('x',)])
    d = c.copy(); d['x'][:] = 0.0
    d[:] = c
    assert np.array_equal(c['x'], d['x'])
    pool, workers = bv._copy_pool()
    assert 2 <= workers <= 8
