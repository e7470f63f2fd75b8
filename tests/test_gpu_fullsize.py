"""GPU property tests at benchmark size (no oracle: size-independent identities).

For the 2D Newmark Kelvin-Voigt residual F_u is affine in u1, so
F(u) - F(u') = J (u - u') must hold to round-off with J from the assembly kernel and the
product from the SpMV kernel; Dirichlet rows are identity rows; a rigid translation at
constant velocity produces no interior force."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def big():
    import bench
    return bench.build_big_model(6, seed=3)  # 1.0e6 triangles, 1.0e6 DOF


def test_affine_identity_assembly_vs_spmv(big):
    import torch
    model = big
    eng = model.engine
    N = eng.N
    rng = np.random.default_rng(0)
    model._push_all()
    eng.assemble(0, res=True, jac=True, dt=model.dt)
    F1 = eng.view('F').clone()
    u1 = eng.view('u1').clone()
    du = torch.as_tensor(rng.uniform(-1e-3, 1e-3, N), device='cuda')
    fixed = torch.as_tensor(model.residual.fixed_dofs(), device='cuda')
    du[fixed] = 0.0
    eng.view('u1').copy_(u1 + du)
    eng.assemble(0, res=True, jac=False, dt=model.dt)
    F2 = eng.view('F').clone()
    y = torch.empty_like(du)
    eng.spmv(du, y)
    err = torch.max(torch.abs((F2 - F1) - y)).item()
    scale = torch.max(torch.abs(y)).item()
    assert err <= 1e-9 * scale, (err, scale)
    # Dirichlet rows: J e_k = e_k on fixed rows, residual zero there
    assert torch.all(F1[fixed] == 0)
    x = torch.zeros(N, dtype=torch.float64, device='cuda')
    x[fixed] = 1.0
    eng.spmv(x, y)
    assert torch.all(y[fixed] == 1.0)


def test_bitwise_reproducible_and_rigid_motion(big):
    import torch
    model = big
    eng = model.engine
    N, nn = eng.N, eng.nn
    eng.assemble(0, res=True, jac=True, dt=model.dt)
    J1 = eng.view('J').clone(); F1 = eng.view('F').clone()
    eng.assemble(0, res=True, jac=True, dt=model.dt)
    assert torch.equal(J1, eng.view('J')) and torch.equal(F1, eng.view('F'))
    # rigid translation at constant velocity, zero pressure: no force on free nodes
    dt = model.dt
    c = torch.tensor([0.3, -0.2], dtype=torch.float64, device='cuda')
    u0 = (0.01 * c).repeat(nn); v0 = c.repeat(nn)
    eng.view('u0').copy_(u0); eng.view('v0').copy_(v0); eng.view('a0').zero_()
    eng.view('u1').copy_(u0 + dt * v0); eng.view('p1').zero_()
    eng.assemble(0, res=True, jac=False, dt=dt)
    F = eng.view('F')
    eng.assemble(0, res=False, jac=True, dt=dt)
    diag_scale = torch.max(torch.abs(eng.view('J'))).item() * 0.01
    assert torch.max(torch.abs(F)).item() <= 1e-9 * diag_scale
