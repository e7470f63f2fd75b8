"""Grid-wide GMRES on the refined benchmark meshes: iterations / time per preconditioner.
    python profiles/grid_solve_probe.py [levels] [rtol] [maxiter]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import torch
import bench
from femvf_b200.gridsolve import GridSolver
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-10
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
model = bench.build_big_model(levels, seed=0)
eng = model.engine
model._push_all()
eng.assemble(0, res=False, jac=True, dt=model.dt)
g = torch.Generator(device='cuda'); g.manual_seed(0)
xs = torch.randn(eng.N, dtype=torch.float64, device='cuda', generator=g)
b = torch.empty_like(xs)
eng.spmv(xs, b)
for precond in ('ilu0', 'jacobi'):
    gs = GridSolver(eng, precond=precond)
    x = torch.empty_like(xs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    info = gs.linear_solve(b, x, rtol=rtol, maxiter=maxiter)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    err = (torch.linalg.vector_norm(x - xs) / torch.linalg.vector_norm(xs)).item()
    print(json.dumps({'levels': levels, 'dof': eng.N, 'precond': precond, 'rtol': rtol,
                      'iterations': info['iterations'], 'restarts': info['restarts'],
                      'rel_resid': info['residual'] / info['bnorm'], 'x_err': err,
                      'seconds': round(dt, 3), 'ms_per_iteration': round(1e3 * dt / max(info['iterations'], 1), 3)}), flush=True)
    del gs
