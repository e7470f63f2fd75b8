"""Solver defaults: mirror of ``/root/reference/src/femvf/solverconst.py:1-14``."""

DEFAULT_NEWTON_SOLVER_PRM = {
    'linear_solver': 'petsc',  # accepted for compatibility; the device GMRES is always used
    'absolute_tolerance': 1e-8,
    'relative_tolerance': 1e-10,
    'maximum_iterations': 50,
}

FIXEDPOINT_SOLVER_PRM = {'absolute_tolerance': 1e-8, 'relative_tolerance': 1e-11}

# Controls of the block-Jacobi GMRES that stands in for the PETSc LU
# (models/transient.py:487).  Not present in the reference.
DEFAULT_LINEAR_SOLVER_PRM = {
    'gmres_relative_tolerance': 1e-13,
    'gmres_absolute_tolerance': 0.0,
    'gmres_maximum_iterations': 4000,
    # Neumann-series degree of the polynomial acceleration of block-Jacobi (0 = plain)
    'polynomial_degree': 3,
}
