"""Regenerates vf-fem_b200/csrc/p2_tables.h: exact reference integrals of the P2 triangle.
    python profiles/gen_p2_tables.py"""
import os
import sympy as sy

L0, L1 = sy.symbols('L0 L1')
L2 = 1 - L0 - L1
L = [L0, L1, L2]
phi = [L[0] * (2 * L[0] - 1), L[1] * (2 * L[1] - 1), L[2] * (2 * L[2] - 1),
       4 * L[1] * L[2], 4 * L[0] * L[2], 4 * L[0] * L[1]]
a0, a1, a2 = sy.symbols('a0 a1 a2')   # the barycentric coordinates as independent variables
ph = [a0 * (2 * a0 - 1), a1 * (2 * a1 - 1), a2 * (2 * a2 - 1), 4 * a1 * a2, 4 * a0 * a2, 4 * a0 * a1]
D = [[sy.diff(ph[a], v).subs({a0: L0, a1: L1, a2: L2}) for v in (a0, a1, a2)] for a in range(6)]


def integ(f):
    """Integral over the reference triangle as a fraction of its area."""
    return 2 * sy.integrate(sy.integrate(f, (L1, 0, 1 - L0)), (L0, 0, 1))


def fmt(x):
    p, q = sy.nsimplify(x).as_numer_denom()
    return f"{int(p)}.0 / {int(q)}.0" if q != 1 else f"{int(p)}.0"


W = [[[[integ(D[a][k] * D[b][l]) for l in range(3)] for k in range(3)] for b in range(6)]
     for a in range(6)]
Mm = [[integ(phi[a] * phi[b]) for b in range(6)] for a in range(6)]
out = ["// Reference tensors of the P2 triangle (exact rationals, generated with sympy by\n"
       "// profiles/gen_p2_tables.py):\n"
       "//   kP2W[a][b][k][l] = (1/|K|) int dphi_a/dL_k dphi_b/dL_l   (grad phi_a = sum_k dphi_a/dL_k G_k)\n"
       "//   kP2M[a][b]       = (1/|K|) int phi_a phi_b\n#pragma once\n\nnamespace vf {\n",
       "__device__ __constant__ double kP2W[6][6][3][3] = {"]
for a in range(6):
    out.append(" {")
    for b in range(6):
        out.append("  {" + ", ".join("{" + ", ".join(fmt(W[a][b][k][l]) for l in range(3)) + "}"
                                     for k in range(3)) + "},")
    out.append(" },")
out.append("};\n")
out.append("__device__ __constant__ double kP2M[6][6] = {")
for a in range(6):
    out.append(" {" + ", ".join(fmt(Mm[a][b]) for b in range(6)) + "},")
out.append("};\n\n}  // namespace vf\n")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'vf-fem_b200',
                    'csrc', 'p2_tables.h')
open(path, 'w').write("\n".join(out))
