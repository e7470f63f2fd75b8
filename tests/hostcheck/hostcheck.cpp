// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
// Runs vf::assemble_node (the arithmetic the CUDA kernels execute) in a plain CPU loop so
// that the element math can be checked against the oracle without a GPU.
#include <cmath>
using std::sqrt; using std::fabs; using std::isinf;
#include "../../vf-fem_b200/csrc/node_assembly.cuh"

template <int D>
static void run(const vf::MeshView& m, const vf::PropView& p, const vf::StateView& s,
                double* J, double* F) {
  for (int i = 0; i < m.nn; ++i) {
    double res[D];
    vf::assemble_node<D, true, true>(i, m, p, s, J + D * D * m.brptr[i], res);
    for (int c = 0; c < D; ++c) F[D * i + c] = res[c];
  }
}

extern "C" int hostcheck_assemble(
    int dim, int nn, int ne, int nfp, const double* xyz, const int* cells, const int* brptr,
    const int* bcol, const int* n2e_ptr, const int* n2e, const int* n2f_ptr, const int* n2f,
    const int* pf_cell, const int* pf_opp, const unsigned char* bc, const double* rho,
    const double* eta, const double* emod, const double* scal, const double* emod_m,
    const double* nu_m, const double* th_m, int contact, int membrane, const double* u1,
    const double* u0, const double* v0, const double* a0, const double* p1, double dt,
    double* J, double* F) {
  vf::MeshView m{dim, nn, ne, nfp, xyz, cells, brptr, bcol, n2e_ptr, n2e,
                 n2f_ptr, n2f, pf_cell, pf_opp, bc};
  vf::PropView p{rho, eta, emod, scal, emod_m, nu_m, th_m, contact, membrane};
  vf::StateView s{u1, u0, v0, a0, p1, dt, 0};
  if (dim == 2) run<2>(m, p, s, J, F);
  else if (dim == 3) run<3>(m, p, s, J, F);
  else return 1;
  return 0;
}
