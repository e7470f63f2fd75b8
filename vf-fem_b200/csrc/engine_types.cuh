// Device-side engine description shared by every translation unit: arena layout, solver
// options and the EngineDev handle passed to the kernels by value.
#pragma once

#include "fluid.cuh"
#include "node_assembly.cuh"

namespace vf {

constexpr int kMaxRestart = 128;
constexpr int kMaxDenseN = 1024;  // dense inverse preconditioner: largest system
// asm_node_warp_kernel (assembly.cu): largest shared-memory row group it is launched with, and
// whether Jacobian launches of the thread-per-node path use it by default (VF_NODE_WARP overrides)
constexpr int kNodeWarpMaxSmem = 160 * 1024;
// Measured on B200, 0.99 M tetrahedra, residual + Jacobian (profiles/r2_variants_ab.json): global
// rows 1.98 ms, shared-memory rows 1.90, global rows + gather tables 1.59, both 0.758 ms.
constexpr bool kNodeWarpDefault = true;
// table-driven gathers of the tetrahedral node kernels (tet_tables.h; VF_TET_TABLES overrides)
constexpr bool kTetTablesDefault = true;
constexpr int kDenseNb = 8;       // ... and the pivot block of its Gauss-Jordan
// leading dimension of the fp32 inverse: rows padded to 128 bytes for aligned float4 loads
__host__ __device__ __forceinline__ int dense_ldp(int N) { return (N + 31) & ~31; }
constexpr int kInfoCount = 16;
enum InfoSlot { INFO_NUM_ITER = 0, INFO_ABS_ERR = 1, INFO_REL_ERR = 2, INFO_GMRES_ITERS = 3,
                INFO_GMRES_RESID = 4, INFO_MIN_AREA = 5, INFO_BNORM = 6 };

struct Layout {
  size_t off[32];   // offsets (doubles) of the public arrays inside a member block
  size_t cnt[32];
  size_t Dinv, V, w, z, H, cs, sn, g, y, xk;  // solver workspace
  size_t Pinv, Pf, Pscr, pstate;  // dense inverse preconditioner: fp64 work matrix, fp32
                                  // transposed inverse, mat-vec scratch, state
  size_t stride;    // member block size (doubles)
};

struct SolverOpts {
  double newton_abs_tol, newton_rel_tol;
  int newton_max_iter;
  double gmres_rel_tol, gmres_abs_tol;
  int gmres_max_iter;
  int is_static;
  int poly_degree;  // Neumann-series degree of the polynomial preconditioner (0 = block-Jacobi)
};

struct EngineDev {
  MeshView mesh;
  int d, N, n_fluid, ns, n_fsi, n_fsip, fluid_kind, idx_sep, contact, membrane, damping,
      restart, dense;  // dense != 0: the member blocks carry Pinv
  long long nnz;
  const double* s;
  const int* fsi_solid;   // area gather map (unique fluid DOFs)
  const int* fsi_fluid;
  const int* fsip_solid;  // pressure scatter map (unique solid DOFs)
  const int* fsip_fluid;
  // record-based in-CTA assembly (triangles, < 4096 cells): packed (node, cell) pair info with
  // GLOBAL cell ids (tables.build_tile_elem_tables over one tile), nodes with facet / BC work
  const unsigned* gpair;
  const int* touch;
  int n_touch;
  double* members;
  Layout L;
};


template <int D>
__device__ __forceinline__ PropView member_props(const EngineDev& E, double* mb) {
  const Layout& L = E.L;
  PropView p;
  p.rho = mb + L.off[VF_RHO];
  p.eta = mb + L.off[VF_ETA];
  p.emod = mb + L.off[VF_EMOD];
  p.scal = mb + L.off[VF_SCAL];
  p.emod_m = mb + L.off[VF_EMOD_M];
  p.nu_m = mb + L.off[VF_NU_M];
  p.th_m = mb + L.off[VF_TH_M];
  p.contact = E.contact;
  p.membrane = E.membrane;
  p.damping = E.damping;
  return p;
}

}  // namespace vf
