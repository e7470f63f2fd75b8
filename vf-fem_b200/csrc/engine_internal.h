// Internal declarations shared by the translation units of libvffem_b200.so (not part of the
// C ABI): the engine object behind the opaque vf_engine handle and small host helpers.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vffem_b200.h"
#include "engine_types.cuh"

namespace vf {

// sets the thread's last-error message (vf_last_error) and returns 1
int fail(const std::string& msg);

#define VF_CUDA(call)                                                            \
  do {                                                                           \
    cudaError_t _e = (call);                                                     \
    if (_e != cudaSuccess)                                                       \
      return ::vf::fail(std::string(#call) + ": " + cudaGetErrorString(_e));     \
  } while (0)

struct FanTablesDev {
  const int4* desc;      // (ntiles, 3) int4
  const unsigned* ring;  // header + ring words, per tile rows x tile_nodes
  const int* halo;       // ring vertices outside each tile's node range
  const int* tcell;      // cells touched by each tile (padded to even counts)
  double* mat;           // per member: tile-ordered [emod | eta | rho] blocks (3 * n_tcell)
  const unsigned char* prec;  // per tile: [descriptor | pad | halo vertex ids], prec_stride bytes
  int prec_stride;
  size_t n_tcell;
  int tile_nodes, ntiles, max_verts, max_rows, max_cells, max_blocks;
};

// Boundary-node record of facet_bc_fast_kernel (triangles, follower pressure + Dirichlet only):
// everything the node's exterior-facet terms need that does not depend on the state, so that the
// kernel's dependent-load chain is record -> nodal values / row blocks -> stores.
struct FacetRec2D {
  int node, b0, deg, self;   // node id, first block of its row, blocks in the row, own slot
  int bc, nfac, pad0, pad1;  // bc: bit c set = component c is fixed
  struct Facet {
    int nd[3];               // nodes of the parent cell
    int slot[3];             // their CSR slots in this node's block row
    int a, o;                // local index of the node / of the vertex opposite the facet
    double N[2], meas;       // outward unit normal, facet length
    double G[3][2];          // P1 gradients of the parent cell
  } f[2];
};

// multicolour block ILU(0) (ilu.cu): factor storage and colouring tables, one device allocation
struct IluState {
  char* mem = nullptr;
  double* LU = nullptr;    // pattern and layout of J
  double* Dinv = nullptr;  // inverses of the diagonal blocks of U
  int* color = nullptr;    // per node: colour, -1 outside the solved range
  int* rows = nullptr;     // nodes of the solved range grouped by colour
  int node0 = 0, node1 = 0, ncolors = 0;
  std::vector<int32_t> color_ptr;
};
void ilu_release(struct ::vf_engine* e);

// banded LU (band.cu): row-major band storage in a bandwidth-reducing ordering
struct BandState {
  char* mem = nullptr;
  double* AB = nullptr;  // N x (2 b + 1)
  int* perm = nullptr;   // band index of scalar DOF i
  double* x = nullptr;   // work vector in the band ordering
  int b = 0, N = 0;
};
void band_release(struct ::vf_engine* e);

}  // namespace vf

struct vf_engine {
  vf_problem_desc desc;  // scalar fields only are valid after create
  vf::EngineDev dev;
  char* arena;
  size_t arena_bytes;
  int* tile_start_dev;
  int* te_ptr_dev;
  int* te_elem_dev;
  unsigned* pair_info_dev;
  int4* tile_desc_dev;
  int4* te_quad_dev;
  int* tile_halo_dev;
  int* touch_dev;
  int n_touch;
  vf::FacetRec2D* facet_rec_dev;  // one per touched node, or null (generic facet_bc_kernel)
  bool two_phase;
  bool fan_ok;
  std::vector<int32_t> brptr, bcol;
  std::vector<int32_t> pf_nodes;  // (nfp, dim): vertices of every pressure facet, parent-cell order
  int member_threads;
  int64_t launches;
  // node-centric fan assembly (vf_set_fan_tables): tables in their own device allocation
  vf::FanTablesDev fan;
  void* fan_mem;
  std::vector<char> fan_dirty;  // per member: the tile-ordered property copy is stale
  int fan_max_wblocks;          // most CSR blocks owned by 32 consecutive nodes
  void* node_mem = nullptr;     // tetrahedral gather tables (tet_tables.h), own allocation
  int node_warp_blocks = 0;     // same count over ALL nodes (asm_node_warp_kernel's shared memory)
  bool pool_user;               // this engine holds a reference on the raised mempool threshold
  vf::IluState ilu;
  vf::BandState band;
};

namespace vf {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Dense inverse preconditioner of the member solver (member_solver.cuh).  Returns 0 (off), 1 (used
// for static solves only: the default) or 2 (also in the time loop: VF_DENSE_PREC=1).
// Measured on config 1/2 (N = 296, profiles/README.md):
//   * static solve with contact (no mass term, stiff): 2326 -> 11 GMRES iterations,
//     18.2 -> 3.8 ms: this is the PETSc-LU stand-in, on by default when the storage is small;
//   * transient steps (inverse of the state-independent part, kept across launches): 10 -> 6
//     iterations, but one mat-vec with the 350 KB fp32 inverse costs ~24 k cycles on ONE SM
//     (L2 -> SM at ~15 B/clk with 8 warps) against ~10 k for the degree-3 polynomial step out
//     of shared memory: 5.9 k vs 7.1 k steps/s on the same box.  Opt-in only.
inline int dense_prec_mode(const vf_problem_desc& d) {
  const char* env = getenv("VF_DENSE_PREC");
  const int want = env ? atoi(env) : -1;
  if (want == 0) return 0;
  const size_t N = (size_t)d.dim * d.nn;
  if (N > (size_t)kMaxDenseN) return 0;
  // the pivot panels (2 * kDenseNb * N doubles) live in the Krylov basis storage
  if ((size_t)d.gmres_restart + 1 < 2 * (size_t)kDenseNb) return 0;
  const size_t bytes = 12 * N * N * (size_t)d.n_members;
  if (want == 1) return bytes <= ((size_t)24 << 30) ? 2 : 0;
  return bytes <= ((size_t)2 << 30) ? 1 : 0;
}
inline bool dense_prec_enabled(const vf_problem_desc& d) { return dense_prec_mode(d) != 0; }

// assembly.cu: per-engine kernel attributes (dynamic shared memory opt-in)
int assembly_configure(vf_engine* e, const vf_problem_desc& d, bool two_phase);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline double* member_array(vf_engine* e, int id, int member) {
  return e->dev.members + (size_t)member * e->dev.L.stride + e->dev.L.off[id];
}

inline SolverOpts to_opts(const vf_solver_opts* o) {
  SolverOpts s;
  if (o) {
    s.newton_abs_tol = o->newton_abs_tol;
    s.newton_rel_tol = o->newton_rel_tol;
    s.newton_max_iter = o->newton_max_iter;
    s.gmres_rel_tol = o->gmres_rel_tol;
    s.gmres_abs_tol = o->gmres_abs_tol;
    s.gmres_max_iter = o->gmres_max_iter;
    s.is_static = o->is_static;
    s.poly_degree = std::min(std::max(o->poly_degree, 0), 8);
  } else {
    s.newton_abs_tol = 1e-8;   // solverconst.py:1-6
    s.newton_rel_tol = 1e-10;
    s.newton_max_iter = 50;
    s.gmres_rel_tol = 1e-13;
    s.gmres_abs_tol = 0.0;
    s.gmres_max_iter = 2000;
    s.is_static = 0;
    s.poly_degree = 3;
  }
  return s;
}

}  // namespace vf
