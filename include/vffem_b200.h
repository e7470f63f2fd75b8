/*
 * vffem_b200 -- C ABI of the B200-native hot path of femvf.forward.integrate.
 *
 * The reference (jon-deng/vf-fem) is pure Python; it has no FFI seam of its own
 * (SURVEY.md section 8b).  Each entry point below replaces the third-party native call the
 * reference makes at the cited site; INTEGRATION.md shows the ctypes stubs a maintainer
 * would add at those sites.
 *
 * Conventions: plain pointers and sizes only.  Pointers named *_host are host memory,
 * *_dev device memory.  `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Every function returns 0 on success, non-zero on error; the message is
 * available from vf_last_error().  Nothing here falls back to the CPU: if no CUDA device
 * is usable the calls fail.
 *
 * Data layout (all fp64, indices int32): vector DOFs node-major interleaved (d*i + c);
 * J_uu in scalar CSR with the canonical pattern (columns ascending, full d x d block per
 * vertex pair sharing a cell, explicit zeros kept) stored block-row by block-row; DG0
 * properties one value per cell; ensemble members are stored member-major (each member's
 * arrays contiguous).
 */
#ifndef VFFEM_B200_H
#define VFFEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vf_engine vf_engine;

/* Mesh, index tables (built once on the host: femvf_b200/tables.py) and model switches. */
typedef struct vf_problem_desc {
  int32_t dim, nn, ne, nfp;   /* dimension, #vertices, #cells, #pressure facets */
  const double* xyz_host;     /* (dim, nn) SoA coordinates */
  const int32_t* cells_host;  /* (dim+1, ne) SoA connectivity */
  const int32_t* brptr_host;  /* (nn+1) node graph = block pattern of J */
  const int32_t* bcol_host;   /* (nnzb) */
  const int32_t* n2e_ptr_host;
  const int32_t* n2e_host;    /* node -> cell*4 + local index */
  const int32_t* n2f_ptr_host;
  const int32_t* n2f_host;    /* node -> pressure facet*4 + local index in parent cell */
  const int32_t* pf_cell_host;
  const int32_t* pf_opp_host;
  const uint8_t* bc_host;     /* (dim*nn) Dirichlet flag per DOF */
  const int32_t* tile_start_host; /* (ntiles+1) node ranges of the assembly tiles */
  int32_t ntiles;
  int32_t tile_max_values;    /* max #doubles of J in one tile (shared-memory CSR slice) */
  int32_t tile_threads;       /* CTA size of the tile kernel (>= max nodes per tile) */
  /* two-phase tile kernel (triangles): cells touching each tile, packed (node, cell) pair info
   * (femvf_b200/tables.py build_tile_elem_tables); te_ptr_host == NULL disables it */
  const int32_t* te_ptr_host;     /* (ntiles+1) */
  const int32_t* te_elem_host;    /* (te_ptr[ntiles]) */
  const uint32_t* pair_info_host; /* (n2e_ptr[nn]) */
  const int32_t* tile_desc_host;  /* (ntiles, 12): i0 te0 pair0 blk0 halo0 nT|nH<<16 ncell|npair<<16 nblk cell_lo cell_cnt 0 0 */
  const int32_t* te_quad_host;    /* (te_ptr[ntiles], 4): local slots of the cell's 3 vertices + cell id */
  const int32_t* tile_halo_host;  /* (n_tile_halo) halo vertices of the tiles, ascending per tile */
  int32_t max_tile_elems;
  int32_t max_tile_pairs;
  int32_t n_tile_halo;
  int32_t max_tile_verts;     /* max own + halo vertices of a tile */
  int32_t tile2_threads;
  int32_t fan_ok;             /* n2e lists are counter-clockwise fans (tables.order_fans_2d) */
  /* pair_info of ONE tile covering the whole mesh (cell index = cell id; needs ne < 4096): lets
   * the per-member time-loop kernel assemble by records too.  NULL disables. */
  const uint32_t* gpair_host;     /* (n2e_ptr[nn]) */
  /* 1D fluid + FSI map (models/fsi.py:18-88) */
  int32_t n_fluid, ns, n_fsi;
  const double* s_host;           /* (n_fluid, ns) arclength coordinates */
  /* area gather fluid_area[fsi_fluid[k]] = solid_area[fsi_solid[k]] (fsi.py:69-70); fluid DOFs
   * unique (the host keeps the last occurrence, numpy's assignment semantics) */
  const int32_t* fsi_solid_host;  /* (n_fsi) scalar solid DOFs */
  const int32_t* fsi_fluid_host;  /* (n_fsi) fluid DOFs */
  /* pressure scatter solid_p[fsip_solid[k]] = fluid_p[fsip_fluid[k]] (fsi.py:66-67); solid DOFs
   * unique (last occurrence kept) */
  int32_t n_fsip;
  const int32_t* fsip_solid_host;
  const int32_t* fsip_fluid_host;
  int32_t fluid_kind;         /* 0 area-ratio sep, 1 fixed sep, 2 smooth-min sep */
  int32_t idx_sep;
  /* model switches */
  int32_t contact;            /* NodalContactModel semantics */
  int32_t membrane;           /* KelvinVoigtWEpithelium membrane term */
  int32_t damping;            /* 0 Kelvin-Voigt (form.py:965-990), 1 Rayleigh (form.py:918-956) */
  /* ensemble / solver workspace */
  int32_t n_members;
  int32_t gmres_restart;
} vf_problem_desc;

/* Newton options (solverconst.py:1-6) and linear-solver controls. */
typedef struct vf_solver_opts {
  double newton_abs_tol, newton_rel_tol;
  int32_t newton_max_iter;
  double gmres_rel_tol, gmres_abs_tol;
  int32_t gmres_max_iter;
  int32_t is_static;          /* static.py:68-168: u0 == u1, v0 = a0 = 0 */
  int32_t poly_degree;        /* Neumann-series degree of the polynomial preconditioner (0..8) */
  int32_t reserved;
} vf_solver_opts;

/* Named per-member arrays inside the arena. */
enum vf_array_id {
  VF_U0 = 0, VF_V0, VF_A0, VF_Q0, VF_P0,      /* state0: u, v, a (N), q (n_fluid), p (n_fluid*ns) */
  VF_U1, VF_V1, VF_A1, VF_Q1, VF_PF1,         /* state1 */
  VF_PSUB, VF_PSUP,                           /* control (n_fluid each) */
  VF_P1,                                      /* solid control 'p1' (nn) */
  VF_AREA,                                    /* fluid control 'area' (n_fluid*ns) */
  VF_RHO, VF_ETA, VF_EMOD,                    /* DG0 (ne) */
  VF_EMOD_M, VF_NU_M, VF_TH_M,                /* membrane DG0 (ne) */
  VF_SCAL,                                    /* nu, ycontact, kcontact, ncontact[3], ymid, rayleigh_m, rayleigh_k, pad (10) */
  VF_FPROP,                                   /* (n_fluid, 5): rho_air r_sep area_lb zeta_min zeta_sep */
  VF_F,                                       /* residual F_u (N) */
  VF_J,                                       /* CSR values of J_uu (nnz) */
  VF_DX,                                      /* last Newton update (N) */
  VF_INFO,                                    /* num_iter abs_err rel_err gmres_iters gmres_resid ... (16) */
  VF_ARRAY_COUNT
};

const char* vf_last_error(void);
int vf_device_count(void);

/* ---- lifecycle ------------------------------------------------------------------- */
/* Bytes of device memory the engine needs; the caller allocates them (e.g. a torch uint8
 * tensor) and hands the pointer to vf_create. */
size_t vf_arena_bytes(const vf_problem_desc* desc);
int vf_create(const vf_problem_desc* desc, void* arena_dev, size_t arena_bytes, void* stream,
              vf_engine** out);
void vf_destroy(vf_engine* e);

/* Byte offset into the arena and element count of a named per-member array. */
int vf_array_info(const vf_engine* e, int array_id, int member, size_t* byte_offset,
                  size_t* count);
/* Host <-> device copies of a named array (cudaMemcpyAsync on `stream` + synchronize). */
int vf_upload(vf_engine* e, int array_id, int member, const double* src_host, size_t count,
              void* stream);
int vf_download(vf_engine* e, int array_id, int member, double* dst_host, size_t count,
                void* stream);
/* Scalar CSR pattern implied by the node graph (host output; sizes N+1 and nnz). */
int vf_csr_pattern(const vf_engine* e, int32_t* rowptr_host, int32_t* colidx_host);
int64_t vf_nnz(const vf_engine* e);

/* ---- the hot path ---------------------------------------------------------------- */
/* FenicsModel.assem_res / assem_dres_dstate1 (models/transient.py:363-406): assemble F_u
 * (flags & 1) and/or J_uu (flags & 2) of `member` from its current state1/state0/controls/
 * props with Dirichlet rows applied.  Replaces dfn.assemble + bc.apply
 * (models/assemblyutils.py:49-50, transient.py:379-380, 398-399). */
int vf_assemble(vf_engine* e, int member, int flags, double dt, int is_static, void* stream);

/* Tables of the node-centric fan assembly kernel (triangles whose vertex stars are single
 * counter-clockwise fans; femvf_b200/tables.py build_fan_tables).  Once set, vf_assemble /
 * vf_assemble_mix run asm_fan_kernel: one thread per vertex walks its fan out of shared memory;
 * the tile's ring table, cell properties and CSR slice move by bulk asynchronous copies.  Still
 * the same dfn.assemble call sites (models/assemblyutils.py:49-50).
 *   desc_host  (ntiles, 12) int32: i0, nT | nH << 16, halo0, ring0, rows, tcell0, padded cell
 *              count, brptr[i0], #blocks of the tile, 0, 0, 0
 *   ring_host  (n_ring) uint32: per tile rows x tile_nodes words (header row + ring rows)
 *   halo_host  (n_halo) int32: ring vertices outside each tile's own node range
 *   tcell_host (n_tcell) int32: cells touched by each tile, lists padded to even counts
 * The tables are copied into a device allocation owned by the engine, which also holds a
 * tile-ordered copy of emod / eta / rho per member (refreshed when vf_upload changes them or
 * vf_props_changed is called). */
int vf_set_fan_tables(vf_engine* e, int tile_nodes, int ntiles, const int32_t* desc_host,
                      const uint32_t* ring_host, size_t n_ring, const int32_t* halo_host,
                      size_t n_halo, const int32_t* tcell_host, size_t n_tcell, int max_verts,
                      int max_rows, int max_cells, int max_blocks, void* stream);
/* Tell the engine that VF_RHO / VF_ETA / VF_EMOD of `member` (-1: every member) were written
 * directly in the arena (not through vf_upload). */
int vf_props_changed(vf_engine* e, int member);

/* The state0 sensitivities of FenicsModel.assem_dres_dstate0 (models/transient.py:408-421:
 * assemble_derivative(form, 'state/u0' | 'state/v0' | 'state/a0'), no bc.apply).  F_u depends on
 * state0 only through v_nmk, a_nmk, so each block is a mix of the matrices the Jacobian is made
 * of: VF_J := coef4[0] K + coef4[1] C + coef4[2] M + coef4[3] K_p on the Jacobian's CSR pattern,
 * C = dF_u/dv1, M = dF_u/da1; Dirichlet rows only when apply_bc != 0. */
int vf_assemble_mix(vf_engine* e, int member, double dt, const double* coef4, int apply_bc,
                    void* stream);

/* FenicsModel.assem_dres_dcontrol (models/transient.py:423-435): dF_u/dp1 of the follower
 * pressure term.  One (dim x 1) block per ordered pair (a, b) of vertices of every pressure
 * facet, written as nfp * dim * dim blocks of dim doubles to out_dev in the order
 * facet-major, a, b (local facet vertex order = parent cell order without the opposite
 * vertex); rows_host / cols_host (optional, nfp*dim*dim each) receive the vertex a and b. */
int vf_pressure_control_blocks(vf_engine* e, int member, double* out_dev, int32_t* rows_host,
                               int32_t* cols_host, void* stream);

/* y = J_uu x for `member` (PETSc MatMult; transient.py:488-489).  x, y: device, N doubles. */
int vf_spmv(vf_engine* e, int member, const double* x_dev, double* y_dev, void* stream);

/* Grid-wide Krylov building blocks for one large mesh or one partition of it (the PETSc KSP
 * internals behind dfn.solve, transient.py:487): the same products on a node-row range, the
 * block-Jacobi preconditioner, V^T w for a set of basis vectors (fixed-order reductions:
 * bit-reproducible), w -= V h, and y = alpha x + beta y.  femvf_b200/distributed.py drives a
 * GMRES over them, with NCCL halo exchange / all-reduce between partitions.
 * vf_multidot needs a scratch buffer of at least 592 * nvec doubles. */
int vf_spmv_rows(vf_engine* e, int member, const double* x_dev, double* y_dev, int node0,
                 int node1, void* stream);
int vf_block_jacobi_setup(vf_engine* e, int member, int node0, int node1, void* stream);
int vf_block_jacobi_apply(vf_engine* e, int member, const double* r_dev, double* z_dev, int node0,
                          int node1, void* stream);
int vf_multidot(vf_engine* e, const double* V_dev, size_t ldv, int nvec, const double* w_dev,
                size_t n, double* out_dev, double* scratch_dev, size_t scratch_count,
                void* stream);
int vf_multi_axpy(vf_engine* e, const double* V_dev, size_t ldv, int nvec, const double* h_dev,
                  double* w_dev, size_t n, void* stream);
int vf_axpby(vf_engine* e, double alpha, const double* x_dev, double beta, double* y_dev, size_t n,
             void* stream);

/* Glottal-width time series (postprocess/solid.py:487-501 MeanGlottalWidth evaluated by
 * postprocess/base.py:138-161 TimeSeries over a StateFile): out[t] = min over the fluid area
 * vector obtained from the stored displacement u_hist[t] (nt rows of ldu >= N doubles, device)
 * through the FSI area map, using the member's ymid and current fluid area as the base. */
int vf_glottal_width_series(vf_engine* e, int member, int nt, const double* u_hist_dev, size_t ldu,
                            double* out_dev, void* stream);

/* Banded LU of J_uu (no pivoting) in a bandwidth-reducing ordering: the direct stand-in for the
 * reference's sparse LU (dfn.solve(A, x, b, 'petsc'), models/transient.py:487, static.py:140) on
 * meshes of up to a few 1e4 DOF that do not fit the one-CTA solver.
 *   perm_host (dim * nn) int32: band index of every scalar DOF (a permutation; reverse
 *   Cuthill-McKee of the node graph in femvf_b200/gridsolve.py); half_bandwidth = max |perm[i] -
 *   perm[j]| over the non-zeros.  vf_band_factor scatters the member's current J into the band
 *   and factorises it in place; vf_band_solve computes x = J^-1 b (device vectors, may alias). */
int vf_band_setup(vf_engine* e, const int32_t* perm_host, int half_bandwidth, void* stream);
int vf_band_factor(vf_engine* e, int member, void* stream);
int vf_band_solve(vf_engine* e, const double* b_dev, double* x_dev, void* stream);

/* P2 (6-node) triangle residual + Jacobian assembly: the P2 extension named by BASELINE.json
 * (north_star subsystem 1, configs[2]); the reference itself is P1 only (equations/form.py:521-524).
 * Same call sites as vf_assemble (models/assemblyutils.py:49-50, models/transient.py:363-406) on a
 * P2 function space: F_u = M a_nmk + C v_nmk + K u1 + follower pressure, J = ca M + cv C + K + K_p,
 * Dirichlet rows applied.  Self-contained object; every vector is a caller-owned device array.
 *   coords_host (nn, 2) incl. mid-edge nodes; cells6_host (ne, 6), local order v0 v1 v2 e(12) e(02) e(01)
 *   brptr/bcol: node graph (block CSR pattern, columns ascending, <= 31 blocks per row)
 *   n2e_ptr/n2e: per node its (cell * 8 + local node) pairs; n2e_slots: per pair the CSR slots of
 *                the cell's six nodes in that node's row, 5 bits each
 *   n2f_ptr/n2f: per node its pressure edges (edge * 4 + position 0/1/2 = va/vb/mid);
 *                n2f_pair: index of the (node, parent cell) pair in n2e
 *   pf_cell (nfp), pf_loc (nfp, 3) local nodes (va, vb, mid) in the parent cell,
 *   pf_geo (nfp, 3) = outward unit normal, edge length; fixed_host (nn) Dirichlet flags
 *   order_host (nn): thread -> node map, the n_class0 vertex nodes first, then the mid-edge nodes
 * J_dev receives vf_p2_nnz doubles in scalar CSR order of the node-major interleaved DOFs. */
typedef struct vf_p2 vf_p2;
int vf_p2_create(int nn, int ne, const double* coords_host, const int32_t* cells6_host,
                 const int32_t* brptr_host, const int32_t* bcol_host, const int32_t* n2e_ptr_host,
                 const int32_t* n2e_host, const uint32_t* n2e_slots_host,
                 const int32_t* n2f_ptr_host, const int32_t* n2f_host,
                 const int32_t* n2f_pair_host, int nfp, const int32_t* pf_cell_host,
                 const int32_t* pf_loc_host, const double* pf_geo_host, const uint8_t* fixed_host,
                 const int32_t* order_host, int n_class0, void* stream, vf_p2** out);
void vf_p2_destroy(vf_p2* p);
long long vf_p2_nnz(const vf_p2* p);
int vf_p2_assemble(vf_p2* p, int flags, double dt, double nu, const double* emod_dev,
                   const double* eta_dev, const double* rho_dev, const double* u1_dev,
                   const double* u0_dev, const double* v0_dev, const double* a0_dev,
                   const double* p1_dev, double* F_dev, double* J_dev, void* stream);

/* Multicolour block ILU(0) of J_uu on node rows [node0, node1): the cuSPARSE-free stand-in for
 * the reference's sparse LU (dfn.solve(A, x, b, 'petsc'), models/transient.py:487, static.py:140)
 * as preconditioner of the grid-wide GMRES on meshes that do not fit one CTA.
 *   rows_host      (node1 - node0) int32: the nodes of the range grouped by colour
 *   color_ptr_host (ncolors + 1) int32: offsets of the colour classes in rows_host
 *   color_host     (nn) int32: colour of every local node, -1 outside the range; no two nodes of
 *                  one colour may share a cell (femvf_b200/tables.py color_node_graph)
 * vf_ilu_factor copies the member's current J and factorises it in place (ncolors launches);
 * vf_ilu_apply computes z = U^-1 L^-1 r on the DOFs of the range (2 ncolors - 1 launches; r and z
 * are full local vectors and may alias). */
int vf_ilu_setup(vf_engine* e, int node0, int node1, int ncolors, const int32_t* rows_host,
                 const int32_t* color_ptr_host, const int32_t* color_host, void* stream);
int vf_ilu_factor(vf_engine* e, int member, void* stream);
int vf_ilu_apply(vf_engine* e, const double* r_dev, double* z_dev, void* stream);

/* y = x / sqrt(s), s = *s2_dev - sum_{i<nsub} sub_dev[i]^2, all read on the device; s is also
 * stored to *s_out_dev when non-null (must not alias s2_dev).  Krylov-vector normalisation
 * with the squared norm left on the device by vf_multidot (KSPGMRES's VecNormalize without the
 * host round trip); the subtraction is the norm update of the second Gram-Schmidt pass. */
int vf_scale_rsqrt(vf_engine* e, const double* x_dev, const double* s2_dev, const double* sub_dev,
                   int nsub, double* s_out_dev, double* y_dev, size_t n, void* stream);

/* Nodal Newmark residuals F_v = v1 - newmark_v(u1, u0, v0, a0, dt), F_a = a1 - newmark_a(...)
 * of FenicsModel.assem_res (models/transient.py:374-377, equations/newmark.py:8-73), from the
 * member's device-resident state.  fv, fa: device, N doubles. */
int vf_newmark_residual(vf_engine* e, int member, double dt, double* fv_dev, double* fa_dev,
                        void* stream);

/* Solve J_uu x = b with the block-Jacobi preconditioned GMRES that stands in for the PETSc
 * LU of dfn.solve(A, x, b, 'petsc') (transient.py:487).  b, x: device, N doubles.
 * info_host[0] = iterations, [1] = final residual norm, [2] = ||b||. */
int vf_linear_solve(vf_engine* e, int member, const double* b_dev, double* x_dev,
                    const vf_solver_opts* opts, double* info_host, void* stream);

/* FenicsModel.solve_state1 (transient.py:441-468): Newton on F_u(u1) = 0 from the guess in
 * VF_U1, then v1, a1 from the Newmark relations, for members [member0, member0+count). */
int vf_solve_state1(vf_engine* e, int member0, int count, double dt, const vf_solver_opts* opts,
                    void* stream);

/* JaxModel.solve_state1 (transient.py:667-672): (q, p) from VF_AREA, VF_PSUB, VF_PSUP into
 * VF_Q1 / VF_PF1. */
int vf_fluid_solve(vf_engine* e, int member0, int count, void* stream);

/* forward.integrate_steps with ExplicitFSIModel.solve_state1 (forward.py:139-186,
 * transient.py:899-920), device resident for all members: nsteps steps from state0.
 *   dts_host        (nsteps) time step sizes
 *   controls_host   (ncontrols, 2, n_fluid) psub, psup per control; control min(n, ncontrols-1)
 *                   is used at step n (forward.py:170)
 *   hist_state_dev  optional (n_members, nsteps+1, 3N + n_fluid + n_fluid*ns) state history,
 *                   row 0 = initial state (forward.py:75-86); NULL to skip
 *   hist_info_dev   optional (n_members, nsteps+1, 4): num_iter, abs_err, rel_err, min area
 * On return state0 holds the final state of every member. */
int vf_integrate(vf_engine* e, int nsteps, const double* dts_host, int ncontrols,
                 const double* controls_host, const vf_solver_opts* opts,
                 double* hist_state_dev, double* hist_info_dev, void* stream);

/* Same with host buffers end to end: uploads the initial state (n_members, state_size) and
 * per-member DG0 properties emod/eta (n_members, ne each; NULL keeps the resident ones),
 * integrates, downloads the final state and the (n_members, nsteps+1, 4) info series. */
int vf_integrate_host(vf_engine* e, int nsteps, const double* dts_host, int ncontrols,
                      const double* controls_host, const vf_solver_opts* opts,
                      const double* ini_state_host, const double* emod_host,
                      const double* eta_host, double* fin_state_host, double* info_series_host,
                      void* stream);

/* Number of kernel launches issued by this engine so far (bench.py's gpu_launches). */
int64_t vf_launch_count(const vf_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* VFFEM_B200_H */
