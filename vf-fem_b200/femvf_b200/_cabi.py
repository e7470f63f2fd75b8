"""
ctypes binding of the C ABI in ``include/vffem_b200.h`` (library ``lib/libvffem_b200.so``).

The library is the only compute path: importing this module fails loudly when it has not
been built (``__graft_entry__.build()`` / ``make -C vf-fem_b200``); there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VF_LIB_PATH: developer aid for A/B timing of two builds of the same library
LIB_PATH = os.environ.get('VF_LIB_PATH') or \
    os.path.join(os.path.dirname(_HERE), 'lib', 'libvffem_b200.so')

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class ProblemDesc(C.Structure):
    _fields_ = [
        ('dim', C.c_int32), ('nn', C.c_int32), ('ne', C.c_int32), ('nfp', C.c_int32),
        ('xyz_host', C.c_void_p), ('cells_host', C.c_void_p),
        ('brptr_host', C.c_void_p), ('bcol_host', C.c_void_p),
        ('n2e_ptr_host', C.c_void_p), ('n2e_host', C.c_void_p),
        ('n2f_ptr_host', C.c_void_p), ('n2f_host', C.c_void_p),
        ('pf_cell_host', C.c_void_p), ('pf_opp_host', C.c_void_p),
        ('bc_host', C.c_void_p),
        ('tile_start_host', C.c_void_p),
        ('ntiles', C.c_int32), ('tile_max_values', C.c_int32), ('tile_threads', C.c_int32),
        ('te_ptr_host', C.c_void_p), ('te_elem_host', C.c_void_p),
        ('pair_info_host', C.c_void_p), ('tile_desc_host', C.c_void_p),
        ('te_quad_host', C.c_void_p), ('tile_halo_host', C.c_void_p),
        ('max_tile_elems', C.c_int32), ('max_tile_pairs', C.c_int32),
        ('n_tile_halo', C.c_int32), ('max_tile_verts', C.c_int32),
        ('tile2_threads', C.c_int32), ('fan_ok', C.c_int32),
        ('gpair_host', C.c_void_p),
        ('n_fluid', C.c_int32), ('ns', C.c_int32), ('n_fsi', C.c_int32),
        ('s_host', C.c_void_p), ('fsi_solid_host', C.c_void_p), ('fsi_fluid_host', C.c_void_p),
        ('n_fsip', C.c_int32), ('fsip_solid_host', C.c_void_p), ('fsip_fluid_host', C.c_void_p),
        ('fluid_kind', C.c_int32), ('idx_sep', C.c_int32),
        ('contact', C.c_int32), ('membrane', C.c_int32), ('damping', C.c_int32),
        ('n_members', C.c_int32), ('gmres_restart', C.c_int32),
    ]


class SolverOpts(C.Structure):
    _fields_ = [
        ('newton_abs_tol', C.c_double), ('newton_rel_tol', C.c_double),
        ('newton_max_iter', C.c_int32),
        ('gmres_rel_tol', C.c_double), ('gmres_abs_tol', C.c_double),
        ('gmres_max_iter', C.c_int32),
        ('is_static', C.c_int32),
        ('poly_degree', C.c_int32), ('reserved', C.c_int32),
    ]


ARRAY_IDS = {name: i for i, name in enumerate([
    'u0', 'v0', 'a0', 'q0', 'p0', 'u1', 'v1', 'a1', 'q1', 'pf1', 'psub', 'psup', 'p1', 'area',
    'rho', 'eta', 'emod', 'emod_membrane', 'nu_membrane', 'th_membrane', 'scal', 'fprop',
    'F', 'J', 'dx', 'info',
])}

# every symbol include/vffem_b200.h declares
EXPORTED_SYMBOLS = [
    'vf_last_error', 'vf_device_count', 'vf_arena_bytes', 'vf_create', 'vf_destroy',
    'vf_array_info', 'vf_upload', 'vf_download', 'vf_csr_pattern', 'vf_nnz', 'vf_assemble',
    'vf_spmv', 'vf_linear_solve', 'vf_solve_state1', 'vf_fluid_solve', 'vf_integrate',
    'vf_integrate_host', 'vf_launch_count', 'vf_spmv_rows', 'vf_block_jacobi_setup',
    'vf_block_jacobi_apply', 'vf_multidot', 'vf_multi_axpy', 'vf_axpby',
    'vf_newmark_residual', 'vf_scale_rsqrt', 'vf_glottal_width_series',
    'vf_assemble_mix', 'vf_pressure_control_blocks', 'vf_set_fan_tables',
    'vf_props_changed', 'vf_ilu_setup', 'vf_ilu_factor', 'vf_ilu_apply',
    'vf_p2_create', 'vf_p2_destroy', 'vf_p2_nnz', 'vf_p2_assemble',
    'vf_band_setup', 'vf_band_factor', 'vf_band_solve',
]

_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()').  "
            "femvf_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.vf_last_error.restype = C.c_char_p
    lib.vf_device_count.restype = C.c_int
    lib.vf_arena_bytes.restype = C.c_size_t
    lib.vf_arena_bytes.argtypes = [C.POINTER(ProblemDesc)]
    lib.vf_create.argtypes = [C.POINTER(ProblemDesc), C.c_void_p, C.c_size_t, C.c_void_p,
                              C.POINTER(C.c_void_p)]
    lib.vf_destroy.argtypes = [C.c_void_p]
    lib.vf_destroy.restype = None
    lib.vf_array_info.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t)]
    lib.vf_upload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vf_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vf_csr_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vf_nnz.argtypes = [C.c_void_p]
    lib.vf_nnz.restype = C.c_int64
    lib.vf_assemble.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]
    lib.vf_spmv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vf_spmv_rows.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p]
    lib.vf_block_jacobi_setup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.vf_block_jacobi_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p]
    lib.vf_multidot.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vf_multi_axpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vf_axpby.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p,
                             C.c_size_t, C.c_void_p]
    lib.vf_scale_rsqrt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vf_assemble_mix.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int,
                                    C.c_void_p]
    lib.vf_pressure_control_blocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
    lib.vf_glottal_width_series.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                            C.c_void_p, C.c_void_p]
    lib.vf_newmark_residual.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    lib.vf_linear_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.POINTER(SolverOpts), C.c_void_p, C.c_void_p]
    lib.vf_solve_state1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double,
                                    C.POINTER(SolverOpts), C.c_void_p]
    lib.vf_fluid_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.vf_integrate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                 C.POINTER(SolverOpts), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vf_integrate_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                      C.POINTER(SolverOpts), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vf_set_fan_tables.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.vf_props_changed.argtypes = [C.c_void_p, C.c_int]
    lib.vf_ilu_setup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
    lib.vf_ilu_factor.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.vf_ilu_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vf_p2_create.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 10 + [C.c_int] + \
        [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
    lib.vf_band_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.vf_band_factor.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.vf_band_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vf_p2_destroy.argtypes = [C.c_void_p]
    lib.vf_p2_destroy.restype = None
    lib.vf_p2_nnz.argtypes = [C.c_void_p]
    lib.vf_p2_nnz.restype = C.c_longlong
    lib.vf_p2_assemble.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double] + \
        [C.c_void_p] * 11
    lib.vf_launch_count.argtypes = [C.c_void_p]
    lib.vf_launch_count.restype = C.c_int64
    _lib = lib
    return lib


class VFError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise VFError(load_library().vf_last_error().decode())
