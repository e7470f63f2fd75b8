"""The C-ABI shared library loads and exports every symbol include/vffem_b200.h declares
(no compute calls: this runs without a GPU)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    from femvf_b200 import _cabi
    return _cabi


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'vffem_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vf_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol(built):
    lib = built.load_library()
    names = header_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(built.EXPORTED_SYMBOLS) == names


def test_array_ids_match_header_enum(built):
    text = open(os.path.join(ROOT, 'include', 'vffem_b200.h')).read()
    body = re.search(r'enum vf_array_id \{(.*?)\};', text, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    ids = [t.strip().split('=')[0].strip() for t in body.split(',') if t.strip()]
    assert ids[-1] == 'VF_ARRAY_COUNT'
    assert len(ids) - 1 == len(built.ARRAY_IDS)


def test_struct_sizes_are_plain_c(built):
    # pointers and sizes only: the descriptor is a POD of int32 / pointer fields
    assert ctypes.sizeof(built.SolverOpts) == 56  # static_assert-ed in csrc/vffem_b200.cu
    assert ctypes.sizeof(built.ProblemDesc) % 8 == 0


def test_error_reporting_without_gpu(built):
    import torch
    lib = built.load_library()
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a CPU-only box")
    assert lib.vf_device_count() == 0
    desc = built.ProblemDesc()
    assert lib.vf_arena_bytes(ctypes.byref(desc)) == 0
    assert b'dim' in lib.vf_last_error()
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    from helpers import mesh_tuples
    model = transient.FenicsModel(slr.KelvinVoigt(*mesh_tuples()['square5']()))
    with pytest.raises(built.VFError):
        model.assem_res()  # no CUDA device -> loud failure, never a CPU fallback
