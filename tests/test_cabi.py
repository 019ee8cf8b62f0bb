"""CPU: the C-ABI library builds/loads and exports every symbol include/b2m.h declares (no compute calls)."""
import os
import re

from box2mask_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b2m.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2m_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header():
    build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 27
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms            # the ctypes table mirrors the header one to one
    assert lib.b2m_version() == 3
    assert lib.b2m_error_string(-4) == b"unsupported shape"
    assert lib.b2m_hash_capacity(1000) == 2048 and lib.b2m_hash_capacity(0) == 1024
    assert lib.b2m_packed_weight_bytes(27, 96, 128, 0) == 27 * 128 * (128 + 64)   # one SW128 chunk + one SW64 remainder
    assert lib.b2m_map_pitch(129) == 256 and lib.b2m_map_pitch(0) == 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "box2mask_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_launch_list_command_layout():
    """The packed command records of ops.LaunchList match b2m_command_t: argument validation (which happens before any
    CUDA call) must see the values where the header says they are."""
    import ctypes

    from box2mask_b200.ops import LaunchList
    build.build()
    lib = _lib.load()
    assert LaunchList._S.size == 8 + 28 * 8 + 2 * 8           # int32 op, int32 stream, int64 a[28], double f[2]
    ll = LaunchList("cpu")

    def run():
        failed = ctypes.c_int64(-7)
        n, ll.n, ll.kinds = ll.n, 0, []
        return lib.b2m_run_commands(ll.cbuf, n, None, None, ctypes.byref(failed)), failed.value

    ll.add("bogus", 99, ())
    assert run() == (-1, 0)                                    # unknown op: invalid argument at command 0
    # b2m_copy_columns(src, src_ld, dst, dst_ld, n, width): n = 0 is a no-op, width 4 (not a multiple of 8) unsupported
    ll.add("copy_columns", LaunchList.OP_COPY_COLUMNS, (64, 8, 128, 8, 0, 8))
    ll.add("copy_columns", LaunchList.OP_COPY_COLUMNS, (64, 8, 128, 8, 5, 4))
    assert run() == (-4, 1)
    ll.add("copy_columns", LaunchList.OP_COPY_COLUMNS, (64, 8, 128, 8, -5, 8))
    assert run() == (-1, 0)                                    # negative row count
    ll.stream = 1
    ll.add("copy_columns", LaunchList.OP_COPY_COLUMNS, (64, 8, 128, 8, 0, 8))
    assert run() == (-1, 0)                                    # side-stream command without a side stream
    assert run() == (0, -1)                                    # empty list
