"""The C-ABI library builds, loads, and exports every symbol include/maxstyle_b200.h declares.
No compute calls (there is no GPU here); only pure host entry points are exercised."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "maxstyle_b200.h")


@pytest.fixture(scope="module")
def lib():
    from maxstyle_b200 import build, _lib
    build.build()
    return _lib.get_lib()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(maxstyle_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    from maxstyle_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding table and header disagree"


def test_library_is_sm100a_only_and_uses_256bit_accesses():
    from maxstyle_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    # Blackwell-only forms: 256-bit global loads/stores carrying a run-time L2 cache policy
    assert re.search(r"LDG\.E\.NA\.ENL2\.256", sass) and re.search(r"STG\.E\.NA\.ENL2\.256", sass)


def test_host_entry_points(lib):
    assert b"sm_100a" in lib.maxstyle_version()
    assert lib.maxstyle_strerror(0) == b"ok"
    assert b"workspace" in lib.maxstyle_strerror(3)
    n = lib.maxstyle_workspace_bytes(20, 64, 224, 224, 0, 0)
    assert n > 0 and n % 256 == 0
    assert lib.maxstyle_workspace_bytes(20, 64, 224, 224, 0, 1) % 256 == 0 and lib.maxstyle_workspace_bytes(20, 64, 224, 224, 0, 1) > 0   # NHWC
    assert lib.maxstyle_workspace_bytes(0, 64, 224, 224, 0, 0) == 0               # bad shape
    assert lib.maxstyle_workspace_bytes(4, 4, 1, 1, 0, 0) == 0                    # M < 2: identity case upstream
    assert lib.maxstyle_workspace_bytes(4, 4, 8, 8, 7, 0) == 0                    # unknown dtype


def test_bad_arguments_return_codes_not_crashes(lib):
    # null pointers are rejected before any CUDA call
    assert lib.maxstyle_stats(None, None, None, 4, 0, 4, 4, 8, 8, 0, 0, 1e-6, 0, None, 0, None) == 1
    assert lib.maxstyle_apply(None, None, None, 4, 0, None, None, 4, 4, 8, 8, 0, 0, 0, None) == 1
    assert lib.maxstyle_tables(None, None, 4, 4, 0, 4, 4, None, None, None, None, None, None, 0, None, None, None) == 1
    assert lib.maxstyle_stats(None, None, None, 4, 0, 4, 4, 8, 8, 3, 0, 1e-6, 0, None, 0, None) == 2   # dtype
    assert lib.maxstyle_stats(None, None, None, 4, 0, 4, 4, 1, 1, 0, 0, 1e-6, 0, None, 0, None) == 1   # M < 2


def test_new_entry_points_validate_before_launching(lib):
    """Peer-memory exchange, one-kernel multi-GPU forward, rank barrier and cross entropy: sizes and argument checks
    (everything here returns before any CUDA call, so it runs on the CPU box)."""
    # exchange buffer: two parities of 8-byte {value, epoch} words per (global row, mu|sig, channel) + two parities of barrier words
    assert lib.maxstyle_p2p_bytes(20, 64, 2) == (2 * 40 * 2 * 64 + 2 * 2) * 8
    assert lib.maxstyle_p2p_bytes(20, 64, 8) == (2 * 160 * 2 * 64 + 2 * 8) * 8
    assert lib.maxstyle_p2p_bytes(0, 64, 2) == 0 and lib.maxstyle_p2p_bytes(20, 64, 0) == 0
    assert lib.maxstyle_tables_p2p(None, 0, 2, None, None, None, None, None, 4, 8, 0, 4, 4, None, None, None, None, None, None, 0,
                                   None, None, None) == 1
    assert lib.maxstyle_rank_barrier(None, 0, 2, 4, 4, None, None, None) == 1
    assert lib.maxstyle_fwd_p2p(None, None, None, None, 4, 8, 0, None, None, None, None, None, None, None, None, 4, 4, 8, 8, 0, 0, 1,
                                1e-6, 0, None, 0, 2, None, None, 0, None) == 1
    assert lib.maxstyle_ce2d_workspace_bytes(20, 4, 224, 224) >= 256 + 148 * 8 * 4
    assert lib.maxstyle_ce2d_workspace_bytes(0, 4, 224, 224) == 0
    assert lib.maxstyle_ce2d_fwd(None, None, None, None, None, 2, 3, 4, 4, 0, 1, None, 0, None) == 1
    assert lib.maxstyle_ce2d_bwd(None, None, None, None, None, None, 2, 3, 4, 4, 0, 1, None) == 1


def test_step_struct_layout_matches_header():
    from maxstyle_b200._lib import StepStruct
    # 2 int32 + 4 double + 4 int32 + 10 pointers
    assert ctypes.sizeof(StepStruct) == 8 + 32 + 16 + 80
    assert StepStruct.lr.offset == 8 and StepStruct.t.offset == 40 and StepStruct.step_dev.offset == 56
