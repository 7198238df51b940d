"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol declared in
include/gsrast.h, answers size queries and rejects malformed arguments before touching CUDA."""
import ctypes
import os
import re
from ctypes import byref, c_int, c_size_t

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsrast.h")
LIB = os.path.join(ROOT, "wild-gaussians_b200", "lib", "libgsrast.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsr_[a-z_0-9]+)\s*\(", src)) - {"gsr_alloc_fn"})


def test_header_declares_expected_surface():
    fns = declared_functions()
    for must in ("gsr_forward", "gsr_forward_sizes", "gsr_forward_geometry", "gsr_binning_sizes",
                 "gsr_forward_render", "gsr_backward", "gsr_mark_visible", "gsr_last_error", "gsr_abi_version"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    for fn in declared_functions():
        assert hasattr(lib, fn), f"{fn} declared in gsrast.h but not exported"
    assert lib.gsr_abi_version() == 2


def test_no_torch_types_in_boundary():
    src = open(HEADER).read()
    assert "torch" not in src.replace("no torch types", "") and "at::" not in src and "#include <cuda" not in src


def test_size_queries_need_no_gpu():
    lib = ctypes.CDLL(LIB)
    lib.gsr_forward_sizes.argtypes = [c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]
    g, i = c_size_t(0), c_size_t(0)
    assert lib.gsr_forward_sizes(1000, 0, 640, 480, byref(g), byref(i)) == 0
    assert g.value >= 1000 * (32 + 4 + 4 + 8 + 16) and i.value >= 640 * 480 * 8
    g2 = c_size_t(0)
    assert lib.gsr_forward_sizes(1000, 16, 640, 480, byref(g2), byref(i)) == 0
    assert g2.value > g.value      # SH path keeps rgb + clamp flags
    lib.gsr_binning_sizes.argtypes = [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]
    b, s = c_size_t(0), c_size_t(0)
    assert lib.gsr_binning_sizes(1000, 640, 480, 5000, 1500, byref(b), byref(s)) == 0
    assert b.value >= 5000 * 4 and s.value >= 1500 * 16
    lib.gsr_backward_scratch_bytes.restype = c_size_t
    assert lib.gsr_backward_scratch_bytes(1000) >= 48000


def test_invalid_arguments_are_rejected_with_a_message():
    import diff_gaussian_rasterization._C as C   # ctypes structures
    lib = C._lib
    a = C.GsrForwardArgs()
    a.P, a.W, a.H = 10, 64, 64            # required pointers left NULL
    R = c_int(0)
    N1 = c_int(0)
    rc = lib.gsr_forward_geometry(byref(a), None, None, None, byref(R), byref(N1))
    assert rc == -1
    assert b"NULL" in lib.gsr_last_error()
    assert lib.gsr_forward_sizes(-1, 0, 64, 64, None, None) == -1
    b = C.GsrBackwardArgs()
    b.P, b.W, b.H = 10, 64, 64
    assert lib.gsr_backward(byref(b), None) == -1
    # P == 0 is legal and does nothing
    a0 = C.GsrForwardArgs(); a0.P, a0.W, a0.H = 0, 64, 64
    assert lib.gsr_forward_geometry(byref(a0), None, None, None, byref(R), byref(N1)) == 0 and R.value == 0


def test_peer_reduction_entry_point_validates_its_arguments():
    """gsr_backward_partials_peers: without peer pointers and without a multicast address there is nothing to reduce
    into -- rejected with a message, before any CUDA call."""
    import diff_gaussian_rasterization._C as C
    lib = C._lib
    b = C.GsrBackwardArgs()
    b.P, b.W, b.H = 10, 64, 64
    assert lib.gsr_backward_partials_peers(byref(b), None, 0, None, None) == -1
    assert b"peer" in lib.gsr_last_error()
    # with a (dummy) multicast address the ordinary argument checks apply next: required pointers are NULL
    assert lib.gsr_backward_partials_peers(byref(b), None, 0, ctypes.c_void_p(256), None) == -1
    assert b"NULL" in lib.gsr_last_error()
